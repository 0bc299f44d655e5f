"""Drop-in for ``knn_cuda.KNN`` (third-party KNN_CUDA 0.2; used at run_robot.py:65-66,122,138)."""
from __future__ import annotations

import torch

from . import ops


class KNN(torch.nn.Module):
    """``KNN(k, transpose_mode)(ref, query) -> (dist, idx)`` with EUCLIDEAN distances, ascending.

    transpose_mode=True:  ref [B,N,D], query [B,M,D] -> [B,M,k]
    transpose_mode=False: ref [B,D,N], query [B,D,M] -> [B,k,M]   (utils/flow_utils.py:127)
    D == 3, k <= 8.
    """

    def __init__(self, k, transpose_mode=False):
        super().__init__()
        self.k = k
        self._t = transpose_mode

    def forward(self, ref, query):
        if not self._t:
            ref, query = ref.transpose(1, 2), query.transpose(1, 2)
        if ref.shape[-1] != 3:
            raise ValueError("reart_b200 KNN is specialised for 3-D points")
        dist, idx = ops.knn(ref, query, self.k)
        if not self._t:
            dist, idx = dist.transpose(1, 2).contiguous(), idx.transpose(1, 2).contiguous()
        return dist, idx
