"""Import the UNMODIFIED reference (/root/reference) in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  /root/reference does not exist on the
GPU box, so nothing that runs there may import this module; it is used by
``oracle/make_golden.py`` (fixture generation) and by container-only tests that are skipped
when the reference tree is absent.

The reference needs three native/third-party modules that are not installed here
(SURVEY.md section 8c / Appendix B).  We register stand-ins in ``sys.modules`` *before*
importing any reference module:

* ``chamferdist._C``  -- torch restatement of PyTorch3D's brute-force K-NN as the reference
  calls it (utils/chamfer.py:174,206): squared L2 by direct differences, K=1..,
  lowest index wins ties; backward ``g1 = 2 g (p1 - p2[idx])``, ``g2[idx] -= ...``.
* ``knn_cuda.KNN``    -- ``torch.cdist`` + ``topk(largest=False)``, Euclidean distances,
  both ``transpose_mode``s (utils/flow_utils.py:158, utils/model_utils.py:42).
* viz / apted / trimesh -- ``MagicMock`` (never on the hot path).
"""
from __future__ import annotations

import os
import pickle
import sys
import types
from unittest.mock import MagicMock

import torch

REFERENCE_ROOT = os.environ.get("REART_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "utils"))


# ----------------------------------------------------------------------------- stand-ins
def _knn_points_idx(p1, p2, lengths1, lengths2, K, version):
    """Stand-in for chamferdist._C.knn_points_idx (call site utils/chamfer.py:174)."""
    B, P1, D = p1.shape
    P2 = p2.shape[1]
    idx = torch.zeros(B, P1, K, dtype=torch.int64, device=p1.device)
    dists = torch.zeros(B, P1, K, dtype=p1.dtype, device=p1.device)
    chunk = max(1, (1 << 24) // max(P2 * D, 1))
    for b in range(B):
        for s in range(0, P1, chunk):
            diff = p1[b, s:s + chunk, None, :] - p2[b, None, :, :]
            d2 = (diff * diff).sum(-1)                       # [c, P2]
            if K == 1:
                dmin, imin = d2.min(dim=1)
                # torch.min may return any minimal index on ties: force the lowest one
                first = (d2 == dmin[:, None]).to(torch.int8).argmax(dim=1)
                dists[b, s:s + chunk, 0] = dmin
                idx[b, s:s + chunk, 0] = first
            else:
                dk, ik = torch.topk(d2, K, dim=1, largest=False, sorted=True)
                dists[b, s:s + chunk] = dk
                idx[b, s:s + chunk] = ik
    return idx, dists


def _knn_points_backward(p1, p2, lengths1, lengths2, idx, grad_dists):
    """Stand-in for chamferdist._C.knn_points_backward (call site utils/chamfer.py:206)."""
    B, P1, D = p1.shape
    K = idx.shape[2]
    g1 = torch.zeros_like(p1)
    g2 = torch.zeros_like(p2)
    for k in range(K):
        j = idx[:, :, k]                                     # [B,P1]
        nb = torch.gather(p2, 1, j[:, :, None].expand(-1, -1, D))
        diff = 2.0 * grad_dists[:, :, k, None] * (p1 - nb)
        g1 += diff
        g2.scatter_add_(1, j[:, :, None].expand(-1, -1, D), -diff)
    return g1, g2


class _KNN:
    """Stand-in for knn_cuda.KNN (KNN_CUDA 0.2): Euclidean distances, ascending."""

    def __init__(self, k, transpose_mode=False):
        self.k = k
        self._t = transpose_mode

    def __call__(self, ref, query):
        if not self._t:
            ref, query = ref.transpose(1, 2), query.transpose(1, 2)
        d = torch.cdist(query, ref, compute_mode="donot_use_mm_for_euclid_dist")   # [B,M,N]
        dist, idx = torch.topk(d, self.k, dim=2, largest=False, sorted=True)
        if not self._t:
            dist, idx = dist.transpose(1, 2).contiguous(), idx.transpose(1, 2).contiguous()
        return dist, idx

    forward = __call__


def install_stubs() -> None:
    if "chamferdist" not in sys.modules:
        pkg = types.ModuleType("chamferdist")
        c = types.ModuleType("chamferdist._C")
        c.knn_points_idx = _knn_points_idx
        c.knn_points_backward = _knn_points_backward
        pkg._C = c
        sys.modules["chamferdist"] = pkg
        sys.modules["chamferdist._C"] = c
    if "knn_cuda" not in sys.modules:
        k = types.ModuleType("knn_cuda")
        k.KNN = _KNN
        sys.modules["knn_cuda"] = k
    if "pointnet2_cuda" not in sys.modules:
        # on a CUDA box the reference imports its compiled PointNet++ extension unconditionally
        # (networks/pointnet2_utils.py:7-12); the CPU timings made through this harness never call it
        sys.modules["pointnet2_cuda"] = MagicMock()
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm", "plotly",
                 "plotly.graph_objects", "plotly.express", "imageio", "apted", "apted.helpers", "trimesh", "kaleido"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock()
    os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")
    try:
        import networkx as nx
        if not hasattr(nx, "read_gpickle"):
            nx.read_gpickle = lambda p: pickle.load(open(p, "rb"))
    except Exception:
        pass


def import_reference():
    """Return a namespace of the reference modules on the hot path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    ns = types.SimpleNamespace()
    ns.chamfer = importlib.import_module("utils.chamfer")
    ns.model_utils = importlib.import_module("utils.model_utils")
    ns.screw_se3 = importlib.import_module("screw_se3")
    ns.kinematic_utils = importlib.import_module("utils.kinematic_utils")
    ns.flow_utils = importlib.import_module("utils.flow_utils")
    ns.loss = importlib.import_module("networks.loss")
    ns.model = importlib.import_module("networks.model")
    ns.KNN = _KNN
    return ns


def load_nao():
    """The shipped demo sequence (demo_data/data/nao/state_i.pkl) as float32 arrays."""
    import numpy as np
    root = os.path.join(REFERENCE_ROOT, "demo_data", "data", "nao")
    pcs, parts = [], []
    for i in range(10):
        with open(os.path.join(root, f"state_{i}.pkl"), "rb") as f:
            s = pickle.load(f)
        pcs.append(np.asarray(s["pc"], dtype=np.float64))
        parts.append(np.asarray(s["part_id"], dtype=np.int64))
    return np.stack(pcs), np.stack(parts)
