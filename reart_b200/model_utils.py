"""Drop-in for the hot-path helpers of the reference's ``utils/model_utils.py``."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import ops


def th_with_zeros(tensor: torch.Tensor) -> torch.Tensor:
    """utils/model_utils.py:12-19 -- append the [0,0,0,1] row: (B,3,4) -> (B,4,4)."""
    pad = tensor.new_tensor([0.0, 0.0, 0.0, 1.0]).view(1, 1, 4).expand(tensor.shape[0], 1, 4)
    return torch.cat([tensor, pad], dim=1)


def create_transformation(rotation: torch.Tensor, translation: torch.Tensor) -> torch.Tensor:
    """utils/model_utils.py:22-30 -- rotation (B,3,3), translation (B,3,1) -> (B,4,4)."""
    B = rotation.shape[0]
    last_row = torch.zeros(B, 1, 4, device=rotation.device, dtype=rotation.dtype)
    last_row[:, :, 3] = 1
    return torch.cat([torch.cat([rotation, translation], dim=2), last_row], dim=1)


def tau_cosine(cur_iter, max_iter, end_temp, start_temp):
    """utils/model_utils.py:33-37."""
    assert end_temp <= start_temp
    return end_temp + (start_temp - end_temp) * (math.cos(math.pi * cur_iter / max_iter) + 1.0) * 0.5


def knn_query(query_pc, src_pc, src_input, knn):
    """utils/model_utils.py:41-51 -- transfer labels ([n] int) or features ([n,C]) from src_pc to query_pc through
    the k nearest neighbours: features are averaged, labels take the mode (for k == 1 simply the neighbour's)."""
    k = knn.k
    nn_idx = knn(ref=src_pc[None], query=query_pc[None])[1][0].reshape(-1)        # [m*k]
    picked = src_input[nn_idx]
    if src_input.dim() == 2:
        return picked.reshape(src_input.shape[0], k, src_input.shape[1]).mean(dim=1)
    picked = picked.reshape(-1, k)
    return picked[:, 0] if k == 1 else torch.mode(picked, dim=1)[0]


def compute_pc_transform(cano_pc, pose_list, cano_part):
    """utils/model_utils.py:54-67 -- hard-label skinning: cano (N,3), pose (T,P,4,4), part (N,) -> (T,N,3)."""
    num_parts = pose_list.shape[1]
    R = pose_list[:, :, :3, :3]
    tr = pose_list[:, :, :3, 3]
    W = F.one_hot(cano_part, num_classes=num_parts)
    return ops.skin(cano_pc, W, R, tr)


def compute_align_trans(trans_list: torch.Tensor, root_trans: torch.Tensor) -> torch.Tensor:
    """utils/model_utils.py:121-126 -- express every part motion in the root's frame: (T,P,4,4), (T,4,4) -> (T,P,4,4)."""
    from .screw_se3 import inverse_transformation
    return torch.matmul(inverse_transformation(root_trans)[:, None], trans_list)


def compute_ass_err(pc_trans_list: torch.Tensor, pc_list: torch.Tensor, use_nproc: bool = True) -> torch.Tensor:
    """utils/model_utils.py:92-103 -- model-selection term (run_robot.py:306): mean squared distance of the optimal
    one-to-one matching between each posed cloud and its observed frame, (T,N,3) x2 -> scalar.

    On CUDA with N <= 4096 the T assignments are solved on the GPU (``reart_lap``: exact, costs on the fly, no N x N
    matrix and no D2H -- the reference copies T float64 N x N matrices to the host and runs scipy in a process pool);
    otherwise the reference's host path (cdist -> scipy) in a thread pool.  ``use_nproc`` only matters on the host path.
    """
    T, N = pc_list.shape[0], pc_list.shape[1]
    if pc_list.is_cuda and N <= 4096:
        from .assign import lap_assign
        cols = lap_assign(pc_trans_list, pc_list).long()
        b = torch.gather(pc_list, 1, cols[:, :, None].expand(-1, -1, 3))
        return (pc_trans_list - b).square().sum(dim=-1).mean()
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from scipy.optimize import linear_sum_assignment
    cost = torch.cdist(pc_trans_list, pc_list).cpu().numpy()
    if use_nproc and len(cost) > 1:
        with ThreadPoolExecutor(max_workers=min(len(cost), 16)) as pool:
            pairs = list(pool.map(linear_sum_assignment, cost))
    else:
        pairs = [linear_sum_assignment(c) for c in cost]
    rows = torch.from_numpy(np.stack([r for r, _ in pairs])).to(pc_list.device)
    cols = torch.from_numpy(np.stack([c for _, c in pairs])).to(pc_list.device)
    a = torch.gather(pc_trans_list, 1, rows[:, :, None].expand(-1, -1, 3))
    b = torch.gather(pc_list, 1, cols[:, :, None].expand(-1, -1, 3))
    return (a - b).square().sum(dim=-1).mean()


def compute_group_temporal_err(pc_list: torch.Tensor, seg_part: torch.Tensor) -> torch.Tensor:
    """utils/model_utils.py:106-118 -- model-selection term (run_robot.py:311): the largest, over parts, mean squared
    distance of a part's points to the part's per-frame centroid: (T,N,3), (N,) -> scalar.  One segmented reduction
    over all parts (the reference loops over parts with an ``.item()`` each)."""
    labels, inverse = torch.unique(seg_part, sorted=True, return_inverse=True)
    P, T = labels.numel(), pc_list.shape[0]
    counts = torch.bincount(inverse, minlength=P).to(pc_list.dtype)
    sums = torch.zeros(T, P, 3, dtype=pc_list.dtype, device=pc_list.device).index_add_(1, inverse, pc_list)
    centroid = sums / counts[None, :, None]
    sq = (pc_list - centroid[:, inverse]).square().sum(dim=2)                                  # (T,N)
    per_part = torch.zeros(T, P, dtype=pc_list.dtype, device=pc_list.device).index_add_(1, inverse, sq).sum(dim=0)
    return (per_part / (counts * T)).max()
