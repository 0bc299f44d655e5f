"""Snapshot-evaluation helpers on the GPU.

Same names and argument meaning as the reference's ``utils/eval_utils.py``:
``compute_chamfer`` / ``compute_chamfer_list`` (:39-66, two scipy KD-trees per frame on the CPU there) run on the
symmetric search kernel for all frames at once, and ``eval_seg`` (:25-36, two N x N matmuls there) uses the
contingency table, O(N + s^2).  Called every ``snapshot_gap`` iterations by the run scripts (run_robot.py:247-268).
"""
from __future__ import annotations

import numpy as np
import torch

from .chamfer import _ChamferBidir


def _as_cuda(a, device=None) -> torch.Tensor:
    if not torch.is_tensor(a):
        a = torch.from_numpy(np.ascontiguousarray(a))
    return a.to(device or "cuda", dtype=torch.float32)


def compute_chamfer_list(points_set1, points_set2, reduction="sum"):
    """utils/eval_utils.py:54-66: per-frame bidirectional Chamfer (sum or mean of squared NN distances per
    direction), then sum / mean / none over frames.  [T,N,3] x [T,M,3] (numpy or torch) -> python float / [T]."""
    a, b = _as_cuda(points_set1), _as_cuda(points_set2)
    d_f, d_b, _, _ = _ChamferBidir.apply(a, b)
    if reduction == "mean":
        per = d_f.double().mean(dim=1) + d_b.double().mean(dim=1)
        return float(per.mean())
    per = d_f.double().sum(dim=1) + d_b.double().sum(dim=1)
    if reduction == "sum":
        return float(per.sum())
    return per.cpu().numpy()


def compute_chamfer(points_1, points_2, reduction="sum"):
    """utils/eval_utils.py:39-51 for one pair of clouds."""
    out = compute_chamfer_list(_as_cuda(points_1)[None], _as_cuda(points_2)[None],
                               reduction="mean" if reduction == "mean" else "none")
    return float(out) if reduction == "mean" else float(out[0])


def eval_seg(gt_segm: torch.Tensor, pd_segm: torch.Tensor):
    """Rand index of utils/eval_utils.py:25-36 without the two N x N matrices: with contingency counts n_ij,
    row sums a_i and column sums b_j, the number of ordered pairs (incl. i == j) on which both labelings agree is
    N^2 - sum a_i^2 - sum b_j^2 + 2 sum n_ij^2."""
    n = gt_segm.shape[0]
    s = int(max(int(gt_segm.max()), int(pd_segm.max())) + 1)
    table = torch.bincount(gt_segm.long() * s + pd_segm.long(), minlength=s * s).reshape(s, s).double()
    a, b = table.sum(1), table.sum(0)
    agree = n * n - (a * a).sum() - (b * b).sum() + 2.0 * (table * table).sum()
    return np.float32((agree / float(n * n)).item())


def eval_flow(pred_flow_list, gt_flow_list, acc1_thre=0.05, acc2_thre=0.1):
    """utils/eval_utils.py:6-22 (EPE, Acc_5, Acc_10, mean angle) on torch tensors or numpy arrays."""
    p, g = _as_cuda(pred_flow_list).double(), _as_cuda(gt_flow_list).double()
    error = torch.sqrt(((p - g) ** 2).sum(2) + 1e-20)
    glen = torch.sqrt((g * g).sum(2) + 1e-20)
    acc1 = torch.logical_or(error <= acc1_thre, error / glen <= acc1_thre).double().mean(1).mean()
    acc2 = torch.logical_or(error <= acc2_thre, error / glen <= acc2_thre).double().mean(1).mean()
    epe = error.mean()
    ul = g / g.norm(dim=-1, keepdim=True)
    up = p / p.norm(dim=-1, keepdim=True)
    eps = 1e-7
    dot = (ul * up).sum(2).clamp(-1 + eps, 1 - eps)
    dot = torch.where(torch.isnan(dot), torch.ones_like(dot), dot)
    ang = torch.arccos(dot).mean(1).mean()
    return float(epe), float(acc1), float(acc2), float(ang)
