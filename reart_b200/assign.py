"""Assignment loss of the run scripts (run_robot.py:164-187, run_real.py:180-203, run_sapien.py:179-203).

Every ``assign_gap`` iterations the reference FPS-samples the canonical cloud and every observed frame, builds
the [T,n,n] Euclidean cost, solves one Hungarian assignment per frame on the CPU inside a freshly spawned
``multiprocessing.Pool`` (utils/model_utils.py:85-89) and then penalises the squared distance of the
matched pairs.  Here: FPS for all frames in ONE launch (``reart_fps``), the cost matrix stays a torch op, and
the Hungarian solves run in a persistent worker pool created once (a GPU LAP is future work).
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Optional

import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

from . import ops


def index_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """networks/pointnet2_utils.py:60-76: points [B,N,C], idx [B,S] -> [B,S,C]."""
    return torch.gather(points, 1, idx[:, :, None].expand(-1, -1, points.shape[2]))


def farthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """networks/pointnet2_utils.py:74-99 on CUDA: [B,N,3] -> [B,npoint] int64, starting at index 0."""
    return ops.fps(xyz, npoint)


class AssignLoss:
    """lambda_assign * sum ||src_matched - tgt_matched||^2 with assignments refreshed every ``assign_gap`` calls."""

    def __init__(self, cano_pc: torch.Tensor, pc_list: torch.Tensor, downsample: int = 4, assign_gap: int = 5,
                 lambda_assign: float = 3e-1, workers: Optional[int] = None):
        self.cano_pc, self.pc_list = cano_pc, pc_list
        self.T, N = pc_list.shape[0], pc_list.shape[1]
        self.num_fps = N // downsample
        self.assign_gap, self.lambda_assign = assign_gap, lambda_assign
        self.calls = 0
        # the sample indices never change (FPS is deterministic from index 0): compute them once
        self.src_idx = farthest_point_sample(cano_pc[None], self.num_fps).expand(self.T, self.num_fps)
        self.tgt_idx = farthest_point_sample(pc_list, self.num_fps)
        self.pc_tgt = index_points(pc_list, self.tgt_idx)
        self.pool = ThreadPoolExecutor(max_workers=workers or min(self.T, 16))
        self.match_src = self.match_tgt = None

    def refresh(self, pc_src: torch.Tensor) -> None:
        with torch.no_grad():
            cost = torch.cdist(pc_src, self.pc_tgt).cpu().numpy()
        res = list(self.pool.map(linear_sum_assignment, cost))
        dev = pc_src.device
        self.match_src = torch.from_numpy(np.stack([r[0] for r in res])).to(dev)
        self.match_tgt = torch.from_numpy(np.stack([r[1] for r in res])).to(dev)

    def __call__(self, pc_trans_list: torch.Tensor) -> torch.Tensor:
        pc_src = index_points(pc_trans_list, self.src_idx)
        if self.match_src is None or self.calls % self.assign_gap == 0:
            self.refresh(pc_src.detach())
        self.calls += 1
        a = index_points(pc_src, self.match_src)
        b = index_points(self.pc_tgt, self.match_tgt)
        return self.lambda_assign * ((a - b) ** 2).sum(dim=-1).sum()
