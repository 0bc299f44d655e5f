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
