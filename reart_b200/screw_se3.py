"""Drop-in for the per-iteration part of the reference's ``screw_se3`` package.

Same names and argument meaning as screw_se3/geo_utils.py and screw_se3/screw_utils.py for the functions
on the optimisation path.  CUDA tensors run the fused kernels of
``csrc/se3.cu``; shapes/dtypes follow the reference.  The init-only helpers (dual quaternions,
log maps, quaternion conversions) are outside the hot path and are not re-implemented here.
"""
from __future__ import annotations

import torch

from . import ops


def rotation_6d_to_matrix(d6: torch.Tensor) -> torch.Tensor:
    """screw_se3/geo_utils.py:632-651 -- Gram-Schmidt, rows (b1, b2, b1 x b2); (*,6) -> (*,3,3)."""
    return ops.rot6d(d6)


def matrix_to_rotation_6d(matrix: torch.Tensor) -> torch.Tensor:
    """screw_se3/geo_utils.py:654-668 -- drop the last row; (*,3,3) -> (*,6)."""
    return matrix[..., :2, :].clone().reshape(matrix.shape[:-2] + (6,))


def screw_to_transform(l: torch.Tensor, m: torch.Tensor, theta: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    """Fused ``transform_from_exponential_coordinates(screw_param_to_exponential_coordinates(l, m, theta, d))``
    (screw_se3/screw_utils.py:6-30): l, m (B,3), theta, d (B,) -> (B,4,4)."""
    return ops.screw_to_transform(l, m, theta, d)


def screw_param_to_exponential_coordinates(l, m, theta, d):
    """screw_se3/screw_utils.py:6-23 -- (B,6) exponential coordinates [w*theta, v*theta] (rotation first).

    Branch-free restatement (no boolean-mask indexing, hence no host sync): the no-rot set is
    |theta| < 1e-6 or |theta - pi| < 1e-6; there w = 0, v = l and d is ignored.
    """
    import math
    eps = 1e-6
    no_rot = torch.logical_or(theta.abs() < eps, (theta - math.pi).abs() < eps)[:, None]
    q = torch.cross(l, m, dim=-1)
    safe_theta = torch.where(no_rot[:, 0], torch.ones_like(theta), theta)
    h = (d / safe_theta)[:, None]
    v_rot = torch.cross(q, l, dim=-1) + h * l
    w = torch.where(no_rot, torch.zeros_like(l), l)
    v = torch.where(no_rot, l, v_rot)
    return torch.cat((w, v), dim=1) * theta[:, None]


def transform_from_exponential_coordinates(log_transform: torch.Tensor) -> torch.Tensor:
    """screw_se3/screw_utils.py:27-30 over se3_exp_map (geo_utils.py:147-222): (B,6) [rot, trans] -> (B,4,4).

    theta^2 is clamped at 1e-4 before the square root; K is not normalised.
    Thin torch composition (this entry is not on the per-iteration path; ``fk`` uses the fused kernel).
    """
    w, v = log_transform[:, :3], log_transform[:, 3:]
    nrm = (w * w).sum(1)
    ang = torch.clamp(nrm, 1e-4).sqrt()
    inv = 1.0 / ang
    fac1 = inv * ang.sin()
    fac2 = inv * inv * (1.0 - ang.cos())
    zero = torch.zeros_like(ang)
    K = torch.stack([zero, -w[:, 2], w[:, 1], w[:, 2], zero, -w[:, 0], -w[:, 1], w[:, 0], zero], dim=1).reshape(-1, 3, 3)
    K2 = torch.bmm(K, K)
    eye = torch.eye(3, dtype=w.dtype, device=w.device)[None]
    Rm = fac1[:, None, None] * K + fac2[:, None, None] * K2 + eye
    V = eye + K * ((1 - ang.cos()) / ang ** 2)[:, None, None] + K2 * ((ang - ang.sin()) / ang ** 3)[:, None, None]
    t = torch.bmm(V, v[:, :, None])
    top = torch.cat([Rm, t], dim=2)
    bottom = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=w.dtype, device=w.device).expand(w.shape[0], 1, 4)
    return torch.cat([top, bottom], dim=1)


def inverse_transformation(trans_12: torch.Tensor) -> torch.Tensor:
    """screw_se3/geo_utils.py:9-53 -- [R t; 0 1]^-1 = [R^T  -R^T t; 0 1] for (N,4,4) or (4,4)."""
    if not torch.is_tensor(trans_12):
        raise TypeError("Input type is not a torch.Tensor. Got {}".format(type(trans_12)))
    Rt = trans_12[..., :3, :3].transpose(-1, -2)
    t = -torch.matmul(Rt, trans_12[..., :3, 3:4])
    out = torch.zeros_like(trans_12)
    out[..., :3, :3] = Rt
    out[..., :3, 3:4] = t
    out[..., 3, 3] = 1.0
    return out
