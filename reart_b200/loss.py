"""Drop-in for the live losses of the reference's ``networks/loss.py`` (recon_loss :24-29, flow_loss :10-21)."""
from __future__ import annotations

import torch


def flow_loss(gt_flow_list, pred_flow_list, flow_mask_list=None, robust=False, smooth_weight=1e-2):
    """networks/loss.py:10-21 (argument order gt, pred -- SURVEY Q19).

    Per point: the data term sum_xyz MSE (or Huber, delta 1) where the mask is set, and
    ``smooth_weight * |pred|^2`` where it is not; everything summed.  [T,N,3] x2, mask [T,N] -> scalar.
    """
    err = pred_flow_list - gt_flow_list
    if robust:
        a = err.abs()
        per = torch.where(a < 1.0, 0.5 * err * err, a - 0.5).sum(dim=2)
    else:
        per = (err * err).sum(dim=2)
    if flow_mask_list is None:
        return per.sum()
    keep = flow_mask_list.to(per.dtype)
    drop = torch.logical_not(flow_mask_list).to(per.dtype)
    mag = (pred_flow_list * pred_flow_list).sum(dim=2)
    return (keep * per + smooth_weight * (drop * mag)).sum()


def recon_loss(pc_trans_list, pc_list, chamfer_dist):
    """networks/loss.py:24-29 -- sum of the bidirectional per-point Chamfer distances."""
    cd = chamfer_dist(pc_trans_list, pc_list, bidirectional=True)
    return torch.sum(cd)
