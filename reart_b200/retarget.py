"""Retargeting by inverse kinematics (``ik`` of the reference's utils/kinematic_utils.py:201-266).

The reference fits each novel state in its own 200-iteration Adam loop, one after the other, on a sparse cloud of one
point per ground-truth part (T = 1, N ~ 14), every iteration walking the dict-based ``fk`` (~10 k tiny ops).  The
sparse canonical points are the same for every state (utils/dataset_utils.py:75 fixes the sample index), the states
are independent and Adam/AMSGrad is element-wise, so here ALL S states are fitted at once: the per-state unknowns are
stacked along the model's frame axis (``theta_list [S,E]`` or ``proposal_6d/proposal_t [S,P,.]``), one forward poses
the sparse cloud S times (fused FK + skinning kernels), the summed MSE back-propagates to S independent rows, and a
single optimiser steps them -- the same trajectories as S separate loops, in 1/S of the launches.
"""
from __future__ import annotations

from typing import Dict

import torch


def _is_relaxation_model(model) -> bool:
    return hasattr(model, "proposal_6d") and hasattr(model, "proposal_t")


def init_unknowns(model, num_states: int, device) -> Dict[str, torch.Tensor]:
    """Start values of kinematic_utils.py:217-234: identity 6D / zero translation per part, or theta = 1e-6 per joint."""
    if _is_relaxation_model(model):
        six = torch.tensor([1.0, 0, 0, 0, 1, 0], device=device).repeat(num_states, model.num_parts, 1)
        return {"proposal_6d": six.requires_grad_(True),
                "proposal_t": torch.zeros(num_states, model.num_parts, 3, device=device, requires_grad=True)}
    E = model.axis_list.shape[0]
    return {"theta_list": torch.full((num_states, E), 1e-6, device=device, requires_grad=True)}


def retarget(model, sparse_cano_pc: torch.Tensor, sparse_novel_pc: torch.Tensor, n_iter: int = 200, lr: float = 1e-1,
             tau: float = 1.0, verbose: bool = False) -> Dict[str, torch.Tensor]:
    """Fit the per-state unknowns so that ``model(sparse_cano_pc)`` lands on ``sparse_novel_pc``.

    sparse_cano_pc (n,3); sparse_novel_pc (S,n,3), one row per novel state.  Returns the fitted keyword arguments
    (detached) to pass back into ``model(cloud, **kwargs)``.  Loss, optimiser and iteration count follow
    kinematic_utils.py:236-246: sum-reduced MSE, Adam(lr=0.1, amsgrad=True), 200 iterations.
    """
    if sparse_novel_pc.dim() == 2:
        sparse_novel_pc = sparse_novel_pc[None]
    S = sparse_novel_pc.shape[0]
    unknowns = init_unknowns(model, S, sparse_cano_pc.device)
    extra = {"tau": tau} if _is_relaxation_model(model) else {}
    optimizer = torch.optim.Adam(list(unknowns.values()), lr=lr, amsgrad=True)
    for it in range(n_iter):
        pc_trans = model(sparse_cano_pc, **extra, **unknowns)[0]
        loss = (pc_trans - sparse_novel_pc).square().sum()
        optimizer.zero_grad(set_to_none=True)
        loss.backward()
        optimizer.step()
        if verbose and (it % 50 == 0 or it == n_iter - 1):
            print(f"retarget iter {it}: loss {loss.item():.6f}")
    return {**extra, **{k: v.detach() for k, v in unknowns.items()}}


@torch.no_grad()
def retarget_error(model, cano_pc: torch.Tensor, novel_pc: torch.Tensor, fitted: Dict[str, torch.Tensor]):
    """Per-state error of kinematic_utils.py:248-255: 100 x mean Euclidean distance between the posed dense cloud and
    the ground-truth novel cloud.  cano_pc (N,3), novel_pc (S,N,3) -> (errors (S,), posed (S,N,3), seg_part (N,))."""
    pc_trans, seg_part, _ = model(cano_pc, **fitted)
    err = 100.0 * (pc_trans - novel_pc).square().sum(dim=-1).sqrt().mean(dim=1)
    return err, pc_trans, seg_part


def ik(dataset, model, device, verbose=True, vis=True, save_dir=None, sampler=None, **ikargs):
    """Same call as the reference's ``ik`` (robot sequences only): mean retarget error over ``dataset.novel_pose_list``.

    ``sampler`` is the reference's ``utils.dataset_utils.sparse_sample_novel_state`` (dataset formats are outside this
    package; it is imported from the reference tree when not given).  Visualisation files are the caller's business:
    ``vis``/``save_dir`` are accepted and ignored.
    """
    import numpy as np
    if sampler is None:
        from utils.dataset_utils import sparse_sample_novel_state as sampler           # the reference tree's loader
    sample = dataset[0]
    cano_pose = dataset.pose_list[dataset.cano_idx]
    states = [sampler(sample["cano_pc"], sample["gt_cano_part"], cano_pose, novel_pose, sparse_sample_per_part=1)
              for novel_pose in dataset.novel_pose_list]
    as_dev = lambda a: torch.from_numpy(np.asarray(a)).float().to(device)
    sparse_cano = as_dev(states[0]["sparse_cano_pc"])
    sparse_novel = torch.stack([as_dev(s["sparse_novel_pc"]) for s in states])
    novel = torch.stack([as_dev(s["novel_pc"]) for s in states])
    fitted = retarget(model, sparse_cano, sparse_novel, n_iter=ikargs.get("n_iter", 200), tau=ikargs.get("tau", 1.0),
                      verbose=verbose)
    err, _, _ = retarget_error(model, as_dev(sample["cano_pc"]), novel, fitted)
    if verbose:
        for s, e in enumerate(err.tolist()):
            print(f"Novel retarget err {s}: {e:.3f}")
    return float(err.mean())
