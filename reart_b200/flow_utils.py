"""Drop-in for ``blend_anchor_motion`` of the reference's ``utils/flow_utils.py`` (:147-170)."""
from __future__ import annotations

import torch

from . import ops


def blend_anchor_motion(query_loc, reference_loc, reference_flow, knn, return_mask=False):
    """Same signature as utils/flow_utils.py:147.  k = 3 runs the fused kernel (search + clamp + inverse
    distance weights + blend + mask in one launch); other k fall back to the generic k-NN kernel + torch tail."""
    k = knn.k if knn is not None else 3
    if k == 3:
        off = torch.tensor([0, reference_loc.shape[0]], dtype=torch.int64, device=query_loc.device)
        blended, mask = ops.knn3_blend(query_loc[None], reference_loc, reference_flow, off)
        return (blended[0], mask[0]) if return_mask else blended[0]
    dists, idx = ops.knn(reference_loc[None], query_loc[None], k)
    dists, idx = dists[0], idx[0]
    dists = dists.clamp_min(1e-10)
    weight = 1.0 / dists
    weight = weight / weight.sum(dim=-1, keepdim=True)
    blended = (reference_flow[idx] * weight[:, :, None]).sum(dim=1)
    if return_mask:
        min_d = dists.min(dim=-1)[0]
        flow_d = (reference_flow[idx] ** 2).sum(dim=-1).max(dim=1)[0]
        return blended, torch.logical_or(min_d <= flow_d, min_d <= 0.05)
    return blended


class FlowReference:
    """The per-pair reference lists of run_robot.py:78-84 concatenated once for the batched kernel."""

    def __init__(self, pc_ref_list, flow_ref_list):
        lens = [int(p.shape[0]) for p in pc_ref_list]
        dev = pc_ref_list[0].device
        self.ref_cat = torch.cat(list(pc_ref_list), dim=0).float().contiguous()
        self.flow_cat = torch.cat(list(flow_ref_list), dim=0).float().contiguous()
        off = [0]
        for n in lens:
            off.append(off[-1] + n)
        self.offsets = torch.tensor(off, dtype=torch.int64, device=dev)
        self.T = len(lens)
        # the sets never change during a fit: sort them along x once for the windowed exact kernel (worth it when the
        # sets are large; small ones stay on the brute-force kernel, which needs no per-call query ordering)
        self.max_refs = max(lens) if lens else 0
        self.sorted_refs = None
        if self.ref_cat.is_cuda and 2048 <= self.max_refs <= 16384:
            self.sorted_refs = ops.flow_refs_sort(self.ref_cat, self.offsets, self.max_refs)

    def slice(self, p0: int, p1: int) -> "FlowReference":
        """The pairs [p0, p1) as a view (shared storage; offsets stay relative to the concatenated arrays)."""
        sub = object.__new__(FlowReference)
        sub.ref_cat, sub.flow_cat = self.ref_cat, self.flow_cat
        sub.offsets = self.offsets[p0:p1 + 1].contiguous()
        sub.T = p1 - p0
        sub.max_refs = self.max_refs
        sub.sorted_refs = None if self.sorted_refs is None else (self.sorted_refs[0], self.sorted_refs[1][p0:p1].contiguous())
        return sub


def blend_anchor_motion_batched(query_list: torch.Tensor, ref: FlowReference):
    """All T frame pairs of run_robot.py:199-202 in one launch: query [T,m,3] -> (flow [T,m,3], mask [T,m])."""
    assert query_list.shape[0] == ref.T
    windowed = ref.sorted_refs is not None and query_list.shape[1] >= 2048
    return ops.knn3_blend(query_list, ref.ref_cat, ref.flow_cat, ref.offsets, sorted_refs=ref.sorted_refs if windowed else None)
