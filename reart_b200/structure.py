"""Structure extraction: relaxation result -> part merging -> spanning tree -> joint screws.

Mirrors the public functions of the reference's ``utils/graph_utils.py`` (:62-421), the tree-building half of
``utils/kinematic_utils.py`` (:20-148) and the dual-quaternion screw extraction of ``screw_se3/dq_utils.py``
(:134-182): same names, argument meaning and return values, so ``run_robot.py:101-124,224-243`` can call them
unchanged.  The stage runs once per fit (kinematic init and the last iteration), on [T,P,P] relative transforms.

What is different from the reference is the shape of the computation, not the arithmetic:

* every function is batched over (T, E) with ``torch.where`` selections -- no boolean-mask indexing, no per-edge or
  per-part Python loops with ``.item()``, hence no host synchronisation inside the tensor code;
* the revolute/prismatic joint fit that the reference spells out three times (``compute_geo_cost`` :127-160,
  ``compute_screw_trans`` :222-266, ``build_graph`` kinematic_utils.py:88-123) is one routine, ``_joint_fit``;
* per-part furthest point sampling is ONE padded batch (``reart_fps`` launch) instead of one call per part;
* the greedy spanning tree and the merge pass run on the host from a single D2H copy of the PxP cost matrix.

Tensor code here is device-agnostic torch; the Chamfer search and FPS arrive as callables (``chamfer_dist`` exactly
as in the reference; ``fps`` defaults to the CUDA kernel and raises on CPU tensors like every kernel of this package).
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import networkx as nx
import numpy as np
import torch

from .screw_se3 import (inverse_transformation, screw_param_to_exponential_coordinates,
                        transform_from_exponential_coordinates)

_EPS = 1e-6


# ----------------------------------------------------------------------------------------------- dual quaternions
def _qmul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Hamilton product a*b, real part first (what dq_utils.py:66-84 computes through an outer product)."""
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz,
                        aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw), dim=-1)


def matrix_to_quaternion(matrix: torch.Tensor) -> torch.Tensor:
    """screw_se3/geo_utils.py:536-587 -- (...,3,3) -> (...,4), real first; the best-conditioned of the four
    candidates (largest |q_i|) is picked with a gather instead of a boolean mask."""
    if matrix.size(-1) != 3 or matrix.size(-2) != 3:
        raise ValueError(f"Invalid rotation matrix shape {matrix.shape}.")
    m = matrix.reshape(matrix.shape[:-2] + (9,))
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = m.unbind(-1)
    q_abs = torch.stack((1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
                         1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22), dim=-1).clamp_min(0.0).sqrt()
    a, b, c, d = (q_abs ** 2).unbind(-1)
    cand = torch.stack((torch.stack((a, m21 - m12, m02 - m20, m10 - m01), -1),
                        torch.stack((m21 - m12, b, m10 + m01, m02 + m20), -1),
                        torch.stack((m02 - m20, m10 + m01, c, m12 + m21), -1),
                        torch.stack((m10 - m01, m20 + m02, m21 + m12, d), -1)), dim=-2)
    cand = cand / (2.0 * q_abs.clamp_min(0.1))[..., None]
    best = q_abs.argmax(dim=-1)
    return torch.gather(cand, -2, best[..., None, None].expand(best.shape + (1, 4))).squeeze(-2)


def transform_to_dq(T: torch.Tensor) -> torch.Tensor:
    """screw_se3/dq_utils.py:134-139 -- (B,4,4) -> (B,8) unit dual quaternion [q_r, q_d], q_d = (0,t) q_r / 2."""
    q_r = matrix_to_quaternion(T[:, :3, :3])
    t = torch.cat((torch.zeros_like(T[:, :1, 3]), T[:, :3, 3]), dim=1)
    return torch.cat((q_r, 0.5 * _qmul(t, q_r)), dim=1)


def dq_to_screw(dq: torch.Tensor, eps: float = _EPS):
    """screw_se3/dq_utils.py:142-182 -- (*,8) -> Pluecker axis l, moment m (*,3), rotation theta and translation d
    (*) about/along it.  Conventions kept: the axis is flipped into the half space l.(1,1,1) >= 0 (theta, d follow);
    |theta| < eps or |theta - pi| < eps counts as "no rotation" (axis = translation direction, theta := eps); the
    identity gets axis component 0 set to 1.  (The reference also emits a warning for identities; that needs a host
    sync and is dropped.)"""
    assert dq.shape[-1] == 8
    q_r, q_d = dq[..., :4], dq[..., 4:]
    qn = q_r / q_r.square().sum(-1, keepdim=True).sqrt()
    theta = 2.0 * torch.atan2(torch.linalg.norm(qn[..., 1:], dim=-1), qn[..., 0])
    no_rot = torch.logical_or(theta.abs() < eps, (theta - math.pi).abs() < eps)
    conj = q_r * q_r.new_tensor([1.0, -1.0, -1.0, -1.0])
    t = _qmul(2.0 * q_d, conj)[..., 1:]

    t_norm = torch.linalg.norm(t, dim=-1)
    l = torch.where(no_rot[..., None], t / (t_norm[..., None] + 1e-10), q_r[..., 1:] / torch.sin(theta / 2)[..., None])
    flip = l.sum(dim=-1) < 0
    theta = torch.where(flip, -theta, theta)
    l = torch.where(flip[..., None], -l, l)
    d = torch.where(no_rot, torch.where(flip, -t_norm, t_norm), (t * l).sum(dim=-1))

    unit = torch.logical_and(no_rot, torch.isclose(d, torch.zeros_like(d)))
    l = torch.cat((torch.where(unit, torch.ones_like(d), l[..., 0])[..., None], l[..., 1:]), dim=-1)
    theta = torch.where(no_rot, torch.full_like(theta, eps), theta)
    t_x_l = torch.cross(t, l, dim=-1)
    m = 0.5 * (t_x_l + torch.cross(l, t_x_l / torch.tan(theta / 2)[..., None], dim=-1))
    return l, m, theta, d


# ------------------------------------------------------------------------------------------------ geometric costs
def frobenius_cost(predict: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """graph_utils.py:180-186 -- || predict gt^-1 - I ||_F^2 per transform: (B,4,4) x2 -> (B,)."""
    err = torch.bmm(predict, inverse_transformation(gt))
    eye = torch.eye(4, dtype=predict.dtype, device=predict.device)
    return (err - eye).square().sum(dim=(-2, -1))


def compute_root_cost(trans_list: torch.Tensor) -> torch.Tensor:
    """graph_utils.py:189-193 -- how far each part's motion is from static: (T,P,4,4) -> (P,)."""
    eye = torch.eye(4, dtype=trans_list.dtype, device=trans_list.device)
    return (trans_list - eye).square().sum(dim=(2, 3)).mean(dim=0)


def compute_mean_screw_param(s_axis, moment, theta, distance, eps_tol: float = 1e-5):
    """graph_utils.py:196-219 -- time-average of the axis and moment of each edge, (T,E,3) -> (E,3), skipping the
    frames whose transform is the identity (their axis is arbitrary) unless every frame is.  One masked mean over
    all edges; a single edge (E <= 1) is the plain mean, as in the reference."""
    assert s_axis.dim() == 3 and moment.dim() == 3
    if s_axis.shape[1] <= 1:
        return s_axis.mean(dim=0), moment.mean(dim=0)
    no_rot = torch.logical_or(theta.abs() <= eps_tol, (theta - math.pi).abs() <= eps_tol)
    unit = torch.logical_and(no_rot, distance <= eps_tol)
    keep = torch.logical_or(~unit, unit.all(dim=0, keepdim=True))[..., None]                       # (T,E,1)
    count = keep.sum(dim=0).to(s_axis.dtype)
    zero = torch.zeros_like(s_axis)
    return torch.where(keep, s_axis, zero).sum(dim=0) / count, torch.where(keep, moment, zero).sum(dim=0) / count


def _joint_fit(rel_trans, mean_axis, mean_moment, theta, distance):
    """One-dof fit of a relative motion by a revolute and by a prismatic joint about a fixed screw axis.

    rel_trans (T,E,4,4); mean_axis, mean_moment (E,3); theta, distance (T,E).  Returns
    ``recon_r, recon_p`` (T,E,4,4), ``cost_r`` (E,) = sum_t frob(recon_r, rel), ``cost_p1`` (E,) = sum_t
    frob(recon_p, rel with R := I) and ``rot_res`` (T,E) = mean squared difference of the 3x3 blocks of recon_p and
    rel (the callers average it over different index sets: graph_utils.py:157, :262, kinematic_utils.py:111).
    The unused degree of freedom is pinned at 1e-6, not 0 (graph_utils.py:135,147).
    """
    T, E = theta.shape
    ax = mean_axis[None].expand(T, E, 3).reshape(-1, 3)
    mo = mean_moment[None].expand(T, E, 3).reshape(-1, 3)
    pinned = torch.full_like(theta.reshape(-1), _EPS)
    rel = rel_trans.reshape(-1, 4, 4)

    recon_r = transform_from_exponential_coordinates(
        screw_param_to_exponential_coordinates(ax, mo, theta.reshape(-1), pinned))
    cost_r = frobenius_cost(recon_r, rel).reshape(T, E).sum(dim=0)

    recon_p = transform_from_exponential_coordinates(
        screw_param_to_exponential_coordinates(ax, mo, pinned, distance.reshape(-1)))
    rel_no_rot = rel.clone()
    rel_no_rot[:, :3, :3] = torch.eye(3, dtype=rel.dtype, device=rel.device)
    cost_p1 = frobenius_cost(recon_p, rel_no_rot).reshape(T, E).sum(dim=0)
    rot_res = (recon_p[:, :3, :3] - rel[:, :3, :3]).square().mean(dim=(1, 2)).reshape(T, E)
    return recon_r.reshape(T, E, 4, 4), recon_p.reshape(T, E, 4, 4), cost_r, cost_p1, rot_res


def compute_relative_trans(trans_list: torch.Tensor, return_trans: bool = False):
    """graph_utils.py:163-177 -- screw parameters of every ordered part pair: (T,P,4,4) ->
    axis, moment (T,P,P,3), theta, distance (T,P,P) [, rel_trans (T,P,P,4,4)] with rel[t,i,j] = T_i^-1 T_j."""
    T, P = trans_list.shape[:2]
    inv = inverse_transformation(trans_list.reshape(-1, 4, 4)).reshape(T, P, 4, 4)
    rel = torch.matmul(inv[:, :, None], trans_list[:, None, :]).reshape(-1, 4, 4)
    l, m, theta, d = dq_to_screw(transform_to_dq(rel))
    out = (l.reshape(T, P, P, 3), m.reshape(T, P, P, 3), theta.reshape(T, P, P), d.reshape(T, P, P))
    return out + (rel.reshape(T, P, P, 4, 4),) if return_trans else out


def compute_geo_cost(rel_trans, axis, moment, theta, distance) -> torch.Tensor:
    """graph_utils.py:127-160 -- (P,P) cost of explaining each pair's relative motion by ONE 1-dof joint:
    min(revolute, prismatic) residual summed over time."""
    T, P = axis.shape[:2]
    mean_axis, mean_moment = compute_mean_screw_param(axis.reshape(T, -1, 3), moment.reshape(T, -1, 3),
                                                      theta.reshape(T, -1), distance.reshape(T, -1))
    _, _, cost_r, cost_p1, rot_res = _joint_fit(rel_trans.reshape(T, P * P, 4, 4), mean_axis, mean_moment,
                                                theta.reshape(T, -1), distance.reshape(T, -1))
    return torch.min(cost_r, cost_p1 + rot_res.mean()).reshape(P, P)


def compute_screw_trans(trans_list: torch.Tensor, return_cost: bool = False):
    """graph_utils.py:222-266 -- project each edge's motion (T,E,4,4) onto its best 1-dof joint; optionally also
    the mean residual / T."""
    T, E = trans_list.shape[:2]
    l, m, theta, d = dq_to_screw(transform_to_dq(trans_list.reshape(-1, 4, 4)))
    l, m, theta, d = l.reshape(T, E, 3), m.reshape(T, E, 3), theta.reshape(T, E), d.reshape(T, E)
    mean_axis, mean_moment = compute_mean_screw_param(l, m, theta, d)
    recon_r, recon_p, cost_r, cost_p1, rot_res = _joint_fit(trans_list, mean_axis, mean_moment, theta, d)
    cost_p = cost_p1 + rot_res.mean()
    T_recon = torch.where((cost_p <= cost_r)[None, :, None, None], recon_p, recon_r)
    if return_cost:
        return T_recon, torch.min(cost_r, cost_p).mean() / T
    return T_recon


def compute_screw_cost(pred_trans_list: torch.Tensor, pred_connection: torch.Tensor) -> torch.Tensor:
    """graph_utils.py:269-275 -- 1-dof residual of the chosen edges (part of the model-selection energy,
    run_robot.py:309)."""
    T = pred_trans_list.shape[0]
    src, tgt = pred_trans_list[:, pred_connection[:, 0]], pred_trans_list[:, pred_connection[:, 1]]
    inv_src = inverse_transformation(src.reshape(-1, 4, 4)).reshape(T, -1, 4, 4)
    return compute_screw_trans(torch.matmul(inv_src, tgt), return_cost=True)[1]


# ------------------------------------------------------------------------------------- sampling / spatial costs
def _default_fps(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    from .assign import farthest_point_sample
    return farthest_point_sample(xyz, npoint)


def fps_sample_cano(cano_pc, cano_part, uni_label, num_fps: int = 20, fps: Optional[Callable] = None):
    """graph_utils.py:38-56 -- ``num_fps`` furthest-point samples of every part of the canonical cloud:
    -> points (P,num_fps,3), indices into cano_pc (P,num_fps).

    All parts go through ONE batched FPS call: the cloud is ordered by part, every part is padded to the largest
    one by repeating its own first point (distance 0 to the start sample, so padding is never selected), and the
    sampled positions are mapped back.  Same samples as the per-part calls of the reference on CUDA (start index 0).
    """
    fps = fps or _default_fps
    uni_label = torch.as_tensor(uni_label, device=cano_part.device)
    P = uni_label.numel()
    slot = torch.searchsorted(uni_label, cano_part)
    slot = torch.where(uni_label[slot.clamp_max(P - 1)] == cano_part, slot, torch.full_like(slot, P))   # others last
    order = torch.argsort(slot, stable=True)
    counts = torch.bincount(slot, minlength=P + 1)[:P]
    counts_h = counts.cpu()
    if int(counts_h.min()) < num_fps:
        bad = int(counts_h.argmin())
        raise ValueError("part id {} too small, only {} points".format(uni_label[bad].item(), int(counts_h[bad])))
    start = torch.cumsum(counts, 0) - counts
    col = torch.arange(int(counts_h.max()), device=cano_part.device)
    pos = torch.where(col[None, :] < counts[:, None], col[None, :], torch.zeros_like(col)[None, :]) + start[:, None]
    src_idx = order[pos]                                                     # (P, Nmax) indices into cano_pc
    picked = fps(cano_pc[src_idx].contiguous(), num_fps).long()              # (P, num_fps) positions in the padded part
    part_idx = torch.gather(src_idx, 1, picked)
    return cano_pc[part_idx], part_idx


def fps_index_list(pc_trans_list: torch.Tensor, cano_part_idx_list: torch.Tensor) -> torch.Tensor:
    """graph_utils.py:59-71 -- the same samples in every posed frame: (T,N,3), (P,S) -> (T,P,S,3)."""
    return pc_trans_list[:, cano_part_idx_list]


def compute_spatial_cost(cano_part_fps_list: torch.Tensor, chamfer_dist, return_index: bool = False):
    """graph_utils.py:74-88 -- closest approach of every ordered part pair in the canonical frame, from a P^2-batch
    one-directional Chamfer search over the samples: -> (P,P) squared distance [, (P,P,2) the sample pair]."""
    P, S = cano_part_fps_list.shape[:2]
    src = cano_part_fps_list[:, None].expand(P, P, S, 3).reshape(-1, S, 3)
    tgt = cano_part_fps_list[None, :].expand(P, P, S, 3).reshape(-1, S, 3)
    dist, nn_idx = chamfer_dist(src, tgt, return_index=True)
    dist_cost, src_idx = dist.reshape(P, P, S).min(dim=2)
    if not return_index:
        return dist_cost
    tgt_idx = torch.gather(nn_idx.reshape(P, P, S), 2, src_idx[:, :, None]).squeeze(2)
    return dist_cost, torch.stack((src_idx, tgt_idx), dim=2)


def compute_joint_cost(part_fps_list, joint_connection, edge_pair_indices) -> torch.Tensor:
    """graph_utils.py:91-104 -- squared distance between the two joint-anchor samples of each edge in every frame:
    (T,P,S,3), (E,2) part ids, (E,2) sample ids -> (T,E) (or (E,) without the time axis)."""
    a = part_fps_list[..., joint_connection[:, 0], edge_pair_indices[:, 0], :]
    b = part_fps_list[..., joint_connection[:, 1], edge_pair_indices[:, 1], :]
    return (a - b).square().sum(dim=-1)


def filter_seg_label(cano_part: torch.Tensor, min_num: int = 10) -> torch.Tensor:
    """graph_utils.py:107-117 -- sorted labels that own at least ``min_num`` points."""
    labels, counts = torch.unique(cano_part, sorted=True, return_counts=True)
    return labels[counts >= min_num]


def denoise_seg_label(cano_part, cano_pc, knn, min_num: int = 10):
    """graph_utils.py:120-128 -- points of parts smaller than ``min_num`` take the label of their nearest neighbour
    among the other points (in place, like the reference)."""
    from .model_utils import knn_query
    labels, inverse, counts = torch.unique(cano_part, sorted=True, return_inverse=True, return_counts=True)
    small = (counts < min_num)[inverse]
    if bool(small.any()):
        cano_part[small] = knn_query(cano_pc[small], cano_pc[~small], cano_part[~small], knn)
    return cano_part


# ----------------------------------------------------------------------------------------- tree search and merging
def mst(cost: torch.Tensor, uni_label=None, max_cost=None, keep_index: bool = False, verbose: bool = False):
    """graph_utils.py:278-309 -- greedy (Kruskal) spanning tree over a dense, possibly asymmetric (P,P) cost: repeatedly
    take the cheapest ordered pair whose ends are in different components (lowest flat index on ties); edges keep the
    orientation (row, column) they were found in.  ``max_cost`` stops early; ``uni_label`` renames the ends.
    One D2H copy of the matrix, then a host loop over P-1 picks (the reference does P-1 device argmins + syncs)."""
    n = cost.shape[0]
    if uni_label is not None:
        assert n == len(uni_label)
    c = cost.detach().cpu().numpy()
    blocked = c.dtype.type(1e10)
    comp = np.arange(n)
    edges = []
    for _ in range(n - 1):
        masked = np.where(comp[:, None] == comp[None, :], blocked, c)
        flat = int(np.argmin(masked))
        i, j = divmod(flat, n)
        if max_cost is not None and masked[i, j] > max_cost:
            break
        if verbose:
            print(i if uni_label is None else int(uni_label[i]), j if uni_label is None else int(uni_label[j]),
                  float(masked[i, j]))
        comp[comp == comp[j]] = comp[i]
        edges.append((i, j))
    out = torch.tensor(edges, dtype=torch.long).reshape(-1, 2)
    if uni_label is not None and not keep_index:
        out = torch.as_tensor(uni_label).to("cpu", torch.long)[out]
    return out.to(cost.device)


def merge_graph(seg_part, joint_connection, trans_list, merge_thr, verbose: bool = True):
    """graph_utils.py:312-366 -- contract tree edges whose two parts move (nearly) rigidly together: an edge's cost
    is the time-mean of || T_src^-1 T_tgt - I ||_F^2; walking the tree in topological order, every edge below
    ``merge_thr`` is contracted and the absorbed part's points relabelled.  -> (new seg_part, new (E',2) edges)."""
    T = trans_list.shape[0]
    src, tgt = trans_list[:, joint_connection[:, 0]], trans_list[:, joint_connection[:, 1]]
    rel = torch.matmul(inverse_transformation(src.reshape(-1, 4, 4)).reshape(T, -1, 4, 4), tgt).reshape(-1, 4, 4)
    eye = torch.eye(4, device=rel.device, dtype=rel.dtype)[None].expand(rel.shape[0], 4, 4)
    edge_cost = frobenius_cost(rel, eye).reshape(T, -1).mean(dim=0).cpu().tolist()
    edges = joint_connection.cpu().tolist()

    tree = nx.DiGraph()
    tree.add_nodes_from(sorted({p for e in edges for p in e}))
    for (a, b), c in zip(edges, edge_cost):
        tree.add_edge(a, b, cost=c)
        if verbose:
            print("add edge {}-{}: cost {}".format(a, b, c))

    merged = tree.copy()
    absorbed = {}                                                 # part id -> the part that swallowed it
    for node in nx.topological_sort(tree):
        if not merged.has_node(node):
            continue
        for a, b in list(merged.edges(node)):
            if merged.has_node(b) and merged.get_edge_data(a, b)["cost"] < merge_thr:
                if verbose:
                    print("merge edge {}-{}: cost {}".format(b, a, merged.get_edge_data(a, b)["cost"]))
                merged = nx.contracted_edge(merged, (a, b), self_loops=False)
                absorbed[b] = a
    if not nx.is_weakly_connected(merged):
        raise ValueError("New graph are not all connected.")
    if not nx.is_directed_acyclic_graph(merged):
        raise ValueError("There are cycles in the link graph")

    new_part = seg_part.clone()
    if absorbed:
        hi = int(max(max(absorbed), int(seg_part.max()))) + 1
        remap = list(range(hi))
        for b, a in absorbed.items():                             # in merge order: later merges see earlier ones
            remap = [a if r == b else r for r in remap]
        new_part = torch.tensor(remap, dtype=seg_part.dtype, device=seg_part.device)[seg_part]
    if verbose:
        for a, b in merged.edges:
            print("remain edge {}-{}: cost {}".format(a, b, merged.get_edge_data(a, b)["cost"]))
    new_connection = torch.tensor([[a, b] for a, b in merged.edges], device=joint_connection.device,
                                  dtype=joint_connection.dtype)
    return new_part, new_connection


def _pair_costs(seg_part, trans_list, cano_pc, chamfer_dist, num_fps, fps, pose_labels=None):
    """Shared front half of merging_wrapper / mst_wrapper: canonical closest approach of every part pair and the
    drift of that joint-anchor pair over time.  Labels are hard, so a sample moves with ONE transform: only the
    P x num_fps samples are posed (the reference skins the whole cloud first, graph_utils.py:370,399).
    ``pose_labels`` are the per-point labels the reference skinned the cloud with: in merging_wrapper that happens
    ONCE before the merge rounds (graph_utils.py:370), so from round 2 on a point absorbed from part b still moves
    with T_b, not with the transform of the part that swallowed it."""
    uni_label = torch.unique(seg_part, sorted=True)
    P = uni_label.numel()
    fps_pts, fps_idx = fps_sample_cano(cano_pc, seg_part, uni_label, num_fps=num_fps, fps=fps)
    mover = (seg_part if pose_labels is None else pose_labels)[fps_idx]                 # (P, S) transform id per sample
    R, t = trans_list[:, mover, :3, :3], trans_list[:, mover, :3, 3]                   # (T,P,S,3,3), (T,P,S,3)
    part_fps = torch.einsum("tpsij,psj->tpsi", R, fps_pts) + t
    cano_dist, pair = compute_spatial_cost(fps_pts, chamfer_dist, return_index=True)
    ids = torch.arange(P, device=pair.device)
    all_pairs = torch.stack(torch.meshgrid(ids, ids, indexing="ij"), dim=2).reshape(-1, 2)
    joint_cost = compute_joint_cost(part_fps, all_pairs, pair.reshape(-1, 2)).reshape(-1, P, P).sum(dim=0)
    return uni_label, cano_dist, joint_cost


def merging_wrapper(seg_part, trans_list, cano_pc, chamfer_dist, merge_thr, n_it: int = 2, fps=None):
    """graph_utils.py:369-394 -- ``n_it`` rounds of: spanning tree on (closest approach + joint drift), then
    ``merge_graph``.  The cloud is posed with the labels at entry in every round (graph_utils.py:370)."""
    entry_labels = seg_part
    for _ in range(n_it):
        uni_label, cano_dist, joint_cost = _pair_costs(seg_part, trans_list, cano_pc, chamfer_dist, 20, fps,
                                                       pose_labels=entry_labels)
        cost = cano_dist + joint_cost
        cost = cost + 1e4 * torch.eye(cost.shape[0], device=cost.device, dtype=cost.dtype)
        candidates = mst(cost, uni_label=uni_label)
        seg_part, _ = merge_graph(seg_part, candidates, trans_list, merge_thr, verbose=False)
        if not len(torch.unique(seg_part)) > 1:
            break
    return seg_part


def mst_wrapper(seg_part, trans, cano_pc, chamfer_dist, verbose: bool = False, num_fps: int = 20,
                cano_dist_thr: float = 1e-2, joint_cost_weight: float = 100, fps=None):
    """graph_utils.py:397-421 -- the kinematic tree: spanning tree on adjacency gate (closest approach under
    ``cano_dist_thr``) + 1-dof geometric cost + weighted joint drift.  -> (P-1,2) edges in original part ids."""
    uni_label, cano_dist, joint_cost = _pair_costs(seg_part, trans, cano_pc, chamfer_dist, num_fps, fps)
    axis, moment, theta, distance, rel = compute_relative_trans(trans, return_trans=True)
    pick = lambda x: x[:, uni_label][:, :, uni_label]
    geo_cost = compute_geo_cost(pick(rel), pick(axis), pick(moment), pick(theta), pick(distance))
    dist_cost = 1e4 * (cano_dist >= cano_dist_thr).to(geo_cost.dtype)
    cost = dist_cost + geo_cost + joint_cost_weight * joint_cost
    cost = cost + 1e4 * torch.eye(cost.shape[0], device=cost.device, dtype=cost.dtype)
    return mst(cost, uni_label=uni_label, verbose=verbose)


# ------------------------------------------------------------------------------------------ kinematic tree set-up
def extract_kinematic(seg_part, trans_list, joint_connection):
    """kinematic_utils.py:20-36 -- renumber the surviving parts 0..P-1 (sorted order) in the labels, the transform
    list and the edges.  (The reference also rewrites ``joint_connection`` in place; a new tensor is returned here.)"""
    uni_label = torch.unique(seg_part, sorted=True)
    assert torch.equal(torch.unique(joint_connection, sorted=True), uni_label)
    return (torch.searchsorted(uni_label, seg_part), trans_list[:, uni_label],
            torch.searchsorted(uni_label, joint_connection.contiguous()))


def to_DAG(G: nx.Graph, root_node) -> nx.DiGraph:
    """kinematic_utils.py:39-54 -- orient an undirected tree child -> parent towards ``root_node``; edges are listed
    in the order the walk from each node (in G's node order) to the root first meets them."""
    parent = {root_node: None}
    frontier = [root_node]
    while frontier:
        nxt = []
        for u in frontier:
            for v in G.neighbors(u):
                if v not in parent:
                    parent[v] = u
                    nxt.append(v)
        frontier = nxt
    edges, seen = [], set()
    for node in G.nodes:
        while parent[node] is not None:
            if (node, parent[node]) not in seen:
                seen.add((node, parent[node]))
                edges.append((node, parent[node]))
            node = parent[node]
    assert len(edges) == G.number_of_nodes() - 1, "invalid tree structure"
    D = nx.from_edgelist(edges, create_using=nx.DiGraph())
    assert len(nx.descendants(D, root_node)) == 0
    return D


def build_graph(edges_list, trans_list, verbose: bool = False, root_part=None, revolute_only: bool = True,
                return_joint_type: bool = False):
    """kinematic_utils.py:57-139 -- initial kinematic model from the tree and the per-part motions.

    edges_list (E,2) part ids 0..P-1, trans_list (T,P,4,4).  The root is the most static part unless given.  Returns
    ``G, root_part, axis_list (E,3), moment_list (E,3), theta_list (T,E)[, distance_list (T,E)], edge_index
    {"child_parent": e}[, joint_type_list]`` exactly like the reference; all E edges are processed in one batch.
    """
    init_G = nx.from_edgelist(edges_list.cpu().numpy(), create_using=nx.Graph())
    uni_label = torch.unique(edges_list, sorted=True)
    assert torch.equal(uni_label, torch.arange(trans_list.shape[1], dtype=uni_label.dtype, device=uni_label.device))
    if root_part is None:
        root_part = int(compute_root_cost(trans_list).argmin())
    if verbose:
        print("root part id", root_part)
    G = to_DAG(init_G, root_node=root_part)
    pairs = [(int(c), int(p)) for c, p in G.edges()]
    edge_index = {"{}_{}".format(c, p): e for e, (c, p) in enumerate(pairs)}
    T, E = trans_list.shape[0], len(pairs)
    child = torch.tensor([c for c, _ in pairs], dtype=torch.long, device=trans_list.device)
    par = torch.tensor([p for _, p in pairs], dtype=torch.long, device=trans_list.device)

    inv_parent = inverse_transformation(trans_list[:, par].reshape(-1, 4, 4)).reshape(T, E, 4, 4)
    rel = torch.matmul(inv_parent, trans_list[:, child])
    l, m, theta, d = dq_to_screw(transform_to_dq(rel.reshape(-1, 4, 4)))
    l, m, theta, d = l.reshape(T, E, 3), m.reshape(T, E, 3), theta.reshape(T, E), d.reshape(T, E)
    axis_list, moment_list = l.mean(dim=0), m.mean(dim=0)            # per-edge plain time mean (E == 1 case of :196)

    if revolute_only:
        no_rot = torch.logical_or(theta.abs() < _EPS, (theta - math.pi).abs() < _EPS)
        assert int(no_rot.sum()) == 0
        joint_types = ["revolute"] * E
        if verbose:
            print("joint types at each edge: {}".format(joint_types))
        return G, root_part, axis_list, moment_list, theta, edge_index

    _, _, cost_r, cost_p1, rot_res = _joint_fit(rel, axis_list, moment_list, theta, d)
    prismatic = (cost_p1 + rot_res.mean(dim=0)) <= cost_r
    pinned = torch.full_like(theta, _EPS)
    theta_list = torch.where(prismatic[None], pinned, theta)
    distance_list = torch.where(prismatic[None], d, pinned)
    joint_types = ["prismatic" if p else "revolute" for p in prismatic.cpu().tolist()]
    if verbose:
        print("joint types at each edge: {}".format(joint_types))
    if return_joint_type:
        return G, root_part, axis_list, moment_list, theta_list, distance_list, edge_index, joint_types
    return G, root_part, axis_list, moment_list, theta_list, distance_list, edge_index


def edge_index2edges(edge_index) -> list:
    """kinematic_utils.py:142-148 -- {"child_parent": e} -> [[child, parent], ...] in dict order."""
    return [[int(x) for x in name.split("_")] for name in edge_index]
