"""Deterministic synthetic sequences of the shapes named in BASELINE.json.

An articulated object = P boxes on a random joint tree, surface-sampled to N canonical points in roughly
[-0.35, 0.35]^3 (the nao range); every frame poses the parts by composing per-joint rotations about the
joint anchors (theta ~ U(-1,1) rad) and re-samples M observed points with N(0, 1e-3) noise.  numpy only,
seed 2 by default (the reference's --manual_seed default, run_robot.py:364).
"""
from __future__ import annotations

import numpy as np


def _rodrigues(axis, angle):
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def make_sequence(T: int, N: int, P: int, seed: int = 2, M: int | None = None, noise: float = 1e-3):
    """Returns dict(cano [N,3] f32, part [N] i64, frames [T,M,3] f32, pose [T,P,4,4] f32 GT part poses,
    parent [P], anchors [P,3], axes [P,3])."""
    M = N if M is None else M
    rng = np.random.default_rng(seed)
    parent = np.full(P, -1, np.int64)
    centers = np.zeros((P, 3))
    half = np.zeros((P, 3))
    anchors = np.zeros((P, 3))
    axes = rng.standard_normal((P, 3))
    half[0] = rng.uniform(0.04, 0.09, 3)
    for p in range(1, P):
        par = int(rng.integers(0, p))
        parent[p] = par
        half[p] = rng.uniform(0.02, 0.07, 3)
        direction = rng.standard_normal(3)
        direction /= np.linalg.norm(direction)
        anchors[p] = centers[par] + direction * half[par]
        centers[p] = anchors[p] + direction * half[p]
    scale = 0.33 / max(np.abs(centers).max() + half.max(), 1e-6)
    centers *= scale; half *= scale; anchors *= scale

    def sample(n, r):
        part = r.integers(0, P, n)
        face = r.integers(0, 6, n)
        u = r.uniform(-1, 1, (n, 3))
        ax = face % 3
        u[np.arange(n), ax] = np.where(face < 3, -1.0, 1.0)
        return centers[part] + u * half[part], part

    cano, part = sample(N, rng)
    theta = rng.uniform(-1, 1, (T, P))
    pose = np.tile(np.eye(4), (T, P, 1, 1))
    order = list(range(P))                                    # parents have smaller ids by construction
    for t in range(T):
        for p in order[1:]:
            Rl = _rodrigues(axes[p], theta[t, p])
            local = np.eye(4)
            local[:3, :3] = Rl
            local[:3, 3] = anchors[p] - Rl @ anchors[p]
            pose[t, p] = pose[t, parent[p]] @ local
    frames = np.zeros((T, M, 3))
    for t in range(T):
        pts, prt = sample(M, rng)
        Rm = pose[t, prt, :3, :3]
        frames[t] = np.einsum("nij,nj->ni", Rm, pts) + pose[t, prt, :3, 3] + rng.normal(0, noise, (M, 3))
    return dict(cano=cano.astype(np.float32), part=part.astype(np.int64), frames=frames.astype(np.float32),
                pose=pose.astype(np.float32), parent=parent, anchors=anchors.astype(np.float32),
                axes=axes.astype(np.float32), theta=theta.astype(np.float32))


def make_flow_reference(seq: dict, cano_idx: int = 0, n_ref: int | None = None, seed: int = 2, noise: float = 5e-4):
    """Synthetic stand-in for the correspondence precompute of run_robot.py:75-84: for every consecutive pair of
    the COMPLETE sequence (canonical frame inserted at cano_idx) a set of reference points on frame t and their
    flow to frame t+1, obtained by posing canonical surface points with the ground-truth part poses.
    Returns (pc_ref_list, flow_ref_list): lists of [n_t,3] float32 arrays with ragged n_t."""
    rng = np.random.default_rng(seed + 17)
    cano, part, pose = seq["cano"], seq["part"], seq["pose"]
    T = pose.shape[0]
    n_ref = n_ref or max(16, cano.shape[0] // 4)
    ident = np.tile(np.eye(4, dtype=np.float32), (1, pose.shape[1], 1, 1))
    complete = np.concatenate([pose[:cano_idx], ident, pose[cano_idx:]], axis=0)          # [T+1,P,4,4]
    refs, flows = [], []
    for t in range(T):
        n = int(n_ref * rng.uniform(0.7, 1.0))
        idx = rng.choice(cano.shape[0], n, replace=False)
        def posed(k):
            Rm = complete[k, part[idx], :3, :3]
            return np.einsum("nij,nj->ni", Rm, cano[idx]) + complete[k, part[idx], :3, 3]
        a, b = posed(t), posed(t + 1)
        refs.append((a + rng.normal(0, noise, a.shape)).astype(np.float32))
        flows.append((b - a + rng.normal(0, noise, a.shape)).astype(np.float32))
    return refs, flows


def kinematic_init(seq: dict, theta_noise: float = 0.1, seed: int = 2):
    """Screw parameters of the synthetic joint tree in the reference's KinematicModel convention
    (networks/model.py:73-135): part p hangs on parent[p] by a revolute joint about `axes[p]` through `anchors[p]`
    => l = axis / |axis|, m = anchor x l, theta[t, e] = joint angle.  Returns kwargs for KinematicModel /
    KinematicEngine (numpy arrays; thetas perturbed by N(0, theta_noise)) and the part labels."""
    rng = np.random.default_rng(seed + 31)
    parent, axes, anchors, theta = seq["parent"], seq["axes"], seq["anchors"], seq["theta"]
    P = len(parent)
    l = axes / np.linalg.norm(axes, axis=1, keepdims=True)
    edge_index, axis_list, moment_list, cols = {}, [], [], []
    for p in range(1, P):
        edge_index[f"{p}_{int(parent[p])}"] = p - 1
        axis_list.append(l[p]); moment_list.append(np.cross(anchors[p], l[p])); cols.append(p)
    paths = {}
    for p in range(P):
        path, x = [p], p
        while parent[x] >= 0:
            x = int(parent[x]); path.append(x)
        paths[p] = path
    th = theta[:, cols] + rng.normal(0, theta_noise, (theta.shape[0], P - 1))
    return dict(edge_index=edge_index, paths_to_base=paths, reverse_topo=list(range(P)),
                axis_list=np.asarray(axis_list, np.float32), moment_list=np.asarray(moment_list, np.float32),
                theta_list=th.astype(np.float32))
