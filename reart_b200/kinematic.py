"""Drop-in for ``fk`` of the reference's ``utils/kinematic_utils.py`` (:151-198)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops

_JOINT_CODE = {"revolute": 1, "prismatic": 2}


def flatten_tree(paths_to_base, reverse_topo, edge_index):
    """dict-based joint tree -> (order, parent, edge) int32 numpy arrays (see include/reart_b200.h: reart_fk_fwd)."""
    P = len(reverse_topo)
    assert sorted(int(p) for p in reverse_topo) == list(range(P))
    order = np.array([int(p) for p in reverse_topo], np.int32)
    parent = -np.ones(P, np.int32)
    edge = -np.ones(P, np.int32)
    for part, path in paths_to_base.items():
        part = int(part)
        if len(path) > 1:
            par = int(path[1])
            parent[part] = par
            edge[part] = int(edge_index["_".join([str(part), str(par)])])
    seen = set()
    for c in order:                                   # parents must precede children (reverse_topo is root-first)
        assert parent[c] < 0 or int(parent[c]) in seen, "reverse_topo must list parents before children"
        seen.add(int(c))
    return order, parent, edge


class FlatTree:
    """Device-resident flattened tree, built once per model (the reference walks dicts every call)."""

    def __init__(self, paths_to_base, reverse_topo, edge_index, joint_type_list=None, device="cuda"):
        order, parent, edge = flatten_tree(paths_to_base, reverse_topo, edge_index)
        self.P = len(order)
        self.order = torch.from_numpy(order).to(device)
        self.parent = torch.from_numpy(parent).to(device)
        self.edge = torch.from_numpy(edge).to(device)
        self.joint_type = None
        if joint_type_list is not None:
            jt = np.array([_JOINT_CODE.get(j, 1) for j in joint_type_list], np.int32)
            self.joint_type = torch.from_numpy(jt).to(device)

    def to(self, device):
        for k in ("order", "parent", "edge", "joint_type"):
            v = getattr(self, k)
            if v is not None:
                setattr(self, k, v.to(device))
        return self


_tree_cache = {}


def fk(paths_to_base, reverse_topo, edge_index, axis_list, moment_list, theta_list,
       distance_list=None, joint_type_list=None):
    """Same signature as utils/kinematic_utils.py:151-152 -> fk_trans_list (T, P, 4, 4), ordered by part id.

    fk[root] = I, fk[c] = fk[parent(c)] @ exp(xi_e(theta[t,e], d[t,e])); one fused kernel forward and one
    backward instead of ~10.5k ATen ops.
    """
    # flattening is a few dozen dict look-ups; the device copies are cached by CONTENT (ids of dicts can be recycled)
    order, parent, edge = flatten_tree(paths_to_base, reverse_topo, edge_index)
    key = (order.tobytes(), parent.tobytes(), edge.tobytes(),
           None if joint_type_list is None else tuple(joint_type_list), str(theta_list.device))
    tree = _tree_cache.get(key)
    if tree is None:
        tree = FlatTree(paths_to_base, reverse_topo, edge_index, joint_type_list, device=theta_list.device)
        if len(_tree_cache) > 64:
            _tree_cache.clear()
        _tree_cache[key] = tree
    if joint_type_list is not None and distance_list is None and any(j == "prismatic" for j in joint_type_list):
        raise ValueError("prismatic joints need distance_list (utils/kinematic_utils.py:176-179)")
    return ops.fk_flat(axis_list, moment_list, theta_list, distance_list, tree.order, tree.parent, tree.edge,
                       tree.joint_type)
