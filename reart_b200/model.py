"""Drop-in for the reference's ``networks/model.py``: ``BaseModel`` (:11-70) and ``KinematicModel`` (:73-166).

Same constructor arguments, ``forward`` keyword arguments, return triple
``(pc_trans_list [T,N,3], seg_part [N], trans_list [T,P,4,4])`` and ``state_dict`` keys (so the shipped
checkpoints load), with the skinning, 6D->R and forward kinematics running in the fused sm_100a kernels.
RNG (``F.gumbel_softmax``), the seg MLP and the optimiser stay in torch.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .kinematic import FlatTree
from .model_utils import create_transformation, knn_query, th_with_zeros
from .screw_se3 import matrix_to_rotation_6d, rotation_6d_to_matrix


class SegHead(nn.Module):
    """MLPConv1d(3, (128, P), bn=False, gn=False, last_activation='none') of networks/blocks.py:99-118:
    Conv1d(3,128,1,bias) -> ReLU -> Conv1d(128,P,1,no bias); parameter names ``model.0.*`` / ``model.2.weight``."""

    def __init__(self, in_channel: int, hidden: int, out_channel: int):
        super().__init__()
        self.model = nn.Sequential(
            nn.Conv1d(in_channel, hidden, kernel_size=1, bias=True),
            nn.ReLU(inplace=True),
            nn.Conv1d(hidden, out_channel, kernel_size=1, bias=False),
        )
        self.out_channel = out_channel

    def forward(self, x):
        return self.model(x)

    def logits(self, points: torch.Tensor) -> torch.Tensor:
        """[N,3] -> [N,P].  A k=1 Conv1d IS a per-point matmul: the two layers run as one fused fp32 kernel each way
        (csrc/segmlp.cu) on the conv parameters themselves -- identical state_dict; it also avoids cuDNN's default
        TF32, so the logits match the reference's CPU arithmetic to fp32 round-off.  CUDA tensors only (no fallback)."""
        w0, b0, w2 = self.model[0].weight[:, :, 0], self.model[0].bias, self.model[2].weight[:, :, 0]
        return ops.seg_mlp(points, w0, b0, w2, getattr(self, "grad_sink", None))


def _assemble(R: torch.Tensor, tr: torch.Tensor) -> torch.Tensor:
    """[T,P,3,3], [T,P,3] -> [T,P,4,4] (th_with_zeros of networks/model.py:66-68)."""
    T, P = R.shape[:2]
    top = torch.cat((R.reshape(T * P, 3, 3), tr.reshape(T * P, 3, 1)), dim=2)
    return th_with_zeros(top).reshape(T, P, 4, 4)


class BaseModel(nn.Module):
    """Relaxation model (networks/model.py:11-70)."""

    def __init__(self, num_parts, pose_len, joint_trajectory=None, init_6d=None, init_t=None):
        super().__init__()
        if joint_trajectory is not None:
            raise NotImplementedError("joint_trajectory is a dead feature of the reference (SURVEY Q6)")
        self.num_parts = num_parts
        self.pose_len = pose_len
        initial_connection = torch.stack([torch.arange(num_parts - 1), torch.arange(num_parts - 1) + 1], dim=1)
        self.register_buffer("joint_connection", initial_connection.long())
        self.seg_head = SegHead(3, 128, num_parts)
        self.joint_trajectory = None
        if init_6d is None:
            ident = torch.tensor([[[1.0, 0, 0, 0, 1, 0]]]).repeat(pose_len, num_parts, 1)
            self.proposal_6d = nn.Parameter(ident, requires_grad=True)
        else:
            self.proposal_6d = nn.Parameter(init_6d, requires_grad=False)
        if init_t is None:
            self.proposal_t = nn.Parameter(torch.zeros(pose_len, num_parts, 3), requires_grad=True)
        else:
            self.proposal_t = nn.Parameter(init_t, requires_grad=False)

    def seg_logits(self, cano_pc):
        return self.seg_head.logits(cano_pc)                         # [N,P] (networks/model.py:42-43)

    def seg_forward(self, cano_pc, **kwargs):
        seg = self.seg_logits(cano_pc)
        return seg.argmax(dim=-1) if kwargs.get("argmax") else seg

    def pose(self, **kwargs):
        """(R [T,P,3,3], t [T,P,3]) from the 6D proposals (or the kwargs overrides used by ik())."""
        proposal_6d = kwargs.get("proposal_6d", self.proposal_6d)
        proposal_t = kwargs.get("proposal_t", self.proposal_t)
        T = proposal_6d.shape[0]
        R = rotation_6d_to_matrix(proposal_6d.reshape(-1, 6)).reshape(T, self.num_parts, 3, 3)
        return R, proposal_t.reshape(T, self.num_parts, 3)

    def weights(self, cano_pc, tau=1.0):
        """Straight-through gumbel-softmax assignment (networks/model.py:42-44); draws RNG every call."""
        seg = self.seg_logits(cano_pc)
        return seg, ops.gumbel_softmax_st(seg, tau)

    def forward(self, cano_pc, **kwargs):
        seg, weight = self.weights(cano_pc, tau=kwargs.get("tau", 1.0))
        R, tr = self.pose(**kwargs)
        pc_trans_list = ops.skin(cano_pc, weight, R, tr)              # [T,N,3]
        return pc_trans_list, seg.argmax(dim=-1), _assemble(R, tr)


class KinematicModel(nn.Module):
    """Projection model (networks/model.py:73-166): fixed segmentation, screw joints on a tree."""

    def __init__(self, pose_len, seg_part, cano_pc, knn, **kwargs):
        super().__init__()
        self.seg_part = seg_part.long()
        self.cano_pc = cano_pc
        self.pose_len = pose_len
        self.num_parts = len(torch.unique(self.seg_part))
        self.knn = knn
        if knn is not None:
            assert self.knn.k == 1
        self.edge_index = kwargs["edge_index"]
        self.paths_to_base = kwargs["paths_to_base"]
        self.reverse_topo = kwargs["reverse_topo"]
        E = len(self.edge_index)
        assert self.num_parts == E + 1

        def param(name, default):
            if name in kwargs:
                return nn.Parameter(kwargs[name], requires_grad=True)
            return nn.Parameter(default, requires_grad=True)

        if "axis_list" in kwargs:
            assert self.num_parts == kwargs["axis_list"].shape[0] + 1
        if "moment_list" in kwargs:
            assert self.num_parts == kwargs["moment_list"].shape[0] + 1
        self.axis_list = param("axis_list", torch.zeros(E, 3))
        self.moment_list = param("moment_list", torch.zeros(E, 3))
        self.theta_list = param("theta_list", torch.zeros(pose_len, E))
        if "distance_list" in kwargs:
            self.distance_list = nn.Parameter(kwargs["distance_list"], requires_grad=True)
        elif kwargs.get("load_distance"):
            self.distance_list = nn.Parameter(torch.zeros(pose_len, E), requires_grad=True)
        if "root_trans" in kwargs:
            self.root_6d = nn.Parameter(matrix_to_rotation_6d(kwargs["root_trans"][:, :3, :3]), requires_grad=True)
            self.root_t = nn.Parameter(kwargs["root_trans"][:, :3, 3], requires_grad=True)
        elif kwargs.get("load_root_trans"):
            self.root_6d = nn.Parameter(torch.tensor([[1.0, 0, 0, 0, 1, 0]]).repeat(pose_len, 1), requires_grad=True)
            self.root_t = nn.Parameter(torch.zeros(pose_len, 3), requires_grad=True)
        self.joint_type_list = kwargs.get("joint_type_list")
        self._tree = None
        self._label_cache = None

    def _apply(self, fn, *args, **kwargs):
        # .to(device) must also move the plain-tensor attributes the reference keeps outside state_dict
        out = super()._apply(fn, *args, **kwargs)
        self.seg_part = fn(self.seg_part)
        self.cano_pc = fn(self.cano_pc)
        self._tree = None
        self._label_cache = None
        return out

    def tree(self, device) -> FlatTree:
        if self._tree is None or self._tree.order.device != device:
            self._tree = FlatTree(self.paths_to_base, self.reverse_topo, self.edge_index, self.joint_type_list,
                                  device=device)
        return self._tree

    def seg_forward(self, input_pc, **kwargs):
        return self._labels(input_pc)

    def _labels(self, input_pc):
        """k=1 label transfer (networks/model.py:138).  When the query IS the stored canonical cloud the
        answer is constant over the optimisation: it is computed once and cached."""
        if input_pc is self.cano_pc or (input_pc.shape == self.cano_pc.shape and input_pc.data_ptr() == self.cano_pc.data_ptr()):
            key = (self.cano_pc.data_ptr(), self.cano_pc._version, self.seg_part.data_ptr(), self.seg_part._version)
            if self._label_cache is None or self._label_cache[0] != key:      # also invalidated by reassigning seg_part
                self._label_cache = (key, knn_query(input_pc, self.cano_pc, self.seg_part, self.knn))
            return self._label_cache[1]
        return knn_query(input_pc, self.cano_pc, self.seg_part, self.knn)

    def transforms(self, **kwargs):
        """trans_list [T,P,4,4]: fused tree FK (+ optional learned root pose, networks/model.py:151-159)."""
        theta_list = kwargs.get("theta_list", self.theta_list)
        distance_list = self.distance_list if hasattr(self, "distance_list") else None
        tree = self.tree(theta_list.device)
        trans_list = ops.fk_flat(self.axis_list, self.moment_list, theta_list, distance_list, tree.order, tree.parent,
                                 tree.edge, tree.joint_type)
        if hasattr(self, "root_6d") and hasattr(self, "root_t"):
            root = create_transformation(rotation_6d_to_matrix(self.root_6d), self.root_t[:, :, None])
            trans_list = torch.matmul(root[:, None, :, :].expand(trans_list.shape), trans_list)
        return trans_list

    def forward(self, input_pc, **kwargs):
        seg_part = self._labels(input_pc)
        weight = F.one_hot(seg_part, num_classes=self.num_parts)       # int64, SURVEY Q7
        trans_list = self.transforms(**kwargs)
        pc_trans_list = ops.skin(input_pc, weight, trans_list[:, :, :3, :3], trans_list[:, :, :3, 3])
        return pc_trans_list, seg_part, trans_list
