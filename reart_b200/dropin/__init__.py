"""Native-module drop-ins so the UNMODIFIED reference scripts run on a B200 box.

The reference hard-imports three compiled modules by name:
``chamferdist`` (with ``_C``), ``knn_cuda`` (with ``KNN``) and ``pointnet2_cuda``.
``install()`` registers equivalents backed by ``libreart_b200.so`` in ``sys.modules``; after that
``run_robot.py`` / ``run_real.py`` / ``run_sapien.py`` import and run unchanged with the reference tree on
``sys.path``.
"""
from __future__ import annotations

import sys
import types

import torch

from .. import _lib
from ..knn_module import KNN


# The reference's ChamferDistance.forward runs TWO searches back to back, knn_points(src, tgt) then
# knn_points(tgt, src) (utils/chamfer.py:78-94).  Every distance is symmetric, so the first call evaluates each
# pair once for both directions (chamfer_sym.cu) and parks the reverse result here; the second call is answered from
# the cache.  The pairing is pinned to ONE invocation of ChamferDistance.forward: the cache entry remembers the Python
# frame object of the enclosing forward() and is honoured only by a call made from that very frame (plus the swapped
# (data_ptr, version, shape) signature).  Outside a ChamferDistance.forward nothing is cached, so raw-pointer writers
# that do not bump tensor versions can never be served a stale result by a later, unrelated call.  The entry holds
# strong references to both tensors (their storage cannot be recycled while it is alive); any other call drops it.
_reverse_cache = {"sig": None, "idx": None, "dists": None, "keep": None, "frame": None}


def _sig(a, b):
    return (a.data_ptr(), a._version, tuple(a.shape), b.data_ptr(), b._version, tuple(b.shape), a.device.index)


def _enclosing_chamfer_forward():
    """The frame of the ChamferDistance.forward this call is (transitively) made from, or None."""
    f = sys._getframe(2)
    for _ in range(16):
        if f is None:
            return None
        if f.f_code.co_name == "forward" and type(f.f_locals.get("self")).__name__ == "ChamferDistance":
            return f
        f = f.f_back
    return None


def _drop_cache():
    _reverse_cache.update(sig=None, idx=None, dists=None, keep=None, frame=None)


def _knn_points_idx(p1, p2, lengths1, lengths2, K, version):
    """chamferdist._C.knn_points_idx as called at utils/chamfer.py:174 -> (idx [B,P1,K], dists [B,P1,K])."""
    if K != 1:
        raise NotImplementedError("reart_b200 chamferdist._C drop-in implements K == 1 (the reference's only use)")
    if p1.shape[2] != 3:
        raise ValueError("reart_b200 chamferdist._C drop-in is specialised for D == 3")
    _lib.require_cuda(p1, p2)
    L = _lib.lib()
    p1c, p2c = p1.float().contiguous(), p2.float().contiguous()
    B, P1, _ = p1c.shape
    P2 = p2c.shape[1]
    fwd_frame = _enclosing_chamfer_forward()
    if fwd_frame is not None and _reverse_cache["frame"] is fwd_frame and _reverse_cache["sig"] == _sig(p1c, p2c):
        idx, dists = _reverse_cache["idx"], _reverse_cache["dists"]
        _drop_cache()
        return idx, dists
    _drop_cache()
    dists = torch.empty(B, P1, 1, dtype=torch.float32, device=p1.device)
    idx = torch.empty(B, P1, 1, dtype=torch.int64, device=p1.device)
    if fwd_frame is not None and P1 >= 256 and P2 >= 256:
        # the first search of a ChamferDistance.forward: compute both directions now
        rd = torch.empty(B, P2, 1, dtype=torch.float32, device=p1.device)
        ri = torch.empty(B, P2, 1, dtype=torch.int64, device=p1.device)
        nbytes = L.reart_chamfer_workspace_bytes(B, P1, P2)
        ws = _lib.workspace(nbytes, p1.device)
        with torch.cuda.device(p1.device):
            _lib.check(L.reart_chamfer_bidir_fwd(_lib.ptr(p1c), _lib.ptr(p2c), B, P1, P2, _lib.ptr(dists), _lib.ptr(idx),
                                                 _lib.ptr(rd), _lib.ptr(ri), _lib.ptr(ws), nbytes, _lib.stream_ptr()),
                       "reart_chamfer_bidir_fwd")
        _reverse_cache.update(sig=_sig(p2c, p1c), idx=ri, dists=rd, keep=(p1c, p2c), frame=fwd_frame)
        return idx, dists
    nbytes = L.reart_knn1_workspace_bytes(B, P1, P2)
    ws = _lib.workspace(nbytes, p1.device)
    with torch.cuda.device(p1.device):
        _lib.check(L.reart_knn1_fwd(_lib.ptr(p1c), _lib.ptr(p2c), B, P1, P2, _lib.ptr(dists), _lib.ptr(idx),
                                    _lib.ptr(ws), nbytes, _lib.stream_ptr()), "reart_knn1_fwd")
    return idx, dists


def _knn_points_backward(p1, p2, lengths1, lengths2, idx, grad_dists):
    """chamferdist._C.knn_points_backward as called at utils/chamfer.py:206-208 -> (grad_p1, grad_p2)."""
    if idx.shape[2] != 1:
        raise NotImplementedError("reart_b200 chamferdist._C drop-in implements K == 1")
    _lib.require_cuda(p1, p2, idx, grad_dists)
    L = _lib.lib()
    p1c, p2c = p1.float().contiguous(), p2.float().contiguous()
    B, P1, _ = p1c.shape
    P2 = p2c.shape[1]
    g = grad_dists.float().contiguous()
    idxc = idx.contiguous()
    g1, g2 = torch.empty_like(p1c), torch.empty_like(p2c)
    with torch.cuda.device(p1.device):
        _lib.check(L.reart_knn1_bwd(_lib.ptr(p1c), _lib.ptr(p2c), _lib.ptr(idxc), _lib.ptr(g), B, P1, P2, _lib.ptr(g1),
                                    _lib.ptr(g2), _lib.stream_ptr()), "reart_knn1_bwd")
    return g1, g2


def _furthest_point_sampling_wrapper(B, N, npoint, xyz, temp, out):
    """pointnet2_cuda.furthest_point_sampling_wrapper (networks/pointnet_lib/pointnet2_utils.py:29)."""
    from ..ops import fps_into
    fps_into(xyz, int(npoint), out, temp=temp if torch.is_tensor(temp) else None)
    return 1


def _ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx):
    """pointnet2_cuda.ball_query_wrapper (networks/pointnet_lib/pointnet2_utils.py:263)."""
    from ..ops import ball_query_into
    ball_query_into(new_xyz, xyz, float(radius), int(nsample), idx)
    return 1


def install(force: bool = False) -> None:
    """Register ``chamferdist``, ``chamferdist._C``, ``knn_cuda`` and ``pointnet2_cuda`` in ``sys.modules``."""
    if force or "chamferdist" not in sys.modules:
        pkg = types.ModuleType("chamferdist")
        c = types.ModuleType("chamferdist._C")
        c.knn_points_idx = _knn_points_idx
        c.knn_points_backward = _knn_points_backward
        pkg._C = c
        pkg.__reart_b200__ = True
        sys.modules["chamferdist"] = pkg
        sys.modules["chamferdist._C"] = c
    if force or "knn_cuda" not in sys.modules:
        k = types.ModuleType("knn_cuda")
        k.KNN = KNN
        k.__reart_b200__ = True
        sys.modules["knn_cuda"] = k
    if force or "pointnet2_cuda" not in sys.modules:
        pn = types.ModuleType("pointnet2_cuda")
        pn.furthest_point_sampling_wrapper = _furthest_point_sampling_wrapper
        pn.ball_query_wrapper = _ball_query_wrapper
        pn.__reart_b200__ = True
        sys.modules["pointnet2_cuda"] = pn
