"""CPU oracle for the reart hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.  The product
(``reart_b200``) never imports it and has no CPU fallback.

``oracle/reart_oracle.c`` is the C restatement (each function cites the reference
file:line it follows); this module is its ctypes/numpy binding.  Parity status:
the K=1 / k-NN boundary (third-party ``chamferdist._C`` and ``knn_cuda``, absent from
the reference tree and unpinned there) is "parity unpinned" by the reference itself;
everything else is pinned by ``tests/golden/`` vectors generated from the reference's
own Python (``oracle/make_golden.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libreart_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force: bool = False) -> str:
    """Compile oracle/reart_oracle.c -> oracle/_build/libreart_oracle.so (gcc, OpenMP)."""
    src = os.path.join(_HERE, "reart_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_chamfer_bidir_fwd_bwd.restype = ctypes.c_double
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(int(n))


def knn1(p1, p2):
    """utils/chamfer.py:174 -- returns (dists [B,P1] f32, idx [B,P1] i64)."""
    p1, p2 = _f32(p1), _f32(p2)
    B, P1, _ = p1.shape
    P2 = p2.shape[1]
    d = np.empty((B, P1), np.float32)
    i = np.empty((B, P1), np.int64)
    lib().oracle_knn1(_p(p1, _f32p), _p(p2, _f32p), ctypes.c_int64(B), ctypes.c_int64(P1), ctypes.c_int64(P2),
                      _p(d, _f32p), _p(i, _i64p))
    return d, i


def knn1_bwd(p1, p2, idx, grad_dists):
    """utils/chamfer.py:206-208 -- returns (grad_p1, grad_p2)."""
    p1, p2, g = _f32(p1), _f32(p2), _f32(grad_dists)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    B, P1, _ = p1.shape
    P2 = p2.shape[1]
    g1 = np.empty_like(p1)
    g2 = np.empty_like(p2)
    lib().oracle_knn1_bwd(_p(p1, _f32p), _p(p2, _f32p), _p(idx, _i64p), _p(g, _f32p), ctypes.c_int64(B),
                          ctypes.c_int64(P1), ctypes.c_int64(P2), _p(g1, _f32p), _p(g2, _f32p))
    return g1, g2


def chamfer_bidir_fwd_bwd(src, tgt, want_grad=True):
    """networks/loss.py:24-29 -- returns dict(loss, d_fwd, i_fwd, d_bwd, i_bwd, grad_src, grad_tgt)."""
    src, tgt = _f32(src), _f32(tgt)
    B, N, _ = src.shape
    M = tgt.shape[1]
    d_f = np.empty((B, N), np.float32); i_f = np.empty((B, N), np.int64)
    d_b = np.empty((B, M), np.float32); i_b = np.empty((B, M), np.int64)
    gs = np.empty_like(src) if want_grad else None
    gt = np.empty_like(tgt) if want_grad else None
    loss = lib().oracle_chamfer_bidir_fwd_bwd(_p(src, _f32p), _p(tgt, _f32p), ctypes.c_int64(B), ctypes.c_int64(N),
                                              ctypes.c_int64(M), _p(d_f, _f32p), _p(i_f, _i64p), _p(d_b, _f32p),
                                              _p(i_b, _i64p), _p(gs, _f32p), _p(gt, _f32p))
    return dict(loss=float(loss), d_fwd=d_f, i_fwd=i_f, d_bwd=d_b, i_bwd=i_b, grad_src=gs, grad_tgt=gt)


def skin_fwd(cano, W, R, tr):
    """networks/model.py:63-69 -- cano [N,3], W [N,P], R [T,P,3,3], tr [T,P,3] -> [T,N,3]."""
    cano, W, R, tr = _f32(cano), _f32(W), _f32(R), _f32(tr)
    T, P = R.shape[:2]
    N = cano.shape[0]
    out = np.empty((T, N, 3), np.float32)
    lib().oracle_skin_fwd(_p(cano, _f32p), _p(W, _f32p), _p(R, _f32p), _p(tr, _f32p), ctypes.c_int64(T),
                          ctypes.c_int64(N), ctypes.c_int64(P), _p(out, _f32p))
    return out


def skin_bwd(cano, W, R, tr, g):
    cano, W, R, tr, g = _f32(cano), _f32(W), _f32(R), _f32(tr), _f32(g)
    T, P = R.shape[:2]
    N = cano.shape[0]
    gW = np.empty((N, P), np.float32); gR = np.empty((T, P, 3, 3), np.float32); gt = np.empty((T, P, 3), np.float32)
    lib().oracle_skin_bwd(_p(cano, _f32p), _p(W, _f32p), _p(R, _f32p), _p(tr, _f32p), _p(g, _f32p), ctypes.c_int64(T),
                          ctypes.c_int64(N), ctypes.c_int64(P), _p(gW, _f32p), _p(gR, _f32p), _p(gt, _f32p))
    return gW, gR, gt


def rot6d(d6):
    """screw_se3/geo_utils.py:632-651."""
    d6 = _f32(d6)
    shp = d6.shape[:-1]
    flat = d6.reshape(-1, 6)
    R = np.empty((flat.shape[0], 3, 3), np.float32)
    lib().oracle_rot6d(_p(flat, _f32p), ctypes.c_int64(flat.shape[0]), _p(R, _f32p))
    return R.reshape(shp + (3, 3))


def screw_to_transform(l, m, theta, d):
    """screw_se3/screw_utils.py:6-30 composed -- [B,3],[B,3],[B],[B] -> [B,4,4]."""
    l, m, theta, d = _f32(l), _f32(m), _f32(theta), _f32(d)
    B = l.shape[0]
    M = np.empty((B, 4, 4), np.float32)
    lib().oracle_screw_to_transform(_p(l, _f32p), _p(m, _f32p), _p(theta, _f32p), _p(d, _f32p), ctypes.c_int64(B),
                                    _p(M, _f32p))
    return M


def fk(axis, moment, theta, distance, order, parent, edge, joint_type=None):
    """utils/kinematic_utils.py:151-198 on the flattened tree -> [T,P,4,4]."""
    axis, moment, theta = _f32(axis), _f32(moment), _f32(theta)
    distance = _f32(distance) if distance is not None else None
    order = np.ascontiguousarray(order, np.int32); parent = np.ascontiguousarray(parent, np.int32)
    edge = np.ascontiguousarray(edge, np.int32)
    jt = np.ascontiguousarray(joint_type, np.int32) if joint_type is not None else None
    T = theta.shape[0]
    P = order.shape[0]
    out = np.empty((T, P, 4, 4), np.float32)
    lib().oracle_fk(_p(axis, _f32p), _p(moment, _f32p), _p(theta, _f32p), _p(distance, _f32p), _p(order, _i32p),
                    _p(parent, _i32p), _p(edge, _i32p), _p(jt, _i32p), ctypes.c_int64(T), ctypes.c_int64(P),
                    _p(out, _f32p))
    return out


def knn(ref, query, k):
    """knn_cuda.KNN(k, transpose_mode=True)(ref, query) for one batch -> (dist [m,k] Euclidean, idx [m,k])."""
    ref, query = _f32(ref), _f32(query)
    n, m = ref.shape[0], query.shape[0]
    assert 1 <= k <= 8
    d = np.empty((m, k), np.float32); i = np.empty((m, k), np.int64)
    lib().oracle_knn(_p(ref, _f32p), _p(query, _f32p), ctypes.c_int64(n), ctypes.c_int64(m), ctypes.c_int(k),
                     _p(d, _f32p), _p(i, _i64p))
    return d, i


def blend_anchor_motion(query, ref, flow, k=3):
    """utils/flow_utils.py:147-170 with return_mask=True -> (blended [m,3], mask [m] bool)."""
    query, ref, flow = _f32(query), _f32(ref), _f32(flow)
    m, n = query.shape[0], ref.shape[0]
    out = np.empty((m, 3), np.float32); mask = np.empty((m,), np.uint8)
    lib().oracle_blend_anchor_motion(_p(query, _f32p), _p(ref, _f32p), _p(flow, _f32p), ctypes.c_int64(m),
                                     ctypes.c_int64(n), ctypes.c_int(k), _p(out, _f32p), _p(mask, _u8p))
    return out, mask.astype(bool)


def fps(xyz, m):
    """networks/pointnet_lib/src/sampling_gpu.cu:93-209 -> idx [B,m] int32 (starts at 0)."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    out = np.empty((B, m), np.int32)
    lib().oracle_fps(_p(xyz, _f32p), ctypes.c_int64(B), ctypes.c_int64(N), ctypes.c_int64(m), _p(out, _i32p))
    return out
