"""torch.autograd bindings of the C-ABI kernels (one Function per differentiable op).

Every Function passes raw device pointers on the current stream to ``libreart_b200.so`` and raises
``ReartError`` on CPU tensors -- there is no fallback path.
"""
from __future__ import annotations

import ctypes

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import check, ptr, stream_ptr


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# --------------------------------------------------------------------------------------- packed clouds
def pack_cloud(pts: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """[B,P,3] -> packed float buffer for the searches (include/reart_b200.h: reart_pack_cloud); ``out``: write into an
    existing packed buffer of the same shape (e.g. the engine's, when new frames arrive)."""
    _lib.require_cuda(pts)
    L = _lib.lib()
    pts = _f32c(pts)
    B, P, _ = pts.shape
    n = L.reart_packed_bytes(B, P) // 4
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=pts.device)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == n and out.device == pts.device
    with torch.cuda.device(pts.device):
        check(L.reart_pack_cloud(ptr(pts), B, P, ptr(out), stream_ptr()), "reart_pack_cloud")
    return out


# --------------------------------------------------------------------------------------- skinning
class _Skin(Function):
    @staticmethod
    def forward(ctx, cano, W, R, tr):
        _lib.require_cuda(cano, W, R, tr)
        L = _lib.lib()
        cano_c, W_c, R_c, tr_c = _f32c(cano), _f32c(W), _f32c(R), _f32c(tr)
        T, P = R_c.shape[0], R_c.shape[1]
        N = cano_c.shape[0]
        out = torch.empty(T, N, 3, dtype=torch.float32, device=cano.device)
        with torch.cuda.device(cano.device):
            check(L.reart_skin_fwd(ptr(cano_c), ptr(W_c), ptr(R_c), ptr(tr_c), T, N, P, ptr(out), stream_ptr()),
                  "reart_skin_fwd")
        ctx.save_for_backward(cano_c, W_c, R_c, tr_c)
        ctx.w_dtype = W.dtype
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        cano, W, R, tr = ctx.saved_tensors
        L = _lib.lib()
        T, P = R.shape[0], R.shape[1]
        N = cano.shape[0]
        g = _f32c(g)
        gW = torch.empty_like(W)
        gR = torch.empty_like(R)
        gtr = torch.empty_like(tr)
        nbytes = L.reart_skin_bwd_workspace_bytes(T, N, P)
        ws = _lib.workspace(nbytes, cano.device)
        with torch.cuda.device(cano.device):
            check(L.reart_skin_bwd(ptr(cano), ptr(W), ptr(R), ptr(tr), ptr(g), T, N, P, ptr(gW), ptr(gR), ptr(gtr),
                                   ptr(ws), nbytes, stream_ptr()), "reart_skin_bwd")
        gW_out = gW if ctx.needs_input_grad[1] and ctx.w_dtype.is_floating_point else None
        return None, gW_out, gR, gtr


def skin(cano: torch.Tensor, W: torch.Tensor, R: torch.Tensor, tr: torch.Tensor) -> torch.Tensor:
    """out[t,n] = sum_p W[n,p] (R[t,p] cano[n] + tr[t,p])   (networks/model.py:63-69).

    cano [N,3] (no gradient: it is data), W [N,P] (float or the int64 one_hot of SURVEY Q7),
    R [T,P,3,3], tr [T,P,3] -> [T,N,3].
    """
    return _Skin.apply(cano, W, R, tr)


# --------------------------------------------------------------------------------------- seg MLP + gumbel ST
class _SegMlp(Function):
    @staticmethod
    def forward(ctx, x, w0, b0, w2, sink):
        _lib.require_cuda(x, w0, b0, w2)
        L = _lib.lib()
        xc, w0c, b0c, w2c = _f32c(x), _f32c(w0), _f32c(b0), _f32c(w2)
        N, H, P = xc.shape[0], w0c.shape[0], w2c.shape[0]
        logits = torch.empty(N, P, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(L.reart_segmlp_fwd(ptr(xc), ptr(w0c), ptr(b0c), ptr(w2c), N, H, P, ptr(logits), stream_ptr()),
                  "reart_segmlp_fwd")
        ctx.save_for_backward(xc, w0c, b0c, w2c)
        ctx.sink = sink
        return logits

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w0, b0, w2 = ctx.saved_tensors
        L = _lib.lib()
        N, H, P = x.shape[0], w0.shape[0], w2.shape[0]
        g = _f32c(g)
        sink = ctx.sink
        if sink is not None and not sink.active:
            sink = None                                        # outside a synchronized engine step: plain gradients, no collective
        if sink is None:
            flat = torch.empty(4 * H + P * H, dtype=torch.float32, device=x.device)
            gw0, gb0, gw2 = flat[:3 * H].view(H, 3), flat[3 * H:4 * H], flat[4 * H:].view(P, H)
        else:
            # gradients land directly in the caller's flat bucket (multi-GPU: reduced in place, no pack/unpack copies)
            flat = sink.flat
            gw0 = flat[:3 * H].view(H, 3)
            gb0 = flat[3 * H:4 * H]
            gw2 = flat[4 * H:4 * H + P * H].view(P, H)
        with torch.cuda.device(x.device):
            check(L.reart_segmlp_bwd(ptr(x), ptr(w0), ptr(b0), ptr(w2), ptr(g), N, H, P, ptr(gw0), ptr(gb0), ptr(gw2),
                                     stream_ptr()), "reart_segmlp_bwd")
        if sink is not None:
            sink.reduce()
        return None, gw0, gb0, gw2, None


class GradSink:
    """Flat gradient buffer of the seg MLP [w0 | b0 | w2 | extra scalars] plus the reduction to run on it once the
    backward kernel has filled it (an NCCL all-reduce under frame sharding)."""

    def __init__(self, H: int, P: int, device, extra_scalars: int = 1, reducer=None):
        self.n_grad = 4 * H + P * H
        self.flat = torch.zeros(self.n_grad + extra_scalars, dtype=torch.float32, device=device)
        self.extra = self.flat[self.n_grad:]
        self.reducer = reducer
        # The bucket is written in place and reduced (a collective!) from inside the backward, and the returned gradients
        # alias it.  That is only sound inside the engine's synchronized step (all ranks run the same backward, grads are
        # consumed by the optimiser before the next one); any other backward through the seg logits -- ik() on a BaseModel,
        # an evaluation on one rank -- gets fresh gradient tensors and no collective.
        self.active = False

    def reduce(self):
        if self.reducer is not None:
            self.reducer(self.flat)


def seg_mlp(x, w0, b0, w2, sink: "GradSink | None" = None):
    """logits = W2 relu(W0 x + b0): x [N,3] (no gradient), w0 [H,3], b0 [H], w2 [P,H] -> [N,P]."""
    return _SegMlp.apply(x, w0, b0, w2, sink)


class _GumbelST(Function):
    @staticmethod
    def forward(ctx, logits, tau):
        _lib.require_cuda(logits, tau)
        L = _lib.lib()
        lg = _f32c(logits)
        N, P = lg.shape
        expo = torch.empty_like(lg).exponential_()            # the same RNG draw F.gumbel_softmax makes
        W = torch.empty_like(lg)
        ysoft = torch.empty_like(lg)
        with torch.cuda.device(lg.device):
            check(L.reart_gumbel_st_fwd(ptr(lg), ptr(expo), ptr(tau), N, P, ptr(W), ptr(ysoft), stream_ptr()),
                  "reart_gumbel_st_fwd")
        ctx.save_for_backward(ysoft, tau)
        return W

    @staticmethod
    @once_differentiable
    def backward(ctx, gW):
        ysoft, tau = ctx.saved_tensors
        L = _lib.lib()
        N, P = ysoft.shape
        g = _f32c(gW)
        out = torch.empty_like(ysoft)
        with torch.cuda.device(ysoft.device):
            check(L.reart_gumbel_st_bwd(ptr(ysoft), ptr(tau), ptr(g), N, P, ptr(out), stream_ptr()),
                  "reart_gumbel_st_bwd")
        return out, None


def gumbel_softmax_st(logits: torch.Tensor, tau) -> torch.Tensor:
    """F.gumbel_softmax(logits, tau=tau, hard=True) fused (networks/model.py:44); tau: float or 0-dim CUDA tensor."""
    if not torch.is_tensor(tau):
        tau = torch.full((1,), float(tau), dtype=torch.float32, device=logits.device)
    else:
        tau = tau.detach().to(device=logits.device, dtype=torch.float32).reshape(1)
    return _GumbelST.apply(logits, tau)


# --------------------------------------------------------------------------------------- 6D -> R
class _Rot6d(Function):
    @staticmethod
    def forward(ctx, d6):
        _lib.require_cuda(d6)
        L = _lib.lib()
        flat = _f32c(d6).reshape(-1, 6)
        B = flat.shape[0]
        R = torch.empty(B, 3, 3, dtype=torch.float32, device=d6.device)
        with torch.cuda.device(d6.device):
            check(L.reart_rot6d_fwd(ptr(flat), B, ptr(R), stream_ptr()), "reart_rot6d_fwd")
        ctx.save_for_backward(flat)
        ctx.in_shape = d6.shape
        return R.reshape(d6.shape[:-1] + (3, 3))

    @staticmethod
    @once_differentiable
    def backward(ctx, gR):
        (flat,) = ctx.saved_tensors
        L = _lib.lib()
        B = flat.shape[0]
        g = _f32c(gR).reshape(B, 9)
        gd6 = torch.empty_like(flat)
        with torch.cuda.device(flat.device):
            check(L.reart_rot6d_bwd(ptr(flat), ptr(g), B, ptr(gd6), stream_ptr()), "reart_rot6d_bwd")
        return gd6.reshape(ctx.in_shape)


def rot6d(d6: torch.Tensor) -> torch.Tensor:
    return _Rot6d.apply(d6)


# --------------------------------------------------------------------------------------- screw -> 4x4
class _ScrewToTransform(Function):
    @staticmethod
    def forward(ctx, l, m, theta, d):
        _lib.require_cuda(l, m, theta, d)
        L = _lib.lib()
        l_c, m_c, th_c, d_c = _f32c(l), _f32c(m), _f32c(theta), _f32c(d)
        B = l_c.shape[0]
        M = torch.empty(B, 4, 4, dtype=torch.float32, device=l.device)
        with torch.cuda.device(l.device):
            check(L.reart_screw_to_transform_fwd(ptr(l_c), ptr(m_c), ptr(th_c), ptr(d_c), B, ptr(M), stream_ptr()),
                  "reart_screw_to_transform_fwd")
        ctx.save_for_backward(l_c, m_c, th_c, d_c)
        return M

    @staticmethod
    @once_differentiable
    def backward(ctx, gM):
        l, m, th, d = ctx.saved_tensors
        L = _lib.lib()
        B = l.shape[0]
        g = _f32c(gM)
        gl, gm, gth, gd = torch.empty_like(l), torch.empty_like(m), torch.empty_like(th), torch.empty_like(d)
        with torch.cuda.device(l.device):
            check(L.reart_screw_to_transform_bwd(ptr(l), ptr(m), ptr(th), ptr(d), ptr(g), B, ptr(gl), ptr(gm), ptr(gth),
                                                 ptr(gd), stream_ptr()), "reart_screw_to_transform_bwd")
        return gl, gm, gth, gd


def screw_to_transform(l, m, theta, d) -> torch.Tensor:
    """transform_from_exponential_coordinates(screw_param_to_exponential_coordinates(l, m, theta, d)) fused."""
    return _ScrewToTransform.apply(l, m, theta, d)


# --------------------------------------------------------------------------------------- forward kinematics
class _Fk(Function):
    @staticmethod
    def forward(ctx, axis, moment, theta, distance, order, parent, edge, joint_type):
        _lib.require_cuda(axis, moment, theta)
        L = _lib.lib()
        a, mo, th = _f32c(axis), _f32c(moment), _f32c(theta)
        di = _f32c(distance) if distance is not None else None
        T = th.shape[0]
        P = order.shape[0]
        out = torch.empty(T, P, 4, 4, dtype=torch.float32, device=axis.device)
        with torch.cuda.device(axis.device):
            check(L.reart_fk_fwd(ptr(a), ptr(mo), ptr(th), ptr(di), ptr(order), ptr(parent), ptr(edge), ptr(joint_type),
                                 T, P, ptr(out), stream_ptr()), "reart_fk_fwd")
        ctx.save_for_backward(a, mo, th, out, order, parent, edge)
        ctx.di = di
        ctx.jt = joint_type
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g_out):
        a, mo, th, out, order, parent, edge = ctx.saved_tensors
        di, jt = ctx.di, ctx.jt
        L = _lib.lib()
        T, P = out.shape[0], out.shape[1]
        g = _f32c(g_out)
        ga, gm, gth = torch.zeros_like(a), torch.zeros_like(mo), torch.zeros_like(th)
        gdi = torch.zeros_like(di) if di is not None else None
        ws = torch.empty(T * P * 16, dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device):
            check(L.reart_fk_bwd(ptr(a), ptr(mo), ptr(th), ptr(di), ptr(order), ptr(parent), ptr(edge), ptr(jt), T, P,
                                 ptr(out), ptr(g), ptr(ga), ptr(gm), ptr(gth), ptr(gdi), ptr(ws), stream_ptr()),
                  "reart_fk_bwd")
        return ga, gm, gth, gdi, None, None, None, None


def fk_flat(axis, moment, theta, distance, order, parent, edge, joint_type=None) -> torch.Tensor:
    """Forward kinematics on the flattened tree (device int32 arrays) -> [T,P,4,4]."""
    return _Fk.apply(axis, moment, theta, distance, order, parent, edge, joint_type)


# --------------------------------------------------------------------------------------- k-NN helpers (no grad)
@torch.no_grad()
def knn(ref: torch.Tensor, query: torch.Tensor, k: int, squared: bool = False):
    """ref [B,n,3], query [B,m,3] -> (dist [B,m,k] Euclidean -- or squared -- , idx [B,m,k] int64), ascending."""
    _lib.require_cuda(ref, query)
    L = _lib.lib()
    ref, query = _f32c(ref), _f32c(query)
    B, n, _ = ref.shape
    m = query.shape[1]
    dist = torch.empty(B, m, k, dtype=torch.float32, device=ref.device)
    idx = torch.empty(B, m, k, dtype=torch.int64, device=ref.device)
    fn = L.reart_knn_sq if squared else L.reart_knn
    with torch.cuda.device(ref.device):
        check(fn(ptr(ref), ptr(query), B, n, m, int(k), ptr(dist), ptr(idx), stream_ptr()), "reart_knn")
    return dist, idx


@torch.no_grad()
def flow_refs_sort(ref_cat: torch.Tensor, ref_offsets: torch.Tensor, max_refs: int):
    """Sort every pair's reference set along x once (include/reart_b200.h: reart_flow_refs_sort) ->
    (sorted uint8 buffer, sorted_offsets [T] int64) for ``knn3_blend(..., sorted_refs=...)``; None when a set is too
    large for the windowed kernel (max_refs > 16384)."""
    _lib.require_cuda(ref_cat, ref_offsets)
    L = _lib.lib()
    ref_cat = _f32c(ref_cat)
    ref_offsets = ref_offsets.to(torch.int64).contiguous()
    T = ref_offsets.numel() - 1
    if T <= 0 or max_refs > 16384:
        return None
    total = int(ref_cat.shape[0])
    nbytes = L.reart_flow_refs_sorted_bytes(total, T)
    buf = torch.empty(nbytes, dtype=torch.uint8, device=ref_cat.device)
    soff = torch.empty(T, dtype=torch.int64, device=ref_cat.device)
    with torch.cuda.device(ref_cat.device):
        check(L.reart_flow_refs_sort(ptr(ref_cat), ptr(ref_offsets), T, int(max_refs), ptr(buf), nbytes, total, ptr(soff),
                                     stream_ptr()), "reart_flow_refs_sort")
    return buf, soff


def knn3_blend(query: torch.Tensor, ref_cat: torch.Tensor, flow_cat: torch.Tensor, ref_offsets: torch.Tensor,
               sorted_refs=None):
    """All frame pairs of the flow loss at once -> (blended [T,m,3], mask [T,m] bool).  With ``sorted_refs`` (from
    ``flow_refs_sort`` of the same reference sets) the windowed exact kernel runs: same bits, a fraction of the pairs."""
    _lib.require_cuda(query, ref_cat, flow_cat, ref_offsets)
    L = _lib.lib()
    query, ref_cat, flow_cat = _f32c(query), _f32c(ref_cat), _f32c(flow_cat)
    ref_offsets = ref_offsets.to(torch.int64).contiguous()
    T, m, _ = query.shape
    blended = torch.empty(T, m, 3, dtype=torch.float32, device=query.device)
    mask = torch.empty(T, m, dtype=torch.uint8, device=query.device)
    with torch.cuda.device(query.device):
        if sorted_refs is not None:
            buf, soff = sorted_refs
            order = torch.empty(T, m, dtype=torch.int32, device=query.device)
            check(L.reart_knn3_blend_sorted(ptr(query), ptr(buf), ptr(soff), ptr(flow_cat), ptr(ref_offsets), T, m, ptr(order),
                                            ptr(blended), ptr(mask), stream_ptr()), "reart_knn3_blend_sorted")
        else:
            check(L.reart_knn3_blend(ptr(query), ptr(ref_cat), ptr(flow_cat), ptr(ref_offsets), T, m, ptr(blended),
                                     ptr(mask), stream_ptr()), "reart_knn3_blend")
    return blended, mask.bool()


# --------------------------------------------------------------------------------------- FPS / ball query (no grad)
@torch.no_grad()
def fps_into(xyz: torch.Tensor, npoint: int, out: torch.Tensor, temp: "torch.Tensor | None" = None) -> torch.Tensor:
    """xyz [B,N,3] -> out [B,npoint] int32 (written in place), starting at index 0.  Clouds above 32 768 points keep
    their running distances in ``temp`` [B,N] float32 (allocated here when not given), as the reference's kernel does."""
    _lib.require_cuda(xyz, out)
    L = _lib.lib()
    xyz = _f32c(xyz)
    B, N, _ = xyz.shape
    assert out.dtype == torch.int32 and out.is_contiguous() and out.shape == (B, npoint)
    if N > 32768 and (temp is None or temp.dtype != torch.float32 or not temp.is_contiguous() or temp.numel() < B * N):
        temp = torch.empty(B, N, dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        check(L.reart_fps_temp(ptr(xyz), B, N, int(npoint), ptr(temp) if N > 32768 else None, ptr(out), stream_ptr()),
              "reart_fps_temp")
    return out


def fps(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """furthest_point_sample(xyz, npoint).long() of networks/pointnet2_utils.py:85 -> [B,npoint] int64."""
    out = torch.empty(xyz.shape[0], npoint, dtype=torch.int32, device=xyz.device)
    return fps_into(xyz, npoint, out).long()


@torch.no_grad()
def ball_query_into(new_xyz, xyz, radius: float, nsample: int, idx: torch.Tensor) -> torch.Tensor:
    _lib.require_cuda(new_xyz, xyz, idx)
    L = _lib.lib()
    new_xyz, xyz = _f32c(new_xyz), _f32c(xyz)
    B, N, _ = xyz.shape
    m = new_xyz.shape[1]
    assert idx.dtype == torch.int32 and idx.is_contiguous()
    with torch.cuda.device(xyz.device):
        check(L.reart_ball_query(ptr(new_xyz), ptr(xyz), B, N, m, float(radius), int(nsample), ptr(idx), stream_ptr()),
              "reart_ball_query")
    return idx


# --------------------------------------------------------------------------------------- fused energy
class _SkinnedChamfer(Function):
    """skin -> bidirectional Chamfer -> sum, with the backward computed in the same C call."""

    @staticmethod
    def forward(ctx, cano, W, R, tr, tgt, tgt_packed, unit_grad):
        _lib.require_cuda(cano, W, R, tr, tgt, tgt_packed)
        ctx.unit_grad = bool(unit_grad)
        L = _lib.lib()
        cano_c, W_c, R_c, tr_c, tgt_c = _f32c(cano), _f32c(W), _f32c(R), _f32c(tr), _f32c(tgt)
        T, P = R_c.shape[0], R_c.shape[1]
        N, M = cano_c.shape[0], tgt_c.shape[1]
        dev = cano.device
        skinned = torch.empty(T, N, 3, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float64, device=dev)
        need = any(ctx.needs_input_grad[1:4])
        gW = torch.empty_like(W_c) if need else None
        gR = gtr = None
        if need:                                               # one flat [gR | gtr] buffer => one memset in the C call
            flat = torch.empty(T * P * 12, dtype=torch.float32, device=dev)
            gR, gtr = flat[:T * P * 9].view(T, P, 3, 3), flat[T * P * 9:].view(T, P, 3)
        nbytes = L.reart_energy_workspace_bytes(T, N, M)
        ws = _lib.workspace(nbytes, dev)
        with torch.cuda.device(dev):
            check(L.reart_skinned_chamfer_fwd_bwd(ptr(cano_c), ptr(W_c), ptr(R_c), ptr(tr_c), ptr(tgt_c), ptr(tgt_packed),
                                                  T, N, M, P, ptr(skinned), ptr(loss), ptr(gW), ptr(gR), ptr(gtr), None,
                                                  1 if need else 0, ptr(ws), nbytes, stream_ptr()),
                  "reart_skinned_chamfer_fwd_bwd")
        ctx.grads = (gW, gR, gtr)
        ctx.w_float = W.dtype.is_floating_point
        ctx.save_for_backward(cano_c, W_c, R_c, tr_c)
        ctx.set_materialize_grads(False)                      # g_skinned stays None when nothing else consumes it
        return loss.to(torch.float32)[0], skinned

    @staticmethod
    @once_differentiable
    def backward(ctx, g_loss, g_skinned):
        gW, gR, gtr = ctx.grads
        if gW is None:
            return None, None, None, None, None, None, None
        if g_loss is None:                                     # the loss itself is not part of the objective
            gW, gR, gtr = torch.zeros_like(gW), torch.zeros_like(gR), torch.zeros_like(gtr)
        elif ctx.unit_grad:
            pass                                               # caller promised d(objective)/d(loss) == 1: no scaling launch
        else:
            gW, gR, gtr = torch._foreach_mul([gW, gR, gtr], g_loss)          # one launch for the three
        if g_skinned is not None:
            # other consumers of the skinned cloud (flow / assign losses): one more skin backward, added by linearity
            cano, W, R, tr = ctx.saved_tensors
            L = _lib.lib()
            T, P = R.shape[0], R.shape[1]
            N = cano.shape[0]
            g = _f32c(g_skinned)
            eW, eR, et = torch.empty_like(W), torch.empty_like(R), torch.empty_like(tr)
            nbytes = L.reart_skin_bwd_workspace_bytes(T, N, P)
            ws = _lib.workspace(nbytes, cano.device)
            with torch.cuda.device(cano.device):
                check(L.reart_skin_bwd(ptr(cano), ptr(W), ptr(R), ptr(tr), ptr(g), T, N, P, ptr(eW), ptr(eR), ptr(et),
                                       ptr(ws), nbytes, stream_ptr()), "reart_skin_bwd")
            gW, gR, gtr = gW + eW, gR + eR, gtr + et
        return (None, gW if ctx.w_float else None, gR, gtr, None, None, None)


def skinned_chamfer_loss(cano, W, R, tr, tgt, tgt_packed=None, unit_grad: bool = False):
    """Fused recon_loss(model skin(cano), pc_list): returns (loss scalar, skinned [T,N,3]).  `skinned` stays
    differentiable for further consumers (flow / assign losses); their gradient costs one extra skin backward.
    unit_grad=True promises that the loss enters the objective with weight exactly 1 (the run scripts' case) and
    skips the gradient-scaling launch."""
    if tgt_packed is None:
        tgt_packed = pack_cloud(tgt)
    return _SkinnedChamfer.apply(cano, W, R, tr, tgt, tgt_packed, unit_grad)


# --------------------------------------------------------------------------------------- k-d leaf ordering (culling)
@torch.no_grad()
def kd_order(points: torch.Tensor, leaf: int) -> torch.Tensor:
    """Permutation [B,N] that orders every cloud of ``points`` [B,N,3] into k-d leaves of ``leaf`` consecutive points
    (recursive median split along the longest axis of each segment), so that consecutive index ranges are spatially
    compact.  Pure torch, one batched sort per level; meant to run ONCE per cloud at set-up.  The exact culled search
    (``reart_skinned_chamfer_fwd_bwd_culled``) is correct for any order and fast for this one."""
    B, N, _ = points.shape
    levels = 0
    while (leaf << levels) < N:
        levels += 1
    n_pad = leaf << levels
    pts = points.float()
    if n_pad != N:                                             # far-away padding: sorts last, is dropped at the end
        pts = torch.cat((pts, torch.full((B, n_pad - N, 3), 1e30, dtype=pts.dtype, device=pts.device)), dim=1)
    idx = torch.arange(n_pad, device=pts.device)[None].expand(B, n_pad).contiguous()
    for level in range(levels):
        segs, seg = 1 << level, n_pad >> level
        P = torch.gather(pts, 1, idx[:, :, None].expand(-1, -1, 3)).view(B, segs, seg, 3)
        axis = (P.amax(dim=2) - P.amin(dim=2)).argmax(dim=-1)                                  # [B, segs]
        key = torch.gather(P, 3, axis[:, :, None, None].expand(-1, -1, seg, 1))[..., 0]         # [B, segs, seg]
        order = key.argsort(dim=-1, stable=True)
        idx = torch.gather(idx.view(B, segs, seg), 2, order).view(B, n_pad)
    if n_pad != N:
        keep = idx < N
        idx = idx[keep].view(B, N)
    return idx


def fp32_probe(variant: int, iters: int = 2000, blocks: int | None = None, device=None):
    """Run one FP32 pipe micro-benchmark; returns (ms, lane_ops_total)."""
    L = _lib.lib()
    device = torch.device(device or "cuda")
    if blocks is None:
        blocks = torch.cuda.get_device_properties(device).multi_processor_count * 8
    sin = torch.rand(1024, device=device) + 0.5
    sout = torch.empty(blocks * 256, device=device)
    ms = ctypes.c_double()
    ops = ctypes.c_double()
    with torch.cuda.device(device):
        check(L.reart_fp32_probe(int(variant), int(iters), int(blocks), ptr(sin), ptr(sout), ctypes.byref(ms),
                                 ctypes.byref(ops), stream_ptr()), "reart_fp32_probe")
    return ms.value, ops.value * blocks * 256
