"""Exact tile culling of the search (csrc/cull.cu, chamfer_sym.cu CULL): the culled evaluation must be BIT-IDENTICAL to
the brute-force one -- loss, per-point distances, arg-min indices, gradients -- for any point order and any seeds, and it
must actually skip work when the clouds are in k-d leaf order and the seeds come from a nearby previous evaluation."""
import numpy as np
import pytest
import torch

from conftest import synthetic_sequence

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda")


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev())
    return t if dtype is None else t.to(dtype)


def evaluate(cano, W, R, tr, frames, nn=None):
    """One fused evaluation through the C ABI; nn = (nn_rows, nn_cols, stats) switches the culled schedule on."""
    from reart_b200 import _lib, ops
    L = _lib.lib()
    T, P = R.shape[0], R.shape[1]
    N, M = cano.shape[0], frames.shape[1]
    out = dict(skinned=torch.empty(T, N, 3, device=dev()), loss=torch.empty(1, dtype=torch.float64, device=dev()),
               gW=torch.empty(N, P, device=dev()), gpose=torch.empty(T * P * 12, device=dev()),
               gs=torch.empty(T, N, 3, device=dev()), d_f=torch.empty(T, N, device=dev()),
               i_f=torch.empty(T, N, dtype=torch.int64, device=dev()), d_b=torch.empty(T, M, device=dev()),
               i_b=torch.empty(T, M, dtype=torch.int64, device=dev()))
    nbytes = L.reart_energy_workspace_bytes(T, N, M)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev())
    packed = ops.pack_cloud(frames)
    p = _lib.ptr
    gR, gtr = out["gpose"][:T * P * 9], out["gpose"][T * P * 9:]
    args = [p(cano), p(W), p(R), p(tr), p(frames), p(packed), T, N, M, P, p(out["skinned"]), p(out["loss"]), p(out["gW"]), p(gR), p(gtr),
            p(out["gs"]), 1, p(out["d_f"]), p(out["i_f"]), p(out["d_b"]), p(out["i_b"])]
    if nn is None:
        _lib.check(L.reart_skinned_chamfer_fwd_bwd_ex(*args, p(ws), nbytes, _lib.stream_ptr()), "ex")
    else:
        _lib.check(L.reart_skinned_chamfer_fwd_bwd_culled(*args, p(nn[0]), p(nn[1]), p(nn[2]), p(ws), nbytes, _lib.stream_ptr()), "culled")
    torch.cuda.synchronize()
    return out


def same(a, b):
    for k in ("skinned", "loss", "gW", "gpose", "gs", "d_f", "i_f", "d_b", "i_b"):
        assert torch.equal(a[k], b[k]), k


def scene(T, N, P, seed, ordered):
    from reart_b200 import ops
    seq = synthetic_sequence(T, N, P, seed=seed)
    cano, frames = cu(seq["cano"]), cu(seq["frames"])
    part = cu(seq["part"].astype(np.int64))
    if ordered:
        pc = ops.kd_order(cano[None], 256)[0]
        cano, part = cano[pc].contiguous(), part[pc]
        pf = ops.kd_order(frames, 32)
        frames = torch.gather(frames, 1, pf[:, :, None].expand(-1, -1, 3)).contiguous()
    W = torch.eye(P, device=dev())[part].contiguous()
    R = cu(np.ascontiguousarray(seq["pose"][:, :, :3, :3]))
    tr = cu(np.ascontiguousarray(seq["pose"][:, :, :3, 3]))
    return cano, W, R, tr, frames


@pytest.mark.parametrize("T,N,P,ordered", [(3, 4096, 5, True), (2, 5000, 4, True), (3, 4096, 5, False), (2, 300, 3, True),
                                           (1, 16384, 8, True)])
def test_culled_evaluation_is_bit_identical_and_skips_work(T, N, P, ordered):
    cano, W, R, tr, frames = scene(T, N, P, 11, ordered)
    M = frames.shape[1]
    nn = (torch.full((T, N), -1, dtype=torch.int32, device=dev()), torch.full((T, M), -1, dtype=torch.int32, device=dev()),
          torch.zeros(2, dtype=torch.int64, device=dev()))
    brute = evaluate(cano, W, R, tr, frames)
    first = evaluate(cano, W, R, tr, frames, nn)                       # no seeds yet
    same(brute, first)
    ev, off = nn[2].tolist()
    # small clouds: infinite bounds, nothing skipped; from 8192 points on the history-free coarse bounds of cull.cu already
    # cull on the very first evaluation
    assert off > 0 and (ev == off if min(N, M) < 8192 else ev < off)
    assert torch.equal(nn[0].long(), brute["i_f"]) and torch.equal(nn[1].long(), brute["i_b"])   # the seeds ARE the arg-mins
    # the optimiser moves the poses a little; seeds from the previous evaluation
    g = torch.Generator(device="cuda").manual_seed(3)
    tr2 = tr + 0.004 * torch.randn(tr.shape, device=dev(), generator=g)
    nn[2].zero_()
    brute2 = evaluate(cano, W, R, tr2, frames)
    culled2 = evaluate(cano, W, R, tr2, frames, nn)
    same(brute2, culled2)
    ev, off = nn[2].tolist()
    if ordered and N >= 4096:
        assert ev < 0.6 * off, (ev, off)                                 # compact leaves + good seeds: most blocks skipped
    # garbage seeds (any valid index is an admissible upper bound) and out-of-range seeds: still exact
    nn[0].random_(0, M); nn[1].random_(0, N)
    same(brute2, evaluate(cano, W, R, tr2, frames, nn))
    nn[0].fill_(M + 5); nn[1].fill_(-7)
    same(brute2, evaluate(cano, W, R, tr2, frames, nn))


def test_culled_evaluation_keeps_lowest_index_ties_on_lattices():
    """Integer lattices: many exactly equal distances and duplicated points; the strict '>' of the skip test must keep
    every tie alive so that the lowest index still wins."""
    rng = np.random.default_rng(5)
    T, N, P = 2, 2048, 2
    cano = cu(rng.integers(-4, 5, (N, 3)).astype(np.float32))
    frames = cu(rng.integers(-4, 5, (T, N, 3)).astype(np.float32))
    W = torch.zeros(N, P, device=dev()); W[:, 0] = 1
    R = torch.eye(3, device=dev()).repeat(T, P, 1, 1).contiguous()
    tr = torch.zeros(T, P, 3, device=dev())
    nn = (torch.full((T, N), -1, dtype=torch.int32, device=dev()), torch.full((T, N), -1, dtype=torch.int32, device=dev()),
          torch.zeros(2, dtype=torch.int64, device=dev()))
    brute = evaluate(cano, W, R, tr, frames)
    same(brute, evaluate(cano, W, R, tr, frames, nn))
    same(brute, evaluate(cano, W, R, tr, frames, nn))                   # seeded by exact previous arg-mins: bounds are tight (= minima)


def test_engine_with_culling_follows_the_brute_force_optimisation_bit_for_bit():
    from reart_b200.engine import RelaxationEngine, tau_schedule
    seq = synthetic_sequence(6, 4096, 6, seed=4)
    cano, frames = cu(seq["cano"]), cu(seq["frames"])
    runs = {}
    for cull in (True, False):
        eng = RelaxationEngine(cano, frames, num_parts=6, use_graph=cull, seed=2, cull=cull)
        torch.manual_seed(9)
        runs[cull] = [float(eng.step(tau_schedule(i, 100, 5.0, 1.0))) for i in range(25)]
        if cull:
            ev, off = eng.culling_stats()
            assert ev < off
            # engine order -> caller order
            sk = torch.empty_like(eng.skinned); sk[:, eng.perm_cano] = eng.skinned
            assert sk.shape == (6, 4096, 3)
        eng.release()
    # k-d ordering permutes the points (and so the order of every float sum): same optimisation, not the same bits
    np.testing.assert_allclose(runs[True][:5], runs[False][:5], rtol=1e-5)
    np.testing.assert_allclose(runs[True], runs[False], rtol=5e-3)


def test_engine_with_culling_at_coarse_bound_sizes_and_with_the_assignment_loss():
    """8192 points per cloud: the history-free coarse bounds of cull.cu are active inside the engine (they cull from the
    first step on); and culling together with the assignment loss added to Chamfer (run_real.py form) under the graph."""
    from reart_b200.engine import RelaxationEngine, tau_schedule
    seq = synthetic_sequence(4, 8192, 6, seed=8)
    cano, frames = cu(seq["cano"]), cu(seq["frames"])
    for assign in (None, dict(downsample=8, assign_gap=3, lambda_assign=0.3, assign_iter=2, mode="add")):
        runs = {}
        for cull in (True, False):
            eng = RelaxationEngine(cano, frames, num_parts=6, use_graph=cull, seed=2, cull=cull, assign=assign)
            torch.manual_seed(9)
            runs[cull] = [float(eng.step(tau_schedule(i, 100, 5.0, 1.0))) for i in range(10)]
            if cull:
                ev, off = eng.culling_stats()
                assert 0 < ev < off
            eng.release()
        np.testing.assert_allclose(runs[True][:4], runs[False][:4], rtol=2e-5)
        np.testing.assert_allclose(runs[True], runs[False], rtol=5e-3)
