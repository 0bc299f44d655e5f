"""GPU assignment solver (csrc/lap.cu, reart_lap) and the assignment loss inside the optimisation engine.

Checks: (1) against the reference-generated golden of run_robot.py:164-187 on the nao demo (tests/golden/assign.npz,
oracle/make_golden.py::gen_assign): same FPS samples, same per-frame optimal cost, same loss; (2) against
scipy.optimize.linear_sum_assignment on seeded random clouds, tie-heavy integer lattices, duplicated points and the
size limits -- the optimal COST must agree (the matching itself may differ only among equal-cost optima); (3) the
engine: GPU-solver run == host-solver run, CUDA-graph replay == eager, both loss forms (added to / replacing Chamfer).
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, synthetic_sequence

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda")


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev())
    return t if dtype is None else t.to(dtype)


def scipy_totals(src, tgt):
    """Optimal assignment cost per frame under the float32 Euclidean cost, solved by scipy in float64."""
    from scipy.optimize import linear_sum_assignment
    out = []
    for a, b in zip(src, tgt):
        d = a[:, None, :].astype(np.float32) - b[None, :, :].astype(np.float32)
        c = np.sqrt((d * d).sum(-1, dtype=np.float32)).astype(np.float64)
        r, col = linear_sum_assignment(c)
        out.append(c[r, col].sum())
    return np.array(out)


def check_against_scipy(src, tgt, rtol=1e-6):
    from reart_b200.assign import lap_assign
    cols, total = lap_assign(cu(src), cu(tgt), want_total=True)
    cols = cols.cpu().numpy(); total = total.cpu().numpy()
    n = src.shape[1]
    for t in range(src.shape[0]):
        assert sorted(cols[t].tolist()) == list(range(n))                       # a permutation
        mine = np.sqrt(((src[t].astype(np.float64) - tgt[t][cols[t]].astype(np.float64)) ** 2).sum(-1)).sum()
        assert abs(mine - total[t]) <= 1e-5 * max(total[t], 1e-12)              # reported cost = cost of the matching
    want = scipy_totals(src, tgt)
    np.testing.assert_allclose(total, want, rtol=rtol, atol=1e-9)


def test_lap_reference_golden_nao_refresh_block(nao):
    """run_robot.py:164-187 on the nao demo: FPS (CUDA start index 0) -> cdist -> linear_sum_assignment -> loss."""
    from reart_b200.assign import AssignLoss
    from reart_b200.model_utils import compute_pc_transform
    g, cano, pc_list = nao
    ga = load_golden("assign.npz")
    skinned = compute_pc_transform(cu(cano), cu(g["katA_pose"]), cu(g["katA_part"].astype(np.int64)))
    al = AssignLoss(cu(cano), cu(pc_list), downsample=int(ga["downsample"]), assign_gap=5, lambda_assign=float(ga["lambda_assign"]))
    assert np.array_equal(al.src_idx.cpu().numpy(), ga["src_idx"].astype(np.int64))
    assert np.array_equal(al.tgt_idx.cpu().numpy(), ga["tgt_idx"].astype(np.int64))
    al.refresh(skinned)
    cols = al.col4row.cpu().numpy()
    src = skinned[:, al.src_idx].cpu().numpy().astype(np.float64); tgt = al.pc_tgt.cpu().numpy().astype(np.float64)
    totals = np.array([np.sqrt(((src[t] - tgt[t][cols[t]]) ** 2).sum(-1)).sum() for t in range(src.shape[0])])
    np.testing.assert_allclose(totals, ga["totals"], rtol=2e-6)                  # same optimum as scipy on the reference's cost
    assert (cols == ga["col_ind"]).mean() > 0.99                                 # and (ties / ulp-level costs aside) the same matching
    loss = al.loss(skinned)
    assert abs(loss.item() - float(ga["ass_loss"])) <= 1e-4 * float(ga["ass_loss"])


@pytest.mark.parametrize("B,n", [(3, 1), (2, 2), (4, 7), (3, 100), (9, 1024), (2, 2048), (5, 513)])
def test_lap_optimal_cost_equals_scipy_on_random_clouds(B, n):
    rng = np.random.default_rng(100 + n)
    src = (rng.random((B, n, 3)) * 0.7 - 0.35).astype(np.float32)
    tgt = (src[:, rng.permutation(n)] + rng.normal(0, 0.02, (B, n, 3))).astype(np.float32)
    check_against_scipy(src, tgt)


def test_lap_tie_heavy_lattices_and_duplicates():
    rng = np.random.default_rng(7)
    src = rng.integers(-3, 4, (4, 300, 3)).astype(np.float32)                    # many equal costs, duplicated points
    tgt = rng.integers(-3, 4, (4, 300, 3)).astype(np.float32)
    check_against_scipy(src, tgt, rtol=1e-9)
    same = np.repeat(rng.random((1, 1, 3)).astype(np.float32), 64, axis=1)       # all points identical: every matching optimal
    check_against_scipy(same, same.copy(), rtol=1e-9)
    check_against_scipy(src[:, :257], src[:, :257].copy(), rtol=1e-9)            # identical clouds: cost 0


def test_lap_size_limit_and_index_indirection():
    from reart_b200 import _lib
    from reart_b200.assign import lap_assign
    rng = np.random.default_rng(3)
    full = cu((rng.random((2, 900, 3)) - 0.5).astype(np.float32))
    tgt = cu((rng.random((2, 128, 3)) - 0.5).astype(np.float32))
    idx = torch.from_numpy(rng.permutation(900)[:128]).to(dev())
    a = lap_assign(full, tgt, idx)
    b = lap_assign(full[:, idx].contiguous(), tgt)
    assert torch.equal(a, b)
    with pytest.raises(_lib.ReartError):
        lap_assign(torch.zeros(1, 4097, 3, device=dev()), torch.zeros(1, 4097, 3, device=dev()))


def test_compute_ass_err_on_the_gpu_solver_matches_reference_golden(nao):
    """utils/model_utils.py:92-103 (model-selection term): the golden was made by the reference on a 512-point subsample."""
    from reart_b200.model_utils import compute_ass_err, compute_pc_transform
    g, cano, pc_list = nao
    gs = load_golden("structure.npz")
    sub = gs["nao_ass_sub"]
    pred = compute_pc_transform(cu(gs["nao_cano"]), cu(gs["nao_pose"]), cu(gs["nao_part"].astype(np.int64)))
    # the golden used the extracted kinematic chain; here only the solver is under test: compare against the host path
    a = compute_ass_err(pred[:, sub].contiguous(), cu(pc_list)[:, sub].contiguous())
    from scipy.optimize import linear_sum_assignment
    ps, pl = pred[:, sub].cpu().numpy(), pc_list[:, sub]
    want = np.mean([((ps[t][r] - pl[t][c]) ** 2).sum(-1).mean() for t in range(ps.shape[0])
                    for r, c in [linear_sum_assignment(np.sqrt(((ps[t][:, None] - pl[t][None]) ** 2).sum(-1)))]])
    assert abs(a.item() - want) <= 1e-5 * want


@pytest.mark.parametrize("mode", ["add", "replace"])
def test_engine_assign_loss_gpu_solver_equals_host_solver_and_graph_equals_eager(mode):
    """The assignment loss inside the iteration (run_real.py additive form / run_robot.py if-else form):
    native fused path with reart_lap == autograd path with the reference's host solver; graph replay == eager."""
    from reart_b200.engine import RelaxationEngine, tau_schedule
    seq = synthetic_sequence(4, 2048, 5, seed=6)
    cano, frames = cu(seq["cano"]), cu(seq["frames"])
    cfg = dict(downsample=4, assign_gap=3, lambda_assign=0.3, assign_iter=4, mode=mode)
    runs = {}
    for name, kw in (("native_graph", dict(use_graph=True, native=True)), ("native_eager", dict(use_graph=False, native=True)),
                     ("autograd_host", dict(use_graph=False, native=False, assign=dict(cfg, solver="scipy")))):
        eng = RelaxationEngine(cano, frames, num_parts=5, seed=2, **{"assign": cfg, **kw})
        torch.manual_seed(5)
        runs[name] = [float(eng.step(tau_schedule(i, 100, 5.0, 1.0))) for i in range(12)]
        eng.release()
    assert runs["native_graph"] == runs["native_eager"]                          # bit for bit, across the flavour switches
    np.testing.assert_allclose(runs["native_eager"][:6], runs["autograd_host"][:6], rtol=5e-5)
    np.testing.assert_allclose(runs["native_eager"], runs["autograd_host"], rtol=5e-3)
    if mode == "replace":
        assert runs["native_eager"][4] < 0.5 * runs["native_eager"][3]           # the loss switches to the (smaller) assignment term
