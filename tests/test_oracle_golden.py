"""CPU: the C oracle (oracle/reart_oracle.c) against the golden vectors generated from the reference's own
Python (oracle/make_golden.py).  This is what pins the oracle (task section 3)."""
import numpy as np
import pytest

import oracle
from conftest import load_golden


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_knn1_and_backward_match_reference(tag):
    g = load_golden("chamfer_small.npz")
    s, t = g[f"{tag}_src"], g[f"{tag}_tgt"]
    d, i = oracle.knn1(s, t)
    assert np.array_equal(i, g[f"{tag}_i_fwd"])
    np.testing.assert_allclose(d, g[f"{tag}_d_fwd"], rtol=1e-5, atol=1e-9)
    d2, i2 = oracle.knn1(t, s)
    assert np.array_equal(i2, g[f"{tag}_i_bwd"])
    g1a, g2a = oracle.knn1_bwd(s, t, i, g[f"{tag}_wf"])
    g1b, g2b = oracle.knn1_bwd(t, s, i2, g[f"{tag}_wb"])
    np.testing.assert_allclose(g1a + g2b, g[f"{tag}_g_src"], rtol=1e-5, atol=1e-5 * np.abs(g[f"{tag}_g_src"]).max())
    np.testing.assert_allclose(g2a + g1b, g[f"{tag}_g_tgt"], rtol=1e-5, atol=1e-5 * np.abs(g[f"{tag}_g_tgt"]).max())


def test_ties_lowest_index():
    g = load_golden("chamfer_small.npz")
    d, i = oracle.knn1(g["tie_src"], g["tie_tgt"])
    assert np.array_equal(i, g["tie_i_fwd"]) and np.array_equal(d, g["tie_d_fwd"])


def test_empty_and_single():
    d, i = oracle.knn1(np.zeros((2, 3, 3), np.float32), np.zeros((2, 0, 3), np.float32))
    assert d.shape == (2, 3) and not d.any() and not i.any()
    d, i = oracle.knn1(np.ones((1, 1, 3), np.float32), np.zeros((1, 1, 3), np.float32))
    assert d[0, 0] == 3.0 and i[0, 0] == 0


def test_knn1_independent_of_thread_count():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((2, 500, 3)).astype(np.float32); b = rng.standard_normal((2, 600, 3)).astype(np.float32)
    n = oracle.num_threads()
    oracle.set_num_threads(1)
    d1, i1 = oracle.knn1(a, b)
    oracle.set_num_threads(n)
    d2, i2 = oracle.knn1(a, b)
    assert np.array_equal(d1, d2) and np.array_equal(i1, i2)
    dd = ((a[:, :, None, :].astype(np.float64) - b[:, None, :, :]) ** 2).sum(-1)
    assert (dd.argmin(-1) == i1).mean() > 0.999


def test_se3_matches_reference():
    g = load_golden("se3.npz")
    M = oracle.screw_to_transform(g["l"], g["m"], g["theta"], g["d"])
    np.testing.assert_allclose(M, g["M"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(oracle.rot6d(g["d6"]), g["R"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag", ["plain", "dist", "typed"])
def test_fk_matches_reference(tag):
    g = load_golden("fk.npz")
    dist = g["dist"] if tag != "plain" else None
    jt = g["joint_type"] if tag == "typed" else None
    out = oracle.fk(g["axis"], g["moment"], g["theta"], dist, g["order"], g["parent"], g["edge"], jt)
    np.testing.assert_allclose(out, g[f"{tag}_out"], rtol=1e-5, atol=1e-6)


def test_flow_blend_matches_reference():
    g = load_golden("flow.npz")
    for t in range(g["query"].shape[0]):
        b, m = oracle.blend_anchor_motion(g["query"][t], g["ref"][t], g["flow"][t])
        np.testing.assert_allclose(b, g["blended"][t], rtol=1e-5, atol=1e-8)
        assert np.array_equal(m, g["mask"][t])


def test_nao_known_answers(nao):
    """KAT-A / KAT-B / KAT-C of SURVEY.md section 8c through the oracle."""
    g, cano, pc_list = nao
    P = g["katA_pose"].shape[1]
    W = np.eye(P, dtype=np.float32)[g["katA_part"]]
    sk = oracle.skin_fwd(cano, W, g["katA_pose"][:, :, :3, :3], g["katA_pose"][:, :, :3, 3])
    np.testing.assert_allclose(sk[:, ::16], g["katA_skinned_s16"], rtol=1e-5, atol=1e-6)
    r = oracle.chamfer_bidir_fwd_bwd(sk, pc_list, want_grad=False)
    assert abs(r["d_fwd"].astype(np.float64).sum() - float(g["katA_sum_fwd"])) < 1e-5 * 2.2956   # 2.295648
    assert abs(r["d_bwd"].astype(np.float64).sum() - float(g["katA_sum_bwd"])) < 1e-5 * 2.3198   # 2.319767
    assert (r["i_fwd"] == g["katA_idx_fwd"]).mean() > 0.999 and (r["i_bwd"] == g["katA_idx_bwd"]).mean() > 0.999
    cpl = g["complete_pc_list"]
    rb = oracle.chamfer_bidir_fwd_bwd(cpl[0:1], cpl[1:2], want_grad=False)
    assert np.array_equal(rb["i_fwd"], g["katB_idx_fwd"]) and np.array_equal(rb["i_bwd"], g["katB_idx_bwd"])
    assert list(rb["i_fwd"][0, :8]) == [3147, 3248, 3267, 1138, 451, 3926, 1775, 2190]
    assert abs(rb["loss"] - float(g["katB_sum"])) < 1e-5 * 1.0783                                   # 1.078280
    # KAT-C: fk -> skin -> recon loss = 4.810586452
    trans = oracle.fk(g["katC_axis"], g["katC_moment"], g["katC_theta"], None, g["katC_order"], g["katC_parent"], g["katC_edge"])
    np.testing.assert_allclose(trans, g["katC_trans_list"], rtol=1e-5, atol=1e-6)
    Wc = np.eye(trans.shape[1], dtype=np.float32)[g["katC_seg_part"]]
    skc = oracle.skin_fwd(cano, Wc, trans[:, :, :3, :3], trans[:, :, :3, 3])
    rc = oracle.chamfer_bidir_fwd_bwd(skc, pc_list, want_grad=False)
    assert abs(rc["loss"] - float(g["katC_loss"])) < 1e-5 * 4.8106


def test_skin_backward_is_the_adjoint():
    rng = np.random.default_rng(1)
    T, N, P = 3, 40, 4
    cano = rng.standard_normal((N, 3)).astype(np.float32); W = rng.random((N, P)).astype(np.float32)
    R = rng.standard_normal((T, P, 3, 3)).astype(np.float32); tr = rng.standard_normal((T, P, 3)).astype(np.float32)
    g = rng.standard_normal((T, N, 3)).astype(np.float32)
    gW, gR, gt = oracle.skin_bwd(cano, W, R, tr, g)
    eps = 1e-2
    for arr, grad in ((W, gW), (R, gR), (tr, gt)):
        d = rng.standard_normal(arr.shape).astype(np.float32)
        args = {"W": W, "R": R, "tr": tr}
        def f(x):
            a = dict(args); a[[k for k, v in args.items() if v is arr][0]] = x
            return (oracle.skin_fwd(cano, a["W"], a["R"], a["tr"]).astype(np.float64) * g).sum()
        num = (f(arr + eps * d) - f(arr - eps * d)) / (2 * eps)
        assert abs(num - (grad.astype(np.float64) * d).sum()) < 2e-3 * max(1.0, abs(num))


def test_fps_starts_at_zero_and_spreads():
    rng = np.random.default_rng(2)
    xyz = rng.standard_normal((2, 300, 3)).astype(np.float32)
    idx = oracle.fps(xyz, 10)
    assert (idx[:, 0] == 0).all() and all(len(set(r)) == 10 for r in idx)
    d = ((xyz[0] - xyz[0, 0]) ** 2).sum(-1)
    assert idx[0, 1] == d.argmax()
