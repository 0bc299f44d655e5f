"""Structure-extraction stage (reart_b200/structure.py) against vectors produced by the reference's own
utils/graph_utils.py, utils/kinematic_utils.py and screw_se3/dq_utils.py (oracle/make_golden.py:gen_structure).

The tensor code is device-agnostic torch, so the CPU run here checks the same code the GPU executes; the Chamfer
search and FPS stand-ins are the C oracle's (lowest index on ties, FPS from index 0 = the CUDA semantics, Q13).
"""
import math

import numpy as np
import pytest
import torch

import oracle
from reart_b200 import structure as st

from conftest import load_golden


@pytest.fixture(scope="module")
def g():
    return load_golden("structure.npz")


class OracleChamfer:
    """One-directional K=1 search with the ChamferDistance call signature (utils/chamfer.py:24-32)."""

    def __call__(self, src, tgt, bidirectional=False, reverse=False, reduction="mean", return_index=False):
        assert not bidirectional and not reverse
        d, i = oracle.knn1(src.numpy(), tgt.numpy())
        d, i = torch.from_numpy(d), torch.from_numpy(i.astype(np.int64))
        return (d, i) if return_index else d


def oracle_fps(xyz, npoint):
    return torch.from_numpy(oracle.fps(xyz.numpy(), npoint).astype(np.int64))


def t(a):
    return torch.from_numpy(np.asarray(a))


def test_dual_quaternion_screw_extraction_matches_reference(g):
    pose = t(g["nao_pose"]).reshape(-1, 4, 4)
    dq = st.transform_to_dq(pose)
    np.testing.assert_allclose(dq.numpy(), g["nao_dq"], rtol=1e-5, atol=1e-6)
    l, m, th, d = st.dq_to_screw(dq)
    np.testing.assert_allclose(th.numpy(), g["nao_th"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(d.numpy(), g["nao_d"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(l.numpy(), g["nao_l"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(m.numpy(), g["nao_m"], rtol=1e-4, atol=1e-5)


def test_screw_round_trip_and_branches():
    from reart_b200.screw_se3 import screw_param_to_exponential_coordinates as to_exp
    from reart_b200.screw_se3 import transform_from_exponential_coordinates as from_exp
    gen = torch.Generator().manual_seed(3)
    l = torch.nn.functional.normalize(torch.randn(64, 3, generator=gen), dim=-1)
    l = torch.where(l.sum(-1, keepdim=True) < 0, -l, l)                       # the extraction's half-space convention
    m = torch.cross(0.3 * torch.randn(64, 3, generator=gen), l, dim=-1)
    th = 0.2 + 1.3 * torch.rand(64, generator=gen)      # < pi/2: the real quaternion part is the pivot
    d = 0.2 * torch.randn(64, generator=gen)
    l2, m2, th2, d2 = st.dq_to_screw(st.transform_to_dq(from_exp(to_exp(l, m, th, d))))
    for a, b in ((l2, l), (m2, m), (th2, th), (d2, d)):
        torch.testing.assert_close(a, b, rtol=2e-4, atol=2e-5)
    # pure translation: axis = direction (flipped into the l.(1,1,1) >= 0 half space), theta pinned at 1e-6
    T = torch.eye(4).repeat(2, 1, 1)
    T[0, :3, 3] = torch.tensor([0.0, 0.3, 0.4])
    T[1, :3, 3] = torch.tensor([0.0, -0.3, -0.4])
    l3, _, th3, d3 = st.dq_to_screw(st.transform_to_dq(T))
    torch.testing.assert_close(l3, torch.tensor([[0.0, 0.6, 0.8], [0.0, 0.6, 0.8]]))
    torch.testing.assert_close(d3, torch.tensor([0.5, -0.5]))
    torch.testing.assert_close(th3, torch.full((2,), 1e-6))
    # identity: axis component 0 forced to 1, nothing is NaN
    l4, m4, th4, d4 = st.dq_to_screw(st.transform_to_dq(torch.eye(4)[None]))
    assert l4.tolist() == [[1.0, 0.0, 0.0]] and d4.item() == 0.0 and th4.item() == pytest.approx(1e-6)
    assert torch.isfinite(m4).all()


def test_relative_motion_and_geometric_cost_match_reference(g):
    pose = t(g["nao_pose"])
    ax, mo, th, di, rel = st.compute_relative_trans(pose, return_trans=True)
    off = ~torch.eye(pose.shape[1], dtype=torch.bool)                          # diagonals are identities: axis arbitrary
    np.testing.assert_allclose(th[:, off].numpy(), g["nao_rel_theta"][:, off.numpy()], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(di[:, off].numpy(), g["nao_rel_dist"][:, off.numpy()], rtol=1e-3, atol=2e-6)
    np.testing.assert_allclose(ax[:, off].numpy(), g["nao_rel_axis"][:, off.numpy()], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(mo[:, off].numpy(), g["nao_rel_moment"][:, off.numpy()], rtol=1e-3, atol=1e-4)
    geo = st.compute_geo_cost(rel, ax, mo, th, di)
    np.testing.assert_allclose(geo.numpy(), g["nao_geo_cost"], rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(st.compute_root_cost(pose).numpy(), g["nao_root_cost"], rtol=1e-5)

    world = t(g["syn_trans"])
    ax, mo, th, di, rel = st.compute_relative_trans(world, return_trans=True)
    np.testing.assert_allclose(st.compute_geo_cost(rel, ax, mo, th, di).numpy(), g["syn_geo_cost"], rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(st.compute_root_cost(world).numpy(), g["syn_root_cost"], rtol=1e-5)


def test_one_dof_projection_and_screw_cost_match_reference(g):
    recon, cost = st.compute_screw_trans(t(g["syn_local"]), return_cost=True)
    np.testing.assert_allclose(recon.numpy(), g["syn_recon"], rtol=1e-4, atol=1e-5)
    assert cost.item() == pytest.approx(float(g["syn_recon_cost"]), rel=1e-3, abs=1e-9)
    assert st.compute_screw_cost(t(g["syn_trans"]), t(g["syn_conn"])).item() == \
        pytest.approx(float(g["syn_screw_cost"]), rel=1e-3, abs=1e-9)
    new_seg, new_trans, new_conn = st.extract_kinematic(t(g["nao_merged_part"]).long(), t(g["nao_pose"]),
                                                        t(g["nao_connection"]))
    assert st.compute_screw_cost(new_trans, new_conn).item() == pytest.approx(float(g["nao_screw_cost"]), rel=1e-3)


def test_masked_screw_mean_skips_identity_frames():
    axis = torch.tensor([[[1.0, 0, 0], [0, 1.0, 0]], [[0, 0, 1.0], [0, 1.0, 0]], [[0, 0, 1.0], [0, 1.0, 0]]])
    theta = torch.tensor([[1e-6, 1e-6], [0.5, 1e-6], [0.7, 1e-6]])
    dist = torch.zeros(3, 2)
    mean_axis, mean_moment = st.compute_mean_screw_param(axis, 2 * axis, theta, dist)
    torch.testing.assert_close(mean_axis, torch.tensor([[0, 0, 1.0], [0, 1.0, 0]]))       # edge 1 is identity throughout
    torch.testing.assert_close(mean_moment, 2 * mean_axis)
    one, _ = st.compute_mean_screw_param(axis[:, :1], axis[:, :1], theta[:, :1], dist[:, :1])
    torch.testing.assert_close(one, axis[:, :1].mean(0))                                 # single edge: plain mean


def test_greedy_spanning_tree_matches_reference(g):
    cost, lab = t(g["mst_cost"]), t(g["mst_labels"])
    assert st.mst(cost).tolist() == g["mst_plain"].tolist()
    assert st.mst(cost, uni_label=lab).tolist() == g["mst_relabelled"].tolist()
    assert st.mst(cost, uni_label=lab, keep_index=True).tolist() == g["mst_plain"].tolist()
    assert st.mst(cost, uni_label=lab, max_cost=float(g["mst_cap"])).tolist() == g["mst_capped"].tolist()
    edges = st.mst(cost).tolist()                                             # a spanning tree: n-1 edges, connected
    comp = list(range(7))
    for a, b in edges:
        assert comp[a] != comp[b]
        comp = [comp[a] if c == comp[b] else c for c in comp]
    assert len(set(comp)) == 1


def test_sampling_and_spatial_costs_match_reference(g):
    cano, part, uni = t(g["nao_cano"]), t(g["nao_part"]).long(), t(g["nao_uni"])
    pts, idx = st.fps_sample_cano(cano, part, uni, num_fps=20, fps=oracle_fps)
    assert idx.tolist() == g["nao_fps_idx"].tolist()
    assert torch.equal(pts, cano[idx]) and bool((part[idx] == uni[:, None]).all())
    dist, pair = st.compute_spatial_cost(pts, OracleChamfer(), return_index=True)
    np.testing.assert_allclose(dist.numpy(), g["nao_cano_dist"], rtol=1e-5, atol=1e-9)
    assert pair.tolist() == g["nao_pair"].tolist()
    assert torch.equal(st.compute_spatial_cost(pts, OracleChamfer()), dist)
    uni_label, cano_dist, joint = st._pair_costs(part, t(g["nao_pose"]), cano, OracleChamfer(), 20, oracle_fps)
    np.testing.assert_allclose(joint.numpy(), g["nao_joint_cost"], rtol=1e-3, atol=1e-7)
    with pytest.raises(ValueError, match="too small"):
        st.fps_sample_cano(cano, part, uni, num_fps=4096, fps=oracle_fps)


def test_fps_index_list_and_joint_cost_shapes():
    gen = torch.Generator().manual_seed(0)
    pcs = torch.randn(3, 50, 3, generator=gen)
    idx = torch.randint(0, 50, (4, 5), generator=gen)
    out = st.fps_index_list(pcs, idx)
    assert out.shape == (3, 4, 5, 3) and torch.equal(out[2, 1, 3], pcs[2, idx[1, 3]])
    conn = torch.tensor([[0, 1], [2, 3], [1, 3]])
    pair = torch.tensor([[0, 4], [2, 2], [1, 0]])
    jc = st.compute_joint_cost(out, conn, pair)
    assert jc.shape == (3, 3)
    torch.testing.assert_close(jc[1, 2], ((out[1, 1, 1] - out[1, 3, 0]) ** 2).sum())
    assert st.compute_joint_cost(out[0], conn, pair).shape == (3,)


def test_merging_and_tree_search_match_reference_on_nao(g):
    cano, part, pose = t(g["nao_cano"]), t(g["nao_part"]).long(), t(g["nao_pose"])
    merged = st.merging_wrapper(part.clone(), pose, cano, OracleChamfer(), 3e-2, n_it=2, fps=oracle_fps)
    assert torch.equal(merged, t(g["nao_merged_part"]).long())
    conn = st.mst_wrapper(merged, pose, cano, OracleChamfer(), num_fps=20, cano_dist_thr=1e-2, joint_cost_weight=100,
                          fps=oracle_fps)
    assert conn.tolist() == g["nao_connection"].tolist()
    new_seg, new_trans, new_conn = st.extract_kinematic(merged, pose, conn)
    assert torch.equal(new_seg, t(g["nao_new_seg"]).long()) and new_conn.tolist() == g["nao_new_conn"].tolist()


def test_merge_graph_contracts_rigidly_coupled_parts():
    gen = torch.Generator().manual_seed(5)
    T = 4
    base = torch.eye(4).repeat(T, 4, 1, 1)
    ang = torch.rand(T, generator=gen) + 0.3
    base[:, 2, 0, 0], base[:, 2, 0, 1] = torch.cos(ang), -torch.sin(ang)
    base[:, 2, 1, 0], base[:, 2, 1, 1] = torch.sin(ang), torch.cos(ang)
    base[:, 3] = base[:, 2]                                                   # parts 2 and 3 move together, 0 and 1 static
    seg = torch.tensor([0, 0, 1, 1, 2, 2, 3, 3, 3])
    conn = torch.tensor([[0, 1], [1, 2], [2, 3]])
    new_seg, new_conn = st.merge_graph(seg, conn, base, merge_thr=3e-2, verbose=False)
    assert new_seg.tolist() == [0, 0, 0, 0, 2, 2, 2, 2, 2]
    assert new_conn.tolist() == [[0, 2]]


def test_merge_graph_matches_reference_contractions(g):
    new_seg, new_conn = st.merge_graph(t(g["mrg_seg"]), t(g["mrg_conn"]), t(g["mrg_trans"]), 3e-2, verbose=False)
    assert new_seg.tolist() == g["mrg_new_seg"].tolist()
    assert new_conn.tolist() == g["mrg_new_conn"].tolist()


def test_build_graph_matches_reference(g):
    import networkx as nx
    new_seg, new_trans, new_conn = st.extract_kinematic(t(g["nao_merged_part"]).long(), t(g["nao_pose"]),
                                                        t(g["nao_connection"]))
    G, root, axis, moment, theta, edge_index = st.build_graph(new_conn, new_trans)
    assert root == int(g["nao_root"])
    assert st.edge_index2edges(edge_index) == g["nao_edges"].tolist()
    assert list(edge_index.values()) == g["nao_edge_ids"].tolist()
    assert list(reversed(list(nx.topological_sort(G)))) == g["nao_reverse_topo"].tolist()
    np.testing.assert_allclose(axis.numpy(), g["nao_axis_list"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(moment.numpy(), g["nao_moment_list"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(theta.numpy(), g["nao_theta_list"], rtol=1e-4, atol=1e-6)

    out = st.build_graph(t(g["syn_conn"]), t(g["syn_trans"]), revolute_only=False, return_joint_type=True)
    G, root, axis, moment, theta, dist, edge_index, types = out
    assert root == int(g["syn_root"]) and st.edge_index2edges(edge_index) == g["syn_edges"].tolist()
    assert [x == "prismatic" for x in types] == g["syn_joint_prismatic"].tolist()
    np.testing.assert_allclose(theta.numpy(), g["syn_theta_list"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(dist.numpy(), g["syn_distance_list"], rtol=1e-3, atol=2e-6)
    rev = ~g["syn_joint_prismatic"]
    np.testing.assert_allclose(axis.numpy(), g["syn_axis_list"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(moment.numpy()[rev], g["syn_moment_list"][rev], rtol=1e-3, atol=1e-4)
    assert len(st.build_graph(t(g["syn_conn"]), t(g["syn_trans"]), revolute_only=False)) == 7


def test_label_filters():
    part = torch.tensor([0] * 12 + [3] * 2 + [5] * 30)
    assert st.filter_seg_label(part, min_num=10).tolist() == [0, 5]

    class NearestStub:
        k = 1

        def __call__(self, ref, query):
            d = torch.cdist(query, ref)
            return d.min(dim=2, keepdim=True)

    pc = torch.cat([torch.zeros(12, 3), torch.tensor([[0.9, 0.9, 0.9], [0.1, 0, 0]]), torch.ones(30, 3)])
    out = st.denoise_seg_label(part.clone(), pc, NearestStub(), min_num=10)
    assert out[12].item() == 5 and out[13].item() == 0 and out.unique().tolist() == [0, 5]


def test_model_selection_energy_terms_match_reference(g):
    """ass_err + screw_err + group_err is the energy the candidate fits are ranked by (run_robot.py:306-314)."""
    from reart_b200 import model_utils as mu
    cano, pose = t(g["nao_cano"]), t(g["nao_pose"])
    new_seg, new_trans, _ = st.extract_kinematic(t(g["nao_merged_part"]).long(), pose, t(g["nao_connection"]))
    R, tr = new_trans[:, new_seg, :3, :3], new_trans[:, new_seg, :3, 3]
    pred = torch.einsum("tnij,nj->tni", R, cano) + tr                         # hard-label skin (compute_pc_transform)
    nao = load_golden("nao.npz")
    complete_gt = t(nao["complete_pc_list"])
    cidx = int(nao["cano_idx"])
    complete = torch.cat((pred[:cidx], cano[None], pred[cidx:]), dim=0)
    assert mu.compute_group_temporal_err(complete, new_seg).item() == pytest.approx(float(g["nao_group_err"]), rel=1e-4)
    pc_list = torch.cat((complete_gt[:cidx], complete_gt[cidx + 1:]))
    sub = t(g["nao_ass_sub"])
    for pooled in (False, True):
        err = mu.compute_ass_err(pred[:, sub].contiguous(), pc_list[:, sub].contiguous(), use_nproc=pooled)
        assert err.item() == pytest.approx(float(g["nao_ass_err"]), rel=1e-4)
    root = new_trans[:, 3]
    aligned = mu.compute_align_trans(new_trans, root)
    torch.testing.assert_close(aligned[:, 3], torch.eye(4).expand(new_trans.shape[0], 4, 4), atol=1e-5, rtol=0)
