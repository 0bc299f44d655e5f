"""Structure extraction and retargeting on the GPU: the same torch code as tests/test_structure.py, now driving the
sm_100a kernels (P^2-batch Chamfer search at 20 points, padded batched FPS, fused FK + skinning in the IK loop)."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_structure_stage_on_gpu_matches_reference_golden():
    from reart_b200 import structure as st
    from reart_b200.chamfer import ChamferDistance
    g = load_golden("structure.npz")
    cano, part, pose = cu(g["nao_cano"]), cu(g["nao_part"]).long(), cu(g["nao_pose"])
    cd = ChamferDistance()

    pts, idx = st.fps_sample_cano(cano, part, cu(g["nao_uni"]), num_fps=20)          # one padded reart_fps launch
    assert idx.cpu().tolist() == g["nao_fps_idx"].tolist()
    dist, pair = st.compute_spatial_cost(pts, cd, return_index=True)                 # B = P^2 = 100, N = M = 20
    np.testing.assert_allclose(dist.cpu().numpy(), g["nao_cano_dist"], rtol=1e-5, atol=1e-9)
    assert pair.cpu().tolist() == g["nao_pair"].tolist()

    ax, mo, th, di, rel = st.compute_relative_trans(pose, return_trans=True)
    geo = st.compute_geo_cost(rel, ax, mo, th, di)
    np.testing.assert_allclose(geo.cpu().numpy(), g["nao_geo_cost"], rtol=2e-3, atol=1e-6)

    merged = st.merging_wrapper(part.clone(), pose, cano, cd, 3e-2, n_it=2)
    assert torch.equal(merged.cpu(), torch.from_numpy(g["nao_merged_part"].astype(np.int64)))
    conn = st.mst_wrapper(merged, pose, cano, cd, num_fps=20, cano_dist_thr=1e-2, joint_cost_weight=100)
    assert conn.cpu().tolist() == g["nao_connection"].tolist()
    new_seg, new_trans, new_conn = st.extract_kinematic(merged, pose, conn)
    G, root, axis, moment, theta, edge_index = st.build_graph(new_conn, new_trans)
    assert root == int(g["nao_root"]) and st.edge_index2edges(edge_index) == g["nao_edges"].tolist()
    np.testing.assert_allclose(theta.cpu().numpy(), g["nao_theta_list"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(axis.cpu().numpy(), g["nao_axis_list"], rtol=2e-3, atol=2e-4)
    assert st.compute_screw_cost(new_trans, new_conn).item() == pytest.approx(float(g["nao_screw_cost"]), rel=5e-3)


def test_retarget_recovers_joint_angles_with_the_fused_kernels():
    from reart_b200 import retarget as rt
    from reart_b200.knn_module import KNN
    from reart_b200.model import KinematicModel
    rng = np.random.default_rng(5)
    N, P, S = 2000, 4, 3
    E = P - 1
    axis = rng.standard_normal((E, 3)); axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    anchor = np.cumsum(np.full((E, 3), 0.08), axis=0) * np.array([1.0, 0.2, 0.1])
    moment = np.cross(anchor, axis)
    part = rng.integers(0, P, N)
    cano = (rng.uniform(-0.04, 0.04, (N, 3)) + np.concatenate([[np.zeros(3)], anchor])[part]).astype(np.float32)
    model = KinematicModel(pose_len=S, seg_part=cu(part), cano_pc=cu(cano), knn=KNN(k=1, transpose_mode=True),
                           edge_index={f"{c}_{c-1}": c - 1 for c in range(1, P)},
                           paths_to_base={c: list(range(c, -1, -1)) for c in range(P)}, reverse_topo=list(range(P)),
                           axis_list=cu(axis.astype(np.float32)), moment_list=cu(moment.astype(np.float32))).cuda()
    truth = rng.uniform(-0.8, 0.8, (S, E)).astype(np.float32)
    order = np.arange(P, dtype=np.int32); parent = np.arange(-1, P - 1, dtype=np.int32)
    poses = oracle.fk(axis, moment, truth, None, order, parent, parent.copy())
    novel = oracle.skin_fwd(cano, np.eye(P, dtype=np.float32)[part], poses[:, :, :3, :3], poses[:, :, :3, 3])
    pick = np.array([np.flatnonzero(part == p)[10] for p in range(P)])             # one point per part, index 10
    fitted = rt.retarget(model, cu(cano[pick]), cu(novel[:, pick]), n_iter=200)
    err, posed, seg = rt.retarget_error(model, cu(cano), cu(novel), fitted)
    assert posed.shape == (S, N, 3) and torch.equal(seg.cpu(), torch.from_numpy(part))
    assert float(err.max()) < 1.0, err                                             # 100 x metres
    single = rt.retarget(model, cu(cano[pick]), cu(novel[1, pick]), n_iter=200)["theta_list"]
    torch.testing.assert_close(fitted["theta_list"][1:2], single, rtol=1e-3, atol=1e-4)
