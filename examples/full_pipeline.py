"""The two-stage flow of the reference's run_robot.py on a synthetic articulated sequence, end to end on one GPU:

    relaxation fit (--model=base)  ->  denoise / merge parts  ->  spanning tree  ->  build_graph
    ->  projection fit (--model=kinematic)  ->  IK retargeting to unseen poses

    python examples/full_pipeline.py

Every stage runs on this package's kernels; the ground-truth tree of the generator is only used to report how much of
it was recovered.
"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import networkx as nx
import numpy as np
import torch

from reart_b200 import eval_utils, retarget, structure as st
from reart_b200.chamfer import ChamferDistance
from reart_b200.engine import KinematicEngine, fit_relaxation
from reart_b200.knn_module import KNN
from reart_b200.synth import _rodrigues, make_sequence


def pose_from_theta(seq, theta):
    """[S,P] joint angles -> [S,P,4,4] part poses on the generator's tree (same recursion as synth.make_sequence)."""
    S, P = theta.shape
    pose = np.tile(np.eye(4), (S, P, 1, 1))
    for s in range(S):
        for p in range(1, P):
            Rl = _rodrigues(seq["axes"][p], theta[s, p])
            local = np.eye(4)
            local[:3, :3], local[:3, 3] = Rl, seq["anchors"][p] - Rl @ seq["anchors"][p]
            pose[s, p] = pose[s, seq["parent"][p]] @ local
    return pose.astype(np.float32)


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(2)
    T, N, P = 16, 4096, 6
    seq = make_sequence(T=T, N=N, P=P, seed=3)
    cano, frames = torch.from_numpy(seq["cano"]).to(dev), torch.from_numpy(seq["frames"]).to(dev)
    gt_part = torch.from_numpy(seq["part"]).to(dev)
    gt_edges = {frozenset((p, int(seq["parent"][p]))) for p in range(1, P)}

    # ---- stage 1: relaxation (run_robot.py:153-221, recon loss, 2P part proposals as in the reference's 20 for ~10)
    t0 = time.perf_counter()
    engine, losses = fit_relaxation(cano, frames, num_parts=2 * P, n_iter=1500, log_every=500)
    with torch.no_grad():
        seg = engine.model.seg_forward(cano, argmax=True)
        R, tr = engine.model.pose()
        trans = torch.zeros(T, 2 * P, 4, 4, device=dev)
        trans[:, :, :3, :3], trans[:, :, :3, 3], trans[:, :, 3, 3] = R, tr, 1.0
    engine.release()
    torch.cuda.synchronize()
    print(f"[relax] {time.perf_counter() - t0:.2f} s, recon loss {[round(x, 3) for x in losses]}, "
          f"{seg.unique().numel()} parts used, Rand index {float(eval_utils.eval_seg(gt_part, seg)):.3f}")

    try:
        # ---- structure extraction (run_robot.py:232-243, :101-124)
        t0 = time.perf_counter()
        cd, knn = ChamferDistance(), KNN(k=1, transpose_mode=True)
        seg = st.denoise_seg_label(seg.clone(), cano, knn, min_num=20)
        if seg.unique().numel() > 1:
            seg = st.merging_wrapper(seg, trans, cano, cd, 3e-2, n_it=2)
        conn = st.mst_wrapper(seg, trans, cano, cd, num_fps=20, cano_dist_thr=1e-2, joint_cost_weight=100)
        new_seg, new_trans, new_conn = st.extract_kinematic(seg, trans, conn)
        G, root, axis, moment, theta, edge_index = st.build_graph(new_conn, new_trans)
        torch.cuda.synchronize()
        # name each recovered part by the ground-truth part most of its points belong to
        votes = torch.zeros(int(new_seg.max()) + 1, P, device=dev).index_put_((new_seg, gt_part), torch.ones(N, device=dev),
                                                                               accumulate=True)
        name = votes.argmax(dim=1).tolist()
        got = {frozenset((name[a], name[b])) for a, b in new_conn.tolist() if name[a] != name[b]}
        print(f"[structure] {1e3 * (time.perf_counter() - t0):.1f} ms: {new_seg.unique().numel()} parts after merging, "
              f"Rand index {float(eval_utils.eval_seg(gt_part, new_seg)):.3f}, root -> gt part {name[root]}, "
              f"{len(got & gt_edges)} of {P - 1} ground-truth edges recovered, "
              f"screw cost {float(st.compute_screw_cost(new_trans, new_conn)):.2e}")

        # ---- stage 2: projection onto the kinematic model (run_robot.py --model=kinematic)
        t0 = time.perf_counter()
        kw = dict(edge_index=edge_index, paths_to_base=nx.shortest_path(G, target=root),
                  reverse_topo=list(reversed(list(nx.topological_sort(G)))), axis_list=axis, moment_list=moment,
                  theta_list=theta)
        keng = KinematicEngine(kw, new_seg, cano, frames, lr=1e-2)
        first = float(keng.step())
        for _ in range(400):
            last = keng.step()
        last = float(last)
        with torch.no_grad():
            skinned = keng.model(cano)[0]
        cdist = eval_utils.compute_chamfer_list(skinned, frames, reduction="mean")
        print(f"[kinematic] {time.perf_counter() - t0:.2f} s, recon loss {first:.3f} -> {last:.3f}, "
              f"mean Chamfer to the observed frames {cdist:.3e}")

        # ---- IK retargeting to 3 unseen poses from one point per ground-truth part (kinematic_utils.py:201-266)
        t0 = time.perf_counter()
        novel_theta = np.random.default_rng(11).uniform(-0.8, 0.8, (3, P))
        novel_pose = torch.from_numpy(pose_from_theta(seq, novel_theta)).to(dev)
        per_point = novel_pose[:, gt_part]                                                     # (S,N,4,4)
        novel = torch.einsum("snij,nj->sni", per_point[..., :3, :3], cano) + per_point[..., :3, 3]
        pick = torch.stack([torch.nonzero(gt_part == p)[10, 0] for p in range(P)])
        model = keng.model
        fitted = retarget.retarget(model, cano[pick], novel[:, pick], n_iter=200)
        err, _, _ = retarget.retarget_error(model, cano, novel, fitted)
        torch.cuda.synchronize()
        print(f"[ik] {1e3 * (time.perf_counter() - t0):.0f} ms for 3 poses x 200 iterations, retarget error (cm) "
              f"{[round(e, 2) for e in err.tolist()]}")
        keng.release()
    except Exception:                                                                           # keep the stage-1 report
        traceback.print_exc()


if __name__ == "__main__":
    main()
