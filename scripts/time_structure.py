"""Wall-clock of the structure-extraction stage and the IK retargeting loop on the GPU (nao-sized problem).

    python scripts/time_structure.py            # prints one JSON object, also written to gpurun_out/structure_timing.json

Host-driven stages (spanning tree, merging, graph building) are timed with perf_counter around a device synchronise;
the IK loop is 200 Adam iterations over S novel states at once (reart_b200/retarget.py).
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import networkx as nx                                                                    # noqa: E402
from reart_b200 import retarget as rt, structure as st                                   # noqa: E402
from reart_b200.chamfer import ChamferDistance                                           # noqa: E402
from reart_b200.knn_module import KNN                                                    # noqa: E402
from reart_b200.model import KinematicModel                                              # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "structure.npz"))
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    cano, part, pose = cu(g["nao_cano"]), cu(g["nao_part"]).long(), cu(g["nao_pose"])
    cd = ChamferDistance()
    res = {"T": int(pose.shape[0]), "P": int(pose.shape[1]), "N": int(cano.shape[0])}
    res["merging_wrapper_ms"], merged = timed(lambda: st.merging_wrapper(part.clone(), pose, cano, cd, 3e-2, n_it=2))
    res["mst_wrapper_ms"], conn = timed(lambda: st.mst_wrapper(merged, pose, cano, cd))
    new_seg, new_trans, new_conn = st.extract_kinematic(merged, pose, conn)
    res["build_graph_ms"], built = timed(lambda: st.build_graph(new_conn, new_trans))
    res["screw_cost_ms"], _ = timed(lambda: st.compute_screw_cost(new_trans, new_conn))
    res["relative_trans_geo_cost_ms"], _ = timed(
        lambda: st.compute_geo_cost(*(lambda a, m, t, d, r: (r, a, m, t, d))(*st.compute_relative_trans(pose, True))))

    G, root, axis, moment, theta, edge_index = built
    model = KinematicModel(pose_len=pose.shape[0], seg_part=new_seg, cano_pc=cano, knn=KNN(k=1, transpose_mode=True),
                           edge_index=edge_index, paths_to_base=nx.shortest_path(G, target=root),
                           reverse_topo=list(reversed(list(nx.topological_sort(G)))),
                           axis_list=axis, moment_list=moment, theta_list=theta).cuda()
    S = 3
    with torch.no_grad():
        novel = model(cano, theta_list=theta[:S] * 0.7)[0]
    pick = torch.stack([torch.nonzero(new_seg == p)[10, 0] for p in range(int(new_seg.max()) + 1)])
    rt.retarget(model, cano[pick], novel[:, pick], n_iter=5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fitted = rt.retarget(model, cano[pick], novel[:, pick], n_iter=200)
    torch.cuda.synchronize()
    res["ik_200_iters_3_states_ms"] = (time.perf_counter() - t0) * 1e3
    err, _, _ = rt.retarget_error(model, cano, novel, fitted)
    res["ik_retarget_err_cm"] = [round(e, 4) for e in err.tolist()]
    res["device"] = torch.cuda.get_device_name(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "structure_timing.json"), "w") as f:
        json.dump(res, f)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
