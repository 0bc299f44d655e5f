"""Time the assignment refresh (run_robot.py:164-187): GPU solver (reart_lap) vs the reference's host path
(torch.cdist -> .cpu() -> scipy.optimize.linear_sum_assignment per frame; single process and a T-thread pool)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
from scipy.optimize import linear_sum_assignment
from reart_b200.assign import AssignLoss
from reart_b200.synth import make_sequence

dev = torch.device("cuda")
print("| shape | FPS (once) ms | GPU refresh, cold start ms | GPU refresh, warm start (clouds moved by 1e-3) ms | host refresh ms (serial) | host refresh ms (thread pool) | cost agreement |")
print("|---|---|---|---|---|---|---|")
for T, N, ds in ((9, 4096, 4), (16, 4096, 4), (64, 16384, 16), (9, 4096, 1)):
    seq = make_sequence(T=T, N=N, P=10, seed=2)
    cano, frames = torch.from_numpy(seq["cano"]).to(dev), torch.from_numpy(seq["frames"]).to(dev)
    skinned = cano[None].expand(T, N, 3).contiguous() + 0.002 * torch.randn(T, N, 3, device=dev)   # unconverged: identity pose
    torch.cuda.synchronize(); t0 = time.perf_counter()
    al = AssignLoss(cano, frames, downsample=ds)
    torch.cuda.synchronize(); t_fps = (time.perf_counter() - t0) * 1e3
    n = al.num_fps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    al.dual_u.zero_(); al.refresh(skinned); torch.cuda.synchronize()               # warm-up of the kernel itself
    t_cold = 0.0
    for _ in range(2):
        al.dual_u.zero_()
        e0.record(); al.refresh(skinned); e1.record(); torch.cuda.synchronize()
        t_cold += e0.elapsed_time(e1) / 2
    t_warm = 0.0
    for k in range(3):                                                              # each refresh sees the cloud moved a little further
        skinned = skinned + 1e-3 * torch.randn_like(skinned)
        e0.record(); al.refresh(skinned); e1.record(); torch.cuda.synchronize()
        t_warm += e0.elapsed_time(e1) / 3
    t_gpu = t_cold
    gcols = al.col4row.cpu().numpy()
    pc_src = skinned[:, al.src_idx]
    t0 = time.perf_counter()
    cost = torch.cdist(pc_src, al.pc_tgt).cpu().numpy()
    res = [linear_sum_assignment(c) for c in cost[: min(T, 9)]]
    t_serial = (time.perf_counter() - t0) * 1e3 * (T / min(T, 9))
    t0 = time.perf_counter()
    cost = torch.cdist(pc_src, al.pc_tgt).cpu().numpy()
    with ThreadPoolExecutor(max_workers=min(T, 16)) as pool:
        res_all = list(pool.map(linear_sum_assignment, cost))
    t_pool = (time.perf_counter() - t0) * 1e3
    hc = np.array([cost[t][r, c].astype(np.float64).sum() for t, (r, c) in enumerate(res_all)])
    gc = np.array([cost[t][np.arange(n), gcols[t]].astype(np.float64).sum() for t in range(T)])
    print(f"| T={T} N={N} n={n} | {t_fps:.1f} | {t_cold:.2f} | {t_warm:.2f} | {t_serial:.0f} | {t_pool:.0f} | max rel diff {np.abs(hc - gc).max() / hc.max():.1e} |", flush=True)
