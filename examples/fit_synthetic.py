"""Minimal end-to-end use of reart_b200 on a synthetic articulated sequence (needs a B200 / CUDA device).

    python examples/fit_synthetic.py                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/fit_synthetic.py

Relaxation fit (run_robot.py --model=base, recon loss) -> hard part labels and per-frame part poses, then the
snapshot metrics of utils/eval_utils.py on the GPU.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from reart_b200 import eval_utils
from reart_b200.dist import DistContext
from reart_b200.engine import fit_relaxation
from reart_b200.model_utils import compute_pc_transform
from reart_b200.synth import make_sequence


def main():
    ctx = DistContext.from_env()
    dev = torch.device("cuda", ctx.local_rank)
    torch.cuda.set_device(dev)
    T, N, P = 16, 4096, 6
    seq = make_sequence(T=T, N=N, P=P, seed=2)
    cano = torch.from_numpy(seq["cano"]).to(dev)
    frames = torch.from_numpy(seq["frames"]).to(dev)
    engine, losses = fit_relaxation(cano, frames, num_parts=P, n_iter=600, ctx=ctx, log_every=100)
    if ctx.is_main:
        print("recon loss every 100 iterations:", [round(x, 3) for x in losses])
    lo, hi = engine.frame_range
    with torch.no_grad():
        seg = engine.model.seg_forward(cano, argmax=True)
        R, tr = engine.model.pose()
        pose = torch.zeros(hi - lo, P, 4, 4, device=dev)
        pose[:, :, :3, :3], pose[:, :, :3, 3], pose[:, :, 3, 3] = R, tr, 1.0
        skinned = compute_pc_transform(cano, pose, seg)
        cd = eval_utils.compute_chamfer_list(skinned, frames[lo:hi], reduction="mean")
        ri = eval_utils.eval_seg(torch.from_numpy(seq["part"]).to(dev), seg)
    print(f"[rank {ctx.rank}] frames {lo}..{hi - 1}: mean Chamfer {cd:.3e}, Rand index vs ground-truth parts {float(ri):.3f}")
    engine.release()
    ctx.barrier()
    if ctx.world_size > 1:
        os._exit(0)


if __name__ == "__main__":
    main()
