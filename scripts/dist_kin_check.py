"""2-GPU check: the frame-sharded KinematicEngine (axis/moment shared, theta per frame) reproduces the single-GPU run."""
import faulthandler, os, sys
faulthandler.dump_traceback_later(50, exit=True)
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from reart_b200.dist import DistContext
from reart_b200.engine import KinematicEngine
from reart_b200.synth import make_sequence, kinematic_init
ctx = DistContext.from_env()
dev = torch.device("cuda", ctx.local_rank)
seq = make_sequence(T=8, N=4096, P=6, seed=2)
kw = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in kinematic_init(seq, theta_noise=0.1).items()}
cano, frames, part = torch.from_numpy(seq["cano"]).to(dev), torch.from_numpy(seq["frames"]).to(dev), torch.from_numpy(seq["part"])
def clone(kw): return {k: (v.clone() if torch.is_tensor(v) else v) for k, v in kw.items()}
for graph in (False, True):
    eng = KinematicEngine(clone(kw), part, cano, frames, ctx=ctx, use_graph=graph)
    sharded = [float(eng.step()) for _ in range(10)]
    eng.release()
    if ctx.rank == 0:
        single = KinematicEngine(clone(kw), part, cano, frames, ctx=DistContext(), use_graph=graph)
        ref = [float(single.step()) for _ in range(10)]
        rel = np.abs(np.array(sharded) - np.array(ref)) / np.array(ref)
        print(f"graph={graph} sharded {sharded[0]:.5f} .. {sharded[-1]:.5f} | single {ref[0]:.5f} .. {ref[-1]:.5f} | max rel {rel.max():.2e} first {rel[0]:.2e}", flush=True)
        single.release()
    dist.barrier()
sys.stdout.flush(); os._exit(0)
