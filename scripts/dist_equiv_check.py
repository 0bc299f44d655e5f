"""2-GPU check: the frame-sharded optimisation (one-shot all-reduce inside the graph) reproduces the single-GPU one."""
import faulthandler, os, sys
faulthandler.dump_traceback_later(120, exit=True)
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from reart_b200.dist import DistContext
from reart_b200.engine import RelaxationEngine, tau_schedule
from reart_b200.synth import make_sequence
ctx = DistContext.from_env()
dev = torch.device("cuda", ctx.local_rank)
seq = make_sequence(T=8, N=4096, P=6, seed=2)
cano, frames = torch.from_numpy(seq["cano"]).to(dev), torch.from_numpy(seq["frames"]).to(dev)
steps = 12
for graph in (False, True):
    eng = RelaxationEngine(cano, frames, 6, ctx=ctx, use_graph=graph, seed=2)
    torch.manual_seed(5); torch.cuda.manual_seed_all(5)
    sharded = [float(eng.step(tau_schedule(i, 100, 5.0, 1.0))) for i in range(steps)]
    eng.release()
    if ctx.rank == 0:
        single = RelaxationEngine(cano, frames, 6, ctx=DistContext(), use_graph=graph, seed=2)
        torch.manual_seed(5); torch.cuda.manual_seed_all(5)
        ref = [float(single.step(tau_schedule(i, 100, 5.0, 1.0))) for i in range(steps)]
        rel = np.abs(np.array(sharded) - np.array(ref)) / np.array(ref)
        print(f"graph={graph} sharded {sharded[:3]} ... {sharded[-1]:.4f} | single {ref[:3]} ... {ref[-1]:.4f} | max rel diff {rel.max():.2e} (first step {rel[0]:.2e})", flush=True)
        single.release()
    dist.barrier()
sys.stdout.flush(); os._exit(0)
