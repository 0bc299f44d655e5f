"""2-GPU check: recon + flow losses under frame sharding (halo exchange) reproduce the single-GPU optimisation."""
import faulthandler, os, sys
faulthandler.dump_traceback_later(50, exit=True)
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from reart_b200.dist import DistContext
from reart_b200.engine import RelaxationEngine, tau_schedule
from reart_b200.flow_utils import FlowReference
from reart_b200.synth import make_sequence, make_flow_reference
ctx = DistContext.from_env()
dev = torch.device("cuda", ctx.local_rank)
T, N, P, c = 8, 4096, 6, 3
seq = make_sequence(T=T, N=N, P=P, seed=2)
refs, flows = make_flow_reference(seq, cano_idx=c, n_ref=1024)
cano, frames = torch.from_numpy(seq["cano"]).to(dev), torch.from_numpy(seq["frames"]).to(dev)
fr = FlowReference([torch.from_numpy(r).to(dev) for r in refs], [torch.from_numpy(f).to(dev) for f in flows])
steps = 10
for graph in (False, True):
    eng = RelaxationEngine(cano, frames, P, ctx=ctx, use_graph=graph, seed=2, flow_ref=fr, cano_idx=c, lambda_flow=1.0)
    torch.manual_seed(5); torch.cuda.manual_seed_all(5)
    sharded = [float(eng.step(tau_schedule(i, 100, 5.0, 1.0))) for i in range(steps)]
    eng.release()
    if ctx.rank == 0:
        single = RelaxationEngine(cano, frames, P, ctx=DistContext(), use_graph=graph, seed=2, flow_ref=fr, cano_idx=c, lambda_flow=1.0)
        torch.manual_seed(5); torch.cuda.manual_seed_all(5)
        ref = [float(single.step(tau_schedule(i, 100, 5.0, 1.0))) for i in range(steps)]
        norecon = RelaxationEngine(cano, frames, P, ctx=DistContext(), use_graph=False, seed=2)
        torch.manual_seed(5); torch.cuda.manual_seed_all(5)
        plain = float(norecon.step(tau_schedule(0, 100, 5.0, 1.0)))
        rel = np.abs(np.array(sharded) - np.array(ref)) / np.array(ref)
        print(f"graph={graph} sharded {sharded[0]:.5f} .. {sharded[-1]:.5f} | single {ref[0]:.5f} .. {ref[-1]:.5f} | recon only first {plain:.5f} | max rel {rel.max():.2e} first {rel[0]:.2e}", flush=True)
        single.release()
    dist.barrier()
sys.stdout.flush(); os._exit(0)
