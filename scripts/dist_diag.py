"""Stage-by-stage NCCL diagnostic (each stage prints; run under `timeout`)."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(90, exit=True)
sys.path.insert(0, ".")
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); ws = int(os.environ["WORLD_SIZE"])
def say(*a): print(f"[r{rank} {time.time()%1000:.1f}]", *a, flush=True)
torch.cuda.set_device(lr)
say("init pg")
dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", lr))
say("pg ok")
x = torch.ones(1024, device="cuda") * (rank + 1)
dist.all_reduce(x); torch.cuda.synchronize(); say("eager allreduce", float(x[0]))
dist.barrier(); say("barrier ok")
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        dist.all_reduce(x)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize(); say("side-stream allreduce ok")
g = torch.cuda.CUDAGraph()
y = torch.ones(1024, device="cuda")
with torch.cuda.graph(g):
    y.mul_(2.0)
    dist.all_reduce(y)
say("captured")
g.replay(); torch.cuda.synchronize(); say("replayed", float(y[0]))
from reart_b200.dist import DistContext
from reart_b200.engine import RelaxationEngine
from reart_b200.synth import make_sequence
ctx = DistContext(rank, ws, lr, "nccl")
seq = make_sequence(T=8, N=2048, P=6, seed=2)
eng = RelaxationEngine(torch.from_numpy(seq["cano"]).cuda(), torch.from_numpy(seq["frames"]).cuda(), 6, ctx=ctx, use_graph=False)
l = eng.step(1.0); torch.cuda.synchronize(); say("eager engine step", float(l))
eng2 = RelaxationEngine(torch.from_numpy(seq["cano"]).cuda(), torch.from_numpy(seq["frames"]).cuda(), 6, ctx=ctx, use_graph=True)
l = eng2.step(1.0); torch.cuda.synchronize(); say("graph engine step", float(l))
l = eng2.step(1.0); torch.cuda.synchronize(); say("graph engine step 2", float(l))
dist.barrier(); dist.destroy_process_group(); say("done")
