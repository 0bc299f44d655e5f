"""2+ GPU check of the one-shot all-reduce kernel against NCCL, eager and inside a CUDA graph, plus timing."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(100, exit=True)
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from reart_b200.dist import DistContext, OneShotAllReduce
ctx = DistContext.from_env()
dev = torch.device("cuda", ctx.local_rank)
def say(*a):
    print(f"[r{ctx.rank}]", *a, flush=True)
n = 2433
ar = OneShotAllReduce(ctx, n, dev)
say("rendezvous ok")
g = torch.Generator(device=dev).manual_seed(100 + ctx.rank)
ok = True
for it in range(6):
    x = torch.randn(n, device=dev, generator=g)
    ref = x.clone(); dist.all_reduce(ref)
    y = x.clone(); ar(y); torch.cuda.synchronize()
    err = float((y - ref).abs().max())
    ok &= err < 1e-5
    if it < 2: say("eager iter", it, "max err vs nccl", err)
# identical bits on all ranks
chk = y.clone(); dist.broadcast(chk, src=0)
say("bitwise identical across ranks:", bool(torch.equal(chk, y)), "all ok:", ok)
# inside a CUDA graph
buf = torch.zeros(n, device=dev)
src = torch.randn(n, device=dev, generator=g)
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        buf.copy_(src); ar(buf)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize(); dist.barrier()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    buf.copy_(src); ar(buf)
ref = src.clone(); dist.all_reduce(ref)
for it in range(5):
    gr.replay()
torch.cuda.synchronize()
say("graph replay max err", float((buf - ref).abs().max()))
# timing: one-shot vs NCCL, both graph-captured, 200 replays
def bench(fn):
    g2 = torch.cuda.CUDAGraph()
    s2 = torch.cuda.Stream(); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s2); torch.cuda.synchronize(); dist.barrier()
    with torch.cuda.graph(g2):
        for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g2.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 200 * 1e3, g2
t1, ga = bench(lambda: ar(buf))
t2, gb = bench(lambda: dist.all_reduce(buf))
say(f"one-shot {t1:.1f} us per all-reduce, nccl {t2:.1f} us")
del gr, ga, gb
torch.cuda.synchronize(); dist.barrier()
sys.stdout.flush(); os._exit(0)
