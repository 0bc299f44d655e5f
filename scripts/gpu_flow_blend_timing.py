"""Flow-blend kernel on the 64-pair workload of scripts/gpu_kernel_zoo.py: brute force vs the windowed exact kernel."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from reart_b200 import ops
from reart_b200.flow_utils import FlowReference, blend_anchor_motion_batched
from reart_b200.synth import make_sequence, make_flow_reference

dev = torch.device("cuda")
T, N, P = 64, 16384, 15
seq = make_sequence(T, N, P, seed=2)
cano = torch.from_numpy(seq["cano"]).to(dev); frames = torch.from_numpy(seq["frames"]).to(dev)
refs, flows = make_flow_reference(seq, cano_idx=0, n_ref=N // 4)
ref = FlowReference([torch.from_numpy(r).to(dev) for r in refs], [torch.from_numpy(f).to(dev) for f in flows])
q = torch.cat((cano[None], frames), 0)[:-1].contiguous()
nref = float(sum(r.shape[0] for r in refs))
pairs = N * nref

def timeit(name, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"| {name} | {ms*1e3:9.1f} us | {pairs/(ms*1e-3)/1e12:6.2f}e12 pairs/s credited |", flush=True)

brute = lambda: ops.knn3_blend(q, ref.ref_cat, ref.flow_cat, ref.offsets)
wind = lambda: blend_anchor_motion_batched(q, ref)
Bb, Mb = brute(); Bw, Mw = wind()
print("bit-identical:", bool(torch.equal(Bb, Bw) and torch.equal(Mb, Mw)), "max|diff|", float((Bb - Bw).abs().max()))
timeit("brute force `reart_knn3_blend` (64 pairs x 16384 queries x ~3500 refs)", brute)
timeit("windowed exact `reart_knn3_blend_sorted` (query bucketing + slab walk)", wind)
timeit("one-off `reart_flow_refs_sort` of the 64 reference sets", lambda: ops.flow_refs_sort(ref.ref_cat, ref.offsets, ref.max_refs))
