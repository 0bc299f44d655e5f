"""Time the symmetric search variants (column sub-chunks per warp) in isolation."""
import sys, torch
sys.path.insert(0, ".")
from reart_b200 import _lib, ops
dev = torch.device("cuda"); L = _lib.lib()
for (B, N) in [(64, 16384), (16, 4096)]:
    s = torch.rand(B, N, 3, device=dev) * 0.6 - 0.3; t = torch.rand(B, N, 3, device=dev) * 0.6 - 0.3
    tp = ops.pack_cloud(t)
    ka = torch.empty(B * N, dtype=torch.int64, device=dev); kb = torch.empty(B * N, dtype=torch.int64, device=dev)
    ref = None
    for v in (1, 2, 4, 8):
        def run():
            _lib.check(L.reart_chamfer_sym_search(_lib.ptr(s), _lib.ptr(tp), B, N, N, _lib.ptr(ka), _lib.ptr(kb), None, v, _lib.stream_ptr()), "s")
        for _ in range(2): run()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        dist_a = (ka >> 32).clone(); dist_b = (kb >> 32).clone()
        if ref is None: ref = (dist_a, dist_b)
        same = bool(torch.equal(dist_a, ref[0]) and torch.equal(dist_b, ref[1]))
        print(f"B={B} N={N} S={v}: {ms:.3f} ms  {2.0*B*N*N/ms/1e9:.2f} Tpairs/s  dists equal to S=1: {same}", flush=True)
