"""Time the symmetric search vs the number of target splits on shard-sized problems."""
import sys, torch
sys.path.insert(0, ".")
from reart_b200 import _lib, ops
dev = torch.device("cuda"); L = _lib.lib()
for (B, N, splits) in [(8, 16384, [0, 8, 12, 16, 18, 20, 23, 28, 32, 37]), (16, 4096, [0, 2, 4, 5, 8, 9]), (2, 4096, [0, 4, 8, 16, 18, 32])]:
    s = torch.rand(B, N, 3, device=dev) * 0.6 - 0.3; t = torch.rand(B, N, 3, device=dev) * 0.6 - 0.3
    tp = ops.pack_cloud(t)
    ka = torch.empty(B * N, dtype=torch.int64, device=dev); kb = torch.empty(B * N, dtype=torch.int64, device=dev)
    for sp in splits:
        v = sp * 16
        def run():
            _lib.check(L.reart_chamfer_sym_search(_lib.ptr(s), _lib.ptr(tp), B, N, N, _lib.ptr(ka), _lib.ptr(kb), None, v, _lib.stream_ptr()), "s")
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"B={B} N={N} splits={sp}: {ms*1e3:.1f} us  {2.0*B*N*N/ms/1e9:.2f} Tpairs/s", flush=True)
