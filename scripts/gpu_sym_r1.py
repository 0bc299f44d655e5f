"""Parity + timing of the symmetric bidirectional kernel vs the oracle."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import oracle
from reart_b200 import _lib
from reart_b200.chamfer import _ChamferBidir
dev = torch.device("cuda"); L = _lib.lib()
rng = np.random.default_rng(0)
cases = [(3, 257, 300), (5, 20, 20), (1, 1, 7), (2, 1000, 33), (2, 4096, 4096), (1, 5000, 9000), (3, 2049, 2047), (1, 300, 40000), (1, 40000, 300)]
for (B, N, M) in cases:
    s = (rng.standard_normal((B, N, 3)) * 0.3).astype(np.float32); t = (rng.standard_normal((B, M, 3)) * 0.3).astype(np.float32)
    ref = oracle.chamfer_bidir_fwd_bwd(s, t, want_grad=False)
    S = torch.from_numpy(s).to(dev); T = torch.from_numpy(t).to(dev)
    d_f, d_b, i_f, i_b = _ChamferBidir.apply(S, T)
    ok = [np.array_equal(d_f.cpu().numpy(), ref["d_fwd"]), np.array_equal(i_f.cpu().numpy(), ref["i_fwd"]),
          np.array_equal(d_b.cpu().numpy(), ref["d_bwd"]), np.array_equal(i_b.cpu().numpy(), ref["i_bwd"])]
    print("parity", (B, N, M), ok, flush=True)
# ties: integer lattice
s = rng.integers(-2, 3, (4, 700, 3)).astype(np.float32); t = rng.integers(-2, 3, (4, 900, 3)).astype(np.float32)
ref = oracle.chamfer_bidir_fwd_bwd(s, t, want_grad=False)
d_f, d_b, i_f, i_b = _ChamferBidir.apply(torch.from_numpy(s).to(dev), torch.from_numpy(t).to(dev))
print("ties", np.array_equal(i_f.cpu().numpy(), ref["i_fwd"]), np.array_equal(i_b.cpu().numpy(), ref["i_bwd"]), np.array_equal(d_f.cpu().numpy(), ref["d_fwd"]))

def time_chamfer(B, N, iters=5):
    s = torch.randn(B, N, 3, device=dev) * 0.2; t = torch.randn(B, N, 3, device=dev) * 0.2
    d_f = torch.empty(B, N, device=dev); i_f = torch.empty(B, N, dtype=torch.int64, device=dev)
    d_b = torch.empty(B, N, device=dev); i_b = torch.empty(B, N, dtype=torch.int64, device=dev)
    nbytes = L.reart_chamfer_workspace_bytes(B, N, N); ws = _lib.workspace(nbytes, dev)
    def run():
        _lib.check(L.reart_chamfer_bidir_fwd(_lib.ptr(s), _lib.ptr(t), B, N, N, _lib.ptr(d_f), _lib.ptr(i_f), _lib.ptr(d_b), _lib.ptr(i_b), _lib.ptr(ws), nbytes, _lib.stream_ptr()), "fwd")
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    pairs = 2.0 * B * N * N
    print(f"chamfer sym fwd B={B} N={N}: {ms:.3f} ms  {pairs/ms/1e6:.1f} Gpairs/s  {8*pairs/ms/1e9/74.5:.3f} of 74.5 TF", flush=True)
for (B, N) in [(16, 4096), (64, 4096), (64, 16384), (8, 65536)]:
    time_chamfer(B, N)
