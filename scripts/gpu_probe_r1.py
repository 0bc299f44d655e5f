"""Round-1 GPU exploration: parity spot checks, FP32 pipe probes, first kernel timings."""
import ctypes, json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from reart_b200 import _lib
from reart_b200.chamfer import ChamferDistance

dev = torch.device("cuda")
L = _lib.lib()
print(L.reart_version().decode(), torch.cuda.get_device_name(0))
props = torch.cuda.get_device_properties(0)
print("SMs", props.multi_processor_count)

# ---- parity
rng = np.random.default_rng(0)
cd = ChamferDistance()
import os
for (B, N, M) in [] if os.environ.get('SKIP_PARITY') else [(3, 257, 300), (5, 20, 20), (1, 1, 7), (2, 1000, 33), (2, 4096, 4096), (1, 5000, 9000)]:
    s = rng.standard_normal((B, N, 3)).astype(np.float32); t = rng.standard_normal((B, M, 3)).astype(np.float32)
    ref = oracle.chamfer_bidir_fwd_bwd(s, t)
    S = torch.from_numpy(s).to(dev).requires_grad_(True); T = torch.from_numpy(t).to(dev).requires_grad_(True)
    d_f, i_f = cd(S, T, return_index=True)
    d_b, i_b = cd(S, T, reverse=True, return_index=True)
    ok = [np.array_equal(d_f.detach().cpu().numpy(), ref["d_fwd"]), np.array_equal(i_f.cpu().numpy(), ref["i_fwd"]),
          np.array_equal(d_b.detach().cpu().numpy(), ref["d_bwd"]), np.array_equal(i_b.cpu().numpy(), ref["i_bwd"])]
    if N == M:
        tot, j_f, j_b = cd(S, T, bidirectional=True, return_index=True)
        ok += [np.array_equal(j_f.cpu().numpy(), ref["i_fwd"]), np.array_equal(j_b.cpu().numpy(), ref["i_bwd"])]
        tot.sum().backward()
        gs = S.grad.cpu().numpy(); gt = T.grad.cpu().numpy()
        ok += [np.allclose(gs, ref["grad_src"], rtol=1e-5, atol=1e-6), np.allclose(gt, ref["grad_tgt"], rtol=1e-5, atol=1e-6)]
    print("parity", (B, N, M), ok)

# ---- fp32 probes
sin = torch.rand(1024, device=dev) + 0.5
blocks = props.multi_processor_count * 8
sout = torch.empty(blocks * 256, device=dev)
names = ["FFMA", "FFMA2", "KNNMIX", "FMNMX3", "FADD2", "FFMA2+FMNMX3", "FFMA2+FMNMX", "FFMA2+VIADDMNMX", "FFMA2+VIMNMX3+IADD3", "FFMA+FMNMX", "FMNMX", "CREDUX", "FFMA2x4+CREDUX", "KNNMIX_scalar"]
res = {}
for v in range(len(names)):
    ms = ctypes.c_double(); ops = ctypes.c_double()
    best = 1e9
    for rep in range(3):
        _lib.check(L.reart_fp32_probe(v, 2000, blocks, _lib.ptr(sin), _lib.ptr(sout), ctypes.byref(ms), ctypes.byref(ops), _lib.stream_ptr()), "probe")
        best = min(best, ms.value)
    lane_ops = ops.value * blocks * 256
    rate = lane_ops / (best * 1e-3)
    res[names[v]] = rate
    print(f"probe {names[v]:20s} {best:8.3f} ms  {rate/1e12:8.3f} T lane-ops/s  per SM per clk@1.965GHz: {rate/props.multi_processor_count/1.965e9:6.1f}")
print("FFMA TFLOP/s", 2 * res["FFMA"] / 1e12, "FFMA2 TFLOP/s", 4 * res["FFMA2"] / 1e12, "KNNMIX pairs/s", 2 * res["KNNMIX"] / 1e12, "T -> alg TFLOP/s", 16 * res["KNNMIX"] / 1e12)

# ---- kernel timing
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
    print(f"chamfer fwd B={B} N={N}: {ms:.3f} ms  {pairs/ms/1e6:.1f} Gpairs/s  {8*pairs/ms/1e9/74.5:.3f} of 74.5 TF")
for (B, N) in [(16, 4096), (64, 4096), (64, 16384), (8, 65536)]:
    time_chamfer(B, N)
