import sys; sys.path.insert(0, ".")
import torch
from reart_b200 import ops
names = {1: "FFMA2", 12: "FFMA2x4+CREDUX(x2 per 8)", 14: "FFMA2x8 + 2x(ISETP+VOTE)", 15: "FFMA2x8 + 2xSHFL", 16: "FFMA2x8 + 2xREDUX"}
base = None
for v in (1, 14, 15, 16):
    best = min(ops.fp32_probe(v, 2000)[0] for _ in range(3))
    ms, lane_ops = ops.fp32_probe(v, 2000)
    per = best * 1e-3 * 1.965e9 / (lane_ops / 32 / (148 * 4))        # clk per FFMA2 warp-instruction per SMSP
    print(f"{names[v]:30s} {best:.3f} ms  -> {per:.3f} clk per FFMA2 (8 FFMA2 + extras = {8*per:.2f} clk)")
