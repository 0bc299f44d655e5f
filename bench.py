"""bench.py -- skinned-Chamfer forward+backward throughput of the reart energy evaluation on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg3_16k]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE full optimisation iteration of the relaxation model (run_robot.py:154-221, --model=base,
recon loss): seg MLP -> gumbel weights -> 6D->R -> skin -> bidirectional Chamfer against every observed
frame -> loss -> backward -> (N>1: one NCCL all-reduce of the shared grads) -> Adam, on the synthetic
sequence named by --workload.  Metric: directed point pairs / s = 2*T*N*M per step / time, whole job.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (T frames, N = M points, P parts) -- BASELINE.json configs
    "cfg3_16k": (64, 16384, 15),      # the sequence the north_star quotes its targets on (default)
    "cfg2": (16, 4096, 15),           # nao-shaped relaxation model
    "cfg3_4k": (64, 4096, 15),
    "cfg3_64k": (64, 65536, 15),
    "cfg4": (32, 16384, 8),           # sapien-shaped kinematic projection model (KinematicEngine)
    "cfg5": (64, 32768, 15),
    "cfg3_16k_8f": (8, 16384, 15),    # one rank's shard of cfg3_16k at 8 GPUs (tuning aid)
    "tiny": (8, 2048, 6),
}
METRIC = "skinned_chamfer_fwd_bwd_directed_point_pairs_per_s"
UNIT = "pairs/s"
FLOP_PER_PAIR = 8.0                   # 3 FSUB + 1 FMUL + 2 FFMA


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3_16k", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def config_dict(args, T, N, P, world):
    kind = "kinematic projection model" if args.workload == "cfg4" else "relaxation model (base)"
    return {"workload": f"{args.workload}: {kind}, P={P}, synthetic sequence T={T} frames x "
                        f"N=M={N} points, skin + bidirectional Chamfer recon loss fwd+bwd + Adam",
            "T": T, "N": N, "M": N, "P": P, "frames_per_gpu": T // max(world, 1) if args.scaling == "strong" else T,
            "partitioning": f"frames sharded over {world} rank(s), one all-reduce of shared grads per step",
            "l2": "L2 flushed (256 MiB write) between timed steps", "seed": 2}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_throughput(T, N, budget_s=12.0):
    """Time the oracle port (OpenMP, all host cores) on a bounded sample of the workload: `frames` frames of
    N x N bidirectional Chamfer forward + backward.  Returns (pairs_per_s, cores, sample description)."""
    import numpy as np
    import oracle
    from reart_b200.synth import make_sequence
    oracle.build()
    cores = os.cpu_count() or 1
    oracle.set_num_threads(cores)
    seq = make_sequence(T=2, N=N, P=15, seed=2)
    src = np.ascontiguousarray(seq["frames"][0:1]); tgt = np.ascontiguousarray(seq["frames"][1:2])
    t0 = time.perf_counter()
    oracle.chamfer_bidir_fwd_bwd(src, tgt)
    one = time.perf_counter() - t0
    frames = int(max(1, min(T, budget_s / max(one, 1e-6))))
    src = np.repeat(src, frames, axis=0); tgt = np.repeat(tgt, frames, axis=0)
    t0 = time.perf_counter()
    oracle.chamfer_bidir_fwd_bwd(src, tgt)
    dt = time.perf_counter() - t0
    pairs = 2.0 * frames * N * N
    return pairs / dt, oracle.num_threads(), f"{frames} of {T} frames x {N}x{N} points, bidirectional Chamfer fwd+bwd, oracle C port ({dt:.2f} s)"


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The reference's
    own arithmetic lives in absent third-party CUDA packages, so this is the oracle port (kind 'port')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T, N, P = WORKLOADS[args.workload]
    steps, warm = max(args.steps, 1), max(args.warmup, 0)
    budget = max(1.0, min(12.0, 150.0 / (steps + warm)))
    vals, cores, sample = [], 1, ""
    for i in range(warm + steps):
        v, cores, sample = cpu_reference_throughput(T, N, budget_s=budget)
        if i >= warm:
            vals.append(v)
    value = sum(vals) / len(vals)
    pairs_per_step = 2.0 * T * N * N
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": pairs_per_step / value * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, T, N, P, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region.  The timed region of the default run is < 0.1 s,
    shorter than nvidia-smi's start-up, so the primary source is an in-process NVML thread (the library nvidia-smi
    itself reads) polling every ~4 ms between begin() and end(); an `nvidia-smi -lms 20` subprocess started early is
    the fallback, filtered to the same wall-clock window."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, uuid: str | None = None):
        import threading
        self.samples = []                                  # (wall time, sm MHz, max MHz, watts, reasons tuple)
        self.t0 = self.t1 = None
        self._stop = threading.Event()
        self._thread = None
        self.smi_path, self.smi = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

            def poll():
                while not self._stop.is_set():
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        try:
                            r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.samples.append((time.time(), sm, mx, pw, tuple(k for k, b in bits.items() if r & b)))
                    except Exception:
                        pass
                    time.sleep(0.004)

            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None
        try:
            self.smi_path = tempfile.NamedTemporaryFile(suffix=".csv", delete=False).name
            self.smi = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                         "-lms", "20", "-i", str(gpu_index)], stdout=open(self.smi_path, "w"),
                                        stderr=subprocess.DEVNULL)
        except Exception:
            self.smi = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def _smi_samples(self):
        import datetime
        out = []
        if self.smi is None:
            return out
        self.smi.terminate()
        try:
            self.smi.wait(timeout=5)
        except Exception:
            self.smi.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.smi_path):
            parts = [q.strip() for q in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                out.append((ts, float(parts[1]), float(parts[2]), float(parts[3]),
                            tuple(n for n, v in zip(names, parts[4:8]) if v.lower().startswith("active"))))
            except ValueError:
                continue
        os.unlink(self.smi_path)
        return out

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        source = "nvml"
        sel = [q for q in self.samples if self.t0 is not None and self.t0 <= q[0] <= (self.t1 or time.time())]
        smi = self._smi_samples()
        if len(sel) < 3:
            sel = [q for q in smi if self.t0 is not None and self.t0 - 0.02 <= q[0] <= (self.t1 or time.time()) + 0.02]
            source = "nvidia-smi"
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": len(sel), "source": source}
        if sel:
            sm = sorted(q[1] for q in sel)
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(q[2] for q in sel),
                       reasons=sorted({r for q in sel for r in q[4]}), power_w_max=max(q[3] for q in sel))
        return out


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    from reart_b200 import _lib, ops
    from reart_b200.dist import DistContext
    from reart_b200.engine import RelaxationEngine, tau_schedule
    from reart_b200.synth import make_sequence

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: reart_b200 has no CPU fallback")
    ctx = DistContext.from_env()
    world = ctx.world_size
    assert world == args.gpus or world == 1, f"launched with WORLD_SIZE={world} but --gpus {args.gpus}"
    dev = torch.device("cuda", ctx.local_rank)
    torch.cuda.set_device(dev)
    L = _lib.lib()

    T, N, P = WORKLOADS[args.workload]
    T_total = T if args.scaling == "strong" else T * world
    seq = make_sequence(T=T_total, N=N, P=P, seed=2)
    cano_h = torch.from_numpy(seq["cano"]).pin_memory()
    frames_h = torch.from_numpy(seq["frames"]).pin_memory()
    if args.workload == "cfg4":
        from reart_b200.engine import KinematicEngine
        from reart_b200.synth import kinematic_init
        kw = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in kinematic_init(seq).items()}
        engine = KinematicEngine(kw, torch.from_numpy(seq["part"]), cano_h.to(dev), frames_h.to(dev), ctx=ctx,
                                 use_graph=not args.no_graph)
    else:
        engine = RelaxationEngine(cano_h.to(dev), frames_h.to(dev), num_parts=P, ctx=ctx, use_graph=not args.no_graph)
    lo, hi = engine.frame_range
    frames_local_h = frames_h[lo:hi]
    n_iter = 15000                                            # run_robot.py default; only shapes the tau schedule
    pairs_per_step = 2.0 * T_total * N * N

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(ctx.local_rank, str(torch.cuda.get_device_properties(dev).uuid)) if ctx.is_main else None
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    for i in range(W):
        engine.step(tau_schedule(i, n_iter, 5.0, 1.0))
    torch.cuda.synchronize()

    # ---- timed region: K steps, each bracketed by its own events, L2 flushed in between (untimed)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    ctx.barrier(); torch.cuda.synchronize()
    if sampler:
        sampler.begin()
    t_wall0 = time.perf_counter()
    for i in range(K):
        flush.fill_(i & 0xff)
        engine.tau.fill_(tau_schedule(W + i, n_iter, 5.0, 1.0))
        ev[i][0].record()
        loss = engine.step()
        ev[i][1].record()
    torch.cuda.synchronize(); ctx.barrier()
    wall_s = time.perf_counter() - t_wall0
    if sampler:
        sampler.end()
    clocks = sampler.stop() if sampler else None
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    ctx.all_reduce_max_(tmax)
    total_ms = float(tmax.item())
    ms_per_step = total_ms / K
    value = pairs_per_step / (ms_per_step * 1e-3)
    final_loss = float(loss.item())

    # ---- e2e: host buffers in, loss out, every step (H2D of the step's clouds from pinned memory + D2H of the loss)
    cano_d, frames_d = engine.cano, engine.frames
    h2d = cano_h.numel() * 4 + frames_local_h.numel() * 4
    e2e_ms = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K2 = max(3, min(K, 10))
    for i in range(K2 + 1):
        flush.fill_(i & 0xff)
        torch.cuda.synchronize(); ctx.barrier()
        e0.record()
        cano_d.copy_(cano_h, non_blocking=True)
        frames_d.copy_(frames_local_h, non_blocking=True)
        engine.frames_packed.copy_(ops.pack_cloud(frames_d))          # observed frames arrive fresh: re-pack
        lval = engine.step()
        host_loss = lval.to("cpu", non_blocking=False)                 # D2H read of the step's result
        e1.record(); torch.cuda.synchronize()
        if i > 0:
            e2e_ms += e0.elapsed_time(e1)
    t2 = torch.tensor([e2e_ms / K2], dtype=torch.float64, device=dev)
    ctx.all_reduce_max_(t2)
    e2e_value = pairs_per_step / (float(t2.item()) * 1e-3)
    _ = float(host_loss)

    # ---- roofline of the dominant kernel (chamfer_sym_kernel), timed alone with CUDA events on this stream
    Tl = hi - lo
    keys_a = torch.empty(Tl * N, dtype=torch.int64, device=dev); keys_b = torch.empty(Tl * N, dtype=torch.int64, device=dev)
    skinned = engine.skinned.detach().contiguous()
    def search():
        _lib.check(L.reart_chamfer_sym_search(_lib.ptr(skinned), _lib.ptr(engine.frames_packed), Tl, N, N,
                                              _lib.ptr(keys_a), _lib.ptr(keys_b), None, 0, _lib.stream_ptr()), "search")
    for _ in range(2):
        search()
    reps = 5
    k_ms = 0.0
    for i in range(reps):
        flush.fill_(i)
        e0.record(); search(); e1.record(); torch.cuda.synchronize()
        k_ms += e0.elapsed_time(e1)
    k_ms /= reps
    local_pairs = 2.0 * Tl * N * N
    achieved_tf = FLOP_PER_PAIR * local_pairs / (k_ms * 1e-3) / 1e12
    props = torch.cuda.get_device_properties(dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    peak_tf = props.multi_processor_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    ms_ffma, ops_ffma = ops.fp32_probe(0, iters=2000, device=dev)
    ffma_tf = 2.0 * ops_ffma / (ms_ffma * 1e-3) / 1e12
    alg_bytes = Tl * (12.0 * (N + N) + 8.0 * (N + N))                 # read both clouds once, write both key arrays
    roofline = {"bound": "fp32_fma", "kernel": "chamfer_sym_kernel<8,1>", "achieved": achieved_tf, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "peak_source": f"SMs({props.multi_processor_count}) x 128 lanes x 2 x sm_max_mhz({sm_max_mhz:.0f}) from "
                               "MEASURED_PEAKS.json (it carries no FP32 entry; BASELINE.md section 3)",
                "peak_measured_ffma": ffma_tf, "frac_of_measured_ffma": achieved_tf / ffma_tf,
                "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_per_step,
                "flop_per_pair": FLOP_PER_PAIR, "pairs_per_launch": local_pairs,
                "hbm_achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs"),
                "traffic": None}
    try:        # dram bytes of the same kernel from the committed ncu capture (only valid for the captured shape)
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_sym_traffic.json")))
        if tr.get("workload") == args.workload and tr.get("frames_per_launch") == Tl:
            roofline["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            roofline["traffic_unit"] = "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)"
            roofline["algorithmic_bytes_per_launch"] = alg_bytes
    except Exception:
        pass

    line = None
    if ctx.is_main:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_reference_throughput(T, N, budget_s=12.0)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_dict(args, T_total, N, P, world),
                "iters_per_s": 1e3 / ms_per_step, "wall_s_timed_region": wall_s, "final_loss": final_loss,
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                        "what": "pinned host cano+frames -> H2D -> pack -> full iteration -> loss D2H, per step"},
                # our kernels per step and rank: segmlp fwd, gumbel fwd, rot6d fwd, skin fwd, chamfer_sym, energy rows +
                # columns, skin bwd (w + pose), rot6d bwd, gumbel bwd, segmlp bwd
                # (kinematic model: fk fwd, skin fwd, chamfer_sym, energy rows + columns, skin bwd (w + pose), fk bwd)
                "gpu_launches": (8 if args.workload == "cfg4" else 12) * K, "cuda_graph": not args.no_graph,
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    # teardown: captured graphs reference the NCCL communicator, release them first; with more than one rank
    # leave through os._exit after a final barrier so a slow communicator teardown can never hold the job
    engine.release()
    ctx.barrier()
    if world > 1:
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
