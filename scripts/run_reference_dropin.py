"""Drive the UNMODIFIED reference run script (run_robot.py / run_real.py / run_sapien.py) over reart_b200.

    python scripts/run_reference_dropin.py [--backend dropin|stubs] [--script run_robot.py] [--summary out.json] -- \
        --seq_path=<ref>/demo_data/data/nao --save_root=<dir> --cano_idx=2 --evaluate --model=kinematic \
        --resume=<ref>/demo_data/pretrained/nao/kinematic-2/model.pth.tar

TEST INFRASTRUCTURE ONLY (the product is the kernels behind the native module names; this file proves the
"runs unchanged" contract of BASELINE.json's north_star).

``--backend dropin`` (GPU box): ``reart_b200.dropin.install()`` registers ``chamferdist._C``, ``knn_cuda`` and
``pointnet2_cuda`` backed by ``libreart_b200.so``; the reference's own Python (networks/model.py, utils/chamfer.py,
networks/loss.py, utils/kinematic_utils.py, ...) then runs byte-for-byte as shipped, with its hard-imported native
modules answered by our kernels.  Every native entry point is wrapped with a call counter, printed at the end.
``--backend stubs`` (build container, CPU): the torch stand-ins of oracle/ref_harness.py -- this is how the
KAT-E rows in tests/golden/nao_eval.npz were produced.

What is patched, and why it does not touch the path: matplotlib / plotly / imageio / apted /
trimesh are absent from the image -> MagicMock; the three ``vis_*`` functions -> no-ops; ``compute_ted`` -> nan when
apted is absent; ``nx.read_gpickle`` shim (networkx 3); TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD=1 (torch >= 2.6).
"""
from __future__ import annotations

import argparse
import json
import os
import pickle
import runpy
import sys
import time
from unittest.mock import MagicMock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def default_ref_root() -> str:
    staged = os.path.join(ROOT, "baseline", "_ref", "reart")
    if os.path.isdir(os.path.join(staged, "utils")):
        return staged
    return os.environ.get("REART_REFERENCE_ROOT", "/root/reference")


def mock_absent_modules() -> None:
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm", "plotly",
                 "plotly.graph_objects", "plotly.express", "imageio", "apted", "apted.helpers", "trimesh", "kaleido"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock()
    os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")
    import networkx as nx
    if not hasattr(nx, "read_gpickle"):
        nx.read_gpickle = lambda p: pickle.load(open(p, "rb"))


COUNTS: dict = {}


def _counted(name, fn):
    def wrapper(*a, **k):
        COUNTS[name] = COUNTS.get(name, 0) + 1
        return fn(*a, **k)
    wrapper.__name__ = getattr(fn, "__name__", name)
    return wrapper


def install_backend(backend: str) -> None:
    if backend == "dropin":
        from reart_b200 import dropin
        dropin.install(force=True)
        c = sys.modules["chamferdist._C"]
        c.knn_points_idx = _counted("chamferdist._C.knn_points_idx", c.knn_points_idx)
        c.knn_points_backward = _counted("chamferdist._C.knn_points_backward", c.knn_points_backward)
        pn = sys.modules["pointnet2_cuda"]
        pn.furthest_point_sampling_wrapper = _counted("pointnet2_cuda.furthest_point_sampling_wrapper",
                                                      pn.furthest_point_sampling_wrapper)
        pn.ball_query_wrapper = _counted("pointnet2_cuda.ball_query_wrapper", pn.ball_query_wrapper)
        k = sys.modules["knn_cuda"]
        base = k.KNN

        class CountedKNN(base):
            def __call__(self, *a, **kw):
                COUNTS["knn_cuda.KNN.__call__"] = COUNTS.get("knn_cuda.KNN.__call__", 0) + 1
                return super().__call__(*a, **kw)
        CountedKNN.__name__ = "KNN"
        k.KNN = CountedKNN
    else:
        from oracle import ref_harness
        ref_harness.install_stubs()


def neutralise_side_effects() -> None:
    """After the reference modules are importable: visualisation -> no-ops, TED -> nan without apted."""
    import utils.viz_utils as vz
    for n in ("vis_pc", "vis_structure", "vis_pc_seq"):
        setattr(vz, n, lambda *a, **k: None)
    import utils.kinematic_utils as ku
    ku.vis_pc = lambda *a, **k: None
    import utils.ted_utils as td
    if isinstance(sys.modules.get("apted"), MagicMock):
        td.compute_ted = lambda *a, **k: float("nan")


def parse_result_txt(path: str) -> dict:
    out = {}
    if not os.path.exists(path):
        return out
    for line in open(path):
        for part in line.strip().split("|"):
            if ":" in part:
                k, v = part.split(":", 1)
                try:
                    out[k.strip()] = float(v)
                except ValueError:
                    pass
    return out


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="dropin", choices=["dropin", "stubs"])
    ap.add_argument("--script", default="run_robot.py")
    ap.add_argument("--ref-root", default=None)
    ap.add_argument("--summary", default=None)
    ap.add_argument("--fps-from-zero", action="store_true",
                    help="stubs backend only: give the CPU FPS fallback the CUDA kernel's start index 0 (SURVEY Q13)")
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    args = ap.parse_args()
    ref = os.path.abspath(args.ref_root or default_ref_root())
    if not os.path.isdir(os.path.join(ref, "utils")):
        print(f"reference tree not found at {ref} (run scripts/stage_reference.py in the build container)", file=sys.stderr)
        return 2
    rest = [a.replace("{REF}", ref) for a in args.rest if a != "--"]

    mock_absent_modules()
    install_backend(args.backend)
    sys.path.insert(0, ref)
    neutralise_side_effects()
    if args.fps_from_zero and args.backend == "stubs":
        # the reference's CPU fallback starts FPS at torch.randint, its CUDA kernel at index 0 (sampling_gpu.cu:113-115):
        # the structure stage (merging / spanning tree) therefore differs between devices; this reproduces the CUDA order
        from oracle.make_golden import _fps_from_zero
        import networks.pointnet2_utils as pn2
        import utils.graph_utils as gu
        pn2.farthest_point_sample = _fps_from_zero
        gu.farthest_point_sample = _fps_from_zero

    import torch
    save_root = next((a.split("=", 1)[1] for a in rest if a.startswith("--save_root=")), "exp")
    seq = next((a.split("=", 1)[1] for a in rest if a.startswith("--seq_path=")), "data/robot/nao")
    sys.argv = [os.path.join(ref, args.script)] + rest
    t0 = time.time()
    status = "ok"
    try:
        runpy.run_path(os.path.join(ref, args.script), run_name="__main__")
    except Exception as exc:                                   # keep the evidence even if a late stage fails
        import traceback
        traceback.print_exc()
        status = f"{type(exc).__name__}: {exc}"
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    wall = time.time() - t0
    result = parse_result_txt(os.path.join(save_root, seq.rstrip("/").split("/")[-1], "result.txt"))
    loaded = [l.split()[-1] for l in open("/proc/self/maps") if "libreart_b200" in l]
    summary = {"script": args.script, "backend": args.backend, "argv": rest, "status": status, "wall_s": round(wall, 2),
               "cuda": bool(torch.cuda.is_available()), "native_calls": COUNTS, "result_txt": result,
               "native_so_loaded": sorted(set(loaded))}
    print("REFERENCE_DROPIN_SUMMARY " + json.dumps(summary), flush=True)
    if args.summary:
        os.makedirs(os.path.dirname(os.path.abspath(args.summary)), exist_ok=True)
        json.dump(summary, open(args.summary, "w"), indent=1)
    return 0 if status == "ok" else 1


if __name__ == "__main__":
    raise SystemExit(main())
