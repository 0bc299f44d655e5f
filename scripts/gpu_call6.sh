set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_e.log 2> gpurun_out/r02_bench_e.err; tail -3 gpurun_out/r02_bench_e.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_e.log").read().strip().splitlines()[-1])
print("ms", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches_per_step"], d["gpu_launches_source"])
print("step kernels", d["step_kernels"])
print("sustained", d["sustained"])
print("cpu", d["cpu_baseline"], d["cpu_baseline_reference_python"])
for e in d["sweep"] or []:
    print({k: v for k, v in e.items() if k != "kernels_per_step"})
PY
