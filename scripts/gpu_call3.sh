set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/grad_fp64_analysis.py > gpurun_out/r02_grad_fp64.md 2> gpurun_out/r02_grad_fp64.err; tail -3 gpurun_out/r02_grad_fp64.err; cut -c1-260 gpurun_out/r02_grad_fp64.md
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_b.log 2>&1; tail -1 gpurun_out/r02_bench_b.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_b.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open("gpurun_out/r02_launches_b.csv") if l.startswith('"')]
r=csv.reader(lines); hdr=next(r); rows=list(r)
ki,vi=hdr.index("Kernel Name"),hdr.index("Metric Value")
names=[x[ki] for x in rows]; vals=[float(x[vi].replace(",","")) for x in rows]
sym=[i for i,n in enumerate(names) if "chamfer_sym" in n]
a,b=sym[2],sym[3]
for i in range(a,b): print(f"{vals[i]/1e3:9.1f} us  {names[i][:80]}")
PY
