set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c.log 2>&1; tail -1 gpurun_out/r02_bench_c.log | cut -c1-300
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg3_16k_8f > gpurun_out/r02_bench_c8f.log 2>&1; tail -1 gpurun_out/r02_bench_c8f.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_c.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open("gpurun_out/r02_launches_c.csv") if l.startswith('"')]
r=csv.reader(lines); hdr=next(r); rows=list(r)
ki,vi=hdr.index("Kernel Name"),hdr.index("Metric Value")
names=[x[ki] for x in rows]; vals=[float(x[vi].replace(",","")) for x in rows]
sym=[i for i,n in enumerate(names) if "chamfer_sym" in n]
a,b=sym[2],sym[3]
for i in range(a,b): print(f"{vals[i]/1e3:9.1f} us  {names[i][:80]}")
PY
