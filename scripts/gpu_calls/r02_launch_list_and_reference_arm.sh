set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -2 gpurun_out/r02_bench_default.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['sustained']['ms_per_step'], d['sustained']['clocks'], d['final_loss'], d['gpu_launches_per_step'])
print(d['culling'])
print(d['cpu_baseline'], d['cpu_baseline_reference_python'])
for e in d['sweep']: print({k:v for k,v in e.items() if k not in ('kernels_per_step','what')})
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-sweep --sustained-s 0 > /dev/null 2>&1
python - <<'PY'
import csv
lines=[l for l in open("gpurun_out/r02_launches_final.csv") if l.startswith('"')]
r=csv.reader(lines); hdr=next(r); rows=list(r)
ki,vi=hdr.index("Kernel Name"),hdr.index("Metric Value")
names=[x[ki] for x in rows]; vals=[float(x[vi].replace(",","")) for x in rows]
sym=[i for i,n in enumerate(names) if "chamfer_sym" in n]
a,b=sym[2],sym[3]
for i in range(a,b): print(f"{vals[i]/1e3:9.1f} us  {names[i][:80]}")
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
