set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 70 --csv --log-file gpurun_out/r02_launches_8f.csv python bench.py --workload cfg3_16k_8f --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-sweep --sustained-s 0 > /dev/null 2>&1
python - <<'PY'
import csv
lines=[l for l in open("gpurun_out/r02_launches_8f.csv") if l.startswith('"')]
r=csv.reader(lines); hdr=next(r); rows=list(r)
ki,vi=hdr.index("Kernel Name"),hdr.index("Metric Value")
names=[x[ki] for x in rows]; vals=[float(x[vi].replace(",","")) for x in rows]
sym=[i for i,n in enumerate(names) if "chamfer_sym" in n]
a,b=sym[2],sym[3]
for i in range(a,b): print(f"{vals[i]/1e3:9.1f} us  {names[i][:80]}")
PY
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err; tail -2 gpurun_out/r02_bench_g.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_g.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['sustained']['ms_per_step'], d['final_loss'], d['gpu_launches_per_step'])
print(d['culling'])
print(d['cpu_baseline'], d['cpu_baseline_reference_python'])
for e in d['sweep']: print({k:v for k,v in e.items() if k not in ('kernels_per_step','what')})
"
