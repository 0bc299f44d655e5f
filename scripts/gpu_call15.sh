set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
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
for wl in cfg3_16k_8f cfg3_16k; do python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-sweep --sustained-s 0 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['config']['T'], 'ms', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'culled', d['culling'] and d['culling'].get('ms_per_step'))"; done
