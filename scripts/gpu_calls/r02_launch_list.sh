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
