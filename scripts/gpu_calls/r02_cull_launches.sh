ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chamfer_sym|cull_|energy_|skin_|relax_" -c 200 --csv --log-file gpurun_out/r02_cull_launches.csv python scripts/gpu_cull_launches.py > /dev/null 2>&1
python - <<'PY'
import csv
lines=[l for l in open("gpurun_out/r02_cull_launches.csv") if l.startswith('"')]
r=csv.reader(lines); hdr=next(r); rows=list(r)
ki,vi=hdr.index("Kernel Name"),hdr.index("Metric Value")
names=[x[ki] for x in rows]; vals=[float(x[vi].replace(",","")) for x in rows]
sym=[i for i,n in enumerate(names) if "chamfer_sym" in n]
a,b=sym[-2],sym[-1]
for i in range(a,b): print(f"{vals[i]/1e3:9.1f} us  {names[i][:70]}")
PY
