ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"knn_window|order_queries|knn_topk_tiled" -c 6 --csv --log-file gpurun_out/r02_flow_blend_ncu.csv python scripts/gpu_flow_blend_timing.py > /dev/null 2>&1
python - <<'PY'
import csv
lines=[l for l in open("gpurun_out/r02_flow_blend_ncu.csv") if l.startswith('"')]
r=csv.reader(lines); hdr=next(r); rows=list(r)
ki,mi,vi=hdr.index("Kernel Name"),hdr.index("Metric Name"),hdr.index("Metric Value")
for x in rows: print(x[ki][:40], x[mi], x[vi])
PY
