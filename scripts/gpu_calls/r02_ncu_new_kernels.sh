# ncu --set full of the kernels added late in round 2: the decoupled culled search + its bounds, the warp-sorted skin forward
ncu --set full --clock-control none -k regex:'chamfer_sym_cull|cull_row|cull_col|skin_fwd_sorted8' -s 27 -c 4 -o gpurun_out/r02_cullk python scripts/gpu_cull_launches.py > gpurun_out/r02_ncu_cullk.log 2>&1
ncu -i gpurun_out/r02_cullk.ncu-rep --page raw --csv > gpurun_out/r02_cullk_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:'knn_window|knnw_order' -s 2 -c 2 -o gpurun_out/r02_blend python scripts/gpu_flow_blend_timing.py > gpurun_out/r02_ncu_blend.log 2>&1
ncu -i gpurun_out/r02_blend.ncu-rep --page raw --csv > gpurun_out/r02_blend_raw.csv 2>/dev/null
python - <<'PY'
import csv
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
for f in ("gpurun_out/r02_cullk_raw.csv", "gpurun_out/r02_blend_raw.csv"):
    rows = list(csv.reader(open(f)))
    hdr = rows[0]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(r[ki][:60])
        for w in want:
            if w in hdr: print("   ", w, r[hdr.index(w)], rows[1][hdr.index(w)])
PY
