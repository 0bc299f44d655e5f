set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
# full-metric capture of the dominant kernel (one launch) and of the small kernels of a step
ncu --set full --clock-control none --import-source on -k regex:chamfer_sym -s 2 -c 1 -o gpurun_out/r02_sym python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-sweep --sustained-s 0 > gpurun_out/r02_ncu_sym.log 2>&1
ncu -i gpurun_out/r02_sym.ncu-rep --page raw --csv > gpurun_out/r02_sym_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:'energy_|skin_|relax_' -s 24 -c 8 -o gpurun_out/r02_small python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-sweep --sustained-s 0 > gpurun_out/r02_ncu_small.log 2>&1
ncu -i gpurun_out/r02_small.ncu-rep --page raw --csv > gpurun_out/r02_small_raw.csv 2>/dev/null
# launch lists: cold (default) and warm (--cache-control none)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_step.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-sweep --sustained-s 0 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 60 --csv --log-file gpurun_out/r02_launches_step_warm.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-sweep --sustained-s 0 > /dev/null 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -2 gpurun_out/r02_bench_default.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['value'], d['cpu_baseline_reference_python'])
for e in d['sweep']: print({k:v for k,v in e.items() if k!='kernels_per_step'})
"
