set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python scripts/grad_fp64_analysis.py > gpurun_out/r02_grad_fp64.md 2> gpurun_out/r02_grad_fp64.err; tail -3 gpurun_out/r02_grad_fp64.err; cat gpurun_out/r02_grad_fp64.md
python scripts/gpu_kernel_zoo.py > gpurun_out/r02_zoo.log 2>&1; tail -20 gpurun_out/r02_zoo.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_a.log 2>&1; tail -2 gpurun_out/r02_bench_a.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_a.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
python scripts/launch_shares.py gpurun_out/r02_launches_a.csv 24
