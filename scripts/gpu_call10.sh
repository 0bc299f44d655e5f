set -x
timeout 900 python -m pytest tests/test_cull_gpu.py -m gpu -x -q 2>&1 | tail -25
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
CUDA_LAUNCH_BLOCKING=1 python bench.py --workload cfg5 --steps 5 --no-sweep --sustained-s 0 --no-cpu-baseline 2>&1 | tail -4 | cut -c1-600
python bench.py --steps 20 --warmup 5 --no-sweep --sustained-s 3 --no-cpu-baseline > gpurun_out/r02_bench_cull.json 2> gpurun_out/r02_bench_cull.err; tail -2 gpurun_out/r02_bench_cull.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_cull.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['value'], d['culling'], d['roofline']['kernel_ms'], d['sustained']['ms_per_step'], d['final_loss'])
"
