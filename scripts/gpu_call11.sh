set -x
timeout 900 python -m pytest tests/test_assign_gpu.py -m gpu -x -q -k engine 2>&1 | grep -E "assert|Error|error|rtol|Mismatch|ACTUAL|DESIRED|passed|failed" | head -30
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 --sustained-s 3 --no-cpu-baseline > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err; tail -2 gpurun_out/r02_bench_f.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_f.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['value'], d['culling'] and d['culling']['fraction_evaluated'], d['roofline']['kernel_ms'], d['sustained']['ms_per_step'], d['final_loss'], d['gpu_launches_per_step'])
for e in d['sweep']: print({k:v for k,v in e.items() if k not in ('kernels_per_step','what')})
"
