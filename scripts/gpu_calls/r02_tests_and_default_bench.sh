set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -2 gpurun_out/r02_bench_default.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['sustained']['ms_per_step'], d['sustained']['clocks'], d['final_loss'], d['gpu_launches_per_step'])
print(d['culling'])
for e in d['sweep']: print({k:v for k,v in e.items() if k not in ('kernels_per_step','what')})
"
