set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --sustained-s 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -3 gpurun_out/r02_bench_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['sustained'], d['final_loss'], d['gpu_launches_per_step'], d['config'])
for e in d.get('sweep',[]): print({k:v for k,v in e.items() if k not in ('kernels_per_step',)})
"
