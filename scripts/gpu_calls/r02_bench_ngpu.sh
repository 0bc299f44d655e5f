N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --sustained-s 3 --no-sweep > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1])
print($N, d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['sustained']['ms_per_step'], d['final_loss']); c=d['culling']; print('culled', c['ms_per_step'], c['value'], c['fraction_evaluated'], c['final_loss'])"
