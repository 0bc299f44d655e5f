set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
for N in 1 2 4; do
  if [ $N -eq 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-sweep --sustained-s 3 > gpurun_out/r02_final_n1.json 2> gpurun_out/r02_final_n1.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N bench.py --gpus $N --steps 20 --warmup 5 --no-sweep --sustained-s 3 > gpurun_out/r02_final_n$N.json 2> gpurun_out/r02_final_n$N.err; fi
  tail -1 gpurun_out/r02_final_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'sust', d['sustained'] and d['sustained']['ms_per_step'], 'loss', d['final_loss'], 'culled', d['culling'] and d['culling'].get('ms_per_step'))"
done
