set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-sweep --sustained-s 3 > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
  elif [ $N -eq 8 ]; then timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 5 --sustained-s 5 > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 5 --no-sweep --sustained-s 3 > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err; fi
  tail -1 gpurun_out/r02_scale_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'cull', d['culling'] and d['culling']['fraction_evaluated'], 'sust', d['sustained'] and d['sustained']['ms_per_step'], 'loss', d['final_loss'], 'sweep', d['sweep'])"
done
# brute-force engine (no culling) at 1 and 8 for the scaling table
REART_BENCH_NO_CULL=1 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-sweep --sustained-s 0 > gpurun_out/r02_scale_n1_brute.json 2>/dev/null
REART_BENCH_NO_CULL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 --no-sweep --sustained-s 0 > gpurun_out/r02_scale_n8_brute.json 2>/dev/null
for f in gpurun_out/r02_scale_n1_brute.json gpurun_out/r02_scale_n8_brute.json; do tail -1 $f | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('brute N', d['n_gpus'], 'ms', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'])"; done
