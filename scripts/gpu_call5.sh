set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_d.log 2>&1; tail -1 gpurun_out/r02_bench_d.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['final_loss'])"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg3_16k_8f > gpurun_out/r02_bench_d8f.log 2>&1; tail -1 gpurun_out/r02_bench_d8f.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dist_equiv_check.py 2>&1 | grep -v "^W\|^\*" | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.log 2>&1; tail -1 gpurun_out/r02_bench_n2.log | cut -c1-200
REART_ONESHOT_ALLREDUCE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_nccl.log 2>&1; tail -1 gpurun_out/r02_bench_n2_nccl.log | cut -c1-200
