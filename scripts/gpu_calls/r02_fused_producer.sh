set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
run() { python bench.py --workload $1 --no-sweep --sustained-s 2 --no-cpu-baseline 2>gpurun_out/err_$1.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],4), round(d['sustained']['ms_per_step'],4), d['roofline']['kernel_ms'], d['final_loss'], d['gpu_launches_per_step'], d['step_kernels'])" || tail -5 gpurun_out/err_$1.txt; }
run cfg3_16k
run cfg3_16k_8f
