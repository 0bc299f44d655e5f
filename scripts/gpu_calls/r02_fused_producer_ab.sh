run() { python bench.py --workload $1 --no-sweep --sustained-s 2 --no-cpu-baseline 2>gpurun_out/err_$1.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', round(d['ms_per_step'],4), round(d['sustained']['ms_per_step'],4), d['roofline']['kernel_ms'], d['gpu_launches_per_step'])" || tail -5 gpurun_out/err_$1.txt; }
for w in cfg3_16k cfg3_16k_8f cfg3_4k; do
  REART_NO_FUSED_PRODUCER=1 run $w pipelined
  run $w fused
  REART_NO_FUSED_PRODUCER=1 run $w pipelined_again
  run $w fused_again
done
