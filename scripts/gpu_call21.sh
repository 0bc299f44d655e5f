set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "blend or flow or topk" 2>&1 | tail -15
timeout 300 python scripts/gpu_flow_blend_timing.py 2>&1 | tail -8
