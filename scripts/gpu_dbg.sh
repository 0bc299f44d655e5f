python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chamfer or Chamfer or fused" 2>&1 | grep -v "^$" | tail -40
