set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 600 python scripts/gpu_lap_timing.py 2>&1 | tail -7
