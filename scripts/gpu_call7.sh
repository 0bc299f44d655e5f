set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python scripts/gpu_lap_timing.py 2>&1 | tail -8
