set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python scripts/gpu_kernel_zoo.py 2>&1 | grep "^|" | sort -u | cut -c1-200
