set -x
timeout 900 python -m pytest tests/test_r2d2_gpu.py -m gpu -q --timeout 300 -k "max_steps" 2>&1 | tail -30
