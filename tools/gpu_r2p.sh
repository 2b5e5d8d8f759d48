set -x
timeout 900 python -m pytest tests/test_r2d2_gpu.py -m gpu -q --timeout 600 -k "full_size" 2>&1 | tail -30
