set -x
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 600 -rs -k "r2d2 or torchrun" 2>&1 | tail -8
