set -x
timeout 900 python -m pytest tests/test_invalid_actions_gpu.py -m gpu -q --timeout 300 2>&1 | tail -40
