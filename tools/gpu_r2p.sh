set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "lockstep" 2>&1 | tail -8
