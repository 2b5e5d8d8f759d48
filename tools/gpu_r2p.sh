set -x
timeout 900 python -m pytest tests/test_invalid_actions_gpu.py tests/test_srl_classes_gpu.py -m gpu -q --timeout 300 -k "wide or window" 2>&1 | grep -v "WARNING" | tail -45
