set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_srl_classes_gpu.py -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/r2b_plugin.txt; tail -60 gpurun_out/r2b_plugin.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -30 > gpurun_out/r2b_pytest.txt; tail -30 gpurun_out/r2b_pytest.txt
