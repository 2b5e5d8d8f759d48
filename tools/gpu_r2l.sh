set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_r2d2_gpu.py -m gpu -q --timeout 400 2>&1 | tail -80 > gpurun_out/r2l_r2d2.txt; tail -30 gpurun_out/r2l_r2d2.txt
