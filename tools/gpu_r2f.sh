set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_qnet_tc_gpu.py -m gpu -q --timeout 120 -x 2>&1 | tail -40 > gpurun_out/r2f_tc.txt; tail -40 gpurun_out/r2f_tc.txt
timeout 300 python tools/qnet_tc_bench.py --out gpurun_out/r2f_qnet_tc_bench.json 2>&1 | tail -12
