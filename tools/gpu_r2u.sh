set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 tools/r2d2_dp_check.py 2>&1 | grep -E "R2D2DP|Error|error|assert" | cut -c1-700 | tee gpurun_out/r2u_r2d2_dp_4gpu.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 500 2>&1 | tail -5 | tee gpurun_out/r2u_multi_gpu_tests_4gpu.txt
