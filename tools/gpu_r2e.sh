set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rankbased_gpu.py -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/r2e_rank.txt; tail -40 gpurun_out/r2e_rank.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -10 > gpurun_out/r2e_multi.txt; tail -10 gpurun_out/r2e_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep -E "DPCHECK|Error|error" | tail -5 > gpurun_out/r2e_dpcheck.txt; cut -c1-700 gpurun_out/r2e_dpcheck.txt
