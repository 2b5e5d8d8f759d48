set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ppo_gpu.py -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/r2j_ppo.txt; tail -60 gpurun_out/r2j_ppo.txt | cut -c1-220
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_srl_classes_gpu.py tests/test_rankbased_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -8
