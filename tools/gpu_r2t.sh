set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ppo_gpu.py -m gpu -q --timeout 400 2>&1 | tail -3
timeout 600 python tools/ppo_bench.py --out gpurun_out/r2t_ppo_bench.json 2>&1 | tail -2 | cut -c1-900
