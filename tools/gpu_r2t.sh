set -x
mkdir -p gpurun_out
timeout 600 python tools/ppo_bench.py --out gpurun_out/r2t_ppo_bench.json 2>&1 | tail -3 | cut -c1-900
