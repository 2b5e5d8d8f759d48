set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ppo_gpu.py -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/r2k_ppo.txt; tail -40 gpurun_out/r2k_ppo.txt | cut -c1-220
timeout 600 python tools/ppo_bench.py --out gpurun_out/r2k_ppo_bench.json 2>&1 | grep PPOBENCH | cut -c1-900
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-presample > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['trainer_updates_per_sec'], d['roofline']['us_per_update'], d['roofline_rollout'])"
