set -x
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -6
timeout 400 python bench.py > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; tail -c 600 gpurun_out/bench_h.json
timeout 300 python bench.py --workload dqn --envs 4096 --no-cpu-baseline > gpurun_out/bench_dqn.json 2> gpurun_out/bench_dqn.err; tail -c 1500 gpurun_out/bench_dqn.json
timeout 300 python tools/sumtree_speedtest.py --out gpurun_out/sumtree_speedtest.json 2>&1 | tail -5
# ncu --set full: the rollout-side kernels of the headline workload (first launches after the ring is full)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rollout_kernel|post_step_kernel' -s 300 -c 2 \
  -o gpurun_out/r1_h_rollout -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rollout.log 2>&1
python tools/ncu_summary.py gpurun_out/r1_h_rollout.ncu-rep > gpurun_out/r1_h_rollout_ncu_summary.json; cat gpurun_out/r1_h_rollout_ncu_summary.json | head -60
