set -x
mkdir -p gpurun_out
make -C simple_distributed_rl_b200/csrc 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/phase_clocks.py 2>&1 | tail -3 | tee gpurun_out/phase_r1_c.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:learner_fast_kernel -s 1 -c 1 -o gpurun_out/prof_fast_r1_c -f python tools/prof_learner.py > gpurun_out/ncu_c.log 2>&1; tail -3 gpurun_out/ncu_c.log
