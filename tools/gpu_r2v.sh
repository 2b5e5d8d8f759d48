set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppo_update_kernel -s 1 -c 1 -o gpurun_out/r2v_ppo_upd -f python tools/ppo_bench.py --rollouts 1 --updates 2000 > gpurun_out/r2v_ncu.log 2>&1; tail -2 gpurun_out/r2v_ncu.log
