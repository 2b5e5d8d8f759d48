set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
PROF_UPDATES=256 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"learner_fast_kernel|noise_precompute" -s 2 -c 2 -o gpurun_out/prof_fast_r1_d -f python tools/prof_learner.py > gpurun_out/ncu_d.log 2>&1; tail -2 gpurun_out/ncu_d.log
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r1_d.json
