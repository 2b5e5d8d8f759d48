# Full validation on a GPU box: gpurun --timeout 2700 -- "bash tools/gpu_validate.sh" (tests, both bench workloads, SumTree speedtest, ncu capture of learner_small)
set -x
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -6
timeout 400 python bench.py > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; tail -c 300 gpurun_out/bench_i.err
timeout 300 python bench.py --workload dqn --envs 4096 > gpurun_out/bench_i_dqn.json 2> gpurun_out/bench_i_dqn.err; tail -c 300 gpurun_out/bench_i_dqn.err
timeout 300 python tools/sumtree_speedtest.py --out gpurun_out/sumtree_speedtest.json > gpurun_out/sumtree.log 2>&1; tail -4 gpurun_out/sumtree.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'learner_small_kernel' -s 4 -c 1 \
  -o gpurun_out/r1_i_learner_small -f python bench.py --workload dqn --envs 4096 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_small.log 2>&1
python tools/ncu_summary.py gpurun_out/r1_i_learner_small.ncu-rep > gpurun_out/r1_i_learner_small_ncu_summary.json; cat gpurun_out/r1_i_learner_small_ncu_summary.json | head -30
python - <<PY
import json
for f in ('gpurun_out/bench_i.json','gpurun_out/bench_i_dqn.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['trainer_updates_per_sec'], d['roofline']['us_per_update'], d['roofline']['kernel'], d['e2e']['value'], d.get('cpu_baseline',{}) and d['cpu_baseline']['value'])
PY
