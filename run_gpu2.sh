mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "small or uniform or learner_info or many_updates or lockstep" 2>&1 | tail -15
SRLX_LIB=$PWD/tools/ubench/libsrlx_old.so timeout 120 python tools/tree_update_bench.py 2>&1 | tail -6
timeout 120 python tools/tree_update_bench.py 2>&1 | tail -6
timeout 300 python bench.py --workload dqn --envs 4096 --no-cpu-baseline > gpurun_out/bench_dqn_small.json 2> gpurun_out/bench_dqn_small.err; tail -c 400 gpurun_out/bench_dqn_small.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_dqn_small.json').read().strip().splitlines()[-1]); print(d['value'], d['trainer_updates_per_sec'], d['roofline']['us_per_update'], d['roofline']['kernel'], d['e2e']['value'])
PY
