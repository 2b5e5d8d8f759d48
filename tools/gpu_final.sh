# Final validation of a round on one GPU box: gpurun --timeout 2700 -- "bash tools/gpu_final.sh <tag>"
set -x
TAG=${1:-r2q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -12 > gpurun_out/${TAG}_gpu_tests.txt; cat gpurun_out/${TAG}_gpu_tests.txt
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 300 gpurun_out/${TAG}_bench_reference.err
timeout 400 python bench.py --workload dqn --envs 4096 > gpurun_out/${TAG}_bench_dqn.json 2> gpurun_out/${TAG}_bench_dqn.err; tail -c 300 gpurun_out/${TAG}_bench_dqn.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-presample > gpurun_out/${TAG}_ncu_launches.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_launches.log | cut -c1-300
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launch_summary.txt; head -12 gpurun_out/${TAG}_launch_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'learner_fast_kernel' -s 6 -c 1 -o gpurun_out/${TAG}_learner -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-presample > gpurun_out/${TAG}_ncu_learner.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_learner.log | cut -c1-300
python tools/ncu_summary.py gpurun_out/${TAG}_learner.ncu-rep > gpurun_out/${TAG}_learner_ncu_summary.json; head -c 1500 gpurun_out/${TAG}_learner_ncu_summary.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'learner_small_kernel' -s 4 -c 1 -o gpurun_out/${TAG}_learner_small -f python bench.py --workload dqn --envs 4096 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_small.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_learner_small.ncu-rep > gpurun_out/${TAG}_learner_small_ncu_summary.json
timeout 300 python tools/sumtree_speedtest.py --skip python --out gpurun_out/${TAG}_sumtree_speedtest.json > gpurun_out/${TAG}_sumtree.log 2>&1; tail -3 gpurun_out/${TAG}_sumtree.log | cut -c1-250
timeout 600 python tools/r2d2_bench.py --cpu-baseline 2 --out gpurun_out/${TAG}_r2d2_bench.json 2>&1 | tail -2 | cut -c1-1500
timeout 400 python tools/image_bench.py --out gpurun_out/${TAG}_image_bench.json 2>&1 | tail -1 | cut -c1-300
timeout 400 python bench.py --workload image > gpurun_out/${TAG}_bench_image.json 2>/dev/null; cut -c1-300 gpurun_out/${TAG}_bench_image.json
timeout 400 python bench.py --workload image --impl reference > gpurun_out/${TAG}_bench_image_reference.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_image_reference.json
python - <<PY
import json
for f in ('gpurun_out/${TAG}_bench.json','gpurun_out/${TAG}_bench_dqn.json','gpurun_out/${TAG}_bench_reference.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('us_per_update'), (d.get('roofline') or {}).get('traffic_source'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
