set -x
mkdir -p gpurun_out
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2w_bench_${N}gpu.json 2> gpurun_out/r2w_bench_${N}gpu.err; tail -c 300 gpurun_out/r2w_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2w_bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','trainer_updates_per_sec','gpu_launches')}); print(d.get('single_learner')); print(d.get('e2e'))
PY
