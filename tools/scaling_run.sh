# one node, N GPUs: both arms of bench.py under torchrun, as the driver launches them (usage: bash tools/scaling_run.sh N)
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref_${N}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_${N}gpu.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu.json').read()); r=json.loads(open('gpurun_out/bench_ref_${N}gpu.json').read())
print('own', d['n_gpus'], d['value'], d['trainer_updates_per_sec'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
print('ref', r['value'], r['cpu_baseline']['cores'])
PY
