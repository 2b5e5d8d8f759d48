mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_j_2gpu.json 2> gpurun_out/bench_j_2gpu.err; tail -c 300 gpurun_out/bench_j_2gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_j_2gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['trainer_updates_per_sec'], d['ms_per_step'], d['e2e']['value'], d['roofline']['traffic_source'])
PY
