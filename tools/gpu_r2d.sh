set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -30 > gpurun_out/r2d_multi.txt; tail -30 gpurun_out/r2d_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep -E "DPCHECK|Error|error" | tail -5 > gpurun_out/r2d_dpcheck.txt; cat gpurun_out/r2d_dpcheck.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d_bench_2gpu.json 2> gpurun_out/r2d_bench_2gpu.err; tail -c 800 gpurun_out/r2d_bench_2gpu.err; python -c "
import json; d=json.loads(open('gpurun_out/r2d_bench_2gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['trainer_updates_per_sec']); print(json.dumps(d['single_learner'], indent=1))"
timeout 300 python -m pytest tests/test_srl_classes_gpu.py -m gpu -q --timeout 300 2>&1 | tail -5
