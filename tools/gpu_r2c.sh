set -x
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m | head -12
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -40 > gpurun_out/r2c_multi.txt; tail -40 gpurun_out/r2c_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep -E "DPCHECK|Error|error" | tail -5 > gpurun_out/r2c_dpcheck.txt; cat gpurun_out/r2c_dpcheck.txt
timeout 600 python -m pytest tests/test_srl_classes_gpu.py -m gpu -q --timeout 300 2>&1 | tail -15 > gpurun_out/r2c_plugin.txt; tail -15 gpurun_out/r2c_plugin.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2c_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['trainer_updates_per_sec'], d['roofline']['us_per_update'])"
