set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "presampled" 2>&1 | tail -30 > gpurun_out/r2i_presample.txt; tail -30 gpurun_out/r2i_presample.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -c 300 gpurun_out/r2i_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['trainer_updates_per_sec'], d['roofline']['us_per_update']); print(d['presampled'])"
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep -E "DPCHECK|Error|error" | tail -5 | cut -c1-700
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -8
