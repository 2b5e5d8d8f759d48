set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 200 2>&1 | tail -8 > gpurun_out/r2h_multi.txt; tail -8 gpurun_out/r2h_multi.txt
for n in 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/dp_check.py 2>&1 | grep -E "DPCHECK|Error|error" | tail -3 > gpurun_out/r2h_dpcheck_$n.txt; cut -c1-800 gpurun_out/r2h_dpcheck_$n.txt
done
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2h_bench_${n}gpu.json 2> gpurun_out/r2h_bench_${n}gpu.err; tail -c 300 gpurun_out/r2h_bench_${n}gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2h_bench_${n}gpu.json').read().strip().splitlines()[-1]); print($n, d['value'], d['trainer_updates_per_sec'], d['ms_per_step']); print(d['single_learner'])"
done
