set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/ppo_bench.py --out gpurun_out/r2y_ppo_bench_${N}gpu.json 2>&1 | grep -E "PPOBENCH|Error|error" | cut -c1-900
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 tools/r2d2_dp_check.py 2>&1 | grep -E "R2D2DP|Error|error|assert" | cut -c1-700 | tee gpurun_out/r2y_r2d2_dp_${N}gpu.txt
