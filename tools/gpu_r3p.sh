# 4 GPUs: gpurun --gpus 4 --timeout 1200 -- "bash tools/gpu_r3p.sh"
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 300 2>&1 | tail -4 | tee gpurun_out/r3p_multi_gpu_tests.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/image_dp_check.py 2>&1 | grep IMAGEDP | tee gpurun_out/r3p_image_dp_4gpu.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r3p_bench_4gpu.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3p_bench_4gpu.json').read())
print('own', d['n_gpus'], d['value'], d.get('trainer_updates_per_sec'), d['ms_per_step'], 'e2e', d['e2e']['value'], 'single_learner', d.get('single_learner'))
PY
