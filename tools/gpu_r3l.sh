set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_tc3_gpu.py -m gpu -x -q --timeout 100 2>&1 | tail -5
SRLX_IMAGE_TC3=1 timeout 400 python -m pytest tests/test_image_gpu.py -m gpu -q --timeout 200 -k "reference_trainer or lockstep or gradient" 2>&1 | tail -5
SRLX_IMAGE_TC3=1 timeout 400 python tools/image_bench.py --no-cpu --out gpurun_out/r3l_image_bench_tc3.json 2>&1 | tail -1 | cut -c1-200
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3l_image_bench_tc3.json'))
for k in ('imageq_batch32','imageq_batch256'):
    r=d[k]; print(k, 'u8', round(r['uint8_states']['ms_per_update'],3), 'f32', round(r['float32_states']['ms_per_update'],3), 'fwd', round(r['uint8_states']['ms_per_forward'],3))
PY
