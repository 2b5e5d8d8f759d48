set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 | tee gpurun_out/r3q_gpu_tests.txt
timeout 400 python tools/image_bench.py --out gpurun_out/r3q_image_bench.json 2>&1 | tail -1 | cut -c1-100
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3q_image_bench.json'))
print({k:round(v,3) if isinstance(v,float) else v for k,v in d['pipeline'].items()})
for k in ('imageq_batch32','imageq_batch256'):
    r=d[k]; print(k, 'u8', round(r['uint8_states']['ms_per_update'],3), 'f32', round(r['float32_states']['ms_per_update'],3), 'fwd', round(r['uint8_states']['ms_per_forward'],3), 'e2e', round(r['uint8_states']['e2e_ms_per_update_host_batches'],3), 'cpu', r.get('cpu_port_ms_per_update'))
PY
