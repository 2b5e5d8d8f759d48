set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_image_gpu.py -m gpu -q --timeout 300 2>&1 | tail -3 | tee gpurun_out/r3v_image_tests.txt
timeout 400 python tools/image_bench.py --no-cpu --out gpurun_out/r3v_image_bench.json 2>&1 | tail -1 | cut -c1-80
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3v_image_bench.json'))
for k in ('imageq_batch32','imageq_batch256'):
    r=d[k]; print(k, 'u8', round(r['uint8_states']['ms_per_update'],3), 'f32', round(r['float32_states']['ms_per_update'],3), 'fwd', round(r['uint8_states']['ms_per_forward'],3), 'e2e', round(r['uint8_states']['e2e_ms_per_update_host_batches'],3))
PY
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_image.py > gpurun_out/r3v_racecheck.txt 2>&1; tail -2 gpurun_out/r3v_racecheck.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_image.py > gpurun_out/r3v_memcheck.txt 2>&1; tail -1 gpurun_out/r3v_memcheck.txt
