set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_image_gpu.py -m gpu -q --timeout 300 2>&1 | tail -3 | tee gpurun_out/r3u_image_tests.txt
timeout 400 python bench.py --workload image > gpurun_out/r3u_bench_image.json 2>/dev/null; python -c "import json; d=json.loads(open('gpurun_out/r3u_bench_image.json').read()); print(d['value'], d['roofline']['frac'], d['clocks'], d['e2e']['value'], d['cpu_baseline']['value'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:image_process_staged_kernel -s 5 -c 1 -o gpurun_out/r3u_image_process -f python tools/image_bench.py --no-cpu > gpurun_out/r3u_ncu.log 2>&1; tail -1 gpurun_out/r3u_ncu.log
python tools/ncu_summary.py gpurun_out/r3u_image_process.ncu-rep > gpurun_out/r3u_image_process_ncu_summary.json; cat gpurun_out/r3u_image_process_ncu_summary.json | head -24
timeout 300 python tools/image_bench.py --no-cpu --out gpurun_out/r3u_image_bench.json 2>&1 | tail -1 | cut -c1-560
