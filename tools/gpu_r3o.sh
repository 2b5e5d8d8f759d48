set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_image_gpu.py -m gpu -q --timeout 300 -k "processor" 2>&1 | tail -5 | tee gpurun_out/r3o_image_tests.txt
timeout 400 python tools/image_bench.py --no-cpu --out gpurun_out/r3o_image_bench.json 2>&1 | tail -1 | cut -c1-700
timeout 600 ncu --set full --clock-control none --import-source on -k regex:image_process_staged_kernel -s 5 -c 1 -o gpurun_out/r3o_image_process -f python tools/image_bench.py --no-cpu > gpurun_out/r3o_ncu.log 2>&1; tail -2 gpurun_out/r3o_ncu.log
python tools/ncu_summary.py gpurun_out/r3o_image_process.ncu-rep > gpurun_out/r3o_image_process_ncu_summary.json; head -c 1200 gpurun_out/r3o_image_process_ncu_summary.json
