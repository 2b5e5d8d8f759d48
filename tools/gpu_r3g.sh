set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_tc3_gpu.py -m gpu -x -q --timeout 100 2>&1 | tail -30 | tee gpurun_out/r3g_tc3_tests.txt
timeout 500 python -m pytest tests/test_image_gpu.py -m gpu -q --timeout 200 2>&1 | tail -60 | tee gpurun_out/r3g_image_tests.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3g_imageq_launches.csv python tools/image_prof.py 32 > gpurun_out/r3g_ncu.log 2>&1; tail -2 gpurun_out/r3g_ncu.log
python tools/launch_summary.py gpurun_out/r3g_imageq_launches.csv 55
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3g_imageq_launches256.csv python tools/image_prof.py 256 > gpurun_out/r3g_ncu256.log 2>&1
python tools/launch_summary.py gpurun_out/r3g_imageq_launches256.csv 55
timeout 400 python tools/image_bench.py --no-cpu --out gpurun_out/r3g_image_bench.json 2>&1 | tail -3 | cut -c1-3000
