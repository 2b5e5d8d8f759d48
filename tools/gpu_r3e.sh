# image path: gpurun --timeout 1200 -- "bash tools/gpu_r3e.sh"
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_image_gpu.py -m gpu -q --timeout 200 2>&1 | tail -120 | tee gpurun_out/r3e_image_tests.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3e_imageq_launches.csv python tools/image_prof.py 32 > gpurun_out/r3e_ncu.log 2>&1; tail -2 gpurun_out/r3e_ncu.log
python tools/launch_summary.py gpurun_out/r3e_imageq_launches.csv 55
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3e_imageq_launches256.csv python tools/image_prof.py 256 > gpurun_out/r3e_ncu256.log 2>&1
python tools/launch_summary.py gpurun_out/r3e_imageq_launches256.csv 55
timeout 400 python tools/image_bench.py --out gpurun_out/r3e_image_bench.json 2>&1 | tail -3 | cut -c1-2500
