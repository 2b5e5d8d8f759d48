# image path, first GPU run: gpurun --timeout 1200 -- "bash tools/gpu_r3a.sh"
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_image_gpu.py -m gpu -x -q --timeout 200 2>&1 | tail -30 | tee gpurun_out/r3a_image_tests.txt
timeout 400 python -m pytest tests/test_r2d2_gpu.py -m gpu -x -q --timeout 200 2>&1 | tail -4 | tee gpurun_out/r3a_r2d2_tests.txt
timeout 400 python tools/image_bench.py --out gpurun_out/r3a_image_bench.json 2>&1 | tail -3 | cut -c1-2500
