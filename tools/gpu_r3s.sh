set -x
mkdir -p gpurun_out
timeout 400 python bench.py --workload image > gpurun_out/r3s_bench_image.json 2> gpurun_out/r3s_bench_image.err; tail -c 300 gpurun_out/r3s_bench_image.err; cut -c1-1800 gpurun_out/r3s_bench_image.json
timeout 400 python bench.py --workload image --impl reference > gpurun_out/r3s_bench_image_reference.json 2>/dev/null; cut -c1-700 gpurun_out/r3s_bench_image_reference.json
timeout 400 python -m pytest tests/test_image_gpu.py -m gpu -q -k "learns" 2>&1 | tail -3
