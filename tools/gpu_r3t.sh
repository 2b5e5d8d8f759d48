set -x
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_image.py > gpurun_out/r3t_memcheck.txt 2>&1; tail -3 gpurun_out/r3t_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_image.py > gpurun_out/r3t_racecheck.txt 2>&1; tail -2 gpurun_out/r3t_racecheck.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 | tee gpurun_out/r3t_gpu_tests.txt
