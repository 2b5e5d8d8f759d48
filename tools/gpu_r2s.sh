set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_r2.py r2d2 seam masks ppo rank > gpurun_out/r2s_racecheck.txt 2>&1; tail -6 gpurun_out/r2s_racecheck.txt
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_r2.py ppo rank > gpurun_out/r2s_memcheck2.txt 2>&1; tail -4 gpurun_out/r2s_memcheck2.txt
