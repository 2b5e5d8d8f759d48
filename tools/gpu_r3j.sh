set -x
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_image.py > gpurun_out/r3j_${tool}.txt 2>&1; tail -4 gpurun_out/r3j_${tool}.txt
done
SRLX_IMAGE_TC3=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_image.py > gpurun_out/r3j_memcheck_tc3.txt 2>&1; tail -4 gpurun_out/r3j_memcheck_tc3.txt
SRLX_IMAGE_TC3=1 timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_image.py > gpurun_out/r3j_synccheck_tc3.txt 2>&1; tail -4 gpurun_out/r3j_synccheck_tc3.txt
