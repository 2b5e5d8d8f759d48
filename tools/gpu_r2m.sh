set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sgemm_mma_kernel" -s 4 -c 3 -o gpurun_out/r2m_gemm -f python tools/r2d2_prof.py > gpurun_out/r2m_ncu.log 2>&1; tail -2 gpurun_out/r2m_ncu.log
