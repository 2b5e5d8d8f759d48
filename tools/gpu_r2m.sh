set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lstm_seq_fwd_kernel|r2d2_add_kernel|lstm_seq_bwd_kernel" -s 130 -c 3 -o gpurun_out/r2m_seq -f python tools/r2d2_prof.py > gpurun_out/r2m_ncu.log 2>&1; tail -3 gpurun_out/r2m_ncu.log
