set -x
mkdir -p gpurun_out
SRLX_IMAGE_TC3=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3m_tc3_launches256.csv python tools/image_prof.py 256 > gpurun_out/r3m_ncu256.log 2>&1
python tools/launch_summary.py gpurun_out/r3m_tc3_launches256.csv 47
SRLX_IMAGE_TC3=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3m_tc3_launches32.csv python tools/image_prof.py 32 > gpurun_out/r3m_ncu32.log 2>&1
python tools/launch_summary.py gpurun_out/r3m_tc3_launches32.csv 55
