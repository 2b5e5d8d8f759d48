set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l_r2d2_launches.csv python tools/r2d2_prof.py > gpurun_out/r2l_ncu.log 2>&1; tail -3 gpurun_out/r2l_ncu.log
python tools/launch_summary.py gpurun_out/r2l_r2d2_launches.csv 62
