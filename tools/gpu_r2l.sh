set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_r2d2_gpu.py -m gpu -q --timeout 400 2>&1 | tail -30 > gpurun_out/r2l_r2d2.txt; tail -12 gpurun_out/r2l_r2d2.txt
timeout 600 python tools/r2d2_bench.py --out gpurun_out/r2l_r2d2_bench.json 2>&1 | tail -3 | cut -c1-1500
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l_r2d2_launches.csv python tools/r2d2_prof.py > gpurun_out/r2l_ncu.log 2>&1; tail -3 gpurun_out/r2l_ncu.log
python tools/launch_summary.py gpurun_out/r2l_r2d2_launches.csv 62
