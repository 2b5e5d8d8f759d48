timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1_g.json | cut -c1-300
SRLX_L2_PERSIST=0 timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
