set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 60 2>&1 | tail -12
timeout 300 python tools/phase_clocks.py 2>&1 | tail -1 | tee gpurun_out/phase_r1_i.json
SRLX_L2_PERSIST=0 timeout 300 python tools/phase_clocks.py 2>&1 | tail -1
SRLX_CLUSTER=16 timeout 300 python tools/phase_clocks.py 2>&1 | tail -1
