set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/phase_clocks.py 2>&1 | tail -3 | tee gpurun_out/phase_r1_e.json
