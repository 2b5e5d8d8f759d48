set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -6
SRLX_LIB=$PWD/simple_distributed_rl_b200/libsrlx_stamps.so timeout 300 python tools/phase_clocks.py 2>&1 | tail -1
timeout 300 python tools/phase_clocks.py 2>&1 | tail -1
