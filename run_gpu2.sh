for sc in 0 1 2 3; do
SRLX_SCHED=$sc SRLX_LIB=$PWD/simple_distributed_rl_b200/libsrlx_stamps.so timeout 300 python tools/phase_clocks.py 2>&1 | tail -1
SRLX_SCHED=$sc timeout 300 python tools/phase_clocks.py 2>&1 | tail -1 | cut -c1-60
done
