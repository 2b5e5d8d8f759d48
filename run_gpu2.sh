mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "small or uniform or learner_info or many_updates or lockstep or row_split" 2>&1 | tail -15
PC_WORKLOAD=dqn PC_ENVS=4096 timeout 120 python tools/phase_clocks.py 2>&1 | tail -1
SRLX_SMALL_CLUSTER=8 PC_WORKLOAD=dqn PC_ENVS=4096 timeout 120 python tools/phase_clocks.py 2>&1 | tail -1
SRLX_SMALL_CLUSTER=2 PC_WORKLOAD=dqn PC_ENVS=4096 timeout 120 python tools/phase_clocks.py 2>&1 | tail -1
