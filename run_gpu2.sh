mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "small or uniform or learner_info or many_updates or lockstep" 2>&1 | tail -5
PC_WORKLOAD=dqn PC_ENVS=4096 timeout 120 python tools/phase_clocks.py 2>&1 | tail -2
