mkdir -p gpurun_out/san
timeout 1100 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -5
export SAN_UPDATES=3
for tool in memcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py > gpurun_out/san/r1_i_$tool.txt 2>&1
  tail -8 gpurun_out/san/r1_i_$tool.txt | cut -c1-200
done
PC_WORKLOAD=dqn512 PC_ENVS=4096 timeout 120 python tools/phase_clocks.py 2>&1 | tail -1 | cut -c1-300
PC_WORKLOAD=dqn PC_ENVS=4096 timeout 120 python tools/phase_clocks.py 2>&1 | tail -1 | cut -c1-400
