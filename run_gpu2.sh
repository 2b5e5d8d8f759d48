mkdir -p gpurun_out/san
export SAN_UPDATES=3
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py > gpurun_out/san/r1_i_$tool.txt 2>&1
  tail -8 gpurun_out/san/r1_i_$tool.txt | cut -c1-200
done
PC_WORKLOAD=dqn512 PC_ENVS=4096 timeout 120 python tools/phase_clocks.py 2>&1 | tail -1 | cut -c1-300
SRLX_LIB=$PWD/simple_distributed_rl_b200/libsrlx_stamps.so PC_WORKLOAD=dqn512 PC_ENVS=4096 timeout 120 python tools/phase_clocks.py 2>&1 | tail -1
