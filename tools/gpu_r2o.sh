set -x
mkdir -p gpurun_out
SRLX_LIB=$PWD/simple_distributed_rl_b200/libsrlx_stamps.so timeout 300 python tools/phase_clocks.py > gpurun_out/r2o_phase_clocks.json 2>gpurun_out/r2o_pc.err; tail -2 gpurun_out/r2o_pc.err
SRLX_LIB=$PWD/simple_distributed_rl_b200/libsrlx_stamps.so PC_PRESAMPLE=1 timeout 300 python tools/phase_clocks.py > gpurun_out/r2o_phase_clocks_presample.json 2>gpurun_out/r2o_pc2.err; tail -2 gpurun_out/r2o_pc2.err
cat gpurun_out/r2o_phase_clocks.json gpurun_out/r2o_phase_clocks_presample.json
