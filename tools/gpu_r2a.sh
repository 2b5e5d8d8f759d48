# Round 2, first GPU pass: the new parity tests, the bench with the new step definition, the reference arm, gate exploration.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -40 > gpurun_out/r2a_pytest.txt; tail -15 gpurun_out/r2a_pytest.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -k "drift" -s 2>&1 | grep -E "DRIFT|passed|failed" > gpurun_out/r2a_drift.txt; cat gpurun_out/r2a_drift.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.err; cut -c1-1500 gpurun_out/r2a_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; tail -c 600 gpurun_out/r2a_bench_ref.err; cut -c1-600 gpurun_out/r2a_bench_ref.json
timeout 300 python bench.py --workload dqn --envs 4096 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_dqn.json 2> gpurun_out/r2a_bench_dqn.err; cut -c1-400 gpurun_out/r2a_bench_dqn.json
timeout 600 python tools/explore_grid_gate.py 2>&1 | tail -20
