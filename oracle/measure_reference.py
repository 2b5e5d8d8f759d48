"""BASELINE TOOLING (not part of the product; needs the reference importable, i.e. the build container):
times the UNMODIFIED reference loop -- srl.Runner(...).train() = core_play.play, srl/base/run/core_play.py:115-214 -- on the
bench workload's algorithm (Rainbow: double + dueling(512,) + NoisyNet + 3-step + proportional PER) and on DQN, on the
restated CartPole-v1 / the reference's Grid, next to the oracle's sequential port of the vectorised loop (the `cpu_baseline`
of bench.py).  Protocol of BASELINE.md section 2/4: device CPU, memory.compress = False, warm-up train(), then a timed
train(max_steps=N); rates = state.total_step / dt, state.train_count / dt.

    PYTHONPATH=/root/reference python -m oracle.measure_reference [--steps N] [--threads T]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--threads", type=int, default=1)
    args = ap.parse_args()
    import torch

    torch.set_num_threads(args.threads)
    import srl
    from srl.algorithms import dqn, rainbow

    from oracle.ref_envs import register_restated_envs

    register_restated_envs()
    out = {"host_cores": os.cpu_count(), "torch_threads": args.threads, "steps": args.steps, "runs": []}

    def run(name, env, cfg):
        cfg.memory.warmup_size = 1000
        cfg.memory.compress = False
        runner = srl.Runner(env, cfg)
        runner.set_device("CPU")
        runner.set_seed(1)
        runner.train(max_steps=1200, enable_progress=False)  # warm-up (fills the memory past warmup_size)
        t0 = time.perf_counter()
        st = runner.train(max_steps=args.steps, enable_progress=False)
        dt = time.perf_counter() - t0
        out["runs"].append({"what": name, "env_steps_per_s": st.total_step / dt, "updates_per_s": st.train_count / dt,
                            "seconds": dt})
        print(out["runs"][-1], flush=True)

    c = rainbow.Config(multisteps=3, enable_noisy_dense=True, enable_double_dqn=True)  # hidden block default: dueling (512,)
    c.memory.set_proportional()
    run("reference Runner.train: Rainbow(double+dueling512+noisy+3step+PER python SumTree), CartPole-v1 restated", "CartPole-v1", c)
    c = rainbow.Config(multisteps=3, enable_noisy_dense=True, enable_double_dqn=True)
    c.memory.set_proportional()
    run("reference Runner.train: same Rainbow, Grid", "Grid", c)
    c = dqn.Config()
    c.hidden_block.set((64, 64))
    run("reference Runner.train: DQN MLP[64,64] uniform replay, CartPole-v1 restated", "CartPole-v1", c)

    # the oracle's sequential port of the vectorised loop (bench.py cpu_baseline), 1 update per env step like the reference
    sys.argv = [sys.argv[0]]
    import bench

    r = bench.cpu_port_run(64, 1, steps=10_000, warmup=1, budget_s=15.0, threads=args.threads)
    out["runs"].append({"what": "oracle port (bench.py cpu_baseline), train_interval 1: " + r["sample"],
                        "env_steps_per_s": r["env_steps_per_s"], "updates_per_s": r["updates_per_s"], "seconds": r["seconds"]})
    print(out["runs"][-1], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
