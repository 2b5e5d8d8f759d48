"""PPO on BASELINE configs[4] (Pendulum-v1 continuous, 16 384 env copies per GPU, GAE lambda = 0.95, 200-step on-policy rollout buffer):
env-steps/s of the rollout, of the returns pass, and the cost of one minibatch update; under torchrun every rank runs its own env copies
and the parameters are averaged over NCCL after every rollout's updates (the reference's PPO refuses its distributed mode, ppo.py:295-297).

    python tools/ppo_bench.py [--out gpurun_out/ppo_bench.json]            (or under torchrun)
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_distributed_rl_b200 import parallel  # noqa: E402
from simple_distributed_rl_b200.ppo import PPOConfig, PPORunner  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--rollouts", type=int, default=3)
    ap.add_argument("--updates", type=int, default=20000)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = PPOConfig(env="Pendulum-v1", n_envs=args.envs, horizon=200, gae_discount=0.95, seed=1 + rank, lr_decay_steps=0)
    r = PPORunner(cfg, device=dev)
    eng = r.engine
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    t_roll = t_ret = t_upd = 0.0
    eng.rollout(); eng.finish_rollout(); eng.learn(200)  # warm-up
    torch.cuda.synchronize(dev)
    for _ in range(args.rollouts):
        a, b, c, d = ev(), ev(), ev(), ev()
        a.record(); eng.rollout(); b.record(); eng.finish_rollout(); c.record(); eng.learn(args.updates)
        if world > 1:
            parallel.average_parameters([eng.t["params"]])
        d.record()
        torch.cuda.synchronize(dev)
        t_roll += a.elapsed_time(b); t_ret += b.elapsed_time(c); t_upd += c.elapsed_time(d)
    tt = torch.tensor([t_roll, t_ret, t_upd], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_roll, t_ret, t_upd = [float(x) for x in tt.tolist()]
    n_steps = args.rollouts * 200 * args.envs * world
    st = eng.read_state()
    out = {"config": "PPO Pendulum-v1 continuous, GAE 0.95, horizon 200 (BASELINE configs[4])", "n_gpus": world, "envs_per_gpu": args.envs,
           "rollout_env_steps_per_s": n_steps / (t_roll * 1e-3), "rollout_ms_per_vector_step": t_roll / (args.rollouts * 200),
           "returns_ms_per_rollout": t_ret / args.rollouts, "us_per_update": 1e3 * t_upd / (args.rollouts * args.updates),
           "updates_per_s_per_gpu": args.rollouts * args.updates / (t_upd * 1e-3),
           "env_steps_per_s_with_updates": n_steps / ((t_roll + t_ret + t_upd) * 1e-3), "updates_per_rollout": args.updates,
           "mean_episode_reward": st.episode_reward_sum / max(1, st.episode_count)}
    if rank == 0:
        print("PPOBENCH " + json.dumps(out), flush=True)
        if args.out:
            json.dump(out, open(args.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
