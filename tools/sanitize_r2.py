"""Small workloads of the round-2 kernels for compute-sanitizer (memcheck / racecheck / synccheck): R2D2 (persistent unroll kernels with
their grid barrier, per-step launches, replay add, PER sample / priority update, 3 x TF32 GEMM tiles, split-K), the one-launch
IPriorityMemory seam over mapped host memory, invalid-action masks in the generic learner, PPO, rank-based replay."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_distributed_rl_b200.engine import DeviceEngine, EngineConfig  # noqa: E402
from simple_distributed_rl_b200.memory import DeviceProportionalMemory, DeviceRankBasedMemory  # noqa: E402
from simple_distributed_rl_b200.r2d2 import R2D2Config, R2D2Engine  # noqa: E402

what = set(sys.argv[1:]) or {"r2d2", "seam", "masks", "ppo", "rank"}
rng = np.random.default_rng(0)
if "r2d2" in what:
    for kw in (dict(memory="Proportional", batch_size=32, warmup_size=32, lstm_units=24),
               dict(memory="ReplayBuffer", batch_size=8, warmup_size=8, lstm_units=16, dueling_type=None, hidden_layers=(12, 8)),
               dict(memory="Proportional", batch_size=40, warmup_size=40, lstm_units=40, _persistent=False)):
        persistent = kw.pop("_persistent", True)
        cfg = R2D2Config(env="CartPole-v1", n_envs=9, hidden_layers=kw.pop("hidden_layers", (12,)), burnin=2, sequence_length=3, capacity=9 * 12,
                         epsilon=0.5, **kw)
        eng = R2D2Engine(cfg, persistent=persistent)
        for g in range(30):
            eng.vec_step(True)
            eng.learn(1)
        torch.cuda.synchronize()
        print("r2d2 ok", eng.read_state().train_count, eng.read_state().last_loss)
if "seam" in what:
    m = DeviceProportionalMemory(1000, 0.8, 0.4, 1000, has_duplicate=True)
    for i in range(300):
        m.add(i, float(rng.random()) if i % 3 else None)
    for step in range(5):
        m.add(1000 + step, float(rng.random()))
        b, w, idx = m.sample(64, step)
        m.update(idx, rng.random(64))
    print("seam ok", m.length(), m.max_priority)
if "masks" in what:
    D, A, E, R = 3, 5, 4, 16
    for algo, M in (("dqn", 1), ("rainbow", 3)):
        dev = DeviceEngine(EngineConfig(env="external", env_kwargs=dict(obs_dim=D, n_actions=A), n_envs=E, ring_rows=R, batch_size=8, warmup_size=8,
                                        algo=algo, hidden=(32,), dueling="average" if algo == "rainbow" else None, multisteps=M, mem_kind=1,
                                        invalid_actions=True))
        for g in range(12):
            dev.ext_step(rng.normal(size=(E, D)), rng.normal(size=(E, D)), rng.integers(0, A, E), rng.normal(size=E),
                         (rng.random(E) < 0.1), (rng.random(E) < 0.15), next_invalid=rng.integers(0, 1 << (A - 1), E))
            if g >= 5:
                dev.learn(1)
        torch.cuda.synchronize()
        print("masks ok", algo, dev.read_state().train_count)
if "ppo" in what:
    from simple_distributed_rl_b200.ppo import PPOConfig, PPOEngine

    eng = PPOEngine(PPOConfig(env="Pendulum-v1", n_envs=16, horizon=8, batch_size=8, warmup_size=32))
    eng.rollout()
    eng.finish_rollout()
    eng.learn(3)
    torch.cuda.synchronize()
    print("ppo ok", eng.read_pstate().train_count)
if "rank" in what:
    m = DeviceRankBasedMemory(200, 0.6, 0.4, 1000)
    for i in range(150):
        m.add(i, float(rng.random()))
    b, w, idx = m.sample(16, 3)
    m.update(idx, rng.random(16))
    print("rank ok", m.length())
