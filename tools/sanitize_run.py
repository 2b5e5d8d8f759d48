#!/usr/bin/env python
"""Small learner workload for compute-sanitizer (memcheck / racecheck / synccheck): a few vector steps and updates of the
Rainbow default shape (learner_fast_kernel<16>) and of a generic-kernel shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_distributed_rl_b200.engine import DeviceEngine, EngineConfig  # noqa: E402

kw = dict(env="CartPole-v1", algo="rainbow", hidden=(512,), dueling="average", noisy=True, mem_kind=1, multisteps=3,
          n_envs=64, ring_rows=16, batch_size=32, warmup_size=64, seed=1)
d = DeviceEngine(EngineConfig(**kw))
d.run(20, 0)
d.learn(int(os.environ.get("SAN_UPDATES", "6")))
torch.cuda.synchronize()
print("fast:", d.learner_info(), d.read_state().train_count)
kw2 = dict(env="CartPole-v1", algo="dqn", hidden=(64, 64), mem_kind=1, multisteps=1, n_envs=64, ring_rows=8, batch_size=32,
           warmup_size=64, epsilon=0.2)
g = DeviceEngine(EngineConfig(**kw2))
g.run(10, 0)
g.learn(3)
torch.cuda.synchronize()
print("dqn 64x64 PER (row-split + replay CTA):", g.learner_info(), g.read_state().train_count)
kw2b = dict(env="Grid", algo="rainbow", hidden=(32, 16), dueling=None, noisy=True, mem_kind=1, multisteps=1, n_envs=24, ring_rows=4,
            batch_size=8, warmup_size=24)
gg = DeviceEngine(EngineConfig(**kw2b))
gg.run(6, 0)
gg.learn(3)
torch.cuda.synchronize()
print("generic:", gg.learner_info(), gg.read_state().train_count)
if os.environ.get("SAN_SMALL", "1") == "1":  # learner_small_kernel: uniform replay, plain weights, rows over an 8-CTA cluster
    for kw3 in (dict(env="CartPole-v1", algo="dqn", hidden=(64, 64), mem_kind=0, multisteps=1, n_envs=64, ring_rows=8, batch_size=32,
                     warmup_size=64, epsilon=0.2),
                dict(env="Grid", algo="rainbow", hidden=(32, 16), dueling="average", noisy=False, mem_kind=0, multisteps=3, n_envs=24,
                     ring_rows=9, batch_size=16, warmup_size=48, epsilon=0.3, enable_double_dqn=False)):
        s = DeviceEngine(EngineConfig(**kw3))
        s.run(12, 0)
        s.learn(4)
        s.vec_step()
        s.learn(2)
        torch.cuda.synchronize()
        print("small:", s.learner_info(), s.read_state().train_count)
