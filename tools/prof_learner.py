#!/usr/bin/env python
"""Workload for `ncu -k regex:learner_fast_kernel`: the bench configuration, ring filled, a few launches of 128 updates."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_distributed_rl_b200.engine import DeviceEngine, EngineConfig  # noqa: E402

kw = dict(env="CartPole-v1", algo="rainbow", hidden=(512,), dueling="average", noisy=True, mem_kind=1, multisteps=3,
          n_envs=8192, ring_rows=256, batch_size=32, warmup_size=1000, seed=1)
d = DeviceEngine(EngineConfig(**kw))
d.run(256, 0)
for rep in range(3):
    d.learn(int(os.environ.get("PROF_UPDATES", "128")))
torch.cuda.synchronize()
