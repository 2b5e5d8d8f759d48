"""Exploration for the Grid / Rainbow+PER learning gate: mean greedy reward per seed and training length (GPU)."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from simple_distributed_rl_b200.engine import EngineConfig
from simple_distributed_rl_b200.runner import VecRunner

out = {}
for steps in (600, 1200, 2400):
    for seed in (1, 2, 3, 4, 5, 6):
        kw = dict(env="Grid", algo="rainbow", hidden=(64,), dueling="average", noisy=False, mem_kind=1, multisteps=3, n_envs=256,
                  ring_rows=64, batch_size=32, warmup_size=1000, epsilon=0.1, lr=1e-3, target_update_interval=1000, seed=seed)
        r = VecRunner(EngineConfig(**kw))
        r.train(max_steps=kw["n_envs"] * steps, train_interval=1, steps_per_call=16)
        m = float(np.mean(r.evaluate(max_episodes=100, test_epsilon=0.0)))
        out[f"{steps}_{seed}"] = m
        print(steps, seed, m, flush=True)
json.dump(out, open("gpurun_out/grid_gate.json", "w"))
