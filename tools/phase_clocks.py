#!/usr/bin/env python
"""Learner phase clocks: launches the bench workload's learner with the debug clock tap and prints, for the second-to-last
update of a launch, the clock64() deltas between the SRLX_STAMP points of CTA 0 (csrc/learner.cu).  Diagnostic only."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_distributed_rl_b200.engine import DeviceEngine, EngineConfig  # noqa: E402


def main():
    E = int(os.environ.get("PC_ENVS", "8192"))
    R = int(os.environ.get("PC_ROWS", "256"))
    kw = dict(env="CartPole-v1", algo="rainbow", hidden=(512,), dueling="average", noisy=True, mem_kind=1, multisteps=3,
              n_envs=E, ring_rows=R, batch_size=32, warmup_size=1000, seed=1)
    if os.environ.get("PC_WORKLOAD") == "dqn":  # BASELINE configs[1]
        kw = dict(env="CartPole-v1", algo="dqn", hidden=(64, 64), mem_kind=0, multisteps=1, n_envs=E, ring_rows=R, batch_size=32,
                  warmup_size=1000, seed=1, epsilon=0.1)
    if os.environ.get("PC_WORKLOAD") == "dqn512":  # the reference's default DQN: one hidden layer of 512, uniform replay
        kw = dict(env="CartPole-v1", algo="dqn", hidden=(512,), mem_kind=0, multisteps=1, n_envs=E, ring_rows=R, batch_size=32,
                  warmup_size=1000, seed=1, epsilon=0.1)
    if os.environ.get("PC_WORKLOAD") == "dqn_per":  # DQN MLP[64,64] on proportional replay: the generic learner_kernel
        kw = dict(env="CartPole-v1", algo="dqn", hidden=(64, 64), mem_kind=1, multisteps=1, n_envs=E, ring_rows=R, batch_size=32,
                  warmup_size=1000, seed=1, epsilon=0.1)
    if os.environ.get("PC_PRESAMPLE"):
        kw["presample"] = True
    d = DeviceEngine(EngineConfig(**kw), debug=True)
    for f in ("dbg_q", "dbg_action", "dbg_sample_idx", "dbg_weights", "dbg_target_q", "dbg_q_sa", "dbg_grads", "dbg_windows"):
        setattr(d.c, f, None)  # only the clock tap stays on: the other taps cost time inside the learner
    d.run(R, 0)
    d.learn(64)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d.learn(1024)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c = d.t["dbg_clock"].cpu().numpy().astype("int64")
    nz = [int(x) for x in c if x > 0]
    base = min(nz) if nz else 0  # no stamps: not the -DSRLX_STAMPS build (SRLX_LIB=.../libsrlx_stamps.so)
    rel = {i: int(c[i] - base) for i in range(64) if c[i] > 0}
    out = {"us_per_update": 1e3 * ms / 1024, "learner": d.learner_info(), "stamps_cycles_rel": rel}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
