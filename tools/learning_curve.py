"""Learning curve of the bench configuration on the device path: Rainbow (double + dueling(512,) + NoisyNet + 3-step Retrace +
PER) on vectorised CartPole-v1, `train_interval` 10, evaluated greedily (100 fresh episodes, noise on as in the reference's
evaluation of a NoisyNet agent) every few vector steps.  usage: python tools/learning_curve.py [--out gpurun_out/curve.json]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_distributed_rl_b200.engine import EngineConfig  # noqa: E402
from simple_distributed_rl_b200.runner import VecRunner  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--vec-steps", type=int, default=600)
    ap.add_argument("--eval-every", type=int, default=50)
    ap.add_argument("--train-interval", type=int, default=10)
    a = ap.parse_args()
    kw = dict(env="CartPole-v1", algo="rainbow", hidden=(512,), dueling="average", noisy=True, mem_kind=1, multisteps=3, n_envs=a.envs,
              ring_rows=256, batch_size=32, warmup_size=1000, seed=1, enable_double_dqn=True, target_update_interval=1000, lr=1e-3,
              discount=0.99)
    r = VecRunner(EngineConfig(**kw))
    pts = []
    t_train = 0.0
    done_steps = 0
    while done_steps < a.vec_steps:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = r.train(max_steps=a.envs * a.eval_every, train_interval=a.train_interval, steps_per_call=10)
        torch.cuda.synchronize()
        t_train += time.perf_counter() - t0
        done_steps += a.eval_every
        rew = r.evaluate(max_episodes=100)
        es = r.engine.read_state()
        pts.append({"vec_steps": done_steps, "env_steps": int(es.total_step), "updates": int(es.train_count), "train_seconds": t_train,
                    "eval_mean": float(np.mean(rew)), "eval_min": float(np.min(rew)), "eval_max": float(np.max(rew))})
        print(json.dumps(pts[-1]), flush=True)
    out = {"config": {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()}, "train_interval": a.train_interval, "points": pts}
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
