"""R2D2 on the device learns: trains on CartPole-v1 (E env copies, LSTM 64, dueling 64, burn-in 5, sequence 10, proportional sequence
replay) and evaluates 20 greedy episodes every few hundred updates; wall clock of the training part beside it."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from simple_distributed_rl_b200.r2d2 import R2D2Config, R2D2Runner


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="CartPole-v1")
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--rounds", type=int, default=20)
    ap.add_argument("--updates-per-round", type=int, default=500)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    cfg = R2D2Config(env=a.env, n_envs=a.envs, lstm_units=64, hidden_layers=(64,), dueling_type="average", burnin=5, sequence_length=10,
                     batch_size=64, capacity=a.envs * 400, warmup_size=a.envs * 20, memory="Proportional", enable_rescale=False,
                     enable_retrace=True, lr=1e-3, target_model_update_interval=200, epsilon=0.1, discount=0.99, seed=1)
    r = R2D2Runner(cfg)
    curve, t_train = [], 0.0
    for k in range(a.rounds):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = r.train(max_train_count=a.updates_per_round, updates_per_vec_step=4)
        torch.cuda.synchronize()
        t_train += time.perf_counter() - t0
        s = r.engine.read_state()
        rew = r.evaluate(max_episodes=20)
        curve.append(dict(train_count=int(s.train_count), env_steps=int(s.total_step), train_seconds=round(t_train, 3),
                          eval_mean=float(np.mean(rew)), eval_min=float(np.min(rew)), loss=float(s.last_loss)))
        print(json.dumps(curve[-1]), flush=True)
    if a.out:
        json.dump(dict(config=cfg.__dict__, curve=curve), open(a.out, "w"), indent=1, default=str)


if __name__ == "__main__":
    main()
