"""The image path learns: the reference's Runner (its loop, Worker and processors' space logic) over the device processor, conv Q-network,
trainer and uint8 replay on tests/image_env.py::PixelGrid (30 x 24 RGB frames -> 28 x 36 gray, window 2).  Optimal return 0.94 (7 steps to
the goal, -0.01 per step).  Usage: python tools/image_learning_check.py [--seeds 3] [--train 3000]"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
for p in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
    if os.path.isfile(os.path.join(p, "srl", "__init__.py")):
        sys.path.insert(0, p)
        break

import srl  # noqa: E402
from srl.algorithms import dqn  # noqa: E402
from srl.base.define import SpaceTypes  # noqa: E402

import image_env  # noqa: E402
from simple_distributed_rl_b200 import srl_image  # noqa: E402


def run(seed, n_train, device_memory=True):
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    cfg = dqn.Config(batch_size=32, lr=5e-4, target_model_update_interval=200, discount=0.95)
    cfg.epsilon_scheduler.set_linear(1.0, 0.05, 2000)
    cfg.input_block.image.set_dqn_block(filters=16)
    cfg.input_block.image.processors = [srl_image.DeviceImageProcessor(SpaceTypes.GRAY_HW1, (36, 28), normalize_type="0to1")]
    cfg.hidden_block.set((128,))
    cfg.window_length = 2
    cfg.memory.capacity, cfg.memory.warmup_size, cfg.memory.compress = 10_000, 200, False
    runner = srl.Runner("PixelGrid-b200", cfg)
    runner.set_seed(seed)
    t0 = time.time()
    runner.train(max_train_count=n_train, enable_progress=False)
    t_train = time.time() - t0
    rewards = runner.evaluate(max_episodes=10, enable_progress=False)
    return float(np.mean(rewards)), t_train


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--train", type=int, default=3000)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    image_env.register()
    srl_image.register(device_memory=True)
    res = []
    for seed in range(1, args.seeds + 1):
        r, t = run(seed, args.train)
        res.append({"seed": seed, "mean_eval_reward": r, "train_s": t})
        print("IMAGELEARN", json.dumps(res[-1]), flush=True)
    srl_image.unregister()
    if args.out:
        json.dump({"train_count": args.train, "optimal": 0.94, "runs": res}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
