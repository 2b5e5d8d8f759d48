"""torchrun check of the data-parallel single learner over CUDA IPC (one process per GPU, parallel.link_engine_distributed):
replicas stay bit-identical, the learner learns, and the per-update cost of the in-kernel exchange is measured.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_distributed_rl_b200 import parallel  # noqa: E402
from simple_distributed_rl_b200.engine import EngineConfig  # noqa: E402
from simple_distributed_rl_b200.runner import VecRunner  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    E = int(os.environ.get("DP_ENVS", "8192"))
    kw = dict(env="CartPole-v1", algo="rainbow", hidden=(512,), dueling="average", noisy=True, mem_kind=1, multisteps=3, n_envs=E,
              ring_rows=256, batch_size=32, warmup_size=1000, lr=1e-3, target_update_interval=1000, seed=1 + rank)
    runner = VecRunner(EngineConfig(**kw), device=dev)
    eng = runner.engine
    eng.run(256, 0)
    out = {"world": world, "envs_per_gpu": E}
    # single-GPU-style replica timing first (no exchange), then the linked single learner
    U = E // 10
    for name in ("replica", "single_learner"):
        if name == "single_learner":
            parallel.link_engine_distributed(eng, learner_seed=5)
        for _ in range(3):
            eng.vec_step(); eng.learn(U)
        torch.cuda.synchronize(dev); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st0 = eng.read_state()
        e0.record()
        for _ in range(40):
            eng.vec_step(); eng.learn(U)
        e1.record()
        torch.cuda.synchronize(dev); dist.barrier()
        eng.check_dp_alive()
        st1 = eng.read_state()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        n_upd = st1.train_count - st0.train_count
        out[name] = {"updates": int(n_upd), "ms": float(ms.item()), "us_per_update": 1e3 * float(ms.item()) / n_upd,
                     "env_steps_per_s_all_gpus": world * 40 * E / (float(ms.item()) * 1e-3),
                     "updates_per_s": n_upd / (float(ms.item()) * 1e-3) * (world if name == "replica" else 1),
                     "global_batch": 32 * (world if name == "single_learner" else 1)}
    # replicas identical?
    flat = torch.cat([eng.t[k].reshape(-1) for k in ("params", "params_sigma", "target", "adam_m", "adam_v")])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    out["replicas_bit_identical"] = bool(all(torch.equal(gathered[0], g) for g in gathered))
    # train on and evaluate
    t0 = time.time()
    for _ in range(300):
        eng.vec_step(); eng.learn(U)
    torch.cuda.synchronize(dev)
    eng.check_dp_alive()
    out["greedy_reward_after_training"] = float(np.mean(runner.evaluate(max_episodes=50, test_epsilon=0.0)))
    out["train_seconds"] = time.time() - t0
    dist.barrier()
    if rank == 0:
        print("DPCHECK " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
