"""torchrun check of R2D2 with actor shards + ONE learner (BASELINE configs[3]): every rank rolls out its own env copies into its own
replay shard; every update is one trainer step on the global batch (NCCL all-reduce of the flat gradient between backward and Adam).
Checks that parameters / Adam state / target stay bit-identical across ranks, reports the time per update next to the single-GPU time.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/r2d2_dp_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_distributed_rl_b200.r2d2 import R2D2Config, R2D2Engine  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    small = bool(os.environ.get("R2D2_SMALL"))
    E = 64 if small else 2048
    cfg = R2D2Config(env="CartPole-v1", n_envs=E, lstm_units=32 if small else 512, hidden_layers=(32,) if small else (512,), dueling_type="average",
                     burnin=4 if small else 40, sequence_length=8 if small else 80, batch_size=32 if small else 64, capacity=E * 128,
                     warmup_size=E * 4, memory="Proportional", enable_rescale=True, enable_retrace=False, lr=1e-4, seed=1 + rank)
    eng = R2D2Engine(cfg, device=dev)
    W = cfg.burnin + cfg.sequence_length
    for _ in range(W + 8):
        eng.vec_step(True)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def timed(n):
        torch.cuda.synchronize(dev)
        dist.barrier()
        a, b = ev(), ev()
        a.record()
        eng.learn(n)
        b.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([a.elapsed_time(b) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng.learn(2)
    single_ms = timed(10)          # every rank its own trainer (no exchange)
    eng.link_data_parallel()
    eng.learn(2)
    dp_ms = timed(10)
    for _ in range(5):             # keep acting and learning; the replicas must stay identical
        eng.vec_step(True)
        eng.learn(1)
    torch.cuda.synchronize(dev)
    same = True
    for k in ("params", "target", "adam_m", "adam_v"):
        lo, hi = eng.t[k].clone(), eng.t[k].clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same &= bool(torch.equal(lo, hi))
    st = eng.read_state()
    if rank == 0:
        out = dict(world=world, envs_per_gpu=E, batch_per_rank=cfg.batch_size, global_batch=world * cfg.batch_size, n_params=eng.spec.n_params,
                   update_ms_independent=single_ms, update_ms_single_learner=dp_ms, allreduce_overhead_ms=dp_ms - single_ms,
                   sequences_per_s=world * cfg.batch_size / dp_ms * 1e3, replicas_bit_identical=same, train_count=int(st.train_count),
                   loss=st.last_loss)
        print("R2D2DP " + json.dumps(out))
        assert same, "replicas diverged"
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
