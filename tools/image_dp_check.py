"""torchrun check of the conv Q-network's data-parallel trainer (image.ImageQNet.train_data_parallel): every rank trains on its own shard
of the batch (Atari setting, 32 uint8 states per rank), ONE NCCL all-reduce of the 16 MB flat gradient per update, the same Adam step on
every rank.  Prints ms per update (device-timed, max over ranks) and whether the replicas stayed bit-identical.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/image_dp_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_distributed_rl_b200 import image  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = 32
    spec = image.ImageNetSpec((84, 84, 4), "IMAGE_MAP", 6)
    net = image.ImageQNet(spec, batch_size=B, uint8_states=True, seed=5, device=str(dev))
    gen = torch.Generator(device=dev).manual_seed(100 + rank)  # every rank its own shard
    fr = torch.randint(0, 256, (2, B, 84, 84, 4), dtype=torch.uint8, device=dev, generator=gen)
    a = torch.randint(0, 6, (B,), dtype=torch.int32, device=dev, generator=gen)
    r, ud, w = torch.randn(B, device=dev, generator=gen), torch.ones(B, device=dev), torch.rand(B, device=dev, generator=gen) * 0.7 + 0.3
    for _ in range(5):
        net.train_data_parallel(fr[0], fr[1], a, r, ud, w)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        net.train_data_parallel(fr[0], fr[1], a, r, ud, w)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
    chk = torch.stack([net.params.double().sum(), net.params.double().abs().sum(), net.adam_v.double().sum(), net.target.double().sum()])
    same = True
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        all_chk = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(all_chk, chk)
        same = all(torch.equal(c, all_chk[0]) for c in all_chk)
    if rank == 0:
        print("IMAGEDP " + json.dumps({"n_gpus": world, "batch_per_gpu": B, "global_batch": B * world, "ms_per_update": float(ms[0]),
                                       "samples_per_s": B * world / (float(ms[0]) * 1e-3), "replicas_bit_identical": bool(same),
                                       "gradient_bytes": spec.n_params * 4, "train_count": net.train_count}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
