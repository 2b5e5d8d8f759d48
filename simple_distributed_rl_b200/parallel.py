"""Multi-GPU plumbing: one process per GPU (torch.distributed: NCCL on the GPU box, gloo in the CPU tests).

The path shards by env copies: rank g owns its E env copies, its ring-replay shard, its SumTree shard and a
learner with a full weight replica -- experience never leaves the GPU that generated it.  What crosses NVLink:
  * `average_parameters`: every `sync_interval` steps the replicas' online parameters (mu, sigma; 27-53 KB) are averaged
    with ONE all-reduce on a flat buffer (the device analogue of the reference's parameter "board", a pickled CPU
    state_dict polled once a second: srl/base/run/play_mp.py:140-165,289-318).
  * `reduce_counters`: max-over-ranks time and summed step counts for the whole-job rates
    (srl/runner/callbacks/print_progress.py:224-237 defines the rates per process).
Per-actor exploration follows the reference's Ape-X ladder (srl/rl/functions.py:145-154, hooked by
rainbow.Config.setup_from_actor, srl/algorithms/rainbow/rainbow.py:109-114).
"""
from dataclasses import replace
from typing import List, Sequence

import torch
import torch.distributed as dist


def create_epsilon_list(policy_num: int, epsilon: float = 0.4, alpha: float = 8.0) -> List[float]:
    """srl/rl/functions.py:145-154."""
    assert policy_num > 0
    if policy_num == 1:
        return [epsilon / 4]
    return [epsilon ** (1 + (i / (policy_num - 1)) * alpha) for i in range(policy_num)]


def shard_config(cfg, rank: int, world: int, actor_epsilon: float = None, actor_alpha: float = 7.0):
    """Per-rank EngineConfig: independent Philox streams (seed), optional Ape-X epsilon ladder over the ranks."""
    out = replace(cfg, seed=int(cfg.seed) * 1_000_003 + rank)
    if actor_epsilon is not None:
        out = replace(out, epsilon=create_epsilon_list(world, actor_epsilon, actor_alpha)[rank])
    return out


def average_parameters(tensors: Sequence[torch.Tensor], group=None) -> None:
    """In-place mean over ranks of a list of same-dtype tensors through ONE flat all-reduce."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


def broadcast_parameters(tensors: Sequence[torch.Tensor], src: int = 0, group=None) -> None:
    """Learner -> actors parameter push (play_mp.py:289-318) as one flat broadcast."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


def reduce_counters(times_ms: Sequence[float], counts: Sequence[float], device="cpu", group=None):
    """(max over ranks of each time, sum over ranks of each count)."""
    t = torch.tensor(list(times_ms), dtype=torch.float64, device=device)
    c = torch.tensor(list(counts), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    return [float(x) for x in t.tolist()], [float(x) for x in c.tolist()]
