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

Single-learner mode (`link_engines` / `link_engine_distributed`): the reference has ONE trainer
(srl/base/run/play_mp.py:352-462) fed by all actors.  Here every rank keeps its replay shard and samples `batch_size` items from
it; the gradients of the global batch are summed over NVLink INSIDE the learner kernel on every update (peer stores into the
ranks' exchange buffers + release/acquire flags, csrc/learner_fast.cu), the IS weights use the global N / total / max
(proportional_memory.py:138-167), and every rank applies the identical Adam step, so parameters, moments and the target network
stay bit-identical on all ranks with no separate broadcast.  Replica mode (`average_parameters`) stays available.
"""
import ctypes as C
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


# ---- single-learner mode: exchange buffers of the in-kernel gradient all-reduce ------------------------------------------------
_PARAM_TENSORS = ("params", "params_sigma", "target", "target_sigma", "adam_m", "adam_v")


def link_engines(engines, learner_seed: int = 1) -> None:
    """ONE process driving several engines, one per device: wire them into a data-parallel learner.  Rank r = engines[r].
    Parameters, moments, target network and trainer counters are copied from engines[0]; the exchange buffers are plain zeroed
    device tensors, addressable across devices after cudaDeviceEnablePeerAccess."""
    from . import _lib

    world = len(engines)
    if world < 2:
        return
    lib = engines[0].lib
    for a in engines:
        for b in engines:
            if a is not b and a.device != b.device:
                _lib.check(lib.srlx_dp_enable_peer(a.device.index or 0, b.device.index or 0))
    nbytes = engines[0].dp_bytes()
    if nbytes == 0:
        raise _lib.SrlxError("the data-parallel learner needs the single-hidden-layer cluster kernel (srlx_learner_info == 1)")
    bufs = []
    for e in engines:
        if e.dp_bytes() != nbytes:
            raise ValueError("link_engines: the engines' networks differ")
        e.t["dp_xchg"] = torch.zeros(nbytes, dtype=torch.uint8, device=e.device)
        bufs.append(e.t["dp_xchg"])
    src = engines[0]
    st0 = src.read_state()
    for r, e in enumerate(engines):
        if r > 0:
            for k in _PARAM_TENSORS:
                if k in e.t:
                    e.t[k].copy_(src.t[k])
            st = e.read_state()
            st.train_count, st.adam_step, st.sync_count = st0.train_count, st0.adam_step, st0.sync_count
            e.write_state(st)
        e.set_data_parallel(world, r, [b.data_ptr() for b in bufs], nbytes, learner_seed)
    for e in engines:
        torch.cuda.synchronize(e.device)


def link_engine_distributed(engine, learner_seed: int = 1, group=None) -> None:
    """One process per GPU (torch.distributed initialised, all ranks on one node): allocate this rank's exchange buffer, swap CUDA
    IPC handles with the peers, map theirs, broadcast rank 0's parameters / moments / target / trainer counters."""
    from . import _lib

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world < 2:
        return
    lib = engine.lib
    nbytes = engine.dp_bytes()
    if nbytes == 0:
        raise _lib.SrlxError("the data-parallel learner needs the single-hidden-layer cluster kernel (srlx_learner_info == 1)")
    own, handle = C.c_void_p(0), C.create_string_buffer(64)
    with torch.cuda.device(engine.device):
        _lib.check(lib.srlx_dp_alloc(nbytes, C.byref(own), handle))
    handles = [None] * world
    dist.all_gather_object(handles, bytes(handle.raw), group=group)
    peers = []
    for r in range(world):
        if r == rank:
            peers.append(own.value)
        else:
            p = C.c_void_p(0)
            with torch.cuda.device(engine.device):
                _lib.check(lib.srlx_dp_open(handles[r], C.byref(p)))
            peers.append(p.value)
    engine._dp_own_ptr, engine._dp_peer_ptrs = own.value, peers
    tensors = [engine.t[k] for k in _PARAM_TENSORS if k in engine.t]
    broadcast_parameters(tensors, src=0, group=group)
    st = engine.read_state()
    cnt = torch.tensor([st.train_count, st.adam_step, st.sync_count], dtype=torch.int64, device=engine.device)
    dist.broadcast(cnt, src=0, group=group)
    st.train_count, st.adam_step, st.sync_count = [int(x) for x in cnt.tolist()]
    engine.write_state(st)
    engine.set_data_parallel(world, rank, peers, nbytes, learner_seed)
    torch.cuda.synchronize(engine.device)
    dist.barrier(group=group)
