"""world_size-2 gloo tests (CPU) of the N>1 host logic: shard configs, counter reduction, parameter averaging/broadcast."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simple_distributed_rl_b200 import parallel
    from simple_distributed_rl_b200.engine import EngineConfig

    cfg = parallel.shard_config(EngineConfig(seed=7, epsilon=0.1), rank, world, actor_epsilon=0.4, actor_alpha=7.0)
    mu = torch.full((5,), float(rank + 1))
    sg = torch.arange(3, dtype=torch.float32) * (rank + 1)
    parallel.average_parameters([mu, sg])
    b = torch.full((4,), float(rank))
    parallel.broadcast_parameters([b], src=1)
    t, c = parallel.reduce_counters([10.0 + rank, 3.0 - rank], [100.0 * (rank + 1), 5.0])
    q.put((rank, cfg.seed, cfg.epsilon, mu.tolist(), sg.tolist(), b.tolist(), t, c))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    from simple_distributed_rl_b200 import parallel

    eps = parallel.create_epsilon_list(2, 0.4, 7.0)
    assert eps == [0.4, 0.4 ** 8.0]
    for rank, seed, e, mu, sg, b, t, c in out:
        assert seed == 7 * 1_000_003 + rank and e == pytest.approx(eps[rank])
        assert mu == [1.5] * 5 and sg == [0.0, 1.5, 3.0]      # mean of the two replicas
        assert b == [1.0] * 4                                 # broadcast from rank 1
        assert t == [11.0, 3.0] and c == [300.0, 10.0]        # max of times, sum of counts
    assert out[0][1] != out[1][1]


def test_epsilon_ladder_matches_reference_formula():
    from simple_distributed_rl_b200 import parallel

    assert parallel.create_epsilon_list(1, 0.4, 8.0) == [0.1]
    l = parallel.create_epsilon_list(5, 0.4, 8.0)
    assert l[0] == 0.4 and l[-1] == pytest.approx(0.4 ** 9.0) and all(a > b for a, b in zip(l, l[1:]))
