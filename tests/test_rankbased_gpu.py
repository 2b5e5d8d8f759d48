"""GPU tests of the rank-based replay memory on device (csrc/rankbased.cu, memory.DeviceRankBasedMemory) against the reference's
RankBasedMemory: golden sequences produced by the reference itself (tests/golden/rankbased.npz), the oracle restatement at large
sizes, and the reference's own protocol test restated (tests/quick/rl/memories/test_priority_memories.py:17-91)."""
import collections
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import rankbased as orb  # noqa: E402


def test_device_rankbased_equals_reference_golden(golden_dir):
    """add / sample / update through the IPriorityMemory methods, numpy's uniform stream injected: the sampled items are exactly the
    reference's, IS weights (float64) to 1e-10, priorities after the updates bit-equal."""
    from simple_distributed_rl_b200.memory import DeviceRankBasedMemory

    g = np.load(os.path.join(golden_dir, "rankbased.npz"))
    for c in range(int(g["n_cases"])):
        cap, alpha, beta0, bsteps, B, n_add = g[f"c{c}_cfg"]
        cap, B, n_add = int(cap), int(B), int(n_add)
        m = DeviceRankBasedMemory(cap, alpha, beta0, bsteps)
        for i, p in enumerate(g[f"c{c}_pri0"]):
            m.add(("item", i), float(p))
        assert m.length() == n_add
        for step in range(len(g[f"c{c}_idx"])):
            batches, w, idx = m.sample(B, step * 7, uniforms=g[f"c{c}_u"][step])
            np.testing.assert_array_equal(idx, g[f"c{c}_idx"][step])
            assert [b[1] for b in batches] == list(idx)
            np.testing.assert_allclose(w, g[f"c{c}_w"][step], rtol=1e-10)
            m.update(idx, g[f"c{c}_upd"][step])
        b = m.backup()
        assert b[0] == cap and len(b[1]) == n_add and b[3] == n_add % cap
        np.testing.assert_array_equal(b[2], g[f"c{c}_pri_final"])
        m2 = DeviceRankBasedMemory(cap, alpha, beta0, bsteps)
        m2.restore(b)
        assert m2.length() == n_add and np.array_equal(m2.backup()[2], b[2])


@pytest.mark.parametrize("n", [1, 2, 33, 4096, 4097, 100_003, 1 << 20, (1 << 21) + 5])
def test_device_argsort_equals_numpy(n):
    """The radix sort: descending priority, NaN last, -0.0 == +0.0, equal priorities in ascending item order -- numpy's stable
    argsort of -priorities.  Sizes around the 4096-key tile, ragged, and the 2M items of BASELINE configs[2]."""
    from simple_distributed_rl_b200.memory import DeviceRankBasedMemory

    rng = np.random.default_rng(n)
    p = (rng.standard_normal(n) ** 2).astype(np.float32)
    if n > 8:
        p[rng.integers(0, n, size=n // 8)] = np.float32(0.25)   # ties
        p[rng.integers(0, n, size=max(1, n // 50))] = np.nan     # items added without a priority
        p[rng.integers(0, n, size=3)] = np.float32(-0.0)
        p[rng.integers(0, n, size=3)] = np.float32(0.0)
        p[rng.integers(0, n, size=3)] = np.float32(-1.5)         # negative priorities sort below zero
        p[rng.integers(0, n, size=2)] = np.float32(np.inf)
    m = DeviceRankBasedMemory(n)
    m.buffer = [None] * n
    m._pri.copy_(torch.as_tensor(p))
    got = m.argsort()
    want = np.argsort(-p, kind="stable")
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("n,alpha,B", [(1 << 20, 0.6, 32), (300_000, 1.0, 64), (2_000_000, 0.8, 256)])
def test_device_rankbased_equals_oracle_at_replay_scale(n, alpha, B):
    """A million-item memory (distinct priorities): device sample == oracle sample on the same uniform stream, several
    sample / update rounds, and the cached rank cdf survives the updates."""
    from simple_distributed_rl_b200.memory import DeviceRankBasedMemory

    rng = np.random.default_rng(7)
    p = rng.permutation(n).astype(np.float32) / np.float32(n) * 4 + np.float32(1e-3)
    o = orb.RankBasedMemory(n, alpha, 0.4, 100)
    o.priorities[:] = p
    o.size = n
    m = DeviceRankBasedMemory(n, alpha, 0.4, 100)
    m.buffer = list(range(n))
    m._pri.copy_(torch.as_tensor(p))
    for step in range(3):
        u = rng.random(4 * B)
        want_idx, want_w, ranks, used = o.sample(B, step * 10, u)
        batches, w, idx = m.sample(B, step * 10, uniforms=u)
        np.testing.assert_array_equal(idx, want_idx)
        np.testing.assert_allclose(w, want_w, rtol=1e-10)
        assert batches == list(want_idx) and w.max() == 1.0
        td = (rng.random(B) * 5 + 4.0).astype(np.float32) + np.arange(B, dtype=np.float32) * np.float32(1e-3)  # distinct, above all others
        o.update(want_idx, td)
        m.update(idx, td)
    np.testing.assert_array_equal(m.backup()[2], o.priorities)


def test_reference_protocol_test_restated():
    """tests/quick/rl/memories/test_priority_memories.py:17-91 for the "RankBasedMemory" parametrisation: fill, overwrite with
    priorities 1..10, 20000 x (sample 5 without duplicates -> update with the items' own priority -> backup/restore); every item is
    sampled and higher priority means more samples (monotone counts)."""
    from simple_distributed_rl_b200.memory import DeviceRankBasedMemory

    capacity = 10
    memory = DeviceRankBasedMemory(capacity, 0.8, 1, 10, seed=3)
    for i in range(100):
        memory.add((i, i, i, i), 0)
    assert memory.length() == capacity
    for i in range(10):
        i += 1
        memory.add((i, i, i, i), i)
        assert memory.length() == capacity
    counter = []
    for i in range(20000):
        batches, weights, update_args = memory.sample(5, step=1)
        assert len(batches) == 5 and len(weights) == 5
        assert len(list(set(batches))) == 5, list(set(batches))
        for batch in batches:
            counter.append(batch[0])
        memory.update(update_args, np.array([b[3] for b in batches]))
        assert memory.length() == capacity
        if i % 1000 == 0:
            l1 = memory.length()
            memory.restore(memory.backup())
            assert l1 == memory.length()
    counter = collections.Counter(counter)
    keys = sorted(counter.keys())
    assert keys == [i + 1 for i in range(capacity)]
    vals = [counter[key] for key in keys]
    for i in range(capacity - 1):
        assert vals[i] < vals[i + 1]


def test_reference_runner_trains_on_the_device_rankbased_memory(srl_mod):
    """The seam in place: the reference's own Runner / Trainer / Worker with memory.set_custom(DeviceRankBasedMemory)
    (priority_replay_buffer.py:111-117,149-152)."""
    import srl

    dqn, rainbow = srl_mod
    cfg = dqn.Config(batch_size=8)
    cfg.hidden_block.set((16,))
    cfg.memory.set_custom("simple_distributed_rl_b200.memory:DeviceRankBasedMemory", dict(alpha=0.7, beta_initial=0.5, beta_steps=100))
    cfg.memory.capacity, cfg.memory.warmup_size, cfg.memory.compress = 300, 16, False
    runner = srl.Runner("Grid", cfg)  # (device left on AUTO: the reference fixes the device process-wide at the first run)
    state = runner.train(max_train_count=40)
    assert state.trainer.get_train_count() == 40
    mem = state.memory.memory
    assert type(mem).__name__ == "DeviceRankBasedMemory" and mem.length() >= 40
    pri = mem.backup()[2][: mem.length()]
    assert np.isfinite(pri).sum() >= 8  # updated items carry |td|; never-sampled ones keep NaN (priority None), as in the reference
