"""Checkpoint / wire compatibility: the device state written in the reference's file formats and read
back from reference-written artefacts (simple_distributed_rl_b200/checkpoint.py).  Host-side only -- no GPU needed.

Two kinds of checks:
  * against the committed golden trajectories of the reference Runner (tests/golden/worker_records.npz): the items exported
    from a ring holding that trajectory are the items the reference worker handed to memory.add();
  * against the imported reference (present in the build container only; skipped elsewhere): files interchange both ways,
    the reference's own memory / parameter / trainer classes consume what we write, and we consume what they write.
"""
import os
import pickle
import sys
import zlib

import numpy as np
import pytest

from simple_distributed_rl_b200 import checkpoint as ck

REF = "/root/reference"


@pytest.fixture(scope="module")
def golden_dir():
    return os.path.join(os.path.dirname(__file__), "golden")


def _ring_from_golden(d, name, M, E=1, clip=False, extra_rows=0):
    """One env column holding the golden trajectory of the reference Runner on Grid."""
    T = len(d[f"{name}_a"])
    ring = ck.RingView(E, T + extra_rows, M, 4, 2, vec_steps=T)
    r = d[f"{name}_r"]
    ring.obs[:T], ring.next_obs[:T] = d[f"{name}_s"], d[f"{name}_ns"]
    ring.action[:T], ring.reward[:T] = d[f"{name}_a"], (np.sign(r) if clip else r).astype(np.float32)
    ring.term[:T], ring.done[:T] = d[f"{name}_term"], d[f"{name}_done"]
    return ring


# ---- golden: exported items == the reference worker's records ------------------------------------------------------------
def test_exported_dqn_items_are_the_reference_worker_records(golden_dir):
    d = np.load(os.path.join(golden_dir, "worker_records.npz"))
    ring = _ring_from_golden(d, "dqn", 1, extra_rows=3)
    inner, demo = ck.memory_backup(ring, proportional=False)
    items, idx = inner
    assert demo is None and idx == len(items) == 400  # ReplayBuffer.backup() = [memory, idx]; not full -> idx = len
    n = len(d["dqn_b_a"])
    for j in range(n):
        s, ns, onehot, r, undone, inv = items[j]
        np.testing.assert_array_equal(s, d["dqn_b_s"][j])
        np.testing.assert_array_equal(ns, d["dqn_b_ns"][j])
        assert s.dtype == np.float32 and isinstance(onehot, list) and onehot == [int(k == d["dqn_b_a"][j]) for k in range(4)]
        assert r == float(np.float32(d["dqn_b_r"][j])) and undone == d["dqn_b_undone"][j] and inv == []


@pytest.mark.parametrize("name,M,clip", [("rainbow_m3", 3, False), ("rainbow_m2_clip", 2, True)])
def test_exported_rainbow_items_are_the_reference_worker_windows(name, M, clip, golden_dir):
    d = np.load(os.path.join(golden_dir, "worker_records.npz"))
    ring = _ring_from_golden(d, name, M, clip=clip)
    items, _ = ck.export_items(ring, pad_action=lambda e, g: ck.philox_pad_action(7, e, g, 4))
    b_states, b_a, b_r, b_term = d[f"{name}_b_states"], d[f"{name}_b_a"], d[f"{name}_b_r"], d[f"{name}_b_term"]
    done = d[f"{name}_done"]
    assert len(items) == 400 - M + 1  # windows whose M steps are all in the ring
    for j in range(min(len(items), len(b_a))):
        it = items[j]
        assert len(it) == M + 1 and it[0][1:] == [None, None, None, None]
        np.testing.assert_array_equal(np.stack([e[0] for e in it]), b_states[j])
        ended = False
        for k in range(M):
            _, onehot, r, term, inv = it[k + 1]
            assert term == b_term[j][k] and float(np.float32(b_r[j][k])) == r and inv == []
            if not ended:
                assert int(np.argmax(onehot)) == b_a[j][k]
                ended = bool(done[j + k])
            else:
                assert sum(onehot) == 1 and r == 0 and term == 1  # padded tail (rainbow.py:358-371)


def test_pad_action_is_the_device_philox_stream():
    """Known answers of Philox4x32-10 through oracle/philox.py (pinned against the Random123 KAT in test_oracle_golden.py)."""
    from oracle import philox

    for seed, e, g, A in [(0, 0, 0, 2), (3, 17, 123456, 4), (2**40 + 5, 8191, 2**33 + 1, 10)]:
        w = philox.words(seed, philox.STREAM_PAD_ACTION, e, g & 0xFFFFFFFF, g >> 32)
        assert ck.philox_pad_action(seed, e, g, A) == (int(w[0]) * A) >> 32


# ---- synthetic ring: export -> import is the identity on everything the sampler can reach ----------------------------------
def _synthetic_ring(E, R, M, steps, seed, proportional):
    rng = np.random.default_rng(seed)
    ring = ck.RingView(E, R, M, 3, 2, vec_steps=steps)
    state = rng.normal(size=(E, 2)).astype(np.float32)
    for g in range(steps):
        for e in range(E):
            slot = (g % R) * E + e
            ns = (state[e] + rng.normal(size=2)).astype(np.float32)
            term = rng.random() < 0.15
            trunc = (not term) and rng.random() < 0.1
            ring.obs[slot], ring.next_obs[slot] = state[e], ns
            ring.action[slot], ring.reward[slot] = rng.integers(3), np.float32(rng.normal())
            ring.term[slot], ring.done[slot] = term, term or trunc
            state[e] = rng.normal(size=2).astype(np.float32) if (term or trunc) else ns
    if proportional:
        ring.leaf_priority = np.zeros(ring.capacity)
        g_lo, n_g = ring.valid_rows()
        for g in range(g_lo, g_lo + n_g):
            ring.leaf_priority[(g % R) * E: (g % R) * E + E] = rng.uniform(0.1, 2.0, size=E)
        ring.max_priority = 2.5
    return ring


@pytest.mark.parametrize("E,R,M,steps,prop,zip_items", [(3, 8, 3, 21, True, False), (4, 6, 1, 4, False, True), (2, 9, 2, 9, True, True),
                                                        (5, 7, 4, 30, False, False)])
def test_export_import_round_trip(E, R, M, steps, prop, zip_items):
    ring = _synthetic_ring(E, R, M, steps, seed=E * 100 + R, proportional=prop)
    pad = lambda e, g: ck.philox_pad_action(11, e, g, 3)  # noqa: E731
    backup = ck.memory_backup(ring, prop, pad_action=pad, compress=zip_items)
    backup = pickle.loads(pickle.dumps(backup))  # what a file round trip does
    back = ck.memory_restore(backup, E, R, M, 3, 2, prop)
    g_lo, n_g = ring.valid_rows()
    assert back.valid_rows() == (0, n_g) and back.vec_steps == n_g + M - 1
    a, _ = ck.export_items(ring, pad_action=None)
    b, pri_b = ck.export_items(back, pad_action=None)
    assert len(a) == len(b) == n_g * E
    assert pickle.dumps(a) == pickle.dumps(b)  # every window the sampler can rebuild is the same window
    if prop:
        _, pri_a = ck.export_items(ring)
        np.testing.assert_array_equal(pri_a, pri_b)
        inner = backup[0]
        cap = E * R
        assert inner[0] == cap and inner[1] == 2.5 and inner[2] == n_g * E and inner[3] == (n_g * E) % cap
        tree = np.asarray(inner[4])
        np.testing.assert_array_equal(tree[cap - 1: cap - 1 + n_g * E], pri_a)
        np.testing.assert_array_equal(tree[: cap - 1], tree[1: 2 * cap - 2: 2] + tree[2: 2 * cap - 1: 2])
        assert inner[5][n_g * E:] == [None] * (cap - n_g * E)


@pytest.mark.parametrize("cap", [1, 2, 5, 8, 37, 1000])
def test_build_sum_tree_every_node_is_left_plus_right(cap):
    rng = np.random.default_rng(cap)
    leaves = rng.uniform(0, 3, size=cap)
    tree = ck.build_sum_tree(leaves, cap)
    np.testing.assert_array_equal(tree[cap - 1:], leaves)
    for i in range(cap - 1):
        assert tree[i] == tree[2 * i + 1] + tree[2 * i + 2]


# ---- against the imported reference ---------------------------------------------------------------------------------------
def test_files_interchange_with_the_reference(srl_mod, tmp_path):
    sys.path.insert(0, REF)
    try:
        from srl.utils import common
    finally:
        sys.path.remove(REF)
    obj = {"a": np.arange(5), "b": [1, 2.5, "x"]}
    for compress in (True, False):
        p1, p2 = str(tmp_path / f"ours_{compress}.dat"), str(tmp_path / f"theirs_{compress}.dat")
        ck.save_file(p1, obj, compress)
        common.save_file(p2, obj, compress)
        assert open(p1, "rb").read(6) == open(p2, "rb").read(6)
        for path in (p1, p2):
            for load in (ck.load_file, common.load_file):
                got = load(path)
                np.testing.assert_array_equal(got["a"], obj["a"])
                assert got["b"] == obj["b"]


@pytest.mark.parametrize("algo", ["dqn", "rainbow"])
def test_parameter_files_interchange_with_the_reference(srl_mod, tmp_path, algo):
    """our save -> the reference's Parameter.load -> same Q values; the reference's save -> our load -> same flat parameters."""
    from oracle import nets as onets
    from simple_distributed_rl_b200.netspec import NetSpec

    sys.path.insert(0, REF)
    try:
        import srl
        from oracle.ref_envs import register_restated_envs

        dqn, rainbow = srl_mod
        register_restated_envs()
        if algo == "dqn":
            cfg = dqn.Config()
            cfg.hidden_block.set((64, 64))
            spec = NetSpec(4, (64, 64), 2, None, False, "dqn")
        else:
            cfg = rainbow.Config(enable_noisy_dense=True)
            spec = NetSpec(4, (512,), 2, "average", True, "rainbow")
        runner = srl.Runner("CartPole-v1", cfg)
        runner.set_device("CPU")
        param = runner.make_parameter()
        mu, sigma = spec.init_params(5)
        path = str(tmp_path / "ours.dat")
        ck.save_file(path, ck.parameter_backup(spec, mu, sigma))
        param.load(path)
        sd = param.q_online.state_dict()
        ours = spec.to_state_dict(mu, sigma)
        assert sorted(sd.keys()) == sorted(ours.keys())  # load_state_dict matches by key, not by order
        for k in sd:
            np.testing.assert_array_equal(sd[k].numpy(), ours[k])
        x = np.random.default_rng(0).normal(size=(9, 4)).astype(np.float32)
        ospec = onets.NetSpec(4, spec.hidden, 2, spec.dueling, spec.noisy)
        if not spec.noisy:
            np.testing.assert_allclose(param.pred_q(x), onets.np_forward(ospec, mu, None, None, x), rtol=1e-5, atol=1e-6)
        # the other direction
        path2 = str(tmp_path / "theirs.dat")
        param.save(path2)
        mu2, sigma2 = ck.parameter_restore(spec, ck.load_file(path2))
        np.testing.assert_array_equal(mu2, mu)
        if sigma is not None:
            np.testing.assert_array_equal(sigma2, sigma)
    finally:
        sys.path.remove(REF)


def _reference_memory_after_training(srl, cfg, steps):
    runner = srl.Runner("Grid", cfg)
    runner.set_device("CPU")
    runner.train(max_steps=steps, enable_progress=False)
    return runner, runner.make_memory()


@pytest.mark.parametrize("algo,M,mem,zip_items", [("dqn", 1, "replay", True), ("rainbow", 3, "proportional", True),
                                                  ("rainbow", 2, "replay", False), ("rainbow", 1, "proportional", False)])
def test_reference_memory_imports_and_reexports_unchanged(srl_mod, algo, M, mem, zip_items):
    """A memory the reference Runner filled on Grid -> ring (E columns) -> items again: every re-exported item is the
    reference's item (states, actions of real steps, rewards, terminated flags, leaf priorities), and the reference's own
    memory + trainer accept the re-export."""
    sys.path.insert(0, REF)
    try:
        import srl

        dqn, rainbow = srl_mod
        if algo == "dqn":
            cfg = dqn.Config()
        else:
            cfg = rainbow.Config(multisteps=M, enable_noisy_dense=False)
        cfg.hidden_block.set((16,))
        cfg.memory.warmup_size = 50
        cfg.memory.capacity = 1000
        cfg.memory.compress = zip_items
        if mem == "proportional":
            cfg.memory.set_proportional()
        runner, memory = _reference_memory_after_training(srl, cfg, 330)
        backup = memory.call_backup()
        prop = mem == "proportional"
        E, R = 4, 100
        ring = ck.memory_restore(backup, E, R, M, 4, 2, prop)
        n_g = ring.valid_rows()[1]
        src_items, src_pri, maxp = ck._ordered_items(backup[0], prop)
        assert n_g == min(len(src_items) // E, R - M + 1) and n_g > 60
        src_items = src_items[len(src_items) - n_g * E:]
        out_items, out_pri = ck.export_items(ring)
        assert len(out_items) == n_g * E
        for a, b in zip(src_items, out_items):
            a = pickle.loads(zlib.decompress(a)) if zip_items else a
            if M == 1:
                np.testing.assert_array_equal(a[0], b[0])
                np.testing.assert_array_equal(a[1], b[1])
                assert list(a[2]) == b[2] and float(np.float32(a[3])) == b[3] and a[4] == b[4]
            else:
                np.testing.assert_array_equal(np.stack([e[0] for e in a]), np.stack([e[0] for e in b]))
                ended = False
                for k in range(1, M + 1):
                    assert float(np.float32(a[k][2])) == b[k][2] and a[k][3] == b[k][3]
                    if not ended:
                        assert list(a[k][1]) == b[k][1]
                    ended = ended or bool(a[k][3]) or (k < M and a[k + 1][3] == 1 and a[k + 1][2] == 0 and np.array_equal(a[k + 1][0], a[k][0]))
        if prop:
            np.testing.assert_array_equal(out_pri, src_pri[len(src_pri) - n_g * E:])
            assert ring.max_priority == maxp
        # the reference consumes the re-export: restore into its memory, sample, run its trainer
        ring_backup = ck.memory_backup(ring, prop, compress=zip_items)
        memory2 = runner.make_memory()
        memory2.call_restore(ring_backup)
        assert memory2.length() == n_g * E
        batches, weights, update_args = memory2.sample()
        assert len(batches) == cfg.batch_size and np.all(np.isfinite(weights))
        trainer = runner.make_trainer()
        trainer.memory = memory2
        before = trainer.train_count
        trainer.train()
        assert trainer.train_count == before + 1
    finally:
        sys.path.remove(REF)


def test_restore_edge_cases():
    """Fewer items than env columns -> an empty ring (nothing sampleable, counters at zero); a ReplayBuffer backup that has wrapped
    (idx in the middle of the list) is read oldest first; more items than the ring holds -> the newest survive."""
    E, R, M = 4, 5, 1

    def item(v):
        return [np.full(2, v, np.float32), np.full(2, v + 0.5, np.float32), [0, 1, 0], float(v), 1, []]

    few = ck.memory_restore([[[item(1), item(2)], 2], None], E, R, M, 3, 2, False)
    assert few.vec_steps == 0 and few.valid_rows() == (0, 0)
    # a wrapped ReplayBuffer of capacity 8: memory[idx:] are the oldest
    mem = [item(v) for v in (8, 9, 2, 3, 4, 5, 6, 7)]
    ring = ck.memory_restore([[mem, 2], None], E, R, M, 3, 2, False)
    n_g = ring.valid_rows()[1]
    assert n_g == 2 and ring.vec_steps == 2
    got = sorted(float(ring.reward[s]) for s in range(n_g * E))
    assert got == [2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0]
    # column c holds the c-th contiguous chunk of the oldest-first stream 2..9: (2,3), (4,5), (6,7), (8,9)
    assert [float(ring.reward[0 * E + c]) for c in range(E)] == [2.0, 4.0, 6.0, 8.0]
    assert [float(ring.reward[1 * E + c]) for c in range(E)] == [3.0, 5.0, 7.0, 9.0]
    # 30 items into a ring of 4 x 5: the newest 20 survive
    big = ck.memory_restore([[[item(v) for v in range(30)], 30], None], E, R, M, 3, 2, False)
    assert big.valid_rows()[1] == R and sorted(float(big.reward[s]) for s in range(R * E)) == [float(v) for v in range(10, 30)]


def test_proportional_backup_of_other_capacity_restores_by_item():
    """ProportionalMemory.backup() written with another capacity (proportional_memory.py:196-205 re-adds item by item): the
    items and their leaf priorities land in the ring, the tree is rebuilt over them."""
    E, R, M = 2, 6, 1
    cap_src, n = 7, 6
    leaves = np.array([0.5, 1.5, 0.25, 2.0, 1.0, 0.75, 0.0])
    tree = ck.build_sum_tree(leaves, cap_src)
    data = [[np.full(2, i, np.float32), np.full(2, i + 1, np.float32), [1, 0], float(i), 1, []] for i in range(n)] + [None]
    ring = ck.memory_restore([[cap_src, 2.0, n, n % cap_src, tree.tolist(), data], None], E, R, M, 2, 2, True)
    assert ring.valid_rows() == (0, 3) and ring.max_priority == 2.0
    items, pri = ck.export_items(ring)
    assert [it[3] for it in items] == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]
    np.testing.assert_array_equal(pri, leaves[:n])


def test_interleaved_trajectories_are_refused_for_multistep_memories():
    """A reference memory filled by two actors whose items interleave (play_mp) is not one trajectory per column: with multisteps > 1
    the ring would rebuild windows across unrelated steps, so the import raises instead of silently stitching them."""
    E, R, M, A, D = 2, 12, 3, 4, 2
    rng = np.random.default_rng(0)
    v = ck.RingView(E, R, M, A, D, vec_steps=R)
    cur = rng.normal(size=(E, D)).astype(np.float32)
    for g in range(R):
        sl = slice(g * E, (g + 1) * E)
        nxt = (cur + 1.0).astype(np.float32)
        v.obs[sl], v.next_obs[sl] = cur, nxt
        v.action[sl] = rng.integers(0, A, size=E)
        v.reward[sl] = rng.normal(size=E).astype(np.float32)
        cur = nxt
    items, _ = ck.export_items(v)                      # env-major: trajectory 0 then trajectory 1
    n = len(items) // 2
    back = ck.memory_restore([[items, 0], None], E, R, M, A, D, False)   # contiguous per column: imports
    assert back.vec_steps == R
    inter = [items[(i % 2) * n + i // 2] for i in range(2 * n)]          # actor 0, actor 1, actor 0, ... as play_mp would add them
    with pytest.raises(ck.DiscontinuousMemoryError):
        ck.memory_restore([[inter, 0], None], E, R, M, A, D, False)
    ck.memory_restore([[inter, 0], None], E, R, M, A, D, False, check_continuity=False)  # explicit opt-out still works
    one_step, _ = ck.export_items(ck.RingView(E, R, 1, A, D, vec_steps=R))
    ck.memory_restore([[one_step[::-1], 0], None], E, R, 1, A, D, False)  # 1-step items are self-contained: any order imports


@pytest.mark.parametrize("M", [1, 3])
def test_invalid_action_masks_round_trip_through_the_reference_item_format(M):
    """ring_invalid <-> the next_invalid_actions lists of the reference's records (dqn.py:229-246, rainbow.py:341-351)."""
    from simple_distributed_rl_b200 import checkpoint as ck

    E, R, A, D = 2, 12, 5, 3
    rng = np.random.default_rng(0)
    ring = ck.RingView(E, R, M, A, D, vec_steps=9)
    n = 9 * E
    ring.obs[:n] = rng.normal(size=(n, D))
    ring.next_obs[:n] = rng.normal(size=(n, D))
    for g in range(8):  # one contiguous trajectory per column
        ring.obs[(g + 1) * E:(g + 2) * E] = ring.next_obs[g * E:(g + 1) * E]
    ring.action[:n] = rng.integers(0, A, n)
    ring.reward[:n] = rng.normal(size=n)
    ring.invalid = np.zeros(ring.capacity, dtype=np.uint32)
    ring.invalid[:n] = rng.integers(0, 1 << A, n)
    items, _ = ck.export_items(ring)
    first = items[0]
    lst = first[5] if M == 1 else first[1][4]
    assert lst == [a for a in range(A) if (int(ring.invalid[0]) >> a) & 1]
    back = ck.memory_restore([list(items), 0], E, R, M, A, D, proportional=False)
    g_lo, n_g = ring.valid_rows()
    assert back.invalid is not None
    np.testing.assert_array_equal(back.invalid[: (n_g + M - 1) * E], ring.invalid[: (n_g + M - 1) * E])
