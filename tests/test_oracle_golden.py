"""The oracle restatements vs. the golden vectors produced by executing the reference (tests/golden/make_golden.py)
and vs. the reference's own known-answer tests (tests/quick/rl/memories/test_priority_memories.py)."""
import collections
import glob
import math
import os

import numpy as np
import pytest

from oracle import envs, nets, philox, sumtree, targets


def test_philox_kat():
    # Random123 known-answer vectors for philox4x32-10
    f = lambda *a: [int(x) for x in philox.philox4x32(*a)]
    assert f(0, 0, 0, 0, 0, 0) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert f(*([0xFFFFFFFF] * 6)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert f(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_uniform_conversions():
    assert philox.u01_f32(0xFFFFFFFF) < 1.0 and philox.u01_f32(0) == 0.0
    assert philox.u01_f64(0xFFFFFFFF, 0xFFFFFFFF) < 1.0 and philox.u01_f64(0, 0) == 0.0


# ---- Grid ----------------------------------------------------------------------------------------------
def test_grid_transitions_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "grid_transitions.npz"))
    spec = envs.GridSpec()
    assert list(g["start"]) == [1, 3] and spec.starts == [(1, 3)]
    assert int(g["trunc_steps"]) == spec.trunc_limit == 51
    seen_slip = set()
    for x, y, a, u, nx, ny, r, done, executed in g["recs"]:
        ea = spec.slip(int(a), float(u))
        assert ea == int(executed)
        seen_slip.add((int(a), ea))
        mx, my = spec.move(int(x), int(y), ea)
        assert (mx, my) == (int(nx), int(ny))
        rr, dd = spec.reward_done(mx, my)
        assert rr == r and int(dd) == int(done)
    assert len(seen_slip) == 12  # every (chosen, executed) pair with non-zero probability was exercised


# ---- CartPole (unpinned vs gymnasium; the polynomial must track libm) --------------------------------------
def test_cartpole_poly_sincos_close_to_libm():
    x = np.linspace(-0.6, 0.6, 4001)
    s = np.array([envs.poly_sin(v) for v in x])
    c = np.array([envs.poly_cos(v) for v in x])
    assert np.max(np.abs(s - np.sin(x))) < 4e-16
    assert np.max(np.abs(c - np.cos(x))) < 4e-16


def test_cartpole_physics_matches_textbook_euler():
    """Independent restatement of the published gymnasium CartPole equations with libm sin/cos."""
    spec = envs.CartPoleSpec()
    rng = np.random.default_rng(0)
    for _ in range(200):
        st = rng.uniform(-0.2, 0.2, 4)
        a = int(rng.integers(0, 2))
        x, xd, th, thd = st
        force = 10.0 if a == 1 else -10.0
        ct, sn = math.cos(th), math.sin(th)
        temp = (force + 0.05 * thd * thd * sn) / 1.1
        thacc = (9.8 * sn - ct * temp) / (0.5 * (4.0 / 3.0 - 0.1 * ct * ct / 1.1))
        xacc = temp - 0.05 * thacc * ct / 1.1
        want = np.array([x + 0.02 * xd, xd + 0.02 * xacc, th + 0.02 * thd, thd + 0.02 * thacc])
        got, r, term = spec.step(st, a)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-14)
        assert r == 1.0
        assert term == bool(abs(want[0]) > 2.4 or abs(want[2]) > 12 * 2 * math.pi / 360)


# ---- SumTree / ProportionalMemory -------------------------------------------------------------------------
def test_sumtree_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "sumtree.npz"))
    for case in range(int(g["n_cases"])):
        cap, alpha, beta0, bsteps, dup = g[f"c{case}_cfg"]
        mem = sumtree.ProportionalMemory(int(cap), alpha, beta0, bsteps, has_duplicate=bool(dup))
        for p, none in zip(g[f"c{case}_add_pri"], g[f"c{case}_add_none"]):
            mem.add(None if none else float(p))
        np.testing.assert_array_equal(mem.tree.tree, g[f"c{case}_tree_after_add"])
        assert mem.max_priority == float(g[f"c{case}_maxp_after_add"])
        for it, step in enumerate(g[f"c{case}_steps"]):
            us = g[f"c{case}_uniforms"][it]
            cursor = {"n": 0}

            def uniforms(i, k):
                u = us[cursor["n"]]
                cursor["n"] += 1
                return float(u)

            idx, w, pri, _ = mem.sample(5, int(step), uniforms)
            np.testing.assert_array_equal(idx, g[f"c{case}_idx"][it])
            np.testing.assert_allclose(w, g[f"c{case}_weights"][it], rtol=1e-15)
            mem.update(idx, g[f"c{case}_upd"][it])
            np.testing.assert_array_equal(mem.tree.tree, g[f"c{case}_trees"][it])
            assert mem.max_priority == float(g[f"c{case}_maxp"][it])
        assert mem.size == int(g[f"c{case}_size"]) and mem.tree.write == int(g[f"c{case}_write"])


@pytest.mark.parametrize("alpha", [0, 0.2, 0.5, 0.8, 1.0])
def test_IS_Proportional_kat(alpha):
    """tests/quick/rl/memories/test_priority_memories.py:97-117,150-176 restated for the oracle."""
    epsilon = 0.0001
    mem = sumtree.ProportionalMemory(10, alpha=alpha, beta_initial=1, epsilon=epsilon, has_duplicate=False)
    priorities = [1, 2, 4, 3]
    true_p = [(t + epsilon) ** alpha for t in priorities]
    N = 4
    probs = [p / sum(true_p) for p in true_p]
    tw = np.array([(N * p) ** -1 for p in probs])
    tw /= tw.max()
    for p in priorities:
        mem.add(p)
    idx, w, _, _ = mem.sample(N, 1, sumtree.philox_uniforms(3, 0))
    for i, ti in enumerate(idx):
        assert math.isclose(w[i], tw[ti - 9], rel_tol=1e-7)
    assert len(set(idx.tolist())) == N


def test_priority_memory_monotone_counts():
    """tests/quick/rl/memories/test_priority_memories.py:17-91 (2 000 iterations instead of 20 000 for CPU time)."""
    cap = 10
    mem = sumtree.ProportionalMemory(cap, 0.8, 1, 10, has_duplicate=False)
    for i in range(100):
        mem.add(0)
    assert mem.length() == cap
    for i in range(1, 11):
        mem.add(i)
    counter = collections.Counter()
    for it in range(2000):
        idx, w, pri, _ = mem.sample(5, 1, sumtree.philox_uniforms(11, it))
        assert len(set(idx.tolist())) == 5
        for j in idx:
            counter[int(j) - (cap - 1)] += 1
        mem.update(idx, np.array([((int(j) - (cap - 1)) - mem.tree.write) % cap + 1 for j in idx], dtype=np.float64))
    slot_pri = {(mem.tree.write + k) % cap: k + 1 for k in range(cap)}
    vals = [counter[s] for s, _ in sorted(slot_pri.items(), key=lambda kv: kv[1])]
    assert all(vals[i] < vals[i + 1] for i in range(cap - 1)), vals


def test_uniform_sample_distinct():
    idx = sumtree.uniform_sample_distinct(40, 32, 5, 9)
    assert len(set(idx.tolist())) == 32 and idx.min() >= 0 and idx.max() < 40


# ---- functions ---------------------------------------------------------------------------------------------
def test_rescaling_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "functions.npz"))
    np.testing.assert_array_equal(targets.rescaling(g["x"]), g["rescaling"])
    np.testing.assert_array_equal(targets.inverse_rescaling(g["x"]), g["inverse_rescaling"])
    x = g["x"].astype(np.float64)
    np.testing.assert_allclose(targets.inverse_rescaling(targets.rescaling(x)), x, rtol=1e-3, atol=1e-6)


# ---- trainer update ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "trainer_*.npz"))))
def test_trainer_update_matches_reference(path):
    g = np.load(path)
    algo = str(g["algo"])
    dueling = None if str(g["dueling"]) == "none" else str(g["dueling"])
    noisy = bool(g["noisy"])
    spec = nets.NetSpec(2, tuple(int(h) for h in g["hidden"]), 4, dueling, noisy)
    assert spec.n_params == len(g["mu0"])
    sigma0 = g["sigma0"] if noisy else None
    st = nets.AdamState(spec, g["mu0"], sigma0, lr=float(g["lr"]))
    tmu = g["tmu0"].copy()
    tsig = g["tsigma0"].copy() if noisy else None
    mask = spec.sigma_mask(algo) if noisy else None
    for u in range(len(g["losses"])):
        noise = tuple(g["noise"][u]) if noisy else (None, None, None)
        res = nets.train_update(
            spec, st, tmu, tsig, algo=algo, states=g["states"][u], actions=g["actions"][u], rewards=g["rewards"][u],
            dones=g["terms"][u], weights=g["weights"][u], discount=float(g["discount"]), multisteps=int(g["multisteps"]),
            retrace_h=float(g["retrace_h"]), enable_double_dqn=bool(g["double"]), enable_rescale=bool(g["rescale"]),
            noise=noise, sigma_mask=mask, next_invalid=g["invalid"][u] if "invalid" in g.files and g["invalid"].any() else None)
        np.testing.assert_allclose(res["target_q"], g["target_q"][u], rtol=1e-6, atol=1e-6)
        assert abs(res["loss"] - g["losses"][u]) <= 1e-6 * max(1, abs(g["losses"][u]))
        np.testing.assert_allclose(res["priorities"], g["priorities"][u], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(st.mu.detach().numpy(), g["mu_after"][u], rtol=1e-5, atol=1e-7)
        if noisy:
            np.testing.assert_allclose(st.sigma.detach().numpy(), g["sigma_after"][u], rtol=1e-5, atol=1e-7)
        if u % 1000 == 0:  # first update syncs the target (train_count % interval == 0 before the increment)
            tmu = st.mu.detach().numpy().copy()
            tsig = st.sigma.detach().numpy().copy() if noisy else None
        np.testing.assert_allclose(tmu, g["tmu_after"][u], rtol=1e-5, atol=1e-7)
    assert list(g["update_steps"]) == list(range(len(g["losses"])))


# ---- the whole learner step on a frozen memory: the reference's own sample / Trainer.train / update -------------------------
import learner_cases  # noqa: E402


@pytest.mark.parametrize("path", learner_cases.PATHS, ids=learner_cases.IDS)
def test_oracle_learner_step_matches_reference_on_frozen_memory(path):
    """oracle/engine.py::OracleEngine.learn (sampler + window rebuild + targets + loss + Adam + priority update) against the
    reference's ProportionalMemory.sample / ReplayBuffer + Trainer.train + memory.update executed on the same memory
    (tests/golden/make_learner_golden.py): leaf indices exact, everything else to float32 resolution."""
    from oracle import engine as oeng

    kw, v, g = learner_cases.load_case(path)
    noisy = kw["noisy"]
    orc = oeng.OracleEngine(oeng.EngineConfig(**kw), g["mu0"], g["sigma0"] if noisy else None)
    orc.tgt_mu = g["tmu0"].copy()
    orc.tgt_sigma = g["tsigma0"].copy() if noisy else None
    orc.load_ring(v)
    cap = orc.cap
    for u in range(len(g["loss"])):
        o = orc.learn(1)[0]
        np.testing.assert_array_equal(o["idx"], g["idx"][u])
        np.testing.assert_allclose(o["weights"], g["weights"][u], rtol=1e-6)
        np.testing.assert_allclose(o["target_q"], g["target_q"][u], rtol=1e-5, atol=1e-6)
        assert abs(o["loss"] - g["loss"][u]) <= 1e-5 * max(1, abs(g["loss"][u]))
        np.testing.assert_allclose(o["priorities"], g["td"][u], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(orc.mu, g["mu_after"][u], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(orc.tgt_mu, g["tmu_after"][u], rtol=1e-5, atol=1e-6)
        if noisy:
            np.testing.assert_allclose(orc.sigma, g["sigma_after"][u], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(orc.tgt_sigma, g["tsigma_after"][u], rtol=1e-5, atol=1e-6)
        if kw["mem_kind"]:
            # the reference forms (|td| + eps) ** alpha on the float32 array the trainer hands over (proportional_memory.py:172)
            np.testing.assert_allclose(orc.per.tree.tree[cap - 1:][g["touched"]], g["leaves_after"][u], rtol=1e-5, atol=1e-7)
            assert abs(orc.per.max_priority - g["maxp_after"][u]) <= 1e-5 * g["maxp_after"][u]
            np.testing.assert_allclose(orc.per.tree.total(), g["total_after"][u], rtol=1e-7)
    assert list(g["update_step"]) == list(range(len(g["loss"])))


# ------------------------------------------------------------------------------------------------------------------
# The vectorised engine adds a whole row of E leaves to the SumTree at once (oracle/engine.py::_tree_set_range, device twin
# csrc/rollout.cu::tree_set_row).  Its result has to equal the reference's one-leaf-at-a-time update
# (proportional_memory.py:85-90, restated and golden-pinned in oracle/sumtree.py) up to fp64 association -- including rows
# that straddle the two leaf depths of a non-power-of-two tree, where the index range spans two levels.
@pytest.mark.parametrize("E,R", [(3, 11), (5, 7), (48, 12), (33, 5), (40, 6), (1000, 37), (7, 1), (64, 4)])
def test_bulk_row_set_equals_sequential_reference_updates(E, R):
    from types import SimpleNamespace

    from oracle import engine as oeng
    from oracle import sumtree

    cap = E * R
    bulk = SimpleNamespace(cap=cap, per=SimpleNamespace(tree=sumtree.SumTree(cap)))
    seq = sumtree.SumTree(cap)
    rng = np.random.default_rng(E * 1000 + R)
    for step in range(2 * R + 3):  # wraps the ring twice: rows are overwritten
        row = step % R
        vals = rng.uniform(0.0, 2.0, size=E) if step % 3 else np.zeros(E)
        oeng.OracleEngine._tree_set_range(bulk, row * E, vals)
        for j in range(E):
            seq.update(row * E + j + cap - 1, float(vals[j]))
        np.testing.assert_allclose(bulk.per.tree.tree, seq.tree, rtol=1e-12, atol=1e-9)  # sums of up to E x 2.0
    tree = bulk.per.tree.tree
    if cap > 1:
        np.testing.assert_allclose(tree[: cap - 1], tree[1::2][: cap - 1] + tree[2::2][: cap - 1], rtol=1e-12, atol=1e-9)


# ------------------------------------------------------------------------------------------------------------------
# Worker-side records.  tests/golden/worker_records.npz holds, for real trajectories of the reference Runner on
# Grid, every batch rainbow.Worker._add_batch / dqn.Worker.on_step handed to memory.add().  The vectorised engine stores one
# record per step and rebuilds the n-step window by index at sample time (OracleEngine.window == the device gather): the
# rebuilt windows must be the reference's batches -- which steps start a window, where an episode cuts it, and how the
# tail is padded (repeated last state, reward 0, terminated 1; the padded ACTION is a fresh random draw on both sides and is
# only checked for range) (srl/algorithms/rainbow/rainbow.py:341-400).
@pytest.mark.parametrize("name,M,clip", [("rainbow_m3", 3, False), ("rainbow_m2_clip", 2, True)])
def test_window_rebuild_equals_reference_worker_batches(name, M, clip, golden_dir):
    from oracle import engine as oeng
    from oracle import nets as onets

    d = np.load(os.path.join(golden_dir, "worker_records.npz"))
    s, a, r, ns = d[f"{name}_s"], d[f"{name}_a"], d[f"{name}_r"], d[f"{name}_ns"]
    term, done = d[f"{name}_term"], d[f"{name}_done"]
    T = len(a)
    cfg = oeng.EngineConfig(env="Grid", algo="rainbow", hidden=(16,), dueling="average", noisy=False, mem_kind=0, multisteps=M,
                            n_envs=1, ring_rows=T + M, batch_size=4, warmup_size=1, enable_reward_clip=clip)
    spec = onets.NetSpec(2, (16,), 4, "average", False)
    mu = np.zeros(spec.n_params, dtype=np.float32)  # the window rebuild never looks at the network
    orc = oeng.OracleEngine(cfg, mu, None)
    rr = np.sign(r) if clip else r  # the engine stores the reward after the worker's clip (rainbow.py:343-350)
    orc.ring_obs[:T], orc.ring_next_obs[:T] = s, ns
    orc.ring_action[:T], orc.ring_reward[:T] = a, rr.astype(np.float32)
    orc.ring_term[:T], orc.ring_done[:T] = term, done
    orc.vec_steps = T
    b_states, b_a, b_r, b_term = d[f"{name}_b_states"], d[f"{name}_b_a"], d[f"{name}_b_r"], d[f"{name}_b_term"]
    n = len(b_a)
    assert T - M <= n <= T  # every real step starts exactly one window; the windows of the last < M+1 steps are still open
    n_padded = 0
    for j in range(n):  # the reference emits the windows in the order of their first step
        states, acts, rews, terms = orc.window(j)
        np.testing.assert_array_equal(states, b_states[j])
        np.testing.assert_array_equal(terms.astype(np.int64), b_term[j])
        np.testing.assert_allclose(rews, b_r[j].astype(np.float32), rtol=0, atol=0)
        ended = False
        for k in range(M):
            if not ended:
                assert acts[k] == b_a[j][k]
                ended = bool(done[j + k])
            else:
                n_padded += 1
                assert 0 <= acts[k] < 4 and 0 <= b_a[j][k] < 4 and rews[k] == 0.0 and terms[k] == 1.0
    assert n_padded > 0  # the log contains finished episodes


def test_dqn_record_equals_reference_worker_batches(golden_dir):
    """dqn.Worker.on_step (dqn.py:229-246): one record per step, undone = not TERMINATED (truncation still bootstraps)."""
    d = np.load(os.path.join(golden_dir, "worker_records.npz"))
    n = len(d["dqn_b_a"])
    np.testing.assert_array_equal(d["dqn_b_s"], d["dqn_s"][:n])
    np.testing.assert_array_equal(d["dqn_b_ns"], d["dqn_ns"][:n])
    np.testing.assert_array_equal(d["dqn_b_a"], d["dqn_a"][:n])
    np.testing.assert_array_equal(d["dqn_b_r"], d["dqn_r"][:n])
    np.testing.assert_array_equal(d["dqn_b_undone"], 1 - d["dqn_term"][:n])  # == 1 - ring_term, what the learner multiplies gamma with


def test_grid_truncation_step_matches_reference_envrun(golden_dir):
    """EnvRun truncates when step_num > max_episode_steps (env_run.py:361): on Grid (max 50) every truncated episode of the
    reference log is 51 steps long and no episode is longer -- the engine's trunc_limit (oracle/envs.py, envspec.py)."""
    from oracle import envs as oenvs

    d = np.load(os.path.join(golden_dir, "worker_records.npz"))
    limit = oenvs.make_spec("Grid").trunc_limit
    seen_truncated = False
    for name in ("rainbow_m3", "rainbow_m2_clip", "dqn"):
        done, term = d[f"{name}_done"], d[f"{name}_term"]
        cur = 0
        for i in range(len(done)):
            cur += 1
            assert cur <= limit
            if done[i]:
                if not term[i]:
                    assert cur == limit
                    seen_truncated = True
                cur = 0
    assert seen_truncated


def test_action_division_table_matches_reference_boxspace(golden_dir):
    """Pendulum-v1 on the value-based path: the discrete action set is the reference's BoxSpace division table
    (srl/base/spaces/box.py:317-366), float32 arithmetic included -- both the oracle's and the host spec's restatement."""
    from oracle import envs as oenvs
    from simple_distributed_rl_b200 import envspec

    d = np.load(os.path.join(golden_dir, "spaces.npz"))
    assert int(d["default_action_division_num"][0]) == 10
    for n in (2, 3, 5, 10, 16):
        want = d[f"pendulum_div{n}"]
        np.testing.assert_array_equal(oenvs.division_table(-2.0, 2.0, n), want)
        np.testing.assert_array_equal(np.array(envspec.division_table(-2.0, 2.0, n), dtype=np.float32), want)
    assert oenvs.make_spec("Pendulum-v1").n_actions == 10


def test_pendulum_restatement_properties():
    """gymnasium Pendulum-v1 restated (parity vs gymnasium unpinned): observation = (cos, sin, thdot) of the state within
    1e-6 of libm, speed clipped to 8, reward = -(angle_normalize(th)^2 + .1 thdot^2 + .001 u^2) in [-16.2736, 0]."""
    import math

    from oracle import envs as oenvs

    sp = oenvs.make_spec("Pendulum-v1")
    st = sp.reset(7, 3, 0)
    assert -math.pi <= st[0] < math.pi and -1.0 <= st[1] < 1.0
    for t in range(400):
        a = (t * 7) % 10
        th, thdot = st[0], st[1]
        st2, r, term = sp.step(st, a)
        u = float(sp.action_table[a])
        an = ((th + math.pi) % (2 * math.pi)) - math.pi
        want_r = -(an * an + 0.1 * thdot * thdot + 0.001 * u * u)
        assert abs(r - want_r) < 1e-9 and -16.2736044 <= r <= 0.0 and term is False
        want_thdot = min(max(thdot + (15.0 * math.sin(th) + 3.0 * u) * 0.05, -8.0), 8.0)
        assert abs(st2[1] - want_thdot) < 1e-9 and abs(st2[0] - (th + want_thdot * 0.05)) < 1e-9
        o = sp.obs(st2)
        assert abs(o[0] - math.cos(st2[0])) < 1e-6 and abs(o[1] - math.sin(st2[0])) < 1e-6 and o[2] == np.float32(st2[1])
        st = st2


def test_epsilon_schedule_is_the_reference_linear_phase():
    """oracle.engine.epsilon_at restates Linear.update(step).to_float() (srl/rl/schedulers/schedulers/linear.py:11-21):
    start - ((start - end) / phase) * step below the phase, end_rate from `phase` on, and it only applies to training steps.
    (tests/test_srl_plugin.py compares it with the imported reference class value for value where the reference is present.)"""
    from oracle import engine as oeng
    from oracle import nets as onets

    cfg = oeng.EngineConfig(env="Grid", algo="dqn", hidden=(8,), mem_kind=0, multisteps=1, n_envs=64, ring_rows=4, batch_size=4,
                            warmup_size=8, epsilon=1.0, eps_end=0.0, eps_phase_steps=4)
    spec = onets.NetSpec(2, (8,), 4, None, False)
    rng = np.random.default_rng(0)
    mu = rng.normal(size=spec.n_params).astype(np.float32)
    orc = oeng.OracleEngine(cfg, mu, None)
    assert [orc.epsilon_at(s) for s in (0, 1, 2, 3, 4, 9)] == [1.0, 0.75, 0.5, 0.25, 0.0, 0.0]
    greedy = []
    for s in range(6):
        res = orc.vec_step()
        greedy.append(float(np.mean(res["actions"] == np.argmax(res["q"], axis=1))))
    assert greedy[4] == 1.0 and greedy[5] == 1.0  # epsilon 0 from step `phase` on: every action is the argmax
    assert greedy[0] < 0.6  # epsilon 1 at step 0: uniformly random over 4 actions (argmax hit ~ 1/4)
    c2 = oeng.EngineConfig(**{**cfg.__dict__, "eps_phase_steps": 0, "epsilon": 0.3})
    assert oeng.OracleEngine(c2, mu, None).epsilon_at(123) == 0.3


# ------------------------------------------------------------------------------------------------------------------
# PPO worker-side returns (the worker half): tests/golden/ppo_returns.npz = the "discounted_reward" values the
# reference's ppo.Worker.on_step handed to memory.add() over six episodes (lengths 1..200), GAE and MC, with and without clip.
@pytest.mark.parametrize("name", ["gae_g0.9_l0.9_noclip", "gae_g0.99_l0.95_clip", "mc_g0.9_l0.9_noclip", "mc_g0.997_l0.9_clip"])
def test_ppo_returns_equal_reference_worker(name, golden_dir):
    from oracle import gae as ogae

    d = np.load(os.path.join(golden_dir, "ppo_returns.npz"))
    discount, lam, lo, hi = d[f"{name}_params"]
    clip = None if np.isnan(lo) else (lo, hi)
    method = ogae.METHOD_GAE if name.startswith("gae") else ogae.METHOD_MC
    reward = ogae.clip_reward(d[f"{name}_reward"], clip)
    T = len(reward)
    out, valid = ogae.returns_scan(reward.reshape(T, 1), d[f"{name}_v"].reshape(T, 1), d[f"{name}_nv"].reshape(T, 1),
                                   d[f"{name}_done"].reshape(T, 1), float(discount), float(lam), method)
    assert valid.all() and d[f"{name}_done"][-1] == 1
    np.testing.assert_array_equal(out[:, 0], d[f"{name}_ret"])  # float32 bit for bit


# ------------------------------------------------------------------------------------------------------------------
# R2D2 trainer, per-sequence target loop (the target / Retrace / priority half): tests/golden/r2d2_targets.npz =
# what r2d2.Trainer._train_on_batches computed for six 80-step sequences (reference source executed over TensorFlow stubs).
R2D2_CASES = ["double_retrace", "plain", "double_rescale_retrace", "target_rescale"]


@pytest.mark.parametrize("name", R2D2_CASES)
def test_r2d2_sequence_targets_equal_reference_trainer(name, golden_dir):
    from oracle import r2d2_targets as o

    d = np.load(os.path.join(golden_dir, "r2d2_targets.npz"))
    double, rescale, retrace, h, disc = [float(x) for x in d[f"{name}_params"]]
    tgt, td_mean = o.batch_targets(d[f"{name}_q_on"], d[f"{name}_q_tg"], d[f"{name}_actions"], d[f"{name}_mu"].tolist(),
                                   d[f"{name}_rewards"].tolist(), d[f"{name}_dones"], disc, h, bool(double), bool(rescale), bool(retrace))
    np.testing.assert_array_equal(tgt, d[f"{name}_target"])  # float64, bit for bit
    want = d[f"{name}_td_mean"]
    assert str(np.asarray(td_mean).dtype) == str(d[f"{name}_td_mean_dtype"])
    np.testing.assert_array_equal(np.asarray(td_mean), want)
    # np.mean of the TD errors is numpy's pairwise sum / n: the restated summation order reproduces it exactly
    tds32 = np.random.default_rng(0).normal(size=80).astype(np.float32)
    assert o.np_pairwise_sum(tds32) == np.add.reduce(tds32) and o.np_pairwise_sum(tds32.astype(np.float64)) == np.add.reduce(tds32.astype(np.float64))


# ---- rank-based replay: oracle/rankbased.py vs the reference's RankBasedMemory ------------------------------------
def test_rankbased_oracle_matches_reference_golden(golden_dir):
    """add / sample / update sequences of the reference class (np.random seeded) replayed on the stored uniform stream: sampled item
    indices exact, float64 IS weights to 1e-12, priorities after the updates exact."""
    from oracle import rankbased as orb

    g = np.load(os.path.join(golden_dir, "rankbased.npz"))
    for c in range(int(g["n_cases"])):
        cap, alpha, beta0, bsteps, B, n_add = g[f"c{c}_cfg"]
        cap, B, n_add = int(cap), int(B), int(n_add)
        m = orb.RankBasedMemory(cap, alpha, beta0, bsteps)
        for p in g[f"c{c}_pri0"]:
            m.add(float(p))
        for step in range(len(g[f"c{c}_idx"])):
            idx, w, ranks, used = m.sample(B, step * 7, g[f"c{c}_u"][step])
            np.testing.assert_array_equal(idx, g[f"c{c}_idx"][step])
            np.testing.assert_allclose(w, g[f"c{c}_w"][step], rtol=1e-12)
            assert len(set(idx.tolist())) == B and B <= used <= 4 * B
            m.update(idx, g[f"c{c}_upd"][step])
        np.testing.assert_array_equal(m.priorities, g[f"c{c}_pri_final"])


def test_rankbased_choice_restatement_equals_numpy():
    """oracle.rankbased.choice_without_replacement == np.random.choice(n, size, replace=False, p) for the same MT19937 stream, including
    draws that need several rounds (few heavy ranks: the first round repeats indices)."""
    from oracle import rankbased as orb

    for seed, n, size, alpha in [(0, 50, 20, 2.0), (1, 400, 64, 1.5), (2, 33, 33, 0.3), (3, 1000, 10, 0.0)]:
        p = (1.0 / np.arange(1, n + 1)) ** alpha
        p /= p.sum()
        np.random.seed(seed)
        want = np.random.choice(n, size=size, replace=False, p=p)
        got, used = orb.choice_without_replacement(p, size, np.random.RandomState(seed).random_sample(40 * size))
        np.testing.assert_array_equal(got, want)
        assert used >= size


@pytest.mark.parametrize("E,R", [(2, 2), (8, 5), (64, 12), (1024, 37), (256, 256), (4, 3)])
def test_constant_row_set_on_complete_subtrees_equals_sequential_reference_updates(E, R):
    """Power-of-two E (every BASELINE configuration): a ring row is one complete subtree, set directly (node = value x leaves below)
    with the root's change added to its ancestors -- oracle/engine.py::_tree_set_const, device twin post_step_pow2_kernel.  Against
    the reference's one-leaf-at-a-time SumTree.update (proportional_memory.py:85-90) over a ring that wraps, n-step style (zero the
    row being overwritten, add an older one)."""
    from types import SimpleNamespace

    from oracle import engine as oeng
    from oracle import sumtree

    cap = E * R
    bulk = SimpleNamespace(cap=cap, per=SimpleNamespace(tree=sumtree.SumTree(cap)))
    bulk._tree_set_range = lambda lo, vals: oeng.OracleEngine._tree_set_range(bulk, lo, vals)
    seq = sumtree.SumTree(cap)
    rng = np.random.default_rng(E * 100 + R)
    maxp = 1.0
    for step in range(3 * R + 2):
        for row, val in ((step % R, 0.0), ((step - 2) % R, maxp)):
            oeng.OracleEngine._tree_set_const(bulk, row * E, E, val)
            for j in range(E):
                seq.update(row * E + j + cap - 1, float(val))
            # (the reference's E sequential += carry up to E ulps themselves; the direct set is the more accurate of the two)
            np.testing.assert_allclose(bulk.per.tree.tree, seq.tree, rtol=1e-11, atol=1e-9)
        # some leaves get individual priorities in between (the trainer's updates), max_priority grows
        for _ in range(5):
            leaf = int(rng.integers(0, cap))
            if seq.tree[leaf + cap - 1] > 0:
                p = float(rng.uniform(0.01, 3.0))
                seq.update(leaf + cap - 1, p)
                sumtree.SumTree.update(bulk.per.tree, leaf + cap - 1, p)
                maxp = max(maxp, p)
    tree = bulk.per.tree.tree
    np.testing.assert_allclose(tree[: cap - 1], tree[1::2][: cap - 1] + tree[2::2][: cap - 1], rtol=1e-12, atol=1e-9)
