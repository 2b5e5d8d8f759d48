"""GPU tests of R2D2 on the device (csrc/r2d2.cu, simple_distributed_rl_b200/r2d2.py) against oracle/r2d2.py -- a torch RESTATEMENT
of the reference's TensorFlow code (srl/algorithms/r2d2/r2d2.py; TensorFlow is not available, so the network / optimiser parity of this
row is by restatement; the trainer's per-sequence target loop is pinned separately against goldens from the reference's own loop,
the keras LSTM restatement against torch.nn.LSTM in tests/test_r2d2_cpu.py) -- and the reference's own acceptance gate
(tests/algorithms_/base_r2d2.py:29-44)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import envs as oenvs  # noqa: E402
from oracle import philox  # noqa: E402
from oracle import r2d2 as orc  # noqa: E402
from oracle import sumtree as osum  # noqa: E402


def _cfg(**kw):
    from simple_distributed_rl_b200.r2d2 import R2D2Config

    base = dict(env="CartPole-v1", n_envs=6, lstm_units=16, hidden_layers=(12,), dueling_type="average", burnin=2, sequence_length=3,
                batch_size=8, warmup_size=8, capacity=6 * 64, seed=11, epsilon=0.3, lr=1e-3, target_model_update_interval=2)
    base.update(kw)
    return R2D2Config(**base)


def _engine(cfg, **kw):
    from simple_distributed_rl_b200.r2d2 import R2D2Engine

    return R2D2Engine(cfg, debug=True, **kw)


def _n_hidden(cfg):
    return len(cfg.hidden_layers) if cfg.dueling_type is None else len(cfg.hidden_layers) - 1


@pytest.mark.parametrize("hidden,dueling", [((16, 16), None), ((12,), "average"), ((8, 12), "max"), ((12,), "")])
def test_forward_equals_the_restated_qnetwork(hidden, dueling):
    cfg = _cfg(hidden_layers=hidden, dueling_type=dueling, env="Pendulum-v1", lstm_units=24)
    eng = _engine(cfg)
    rng = np.random.default_rng(0)
    w = [x + rng.normal(0, 0.05, x.shape).astype(np.float32) for x in eng.get_weights()]
    eng.set_weights(w, target_too=False)
    net = orc.QNet(w, _n_hidden(cfg), dueling)
    n = cfg.batch_size
    x, h, c = (rng.normal(size=s).astype(np.float32) for s in ((n, eng.D), (n, eng.u), (n, eng.u)))
    q, h2, c2 = eng.forward(x, h, c)
    with torch.no_grad():
        ho, co = net.step(torch.as_tensor(x), torch.as_tensor(h), torch.as_tensor(c))
        qo = net.head(ho)
    np.testing.assert_allclose(h2.cpu().numpy(), ho.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(c2.cpu().numpy(), co.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(q.cpu().numpy(), qo.numpy(), rtol=1e-4, atol=1e-5)
    qt, _, _ = eng.forward(x, h, c, use_target=True)  # the target network still holds the initial weights
    assert not np.allclose(qt.cpu().numpy(), q.cpu().numpy())


class _Twin:
    """The oracle side of a lockstep rollout: env twins, the reference's policy helpers on the device's Q values (so that action
    indices are comparable bit for bit), one WorkerLists per env copy."""

    def __init__(self, eng):
        cfg = eng.cfg
        self.eng, self.cfg = eng, cfg
        self.env = oenvs.make_spec(cfg.env, **cfg.env_kwargs)
        E = eng.E
        self.state = [None] * E
        self.needs_reset = [True] * E
        self.episode = [0] * E
        self.step_num = [0] * E
        self.ctx = [dict(ep_start=0, c0=0) for _ in range(E)]
        self.workers = [orc.WorkerLists(cfg.burnin, eng.S, eng.D, eng.u, eng.A, self._rand(e)) for e in range(E)]
        self.g = 0
        self.net = None

    def _rand(self, e):
        S, A, seed = self.eng.S, self.eng.A, self.cfg.seed

        def f(kind, j):
            ctx = self.ctx[e]
            pos, tag = (ctx["ep_start"] - S + j, 1) if kind == "reset" else (ctx["c0"] + 1 + j, 0)
            w = philox.words(seed, philox.STREAM_PAD_ACTION, e, pos & 0xFFFFFFFF, tag)
            return (int(w[0]) * A) >> 32

        return f

    def step(self, check_net=True):
        eng, cfg, env, E = self.eng, self.cfg, self.env, self.eng.E
        g = self.g
        h_before = np.zeros((E, eng.u), np.float32)
        c_before = np.zeros((E, eng.u), np.float32)
        prev_h = eng.t["roll_h"].cpu().numpy()[:, :eng.u]
        prev_c = eng.t["roll_c"].cpu().numpy()
        obs = np.zeros((E, eng.D), np.float32)
        for e in range(E):
            if self.needs_reset[e]:
                self.state[e] = env.reset(cfg.seed, e, self.episode[e])
                self.episode[e] += 1
                self.step_num[e] = 0
                self.needs_reset[e] = False
                self.ctx[e]["ep_start"] = len(self.workers[e].items)
                self.workers[e].on_reset(env.obs(self.state[e]))
            else:
                h_before[e], c_before[e] = prev_h[e], prev_c[e]
            obs[e] = env.obs(self.state[e])
        eng.vec_step(True)
        np.testing.assert_array_equal(eng.t["roll_xh"].cpu().numpy()[:, :eng.D], obs)  # env transitions + observation encoding, exact
        q_dev = eng.t["dbg_q"].cpu().numpy()
        a_dev = eng.t["dbg_action"].cpu().numpy()
        h_dev = eng.t["roll_h"].cpu().numpy()[:, :eng.u].copy()
        c_dev = eng.t["roll_c"].cpu().numpy().copy()
        if check_net:
            net = orc.QNet(eng.get_weights(), _n_hidden(cfg), cfg.dueling_type)
            with torch.no_grad():
                ho, co = net.step(torch.as_tensor(obs), torch.as_tensor(h_before), torch.as_tensor(c_before))
                qo = net.head(ho)
            np.testing.assert_allclose(h_dev, ho.numpy(), rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(c_dev, co.numpy(), rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(q_dev, qo.numpy(), rtol=1e-4, atol=1e-5)
        for e in range(E):
            probs = orc.policy_probs(q_dev[e], cfg.epsilon)
            w = philox.words(cfg.seed, philox.STREAM_POLICY, e, g & 0xFFFFFFFF, g >> 32)
            a = orc.choice(probs, float(philox.u01_f64(w[0], w[1])))
            assert a == a_dev[e], (g, e)
            nst, r, terminated = env.step(self.state[e], a, cfg.seed, e, g)
            self.step_num[e] += 1
            truncated = self.step_num[e] >= env.trunc_limit
            if env.trunc_overrides_term:
                terminated = terminated and not truncated
            else:
                truncated = truncated and not terminated
            done = bool(terminated or truncated)
            self.ctx[e]["c0"] = len(self.workers[e].items)
            self.workers[e].on_step(a, probs[a], (r + cfg.reward_shift) * cfg.reward_scale, bool(terminated), done, env.obs(nst),
                                    (h_dev[e], c_dev[e]))
            self.state[e] = nst
            if done:
                self.needs_reset[e] = True
        self.g += 1

    def item_of(self, sel):
        """device batch entry (tree index or slot) -> (env, item)"""
        eng = self.eng
        E, R = eng.E, eng.R
        if eng.per:  # tree leaves are env-major (csrc/r2d2.cu: ring_leaf)
            leaf = int(sel) - (R * E - 1)
            e, row = leaf // R, leaf % R
        else:
            e, row = int(sel) % E, int(sel) // E
        cur = int(eng.t["cursor"][e].item())
        assert cur == len(self.workers[e].items)
        pos = row if cur <= R else (cur - 1) - ((cur - 1 - row) % R)
        assert 0 <= pos < cur
        return e, pos, self.workers[e].items[pos]


def _check_batch(eng, twin):
    """the gathered batch of the last update == the items the restated worker added (bit for bit)"""
    t = {k: eng.t[k].cpu().numpy() for k in ("sel", "xh", "cbuf", "b_actions", "b_mu", "b_rewards", "b_dones")}
    D, u, W = eng.D, eng.u, eng.W
    items = []
    for b in range(eng.B):
        e, pos, it = twin.item_of(t["sel"][b])
        items.append(it)
        for z in range(2):
            np.testing.assert_array_equal(t["xh"][z, :W + 1, b, :D], it["states"], err_msg=f"b={b} e={e} pos={pos}")
        np.testing.assert_array_equal(t["xh"][0, 0, b, D:D + u], it["hidden_states"][0])
        np.testing.assert_array_equal(t["cbuf"][0, 0, b], it["hidden_states"][1])
        np.testing.assert_array_equal(t["cbuf"][1, 0, b], it["hidden_states"][1])
        assert list(t["b_actions"][b]) == it["actions"], (b, e, pos)
        assert list(t["b_mu"][b]) == it["probs"]
        assert list(t["b_rewards"][b]) == it["rewards"]
        assert list(t["b_dones"][b].astype(bool)) == it["dones"]
    return items


def _oracle_update(eng, tr, items, weights):
    out = tr.train_on_batches(np.stack([it["states"] for it in items]), [it["actions"] for it in items], [it["probs"] for it in items],
                              [it["rewards"] for it in items], [it["dones"] for it in items],
                              np.stack([it["hidden_states"][0] for it in items]), np.stack([it["hidden_states"][1] for it in items]), weights)
    return out


def _trainer(eng, w):
    c = eng.cfg
    return orc.Trainer(w, _n_hidden(c), c.dueling_type, c.burnin, c.sequence_length, c.discount, c.lr, c.target_model_update_interval,
                       c.enable_double_dqn, c.enable_rescale, c.enable_retrace, c.retrace_h)


CASES = {
    "cartpole_duel_avg_retrace": dict(),
    "cartpole_plain_rescale_noretrace": dict(hidden_layers=(16, 8), dueling_type=None, enable_rescale=True, enable_retrace=False,
                                             enable_double_dqn=False, reward_shift=0.1, reward_scale=2.0),
    "pendulum_duel_max_per": dict(env="Pendulum-v1", dueling_type="max", memory="Proportional", burnin=0, sequence_length=4,
                                  env_kwargs=dict(action_division_num=5)),
    "grid_naive_per_nodup": dict(env="Grid", dueling_type="", memory="Proportional", per_has_duplicate=False, burnin=3, sequence_length=2,
                                 lstm_units=8, hidden_layers=(8,)),
    "cartpole_batch40_two_row_tiles": dict(batch_size=40, warmup_size=40, n_envs=9, lstm_units=40, hidden_layers=(70,), capacity=9 * 64),
    # batch sizes that are multiples of 32 stage the recurrent input with cp.async
    "cartpole_batch32_cp_async": dict(batch_size=32, warmup_size=32, n_envs=8, lstm_units=24, capacity=8 * 64, memory="Proportional"),
    "pendulum_batch64_cp_async": dict(env="Pendulum-v1", batch_size=64, warmup_size=64, n_envs=12, lstm_units=40, hidden_layers=(20,),
                                      dueling_type=None, capacity=12 * 64, enable_rescale=True),
    # more than 64 env copies: the rollout's LSTM step runs as tiled GEMM + cell update
    "cartpole_80_envs_gemm_rollout": dict(n_envs=80, lstm_units=24, capacity=80 * 24, warmup_size=64, batch_size=16, memory="Proportional"),
    # one launch per time step instead of the persistent unroll kernels (the path shapes outside their limits take)
    "cartpole_step_launches": dict(_persistent=False),
    "pendulum_per_step_launches_batch40": dict(env="Pendulum-v1", memory="Proportional", batch_size=40, warmup_size=40, n_envs=9, lstm_units=40,
                                               capacity=9 * 64, _persistent=False),
}


@pytest.mark.parametrize("name", list(CASES))
def test_lockstep_rollout_replay_and_updates(name):
    """Vector steps and trainer updates interleaved, every function checked where the reference computes it:
    policy (epsilon-greedy probabilities + choice on the device's Q, exact), env transitions (exact), LSTM state / Q against the
    restated network (1e-4), the sampled items against the restated worker's lists (exact: states, actions, probabilities, rewards,
    dones, hidden state, both kinds of padding), PER leaf selection and IS weights against the oracle memory on the device's tree
    (exact / 1e-6), Q / targets / loss / mean TD (1e-4), gradients (1e-3), parameters after keras Adam (1e-4), priorities in the
    tree (1e-12), target sync and counters."""
    kw = dict(CASES[name])
    persistent = kw.pop("_persistent", True)
    cfg = _cfg(**kw)
    eng = _engine(cfg, persistent=persistent)
    twin = _Twin(eng)
    tr = _trainer(eng, eng.get_weights())
    n_upd = 0
    for g in range(40):
        twin.step()
        if eng.read_state().mem_size < cfg.warmup_size:
            eng.learn(1)  # below the warm-up the call must be a no-op (Trainer.train: `if batches is None: return`)
            assert eng.read_state().train_count == 0
            continue
        if g % 3 == 0:
            continue  # several vector steps between some of the updates
        eng.learn(1)
        items = _check_batch(eng, twin)
        dev_w = eng.t["weights"].cpu().numpy()
        out = _oracle_update(eng, tr, items, dev_w)
        q = eng.t["q"].cpu().numpy()
        np.testing.assert_allclose(q[0], out["q"], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(q[1], out["q_target"], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(eng.t["b_target"].cpu().numpy(), out["target"], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(eng.t["b_tdmean"].cpu().numpy(), np.asarray(out["td_mean"], np.float64), rtol=1e-4, atol=2e-5)
        st2 = eng.read_state()
        assert abs(st2.last_loss - out["loss"]) <= 1e-4 * max(1.0, abs(out["loss"]))
        gd = eng.spec.to_keras(eng.t["grads"].cpu().numpy())
        for a, b in zip(gd, out["grads"]):
            np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-6 + 1e-4 * float(np.abs(b).max()))
        tr.apply(out["grads"])
        n_upd += 1
        for a, b in zip(eng.get_weights(), tr.weights()):
            np.testing.assert_allclose(a, b, rtol=1e-4, atol=2e-5)
        tw = eng.spec.to_keras(eng.t["target"].cpu().numpy())
        for a, b in zip(tw, [w.detach().numpy() for w in tr.target.w]):
            np.testing.assert_allclose(a, b, rtol=1e-4, atol=2e-5)
        assert st2.train_count == n_upd and st2.sync_count == tr.sync_count and st2.adam_step == n_upd
        if eng.per:
            sel = eng.t["sel"].cpu().numpy()
            tdm = eng.t["b_tdmean"].cpu().numpy()
            tree = eng.t["tree"].cpu().numpy()
            last = {int(s): float((abs(tdm[i]) + cfg.per_epsilon) ** cfg.per_alpha) for i, s in enumerate(sel)}
            for s, p in last.items():
                assert abs(tree[s] - p) <= 1e-12 * max(1.0, p)
            cap = eng.R * eng.E
            np.testing.assert_allclose(tree[0], tree[cap - 1:].sum(), rtol=1e-9)
        # resynchronise the oracle's parameters with the device so that 1e-4 bounds one update, not the trajectory
        tr.online.w = [torch.tensor(x, requires_grad=True) for x in eng.get_weights()]
        tr.target.w = [torch.tensor(x) for x in tw]
        tr.m = [torch.tensor(x) for x in eng.spec.to_keras(eng.t["adam_m"].cpu().numpy())]
        tr.v = [torch.tensor(x) for x in eng.spec.to_keras(eng.t["adam_v"].cpu().numpy())]
    assert n_upd >= 8
    assert eng.read_state().episode_count > 0 or cfg.env == "Pendulum-v1"


@pytest.mark.parametrize("has_dup", [True, False])
def test_per_selection_and_weights_equal_the_oracle_memory(has_dup):
    """ProportionalMemory.sample on the device's own tree: identical leaves for identical uniforms (the Philox words replayed), IS
    weights to 1e-6, beta from the PREVIOUS update's train_count (priority_replay_buffer.py:228-250); new rows enter at max_priority,
    anchors whose window the ring overwrote sit at 0 and are never drawn."""
    cfg = _cfg(memory="Proportional", per_has_duplicate=has_dup, n_envs=5, capacity=5 * 12, per_beta_steps=20, epsilon=0.5)
    eng = _engine(cfg)
    E, R, W = eng.E, eng.R, eng.W
    cap = R * E
    for g in range(60):
        eng.vec_step(True)
        if g < 4:
            continue
        st = eng.read_state()
        tree = eng.t["tree"].cpu().numpy().copy()
        cur = eng.t["cursor"].cpu().numpy()
        # leaves: written rows whose window is intact carry a priority, cut anchors and unwritten rows are 0
        for e in range(E):
            for row in range(R):
                c = int(cur[e])
                pos = row if c <= R else (c - 1) - ((c - 1 - row) % R)
                valid = pos < c and (c <= R or pos - (W - 1) >= c - R)
                assert (tree[cap - 1 + e * R + row] > 0) == valid, (g, e, row)
        mem = osum.ProportionalMemory(cap, cfg.per_alpha, cfg.per_beta_initial, cfg.per_beta_steps, has_dup, cfg.per_epsilon)
        mem.tree.tree[:] = tree
        mem.size = int(st.mem_size)
        tc = int(st.train_count)
        idx, w, pri, _ = mem.sample(cfg.batch_size, max(tc - 1, 0), osum.philox_uniforms(cfg.seed, tc))
        eng.learn(1)
        np.testing.assert_array_equal(eng.t["sel"].cpu().numpy(), idx)
        np.testing.assert_allclose(eng.t["weights"].cpu().numpy(), w, rtol=1e-6)
        if not has_dup:
            assert len(set(idx.tolist())) == cfg.batch_size
    assert eng.read_state().max_priority >= 1.0


def test_ring_wrap_keeps_windows_intact():
    """A ring of the minimum size (2 x (burnin + seq_len) rows) wrapping many times: every sampled item still equals the restated
    worker's item, uniform replay."""
    cfg = _cfg(n_envs=4, capacity=4 * 6, burnin=2, sequence_length=3, batch_size=4, warmup_size=4)
    eng = _engine(cfg)
    assert eng.R == 10
    twin = _Twin(eng)
    for g in range(70):
        twin.step(check_net=False)
        if g >= 3:
            eng.learn(1)
            _check_batch(eng, twin)
    assert int(eng.t["cursor"].max().item()) > 3 * eng.R


def test_many_updates_in_one_call_equal_one_by_one():
    cfg = _cfg(memory="Proportional")
    a, b = _engine(cfg), _engine(cfg)
    for _ in range(12):
        a.vec_step(True)
        b.vec_step(True)
    a.learn(5)
    for _ in range(5):
        b.learn(1)
    for k in ("params", "target", "adam_m", "adam_v", "tree", "sel", "b_target"):
        assert torch.equal(a.t[k], b.t[k]), k
    assert a.read_state().train_count == 5


def test_update_under_cuda_graph_replay():
    """The whole update is launch-ordered on one stream with the warm-up gate on device: capturable, and a replay equals eager calls."""
    cfg = _cfg()
    a, b = _engine(cfg), _engine(cfg)
    for _ in range(10):
        a.vec_step(True)
        b.vec_step(True)
    a.learn(1)  # sets the kernels' attributes outside the capture
    b.learn(1)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        a.learn(2)
    graph.replay()  # the capture itself does not execute
    graph.replay()
    b.learn(4)
    torch.cuda.synchronize()
    for k in ("params", "adam_m", "adam_v", "target"):
        assert torch.equal(a.t[k], b.t[k]), k
    assert a.read_state().train_count == 5


def test_evaluate_and_runner_counters():
    from simple_distributed_rl_b200.r2d2 import R2D2Runner

    cfg = _cfg(n_envs=16, capacity=16 * 40)
    r = R2D2Runner(cfg)
    st = r.train(max_train_count=30, updates_per_vec_step=4)
    assert st.train_count == 30 and st.end_reason == "max_train_count over."
    assert st.total_step == st.vec_steps * 16 and st.sync == 15
    rewards = r.evaluate(max_episodes=5)
    assert len(rewards) == 5 and all(1 <= x <= 500 for x in rewards)


def test_learning_pendulum_reference_acceptance_gate():
    """tests/algorithms_/base_r2d2.py:29-44: lstm 32, MLP (16, 16), uniform replay, rescale, burn-in 5, sequence 5, no Retrace; the
    reference trains 200 x 35 updates on one env and asks for the env's baseline (Pendulum-v1: mean reward of 10 evaluation episodes
    >= -500, srl/envs gym registration).  Here 32 env copies, one update per env step as the reference's loop does."""
    from simple_distributed_rl_b200.r2d2 import R2D2Config, R2D2Runner

    cfg = R2D2Config(env="Pendulum-v1", n_envs=32, lstm_units=32, hidden_layers=(16, 16), dueling_type=None, memory="ReplayBuffer",
                     target_model_update_interval=100, enable_rescale=True, burnin=5, sequence_length=5, enable_retrace=False, seed=3)
    r = R2D2Runner(cfg)
    r.train(max_train_count=200 * 35 * 2, train_interval=1)
    rewards = r.evaluate(max_episodes=10)
    assert np.mean(rewards) >= -500, rewards


@pytest.mark.parametrize("M,N,K", [(7, 5, 3), (40, 33, 70), (200, 150, 37), (1300, 1100, 129), (64, 2048, 517), (2048, 517, 640), (2048, 1700, 70), (5184, 3, 1025), (300, 8, 70)])
@pytest.mark.parametrize("a_t,b_t", [(False, False), (True, False), (False, True), (True, True)])
def test_strided_sgemm_every_tile_variant(M, N, K, a_t, b_t):
    """srlx_sgemm (the map every R2D2 layer runs on) for row- and column-major operands, odd leading dimensions, ReLU and accumulate,
    against torch fp64."""
    from simple_distributed_rl_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    lda, ldb = (M + 3 if a_t else K + 1), (K + 5 if b_t else N + 2)
    a = torch.randn((K, lda) if a_t else (M, lda), device="cuda", generator=g)
    b = torch.randn((N, ldb) if b_t else (K, ldb), device="cuda", generator=g)
    A = a[:, :M].T if a_t else a[:, :K]
    Bm = b[:, :K].T if b_t else b[:, :N]
    c = torch.randn(M, N + 1, device="cuda", generator=g)
    want = torch.relu(c[:, :N].double() + A.double() @ Bm.double())
    sa = (1, lda) if a_t else (lda, 1)
    sb = (1, ldb) if b_t else (ldb, 1)
    _lib.check(lib.srlx_sgemm(a.data_ptr(), sa[0], sa[1], b.data_ptr(), sb[0], sb[1], c.data_ptr(), N + 1, M, N, K, 1, 1,
                              torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    np.testing.assert_allclose(c[:, :N].cpu().numpy(), want.float().cpu().numpy(), rtol=2e-5, atol=2e-5 * K ** 0.5)


def test_runner_max_steps_on_grid_and_weight_round_trip():
    """R2D2Runner.train(max_steps=...) (counters polled every 8th vector step), Grid (2 observation ints, 4 actions, LSTM units not a
    multiple of 8), Parameter.call_backup / call_restore in keras order through the device."""
    from simple_distributed_rl_b200.r2d2 import R2D2Engine, R2D2Runner

    cfg = _cfg(env="Grid", n_envs=24, lstm_units=20, hidden_layers=(16,), dueling_type="max", capacity=24 * 60, warmup_size=48,
               batch_size=16, memory="Proportional", burnin=1, sequence_length=4)
    r = R2D2Runner(cfg)
    st = r.train(max_steps=24 * 40, updates_per_vec_step=2)
    assert st.end_reason == "max_steps over." and st.total_step == 24 * 40 and st.vec_steps == 40
    s = r.engine.read_state()
    assert st.train_count == s.train_count and 60 <= s.train_count <= 80 and s.episode_count > 0
    w = r.engine.get_weights()
    assert [x.shape for x in w] == [(2, 80), (20, 80), (80,), (20, 16), (16,), (16, 1), (1,), (20, 16), (16,), (16, 4), (4,)]
    other = R2D2Engine(cfg, weights=w)
    assert torch.equal(other.t["params"], r.engine.t["params"]) and torch.equal(other.t["target"], other.t["params"])
    rewards = r.evaluate(max_episodes=4)
    assert len(rewards) == 4 and all(-3.0 <= x <= 1.0 for x in rewards)


def test_full_size_update_equals_the_oracle_and_the_per_step_path():
    """The shape tools/r2d2_bench.py times (BASELINE configs[3]: LSTM 512, dueling 512, burn-in 40, sequence 80, batch 64, proportional
    sequence replay; 256 env copies here): one update of the persistent unroll kernels + tensor-core GEMM tiles against (a) the oracle
    trainer on the batch the device gathered (Q, targets, loss 1e-4; gradients 2e-3 of their scale; parameters after Adam) and (b) the
    same update through one launch per time step (the path the small lockstep cases pin item by item)."""
    cfg = _cfg(env="CartPole-v1", n_envs=256, lstm_units=512, hidden_layers=(512,), dueling_type="average", burnin=40, sequence_length=80,
               batch_size=64, capacity=256 * 300, warmup_size=256 * 4, memory="Proportional", enable_rescale=True, enable_retrace=False,
               lr=1e-4, epsilon=0.4, seed=2)
    a, b = _engine(cfg), _engine(cfg, persistent=False)
    for _ in range(130):
        a.vec_step(True)
        b.vec_step(True)
    assert torch.equal(a.t["tree"], b.t["tree"]) and torch.equal(a.t["ring_h"], b.t["ring_h"])
    w0 = a.get_weights()
    a.learn(1)
    b.learn(1)
    torch.cuda.synchronize()
    assert torch.equal(a.t["sel"], b.t["sel"]) and torch.equal(a.t["b_actions"], b.t["b_actions"])
    for k, tol in (("q", 2e-4), ("b_target", 2e-4), ("grads", 2e-3), ("params", 1e-5)):
        x, y = a.t[k].double(), b.t[k].double()
        assert float((x - y).abs().max()) <= tol * max(1e-3, float(y.abs().max())), k
    # (a) the oracle on the gathered batch
    D, u, W, S, B = a.D, a.u, a.W, a.S, a.B
    xh = a.t["xh"].cpu().numpy()
    states = np.transpose(xh[0, :W + 1, :, :D], (1, 0, 2))                      # [B, W + 1, D]
    tr = _trainer(a, w0)
    out = tr.train_on_batches(states, a.t["b_actions"].cpu().numpy().tolist(), a.t["b_mu"].cpu().numpy().tolist(),
                              a.t["b_rewards"].cpu().numpy().tolist(), a.t["b_dones"].cpu().numpy().astype(bool).tolist(),
                              xh[0, 0, :, D:D + u], a.t["cbuf"].cpu().numpy()[0, 0], a.t["weights"].cpu().numpy())
    q = a.t["q"].cpu().numpy()
    np.testing.assert_allclose(q[0], out["q"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(q[1], out["q_target"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(a.t["b_target"].cpu().numpy(), out["target"], rtol=1e-4, atol=1e-4)
    assert abs(a.read_state().last_loss - out["loss"]) <= 1e-4 * max(1.0, abs(out["loss"]))
    for gd, go in zip(a.spec.to_keras(a.t["grads"].cpu().numpy()), out["grads"]):
        assert float(np.abs(gd - go).max()) <= 2e-3 * max(1e-6, float(np.abs(go).max()))
    tr.apply(out["grads"])
    for wd, wo in zip(a.get_weights(), tr.weights()):
        np.testing.assert_allclose(wd, wo, rtol=1e-4, atol=2e-5)
