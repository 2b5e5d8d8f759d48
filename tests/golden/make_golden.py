"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE (pocokhc/simple_distributed_rl v1.4.5).

Run in the build container only (the reference tree does not exist on the GPU box):

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

Outputs (committed):
    grid_transitions.npz   Grid.step / _move / reward_done_func of srl/envs/grid.py driven with known np.random uniforms,
                           plus the EnvRun truncation length (srl/base/env/env_run.py:360-362)
    sumtree.npz            srl/rl/memories/priority_memories/proportional_memory.py ProportionalMemory driven with injected
                           uniforms: tree arrays, sampled indices, IS weights, max_priority after add/sample/update sequences
    trainer_<case>.npz     srl/algorithms/{dqn,rainbow}/model_torch.py Trainer.train() on frozen batches with injected
                           NoisyLinear noise: target_q, q, loss, priorities, parameters before/after two updates
    functions.npz          srl/rl/functions.py rescaling / inverse_rescaling / create_epsilon_list
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import srl  # noqa: E402  (the reference)
from srl.algorithms import dqn, rainbow  # noqa: E402
from srl.envs import grid as ref_grid  # noqa: E402
from srl.rl import functions as ref_funcs  # noqa: E402
from srl.rl.memories.priority_memories import proportional_memory as ref_pm  # noqa: E402

from oracle import nets  # noqa: E402


# --------------------------------------------------------------------------------------------------------
def gen_grid():
    env = ref_grid.Grid()
    recs = []
    cells = [(x, y) for y in range(env.H) for x in range(env.W) if env.field[y][x] != 9]
    for x, y in cells:
        for a in range(4):
            for seed in range(48):
                np.random.seed(seed * 7919 + a)
                u = np.random.random_sample()
                np.random.seed(seed * 7919 + a)
                env.player_pos = (x, y)
                st, r, done, trunc = env.step(a)
                assert trunc is False
                recs.append((x, y, a, u, st[0], st[1], r, int(done), env.action.value))
    recs = np.array(recs, dtype=np.float64)
    # start position
    st = env.reset()
    # EnvRun truncation: bump into the left wall from the start cell forever (move_prob=1 so no slip)
    e2 = srl.make_env(srl.EnvConfig("Grid", kwargs=dict(move_prob=1.0)))
    e2.setup()
    e2.reset()
    n = 0
    while not e2.done:
        e2.step(0)
        n += 1
    np.savez(os.path.join(HERE, "grid_transitions.npz"), recs=recs, start=np.array(st), trunc_steps=n,
             trunc_done_type=str(e2.done_type.name) if hasattr(e2, "done_type") else "")
    print("grid:", recs.shape, "start", st, "episode length when never terminating:", n)


# --------------------------------------------------------------------------------------------------------
class _Uniforms:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.log = []

    def __call__(self):
        u = float(self.rng.random())
        self.log.append(u)
        return u


def gen_sumtree():
    out = {}
    for case, (cap, alpha, beta0, bsteps, dup) in enumerate(
        [(10, 0.8, 1.0, 10, True), (10, 0.6, 0.4, 1000, False), (37, 0.6, 0.4, 1_000_000, True), (64, 0.5, 0.4, 100, True)]
    ):
        mem = ref_pm.ProportionalMemory(cap, alpha, beta0, bsteps, has_duplicate=dup)
        uni = _Uniforms(100 + case)
        ref_pm.random.random = uni
        rng = np.random.default_rng(case)
        n_add = cap + cap // 2
        add_pri = rng.random(n_add) * 3.0
        use_none = rng.random(n_add) < 0.3
        for i in range(n_add):
            mem.add(i, None if use_none[i] else float(add_pri[i]))
        out[f"c{case}_cfg"] = np.array([cap, alpha, beta0, bsteps, int(dup)], dtype=np.float64)
        out[f"c{case}_add_pri"] = add_pri
        out[f"c{case}_add_none"] = use_none
        out[f"c{case}_tree_after_add"] = np.array(mem.tree.tree, dtype=np.float64)
        out[f"c{case}_maxp_after_add"] = mem.max_priority
        B = 5
        n_iter = 6
        idxs, ws, ups, trees, maxps, ulogs, steps = [], [], [], [], [], [], []
        for it in range(n_iter):
            step = it * 3
            uni.log = []
            batches, weights, indices = mem.sample(B, step)
            ulogs.append(list(uni.log) + [np.nan] * (64 - len(uni.log)))
            idxs.append(indices)
            ws.append(weights)
            up = rng.standard_normal(B) * 2.0
            mem.update(indices, up)
            ups.append(up)
            trees.append(np.array(mem.tree.tree, dtype=np.float64))
            maxps.append(mem.max_priority)
            steps.append(step)
        out[f"c{case}_idx"] = np.array(idxs, dtype=np.int64)
        out[f"c{case}_weights"] = np.array(ws, dtype=np.float64)
        out[f"c{case}_upd"] = np.array(ups, dtype=np.float64)
        out[f"c{case}_trees"] = np.array(trees)
        out[f"c{case}_maxp"] = np.array(maxps)
        out[f"c{case}_uniforms"] = np.array(ulogs, dtype=np.float64)
        out[f"c{case}_steps"] = np.array(steps)
        out[f"c{case}_size"] = mem.size
        out[f"c{case}_write"] = mem.tree.write
    ref_pm.random.random = random.random
    out["n_cases"] = 4
    np.savez(os.path.join(HERE, "sumtree.npz"), **out)
    print("sumtree: cases", out["n_cases"])


# --------------------------------------------------------------------------------------------------------
def gen_functions():
    x = np.concatenate([np.linspace(-50, 50, 41), [-1e-3, 0.0, 1e-3, 123.456]]).astype(np.float32)
    np.savez(
        os.path.join(HERE, "functions.npz"),
        x=x,
        rescaling=ref_funcs.rescaling(x),
        inverse_rescaling=ref_funcs.inverse_rescaling(x),
        eps_list_8=np.array(ref_funcs.create_epsilon_list(8, 0.4, 7.0)),
        eps_list_1=np.array(ref_funcs.create_epsilon_list(1, 0.4, 7.0)),
    )
    print("functions ok")


# --------------------------------------------------------------------------------------------------------
class _StubMemory:
    """Stands in for RLPriorityReplayBuffer: returns the frozen batch, records the priorities written back."""

    def __init__(self, batches_list, weights_list):
        self.batches_list = batches_list
        self.weights_list = weights_list
        self.i = 0
        self.updates = []

    def sample(self):
        b, w = self.batches_list[self.i], self.weights_list[self.i]
        self.i += 1
        return b, w, list(range(len(b)))

    def update(self, update_args, priorities, step):
        self.updates.append((np.array(priorities).copy(), step))

    def length(self):
        return 10**6


def _noise_schedule(spec, model):
    """flat-layout slices in the order NoisyLinear.forward draws them (w then b per module, modules in forward order)."""
    from srl.rl.torch_.modules.noisy_linear import NoisyLinear

    key2off = {}
    for off, shape, kmu, ksig in spec._keys("rainbow"):
        key2off[kmu] = (off, int(np.prod(shape)), shape)
    sched = []
    for name, mod in model.named_modules():
        if isinstance(mod, NoisyLinear):
            sched.append(key2off[name + ".w_mu"])
            sched.append(key2off[name + ".b_mu"])
    return sched


def gen_trainer(case, algo, hidden, dueling, noisy, multisteps, double, rescale, B=8, n_updates=2, retrace_h=1.0,
                seed=0, invalid_p=0.0):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    if algo == "dqn":
        cfg = dqn.Config(batch_size=B, enable_double_dqn=double, enable_rescale=rescale, target_model_update_interval=1000)
        cfg.hidden_block.set(hidden)
    else:
        cfg = rainbow.Config(batch_size=B, enable_double_dqn=double, enable_rescale=rescale, multisteps=multisteps,
                             enable_noisy_dense=noisy, retrace_h=retrace_h, target_model_update_interval=1000)
        if dueling is None:
            cfg.hidden_block.set(hidden)
        else:
            cfg.hidden_block.set_dueling_network(hidden, dueling_type=dueling)
    cfg.memory.compress = False
    cfg.memory.warmup_size = B
    runner = srl.Runner("Grid", cfg)
    runner.set_device("CPU")
    parameter = runner.make_parameter()
    D, A = 2, 4
    spec = nets.NetSpec(D, tuple(hidden), A, dueling, noisy)
    # de-synchronise target from online so the two nets differ
    with torch.no_grad():
        for p in parameter.q_target.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    mu0, sig0 = spec.from_state_dict(parameter.q_online.state_dict(), algo)
    tmu0, tsig0 = spec.from_state_dict(parameter.q_target.state_dict(), algo)
    assert spec.n_params == sum(p.numel() for n, p in parameter.q_online.named_parameters() if "sigma" not in n), (
        spec.n_params, [(n, p.shape) for n, p in parameter.q_online.named_parameters()])

    M = multisteps
    states = rng.uniform(0, 5, size=(n_updates, B, M + 1, D)).astype(np.float32)
    actions = rng.integers(0, A, size=(n_updates, B, M))
    rewards = rng.normal(0, 1, size=(n_updates, B, M)).astype(np.float32)
    terms = (rng.random((n_updates, B, M)) < 0.25).astype(np.int64)
    weights = rng.uniform(0.3, 1.0, size=(n_updates, B)).astype(np.float32)
    # invalid actions of the M next states (never all of them): dqn.py:156-165, rainbow_nomultisteps.py:19-31, rainbow.py:236-249
    invalid = rng.random((n_updates, B, M, A)) < invalid_p
    invalid[..., 0] &= ~invalid.all(axis=-1)
    inv_list = lambda u, i, k: [int(a) for a in np.nonzero(invalid[u, i, k])[0]]  # noqa: E731
    # make a few windows greedy-consistent is not needed: random actions already hit both retrace branches for A=4
    batches_list = []
    for u in range(n_updates):
        bl = []
        for i in range(B):
            if algo == "rainbow" and M > 1:
                steps = [[states[u, i, 0], None, None, None, None]]
                for k in range(M):
                    steps.append([states[u, i, k + 1], np.eye(A, dtype=np.float32)[actions[u, i, k]].tolist(),
                                  float(rewards[u, i, k]), int(terms[u, i, k]), inv_list(u, i, k)])
                bl.append(steps)
            else:
                bl.append([states[u, i, 0], states[u, i, 1], np.eye(A, dtype=np.float32)[actions[u, i, 0]].tolist(),
                           float(rewards[u, i, 0]), int(1 - terms[u, i, 0]), inv_list(u, i, 0)])
        batches_list.append(bl)
    memory = _StubMemory(batches_list, [w for w in weights])
    trainer = cfg.make_trainer(parameter, memory)
    trainer.on_setup()

    # noise injection: pass order inside Trainer.train is online(s') [1], target(s') [2], online(s) [0]
    noise = rng.standard_normal((n_updates, 3, spec.n_params)).astype(np.float32)
    real_randn = torch.randn
    if noisy:
        import srl.rl.torch_.modules.noisy_linear as nl

        sched = _noise_schedule(spec, parameter.q_online)
        state = {"u": 0, "call": 0}
        if algo == "rainbow" and M > 1 and not double:
            order = [1, 2, 0]
        else:
            order = [2, 1, 0] if (algo == "rainbow" and M == 1) or algo == "dqn" else [1, 2, 0]
        # rainbow n-step: pred_q(online) first then pred_target_q (rainbow.py:219-220);
        # 1-step variants: pred_target_q first then pred_q (rainbow_nomultisteps.py:22-26, dqn.py:155-159)

        def fake_randn(size, **kw):
            per_pass = len(sched)
            c = state["call"]
            p = order[c // per_pass]
            off, n, shape = sched[c % per_pass]
            assert tuple(size) == tuple(shape), (size, shape)
            state["call"] += 1
            return torch.tensor(noise[state["u"], p, off : off + n].reshape(shape))

        nl.torch.randn = fake_randn

    # capture target_q
    captured = []
    if algo == "rainbow" and M == 1:
        import srl.algorithms.rainbow.model_torch as mt

        orig = mt.calc_target_q

        def wrap(*a, **k):
            r = orig(*a, **k)
            captured.append(np.array(r[0]).copy())
            return r

        mt.calc_target_q = wrap
    else:
        orig = parameter.calc_target_q

        def wrap(*a, **k):
            r = orig(*a, **k)
            captured.append(np.array(r[0] if isinstance(r, tuple) else r).copy())
            return r

        parameter.calc_target_q = wrap

    losses, mus, sigs, tmus = [], [], [], []
    for u in range(n_updates):
        if noisy:
            state["u"], state["call"] = u, 0
        trainer.train()
        losses.append(trainer.info["loss"])
        m_, s_ = spec.from_state_dict(parameter.q_online.state_dict(), algo)
        mus.append(m_)
        sigs.append(s_ if s_ is not None else np.zeros(0, np.float32))
        tmus.append(spec.from_state_dict(parameter.q_target.state_dict(), algo)[0])
    if noisy:
        nl.torch.randn = real_randn
    if algo == "rainbow" and M == 1:
        mt.calc_target_q = orig

    np.savez(
        os.path.join(HERE, f"trainer_{case}.npz"),
        algo=algo, hidden=np.array(hidden), dueling="none" if dueling is None else dueling, noisy=int(noisy),
        multisteps=M, double=int(double), rescale=int(rescale), retrace_h=retrace_h, discount=cfg.discount, lr=cfg.lr,
        mu0=mu0, sigma0=sig0 if sig0 is not None else np.zeros(0, np.float32),
        tmu0=tmu0, tsigma0=tsig0 if tsig0 is not None else np.zeros(0, np.float32),
        states=states, actions=actions, rewards=rewards, terms=terms, weights=weights, noise=noise if noisy else np.zeros(0),
        invalid=invalid,
        target_q=np.array(captured), losses=np.array(losses), priorities=np.array([p for p, s in memory.updates]),
        update_steps=np.array([s for p, s in memory.updates]), mu_after=np.array(mus), sigma_after=np.array(sigs),
        tmu_after=np.array(tmus), train_count=trainer.train_count, sync_count=trainer.sync_count,
    )
    print(f"trainer_{case}: n_params={spec.n_params} losses={losses}")


# ----------------------------------------------------------------------------------------------------------------------
# Worker-side records (R6): what dqn.Worker.on_step / rainbow.Worker._add_batch hand to memory.add() along a real
# trajectory of the reference Runner on Grid -- the n-step windows incl. the padded tail after an episode ends
# (srl/algorithms/rainbow/rainbow.py:341-400, dqn/dqn.py:229-246).  The trajectory itself is logged from EnvRun.
class _RecordingMemory:
    """IPriorityMemory (srl/rl/memories/priority_memories/imemory.py:7-34) that records every add()."""
    LOG = []

    def __init__(self, capacity, **kw):
        self.n = 0

    def clear(self):
        self.n = 0

    def length(self):
        return self.n

    def add(self, batch, priority=None):
        _RecordingMemory.LOG.append(batch)
        self.n += 1

    def sample(self, batch_size, step):
        raise RuntimeError("rollout only")

    def update(self, indices, priorities):
        pass

    def backup(self):
        return []

    def restore(self, data):
        pass


def gen_worker_records():
    import srl
    from srl.algorithms import dqn, rainbow
    from srl.base.define import DoneTypes
    from srl.base.run.callback import RunCallback

    class Traj(RunCallback):
        def __init__(self):
            self.rows = []  # (episode, s, a, r, s', terminated, done)
            self.ep = -1
            self.prev = None

        def on_episode_begin(self, context, state, **kw):
            self.ep += 1
            self.prev = np.array(state.env.state, dtype=np.float32)

        def on_step_end(self, context, state, **kw):
            env = state.env
            nxt = np.array(env.state, dtype=np.float32)
            self.rows.append((self.ep, self.prev.copy(), int(state.action), float(env.reward),
                              nxt.copy(), int(env.done_type == DoneTypes.TERMINATED), int(env.done)))
            self.prev = nxt

    sys.modules["_srlx_recmem"] = sys.modules[__name__]
    out = {}
    for name, cfg in (("rainbow_m3", rainbow.Config(multisteps=3, enable_noisy_dense=False, epsilon=0.5)),
                      ("rainbow_m2_clip", rainbow.Config(multisteps=2, enable_noisy_dense=False, epsilon=0.7, enable_reward_clip=True)),
                      ("dqn", dqn.Config(epsilon=0.5))):
        if name.startswith("rainbow"):
            cfg.hidden_block.set_dueling_network((16,))
        else:
            cfg.hidden_block.set((16,))
        cfg.memory.set_custom(f"{__name__}:_RecordingMemory", {})
        cfg.memory.warmup_size = 1000
        cfg.memory.compress = False
        _RecordingMemory.LOG = []
        runner = srl.Runner("Grid", cfg)
        runner.set_seed(5)
        tr = Traj()
        runner.rollout(max_steps=400, callbacks=[tr], enable_progress=False)
        rows = tr.rows
        out[f"{name}_ep"] = np.array([r[0] for r in rows], dtype=np.int64)
        out[f"{name}_s"] = np.stack([r[1] for r in rows])
        out[f"{name}_a"] = np.array([r[2] for r in rows], dtype=np.int64)
        out[f"{name}_r"] = np.array([r[3] for r in rows], dtype=np.float64)
        out[f"{name}_ns"] = np.stack([r[4] for r in rows])
        out[f"{name}_term"] = np.array([r[5] for r in rows], dtype=np.int64)
        out[f"{name}_done"] = np.array([r[6] for r in rows], dtype=np.int64)
        log = _RecordingMemory.LOG
        if name == "dqn":
            # dqn record: [state, next_state, onehot action, reward, undone, next_invalid_actions] (dqn.py:229-246)
            out["dqn_b_s"] = np.stack([np.asarray(b[0], dtype=np.float32) for b in log])
            out["dqn_b_ns"] = np.stack([np.asarray(b[1], dtype=np.float32) for b in log])
            out["dqn_b_a"] = np.array([int(np.argmax(b[2])) for b in log], dtype=np.int64)
            out["dqn_b_r"] = np.array([float(b[3]) for b in log], dtype=np.float64)
            out["dqn_b_undone"] = np.array([int(b[4]) for b in log], dtype=np.int64)
        else:
            M = cfg.multisteps
            out[f"{name}_b_states"] = np.stack([np.stack([np.asarray(e[0], dtype=np.float32) for e in b]) for b in log])  # [n][M+1][2]
            out[f"{name}_b_a"] = np.array([[int(np.argmax(e[1])) for e in b[1:]] for b in log], dtype=np.int64)           # [n][M]
            out[f"{name}_b_r"] = np.array([[float(e[2]) for e in b[1:]] for b in log], dtype=np.float64)
            out[f"{name}_b_term"] = np.array([[int(e[3]) for e in b[1:]] for b in log], dtype=np.int64)
            assert all(len(b) == M + 1 for b in log)
    np.savez_compressed(os.path.join(HERE, "worker_records.npz"), **out)
    print("worker_records:", {k: v.shape for k, v in out.items() if k.endswith("_b_a") or k.endswith("_a")})


def gen_rankbased():
    """RankBasedMemory (srl/rl/memories/priority_memories/rankbased_memory.py) with np.random seeded: add / sample / update
    sequences; the uniform stream np.random.choice consumes is RandomState(seed).random_sample, stored next to the outputs."""
    from srl.rl.memories.priority_memories.rankbased_memory import RankBasedMemory

    out = {}
    cases = [(64, 0.6, 0.4, 100, 8), (1000, 1.0, 0.5, 50, 32), (300, 0.0, 0.4, 10, 16), (5000, 0.8, 0.4, 1000, 64)]
    out["n_cases"] = len(cases)
    for c, (cap, alpha, beta0, bsteps, B) in enumerate(cases):
        rng = np.random.default_rng(100 + c)
        m = RankBasedMemory(cap, alpha, beta0, bsteps)
        n_add = cap if c != 2 else cap // 2 + 7  # one case samples from a partly filled memory
        pri0 = (rng.random(n_add) ** 2 * 3).astype(np.float32)
        for i in range(n_add):
            m.add(("item", i), float(pri0[i]))
        out[f"c{c}_cfg"] = np.array([cap, alpha, beta0, bsteps, B, n_add], dtype=np.float64)
        out[f"c{c}_pri0"] = pri0
        idxs, ws, us, upds, sorts = [], [], [], [], []
        for step in range(6):
            seed = 1000 * c + step
            np.random.seed(seed)
            us.append(np.random.RandomState(seed).random_sample(4 * B))
            batches, w, idx = m.sample(B, step * 7)
            assert [b[1] for b in batches] == list(idx)
            sorts.append(np.argsort(-m.priorities[:n_add]))
            idxs.append(np.asarray(idx, dtype=np.int64))
            ws.append(np.asarray(w, dtype=np.float64))
            td = np.abs(rng.normal(size=B)).astype(np.float32)
            upds.append(td)
            m.update(idx, td)
        out[f"c{c}_idx"], out[f"c{c}_w"], out[f"c{c}_u"], out[f"c{c}_upd"] = np.array(idxs), np.array(ws), np.array(us), np.array(upds)
        out[f"c{c}_pri_final"] = m.priorities.copy()
        b = m.backup()
        assert b[0] == cap and len(b) == 4 and b[3] == n_add % cap
    np.savez_compressed(os.path.join(HERE, "rankbased.npz"), **out)
    print("rankbased: ok")


def gen_spaces():
    """BoxSpace.create_division_tbl (srl/base/spaces/box.py:317-366): the discrete action set value-based algorithms see on a
    continuous-action env (Pendulum-v1: Box(1,) in [-2, 2], RLConfig.action_division_num = 10 by default)."""
    from srl.base.rl.config import RLConfig
    from srl.base.spaces.box import BoxSpace

    out = {}
    for n in (2, 3, 5, 10, 16):
        sp = BoxSpace((1,), -2.0, 2.0, np.float32)
        sp.create_division_tbl(n)
        out[f"pendulum_div{n}"] = np.asarray(sp.division_tbl, dtype=np.float32).reshape(-1)
    out["default_action_division_num"] = np.array([RLConfig.__dataclass_fields__["action_division_num"].default])
    np.savez_compressed(os.path.join(HERE, "spaces.npz"), **out)
    print("spaces:", {k: v.tolist() for k, v in out.items() if k.endswith("10") or k.startswith("default")})


def gen_ppo_returns():
    """ppo.Worker.on_step (srl/algorithms/ppo/ppo.py:357-406): the GAE / Monte-Carlo return accumulation at episode end, executed
    from the reference's own source.  The module imports TensorFlow (absent here) only for the network classes, so it is imported
    over permissive stub modules; on_step itself is python + numpy.  The value network is replaced by a stub that returns the
    prepared V(s) / V(s') arrays (what `self.parameter.model(...)` would give), so the golden pins the ARITHMETIC of the
    accumulation (order of operations, float32 vs python-float rounding, the last-step rule delta = r - v), which is what the
    device scan (srlx_returns_scan) restates."""
    import types

    class _Any(type):
        def __getattr__(cls, k):
            return _Stub

    class _Stub(metaclass=_Any):
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Stub()

        def __getattr__(self, k):
            return _Stub()

    def fake(name):
        m = types.ModuleType(name)
        m.__getattr__ = lambda k: _Stub
        m.__path__ = []
        return m

    for n in ["tensorflow", "tensorflow.keras", "tensorflow.keras.layers", "tensorflow_probability"]:
        sys.modules.setdefault(n, fake(n))
    sys.modules["tensorflow"].keras = sys.modules["tensorflow.keras"]
    from srl.algorithms.ppo import ppo

    class _Val:
        def __init__(self, a):
            self.a = a

        def numpy(self):
            return self.a

    rng = np.random.default_rng(11)
    out = {"numpy_version": np.array(np.__version__)}
    for method, discount, lam, clip in (("GAE", 0.9, 0.9, None), ("GAE", 0.99, 0.95, (-1.0, 1.5)), ("MC", 0.9, 0.9, None),
                                        ("MC", 0.997, 0.9, (-2.0, 0.25))):
        name = f"{method.lower()}_g{discount}_l{lam}_{'clip' if clip else 'noclip'}"
        lens = [1, 2, 7, 33, 200, 5]
        T = sum(lens)
        reward = rng.normal(size=T) * 2.0                       # python floats reach the worker (worker.reward)
        v = rng.normal(size=T).astype(np.float32)               # V(s_t)
        nv = rng.normal(size=T).astype(np.float32)              # V(s_{t+1})
        done = np.zeros(T, dtype=np.uint8)
        added = []
        w = object.__new__(ppo.Worker)
        ctx = types.SimpleNamespace(training=True, distributed=False, rl_render_mode="")
        w._RLWorkerGeneric__worker_run = types.SimpleNamespace(_context=ctx)
        w.config = types.SimpleNamespace(reward_clip=clip, experience_collection_method=method, state_clip=None, discount=discount,
                                         gae_discount=lam)
        w.memory = types.SimpleNamespace(add=lambda b: added.append(float(b["discounted_reward"])) or
                                         added_dtype.append(b["discounted_reward"].dtype))
        added_dtype = []
        t = 0
        for n in lens:
            w.on_reset(None)
            ep0 = t
            calls = []

            def model(x, _ep0=ep0, _calls=calls):
                # first call: V of the recent states, second call: V of the recent next states (ppo.py:386-387)
                _calls.append(len(x))
                src = v if len(_calls) == 1 else nv
                return _Val(src[_ep0:_ep0 + len(x)].reshape(-1, 1)), None

            w.parameter = types.SimpleNamespace(model=model)
            for i in range(n):
                w.recent_batch.append({"state": np.zeros(3, np.float32)})   # what policy() appends (ppo.py:325-345)
                last = i == n - 1
                wk = types.SimpleNamespace(reward=float(reward[t]), next_state=np.zeros(3, np.float32), done=last)
                w.on_step(wk)
                done[t] = last
                t += 1
        # the reference emits an episode's items LAST step first (reversed loop): put them back in time order
        ret = np.zeros(T, dtype=np.float32)
        pos, k = 0, 0
        for n in lens:
            for i in reversed(range(n)):
                ret[pos + i] = np.float32(added[k])
                k += 1
            pos += n
        assert k == T and all(d == np.float32 for d in added_dtype)
        out[f"{name}_reward"], out[f"{name}_v"], out[f"{name}_nv"], out[f"{name}_done"], out[f"{name}_ret"] = reward, v, nv, done, ret
        out[f"{name}_params"] = np.array([discount, lam, clip[0] if clip else np.nan, clip[1] if clip else np.nan])
    np.savez_compressed(os.path.join(HERE, "ppo_returns.npz"), **out)
    print("ppo_returns:", [k for k in out if k.endswith("_ret")], "numpy", np.__version__)


def gen_r2d2_targets():
    """r2d2.Trainer._train_on_batches (srl/algorithms/r2d2/r2d2.py:109-215): the per-sequence target / TD-error / Retrace loop
    (:150-203), executed from the reference's own source.  TensorFlow (absent here) is replaced by numpy-backed stubs and the
    two LSTM Q-networks by stubs that return prepared Q arrays, so the golden pins the ARITHMETIC of the loop -- double-DQN
    index from the online Q, value from the target Q, optional rescaling, `target_t = gain_t + retrace * td_{t+1}`,
    `retrace *= discount * retrace_h * min(1, pi/mu)`, the mixed float32 / python-float rounding, the mean TD error per
    sequence -- which is what srlx_sequence_targets restates on the device."""
    import types

    captured = {}

    class _Loss:
        def __init__(self, v=0.0):
            self.v = v

        def __add__(self, o):
            return _Loss(self.v)

        __iadd__ = __add__

        def numpy(self):
            return self.v

    class _Frozen:
        def __init__(self, a):
            self.a = a

        def numpy(self):
            return self.a

    class _Tape:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def gradient(self, loss, variables):
            return []

    tf = types.ModuleType("tensorflow")
    tf.__path__ = []
    tf.__version__ = "2.16.1"
    tf.function = lambda *a, **k: (lambda f: f)
    tf.__getattr__ = lambda k: _Stub
    tf.one_hot = lambda idx, n, axis=2: np.eye(n, dtype=np.float32)[np.asarray(idx)]
    tf.stack = lambda xs: np.stack(xs)
    tf.GradientTape = _Tape
    tf.stop_gradient = lambda x: _Frozen(np.asarray(x))
    tf.reduce_sum = lambda x, axis=None: np.sum(x, axis=axis) if axis is not None else float(np.sum(x))

    class _Any(type):
        def __getattr__(cls, k):
            return _Stub

    class _Stub(metaclass=_Any):
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Stub()

        def __getattr__(self, k):
            return _Stub()

    keras = types.ModuleType("tensorflow.keras")
    keras.__getattr__ = lambda k: _Stub
    keras.__path__ = []
    tf.keras = keras
    saved = {n: sys.modules.get(n) for n in ("tensorflow", "tensorflow.keras")}
    sys.modules["tensorflow"], sys.modules["tensorflow.keras"] = tf, keras
    for n in [m for m in sys.modules if m.startswith("srl.algorithms.r2d2") or m.startswith("srl.rl.tf")]:
        del sys.modules[n]
    lay = types.ModuleType("tensorflow.keras.layers")
    lay.__getattr__ = lambda k: _Stub
    sys.modules["tensorflow.keras.layers"] = lay
    from srl.algorithms.r2d2 import r2d2

    rng = np.random.default_rng(5)
    out = {"numpy_version": np.array(np.__version__)}
    B, T, A = 6, 80, 4
    for name, double, rescale, retrace, h, disc in (("double_retrace", True, False, True, 0.95, 0.997),
                                                   ("plain", False, False, False, 1.0, 0.99),
                                                   ("double_rescale_retrace", True, True, True, 0.9, 0.997),
                                                   ("target_rescale", False, True, False, 1.0, 0.9)):
        q_on = rng.normal(size=(B, T + 1, A)).astype(np.float32) * 2
        q_tg = rng.normal(size=(B, T + 1, A)).astype(np.float32) * 2
        q_on[0, 3, :] = q_on[0, 3, 0]          # ties: argmax takes the first, pi = 1 / (number of maxima)
        q_on[1, 10, 2] = q_on[1, 10, 1] = q_on[1, 10].max() + 1
        actions = rng.integers(A, size=(B, T))
        greedy = np.argmax(q_on[:, :T], axis=2)
        take = rng.random((B, T)) < 0.7         # mostly greedy behaviour so Retrace coefficients are not all zero
        actions = np.where(take, greedy, actions)
        mu = rng.uniform(0.05, 1.0, size=(B, T))
        rewards = rng.normal(size=(B, T))
        dones = (rng.random((B, T)) < 0.05)
        dones[2, T - 1] = True
        batches = []
        for b in range(B):
            batches.append({"states": [np.zeros(3, np.float32)] * (T + 1), "actions": [int(a) for a in actions[b]],
                            "probs": [float(x) for x in mu[b]], "rewards": [float(x) for x in rewards[b]],
                            "dones": [bool(x) for x in dones[b]], "invalid_actions": [[] for _ in range(T + 1)],
                            "hidden_states": [np.zeros(2, np.float32), np.zeros(2, np.float32)]})
        tr = object.__new__(r2d2.Trainer)
        tr.config = types.SimpleNamespace(burnin=0, sequence_length=T, action_space=types.SimpleNamespace(n=A), enable_double_dqn=double,
                                          enable_rescale=rescale, discount=disc, enable_retrace=retrace, retrace_h=h)
        online = types.SimpleNamespace(losses=[], trainable_variables=[])
        tr.parameter = types.SimpleNamespace(
            q_online=type("Q", (), {"__call__": lambda self, x, hs, training=False: (q_on, hs), "losses": [], "trainable_variables": []})(),
            q_target=type("Q", (), {"__call__": lambda self, x, hs, training=False: (_Frozen(q_tg), hs)})())
        tr.loss = lambda target, q: captured.update(target=np.asarray(target), q=np.asarray(q)) or _Loss()
        tr.optimizer = types.SimpleNamespace(apply_gradients=lambda pairs: None)
        td_means, _ = tr._train_on_batches(batches, np.ones((B, 1)))
        tgt = captured["target"]
        assert tgt.shape == (B, T)
        out[f"{name}_q_on"], out[f"{name}_q_tg"] = q_on, q_tg
        out[f"{name}_actions"], out[f"{name}_mu"], out[f"{name}_rewards"], out[f"{name}_dones"] = actions, mu, rewards, dones
        out[f"{name}_target"] = tgt
        out[f"{name}_target_dtype"] = np.array(str(tgt.dtype))
        out[f"{name}_td_mean"] = np.asarray(td_means)
        out[f"{name}_td_mean_dtype"] = np.array(str(np.asarray(td_means).dtype))
        out[f"{name}_q_sa"] = captured["q"]
        out[f"{name}_params"] = np.array([float(double), float(rescale), float(retrace), h, disc])
    for n, m in saved.items():
        if m is None:
            sys.modules.pop(n, None)
        else:
            sys.modules[n] = m
    np.savez_compressed(os.path.join(HERE, "r2d2_targets.npz"), **out)
    print("r2d2_targets:", [(k, str(out[k])) for k in out if k.endswith("dtype")])


if __name__ == "__main__":
    only = set(sys.argv[1:])  # e.g. `make_golden.py worker` regenerates only worker_records.npz
    if not only or "spaces" in only:
        gen_spaces()
    if not only or "rankbased" in only:
        gen_rankbased()
    if not only or "ppo" in only:
        gen_ppo_returns()
    if not only or "r2d2" in only:
        gen_r2d2_targets()
    if not only or "worker" in only:
        gen_worker_records()
    if not only or "invalid" in only:  # the same trainer steps with invalid-action lists in the records
        gen_trainer("dqn_mlp32_invalid_double", "dqn", (32,), None, False, 1, True, False, invalid_p=0.35, seed=3)
        gen_trainer("dqn_mlp32_invalid_nodouble_rescale", "dqn", (32,), None, False, 1, False, True, invalid_p=0.35, seed=4)
        gen_trainer("rainbow_duel32_invalid_m1", "rainbow", (32,), "average", False, 1, True, False, invalid_p=0.35, seed=5)
        gen_trainer("rainbow_duel32_invalid_m3", "rainbow", (32,), "average", False, 3, True, False, invalid_p=0.35, seed=6)
        gen_trainer("rainbow_mlp32_invalid_m2_nodouble", "rainbow", (32,), None, False, 2, False, False, invalid_p=0.35, seed=7, retrace_h=0.9)
    if not only or "base" in only:
        gen_grid()
        gen_sumtree()
        gen_functions()
        gen_trainer("dqn_mlp64x64_double", "dqn", (64, 64), None, False, 1, True, False)
        gen_trainer("dqn_mlp32_plain_rescale", "dqn", (32,), None, False, 1, False, True)
        gen_trainer("rainbow_default_noisy_m3", "rainbow", (512,), "average", True, 3, True, False)
        gen_trainer("rainbow_duel64x64_m3", "rainbow", (64, 64), "average", False, 3, True, False)
        gen_trainer("rainbow_duelmax_m2_nodouble", "rainbow", (32,), "max", False, 2, False, False, retrace_h=0.9)
        gen_trainer("rainbow_naive_noisy_m1", "rainbow", (32,), "", True, 1, True, False)
        gen_trainer("rainbow_mlp_noisy_m3_rescale", "rainbow", (32, 16), None, True, 3, True, True)
