"""Generate tests/golden/learner_<case>.npz by EXECUTING THE REFERENCE (pocokhc/simple_distributed_rl v1.4.5) on a frozen replay
memory -- the whole `Trainer.train()` step INCLUDING the memory's own sample / update:

    PriorityReplayBuffer.sample   srl/rl/memories/priority_replay_buffer.py:228-245
    ProportionalMemory.sample     srl/rl/memories/priority_memories/proportional_memory.py:131-169  (leaf selection, IS weights)
    Trainer.train                 srl/algorithms/dqn/model_torch.py:90-132, srl/algorithms/rainbow/model_torch.py:85-122
    ProportionalMemory.update     proportional_memory.py:171-177

Run in the build container only (the reference tree does not exist on the GPU box):

    PYTHONPATH=/root/reference python tests/golden/make_learner_golden.py

How the reference is made comparable with the device (nothing in the reference is modified; only its random sources are fed):
  * the memory is restored (`memory.call_restore`, the reference's own backup format) from a ring laid out by
    simple_distributed_rl_b200/checkpoint.py, with the items in SLOT order so that the reference's tree has the device's leaf order;
  * `random.random` (proportional_memory.py:147) returns the uniforms of the device's Philox stream (seed, STREAM_SAMPLE, (i | k<<16,
    step)) in the order the reference consumes them; `random.sample` (replay_buffer.py:35) returns the items of the device's distinct
    uniform picks (oracle/sumtree.py::uniform_sample_distinct -- a stated design choice);
  * `torch.randn` inside NoisyLinear (srl/rl/torch_/modules/noisy_linear.py:35-52) returns the device's Philox / Box-Muller draws
    (oracle/engine.py::default_noise_fn, equal to srlx_noise_fill to 2e-5).
The GPU test (tests/test_gpu_parity.py::test_device_learner_equals_reference_trainer_on_frozen_memory) loads the same ring into a
DeviceEngine and compares every update with what the reference did: leaf indices (exact), IS weights, target_q, loss, |td|, the
parameters after Adam, the target network, the leaf priorities after the update.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import srl  # noqa: E402  (the reference)
from srl.algorithms import dqn, rainbow  # noqa: E402

from oracle import engine as oeng  # noqa: E402
from oracle import nets, sumtree  # noqa: E402
from oracle.ref_envs import register_restated_envs  # noqa: E402
from simple_distributed_rl_b200 import checkpoint as ck  # noqa: E402
from synth_ring import ring_seed_of, synth_ring  # noqa: E402  (tests/golden/synth_ring.py: shared with the tests)

CASES = {
    # BASELINE configs[2] shape: learner_fast_kernel<16>; the ring has wrapped (vec_steps > R)
    "rainbow_default_per_m3": dict(env="CartPole-v1", algo="rainbow", hidden=(512,), dueling="average", noisy=True, mem_kind=1,
                                   multisteps=3, n_envs=32, ring_rows=32, batch_size=32, vec_steps=39, zero_leaves=0.05),
    # BASELINE configs[1] shape: learner_small_kernel, uniform replay
    "dqn_mlp64x64_uniform": dict(env="CartPole-v1", algo="dqn", hidden=(64, 64), dueling=None, noisy=False, mem_kind=0, multisteps=1,
                                 n_envs=64, ring_rows=16, batch_size=32, vec_steps=16),
    # learner_small_kernel with its replay CTA, no duplicates in a batch
    "dqn_mlp64x64_per_nodup": dict(env="CartPole-v1", algo="dqn", hidden=(64, 64), dueling=None, noisy=False, mem_kind=1, multisteps=1,
                                   n_envs=64, ring_rows=16, batch_size=32, vec_steps=21, has_duplicate=False, enable_double_dqn=False),
    # generic learner_kernel: NoisyNet with two hidden layers (Grid shapes: 2 observation floats, 4 actions)
    "rainbow_noisy_mlp32x16_per_m1": dict(env="Grid", algo="rainbow", hidden=(32, 16), dueling=None, noisy=True, mem_kind=1, multisteps=1,
                                          n_envs=24, ring_rows=20, batch_size=16, vec_steps=20),
    # dueling (max) head on 5 outputs, 3-step Retrace with h = 0.9, rescaling, no double DQN: learner_small_kernel + replay CTA
    "rainbow_duelmax64x64_per_m3_rescale": dict(env="Grid", algo="rainbow", hidden=(64, 64), dueling="max", noisy=False, mem_kind=1,
                                                multisteps=3, n_envs=16, ring_rows=40, batch_size=16, vec_steps=55, retrace_h=0.9,
                                                enable_double_dqn=False, enable_rescale=True),
    # the reference's default DQN (one hidden layer of 512) on uniform replay: learner_fast_kernel, parallel first-attempt draws
    "dqn_default512_uniform": dict(env="CartPole-v1", algo="dqn", hidden=(512,), dueling=None, noisy=False, mem_kind=0, multisteps=1,
                                   n_envs=48, ring_rows=12, batch_size=32, vec_steps=12),
    # a tree deeper than the 12 levels the fast learner caches in shared memory: 2^16 leaves, one deep round below the cache
    "rainbow_default_per_m3_deep": dict(env="CartPole-v1", algo="rainbow", hidden=(512,), dueling="average", noisy=True, mem_kind=1,
                                        multisteps=3, n_envs=1024, ring_rows=64, batch_size=32, vec_steps=64, store_ring=False),
}
ENGINE_SEED = 11
N_UPDATES = 3
TARGET_INTERVAL = 2  # a sync after update 0 and after update 2 (train_count % interval == 0 before the increment)
ALPHA, BETA0, BETA_STEPS, PER_EPS = 0.6, 0.4, 1000, 1e-4  # beta moves visibly within three updates


def slot_ordered_backup(v, per, seed):
    """The reference memory backup with item i = ring slot i (the device's leaf order); slots whose window is not complete hold no
    item and a zero leaf."""
    E, R = v.E, v.R
    g_lo, n_g = v.valid_rows()
    items, _ = ck.export_items(v, pad_action=lambda e, g: ck.philox_pad_action(seed, e, g, v.A))
    by_slot = [None] * (E * R)
    for e in range(E):
        for j in range(n_g):
            by_slot[((g_lo + j) % R) * E + e] = items[e * n_g + j]
    if not per:
        # uniform replay: the list holds the valid items in the device's pick order (time-major over the valid rows)
        return [[by_slot[((g_lo + j) % R) * E + e] for j in range(n_g) for e in range(E)], 0], by_slot
    cap = E * R
    tree = ck.build_sum_tree(v.leaf_priority, cap)
    return [cap, float(v.max_priority), n_g * E, (v.vec_steps % R) * E % cap, tree.tolist(), by_slot], by_slot


def gen(case, kw):
    torch.manual_seed(7)
    algo, M, B, noisy = kw["algo"], kw["multisteps"], kw["batch_size"], kw["noisy"]
    E, R = kw["n_envs"], kw["ring_rows"]
    cap = E * R
    per = bool(kw["mem_kind"])
    ring_seed = ring_seed_of(case)
    v = synth_ring(kw, ring_seed)
    if algo == "dqn":
        cfg = dqn.Config(batch_size=B, enable_double_dqn=kw.get("enable_double_dqn", True), enable_rescale=kw.get("enable_rescale", False),
                         target_model_update_interval=TARGET_INTERVAL)
        cfg.hidden_block.set(kw["hidden"])
    else:
        cfg = rainbow.Config(batch_size=B, enable_double_dqn=kw.get("enable_double_dqn", True), enable_rescale=kw.get("enable_rescale", False),
                             multisteps=M, enable_noisy_dense=noisy, retrace_h=kw.get("retrace_h", 1.0),
                             target_model_update_interval=TARGET_INTERVAL)
        if kw["dueling"] is None:
            cfg.hidden_block.set(kw["hidden"])
        else:
            cfg.hidden_block.set_dueling_network(kw["hidden"], dueling_type=kw["dueling"])
    if per:
        cfg.memory.set_proportional(alpha=ALPHA, beta_initial=BETA0, beta_steps=BETA_STEPS, has_duplicate=kw.get("has_duplicate", True),
                                    epsilon=PER_EPS)
    else:
        cfg.memory.set_replay_buffer()
    cfg.memory.capacity = cap
    cfg.memory.warmup_size = B
    cfg.memory.compress = False
    runner = srl.Runner(kw["env"], cfg)
    runner.set_device("CPU")
    parameter = runner.make_parameter()
    memory = runner.make_memory()
    D, A = v.D, v.A
    spec = nets.NetSpec(D, tuple(kw["hidden"]), A, kw["dueling"], noisy)
    with torch.no_grad():  # de-synchronise target from online so the two nets differ
        for p in parameter.q_target.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    mu0, sig0 = spec.from_state_dict(parameter.q_online.state_dict(), algo)
    tmu0, tsig0 = spec.from_state_dict(parameter.q_target.state_dict(), algo)
    assert spec.n_params == len(mu0)

    inner, by_slot = slot_ordered_backup(v, per, ENGINE_SEED)
    memory.call_restore([inner, None])
    assert memory.length() == v.valid_rows()[1] * E
    trainer = cfg.make_trainer(parameter, memory)
    trainer.on_setup()

    # ---- random sources ------------------------------------------------------------------------------------------------
    noise_fn = oeng.default_noise_fn(ENGINE_SEED, spec.n_params)
    feed = {"uniforms": [], "picks": None}
    real_randn = torch.randn

    def fake_random():
        return feed["uniforms"].pop(0)

    def fake_sample(population, k):
        assert k == B and len(population) == memory.length()
        return [population[i] for i in feed["picks"]]

    import srl.rl.memories.priority_memories.proportional_memory as pm
    import srl.rl.memories.priority_memories.replay_buffer as rb

    class _FedRandom:  # stands in for the `random` module inside the two memory modules only
        random = staticmethod(fake_random)
        sample = staticmethod(fake_sample)

    real_pm_random, real_rb_random = pm.random, rb.random
    pm.random, rb.random = _FedRandom, _FedRandom
    state = {"u": 0, "call": 0}
    if noisy:
        import srl.rl.torch_.modules.noisy_linear as nl
        from make_golden import _noise_schedule

        sched = _noise_schedule(spec, parameter.q_online)
        # pass order inside Trainer.train: n-step rainbow: online(s') [1], target(s') [2], online(s) [0] (rainbow.py:219-220);
        # 1-step variants: target(s') [2], online(s') [1] (only with double DQN), online(s) [0] (rainbow_nomultisteps.py:22-26)
        if algo == "rainbow" and M > 1:
            order = [1, 2, 0]
        else:
            order = [2, 1, 0] if kw.get("enable_double_dqn", True) else [2, 0]
        noises = {}

        def fake_randn(size, **k2):
            per_pass = len(sched)
            c = state["call"]
            p = order[c // per_pass]
            off, n, shape = sched[c % per_pass]
            assert tuple(size) == tuple(shape), (size, shape)
            state["call"] += 1
            key = (state["u"], p)
            if key not in noises:
                noises[key] = noise_fn(nets.NOISE_KIND_TRAIN, state["u"] * 3 + p)
            return torch.tensor(noises[key][off:off + n].reshape(shape))

        nl.torch.randn = fake_randn

    # ---- taps ------------------------------------------------------------------------------------------------------------
    rec = dict(idx=[], weights=[], target_q=[], loss=[], td=[], mu=[], sigma=[], tmu=[], tsigma=[], leaves=[], maxp=[], total=[],
               update_step=[])
    orig_sample, orig_update = memory.sample, memory.update

    def tap_sample(*a, **k):
        r = orig_sample(*a, **k)
        rec["weights"].append(np.asarray(r[1], dtype=np.float32).copy())
        rec["idx"].append(np.asarray(r[2] if per else feed["slots"], dtype=np.int64).copy())
        return r

    def tap_update(update_args, priorities, step):
        rec["td"].append(np.asarray(priorities).copy())
        rec["update_step"].append(int(step))
        return orig_update(update_args, priorities, step)

    memory.sample, memory.update = tap_sample, tap_update
    if algo == "rainbow" and M == 1:
        import srl.algorithms.rainbow.model_torch as mt

        orig_ct = mt.calc_target_q

        def wrap(*a, **k):
            r = orig_ct(*a, **k)
            rec["target_q"].append(np.array(r[0]).copy())
            return r

        mt.calc_target_q = wrap
    else:
        orig_ct = parameter.calc_target_q

        def wrap(*a, **k):
            r = orig_ct(*a, **k)
            rec["target_q"].append(np.array(r[0] if isinstance(r, tuple) else r).copy())
            return r

        parameter.calc_target_q = wrap

    g_lo, n_g = v.valid_rows()
    try:
        for u in range(N_UPDATES):
            state["u"], state["call"] = u, 0
            if per:
                # replay the reference's own consumption order of random.random(): sample i, attempt k (k grows on a rejection)
                tree_now = np.asarray(memory.memory.tree.tree, dtype=np.float64)
                sim = sumtree.ProportionalMemory(cap, ALPHA, BETA0, BETA_STEPS, kw.get("has_duplicate", True), PER_EPS)
                sim.tree.tree[:] = tree_now
                sim.size = memory.memory.size
                attempts = []
                uf = sumtree.philox_uniforms(ENGINE_SEED, u)

                def logged(i, k):
                    x = uf(i, k)
                    attempts.append(x)
                    return x

                sim_idx, _, _, _ = sim.sample(B, max(u - 1, 0), logged)
                feed["uniforms"] = list(attempts)
            else:
                pick = sumtree.uniform_sample_distinct(n_g * E, B, ENGINE_SEED, u)
                feed["picks"] = [int(p) for p in pick]
                feed["slots"] = ((g_lo + pick // E) % R) * E + pick % E
            trainer.train()
            if per:
                assert not feed["uniforms"], "the reference consumed fewer uniforms than the replay predicted"
                assert list(rec["idx"][-1]) == list(sim_idx), "restated sampler and the reference disagree"
                t = np.asarray(memory.memory.tree.tree, dtype=np.float64)
                rec["leaves"].append(t[cap - 1:].copy())
                rec["maxp"].append(float(memory.memory.max_priority))
                rec["total"].append(float(t[0]))
            rec["loss"].append(float(trainer.info["loss"]))
            m_, s_ = spec.from_state_dict(parameter.q_online.state_dict(), algo)
            tm_, ts_ = spec.from_state_dict(parameter.q_target.state_dict(), algo)
            rec["mu"].append(m_)
            rec["sigma"].append(s_ if s_ is not None else np.zeros(0, np.float32))
            rec["tmu"].append(tm_)
            rec["tsigma"].append(ts_ if ts_ is not None else np.zeros(0, np.float32))
    finally:
        pm.random, rb.random = real_pm_random, real_rb_random
        if noisy:
            nl.torch.randn = real_randn
        if algo == "rainbow" and M == 1:
            mt.calc_target_q = orig_ct
    assert trainer.train_count == N_UPDATES

    z = np.zeros(0, np.float32)
    out = dict(
        kw=np.array(repr({k: v_ for k, v_ in kw.items() if k not in ("vec_steps", "zero_leaves", "store_ring")})),
        vec_steps=kw["vec_steps"], ring_seed=ring_seed, engine_seed=ENGINE_SEED, target_interval=TARGET_INTERVAL,
        per=np.array([ALPHA, BETA0, BETA_STEPS, PER_EPS]), discount=cfg.discount, lr=cfg.lr,
        mu0=mu0, sigma0=sig0 if sig0 is not None else z, tmu0=tmu0, tsigma0=tsig0 if tsig0 is not None else z,
        idx=np.array(rec["idx"]), weights=np.array(rec["weights"]), target_q=np.array(rec["target_q"]), loss=np.array(rec["loss"]),
        td=np.array(rec["td"]), mu_after=np.array(rec["mu"]), sigma_after=np.array(rec["sigma"]), tmu_after=np.array(rec["tmu"]),
        tsigma_after=np.array(rec["tsigma"]), update_step=np.array(rec["update_step"]), maxp_after=np.array(rec["maxp"]),
        total_after=np.array(rec["total"]), max_priority0=v.max_priority,
        ring_checksum=np.array([float(v.obs.astype(np.float64).sum()), float(v.reward.astype(np.float64).sum()), float(v.action.sum()),
                                float(0.0 if v.leaf_priority is None else v.leaf_priority.sum())]),
    )
    if per:
        # leaves the updates touched (all of them for small trees)
        touched = np.unique(np.concatenate(rec["idx"])) - (cap - 1)
        out.update(touched=touched, leaves_after=np.array([l[touched] for l in rec["leaves"]]))
    if kw.get("store_ring", True):
        out.update(ring_obs=v.obs, ring_next_obs=v.next_obs, ring_action=v.action, ring_reward=v.reward, ring_term=v.term,
                   ring_done=v.done, leaf_priority=v.leaf_priority if v.leaf_priority is not None else np.zeros(0))
    np.savez_compressed(os.path.join(HERE, f"learner_{case}.npz"), **out)
    print(f"learner_{case}: n_params={spec.n_params} losses={rec['loss']} idx[0][:6]={rec['idx'][0][:6]}", flush=True)


if __name__ == "__main__":
    register_restated_envs()
    only = sys.argv[1:]
    for case, kw in CASES.items():
        if only and case not in only:
            continue
        gen(case, kw)
