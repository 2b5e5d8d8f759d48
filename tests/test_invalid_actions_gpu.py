"""GPU tests of the invalid-action masks: records enter through srlx_ext_step_masked with the next state's
mask, the generic learner applies the reference's rules -- one-step targets fill the masked entries with the MINIMUM OF THE WHOLE BATCH's
Q matrix (dqn.py:156-165, rainbow_nomultisteps.py:19-31), n-step targets with -inf per window step (rainbow.py:236-249).  The oracle's
masked targets are pinned against the reference's own Trainer.train on records with invalid-action lists
(tests/golden/trainer_*invalid*.npz, CPU suite); here the device is compared with that oracle update by update."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import engine as oeng  # noqa: E402

CASES = {
    "dqn_double": dict(algo="dqn", hidden=(32,), multisteps=1, enable_double_dqn=True, mem_kind=0),
    "dqn_nodouble_rescale_per": dict(algo="dqn", hidden=(32, 16), multisteps=1, enable_double_dqn=False, enable_rescale=True, mem_kind=1),
    "rainbow_duel_m1": dict(algo="rainbow", hidden=(32,), dueling="average", multisteps=1, mem_kind=1),
    "rainbow_duel_m3_per": dict(algo="rainbow", hidden=(64,), dueling="average", multisteps=3, mem_kind=1),
    "rainbow_noisy_m2_nodouble": dict(algo="rainbow", hidden=(32,), dueling="max", noisy=True, multisteps=2, enable_double_dqn=False,
                                      retrace_h=0.9, mem_kind=0),
}


def _records(rng, E, D, A, p_inv=0.4, p_done=0.15):
    obs = rng.normal(size=(E, D)).astype(np.float32)
    nobs = rng.normal(size=(E, D)).astype(np.float32)
    act = rng.integers(0, A, size=E).astype(np.int32)
    rew = rng.normal(size=E).astype(np.float32)
    done = (rng.random(E) < p_done).astype(np.uint8)
    term = (done * (rng.random(E) < 0.7)).astype(np.uint8)
    inv = rng.random((E, A)) < p_inv
    inv[:, 0] &= ~inv.all(axis=1)
    mask = (inv * (1 << np.arange(A))).sum(axis=1).astype(np.uint32)
    return obs, nobs, act, rew, term, done, mask


@pytest.mark.parametrize("name", list(CASES))
def test_masked_targets_equal_the_oracle_update_by_update(name):
    from simple_distributed_rl_b200.engine import DeviceEngine, EngineConfig

    D, A, E, R = 3, 5, 4, 24
    kw = dict(env="external", env_kwargs=dict(obs_dim=D, n_actions=A), n_envs=E, ring_rows=R, batch_size=8, warmup_size=8, seed=3,
              invalid_actions=True, **CASES[name])
    dev = DeviceEngine(EngineConfig(**kw), debug=True)
    assert dev.learner_info()[0] == "learner_kernel"  # masks: the generic learner
    mu, sigma = dev.get_params()
    okw = {k: v for k, v in kw.items() if k != "invalid_actions"}
    orc = oeng.OracleEngine(oeng.EngineConfig(**okw), mu, sigma, noise_fn=lambda kind, cid: dev.noise(kind, cid))
    orc.ring_invalid = np.zeros(orc.cap, dtype=np.uint32)
    rng = np.random.default_rng(1)
    n_upd = 0
    for g in range(60):
        rec = _records(rng, E, D, A)
        dev.ext_step(*rec[:6], next_invalid=rec[6])
        orc.ext_step(*rec)
        if g < 6 or g % 2:
            continue
        dev.learn(1)
        out = orc.learn(1)[0]
        n_upd += 1
        np.testing.assert_array_equal(dev.t["dbg_sample_idx"].cpu().numpy(), out["idx"])
        np.testing.assert_allclose(dev.t["dbg_target_q"].cpu().numpy(), out["target_q"], rtol=1e-4, atol=2e-5)
        st = dev.read_state()
        assert math.isclose(st.last_loss, out["loss"], rel_tol=1e-4, abs_tol=1e-6)
        mu_d, sg_d = dev.get_params()
        np.testing.assert_allclose(mu_d, orc.mu, rtol=1e-4, atol=2e-5)
        # continue from identical parameters so that 1e-4 bounds one update, not the trajectory
        orc.adam.mu.data.copy_(torch.as_tensor(mu_d))
        if sg_d is not None and kw.get("noisy"):
            orc.adam.sigma.data.copy_(torch.as_tensor(sg_d))
        tmu, tsg = dev.get_target()
        orc.tgt_mu = tmu.copy()
        if kw.get("noisy"):
            orc.tgt_sigma = tsg.copy()
    assert n_upd >= 20
    np.testing.assert_array_equal(dev.t["ring_invalid"].cpu().numpy().astype(np.uint32), orc.ring_invalid)


def test_masks_change_the_targets():
    """the same records with and without masks give different targets (the mask path is live), and rows written before the masks
    were switched on count as unmasked"""
    from simple_distributed_rl_b200.engine import DeviceEngine, EngineConfig

    D, A, E, R = 3, 5, 4, 16
    kw = dict(env="external", env_kwargs=dict(obs_dim=D, n_actions=A), n_envs=E, ring_rows=R, batch_size=8, warmup_size=8, seed=3,
              algo="dqn", hidden=(32,), multisteps=1, mem_kind=0)
    a, b = DeviceEngine(EngineConfig(**kw), debug=True), DeviceEngine(EngineConfig(**kw), debug=True)
    rng = np.random.default_rng(2)
    for g in range(10):
        rec = _records(rng, E, D, A, p_inv=0.6)
        a.ext_step(*rec[:6])
        b.ext_step(*rec[:6], next_invalid=rec[6] if g >= 2 else None)
    assert "ring_invalid" not in a.t and a.learner_info()[0] != "learner_kernel"
    assert b.learner_info()[0] == "learner_kernel"
    assert not b.t["ring_invalid"][: 2 * E].any()
    a.learn(1)
    b.learn(1)
    np.testing.assert_array_equal(a.t["dbg_sample_idx"].cpu().numpy(), b.t["dbg_sample_idx"].cpu().numpy())
    assert not np.allclose(a.t["dbg_target_q"].cpu().numpy(), b.t["dbg_target_q"].cpu().numpy())


# ---- through the reference's own Runner -------------------------------------------------------------------------------------------------
def _masked_road_cls():
    from dataclasses import dataclass

    from srl.envs.oneroad import OneRoad

    @dataclass
    class MaskedRoad(OneRoad):
        """OneRoad (srl/envs/oneroad.py) with one forbidden action per state; stepping a forbidden action is an error, so a run that
        finishes proves that the policy never chose one."""

        def get_invalid_actions(self, player_index: int = -1):
            return [1 + (self.player_pos % (self.action - 1))]

        def step(self, action):
            assert action not in self.get_invalid_actions(), f"invalid action {action} at {self.player_pos}"
            return super().step(action)

    return MaskedRoad


MaskedRoad = None


@pytest.fixture()
def masked_env(srl_mod):
    global MaskedRoad
    from srl.base.env import registration

    from simple_distributed_rl_b200 import srl_classes

    if MaskedRoad is None:
        MaskedRoad = _masked_road_cls()
        registration.register(id="MaskedRoad-b200test", entry_point=__name__ + ":MaskedRoad", kwargs={"N": 6, "action": 4, "is_end": False},
                              check_duplicate=False)
    srl_classes.register()
    yield "MaskedRoad-b200test"
    srl_classes.unregister()


@pytest.mark.parametrize("algo,multisteps", [("dqn", 1), ("rainbow", 1), ("rainbow", 3)])
def test_reference_runner_on_an_env_with_invalid_actions(masked_env, srl_mod, algo, multisteps):
    """srl.Runner(env with invalid actions, dqn / rainbow Config).train() over the device classes: the worker masks its policy
    (dqn.py:202-207, rainbow.py:307,319-325), every record carries worker.next_invalid_actions into the ring, and the trainer runs the
    generic learner with the masks."""
    import srl

    dqn, rainbow = srl_mod
    if algo == "dqn":
        cfg = dqn.Config(batch_size=16, lr=1e-3, epsilon=0.5, target_model_update_interval=50)
        cfg.hidden_block.set((32,))
    else:
        cfg = rainbow.Config(batch_size=16, lr=1e-3, epsilon=0.5, target_model_update_interval=50, multisteps=multisteps)
        cfg.hidden_block.set_dueling_network((32,))
        cfg.memory.set_proportional()
    cfg.memory.capacity, cfg.memory.warmup_size, cfg.memory.compress = 400, 32, False
    runner = srl.Runner(masked_env, cfg)
    state = runner.train(max_train_count=150)
    assert type(state.trainer).__name__ == "DeviceTrainer" and state.trainer.get_train_count() == 150
    eng = state.memory.engine
    assert eng.learner_info()[0] == "learner_kernel"
    n = int(eng.read_state().vec_steps)
    assert n >= 150
    n = min(n, eng.R)
    masks = eng.t["ring_invalid"][:n].cpu().numpy().astype(np.int64)
    nobs = eng.t["ring_next_obs"][:n, 0].cpu().numpy().astype(np.int64)
    done = eng.t["ring_done"][:n].cpu().numpy().astype(bool)
    want = 1 << (1 + nobs % 3)
    assert np.array_equal(masks[~done], want[~done]) and masks.any()
    acts = eng.t["ring_action"][:n].cpu().numpy()
    obs = eng.t["ring_obs"][:n, 0].cpu().numpy().astype(np.int64)
    assert not np.any(acts == 1 + obs % 3)  # never the forbidden action of the state it was chosen in
    assert np.isfinite(state.trainer.info["loss"])
    assert len(runner.evaluate(max_episodes=3)) == 3


@pytest.mark.parametrize("force", ["", "generic"])
def test_learners_with_wide_observations_equal_the_oracle(force, monkeypatch):
    """observation vectors of more than 4 floats (stacked states, window_length > 1) run on learner_small_kernel (+ its replay CTA) or
    the generic learner (SRLX_LEARNER=generic): 7 floats, PER, 3-step Rainbow, no masks, against the oracle update by update."""
    from simple_distributed_rl_b200.engine import DeviceEngine, EngineConfig

    if force:
        monkeypatch.setenv("SRLX_LEARNER", force)
    D, A, E, R = 7, 3, 4, 24
    kw = dict(env="external", env_kwargs=dict(obs_dim=D, n_actions=A), n_envs=E, ring_rows=R, batch_size=8, warmup_size=8, seed=5,
              algo="rainbow", hidden=(32,), dueling="average", multisteps=3, mem_kind=1)
    dev = DeviceEngine(EngineConfig(**kw), debug=True)
    assert dev.learner_info()[0] == ("learner_kernel" if force else "learner_small_kernel")
    mu, sigma = dev.get_params()
    orc = oeng.OracleEngine(oeng.EngineConfig(**kw), mu, sigma)
    rng = np.random.default_rng(4)
    n_upd = 0
    for g in range(40):
        rec = _records(rng, E, D, A)
        dev.ext_step(*rec[:6])
        orc.ext_step(*rec[:6])
        if g < 6:
            continue
        dev.learn(1)
        out = orc.learn(1)[0]
        n_upd += 1
        np.testing.assert_array_equal(dev.t["dbg_sample_idx"].cpu().numpy(), out["idx"])
        np.testing.assert_allclose(dev.t["dbg_target_q"].cpu().numpy(), out["target_q"], rtol=1e-4, atol=2e-5)
        mu_d, _ = dev.get_params()
        np.testing.assert_allclose(mu_d, orc.mu, rtol=1e-4, atol=2e-5)
        orc.adam.mu.data.copy_(torch.as_tensor(mu_d))
        orc.tgt_mu = dev.get_target()[0].copy()
    assert n_upd >= 30
