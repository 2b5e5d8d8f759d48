"""GPU tests of PPO on the device (csrc/ppo.cu, simple_distributed_rl_b200/ppo.py) against oracle/ppo.py -- a torch RESTATEMENT of the
reference's TensorFlow code (srl/algorithms/ppo/ppo.py; TensorFlow is not available, so this row's parity is by restatement; the
worker's GAE / MC accumulation is pinned separately against goldens from the reference's own Worker.on_step) -- and the reference's
own acceptance gates (tests/algorithms_/base_ppo.py)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import envs as oenvs  # noqa: E402
from oracle import gae as ogae  # noqa: E402
from oracle import ppo as oppo  # noqa: E402


def _cfg(**kw):
    from simple_distributed_rl_b200.ppo import PPOConfig

    return PPOConfig(**kw)


@pytest.mark.parametrize("env,blocks", [("Pendulum-v1", ((64, 64), (64,), (64,))), ("Pendulum-v1", ((128,), (), (32, 16))),
                                        ("Grid", ((64, 64), (), ())), ("CartPole-v1", ((32,), (16,), (16,)))])
def test_rollout_rows_equal_the_restated_worker(env, blocks):
    """Worker.policy + env.step for E env copies: V(s) and log_prob of every stored step against the oracle forward on the stored
    state (1e-4), the continuous action = loc + exp(log_scale) * N(0,1) with the device's Philox draw, the env transition (next
    stored state = the oracle env stepped with the clipped, rescaled action) bit for bit, rewards with shift / scale."""
    from simple_distributed_rl_b200.ppo import PPOEngine

    cfg = _cfg(env=env, n_envs=40, horizon=12, hidden_block=blocks[0], value_block=blocks[1], policy_block=blocks[2], seed=5,
               reward_shift=1.0, reward_scale=0.1)
    eng = PPOEngine(cfg)
    eng.rollout()
    t = {k: v.cpu().numpy() for k, v in eng.t.items()}
    spec, params = eng.spec, torch.as_tensor(eng.get_params())
    obs = torch.as_tensor(t["buf_obs"].reshape(-1, eng.D))
    with torch.no_grad():
        v, po = oppo.forward(spec.layers, spec.stack_v, spec.stack_p, params, obs)
    np.testing.assert_allclose(t["buf_v"].reshape(-1), v.numpy(), rtol=1e-4, atol=1e-5)
    act = torch.as_tensor(t["buf_action"].reshape(-1))
    if eng.continuous:
        lo, hi = math.log(1e-10), math.log(10)
        ls = torch.clamp(po[:, 1], lo, hi)
        logp = oppo.normal_logprob(act, po[:, 0], ls).numpy()
        z = np.array([[oppo.policy_noise(5, e, g) for e in range(cfg.n_envs)] for g in range(cfg.horizon)], dtype=np.float32).reshape(-1)
        np.testing.assert_allclose(act.numpy(), (po[:, 0] + torch.exp(ls) * torch.as_tensor(z)).numpy(), rtol=1e-4, atol=2e-5)
    else:
        logp = torch.log_softmax(po, dim=-1).gather(1, act.long()[:, None])[:, 0].numpy()
        assert set(np.unique(t["buf_action"])) <= set(float(a) for a in range(eng.env.n_actions))
    np.testing.assert_allclose(t["buf_logp"].reshape(-1), np.maximum(logp, math.log(1e-6)), rtol=1e-4, atol=2e-5)
    # env transitions: stored state of step g+1 == the oracle env stepped from step g's state with the stored action
    spec_env = oenvs.make_spec(env)
    for e in range(0, cfg.n_envs, 7):
        st = spec_env.reset(5, e, 0)
        episode = 0
        for g in range(cfg.horizon):
            np.testing.assert_array_equal(t["buf_obs"][g, e], spec_env.obs(st))
            a = t["buf_action"][g, e]
            if eng.continuous:
                u = min(max((np.float64(a) + 1.0) * 0.5 * 4.0 - 2.0, -2.0), 2.0)
                st, r, term = spec_env.step_torque(st, u)
            else:
                st, r, term = spec_env.step(st, int(a), 5, e, g)
            assert t["buf_reward"][g, e] == np.float32((r + 1.0) * 0.1)
            if t["buf_done"][g, e]:
                episode += 1
                st = spec_env.reset(5, e, episode)
    assert eng.read_state().vec_steps == cfg.horizon


@pytest.mark.parametrize("method,env", [("GAE", "Pendulum-v1"), ("MC", "Pendulum-v1"), ("GAE", "Grid")])
def test_finish_rollout_equals_the_worker_accumulation(method, env):
    """Worker.on_step at episode end (ppo.py:375-404) for the whole buffer: V(s) with the current parameters against the oracle
    forward (1e-4), then the GAE / MC values against oracle/gae.py (bit-exact against the reference's own on_step goldens) on the
    device's values: exact.  Steps of episodes still running when the buffer ends are not emitted."""
    from simple_distributed_rl_b200.ppo import PPOEngine

    T = 200 if env == "Pendulum-v1" else 64
    cfg = _cfg(env=env, n_envs=48, horizon=T, experience_collection_method=method, seed=2, reward_clip=(-3.0, 0.5) if method == "MC" else None)
    eng = PPOEngine(cfg)
    eng.rollout()
    eng.finish_rollout()
    t = {k: v.cpu().numpy() for k, v in eng.t.items()}
    reward = ogae.clip_reward(t["buf_reward"], cfg.reward_clip)
    if method == "GAE":
        with torch.no_grad():
            v, _ = oppo.forward(eng.spec.layers, eng.spec.stack_v, eng.spec.stack_p, torch.as_tensor(eng.get_params()),
                                torch.as_tensor(t["buf_obs"].reshape(-1, eng.D)))
        np.testing.assert_allclose(t["buf_vnew"][:T].reshape(-1), v.numpy(), rtol=1e-4, atol=1e-5)
        want, valid = ogae.returns_scan(reward, t["buf_vnew"][:T], t["buf_vnew"][1:T + 1], t["buf_done"], cfg.discount, cfg.gae_discount, ogae.METHOD_GAE)
    else:
        want, valid = ogae.returns_scan(reward.astype(np.float64), None, None, t["buf_done"], cfg.discount, cfg.gae_discount, ogae.METHOD_MC)
    np.testing.assert_array_equal(t["buf_valid"], valid)
    np.testing.assert_array_equal(t["buf_ret"][valid.astype(bool)], want[valid.astype(bool)])
    if env == "Pendulum-v1":
        assert valid.all()  # 200-step episodes fill the 200-row buffer exactly
    else:
        assert 0 < valid.sum() < valid.size


@pytest.mark.parametrize("kw", [
    dict(env="Pendulum-v1"),                                                              # the reference's defaults: GAE, advantage, clip, value clip
    dict(env="Pendulum-v1", baseline_type="normal", enable_value_clip=False, surrogate_type="", enable_state_normalized=True,
         hidden_block=(128,), value_block=(128,), policy_block=(128,), experience_collection_method="MC", entropy_weight=0.1),
    dict(env="Grid", baseline_type="ave", hidden_block=(64, 64), value_block=(), policy_block=(), lr_decay_steps=3),
    dict(env="CartPole-v1", baseline_type="std", batch_size=16, global_gradient_clip_norm=0.0, lr=1e-3),
], ids=["pendulum_defaults", "pendulum_mc_normal_noclip_statenorm", "grid_ave_decay", "cartpole_std_b16_noclipnorm"])
def test_update_equals_the_restated_trainer(kw):
    """Trainer._train (ppo.py:208-291): six consecutive minibatch updates, each compared with the oracle update on the SAME minibatch
    (the device's own distinct picks): the clipped gradient 1e-3, the three loss terms and the parameters after keras-Adam 1e-4;
    then the oracle continues from the device's parameters (as the Q-learning lockstep tests do)."""
    from simple_distributed_rl_b200.ppo import PPOEngine

    cfg = _cfg(n_envs=32, horizon=200 if kw["env"] == "Pendulum-v1" else 64, seed=3, **kw)
    eng = PPOEngine(cfg, debug=True)
    eng.rollout()
    eng.finish_rollout()
    t = {k: v.cpu().numpy() for k, v in eng.t.items()}
    obs, act = t["buf_obs"].reshape(-1, eng.D), t["buf_action"].reshape(-1)
    oldv, oldlp, ret, valid = t["buf_v"].reshape(-1), t["buf_logp"].reshape(-1), t["buf_ret"].reshape(-1), t["buf_valid"].reshape(-1)
    adam = oppo.KerasAdam(eng.spec.n_params, cfg.lr, cfg.lr_decay_steps, cfg.lr_decay_rate)
    params = eng.get_params()
    seen = set()
    for u in range(6):
        eng.learn(1)
        idx = eng.t["dbg_idx"].cpu().numpy()
        assert len(set(idx.tolist())) == cfg.batch_size and valid[idx].all()  # distinct, only emitted steps
        seen |= set(idx.tolist())
        new_p, info = oppo.train_update(eng.spec, params, adam, cfg, eng.continuous, obs[idx], act[idx], oldv[idx], oldlp[idx], ret[idx])
        ps = eng.read_pstate()
        assert ps.train_count == u + 1 and ps.adam_step == u + 1
        np.testing.assert_allclose(eng.t["dbg_grads"].cpu().numpy(), info["grad"], rtol=1e-3, atol=2e-6)
        assert math.isclose(ps.grad_norm, info["grad_norm"], rel_tol=1e-4)
        for a, b in ((ps.policy_loss, info["policy_loss"]), (ps.value_loss, info["value_loss"]), (ps.entropy_loss, info["entropy_loss"])):
            assert math.isclose(a, b, rel_tol=1e-4, abs_tol=1e-6), (u, a, b)
        got = eng.get_params()
        np.testing.assert_allclose(got, new_p, rtol=1e-4, atol=2e-6)
        params = got  # continue from the device's parameters; the oracle's Adam moments follow its own (1e-3-close) gradients
    assert len(seen) > cfg.batch_size  # different minibatches
    eng2 = PPOEngine(cfg, params=eng.spec.init_params(cfg.seed, eng.continuous))
    eng2.rollout(); eng2.finish_rollout(); eng2.learn(6)   # six updates in one launch == six launches of one
    assert torch.equal(eng2.t["params"], eng.t["params"]) and torch.equal(eng2.t["adam_v"], eng.t["adam_v"])


def test_learning_easygrid_reaches_reference_baseline():
    """tests/algorithms_/base_ppo.py:49-81 (test_EasyGrid1): hidden (64, 64), no value / policy blocks, GAE, no baseline, clip, value
    clip, lr 5e-4, warmup 500, train_num 50; the env's baseline: mean reward >= 0.9 over 100 episodes."""
    from simple_distributed_rl_b200.ppo import PPORunner

    cfg = _cfg(env="EasyGrid", n_envs=256, horizon=64, hidden_block=(64, 64), value_block=(), policy_block=(), experience_collection_method="GAE",
               baseline_type="", surrogate_type="clip", enable_value_clip=True, lr=0.0005, warmup_size=500, train_num=50, lr_decay_steps=0, seed=1)
    r = PPORunner(cfg)
    st = r.train(max_train_count=5000)
    assert st.train_count == 5000 and st.total_step > 0
    assert float(np.mean(r.evaluate(max_episodes=100))) >= 0.9


def test_learning_pendulum_reaches_reference_baseline():
    """tests/algorithms_/base_ppo.py:105-123 (test_Pendulum_continue): MC, baseline "advantage", clip, no value clip, lr 2e-4 (constant
    here: the reference's default staircase decay would divide it by 100 after 2000 updates), blocks of 128, discount 0.9, entropy
    0.1, reward (r + 1) / 10, 40 000 updates; the env's baseline: mean reward >= -500 over 10 evaluation episodes."""
    from simple_distributed_rl_b200.ppo import PPORunner

    cfg = _cfg(env="Pendulum-v1", n_envs=64, horizon=200, hidden_block=(128,), value_block=(128,), policy_block=(128,),
               experience_collection_method="MC", baseline_type="advantage", surrogate_type="clip", enable_value_clip=False, lr=0.0002,
               lr_decay_steps=0, warmup_size=400, train_num=50, discount=0.9, entropy_weight=0.1, reward_shift=1.0, reward_scale=0.1, seed=1)
    r = PPORunner(cfg)
    st = r.train(max_train_count=40_000)
    assert st.train_count >= 40_000
    rewards = r.evaluate(max_episodes=10)
    assert float(np.mean(rewards)) >= -500.0, rewards


def test_baseline_config4_shape_runs():
    """BASELINE configs[4]: 16 384 Pendulum copies, GAE(0.95), 200-step rollout buffer (3.3 M samples): one rollout + returns + updates;
    size-independent checks (every step emitted, returns finite, parameters move, episodes counted)."""
    from simple_distributed_rl_b200.ppo import PPORunner

    cfg = _cfg(env="Pendulum-v1", n_envs=16384, horizon=200, gae_discount=0.95, seed=0)
    r = PPORunner(cfg)
    p0 = r.engine.get_params().copy()
    st = r.train(max_train_count=2000)
    assert st.total_step == 16384 * 200 and st.train_count == 2000 and st.episode_count == 16384
    t = r.engine.t
    assert bool(t["buf_valid"].all()) and bool(torch.isfinite(t["buf_ret"]).all())
    assert np.isfinite(r.engine.get_params()).all() and np.abs(r.engine.get_params() - p0).max() > 0
    assert -2000.0 < st.mean_episode_reward < 0.0
