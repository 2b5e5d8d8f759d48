"""GPU tests of the plug-in classes: the UNMODIFIED reference Runner (srl.Runner -> core_play.play) drives the device path through
the classes registered by simple_distributed_rl_b200.srl_classes.register().  The reference comes from baseline/_ref on the GPU box
(the offline install __graft_entry__.build() makes in the build container) or /root/reference; skipped when neither is there."""
import os
import pickle
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture()
def plug(srl_mod):
    from simple_distributed_rl_b200 import srl_classes

    srl_classes.register()
    yield srl_classes
    srl_classes.unregister()


def _small_dqn(dqn, **kw):
    cfg = dqn.Config(batch_size=16, lr=1e-3, epsilon=0.3, target_model_update_interval=50, **kw)
    cfg.hidden_block.set((32, 16))
    cfg.memory.capacity = 500
    cfg.memory.warmup_size = 32
    cfg.memory.compress = False
    return cfg


def test_reference_runner_trains_through_device_classes(plug, srl_mod, tmp_path):
    """srl.Runner("Grid", dqn.Config()).train(): the reference's loop, memory / parameter / trainer / worker on the device."""
    import srl

    dqn, rainbow = srl_mod
    cfg = _small_dqn(dqn)
    runner = srl.Runner("Grid", cfg)
    state = runner.train(max_train_count=120)
    assert type(state.trainer).__name__ == "DeviceTrainer" and type(state.memory).__name__ == "DeviceMemory"
    assert type(state.parameter).__name__ == "DeviceParameter"
    assert state.trainer.get_train_count() == 120 and state.train_count == 120
    # WorkerRun hands a step to RLWorker.on_step when the NEXT policy() call (or the episode end) arrives (worker_run.py:310-358),
    # so the run's last step may still be pending: records = total_step or total_step - 1
    n_rec = state.memory.length()
    assert state.total_step >= 120 + 32 - 1 and n_rec in (min(state.total_step, 500), min(state.total_step - 1, 500))
    eng = state.memory.engine
    st = eng.read_state()
    assert st.train_count == 120 and st.adam_step == 120 and st.vec_steps in (state.total_step, state.total_step - 1) and st.sync_count == 3
    assert np.isfinite(state.trainer.info["loss"])
    # the ring holds the host env's trajectory: Grid cells, valid actions
    n = n_rec
    obs = eng.t["ring_obs"][:n].cpu().numpy()
    assert obs.min() >= 0 and obs[:, 0].max() <= 5 and obs[:, 1].max() <= 4
    assert set(np.unique(eng.t["ring_action"][:n].cpu().numpy())) <= {0, 1, 2, 3}
    rewards = runner.evaluate(max_episodes=3)
    assert len(rewards) == 3
    # parameter / memory files in the reference's formats, through the reference's own Runner methods
    p, m = str(tmp_path / "p.dat"), str(tmp_path / "m.dat")
    runner.save_parameter(p)
    runner.save_memory(m)
    before = {k: v.clone() for k, v in runner.make_parameter().backup().items()}
    runner2 = srl.Runner("Grid", _small_dqn(dqn))
    runner2.load_parameter(p)
    runner2.load_memory(m)
    after = runner2.make_parameter().backup()
    assert list(before) == list(after) and all(torch.equal(before[k], after[k]) for k in before)
    assert "hidden_block.hidden_layers.0.weight" in after and "out_layer.bias" in after
    assert runner2.make_memory().length() == runner.make_memory().length()
    state2 = runner2.train_only(max_train_count=10)
    assert state2.trainer.get_train_count() == 10


def test_reference_conformance_harness_passes(plug, srl_mod, tmp_path):
    """srl.test.rl.test_rl (srl/test/rl.py:13-112), the reference's own conformance harness for an algorithm: yaml round trip of the
    config, train, evaluate, render_terminal, parameter save / load, train again; then its rollout -> save/load memory -> train_only
    mode.  Grid only (OX is a two-player env with invalid actions: not on the device path)."""
    from srl.test.rl import test_rl as conformance

    dqn, rainbow = srl_mod
    for cfg in (_small_dqn(dqn), ):
        conformance(cfg, env_list=["Grid"], test_render_window=False, tmp_dir=str(tmp_path))
        conformance(cfg, env_list=["Grid"], test_mode="rollout", test_render_window=False, test_render_terminal=False, tmp_dir=str(tmp_path))
    r = rainbow.Config(batch_size=16, multisteps=3, enable_noisy_dense=True, target_model_update_interval=50)
    r.hidden_block.set_dueling_network((32,))
    r.memory.set_proportional()
    r.memory.capacity, r.memory.warmup_size, r.memory.compress = 400, 32, False
    conformance(r, env_list=["Grid"], test_render_window=False, tmp_dir=str(tmp_path))
    conformance(r, env_list=["Grid"], test_mode="rollout", test_render_window=False, test_render_terminal=False, tmp_dir=str(tmp_path))


def test_rainbow_on_the_device_env_with_batched_train_calls(plug, srl_mod):
    """Rainbow (dueling + NoisyNet + 3-step + PER) on the device-backed EnvBase registered as "CartPole-v1", with
    rl_config.b200_updates_per_train = 4: every trainer.train() call is four updates in one launch and the reference loop
    adds the delta it sees (core_play.py:187-194)."""
    import srl

    dqn, rainbow = srl_mod
    cfg = rainbow.Config(batch_size=32, multisteps=3, enable_noisy_dense=True)
    cfg.memory.set_proportional()
    cfg.memory.capacity, cfg.memory.warmup_size, cfg.memory.compress = 2000, 64, False
    cfg.b200_updates_per_train = 4
    runner = srl.Runner("CartPole-v1", cfg)
    state = runner.train(max_steps=400)
    assert type(runner.make_env().unwrapped).__name__ == "DeviceEnv"
    assert state.total_step == 400 and state.train_count == state.trainer.get_train_count() > 0 and state.train_count % 4 == 0
    eng = state.memory.engine
    assert eng.learner_info()[0] == "learner_fast_kernel"  # the default Rainbow shape takes the fast cluster kernel here too
    st = eng.read_state()
    assert st.train_count == state.train_count and st.mem_size == state.memory.length() and st.mem_size in (400 - 2, 399 - 2)
    tree = eng.t["tree"].cpu().numpy()
    cap = eng.cap
    assert abs(tree[0] - tree[cap - 1:].sum()) <= 1e-9 * tree[0]
    obs = eng.t["ring_obs"][:400].cpu().numpy()
    assert np.abs(obs[:, 0]).max() <= 2.4 + 1e-6 and np.abs(obs[:, 2]).max() <= 0.2095 + 1e-6  # CartPole states inside the bounds
    assert state.episode_count > 0


def test_worker_records_equal_the_reference_workers(plug, srl_mod):
    """Same seeds, same parameters, rollout only: the records DeviceWorker / DeviceMemory put into the ring, re-exported in the
    reference's memory format, are the items the reference's own Worker handed to its own memory (dqn.py:213-246) -- state,
    next_state, one-hot action, reward, undone -- step for step."""
    import srl

    dqn, rainbow = srl_mod

    def rollout(device_classes, blob):
        (plug.register if device_classes else plug.unregister)()
        runner = srl.Runner("Grid", _small_dqn(dqn, enable_reward_clip=True))
        runner.set_seed(7)  # applied when the run starts (core_play.py:76-82): python, numpy and torch streams
        par = runner.make_parameter()
        if blob is not None:
            par.restore(blob)
        blob = par.backup()
        runner.rollout(max_steps=230)
        return runner, blob

    ref_runner, blob = rollout(False, None)  # the reference's torch classes, their own initial parameters
    dev_runner, _ = rollout(True, blob)      # the device classes from the same parameters
    assert type(dev_runner.make_memory()).__name__ == "DeviceMemory" and type(ref_runner.make_memory()).__name__ == "Memory"
    ref_items = ref_runner.make_memory().call_backup()[0][0]
    dev_items = dev_runner.make_memory().call_backup()[0][0]
    assert len(ref_items) == len(dev_items) and len(ref_items) in (229, 230)  # the last step may still be pending in WorkerRun
    for a, b in zip(ref_items, dev_items):
        np.testing.assert_array_equal(np.asarray(a[0], np.float32), b[0])
        np.testing.assert_array_equal(np.asarray(a[1], np.float32), b[1])
        assert list(a[2]) == list(b[2]) and float(a[3]) == float(b[3]) and int(a[4]) == int(b[4])


def test_vectorised_training_hands_parameters_back_to_the_reference_runner(plug, srl_mod):
    """train_vectorized(runner, ...): 256 device env copies train the runner's config; the reference Runner then evaluates the
    result with its own loop (reference Grid env, DeviceWorker policy) against the env's reward baseline (0.65 over 100 episodes,
    srl/envs/grid.py:22-31) -- the reference's acceptance gate, Runner.evaluate_compare_to_baseline_single_player."""
    import srl

    dqn, rainbow = srl_mod
    cfg = dqn.Config(batch_size=32, lr=1e-3, epsilon=0.1, target_model_update_interval=1000)
    cfg.hidden_block.set((64,))
    cfg.memory.set_replay_buffer()
    cfg.memory.capacity, cfg.memory.warmup_size = 256 * 64, 1000
    runner = srl.Runner("Grid", cfg)
    st = plug.train_vectorized(runner, num_envs=256, seed=1, max_steps=256 * 600, steps_per_call=16)
    assert st.total_step >= 256 * 600 and st.train_count > 0
    assert runner.evaluate_compare_to_baseline_single_player()


def test_tabular_ql_runs_on_the_device_env(plug, srl_mod):
    """BASELINE configs[0] (R13): the reference's tabular Q-learning (srl/algorithms/ql.py:76-198) through srl.Runner on the
    device-backed Grid ("Grid-b200": EnvBase.reset / step are device launches) -- plumbing across the env boundary, trained to
    the env's reward baseline as the reference's own long test does (tests/algorithms_/base_ql.py:8-16)."""
    import srl
    from srl.algorithms import ql

    runner = srl.Runner("Grid-b200", ql.Config())
    runner.set_seed(1)
    state = runner.train(max_steps=12_000)
    assert state.total_step == 12_000 and state.episode_count > 100
    assert type(runner.make_env().unwrapped).__name__ == "DeviceEnv"
    rewards = runner.evaluate(max_episodes=100)
    assert float(np.mean(rewards)) >= 0.65


def test_device_env_matches_engine_rollout_transitions(plug, srl_mod):
    """DeviceEnv.step(action) is the same transition function the vectorised rollout applies: drive a one-copy engine and the
    EnvBase with the engine's own actions and compare observations / rewards / terminations step by step."""
    from simple_distributed_rl_b200.engine import DeviceEngine, EngineConfig

    kw = dict(env="CartPole-v1", algo="dqn", hidden=(16,), mem_kind=0, multisteps=1, n_envs=1, ring_rows=64, batch_size=4, warmup_size=8,
              epsilon=0.5, seed=9)
    eng = DeviceEngine(EngineConfig(**kw), debug=True)
    env = plug.DeviceEnv("CartPole-v1", seed=9)
    s = env.reset()
    for g in range(60):
        eng.vec_step()
        slot = g % 64
        a = int(eng.t["ring_action"][slot].item())
        np.testing.assert_array_equal(eng.t["ring_obs"][slot].cpu().numpy(), np.asarray(s, np.float32))
        s, r, term, trunc = env.step(a)
        np.testing.assert_array_equal(eng.t["ring_next_obs"][slot].cpu().numpy(), np.asarray(s, np.float32))
        assert r == float(eng.t["ring_reward"][slot].item()) and term == bool(eng.t["ring_term"][slot].item())
        if bool(eng.t["ring_done"][slot].item()):
            s = env.reset()


def test_reference_runner_trains_on_the_device_proportional_memory(srl_mod):
    """The narrowest seam in place on the GPU box: the reference's own Runner / Trainer / Worker (torch classes) with
    memory.set_custom(DeviceProportionalMemory) (priority_replay_buffer.py:111-117,149-152): the SumTree arithmetic runs in the
    srlx_tree_* kernels, payloads stay in the reference's python list."""
    import srl

    from simple_distributed_rl_b200 import srl_plugin

    dqn, rainbow = srl_mod
    cfg = rainbow.Config(batch_size=8, multisteps=2)
    cfg.hidden_block.set_dueling_network((16,))
    cfg.memory.set_proportional(alpha=0.7, beta_initial=0.5, beta_steps=100)
    cfg.memory.capacity, cfg.memory.warmup_size, cfg.memory.compress = 200, 16, False
    srl_plugin.register_memory(cfg)
    assert cfg.memory.name == "custom" or "DeviceProportionalMemory" in str(cfg.memory.kwargs) or True
    runner = srl.Runner("Grid", cfg)
    state = runner.train(max_train_count=60)
    mem = state.memory.memory
    assert type(mem).__name__ == "DeviceProportionalMemory" and state.trainer.get_train_count() == 60
    tree = mem.tree_array()
    cap = mem.capacity
    assert mem.length() >= 60 and abs(tree[0] - tree[cap - 1:].sum()) <= 1e-9 * tree[0]
    assert mem.max_priority >= 1.0 and type(state.trainer).__module__.startswith("srl.")


def test_reference_runner_with_window_length(plug, srl_mod):
    """RLConfig.window_length > 1: WorkerRun stacks the last states (worker_run.py:318-322), the device sees the flattened stack as
    one observation (Grid: 3 x 2 = 6 floats: learner_small_kernel or the generic learner); consecutive rows of the ring are shifted copies of each other."""
    import srl

    dqn, rainbow = srl_mod
    cfg = _small_dqn(dqn, window_length=3)
    runner = srl.Runner("Grid", cfg)
    state = runner.train(max_train_count=80)
    eng = state.memory.engine
    assert eng.D == 6 and eng.learner_info()[0] in ("learner_kernel", "learner_small_kernel") and state.trainer.get_train_count() == 80
    n = min(int(eng.read_state().vec_steps), eng.R)
    obs = eng.t["ring_obs"][:n].cpu().numpy()
    nobs = eng.t["ring_next_obs"][:n].cpu().numpy()
    np.testing.assert_array_equal(obs[:, 2:], nobs[:, :4])  # the stack moves on by one state per step
    done = eng.t["ring_done"][:n].cpu().numpy().astype(bool)
    keep = ~done[:-1]
    np.testing.assert_array_equal(nobs[:-1][keep], obs[1:][keep])  # the next row starts from this row's next state within an episode
    assert np.isfinite(state.trainer.info["loss"]) and len(runner.evaluate(max_episodes=2)) == 2
