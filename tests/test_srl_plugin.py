"""The device path reads the reference's own config objects unchanged (srl_plugin.engine_config_from_srl).
Needs the reference importable (/root/reference: present in the build container only) -> skipped elsewhere."""
import os
import sys

import pytest

REF = "/root/reference"


def test_rainbow_config_maps_field_by_field(srl_mod):
    from simple_distributed_rl_b200 import _lib
    from simple_distributed_rl_b200.srl_plugin import engine_config_from_srl

    _, rainbow = srl_mod
    cfg = rainbow.Config(multisteps=3, enable_noisy_dense=True, lr=5e-4, discount=0.97, retrace_h=0.8, batch_size=64)
    cfg.memory.set_proportional(alpha=0.7, beta_initial=0.5, beta_steps=1234, has_duplicate=False, epsilon=1e-3)
    cfg.memory.capacity = 2_000_000
    cfg.memory.warmup_size = 5000
    cfg.set_torch() if hasattr(cfg, "set_torch") else None
    e = engine_config_from_srl("CartPole-v1", cfg, num_envs=8192)
    assert (e.algo, e.multisteps, e.noisy, e.dueling, e.hidden) == ("rainbow", 3, True, "average", (512,))
    assert (e.n_envs, e.ring_rows, e.batch_size, e.warmup_size) == (8192, 245 + 2, 64, 5000)  # ceil(capacity / E) + (M - 1)
    assert e.mem_kind == _lib.MEM_PROPORTIONAL and not e.has_duplicate
    assert (e.per_alpha, e.per_beta_initial, e.per_beta_steps, e.per_epsilon) == (0.7, 0.5, 1234, 1e-3)
    assert (e.lr, e.discount, e.retrace_h, e.target_update_interval) == (5e-4, 0.97, 0.8, 1000)


def test_dqn_config_maps_and_unsupported_raises(srl_mod):
    from simple_distributed_rl_b200 import _lib
    from simple_distributed_rl_b200.srl_plugin import engine_config_from_srl

    dqn, rainbow = srl_mod
    cfg = dqn.Config(epsilon=0.05, enable_double_dqn=False, enable_rescale=True)
    cfg.hidden_block.set((64, 64))
    cfg.memory.capacity = 1_000_000
    e = engine_config_from_srl("Grid", cfg, num_envs=4096)
    assert (e.algo, e.hidden, e.dueling, e.noisy, e.multisteps) == ("dqn", (64, 64), None, False, 1)
    assert e.mem_kind == _lib.MEM_UNIFORM and e.ring_rows == 245 and not e.enable_double_dqn and e.enable_rescale
    cfg.memory.set_rankbased()
    with pytest.raises(NotImplementedError):
        engine_config_from_srl("Grid", cfg, num_envs=64)
    cfg2 = rainbow.Config()
    cfg2.epsilon_scheduler.add_linear(1.0, 0.1, 1000).add_linear(0.1, 0.01, 1000)  # ListScheduler: tabulated per vector step
    e2 = engine_config_from_srl("Grid", cfg2, num_envs=64)
    assert len(e2.eps_table) == 2002 and e2.eps_table[0] == 1.0 and e2.eps_table[-1] == 0.01
    cfg2.epsilon_scheduler.add_linear(0.01, 0.001, 9_000_000)  # too long to tabulate: refused, not silently ignored
    with pytest.raises(NotImplementedError):
        engine_config_from_srl("Grid", cfg2, num_envs=64)
    cfg3 = dqn.Config()
    cfg3.lr_scheduler.set_step(1000, 0.5)
    with pytest.raises(NotImplementedError):
        engine_config_from_srl("Grid", cfg3, num_envs=64)


def test_epsilon_scheduler_maps_and_oracle_follows_reference_linear(srl_mod):
    """SchedulerConfig.create(config.epsilon) (scheduler.py:232-246): no phases -> Constant(epsilon); one constant -> its rate;
    one linear phase -> Linear (schedulers/linear.py), which the oracle's epsilon_at restates value for value."""
    from oracle import engine as oeng
    from simple_distributed_rl_b200.srl_plugin import engine_config_from_srl

    dqn, rainbow = srl_mod
    cfg = dqn.Config(epsilon=0.07)
    e = engine_config_from_srl("Grid", cfg, num_envs=8)
    assert (e.epsilon, e.eps_end, e.eps_phase_steps) == (0.07, 0.07, 0)
    cfg.epsilon_scheduler.set(0.3)
    e = engine_config_from_srl("Grid", cfg, num_envs=8)
    assert (e.epsilon, e.eps_phase_steps) == (0.3, 0)
    for c in (dqn.Config(), rainbow.Config()):
        c.epsilon_scheduler.set_linear(0.9, 0.05, 37)
        e = engine_config_from_srl("Grid", c, num_envs=8)
        assert (e.epsilon, e.eps_end, e.eps_phase_steps) == (0.9, 0.05, 37)
        sch = c.epsilon_scheduler.create(c.epsilon)
        o = oeng.OracleEngine.__new__(oeng.OracleEngine)
        o.cfg = oeng.EngineConfig(env="Grid", epsilon=e.epsilon, eps_end=e.eps_end, eps_phase_steps=e.eps_phase_steps)
        for step in list(range(0, 45)) + [1000]:
            assert o.epsilon_at(step) == sch.update(step).to_float()
    at = dqn.Config()
    at.set_atari_config()  # dqn.py:89-102 (the image input block is ignored: the env here is a vector env)
    e = engine_config_from_srl("Grid", at, num_envs=8)
    assert (e.epsilon, e.eps_end, e.eps_phase_steps) == (1.0, 0.1, 1_000_000)
    assert (e.hidden, e.enable_reward_clip, e.enable_double_dqn, e.target_update_interval) == ((512,), True, False, 10000)


def test_multi_phase_epsilon_schedule_is_tabulated_from_the_reference_scheduler(srl_mod):
    """Several phases / cosine / polynomial (scheduler.py:232-345, ListScheduler is stateful): the table handed to the device is the
    reference's own scheduler stepped 0, 1, 2, ... as a worker does (rainbow.py:312); the oracle reads the same table."""
    from oracle import engine as oeng
    from simple_distributed_rl_b200.srl_plugin import engine_config_from_srl

    dqn, rainbow = srl_mod
    cfg = rainbow.Config()
    cfg.epsilon_scheduler.clear()
    cfg.epsilon_scheduler.add_linear(1.0, 0.4, 5)
    cfg.epsilon_scheduler.add_cosine(0.4, 0.1, 6)
    cfg.epsilon_scheduler.add_polynomial(0.1, 0.02, 7, power=2.0)
    e = engine_config_from_srl("Grid", cfg, num_envs=8)
    assert e.eps_table is not None and len(e.eps_table) == 5 + 6 + 7 + 2 and e.eps_phase_steps == 0
    fresh = cfg.epsilon_scheduler.create(cfg.epsilon)
    want = [fresh.update(g).to_float() for g in range(40)]
    o = oeng.OracleEngine.__new__(oeng.OracleEngine)
    o.cfg = oeng.EngineConfig(env="Grid", epsilon=e.epsilon, eps_table=e.eps_table)
    assert [o.epsilon_at(g) for g in range(40)] == want
    assert want[0] == 1.0 and want[-1] == 0.02 and len(set(want)) > 12
    one = dqn.Config()
    one.epsilon_scheduler.set_cosine(0.9, 0.05, 11)
    e1 = engine_config_from_srl("Grid", one, num_envs=8)
    s1 = one.epsilon_scheduler.create(one.epsilon)
    assert list(e1.eps_table[:12]) == [s1.update(g).to_float() for g in range(12)] and e1.eps_table[-1] == 0.05


def test_register_memory_uses_reference_custom_seam(srl_mod):
    from simple_distributed_rl_b200.srl_plugin import MEMORY_ENTRY_POINT, register_memory

    dqn, _ = srl_mod
    cfg = dqn.Config()
    cfg.memory.set_proportional(alpha=0.5)
    register_memory(cfg)
    assert cfg.memory.name == "custom" and cfg.memory.kwargs["entry_point"] == MEMORY_ENTRY_POINT
    assert cfg.memory.kwargs["kwargs"]["alpha"] == 0.5


def test_pendulum_takes_the_reference_action_division(srl_mod):
    from simple_distributed_rl_b200.envspec import make_env_spec
    from simple_distributed_rl_b200.srl_plugin import engine_config_from_srl

    dqn, _ = srl_mod
    cfg = dqn.Config(enable_double_dqn=False)
    cfg.hidden_block.set((64, 64))
    e = engine_config_from_srl("Pendulum-v1", cfg, num_envs=64)  # tests/algorithms_/base_dqn.py:30-43
    assert e.env == "Pendulum-v1" and e.env_kwargs["action_division_num"] == cfg.action_division_num == 10
    spec = make_env_spec(e.env, **e.env_kwargs)
    assert (spec.obs_dim, spec.n_actions, spec.trunc_limit) == (3, 10, 200) and spec.reward_baseline["baseline"] == -500
    cfg.action_division_num = 5
    assert engine_config_from_srl("Pendulum-v1", cfg, num_envs=64).env_kwargs["action_division_num"] == 5


def test_reference_runner_runs_on_the_restated_envs(srl_mod):
    """oracle/ref_envs.py: the reference's own loop (core_play.play) steps the CPU restatements of CartPole-v1 / Pendulum-v1
    registered under the gymnasium ids (baseline tooling: BASELINE.md section 2b)."""
    sys.path.insert(0, REF)
    try:
        import srl
        from oracle.ref_envs import register_restated_envs

        dqn, _ = srl_mod
        register_restated_envs()
        for env_id, n_act in (("CartPole-v1", 2), ("Pendulum-v1", 10)):
            cfg = dqn.Config()
            cfg.hidden_block.set((16,))
            cfg.memory.warmup_size = 50
            runner = srl.Runner(env_id, cfg)
            runner.set_device("CPU")
            st = runner.train(max_steps=150, enable_progress=False)
            assert st.total_step == 150 and st.train_count > 0
            assert runner.rl_config.action_space.n == n_act  # Pendulum: the reference's own 10-way division of the torque Box
    finally:
        sys.path.remove(REF)


def test_plugin_classes_register_under_the_reference_keys_and_fail_loudly_without_cuda(srl_mod):
    """srl_classes.register() takes over "DQN:torch" / "Rainbow:torch" / "Rainbow_no_multisteps:torch" in the reference's registry
    (srl/base/rl/registration.py:228-251) and registers the device envs; unregister() restores the reference's classes; without a
    CUDA device the classes refuse to construct (no CPU fallback) instead of silently training on the host."""
    import torch

    import srl
    from srl.base.env import registration as env_reg
    from srl.base.rl import registration as rl_reg

    from simple_distributed_rl_b200 import _lib, srl_classes

    dqn, rainbow = srl_mod
    before = {k: list(v) for k, v in rl_reg._registry.items()}
    srl_classes.register()
    try:
        for key in ("DQN:torch", "Rainbow:torch", "Rainbow_no_multisteps:torch"):
            assert rl_reg._registry[key] == [f"simple_distributed_rl_b200.srl_classes:Device{n}" for n in ("Memory", "Parameter", "Trainer", "Worker")]
        assert rl_reg._registry["DQN:tensorflow"] == before["DQN:tensorflow"]  # other frameworks untouched
        for cls in (srl_classes.DeviceMemory, srl_classes.DeviceParameter, srl_classes.DeviceTrainer, srl_classes.DeviceWorker):
            assert any(b.__module__.startswith("srl.base.rl") for b in cls.__mro__)  # the reference's own base classes
        assert "Grid-b200" in env_reg._registry and "CartPole-v1" in env_reg._registry
        if not torch.cuda.is_available():
            runner = srl.Runner("Grid", dqn.Config())
            with pytest.raises(_lib.SrlxError, match="no CPU fallback"):
                runner.train(max_train_count=1)
    finally:
        srl_classes.unregister()
    assert {k: rl_reg._registry[k] for k in before} == before
