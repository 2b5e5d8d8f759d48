"""Drop-in glue for an installed `srl` (pocokhc/simple_distributed_rl): existing `srl.algorithms.dqn.Config` /
`srl.algorithms.rainbow.Config` objects are read field by field and mapped onto the device engine, unchanged.

    import srl
    from srl.algorithms import rainbow
    from simple_distributed_rl_b200 import srl_plugin

    rl_config = rainbow.Config(multisteps=3, enable_noisy_dense=True)          # the user's existing config
    rl_config.memory.set_proportional(); rl_config.memory.capacity = 2_000_000
    runner = srl_plugin.DeviceRunner("CartPole-v1", rl_config, num_envs=8192)  # instead of srl.Runner(...)
    runner.train(max_steps=10_000_000)                                         # same stop arguments as srl.Runner.train
    rewards = runner.evaluate(max_episodes=100)
    runner.save_parameter_state_dict()  -> reference-compatible state_dict (srl/rl/torch_/helper.py:60-93)

Fields honoured (reference: srl/algorithms/dqn/dqn.py:50-103, srl/algorithms/rainbow/rainbow.py:57-108,
srl/rl/memories/priority_replay_buffer.py:17-117, srl/base/rl/config.py:41-107): batch_size, memory.{capacity,
warmup_size, name, kwargs}, epsilon, test_epsilon, lr, discount, target_model_update_interval, enable_reward_clip,
enable_double_dqn, enable_rescale, enable_noisy_dense, multisteps, retrace_h, hidden_block (MLP / DuelingNetwork layer
sizes + dueling_type), and, when present, reward_scale / reward_shift ... everything else (schedulers, image blocks,
window_length > 1, frameskip, demo memory, rank-based memories) raises NotImplementedError instead of being silently ignored.

`register_memory(rl_config)` alone swaps only the SumTree: `rl_config.memory.set_custom(ENTRY_POINT, kwargs)`
(srl/rl/memories/priority_replay_buffer.py:111-117,149-152) so the reference's own Runner / Trainer drive the device tree.
"""
from typing import Any, List, Optional

from . import _lib
from .engine import EngineConfig
from .runner import VecRunner

MEMORY_ENTRY_POINT = "simple_distributed_rl_b200.memory:DeviceProportionalMemory"


def _get(obj, name, default=None):
    return getattr(obj, name, default)


def _algo_of(rl_config) -> str:
    name = rl_config.get_name() if hasattr(rl_config, "get_name") else type(rl_config).__module__
    name = str(name).split(":")[0]
    if name == "DQN":
        return "dqn"
    if name in ("Rainbow", "Rainbow_no_multisteps"):
        return "rainbow"
    raise NotImplementedError(f"algorithm {name!r} is not on the device path (supported: DQN, Rainbow)")


def engine_config_from_srl(env: Any, rl_config: Any, num_envs: int, seed: int = 0, ring_rows: Optional[int] = None,
                           env_kwargs: Optional[dict] = None, allow_window: bool = False) -> EngineConfig:
    """Map (env id or EnvConfig, dqn/rainbow Config) -> EngineConfig.  Unsupported settings raise.  env = "external" (with
    env_kwargs = {obs_dim, n_actions}) describes an env stepped by a host loop (srl_classes.py)."""
    env_name = env if isinstance(env, str) else _get(env, "name", _get(env, "id", None))
    env_kwargs = dict(env_kwargs or {}) if isinstance(env, str) else {**dict(_get(env, "kwargs", {}) or {}), **dict(env_kwargs or {})}
    algo = _algo_of(rl_config)
    if env_name == "Pendulum-v1":  # continuous action Box: the value-based worker sees RLConfig.action_division_num torques
        env_kwargs.setdefault("action_division_num", int(_get(rl_config, "action_division_num", 10)))
    # ---- things the device path does not implement: refuse loudly
    if _get(rl_config, "window_length", 1) not in (0, 1) and not allow_window:
        # the plug-in classes get the stacked state from the reference's own WorkerRun (allow_window); the device rollout does not stack
        raise NotImplementedError("window_length > 1 is not supported by the device rollout (it is through srl_classes)")
    if _get(rl_config, "frameskip", 0) not in (0, None):
        raise NotImplementedError("frameskip is not supported on the device path")
    eps_sched = _schedule_of(rl_config, "epsilon_scheduler", float(_get(rl_config, "epsilon", 0.1)), allow_linear=True)
    if getattr(_get(rl_config, "lr_scheduler"), "schedule_type", "") not in ("", None):  # LRSchedulerConfig (rl/schedulers/lr_scheduler.py)
        # (the reference's own torch trainers cannot run one either: apply_torch_scheduler returns the torch LR-scheduler object and
        # the trainer then calls zero_grad() / step() on IT, dqn/model_torch.py:79,117-119 -- there is no behaviour to match)
        raise NotImplementedError("lr_scheduler: only the constant learning rate is supported on the device path")
    mem = rl_config.memory
    if _get(mem, "enable_demo_memory", False):
        raise NotImplementedError("demo memory is not supported on the device path")
    mk = dict(_get(mem, "kwargs", {}) or {})
    if mem.name == "ReplayBuffer":
        mem_kind, per = _lib.MEM_UNIFORM, {}
    elif mem.name in ("Proportional", "Proportional_cpp"):
        mem_kind = _lib.MEM_PROPORTIONAL
        per = dict(per_alpha=mk.get("alpha", 0.6), per_beta_initial=mk.get("beta_initial", 0.4),
                   per_beta_steps=mk.get("beta_steps", 1_000_000), per_epsilon=mk.get("epsilon", 0.0001),
                   has_duplicate=mk.get("has_duplicate", True))
    else:
        raise NotImplementedError(f"memory {mem.name!r} is not supported on the device path (ReplayBuffer, Proportional)")
    # ---- network
    hb = rl_config.hidden_block
    hk = dict(_get(hb, "kwargs", {}) or {})
    noisy = bool(_get(rl_config, "enable_noisy_dense", False))
    if hb.name == "MLP":
        hidden, dueling = tuple(hk["layer_sizes"]), None
        if hk.get("activation", "relu") != "relu":
            raise NotImplementedError("only relu activations are supported on the device path")
    elif hb.name == "DuelingNetwork":
        hidden = tuple(hk["layer_sizes"])
        dueling = hk.get("dueling_kwargs", {}).get("dueling_type", "average")
        if hk.get("mlp_kwargs", {}).get("activation", "relu") != "relu":
            raise NotImplementedError("only relu activations are supported on the device path")
    else:
        raise NotImplementedError(f"hidden block {hb.name!r} is not supported on the device path")
    multisteps = int(_get(rl_config, "multisteps", 1)) if algo == "rainbow" else 1
    if _algo_of(rl_config) == "rainbow" and str(rl_config.get_name()).startswith("Rainbow_no_multisteps"):
        multisteps = 1
    cap = int(mem.capacity)
    # an M-step window needs the M-1 rows after its first step: E * (rows - (M - 1)) items are sampleable, so the ring gets M-1
    # rows on top of ceil(capacity / E) and the reference's own checks (priority_replay_buffer.py:195-200) apply to what is reachable
    rows = int(ring_rows) if ring_rows is not None else -(-cap // int(num_envs)) + (multisteps - 1)
    reachable = int(num_envs) * (rows - (multisteps - 1))
    if rows < multisteps or reachable <= 0:
        raise ValueError(f"ring_rows = {rows} cannot hold a {multisteps}-step window")
    if not (int(rl_config.batch_size) <= int(mem.warmup_size) <= reachable):
        raise ValueError(f"assert batch_size ({rl_config.batch_size}) <= memory.warmup_size ({mem.warmup_size}) <= sampleable items "
                         f"({reachable} = {num_envs} envs x ({rows} rows - {multisteps - 1}))")
    return EngineConfig(
        env=env_name, n_envs=int(num_envs), ring_rows=int(rows), multisteps=multisteps, batch_size=int(rl_config.batch_size),
        mem_kind=mem_kind, algo=algo, enable_double_dqn=bool(rl_config.enable_double_dqn),
        enable_rescale=bool(rl_config.enable_rescale), enable_reward_clip=bool(rl_config.enable_reward_clip),
        target_update_interval=int(rl_config.target_model_update_interval), seed=int(seed),
        warmup_size=int(mem.warmup_size), epsilon=eps_sched[0], eps_end=eps_sched[1], eps_phase_steps=eps_sched[2],
        eps_table=eps_sched[3] if len(eps_sched) > 3 else None, discount=float(rl_config.discount),
        lr=float(rl_config.lr), retrace_h=float(_get(rl_config, "retrace_h", 1.0)),
        reward_shift=float(_get(rl_config, "reward_shift", 0.0) or 0.0), reward_scale=float(_get(rl_config, "reward_scale", 1.0) or 1.0),
        hidden=hidden, dueling=dueling, noisy=noisy, env_kwargs=env_kwargs, **per)


def _schedule_of(rl_config: Any, field: str, val: float, allow_linear: bool = True):
    """SchedulerConfig.create(val) (srl/rl/schedulers/scheduler.py:232-246) for what the device implements: no phases or a
    default scheduler -> Constant(val); one constant phase -> its rate; one linear phase -> (start, end, phase_steps)."""
    s = _get(rl_config, field)
    phases = list(getattr(s, "schedulers", None) or []) if s is not None else []
    if not phases or getattr(s, "default_scheduler", False):
        return (val, val, 0)
    if len(phases) == 1 and phases[0].get("name") == "constant":
        r = float(phases[0]["rate"])
        return (r, r, 0)
    if allow_linear and len(phases) == 1 and phases[0].get("name") == "linear":
        ph = phases[0]
        if int(ph["phase_steps"]) <= 0:
            raise ValueError(f"{field}: linear phase_steps must be > 0")
        return (float(ph["start_rate"]), float(ph["end_rate"]), int(ph["phase_steps"]))
    if not allow_linear:
        raise NotImplementedError(f"{field}: only a constant rate is supported on the device path")
    # anything else (several phases, cosine, polynomial, ...): step the reference's own scheduler object 0, 1, 2, ... as a worker
    # would (ListScheduler is stateful, scheduler.py:319-345) and tabulate until every phase is over and the value stands still
    total = sum(int(ph.get("phase_steps", 0) or 0) for ph in phases)
    if total > 8_000_000:
        raise NotImplementedError(f"{field}: a schedule of {total} steps is too long to tabulate")
    sch = s.create(val)
    table = [float(sch.update(g).to_float()) for g in range(total + 2)]
    if float(sch.update(total + 1000).to_float()) != table[-1]:
        raise NotImplementedError(f"{field}: the schedule does not settle after its phases; not supported on the device path")
    return (table[0], table[-1], 0, tuple(table))


class DeviceRunner(VecRunner):
    """`srl.Runner(env, rl_config)`-shaped constructor over the device engine."""

    def __init__(self, env: Any, rl_config: Any, num_envs: int = 4096, seed: int = 0, device="cuda:0", ring_rows=None):
        self.rl_config = rl_config
        super().__init__(engine_config_from_srl(env, rl_config, num_envs, seed, ring_rows), device=device)

    def evaluate(self, max_episodes: int = 10, **kw) -> List[float]:
        return super().evaluate(max_episodes=max_episodes, test_epsilon=float(_get(self.rl_config, "test_epsilon", 0.0)), **kw)

    # RLParameter.call_backup / call_restore interchange (srl/algorithms/dqn/model_torch.py:47-52)
    def save_parameter_state_dict(self):
        return self.state_dict()

    def load_parameter_state_dict(self, sd):
        self.load_state_dict(sd)


def register_memory(rl_config: Any, device: str = "cuda:0", seed: int = 0) -> Any:
    """Keep the reference Runner/Trainer/Worker, swap only the priority memory for the device SumTree."""
    mk = dict(_get(rl_config.memory, "kwargs", {}) or {})
    kw = dict(alpha=mk.get("alpha", 0.6), beta_initial=mk.get("beta_initial", 0.4), beta_steps=mk.get("beta_steps", 1_000_000),
              has_duplicate=mk.get("has_duplicate", True), epsilon=mk.get("epsilon", 0.0001), device=device, seed=seed)
    rl_config.memory.set_custom(MEMORY_ENTRY_POINT, kw)
    return rl_config
