"""The reference's plug-in classes over the device engine: `RLMemory` / `RLParameter` / `RLTrainer` / `RLWorker` implementations and
an `EnvBase`, registered with the reference's own registries so that the UNMODIFIED `srl.Runner(env, rl_config).train()` -- i.e.
`core_play.play` (srl/base/run/core_play.py:115-214) -- drives libsrlx.so.  Needs `srl` importable.

    import srl
    from srl.algorithms import dqn, rainbow                  # registers "DQN:torch", "Rainbow:torch", ...
    from simple_distributed_rl_b200 import srl_classes
    srl_classes.register()                                    # same keys, check_duplicate=False: the device classes take over
    runner = srl.Runner("CartPole-v1", rainbow.Config(...))   # the user's config object, unchanged
    runner.train(max_train_count=10_000)                      # reference loop; replay, sampling, targets, loss, Adam on the GPU
    runner.evaluate(); runner.save_parameter("p.dat")         # reference-format state_dict (interchanges with the torch classes)
    srl_classes.unregister()                                  # the reference's torch classes again

What plugs in where (reference interface -> class here):
  srl/base/rl/registration.py:228-251  register(rl_config, memory_ep, parameter_ep, trainer_ep, worker_ep, check_duplicate=False)
                                        under the reference's own keys "DQN:torch" / "Rainbow:torch" / "Rainbow_no_multisteps:torch"
                                        (a new framework tag is not possible without touching the reference: setup_device asserts
                                        framework in ("tensorflow", "torch"), srl/base/system/device.py:30-33)
  srl/base/rl/memory.py:44-150         DeviceMemory: the ring + SumTree in HBM; add() <- one record per env step (srlx_ext_step);
                                        length(); call_backup()/call_restore() in the reference's memory format (checkpoint.py);
                                        register_worker_func_custom / register_trainer_recv_func / register_trainer_send_func
  srl/base/rl/parameter.py:14-72       DeviceParameter: online/target (mu, sigma) in HBM; call_backup()/call_restore() = the
                                        reference state_dict; pred_q()/pred_target_q() (srlx_qnet_forward)
  srl/base/rl/trainer.py:14-62         DeviceTrainer.train(): `updates_per_train` consecutive Trainer.train() steps in ONE srlx_learn
                                        launch (sample -> targets -> Huber -> Adam -> priority update); train_count grows by that
                                        many -- core_play.play adds whatever delta it sees (core_play.py:187-194)
  srl/base/rl/worker.py:48-69          DeviceWorker: policy() = the reference's epsilon-greedy / noisy argmax (dqn.py:192-211,
                                        rainbow.py:301-331) over the device Q; on_step() = reward clip + one ring record
  srl/base/env/base.py:18-206,         DeviceEnv: EnvBase whose reset()/step(action) run the device's closed-form envs
  srl/base/env/registration.py:116-136 (srlx_env_reset_obs / srlx_env_step_actions); ids "Grid-b200", "CartPole-v1" (a registered id
                                        shadows gymnasium, registration.py:53-62)

This is the drop-in at the REFERENCE'S granularity: one host env, one python iteration per env step, every tensor op on the GPU.
The vectorised loop (thousands of env copies per kernel, no host in the loop) is `srl_plugin.DeviceRunner`; `train_vectorized(runner,
num_envs, ...)` below runs it for an existing `srl.Runner` and hands parameters and memory back to it.  No CPU fallback anywhere:
without CUDA (or with runner.set_device("CPU")) the constructors raise.
"""
import random
from typing import Any, List, Optional

import numpy as np
import torch

from srl.base.env.base import EnvBase
from srl.base.env import registration as env_registration
from srl.base.rl import registration as rl_registration
from srl.base.rl.memory import RLMemory
from srl.base.rl.parameter import RLParameter
from srl.base.rl.trainer import RLTrainer
from srl.base.rl.worker import RLWorker
from srl.base.spaces.array_discrete import ArrayDiscreteSpace
from srl.base.spaces.box import BoxSpace
from srl.base.spaces.discrete import DiscreteSpace

from . import _lib, checkpoint
from .engine import DeviceEngine, EngineConfig
from .envspec import make_env_spec
from .netspec import NetSpec
from .srl_plugin import engine_config_from_srl

_MOD = __name__


# ---------------------------------------------------------------------------------------------------------------------
def _device_of(config) -> torch.device:
    dev = str(getattr(config, "used_device_torch", "cuda") or "cuda")
    if not torch.cuda.is_available() or dev.startswith("cpu"):
        raise _lib.SrlxError(f"the device classes need a CUDA device (used_device_torch = {dev!r}, cuda available = "
                             f"{torch.cuda.is_available()}): there is no CPU fallback; srl_classes.unregister() restores the torch classes")
    return torch.device("cuda:0" if dev == "cuda" else dev)


def _engine_config(config) -> EngineConfig:
    """The user's dqn/rainbow Config -> EngineConfig for ONE host-driven env column (ring rows = memory.capacity + M - 1)."""
    cached = getattr(config, "_b200_engine_config", None)
    if cached is not None:
        return cached
    assert config.is_setup(), "RLConfig.setup(env) must have run (make_memory / make_parameter do it)"
    obs, act = config.observation_space, config.action_space
    if len(obs.shape) < 1 or not hasattr(act, "n"):
        raise NotImplementedError(f"device path: a vector observation and a discrete action are needed (got {obs}, {act})")
    # window_length > 1: WorkerRun stacks the last states itself (worker_run.py:318-322) and the reference's MLP input block flattens
    # them; the device sees the flattened stack as one observation vector
    ecfg = engine_config_from_srl("external", config, num_envs=1, seed=int(getattr(config, "b200_seed", 0)),
                                  env_kwargs=dict(obs_dim=int(np.prod(obs.shape)), n_actions=int(act.n)), allow_window=True)
    object.__setattr__(config, "_b200_engine_config", ecfg)
    return ecfg


# ---------------------------------------------------------------------------------------------------------------------
class DeviceMemory(RLMemory):
    """RLMemory (srl/base/rl/memory.py:44-150) + the add/sample/update protocol of RLPriorityReplayBuffer
    (srl/rl/memories/priority_replay_buffer.py:205-274), HBM-resident.  It owns the DeviceEngine (ring, SumTree, counters);
    DeviceTrainer binds the parameter's tensors into it."""

    def setup(self) -> None:
        self.ecfg = _engine_config(self.config)
        self.device = _device_of(self.config)
        self.engine = DeviceEngine(self.ecfg, device=self.device)
        self._bound_parameter = None
        D = self.engine.D
        # pinned staging of the records a host loop hands over, copied to the device in one go at flush()
        self._cap_stage = 256
        self._stage = torch.zeros((self._cap_stage, 2 * D + 5), dtype=torch.float32).pin_memory()  # ... + invalid-action bit mask
        self._masks = False  # a record with invalid actions has been seen
        self._n_stage = 0
        self._steps = 0  # rows ever added (== the device's vec_steps once flushed)
        self.step = 0    # PriorityReplayBuffer.step (priority_replay_buffer.py:247-250): kept for interface parity
        self.register_worker_func_custom(self.add, self.serialize)
        self.register_trainer_recv_func(self.sample)
        self.register_trainer_send_func(self.update)

    # ---- worker side ------------------------------------------------------------------------------------------------
    def add(self, batch: Any, priority: Optional[float] = None, serialized: bool = False) -> None:
        """batch = one env step (state, next_state, action index, reward, terminated, done[, next_invalid_actions]) -- what
        DeviceWorker.on_step hands over.  n-step windows are rebuilt from consecutive rows at sample time, so records must arrive in
        trajectory order."""
        if serialized:
            import pickle

            batch = pickle.loads(batch)
        s, ns, a, r, term, done = batch[:6]
        mask = 0
        if len(batch) > 6 and batch[6]:
            for ia in batch[6]:
                mask |= 1 << int(ia)
            self._masks = True
        D = self.engine.D
        row = self._stage[self._n_stage]
        row[:D] = torch.as_tensor(np.asarray(s, dtype=np.float32).reshape(-1))
        row[D:2 * D] = torch.as_tensor(np.asarray(ns, dtype=np.float32).reshape(-1))
        row[2 * D], row[2 * D + 1], row[2 * D + 2], row[2 * D + 3] = float(a), float(r), float(bool(term)), float(bool(done))
        row[2 * D + 4] = float(mask)  # <= 16 actions: exact in float32
        self._n_stage += 1
        self._steps += 1
        if self._n_stage == self._cap_stage:
            self.flush()

    def serialize(self, batch: Any, priority: Optional[float] = None) -> Any:
        import pickle

        return (pickle.dumps(batch), priority)

    def flush(self) -> None:
        """Staged records -> ring rows + replay add (srlx_ext_step per row: one record per vector step of the single column)."""
        n = self._n_stage
        if n == 0:
            return
        eng, D = self.engine, self.engine.D
        d = self._stage[:n].to(self.device, non_blocking=True)
        obs, nobs = d[:, :D].contiguous(), d[:, D:2 * D].contiguous()
        act = d[:, 2 * D].to(torch.int32).contiguous()
        rew = d[:, 2 * D + 1].contiguous()
        term, done = d[:, 2 * D + 2].to(torch.uint8).contiguous(), d[:, 2 * D + 3].to(torch.uint8).contiguous()
        inv = None
        if self._masks:  # from the first record with invalid actions on: masks are kept and the generic learner applies them
            eng.enable_invalid_actions()
            inv = d[:, 2 * D + 4].to(torch.int32).contiguous()
        s = eng._stream()
        with torch.cuda.device(self.device):
            for i in range(n):
                _lib.check(eng.lib.srlx_ext_step_masked(eng.c, obs[i].data_ptr(), nobs[i].data_ptr(), act[i:i + 1].data_ptr(),
                                                        rew[i:i + 1].data_ptr(), term[i:i + 1].data_ptr(), done[i:i + 1].data_ptr(),
                                                        None if inv is None else inv[i:i + 1].data_ptr(), s))
        self._keep = (obs, nobs, act, rew, term, done, inv)  # alive until the next flush (the launches are asynchronous)
        eng._holds_data = True
        self._n_stage = 0

    def length(self) -> int:
        M, R = self.engine.M, self.engine.R
        return max(0, min(self._steps - (M - 1), R - (M - 1)))

    def is_warmup_needed(self) -> bool:
        return self.length() < self.ecfg.warmup_size

    # ---- trainer side: in the single-process loop the trainer launches srlx_learn, which samples and updates inside the kernel.
    #      sample()/update() exist for the reference's mp protocol (play_mp.py:248-286 calls the registered functions by name).
    def sample(self, step: int = -1, batch_size: int = -1):
        raise NotImplementedError("DeviceMemory: sampling happens inside srlx_learn (DeviceTrainer.train); the distributed protocol "
                                  "that ships sampled batches between processes is not on the device path")

    def update(self, update_args: List[Any], priorities: np.ndarray, step: int) -> None:
        self.step = step

    # ---- persistence (RLMemory.save/load/backup/restore call these) ---------------------------------------------------------
    def call_backup(self, **kwargs) -> Any:
        self.flush()
        eng = self.engine
        seed, A = int(eng.cfg.seed) & 0xFFFFFFFFFFFFFFFF, eng.A
        return checkpoint.memory_backup(eng.ring_view(), bool(eng.per), compress=bool(getattr(self.config.memory, "compress", False)),
                                        pad_action=lambda e, g: checkpoint.philox_pad_action(seed, e, g, A))

    def call_restore(self, data: Any, **kwargs) -> None:
        eng = self.engine
        self._n_stage = 0
        v = checkpoint.memory_restore(data, eng.E, eng.R, eng.M, eng.A, eng.D, bool(eng.per))
        eng.load_ring(v)
        self._steps = int(v.vec_steps)


# ---------------------------------------------------------------------------------------------------------------------
class DeviceParameter(RLParameter):
    """RLParameter (srl/base/rl/parameter.py:14-72) + the CommonInterfaceParameter inference seam (dqn.py:127-142): the online and
    target networks as flat fp32 buffers in HBM (include/srlx.h srlx_net layout), state_dicts under the reference's module keys."""

    def setup(self) -> None:
        self.ecfg = _engine_config(self.config)
        self.device = _device_of(self.config)
        e = self.ecfg
        env = make_env_spec(e.env, **e.env_kwargs)
        self.spec = NetSpec(env.obs_dim, tuple(e.hidden), env.n_actions, e.dueling, e.noisy, e.algo)
        P, noisy = self.spec.n_params, e.noisy
        z = lambda n: torch.zeros(n, dtype=torch.float32, device=self.device)  # noqa: E731
        self.t = dict(params=z(P), target=z(P))
        if noisy:
            self.t.update(params_sigma=z(P), target_sigma=z(P))
        mu, sigma = self.spec.init_params(int(e.seed))
        self._set(mu, sigma)
        self.lib = _lib.load()
        c = _lib.SrlxEngine()
        env.fill(c)
        c.n_envs, c.seed, c.net = 1, int(e.seed) & 0xFFFFFFFFFFFFFFFF, self.spec.to_c()
        for k, v in self.t.items():
            setattr(c, k, v.data_ptr())
        self.c = c
        self._pred_calls = 0
        self._x = torch.zeros((1, env.obs_dim), dtype=torch.float32).pin_memory()
        self._q = torch.zeros((1, env.n_actions), dtype=torch.float32).pin_memory()

    def _set(self, mu, sigma):
        for name in ("params", "target"):
            self.t[name].copy_(torch.as_tensor(np.asarray(mu, dtype=np.float32)))
        if self.ecfg.noisy:
            for name in ("params_sigma", "target_sigma"):
                self.t[name].copy_(torch.as_tensor(np.asarray(sigma, dtype=np.float32)))

    def call_restore(self, data: Any, from_serialized: bool = False, from_worker: bool = False, **kwargs) -> None:
        mu, sigma = checkpoint.parameter_restore(self.spec, data)
        self._set(mu, sigma)  # model_torch.py:47-49: restore loads q_online and q_target

    def call_backup(self, serialized: bool = False, to_worker: bool = False, **kwargs) -> Any:
        mu = self.t["params"].cpu().numpy()
        sigma = self.t["params_sigma"].cpu().numpy() if self.ecfg.noisy else None
        return checkpoint.parameter_backup(self.spec, mu, sigma)

    def summary(self, **kwargs):
        print(f"DeviceParameter: {self.spec.n_layers} dense layers {self.spec.k_dim[0]} -> {self.spec.out_dim}, "
              f"{self.spec.n_params} parameters (x2 with sigma)" if self.ecfg.noisy else
              f"DeviceParameter: {self.spec.n_layers} dense layers {self.spec.k_dim[0]} -> {self.spec.out_dim}, {self.spec.n_params} parameters")

    # ---- inference seam (dqn/model_torch.py:58-70) ----------------------------------------------------------------------
    def _forward(self, state, use_target: int) -> np.ndarray:
        x = np.ascontiguousarray(state, dtype=np.float32).reshape(-1, self.spec.in_dim)
        n = x.shape[0]
        if n == 1:
            self._x.copy_(torch.from_numpy(x))
            xd = self._x.to(self.device, non_blocking=True)
        else:
            xd = torch.from_numpy(x).to(self.device)
        q = torch.empty((n, self.spec.n_actions), dtype=torch.float32, device=self.device)
        self._pred_calls += 1  # NoisyLinear draws fresh noise on every forward call, evaluation included (noisy_linear.py:35-52)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_qnet_forward(self.c, int(use_target), xd.data_ptr(), n, self._pred_calls, q.data_ptr(),
                                                  torch.cuda.current_stream(self.device).cuda_stream))
        return q.cpu().numpy()

    def pred_q(self, state) -> np.ndarray:
        return self._forward(state, 0)

    def pred_target_q(self, state) -> np.ndarray:
        return self._forward(state, 1)


# ---------------------------------------------------------------------------------------------------------------------
class DeviceTrainer(RLTrainer):
    """RLTrainer (srl/base/rl/trainer.py:14-62).  One train() call = `updates_per_train` reference Trainer.train() steps
    (dqn/model_torch.py:90-132, rainbow/model_torch.py:85-122) inside one srlx_learn launch; set `rl_config.b200_updates_per_train = n`
    on the user's config to batch (default 1 = the reference's cadence)."""

    def on_setup(self) -> None:
        mem, par = self.memory, self.parameter
        if mem._bound_parameter is not par:
            mem.engine.adopt_tensors(par.t)  # the engine now trains the tensors pred_q reads
            mem._bound_parameter = par
        self.engine = mem.engine
        self.updates_per_train = max(1, int(getattr(self.config, "b200_updates_per_train", 1)))
        self.sync_count = 0

    def train(self) -> None:
        mem = self.memory
        mem.flush()
        if mem.length() < mem.ecfg.warmup_size:
            return
        n = self.updates_per_train
        self.engine.learn(n)
        st = self.engine.read_state()  # one 128-byte read: loss / sync for the reference's progress display
        mem.update([], np.zeros(0, np.float32), self.train_count)
        self.train_count = int(st.train_count)
        self.sync_count = int(st.sync_count)
        self.info["loss"] = float(st.last_loss)
        self.info["sync"] = self.sync_count


# ---------------------------------------------------------------------------------------------------------------------
class DeviceWorker(RLWorker):
    """RLWorker (srl/base/rl/worker.py:48-69): the reference's policy (dqn.py:192-211, rainbow.py:301-331: epsilon-greedy with
    python's `random`, or the noisy argmax) over the device Q; on_step stores ONE record per env step (dqn.py:213-246; for n-step
    Rainbow the window + padded tail of rainbow.py:333-400 are rebuilt by index at sample time)."""

    def on_setup(self, worker, context) -> None:
        self.epsilon_sch = self.config.epsilon_scheduler.create(self.config.epsilon)
        self.noisy = bool(getattr(self.config, "enable_noisy_dense", False))
        self.n_actions = int(self.config.action_space.n)

    def policy(self, worker) -> int:
        invalid_actions = worker.invalid_actions
        if not self.noisy:
            epsilon = self.epsilon_sch.update(self.step_in_training).to_float() if self.training else self.config.test_epsilon
            self.info["epsilon"] = epsilon
            if random.random() < epsilon:
                return random.choice([a for a in range(self.n_actions) if a not in invalid_actions])
        q = self.parameter.pred_q(worker.state[np.newaxis, ...])[0]
        q[invalid_actions] = -np.inf  # dqn.py:207, rainbow.py:307,325
        return int(np.argmax(q))

    def on_step(self, worker) -> None:
        if not self.training:
            return
        reward = worker.reward
        if self.config.enable_reward_clip:
            reward = -1 if reward < 0 else (1 if reward > 0 else 0)
        self.memory.add((worker.state, worker.next_state, int(worker.action), float(reward), bool(worker.terminated), bool(worker.done),
                         list(worker.next_invalid_actions)))

    def render_terminal(self, worker, **kwargs) -> None:
        q = self.parameter.pred_q(worker.state[np.newaxis, ...])[0]
        maxa = int(np.argmax(q))
        if self.config.enable_rescale:
            from srl.rl.functions import inverse_rescaling

            q = inverse_rescaling(q)
        worker.print_discrete_action_info(maxa, lambda a: f"{q[a]:7.5f}")


# ---------------------------------------------------------------------------------------------------------------------
class DeviceEnv(EnvBase):
    """EnvBase (srl/base/env/base.py:18-206) over the device's closed-form envs: reset() / step(action) are one tiny launch each
    (srlx_env_reset_obs / srlx_env_step_actions) on a one-copy engine.  Spaces follow the reference's own classes: Grid =
    ArrayDiscreteSpace(2, 0, [W-1, H-1]) + Discrete(4) (srl/envs/grid.py:151-153,667-673), CartPole-v1 = Box(4,) float32 +
    Discrete(2) (gymnasium_wrapper.py space mapping, tests/quick/base/env/test_gymnasium_wrapper.py:30-43)."""

    def __init__(self, name: str = "Grid", seed: int = 0, device: str = "cuda:0", **env_kwargs):
        super().__init__()
        if not torch.cuda.is_available():
            raise _lib.SrlxError("DeviceEnv needs a CUDA device (no CPU fallback)")
        self.name_ = name
        cfg = EngineConfig(env=name, n_envs=1, ring_rows=1, multisteps=1, batch_size=1, mem_kind=_lib.MEM_UNIFORM, hidden=(4,), seed=int(seed),
                           warmup_size=1, env_kwargs=dict(env_kwargs))
        self.engine = DeviceEngine(cfg, device=device)
        self.spec = self.engine.env
        dev = self.engine.device
        self._obs = torch.zeros((1, self.spec.obs_dim), dtype=torch.float32, device=dev)
        self._act = torch.zeros(1, dtype=torch.int32, device=dev)
        self._rew = torch.zeros(1, dtype=torch.float64, device=dev)
        self._term = torch.zeros(1, dtype=torch.uint8, device=dev)
        self._trunc = torch.zeros(1, dtype=torch.uint8, device=dev)

    @property
    def action_space(self):
        return DiscreteSpace(self.spec.n_actions)

    @property
    def observation_space(self):
        if self.spec.env_id == _lib.ENV_GRID:
            return ArrayDiscreteSpace(2, low=0, high=[int(h) for h in self.spec.obs_high])
        low = np.asarray([max(v, -np.finfo(np.float32).max) for v in self.spec.obs_low], dtype=np.float32)
        high = np.asarray([min(v, np.finfo(np.float32).max) for v in self.spec.obs_high], dtype=np.float32)
        return BoxSpace((self.spec.obs_dim,), low, high, np.float32)

    @property
    def max_episode_steps(self) -> int:
        return int(self.spec.max_episode_steps)

    @property
    def player_num(self) -> int:
        return 1

    @property
    def reward_baseline(self):
        return self.spec.reward_baseline

    def _state(self):
        o = self._obs.cpu().numpy()[0]
        return [int(v) for v in o] if self.spec.env_id == _lib.ENV_GRID else o.astype(np.float32)

    def reset(self, *, seed: Optional[int] = None, **kwargs):
        eng = self.engine
        with torch.cuda.device(eng.device):
            _lib.check(eng.lib.srlx_env_reset_obs(eng.c, 1, self._obs.data_ptr(), eng._stream()))
        return self._state()

    def step(self, action):
        eng = self.engine
        self._act.fill_(int(action))
        with torch.cuda.device(eng.device):
            _lib.check(eng.lib.srlx_env_step_actions(eng.c, self._act.data_ptr(), self._obs.data_ptr(), self._rew.data_ptr(),
                                                     self._term.data_ptr(), self._trunc.data_ptr(), eng._stream()))
        # EnvRun applies max_episode_steps itself (env_run.py:360-366): only the env's own termination is reported here
        return self._state(), float(self._rew.item()), bool(self._term.item()), False

    def backup(self) -> Any:
        t = self.engine.t
        return [t["env_state"].cpu().numpy().copy(), int(t["env_step_num"].item()), int(t["env_episode"].item()), self.engine.read_state().vec_steps]

    def restore(self, data: Any) -> None:
        t = self.engine.t
        t["env_state"].copy_(torch.as_tensor(data[0]))
        t["env_step_num"].fill_(int(data[1]))
        t["env_episode"].fill_(int(data[2]))
        st = self.engine.read_state()
        st.vec_steps = int(data[3])
        self.engine.write_state(st)


# ---------------------------------------------------------------------------------------------------------------------
_saved_registry = {}
_KEYS = ("DQN:torch", "Rainbow:torch", "Rainbow_no_multisteps:torch")


def register(envs: bool = True) -> None:
    """Take over the reference's registry keys for DQN / Rainbow (torch framework) and register the device-backed envs."""
    from srl.algorithms import dqn, rainbow  # noqa: F401  (their import registers the reference's own classes first)

    reg = rl_registration._registry
    for k in _KEYS:
        if k in reg and k not in _saved_registry:
            _saved_registry[k] = list(reg[k])
    eps = (f"{_MOD}:DeviceMemory", f"{_MOD}:DeviceParameter", f"{_MOD}:DeviceTrainer", f"{_MOD}:DeviceWorker")
    for cfg in (dqn.Config().set_torch(), rainbow.Config(multisteps=3).set_torch(), rainbow.Config(multisteps=1).set_torch()):
        key = rl_registration._create_registry_key(cfg)
        reg[key] = list(eps)  # what register(..., check_duplicate=False) does, without its overwrite warning
    if envs:
        env_registration.register(id="Grid-b200", entry_point=f"{_MOD}:DeviceEnv", kwargs=dict(name="Grid"), check_duplicate=False)
        env_registration.register(id="EasyGrid-b200", entry_point=f"{_MOD}:DeviceEnv", kwargs=dict(name="EasyGrid"), check_duplicate=False)
        env_registration.register(id="CartPole-v1", entry_point=f"{_MOD}:DeviceEnv", kwargs=dict(name="CartPole-v1"), check_duplicate=False)


def unregister() -> None:
    reg = rl_registration._registry
    for k, v in _saved_registry.items():
        reg[k] = list(v)
    _saved_registry.clear()


def train_vectorized(runner, num_envs: int = 4096, seed: int = 0, device: str = "cuda:0", ring_rows: Optional[int] = None, **train_kwargs):
    """The vectorised loop for an existing `srl.Runner`: trains `runner.rl_config` on `runner.env_config` with `num_envs` device env
    copies (DeviceRunner), starting from the runner's current parameters, and hands the trained parameters back through
    RLParameter.restore -- so `runner.evaluate()`, `runner.save_parameter()`, ... continue from them.  Returns the VecRunState."""
    from .srl_plugin import DeviceRunner

    env_name = runner.env_config.name if hasattr(runner, "env_config") else str(runner.env_config)
    if env_name.endswith("-b200"):
        env_name = env_name[: -len("-b200")]
    dev = DeviceRunner(env_name, runner.rl_config, num_envs=num_envs, seed=seed, device=device, ring_rows=ring_rows)
    par = runner.make_parameter()
    dev.load_state_dict(par.backup())
    state = dev.train(**train_kwargs)
    par.restore(dev.state_dict())
    return state
