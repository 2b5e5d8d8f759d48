"""PPO on the device (BASELINE configs[4]): the host side of csrc/ppo.cu.

`PPOConfig` carries the reference's own field names and defaults (srl/algorithms/ppo/config.py:43-128; the reference's PPO classes are
TensorFlow-only, so they cannot be imported where TensorFlow is absent -- the fields are restated, not read from a config object).
`PPOEngine` owns the HBM-resident buffers and calls the C ABI; `PPORunner` is the Runner-shaped loop:

    rollout of T vector steps (E env copies each)            srlx_ppo_vec_step  x T     Worker.policy + env.step   (ppo.py:307-356)
    V(s) with the current parameters, GAE / MC returns       srlx_ppo_finish_rollout    Worker.on_step at episode end (:357-406)
    train_num minibatch updates per warmup_size samples      srlx_ppo_learn(n)          Trainer.train / _train     (:198-291)
    (the memory is cleared: the next rollout overwrites it)                             Memory.clear               (:51-52)

The reference collects one env's finished episodes until `memory.warmup_size` samples are there, trains `train_num` minibatches of
`batch_size`, and clears.  With E env copies a rollout buffer holds N = (valid steps of) T x E samples at once and gets
`train_num * (N // warmup_size)` minibatch updates, each minibatch drawn from the whole buffer (ReplayBuffer.sample: distinct items).
Steps of episodes still running when the buffer ends are not emitted (the reference would emit them at their episode's end); with
T a multiple of the env's fixed episode length (Pendulum-v1: 200) nothing is lost.  No CPU fallback.
"""
import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .envspec import make_env_spec

BASELINES = {"": 0, "none": 0, "ave": 1, "std": 2, "normal": 3, "advantage": 4, "v": 4}


@dataclass
class PPOConfig:
    env: str = "Pendulum-v1"
    n_envs: int = 1024
    horizon: int = 200                 # rows of the rollout buffer
    # ---- the reference's fields (srl/algorithms/ppo/config.py:43-128), same names and defaults
    batch_size: int = 32
    warmup_size: int = 1000            # memory.warmup_size
    train_num: int = 50
    hidden_block: Tuple[int, ...] = (64, 64)
    value_block: Tuple[int, ...] = (64,)
    policy_block: Tuple[int, ...] = (64,)
    experience_collection_method: str = "GAE"
    discount: float = 0.9
    gae_discount: float = 0.9
    baseline_type: str = "advantage"
    surrogate_type: str = "clip"
    policy_clip_range: float = 0.2
    enable_value_clip: bool = True
    value_clip_range: float = 0.2
    lr: float = 0.0002
    lr_decay_steps: int = 2000         # lr_scheduler.set_step(2000, 0.01): keras ExponentialDecay(staircase=True); 0 = constant
    lr_decay_rate: float = 0.01
    value_loss_weight: float = 1.0
    entropy_weight: float = 0.01
    enable_state_normalized: bool = False
    global_gradient_clip_norm: float = 0.5
    reward_clip: Optional[Tuple[float, float]] = None
    stable_gradients_scale_range: Tuple[float, float] = (1e-10, 10)
    reward_shift: float = 0.0
    reward_scale: float = 1.0
    seed: int = 0
    env_kwargs: dict = field(default_factory=dict)


class PPONetSpec:
    """ActorCriticNetwork (ppo.py:55-101) as a flat fp32 parameter buffer: trunk layers, value block + value output (1), policy block +
    policy output (continuous: loc, log_scale; discrete: logits).  Two `srlx_net` views share the trunk offsets."""

    def __init__(self, in_dim: int, trunk, value, policy, n_out: int):
        self.in_dim, self.trunk, self.value, self.policy, self.n_out = int(in_dim), tuple(trunk), tuple(value), tuple(policy), int(n_out)
        off = 0
        self.layers = []  # (name, out, k, w_off, b_off)

        def add(name, out, k):
            nonlocal off
            w = off
            off += out * k
            b = off
            off += out
            self.layers.append((name, out, k, w, b))

        k = self.in_dim
        for i, h in enumerate(self.trunk):
            add(f"trunk{i}", h, k)
            k = h
        kt = k
        for i, h in enumerate(self.value):
            add(f"value{i}", h, k)
            k = h
        add("value_out", 1, k)
        k = kt
        for i, h in enumerate(self.policy):
            add(f"policy{i}", h, k)
            k = h
        add("policy_out", self.n_out, k)
        self.n_params = off
        nt = len(self.trunk)
        self.stack_v = self.layers[:nt] + [l for l in self.layers if l[0].startswith("value")]
        self.stack_p = self.layers[:nt] + [l for l in self.layers if l[0].startswith("policy")]
        if max(len(self.stack_v), len(self.stack_p)) > _lib.SRLX_MAX_LAYERS:
            raise ValueError("too many layers")

    def _net(self, stack, n_out) -> "_lib.SrlxNet":
        n = _lib.SrlxNet()
        n.n_layers, n.in_dim, n.n_params, n.n_actions, n.dueling, n.noisy = len(stack), self.in_dim, self.n_params, n_out, _lib.DUEL_NONE, 0
        for i, (_, out, k, w, b) in enumerate(stack):
            n.out_dim[i], n.k_dim[i], n.w_off[i], n.b_off[i], n.layer_noisy[i] = out, k, w, b, 0
        return n

    def nets(self):
        return self._net(self.stack_v, 1), self._net(self.stack_p, self.n_out)

    def init_params(self, seed: int, continuous: bool) -> np.ndarray:
        """The reference's initialisers: he_normal + zero bias for the MLP blocks (srl/rl/tf/blocks/mlp_block.py:17-18), orthogonal for
        the value output (ppo.py:62,72), glorot_uniform for loc / log_scale with a truncated-normal loc bias (normal_dist_block.py:108-
        131), zeros for the categorical logits (categorical_dist_block.py:147)."""
        g = torch.Generator().manual_seed(int(seed))
        p = np.zeros(self.n_params, dtype=np.float32)
        for name, out, k, w, b in self.layers:
            if name == "value_out":
                m = torch.empty(out, k)
                torch.nn.init.orthogonal_(m, generator=g)
            elif name == "policy_out":
                if continuous:
                    lim = math.sqrt(6.0 / (k + 1))  # glorot_uniform of each Dense(1)
                    m = (torch.rand(out, k, generator=g) * 2 - 1) * lim
                    p[b] = float(torch.nn.init.trunc_normal_(torch.empty(1), std=0.05, a=-0.1, b=0.1, generator=g))
                else:
                    m = torch.zeros(out, k)
            else:
                m = torch.randn(out, k, generator=g) * math.sqrt(2.0 / k)
            p[w:w + out * k] = m.numpy().reshape(-1)
        return p


class PPOEngine:
    def __init__(self, cfg: PPOConfig, device="cuda:0", debug: bool = False, params=None, track_episodes: bool = False):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.SrlxError("PPOEngine needs a CUDA device (no CPU fallback)")
        if cfg.surrogate_type not in ("clip", ""):
            raise NotImplementedError("surrogate_type 'kl' (adaptive KL penalty) is not on the device path")
        self.cfg, self.device = cfg, torch.device(device)
        self.env = make_env_spec(cfg.env, **cfg.env_kwargs)
        self.continuous = cfg.env == "Pendulum-v1"
        self.E, self.T, self.D, self.B = cfg.n_envs, cfg.horizon, self.env.obs_dim, cfg.batch_size
        n_out = 2 if self.continuous else self.env.n_actions
        self.spec = PPONetSpec(self.D, cfg.hidden_block, cfg.value_block, cfg.policy_block, n_out)
        dev, P, E, T, D = self.device, self.spec.n_params, self.E, self.T, self.D
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
        self.t = dict(
            state=z(C.sizeof(_lib.SrlxState), torch.uint8), pstate=z(C.sizeof(_lib.SrlxPpoState), torch.uint8),
            env_state=z((E, 4), torch.float64), env_step_num=z(E, torch.int32), env_episode=z(E, torch.int32),
            env_ep_reward=z(E, torch.float64), env_needs_reset=torch.ones(E, dtype=torch.uint8, device=dev),
            params=z(P, torch.float32), adam_m=z(P, torch.float32), adam_v=z(P, torch.float32), grad_scratch=z(P, torch.float32),
            buf_obs=z((T, E, D), torch.float32), buf_action=z((T, E), torch.float32), buf_v=z((T, E), torch.float32),
            buf_logp=z((T, E), torch.float32), buf_reward=z((T, E), torch.float32), buf_done=z((T, E), torch.uint8),
            buf_vnew=z((T + 1, E), torch.float32), buf_ret=z((T, E), torch.float32), buf_valid=z((T, E), torch.uint8),
        )
        if track_episodes:
            self.t["env_first_ep_reward"] = z(E, torch.float64)
            self.t["env_last_ep_len"] = z(E, torch.int32)
        if debug:
            self.t["dbg_idx"] = z(self.B, torch.int32)
            self.t["dbg_grads"] = z(P, torch.float32)
        self.c = self._build_struct()
        self.t["params"].copy_(torch.as_tensor(self.spec.init_params(cfg.seed, self.continuous) if params is None else np.asarray(params, np.float32)))

    def _build_struct(self) -> "_lib.SrlxPpo":
        cfg, c = self.cfg, _lib.SrlxPpo()
        self.env.fill(c.env)
        c.env.n_envs, c.env.seed = self.E, int(cfg.seed) & 0xFFFFFFFFFFFFFFFF
        c.env.reward_shift, c.env.reward_scale = float(cfg.reward_shift), float(cfg.reward_scale)
        for k in ("state", "env_state", "env_step_num", "env_episode", "env_ep_reward", "env_needs_reset", "env_first_ep_reward", "env_last_ep_len"):
            if k in self.t:
                setattr(c.env, k, self.t[k].data_ptr())
        c.net_v, c.net_p = self.spec.nets()
        c.n_params, c.continuous, c.horizon, c.batch_size = self.spec.n_params, int(self.continuous), self.T, self.B
        c.baseline_type, c.surrogate_clip = BASELINES[cfg.baseline_type], int(cfg.surrogate_type == "clip")
        c.enable_value_clip, c.state_normalized = int(cfg.enable_value_clip), int(cfg.enable_state_normalized)
        c.method = 0 if cfg.experience_collection_method == "GAE" else 1
        c.reward_clip_enable = int(cfg.reward_clip is not None)
        if cfg.reward_clip is not None:
            c.reward_clip_lo, c.reward_clip_hi = float(cfg.reward_clip[0]), float(cfg.reward_clip[1])
        c.lr_decay_steps, c.lr_decay_rate = int(cfg.lr_decay_steps), float(cfg.lr_decay_rate)
        for k in ("discount", "gae_discount", "policy_clip_range", "value_clip_range", "lr", "value_loss_weight", "entropy_weight"):
            setattr(c, k, float(getattr(cfg, k)))
        c.grad_clip_norm = float(cfg.global_gradient_clip_norm)
        c.adam_beta1, c.adam_beta2, c.adam_eps = 0.9, 0.999, 1e-7  # keras.optimizers.Adam defaults
        c.log_scale_lo, c.log_scale_hi = math.log(cfg.stable_gradients_scale_range[0]), math.log(cfg.stable_gradients_scale_range[1])
        c.action_low, c.action_high = (-2.0, 2.0) if self.continuous else (0.0, 0.0)  # Pendulum-v1's Box(-2, 2)
        for k in ("params", "adam_m", "adam_v", "buf_obs", "buf_action", "buf_v", "buf_logp", "buf_reward", "buf_done", "buf_vnew", "buf_ret",
                  "buf_valid", "pstate", "dbg_idx", "dbg_grads", "grad_scratch"):
            if k in self.t:
                setattr(c, k, self.t[k].data_ptr())
        return c

    def _s(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def vec_step(self, training=True):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_ppo_vec_step(C.byref(self.c), int(training), self._s()))

    def rollout(self, n_steps: Optional[int] = None):
        for _ in range(self.T if n_steps is None else n_steps):
            self.vec_step(True)

    def finish_rollout(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_ppo_finish_rollout(C.byref(self.c), self._s()))

    def learn(self, n_updates: int):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_ppo_learn(C.byref(self.c), int(n_updates), self._s()))

    def values(self, obs: torch.Tensor) -> torch.Tensor:
        x = obs.to(self.device, dtype=torch.float32).reshape(-1, self.D).contiguous()
        out = torch.empty(x.shape[0], dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_ppo_values(C.byref(self.c), x.data_ptr(), x.shape[0], out.data_ptr(), self._s()))
        return out

    def read_state(self) -> "_lib.SrlxState":
        return _lib.SrlxState.from_buffer_copy(self.t["state"].cpu().numpy().tobytes())

    def read_pstate(self) -> "_lib.SrlxPpoState":
        return _lib.SrlxPpoState.from_buffer_copy(self.t["pstate"].cpu().numpy().tobytes())

    def get_params(self) -> np.ndarray:
        return self.t["params"].cpu().numpy()


@dataclass
class PPORunState:
    total_step: int = 0
    train_count: int = 0
    episode_count: int = 0
    rollouts: int = 0
    mean_episode_reward: float = float("nan")
    policy_loss: float = 0.0
    value_loss: float = 0.0
    entropy_loss: float = 0.0
    end_reason: str = ""


class PPORunner:
    """Runner.train / evaluate for PPO on the device (stop arguments as srl.Runner.train: max_steps, max_train_count, max_episodes)."""

    def __init__(self, cfg: PPOConfig, device="cuda:0", debug=False, params=None):
        self.cfg = cfg
        self.engine = PPOEngine(cfg, device=device, debug=debug, params=params)

    def updates_per_rollout(self, n_valid: int) -> int:
        return self.cfg.train_num * max(1, n_valid // max(1, self.cfg.warmup_size))

    def train(self, max_steps=0, max_train_count=0, max_episodes=0, max_rollouts=0, callbacks=None) -> PPORunState:
        assert max_steps > 0 or max_train_count > 0 or max_episodes > 0 or max_rollouts > 0, \
            "Please specify 'max_episodes', 'max_steps', 'max_train_count' or 'max_rollouts'."
        eng, st = self.engine, PPORunState()
        s0, p0 = eng.read_state(), eng.read_pstate()
        prev_ep, prev_r = s0.episode_count, s0.episode_reward_sum
        while True:
            if max_steps > 0 and st.total_step >= max_steps:
                st.end_reason = "max_steps over."
                break
            if max_train_count > 0 and st.train_count >= max_train_count:
                st.end_reason = "max_train_count over."
                break
            if max_episodes > 0 and st.episode_count >= max_episodes:
                st.end_reason = "episode_count over."
                break
            if max_rollouts > 0 and st.rollouts >= max_rollouts:
                st.end_reason = "max_rollouts over."
                break
            eng.rollout()
            eng.finish_rollout()
            n_valid = int(eng.t["buf_valid"].sum().item())
            n_upd = self.updates_per_rollout(n_valid) if n_valid >= eng.B else 0
            if max_train_count > 0:
                n_upd = min(n_upd, max_train_count - st.train_count)
            if n_upd > 0:
                eng.learn(n_upd)
            s, p = eng.read_state(), eng.read_pstate()
            st.total_step = int(s.total_step - s0.total_step)
            st.train_count = int(p.train_count - p0.train_count)
            st.episode_count = int(s.episode_count - s0.episode_count)
            st.rollouts += 1
            if s.episode_count > prev_ep:
                st.mean_episode_reward = float((s.episode_reward_sum - prev_r) / (s.episode_count - prev_ep))
                prev_ep, prev_r = s.episode_count, s.episode_reward_sum
            st.policy_loss, st.value_loss, st.entropy_loss = p.policy_loss, p.value_loss, p.entropy_loss
            for cb in callbacks or []:
                if getattr(cb, "on_step_end", None) and cb.on_step_end(context=None, state=st):
                    st.end_reason = "callback.on_step_end"
                    return st
        return st

    def evaluate(self, max_episodes=10, max_vec_steps=100_000) -> List[float]:
        """Runner.evaluate: fresh env copies, training=False (continuous: the mean action; discrete: sampled, as the reference's worker
        does, ppo.py:319-336); the reward of the first episode each copy finishes."""
        cfg = PPOConfig(**{**self.cfg.__dict__, "n_envs": int(max_episodes), "horizon": 1, "seed": self.cfg.seed + 0x5EED})
        ev = PPOEngine(cfg, device=self.engine.device, params=self.engine.get_params(), track_episodes=True)
        for i in range(max_vec_steps):
            ev.vec_step(training=False)
            if i % 8 == 7 and bool((ev.t["env_last_ep_len"] > 0).all().item()):
                break
        return [float(r) for r in ev.t["env_first_ep_reward"].cpu().numpy()]
