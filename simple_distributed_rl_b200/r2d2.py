"""R2D2 on the device (BASELINE configs[3]): the host side of csrc/r2d2.cu.

`R2D2Config` carries the reference's own field names and defaults (srl/algorithms/r2d2/config.py:45-128; the reference's R2D2 classes
are TensorFlow-only, so they cannot be imported where TensorFlow is absent -- the fields are restated, not read from a config object).
`R2D2Engine` owns the HBM-resident buffers and calls the C ABI; `R2D2Runner` is the Runner-shaped loop:

    one vector step of E env copies                          srlx_r2d2_vec_step   Worker.on_reset / policy / on_step (r2d2.py:221-318)
    n trainer updates                                        srlx_r2d2_learn(n)   Trainer.train / _train_on_batches  (:90-215)
    Q of one step from a given LSTM state                    srlx_r2d2_forward    QNetwork.call                      (:27-63)

Replay: the reference stores one item per step, each a copy of the worker's recent_* lists (burnin + seq_len + 1 states, seq_len
actions / probabilities / rewards / dones, the LSTM state in front of the first state).  Here every env copy owns a ring column with one
ROW per step; an item is an anchor row and its lists are rebuilt from the rows in front of it (csrc/r2d2.cu: r2d2_gather_kernel), the
padding in front of an episode's first step by index.  Capacity: `memory.capacity` items -> ceil(capacity / E) rows per column plus
the burnin + seq_len - 1 rows an anchor needs in front of it.  BASELINE configs[3] names LunarLander-v2, a Box2D simulation with no
closed form: the device runs R2D2 on the closed-form envs (CartPole-v1, Pendulum-v1 -- the reference's own R2D2 acceptance env,
tests/algorithms_/base_r2d2.py:40-44 --, Grid).  No invalid-action masks (these envs have none).  No CPU fallback.
"""
import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .envspec import make_env_spec

DUELING = {None: _lib.DUEL_NONE, "average": _lib.DUEL_AVERAGE, "max": _lib.DUEL_MAX, "": _lib.DUEL_NAIVE}


@dataclass
class R2D2Config:
    env: str = "Pendulum-v1"
    n_envs: int = 256
    # ---- the reference's fields (srl/algorithms/r2d2/config.py:45-87), same names and defaults
    test_epsilon: float = 0.0
    epsilon: float = 0.1
    batch_size: int = 32
    capacity: int = 100_000            # memory.capacity
    warmup_size: int = 1_000           # memory.warmup_size
    memory: str = "ReplayBuffer"       # memory.name: "ReplayBuffer" | "Proportional" (memory.set_proportional)
    per_alpha: float = 0.6
    per_beta_initial: float = 0.4
    per_beta_steps: int = 1_000_000
    per_has_duplicate: bool = True
    per_epsilon: float = 0.0001
    lstm_units: int = 512
    hidden_layers: Tuple[int, ...] = (512,)   # hidden_block.set(layer_sizes) / set_dueling_network(layer_sizes)
    dueling_type: Optional[str] = "average"   # None: hidden_block.set (plain MLP + Dense(A)); "average" | "max" | "": dueling block
    burnin: int = 5
    sequence_length: int = 5
    discount: float = 0.997
    lr: float = 0.001
    target_model_update_interval: int = 1000
    enable_double_dqn: bool = True
    enable_rescale: bool = False
    enable_retrace: bool = True
    retrace_h: float = 1.0
    reward_shift: float = 0.0
    reward_scale: float = 1.0
    seed: int = 0
    env_kwargs: dict = field(default_factory=dict)

    def set_atari_config(self):
        """config.py:97-118"""
        self.lstm_units, self.hidden_layers, self.dueling_type = 512, (512,), "average"
        self.burnin, self.sequence_length = 40, 80
        self.discount, self.lr, self.batch_size, self.target_model_update_interval = 0.997, 0.0001, 64, 2500
        self.enable_double_dqn, self.enable_rescale, self.enable_retrace = True, True, False
        self.capacity, self.memory = 1_000_000, "Proportional"
        self.per_alpha, self.per_beta_initial, self.per_beta_steps = 0.9, 0.6, 1_000_000
        return self


class R2D2NetSpec:
    """QNetwork (r2d2.py:27-63) as one flat fp32 buffer in the layout include/srlx.h documents: LSTM rows unit * 4 + gate over
    [x | h | bias]; every head layer [out][k + 1] with the bias as last column; a dueling block as one hidden layer of width 2H and one
    output layer of 1 + A rows with structurally-zero off-branch blocks."""

    def __init__(self, in_dim: int, lstm_units: int, hidden: Tuple[int, ...], dueling: Optional[str], n_actions: int):
        self.D, self.u, self.A = int(in_dim), int(lstm_units), int(n_actions)
        self.hidden, self.dueling = tuple(int(h) for h in hidden), dueling
        self.K = self.D + self.u + 1
        off = 4 * self.u * self.K
        self.lstm_off = 0
        self.head: List[Tuple[int, int, int]] = []  # (out, k, off)
        k = self.u
        if dueling is None:
            sizes = list(self.hidden) + [self.A]
        else:
            assert len(self.hidden) >= 1, "set_dueling_network needs at least one layer size"
            self.H = self.hidden[-1]
            sizes = list(self.hidden[:-1]) + [2 * self.H, 1 + self.A]
        for out in sizes:
            self.head.append((out, k, off))
            off += out * (k + 1)
            k = out
        self.n_params = off
        if len(self.head) > _lib.SRLX_MAX_LAYERS:
            raise ValueError("too many head layers")
        if dueling is not None:
            out, k, o = self.head[-1]  # k = 2H; row 0 reads [0, H), rows 1.. read [H, 2H); the bias column is real
            H = self.H
            self.zero_mask = np.zeros((out, k + 1), dtype=bool)
            self.zero_mask[0, H:2 * H] = True
            self.zero_mask[1:, 0:H] = True
        else:
            self.zero_mask = None

    # ---- keras get_weights() order: LSTM kernel [D][4u] (gate-major columns i, f, c, o), recurrent kernel [u][4u], bias [4u]; every
    # Dense kernel [k][out], bias [out]; dueling block: v hidden, v out, adv hidden, adv out (tf/blocks/dueling_network.py:37-62)
    def from_keras(self, weights: List[np.ndarray]) -> np.ndarray:
        p = np.zeros(self.n_params, dtype=np.float32)
        D, u, K = self.D, self.u, self.K
        kern, rec, bias = [np.asarray(w, np.float32) for w in weights[:3]]
        Wl = p[:4 * u * K].reshape(u, 4, K)
        Wl[:, :, :D] = kern.reshape(D, 4, u).transpose(2, 1, 0)
        Wl[:, :, D:D + u] = rec.reshape(u, 4, u).transpose(2, 1, 0)
        Wl[:, :, K - 1] = bias.reshape(4, u).T
        i = 3
        n_plain = len(self.head) if self.dueling is None else len(self.head) - 2
        for l in range(n_plain):
            out, k, off = self.head[l]
            Wd = p[off:off + out * (k + 1)].reshape(out, k + 1)
            Wd[:, :k] = np.asarray(weights[i], np.float32).T
            Wd[:, k] = np.asarray(weights[i + 1], np.float32)
            i += 2
        if self.dueling is not None:
            H, A = self.H, self.A
            vhk, vhb, vok, vob, ahk, ahb, aok, aob = [np.asarray(w, np.float32) for w in weights[i:i + 8]]
            out, k, off = self.head[-2]
            Wh = p[off:off + out * (k + 1)].reshape(out, k + 1)
            Wh[:H, :k], Wh[:H, k], Wh[H:, :k], Wh[H:, k] = vhk.T, vhb, ahk.T, ahb
            out, k, off = self.head[-1]
            Wo = p[off:off + out * (k + 1)].reshape(out, k + 1)
            Wo[0, :H], Wo[0, k] = vok[:, 0], vob[0]
            Wo[1:, H:2 * H], Wo[1:, k] = aok.T, aob
        return p

    def to_keras(self, p: np.ndarray) -> List[np.ndarray]:
        p = np.asarray(p, np.float32)
        D, u, K = self.D, self.u, self.K
        Wl = p[:4 * u * K].reshape(u, 4, K)
        out_w = [Wl[:, :, :D].transpose(2, 1, 0).reshape(D, 4 * u).copy(), Wl[:, :, D:D + u].transpose(2, 1, 0).reshape(u, 4 * u).copy(),
                 Wl[:, :, K - 1].T.reshape(4 * u).copy()]
        n_plain = len(self.head) if self.dueling is None else len(self.head) - 2
        for l in range(n_plain):
            out, k, off = self.head[l]
            Wd = p[off:off + out * (k + 1)].reshape(out, k + 1)
            out_w += [Wd[:, :k].T.copy(), Wd[:, k].copy()]
        if self.dueling is not None:
            H = self.H
            out, k, off = self.head[-2]
            Wh = p[off:off + out * (k + 1)].reshape(out, k + 1)
            out2, k2, off2 = self.head[-1]
            Wo = p[off2:off2 + out2 * (k2 + 1)].reshape(out2, k2 + 1)
            out_w += [Wh[:H, :k].T.copy(), Wh[:H, k].copy(), Wo[0:1, :H].T.copy(), Wo[0:1, k2].copy(),
                      Wh[H:, :k].T.copy(), Wh[H:, k].copy(), Wo[1:, H:2 * H].T.copy(), Wo[1:, k2].copy()]
        return out_w

    def init_keras(self, seed: int) -> List[np.ndarray]:
        """The reference's initialisers: LSTM glorot_uniform kernel, orthogonal recurrent kernel, zero bias with unit forget bias (keras
        LSTM defaults); he_normal + zero bias for the MLP and the dueling hidden layers (mlp_block.py:17-18, dueling_network.py:40-58);
        truncated_normal(stddev 0.05) for the output layers."""
        g = torch.Generator().manual_seed(int(seed))
        D, u, A = self.D, self.u, self.A
        lim = math.sqrt(6.0 / (D + 4 * u))
        w = [((torch.rand(D, 4 * u, generator=g) * 2 - 1) * lim).numpy()]
        rec = torch.empty(u, 4 * u)
        torch.nn.init.orthogonal_(rec, generator=g)
        w.append(rec.numpy())
        b = np.zeros(4 * u, np.float32)
        b[u:2 * u] = 1.0
        w.append(b)

        def he(k, out):
            return (torch.nn.init.trunc_normal_(torch.empty(k, out), std=1.0, a=-2.0, b=2.0, generator=g) * (math.sqrt(2.0 / k) / 0.87962566103423978)).numpy()

        def tn(k, out):
            return torch.nn.init.trunc_normal_(torch.empty(k, out), std=0.05, a=-0.1, b=0.1, generator=g).numpy()

        k = u
        if self.dueling is None:
            for h in self.hidden:
                w += [he(k, h), np.zeros(h, np.float32)]
                k = h
            w += [tn(k, A), np.zeros(A, np.float32)]
        else:
            for h in self.hidden[:-1]:
                w += [he(k, h), np.zeros(h, np.float32)]
                k = h
            H = self.H
            w += [he(k, H), np.zeros(H, np.float32), tn(H, 1), np.zeros(1, np.float32),
                  he(k, H), np.zeros(H, np.float32), tn(H, A), np.zeros(A, np.float32)]
        return [np.asarray(x, np.float32) for x in w]


class R2D2Engine:
    def __init__(self, cfg: R2D2Config, device="cuda:0", debug: bool = False, weights=None, track_episodes: bool = False,
                 training: bool = True, persistent: bool = True):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.SrlxError("R2D2Engine needs a CUDA device (no CPU fallback)")
        if cfg.memory not in ("ReplayBuffer", "Proportional"):
            raise NotImplementedError(f"memory {cfg.memory!r}: the device R2D2 replay is 'ReplayBuffer' or 'Proportional'")
        if cfg.dueling_type not in DUELING:
            raise ValueError(f"dueling_type {cfg.dueling_type!r}")
        self.cfg, self.device = cfg, torch.device(device)
        self.env = make_env_spec(cfg.env, **cfg.env_kwargs)
        E, D, A, u, B = cfg.n_envs, self.env.obs_dim, self.env.n_actions, cfg.lstm_units, cfg.batch_size
        S, W = cfg.sequence_length, cfg.burnin + cfg.sequence_length
        self.E, self.D, self.A, self.u, self.B, self.S, self.W = E, D, A, u, B, S, W
        self.spec = R2D2NetSpec(D, u, cfg.hidden_layers, cfg.dueling_type, A)
        self.K = K = self.spec.K
        self.per = cfg.memory == "Proportional"
        self.persistent = bool(persistent)
        self.R = R = max(2 * W, -(-cfg.capacity // E) + W - 1) if training else 2 * W
        if training and cfg.warmup_size > E * (R - (W - 1)):
            raise ValueError(f"warmup_size {cfg.warmup_size} exceeds the reachable memory size {E * (R - (W - 1))}")
        if training and cfg.warmup_size < B:
            raise ValueError("warmup_size must be >= batch_size")
        dev, P = self.device, self.spec.n_params
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731

        def ones_last(shape):  # activations carry a trailing 1 (the bias column's partner)
            t = z(shape, torch.float32)
            t[..., -1] = 1.0
            return t

        self.t = dict(
            state=z(C.sizeof(_lib.SrlxState), torch.uint8),
            env_state=z((E, 4), torch.float64), env_step_num=z(E, torch.int32), env_episode=z(E, torch.int32),
            env_ep_reward=z(E, torch.float64), env_needs_reset=torch.ones(E, dtype=torch.uint8, device=dev),
            params=z(P, torch.float32), target=z(P, torch.float32),
            roll_xh=ones_last((E, K)), roll_h=ones_last((E, u + 1)), roll_c=z((E, u), torch.float32), roll_reset=z(E, torch.uint8),
            roll_gates=z((E, 4 * u) if E > 64 else (1,), torch.float32),
        )
        head = self.spec.head
        for l, (out, k, off) in enumerate(head):
            last = l == len(head) - 1
            self.t[f"roll_act{l}"] = z((E, out), torch.float32) if last else ones_last((E, out + 1))
        if track_episodes:
            self.t["env_first_ep_reward"] = z(E, torch.float64)
            self.t["env_last_ep_len"] = z(E, torch.int32)
        if debug:
            self.t["dbg_q"] = z((E, A), torch.float32)
            self.t["dbg_action"] = z(E, torch.int32)
        if training:
            N = R * E
            self.t.update(
                adam_m=z(P, torch.float32), adam_v=z(P, torch.float32), grads=z(P, torch.float32),
                cursor=z(E, torch.int32), ring_obs=z((N, D), torch.float32), ring_next_obs=z((N, D), torch.float32),
                ring_action=z(N, torch.int32), ring_prob=z(N, torch.float64), ring_reward=z(N, torch.float64), ring_done=z(N, torch.uint8),
                ring_tstep=z(N, torch.int32), ring_h=z((N, u), torch.float32), ring_c=z((N, u), torch.float32),
                new_c0=z(E, torch.int32), new_n=z(E, torch.int32),
                xh=ones_last((2, W + 2, B, K)), cbuf=z((2, W + 2, B, u), torch.float32),
                gates=z((W + 1, B, 4 * u), torch.float32), dgates=z((W + 1, B, 4 * u), torch.float32), dc=z((B, u), torch.float32), bar=z(8, torch.int32),
                gemm_ws=z(32 * max(out * (k + 1) for out, k, _ in head if out <= 32 or k + 1 <= 32) if any(out <= 32 or k + 1 <= 32 for out, k, _ in head) else 1, torch.float32),
                dh=z(((S + 1) * B, u), torch.float32), q=z((2, B, S + 1, A), torch.float32),
                sel=z(B, torch.int64), weights=torch.ones(B, dtype=torch.float32, device=dev),
                b_actions=z((B, S), torch.int32), b_mu=torch.ones((B, S), dtype=torch.float64, device=dev), b_rewards=z((B, S), torch.float64),
                b_dones=z((B, S), torch.uint8), b_target=z((B, S), torch.float64), b_tdmean=z(B, torch.float64), b_tdkind=z(B, torch.uint8),
            )
            for l, (out, k, off) in enumerate(head):
                last = l == len(head) - 1
                self.t[f"act{l}"] = z((2, (S + 1) * B, out), torch.float32) if last else ones_last((2, (S + 1) * B, out + 1))
                self.t[f"dact{l}"] = z(((S + 1) * B, out), torch.float32)
            if self.per:
                self.t["tree"] = z(2 * N - 1, torch.float64)
                self.t["add_idx"] = z(E * (2 * S + W), torch.int64)
                self.t["add_pri"] = z(E * (2 * S + W), torch.float64)
        self.c = self._build_struct()
        self._dp_world, self._dp_group, self._dp_ready = 1, None, False
        self.set_weights(self.spec.init_keras(cfg.seed) if weights is None else weights)
        if training and self.per:
            self._write_state(max_priority=1.0)  # ProportionalMemory.max_priority starts at 1 (proportional_memory.py:116)

    def _build_struct(self) -> "_lib.SrlxR2d2":
        cfg, c = self.cfg, _lib.SrlxR2d2()
        self.env.fill(c.env)
        e = c.env
        e.n_envs, e.seed, e.ring_rows, e.batch_size = self.E, int(cfg.seed) & 0xFFFFFFFFFFFFFFFF, self.R, self.B
        e.mem_kind = _lib.MEM_PROPORTIONAL if self.per else _lib.MEM_UNIFORM
        e.has_duplicate, e.warmup_size = int(cfg.per_has_duplicate), int(cfg.warmup_size)
        e.target_update_interval = int(cfg.target_model_update_interval)
        e.enable_double_dqn, e.enable_rescale = int(cfg.enable_double_dqn), int(cfg.enable_rescale)
        e.epsilon, e.discount, e.lr, e.retrace_h = float(cfg.epsilon), float(cfg.discount), float(cfg.lr), float(cfg.retrace_h)
        e.adam_beta1, e.adam_beta2, e.adam_eps = 0.9, 0.999, 1e-7  # keras.optimizers.Adam defaults
        e.per_alpha, e.per_beta_initial, e.per_beta_steps, e.per_epsilon = (float(cfg.per_alpha), float(cfg.per_beta_initial),
                                                                            float(cfg.per_beta_steps), float(cfg.per_epsilon))
        e.reward_shift, e.reward_scale = float(cfg.reward_shift), float(cfg.reward_scale)
        for k in ("state", "env_state", "env_step_num", "env_episode", "env_ep_reward", "env_needs_reset", "env_first_ep_reward",
                  "env_last_ep_len", "tree", "dbg_q", "dbg_action"):
            if k in self.t:
                setattr(e, k, self.t[k].data_ptr())
        c.lstm_units, c.burnin, c.seq_len, c.enable_retrace = self.u, cfg.burnin, self.S, int(cfg.enable_retrace)
        c.n_head, c.dueling = len(self.spec.head), DUELING[cfg.dueling_type]
        for l, (out, k, off) in enumerate(self.spec.head):
            c.head_out[l], c.head_k[l], c.head_off[l] = out, k, off
            for nm in ("roll_act", "act", "dact"):
                if f"{nm}{l}" in self.t:
                    getattr(c, nm)[l] = self.t[f"{nm}{l}"].data_ptr()
        c.lstm_off, c.n_params, c.test_epsilon = self.spec.lstm_off, self.spec.n_params, float(cfg.test_epsilon)
        c.duel_hidden = self.spec.H if self.spec.zero_mask is not None else 0
        c.no_persistent = int(not self.persistent)
        if "gemm_ws" in self.t:
            c.gemm_ws_floats = self.t["gemm_ws"].numel()
        for k in ("params", "target", "adam_m", "adam_v", "grads", "cursor", "ring_obs", "ring_next_obs", "ring_action", "ring_prob",
                  "ring_reward", "ring_done", "ring_tstep", "ring_h", "ring_c", "roll_xh", "roll_h", "roll_c", "roll_gates", "roll_reset", "new_c0", "new_n",
                  "add_idx", "add_pri", "xh", "cbuf", "gates", "dgates", "dc", "gemm_ws", "bar", "dh", "q", "sel", "weights", "b_actions", "b_mu", "b_rewards",
                  "b_dones", "b_target", "b_tdmean", "b_tdkind"):
            if k in self.t:
                setattr(c, k, self.t[k].data_ptr())
        return c

    # ---- calls
    def _s(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def vec_step(self, training=True):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_r2d2_vec_step(C.byref(self.c), int(training), self._s()))

    def learn(self, n_updates: int = 1):
        with torch.cuda.device(self.device):
            if self._dp_world <= 1:
                _lib.check(self.lib.srlx_r2d2_learn(C.byref(self.c), int(n_updates), self._s()))
                return
            import torch.distributed as dist

            if not self._dp_ready:  # the warm-up gate is per shard on the device: open it for all ranks in the same update
                ok = torch.tensor([int(self.read_state().mem_size >= self.cfg.warmup_size)], device=self.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self._dp_group)
                if not int(ok.item()):
                    return
                self._dp_ready = True
            for _ in range(int(n_updates)):  # ONE trainer over all shards: the gradient of the global batch, the same Adam step everywhere
                _lib.check(self.lib.srlx_r2d2_learn_phase(C.byref(self.c), 1, 1, self._s()))
                dist.all_reduce(self.t["grads"], op=dist.ReduceOp.SUM, group=self._dp_group)
                self.t["grads"].mul_(1.0 / self._dp_world)
                _lib.check(self.lib.srlx_r2d2_learn_phase(C.byref(self.c), 1, 2, self._s()))

    def link_data_parallel(self, group=None):
        """Actor shards + ONE learner (BASELINE configs[3]; the reference's distributed mode has one trainer fed by all actors,
        srl/base/run/play_mp.py:352-462): every rank keeps its env copies and its replay shard, samples batch_size sequences from it,
        and every update applies the gradient of the GLOBAL batch (world x batch_size sequences: keras' mean over all of them) --
        an NCCL all-reduce of the flat gradient over NVLink between the backward pass and Adam (6 MB at LSTM 512: tens of
        microseconds against milliseconds of update).  Parameters, target network and Adam state start from rank 0's and stay
        bit-identical.  IS weights use the shard's own N / total / max.  All ranks must pass the warm-up gate together (same
        warmup_size and env count per rank)."""
        import torch.distributed as dist

        world = dist.get_world_size(group)
        if world > 1:
            for k in ("params", "target", "adam_m", "adam_v"):
                dist.broadcast(self.t[k], src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self._dp_world, self._dp_group, self._dp_ready = world, group, False

    def forward(self, obs, h, c, use_target=False):
        """(q [n][A], h' [n][u], c' [n][u]) of one step (n <= batch_size; runs in the learner workspace)."""
        f = lambda x, w: torch.as_tensor(x, dtype=torch.float32).to(self.device).reshape(-1, w).contiguous()  # noqa: E731
        obs, h, c = f(obs, self.D), f(h, self.u), f(c, self.u)
        n = obs.shape[0]
        q, h2, c2 = (torch.empty((n, w), dtype=torch.float32, device=self.device) for w in (self.A, self.u, self.u))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_r2d2_forward(C.byref(self.c), int(use_target), obs.data_ptr(), h.data_ptr(), c.data_ptr(), n,
                                                  q.data_ptr(), h2.data_ptr(), c2.data_ptr(), self._s()))
        return q, h2, c2

    # ---- state
    def read_state(self) -> "_lib.SrlxState":
        return _lib.SrlxState.from_buffer_copy(self.t["state"].cpu().numpy().tobytes())

    def _write_state(self, **kw):
        st = self.read_state()
        for k, v in kw.items():
            setattr(st, k, v)
        self.t["state"].copy_(torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8))

    def get_params(self) -> np.ndarray:
        return self.t["params"].cpu().numpy()

    def get_weights(self) -> List[np.ndarray]:
        """Parameter.call_backup (r2d2.py:75-76): q_online.get_weights() in keras order."""
        return self.spec.to_keras(self.get_params())

    def set_weights(self, weights, target_too: bool = True):
        """Parameter.call_restore (:71-73): both networks take the weights."""
        p = torch.as_tensor(self.spec.from_keras(weights) if isinstance(weights, (list, tuple)) else np.asarray(weights, np.float32))
        self.t["params"].copy_(p)
        if target_too:
            self.t["target"].copy_(p)


@dataclass
class R2D2RunState:
    total_step: int = 0
    train_count: int = 0
    episode_count: int = 0
    vec_steps: int = 0
    mean_episode_reward: float = float("nan")
    loss: float = 0.0
    sync: int = 0
    end_reason: str = ""


class R2D2Runner:
    """Runner.train / evaluate for R2D2 on the device (stop arguments as srl.Runner.train).  The reference's loop runs one
    Trainer.train() per env step (core_play.play, train_interval 1): one vector step of E copies is followed by E // train_interval
    updates (at least one)."""

    def __init__(self, cfg: R2D2Config, device="cuda:0", debug=False, weights=None):
        self.cfg = cfg
        self.engine = R2D2Engine(cfg, device=device, debug=debug, weights=weights)

    def train(self, max_steps=0, max_train_count=0, max_episodes=0, train_interval: int = 1, updates_per_vec_step: Optional[int] = None,
              callbacks=None) -> R2D2RunState:
        assert max_steps > 0 or max_train_count > 0 or max_episodes > 0, "Please specify 'max_episodes', 'max_steps' or 'max_train_count'."
        eng, st = self.engine, R2D2RunState()
        n_upd = max(1, eng.E // max(1, train_interval)) if updates_per_vec_step is None else int(updates_per_vec_step)
        s0 = eng.read_state()
        poll = 0
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
            eng.vec_step(True)
            n = n_upd if max_train_count <= 0 else min(n_upd, max(0, max_train_count - st.train_count))
            if n > 0:
                eng.learn(n)
            poll += 1
            st.vec_steps += 1
            st.total_step += eng.E
            # the counters live on the device: read them back when a stop condition needs them (every step while warming up or close
            # to the end, else every 8th)
            if max_train_count > 0 or max_episodes > 0 or poll % 8 == 0:
                s = eng.read_state()
                st.train_count = int(s.train_count - s0.train_count)
                st.episode_count = int(s.episode_count - s0.episode_count)
                st.loss, st.sync = s.last_loss, int(s.sync_count)
                if s.episode_count > s0.episode_count:
                    st.mean_episode_reward = float((s.episode_reward_sum - s0.episode_reward_sum) / (s.episode_count - s0.episode_count))
            for cb in callbacks or []:
                if getattr(cb, "on_step_end", None) and cb.on_step_end(context=None, state=st):
                    st.end_reason = "callback.on_step_end"
                    return st
        s = eng.read_state()
        st.train_count, st.episode_count = int(s.train_count - s0.train_count), int(s.episode_count - s0.episode_count)
        st.loss, st.sync = s.last_loss, int(s.sync_count)
        return st

    def evaluate(self, max_episodes=10, max_vec_steps=100_000) -> List[float]:
        """Runner.evaluate: fresh env copies, training=False (test_epsilon); the reward of the first episode each copy finishes."""
        cfg = R2D2Config(**{**self.cfg.__dict__, "n_envs": int(max_episodes), "seed": self.cfg.seed + 0x5EED})
        ev = R2D2Engine(cfg, device=self.engine.device, weights=self.engine.get_params(), track_episodes=True, training=False)
        for i in range(max_vec_steps):
            ev.vec_step(training=False)
            if i % 8 == 7 and bool((ev.t["env_last_ep_len"] > 0).all().item()):
                break
        return [float(r) for r in ev.t["env_first_ep_reward"].cpu().numpy()]
