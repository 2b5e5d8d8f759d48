"""DeviceEngine: owns the HBM-resident buffers (torch CUDA tensors) and drives the libsrlx.so kernels.

Data layout in HBM (one engine per GPU; E = n_envs, R = ring_rows, D = obs_dim, P = n_params):
  env_state      f64 [E,4]     environment state (CartPole x,x_dot,theta,theta_dot; Grid x,y)
  env_*          i32/u32/f64/u8 [E]  EnvRun counters (step_num, episodes started, episode reward, needs_reset)
  ring_obs       f32 [R*E, D]  state            slot(g, e) = (g % R) * E + e  (time-major: one vector step = one row,
  ring_next_obs  f32 [R*E, D]  next_state       so the E writes of a step are one contiguous, coalesced run)
  ring_action    i32 [R*E]     action index
  ring_reward    f32 [R*E]     reward after shift/scale/clip
  ring_term/done u8  [R*E]     worker.terminated / episode ended
  tree           f64 [2*R*E-1] SumTree in the reference's own flat layout (proportional only)
  params/target  f32 [P] (+ sigma [P] when noisy), adam_m/adam_v f32 [P*(1+noisy)]
  state          srlx_state    device-resident counters (RunState / trainer / memory scalars)
PyTorch is used for allocation, streams and host<->device copies only; every computation is a libsrlx kernel.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from . import _lib
from .envspec import make_env_spec
from .netspec import NetSpec


@dataclass
class EngineConfig:
    env: str = "CartPole-v1"
    n_envs: int = 8192
    ring_rows: int = 256
    multisteps: int = 1
    batch_size: int = 32
    mem_kind: int = _lib.MEM_PROPORTIONAL
    algo: str = "dqn"
    enable_double_dqn: bool = True
    enable_rescale: bool = False
    enable_reward_clip: bool = False
    has_duplicate: bool = True
    presample: bool = False  # batch t+1 drawn before update t lands (include/srlx.h srlx_engine.presample)
    target_update_interval: int = 1000
    seed: int = 0
    warmup_size: int = 1000
    epsilon: float = 0.1
    eps_end: float = 0.1  # linear schedule epsilon -> eps_end over eps_phase_steps vector steps (0 = constant epsilon)
    eps_phase_steps: int = 0
    eps_table: Optional[tuple] = None  # any other schedule, tabulated per vector step (include/srlx.h srlx_engine.eps_table)
    discount: float = 0.99
    lr: float = 1e-3
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_eps: float = 1e-8
    retrace_h: float = 1.0
    per_alpha: float = 0.6
    per_beta_initial: float = 0.4
    per_beta_steps: float = 1_000_000
    per_epsilon: float = 1e-4
    reward_shift: float = 0.0
    reward_scale: float = 1.0
    huber_delta: float = 1.0
    hidden: tuple = (64, 64)
    dueling: Optional[str] = None
    noisy: bool = False
    invalid_actions: bool = False  # keep the next state's invalid-action mask per row (external envs; generic learner only)
    env_kwargs: dict = field(default_factory=dict)


class DeviceEngine:
    def __init__(self, cfg: EngineConfig, device="cuda:0", debug: bool = False, params=None, stream=None, track_episodes: bool = False):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.SrlxError("DeviceEngine needs a CUDA device (no CPU fallback)")
        self.cfg = cfg
        self.device = torch.device(device)
        self.env = make_env_spec(cfg.env, **cfg.env_kwargs)
        self.spec = NetSpec(self.env.obs_dim, tuple(cfg.hidden), self.env.n_actions, cfg.dueling, cfg.noisy, cfg.algo)
        self.E, self.R, self.D, self.A, self.M, self.B = cfg.n_envs, cfg.ring_rows, self.env.obs_dim, self.env.n_actions, cfg.multisteps, cfg.batch_size
        self.cap = self.E * self.R
        self.per = cfg.mem_kind == _lib.MEM_PROPORTIONAL
        self.debug = debug
        self.stream = stream
        dev = self.device
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
        P, noisy = self.spec.n_params, cfg.noisy
        self.t = dict(
            state=z(C.sizeof(_lib.SrlxState), torch.uint8),
            env_state=z((self.E, 4), torch.float64),
            env_step_num=z(self.E, torch.int32),
            env_episode=z(self.E, torch.int32),
            env_ep_reward=z(self.E, torch.float64),
            env_needs_reset=z(self.E, torch.uint8),
            ring_obs=z((self.cap, self.D), torch.float32),
            ring_next_obs=z((self.cap, self.D), torch.float32),
            ring_action=z(self.cap, torch.int32),
            ring_reward=z(self.cap, torch.float32),
            ring_term=z(self.cap, torch.uint8),
            ring_done=z(self.cap, torch.uint8),
            params=z(P, torch.float32),
            target=z(P, torch.float32),
            adam_m=z(P * (2 if noisy else 1), torch.float32),
            adam_v=z(P * (2 if noisy else 1), torch.float32),
        )
        if self.per:
            self.t["tree"] = z(2 * self.cap - 1, torch.float64)
            self.t["tree_scratch"] = z(2 * (self.E + 2), torch.float64)
            # blocked copy of the deep SumTree levels for the learner's sampler (csrc/learner_fast.cu), rebuilt per learn call
            self.t["tree_blk"] = z(max(64, self.lib.srlx_tree_blk_bytes(self.cap) // 8), torch.float64)
        if cfg.invalid_actions:
            self.t["ring_invalid"] = z(self.cap, torch.int32)
        if cfg.eps_table is not None and len(cfg.eps_table):
            self.t["eps_table"] = torch.as_tensor(np.asarray(cfg.eps_table, dtype=np.float64)).to(dev)
        if track_episodes:
            self.t["env_first_ep_reward"] = z(self.E, torch.float64)
            self.t["env_last_ep_len"] = z(self.E, torch.int32)
        if noisy:
            self.t["params_sigma"] = z(P, torch.float32)
            self.t["target_sigma"] = z(P, torch.float32)
            # NoisyNet draws of a chunk of 256 trainer updates (csrc/learner_fast.cu): 3 forward calls x the parameters in
            # CTA-local order (a few replicated / padding floats per CTA on top of P) per update; ~40 MB at the default net
            self.t["noise_scratch"] = torch.empty(256 * 3 * (2 * P + 4096), dtype=torch.float32, device=dev)
        if debug:
            B, M, D, A = self.B, self.M, self.D, self.A
            self.t.update(
                dbg_q=z((self.E, A), torch.float32), dbg_action=z(self.E, torch.int32), dbg_sample_idx=z(B, torch.int64),
                dbg_weights=z(B, torch.float32), dbg_target_q=z(B, torch.float32), dbg_q_sa=z(B, torch.float32),
                dbg_grads=z(P * (2 if noisy else 1), torch.float32), dbg_windows=z(B * (M + 1) * D + 3 * B * M, torch.float32),
                dbg_clock=z(64, torch.int64),
            )
        self.c = self._build_struct()
        # pinned host mirror of the device counters (one small D2H per read)
        self._state_host = torch.zeros(C.sizeof(_lib.SrlxState), dtype=torch.uint8).pin_memory()
        self.reset()
        if params is None:
            mu, sigma = self.spec.init_params(cfg.seed)
        else:
            mu, sigma = params
        self.set_params(mu, sigma, also_target=True)

    # ---- struct ---------------------------------------------------------------------------------------------
    def _build_struct(self):
        cfg, c = self.cfg, _lib.SrlxEngine()
        self.env.fill(c)
        c.n_envs, c.ring_rows, c.multisteps, c.batch_size, c.mem_kind = self.E, self.R, self.M, self.B, cfg.mem_kind
        c.enable_double_dqn, c.enable_rescale, c.enable_reward_clip = int(cfg.enable_double_dqn), int(cfg.enable_rescale), int(cfg.enable_reward_clip)
        c.has_duplicate, c.target_update_interval = int(cfg.has_duplicate), int(cfg.target_update_interval)
        c.presample = int(bool(getattr(cfg, 'presample', False)))
        c.seed, c.warmup_size = int(cfg.seed) & 0xFFFFFFFFFFFFFFFF, int(cfg.warmup_size)
        for k in ("epsilon", "discount", "lr", "adam_beta1", "adam_beta2", "adam_eps", "retrace_h", "per_alpha", "per_beta_initial",
                  "per_beta_steps", "per_epsilon", "reward_shift", "reward_scale", "huber_delta"):
            setattr(c, k, float(getattr(cfg, k)))
        c.eps_end, c.eps_phase_steps = float(cfg.eps_end), int(cfg.eps_phase_steps)
        c.net = self.spec.to_c()
        dp = getattr(self, "_dp", None)
        if dp is not None:  # data-parallel learner (parallel.link_engines / link_engine_distributed)
            c.learner_seed, c.dp_world, c.dp_rank, c.dp_bytes = dp["learner_seed"], dp["world"], dp["rank"], dp["nbytes"]
            for r, ptr in enumerate(dp["peers"]):
                c.dp_peer[r] = ptr
        if "noise_scratch" in self.t:
            c.noise_scratch_bytes = self.t["noise_scratch"].numel() * 4
        if "tree_blk" in self.t:
            c.tree_blk_bytes = self.t["tree_blk"].numel() * 8
        if "eps_table" in self.t:
            c.eps_table_len = self.t["eps_table"].numel()
        for name, _ in _lib.SrlxEngine._fields_:
            if name in self.t:
                setattr(c, name, self.t[name].data_ptr())
        return c

    def enable_invalid_actions(self):
        """Start keeping invalid-action masks (include/srlx.h srlx_engine.ring_invalid): rows written so far had none.  From here on
        srlx_learn runs the generic learner, which applies them (dqn.py:156-165, rainbow.py:236-249)."""
        if "ring_invalid" not in self.t:
            self.t["ring_invalid"] = torch.zeros(self.cap, dtype=torch.int32, device=self.device)
            self.c = self._build_struct()

    def ext_step(self, obs, next_obs, action, reward, term, done, next_invalid=None):
        """One row of E records produced by a host loop (srlx_ext_step_masked): float32 [E][D] x 2, int32 [E], float32 [E], uint8 [E] x 2,
        optional uint32 bit masks [E] of the next state's invalid actions."""
        f = lambda x, dt: torch.as_tensor(np.asarray(x)).to(self.device, dtype=dt).contiguous()  # noqa: E731
        bufs = [f(obs, torch.float32).reshape(self.E, self.D), f(next_obs, torch.float32).reshape(self.E, self.D), f(action, torch.int32),
                f(reward, torch.float32), f(term, torch.uint8), f(done, torch.uint8)]
        inv = None
        if next_invalid is not None:
            self.enable_invalid_actions()
            inv = f(np.asarray(next_invalid, dtype=np.int64), torch.int32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_ext_step_masked(self.c, *[b.data_ptr() for b in bufs], None if inv is None else inv.data_ptr(),
                                                     self._stream()))
        self._holds_data = True
        self._keep_ext = (bufs, inv)

    def set_data_parallel(self, world: int, rank: int, peer_ptrs, nbytes: int, learner_seed: int):
        """Make this engine rank `rank` of a `world`-rank data-parallel learner: peer_ptrs[r] = the exchange buffer of rank r as
        addressable from this device (include/srlx.h srlx_engine.dp_peer); world <= 1 turns it off."""
        if world <= 1:
            self._dp = None
        else:
            if not (0 <= rank < world <= 8) or len(peer_ptrs) != world:
                raise ValueError("set_data_parallel: rank / world / peer_ptrs inconsistent (world <= 8)")
            self._dp = dict(world=int(world), rank=int(rank), peers=[int(p) for p in peer_ptrs], nbytes=int(nbytes),
                            learner_seed=int(learner_seed) & 0xFFFFFFFFFFFFFFFF)
        self.c = self._build_struct()

    def dp_bytes(self) -> int:
        with torch.cuda.device(self.device):
            return int(self.lib.srlx_dp_bytes(C.byref(self.c)))

    def check_dp_alive(self):
        """After a data-parallel launch: raises if a peer rank stopped answering inside the kernel (the wait gave up)."""
        if getattr(self, "_dp", None) is not None and self.read_state().reserved[0] != 0:
            raise _lib.SrlxError("data-parallel learner: a peer rank did not answer within the in-kernel timeout; results are invalid")

    def adopt_tensors(self, tensors: dict):
        """Bind caller-owned tensors (same shape and dtype) in place of the engine's own -- e.g. the parameter buffers of an
        RLParameter object (srl_classes.DeviceParameter) -- and rebuild the C struct around them."""
        for k, v in tensors.items():
            old = self.t.get(k)
            if old is None or old.shape != v.shape or old.dtype != v.dtype or v.device != old.device:
                raise ValueError(f"adopt_tensors: {k!r} does not match the engine's buffer")
            self.t[k] = v
        self.c = self._build_struct()

    def _stream(self):
        s = self.stream if self.stream is not None else torch.cuda.current_stream(self.device)
        return s.cuda_stream

    # ---- control ---------------------------------------------------------------------------------------------
    def reset(self):
        self._holds_data = False
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_engine_reset(C.byref(self.c), self._stream()))

    def set_epsilon(self, eps: float):
        self.c.epsilon = float(eps)

    def _guard_eval(self, training):
        # evaluation steps advance the vector step counter, which is also the ring cursor and the base of the sampleable range:
        # mixing them into an engine whose ring holds data would leave unwritten rows inside that range (VecRunner.evaluate
        # builds its own engine for this reason)
        if training:
            self._holds_data = True
        elif getattr(self, "_holds_data", False):
            raise _lib.SrlxError("evaluation steps (training=False) on an engine whose replay ring holds data: use a separate "
                                 "engine (VecRunner.evaluate) or reset() first")

    def vec_step(self, training=True):
        self._guard_eval(training)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_vec_step(C.byref(self.c), int(training), self._stream()))

    def learn(self, n_updates=1):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_learn(C.byref(self.c), int(n_updates), self._stream()))

    def learner_info(self):
        """(kernel, cluster size, shared-memory bytes per CTA) srlx_learn uses for this engine."""
        cs, sm = C.c_int(0), C.c_size_t(0)
        with torch.cuda.device(self.device):
            rc = self.lib.srlx_learner_info(C.byref(self.c), C.byref(cs), C.byref(sm))
        if rc < 0:
            _lib.check(rc)
        return ({1: "learner_fast_kernel", 2: "learner_small_kernel"}.get(rc, "learner_kernel"), cs.value, sm.value)

    def run(self, n_steps, updates_per_step, training=True):
        """n_steps x (one vector step of all E envs + updates_per_step trainer updates), no host sync in between."""
        self._guard_eval(training)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_engine_run(C.byref(self.c), int(n_steps), int(updates_per_step), int(training), self._stream()))

    def pred_q(self, obs: np.ndarray, target=False, noise_call_id=0) -> np.ndarray:
        """RLParameter.pred_q / pred_target_q (srl/algorithms/dqn/model_torch.py:58-70)."""
        x = torch.as_tensor(np.ascontiguousarray(obs, dtype=np.float32)).reshape(-1, self.D).to(self.device)
        q = torch.empty((x.shape[0], self.A), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_qnet_forward(C.byref(self.c), int(target), x.data_ptr(), x.shape[0], int(noise_call_id), q.data_ptr(), self._stream()))
        return q.cpu().numpy()

    def pred_q_tc(self, obs, target=False, noise_call_id=0) -> torch.Tensor:
        """pred_q / pred_target_q for LARGE batches on the tensor cores (csrc/qnet_tc.cu: tcgen05.mma, TMEM accumulators, TMA-staged
        bf16 operands, fp32 accumulation).  An explicit non-parity mode: results agree with `pred_q` to bf16 resolution (~1e-2
        relative), not to the 1e-4 of the fp32 path.  obs: [n, D] numpy array or CUDA tensor; returns a CUDA tensor [n, A]."""
        x = obs if isinstance(obs, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(obs, dtype=np.float32))
        x = x.to(self.device, dtype=torch.float32).reshape(-1, self.D).contiguous()
        n = x.shape[0]
        q = torch.empty((n, self.A), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            need = int(self.lib.srlx_qnet_tc_workspace_bytes(C.byref(self.c), n))
            ws = getattr(self, "_tc_ws", None)
            if ws is None or ws.numel() < need:
                ws = self._tc_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.srlx_qnet_forward_tc(C.byref(self.c), int(target), x.data_ptr(), n, int(noise_call_id), q.data_ptr(),
                                                     ws.data_ptr(), ws.numel(), self._stream()))
        return q

    def noise(self, kind: int, call_id: int) -> np.ndarray:
        out = torch.empty(self.spec.n_params, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_noise_fill(int(self.cfg.seed) & 0xFFFFFFFFFFFFFFFF, int(kind), int(call_id), out.data_ptr(), self.spec.n_params, self._stream()))
        return out.cpu().numpy()

    # ---- state / parameters -------------------------------------------------------------------------------------
    def read_state(self) -> "_lib.SrlxState":
        self._state_host.copy_(self.t["state"], non_blocking=False)
        return _lib.SrlxState.from_buffer_copy(self._state_host.numpy().tobytes())

    def write_state(self, st: "_lib.SrlxState"):
        self.t["state"].copy_(torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8))

    def set_params(self, mu, sigma=None, also_target=False):
        self.t["params"].copy_(torch.as_tensor(np.asarray(mu, dtype=np.float32)))
        if self.cfg.noisy:
            self.t["params_sigma"].copy_(torch.as_tensor(np.asarray(sigma, dtype=np.float32)))
        if also_target:
            self.set_target(mu, sigma)

    def set_target(self, mu, sigma=None):
        self.t["target"].copy_(torch.as_tensor(np.asarray(mu, dtype=np.float32)))
        if self.cfg.noisy:
            self.t["target_sigma"].copy_(torch.as_tensor(np.asarray(sigma, dtype=np.float32)))

    def get_params(self):
        return self.t["params"].cpu().numpy(), (self.t["params_sigma"].cpu().numpy() if self.cfg.noisy else None)

    def get_target(self):
        return self.t["target"].cpu().numpy(), (self.t["target_sigma"].cpu().numpy() if self.cfg.noisy else None)

    def state_dict(self):
        """The reference-compatible state_dict of the online network (RLParameter.call_backup)."""
        mu, sigma = self.get_params()
        return self.spec.to_state_dict(mu, sigma)

    def load_state_dict(self, sd, also_target=True):
        mu, sigma = self.spec.from_state_dict(sd)
        self.set_params(mu, sigma, also_target=also_target)

    # ---- replay memory <-> reference backup format (checkpoint.py) ----------------------------------------
    def ring_view(self):
        """Host copy of the ring (+ SumTree leaves) as a checkpoint.RingView."""
        from . import checkpoint

        st = self.read_state()
        v = checkpoint.RingView(self.E, self.R, self.M, self.A, self.D, vec_steps=int(st.vec_steps))
        v.obs[:] = self.t["ring_obs"].cpu().numpy().reshape(-1, self.D)
        v.next_obs[:] = self.t["ring_next_obs"].cpu().numpy().reshape(-1, self.D)
        v.action[:] = self.t["ring_action"].cpu().numpy().reshape(-1)
        v.reward[:] = self.t["ring_reward"].cpu().numpy().reshape(-1)
        v.term[:] = self.t["ring_term"].cpu().numpy().reshape(-1)
        v.done[:] = self.t["ring_done"].cpu().numpy().reshape(-1)
        if "ring_invalid" in self.t:
            v.invalid = self.t["ring_invalid"].cpu().numpy().reshape(-1).astype(np.uint32)
        if self.per:
            cap = self.E * self.R
            v.leaf_priority = self.t["tree"][cap - 1:].cpu().numpy().astype(np.float64)
            v.max_priority = float(st.max_priority)
        return v

    def load_ring(self, v):
        """Replace the replay contents with a checkpoint.RingView (same E, R, M); counters follow the reference's
        restore: memory length = the restored items, the vector step counter moves to the restored row count."""
        from . import checkpoint

        if (v.E, v.R, v.M, v.D) != (self.E, self.R, self.M, self.D):
            raise ValueError("load_ring: ring geometry differs from the engine's")
        for name, arr in (("ring_obs", v.obs), ("ring_next_obs", v.next_obs), ("ring_action", v.action), ("ring_reward", v.reward),
                          ("ring_term", v.term), ("ring_done", v.done)):
            t = self.t[name]
            t.copy_(torch.as_tensor(np.ascontiguousarray(arr)).reshape(t.shape).to(t.dtype))
        if getattr(v, "invalid", None) is not None:
            self.enable_invalid_actions()
            self.t["ring_invalid"].copy_(torch.as_tensor(np.ascontiguousarray(v.invalid).astype(np.int64)).to(torch.int32))
        elif "ring_invalid" in self.t:
            self.t["ring_invalid"].zero_()
        st = self.read_state()
        g_lo, n_g = v.valid_rows()
        st.vec_steps, st.total_step, st.mem_size = v.vec_steps, v.vec_steps * self.E, n_g * self.E
        if self.per:
            leaves = v.leaf_priority if v.leaf_priority is not None else np.zeros(self.E * self.R)
            self.t["tree"].copy_(torch.as_tensor(checkpoint.build_sum_tree(leaves, self.E * self.R)))
            st.max_priority = float(v.max_priority)
        self.write_state(st)
        self._holds_data = True
        self.t["env_needs_reset"].fill_(1)  # the env copies start fresh episodes after a restore

    def hbm_bytes(self):
        return sum(t.numel() * t.element_size() for t in self.t.values())
