"""Sequential CPU port of the vectorised actor/replay/learner engine (TEST INFRASTRUCTURE).

This is the executable specification the CUDA engine (libsrlx.so: srlx_vec_step / srlx_learn) is compared against,
buffer for buffer, on small sizes.  It composes the per-function restatements:

  vec_step   core_play.play loop body (srl/base/run/core_play.py:115-214) for E independent env copies:
             reset-if-done (:138-159) -> WorkerRun.policy (srl/base/rl/worker_run.py:360) -> dqn/rainbow Worker.policy
             (srl/algorithms/dqn/dqn.py:192-211, srl/algorithms/rainbow/rainbow.py:301-331) -> EnvRun.step
             (srl/base/env/env_run.py:254-366) -> Worker.on_step record (dqn.py:213-246, rainbow.py:333-371)
             -> memory add (srl/rl/memories/priority_replay_buffer.py:205-217)
  learn      Trainer.train (srl/algorithms/dqn/model_torch.py:90-132, srl/algorithms/rainbow/model_torch.py:85-122)

Vectorisation choices that have no single-env counterpart in the reference (stated in DESIGN.md):
  * slot(g, e) = (g % R) * E + e -- the ring is time-major; one vector step fills one row of E slots.
  * n-step (rainbow multisteps=M): per-step records are stored once and the M+1 window of an item is rebuilt by index
    at sample time; an item (row g) becomes sampleable at vector step g+M-1 (the reference emits the window at the same
    env step, rainbow.py:373-400, except that the M-1 tail windows of a finished episode are emitted immediately).
  * RNG: Philox4x32-10 counters (oracle/philox.py) instead of the global Mersenne-Twister streams.
  * NoisyNet: one noise draw per forward CALL (noisy_linear.py:35-52); a vector step is one call with batch E.
"""
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

from . import nets, philox, sumtree, targets
from .envs import make_spec

MEM_UNIFORM, MEM_PROPORTIONAL = 0, 1


@dataclass
class EngineConfig:
    env: str = "Grid"
    n_envs: int = 8
    ring_rows: int = 16
    multisteps: int = 1
    batch_size: int = 4
    mem_kind: int = MEM_PROPORTIONAL
    algo: str = "dqn"  # "dqn" | "rainbow" (state_dict naming only)
    enable_double_dqn: bool = True
    enable_rescale: bool = False
    enable_reward_clip: bool = False
    has_duplicate: bool = True
    presample: bool = False  # batch t+1 is drawn before update t lands (one launch = up to 256 updates), see learn()
    target_update_interval: int = 1000
    seed: int = 0
    warmup_size: int = 16
    epsilon: float = 0.1
    eps_end: float = 0.1  # linear schedule epsilon -> eps_end over eps_phase_steps steps (0 = constant epsilon)
    eps_phase_steps: int = 0
    eps_table: Optional[tuple] = None  # epsilon per vector step, held at its last entry (any scheduler, tabulated)
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
    env_kwargs: dict = field(default_factory=dict)


def default_noise_fn(seed, n_params):
    """numpy Box-Muller over the same Philox words as csrc/philox.cuh::noise4 (float32; the device's logf/sincospif differ
    in the last ulp, so GPU parity tests pull the noise from srlx_noise_fill instead)."""

    def f(kind, call_id):
        nblk = (n_params + 3) // 4
        a = np.arange(nblk, dtype=np.uint32)
        c = np.uint32(((call_id >> 32) & 0x0FFFFFFF) | (kind << 28))
        w0, w1, w2, w3 = philox.words(seed, philox.STREAM_NOISE, a, np.uint32(call_id & 0xFFFFFFFF), c)
        out = np.empty((nblk, 4), dtype=np.float32)
        for j, (wa, wb) in enumerate(((w0, w1), (w2, w3))):
            u1 = ((wa >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)
            u2 = (wb >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
            r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
            ang = (np.float32(2.0) * u2).astype(np.float64) * np.pi
            out[:, 2 * j] = (r * np.cos(ang).astype(np.float32)).astype(np.float32)
            out[:, 2 * j + 1] = (r * np.sin(ang).astype(np.float32)).astype(np.float32)
        return out.reshape(-1)[:n_params].copy()

    return f


class OracleEngine:
    def __init__(self, cfg: EngineConfig, mu, sigma=None, noise_fn: Optional[Callable] = None):
        self.cfg = cfg
        self.env = make_spec(cfg.env, **cfg.env_kwargs)
        E, R, D = cfg.n_envs, cfg.ring_rows, self.env.obs_dim
        self.E, self.R, self.D, self.A, self.M = E, R, D, self.env.n_actions, cfg.multisteps
        self.cap = R * E
        self.spec = nets.NetSpec(D, tuple(cfg.hidden), self.A, cfg.dueling, cfg.noisy)
        self.sigma_mask = self.spec.sigma_mask(cfg.algo) if cfg.noisy else None
        self.adam = nets.AdamState(self.spec, mu, sigma, lr=cfg.lr, betas=(cfg.adam_beta1, cfg.adam_beta2), eps=cfg.adam_eps)
        self.tgt_mu = np.array(mu, dtype=np.float32).copy()
        self.tgt_sigma = None if sigma is None else np.array(sigma, dtype=np.float32).copy()
        self.noise_fn = noise_fn or default_noise_fn(cfg.seed, self.spec.n_params)
        # env state
        self.env_state = np.zeros((E, 4), dtype=np.float64)
        self.step_num = np.zeros(E, dtype=np.int32)
        self.episode = np.zeros(E, dtype=np.uint32)
        self.ep_reward = np.zeros(E, dtype=np.float64)
        self.needs_reset = np.ones(E, dtype=np.uint8)
        # ring
        self.ring_obs = np.zeros((self.cap, D), dtype=np.float32)
        self.ring_next_obs = np.zeros((self.cap, D), dtype=np.float32)
        self.ring_action = np.zeros(self.cap, dtype=np.int32)
        self.ring_reward = np.zeros(self.cap, dtype=np.float32)
        self.ring_term = np.zeros(self.cap, dtype=np.uint8)
        self.ring_done = np.zeros(self.cap, dtype=np.uint8)
        self.ring_invalid = None  # uint32 [cap] bit masks of the next state's invalid actions (external envs; ext_step)
        # memory
        self.per = sumtree.ProportionalMemory(self.cap, cfg.per_alpha, cfg.per_beta_initial, cfg.per_beta_steps,
                                              cfg.has_duplicate, cfg.per_epsilon)
        # counters
        self.vec_steps = 0
        self.total_step = 0
        self.train_count = 0
        self.episode_count = 0
        self.sync_count = 0
        self.mem_size = 0
        self.episode_reward_sum = 0.0
        self.episode_len_sum = 0
        self.last_loss = 0.0
        self.sample_retries = 0

    # ---- parameters -----------------------------------------------------------------------------------
    @property
    def mu(self):
        return self.adam.mu.detach().numpy()

    @property
    def sigma(self):
        return None if self.adam.sigma is None else self.adam.sigma.detach().numpy()

    # ---- replay contents from a ring view (tests: frozen memories, device state downloaded at full size) -----------
    def load_ring(self, v, tree=None):
        """Adopt the replay contents of a checkpoint.RingView-like object (obs, next_obs, action, reward, term, done, vec_steps,
        leaf_priority, max_priority), as DeviceEngine.load_ring does; `tree` = the full node array if the caller has one."""
        assert (v.E, v.R, v.M, v.D) == (self.E, self.R, self.M, self.D)
        self.ring_obs[:], self.ring_next_obs[:] = v.obs, v.next_obs
        self.ring_action[:], self.ring_reward[:], self.ring_term[:], self.ring_done[:] = v.action, v.reward, v.term, v.done
        self.vec_steps = int(v.vec_steps)
        self.total_step = self.vec_steps * self.E
        self.mem_size = self._valid_range()[1] * self.E
        self.per.size = self.mem_size
        if self.cfg.mem_kind == MEM_PROPORTIONAL:
            if tree is not None:
                self.per.tree.tree[:] = tree
            else:
                t = self.per.tree.tree
                t[:] = 0.0
                t[self.cap - 1:] = v.leaf_priority
                for i in range(self.cap - 2, -1, -1):
                    t[i] = t[2 * i + 1] + t[2 * i + 2]
            self.per.max_priority = float(v.max_priority)
        self.needs_reset[:] = 1

    # ---- SumTree bulk row set (device: tree_set_row kernel) ------------------------------------------------
    def _tree_set_range(self, leaf_lo, values):
        """leaves [leaf_lo, leaf_lo+n) <- values; ancestors += pairwise-summed change (left child + right child)."""
        tree = self.per.tree.tree
        a = leaf_lo + self.cap - 1
        b = a + len(values) - 1
        change = np.asarray(values, dtype=np.float64) - tree[a : b + 1]
        tree[a : b + 1] = values
        # a row of a non-power-of-two tree can straddle the two leaf depths, so [a, b] may span two levels: climb until
        # the whole range has collapsed into the root (b == 0); the root has no parent
        while b > 0:
            lo = max(a, 1)
            pa, pb = (lo - 1) // 2, (b - 1) // 2
            pc = np.zeros(pb - pa + 1, dtype=np.float64)
            for i in range(lo, b + 1):
                pc[(i - 1) // 2 - pa] += change[i - a]
            tree[pa : pb + 1] += pc
            a, b, change = pa, pb, pc

    def _tree_set_const(self, leaf_lo, n, value):
        """leaves [leaf_lo, leaf_lo+n) <- value.  When n is a power of two (>= 2) and the range is a ring row, those leaves are the
        leaves of ONE complete subtree rooted at node leaf_lo / n + cap / n - 1: every node below that root is exactly value * (leaves
        under it); the root takes value * n and its ancestors the root's change (device twin: csrc/rollout.cu::post_step_pow2_kernel).
        Otherwise: the level-by-level bulk set."""
        pow2 = n >= 2 and (n & (n - 1)) == 0 and leaf_lo % n == 0 and self.cap % n == 0 and self.cap // n >= 2
        if not pow2:
            return self._tree_set_range(leaf_lo, np.full(n, value, dtype=np.float64))
        tree = self.per.tree.tree
        root = leaf_lo // n + self.cap // n - 1
        logn = n.bit_length() - 1
        for k in range(1, logn + 1):
            first = ((root + 1) << k) - 1
            tree[first:first + (1 << k)] = value * float(1 << (logn - k))
        nv = value * float(n)
        delta = nv - tree[root]
        tree[root] = nv
        p = root
        while p > 0:
            p = (p - 1) // 2
            tree[p] = tree[p] + delta

    # ---- one vector step ---------------------------------------------------------------------------------
    def epsilon_at(self, step):
        """Linear.update(step).to_float() (srl/rl/schedulers/schedulers/linear.py:11-21); phase 0 = Constant (constant.py)."""
        cfg = self.cfg
        if cfg.eps_table is not None and len(cfg.eps_table):
            return float(cfg.eps_table[min(int(step), len(cfg.eps_table) - 1)])
        if not cfg.eps_phase_steps:
            return cfg.epsilon
        if step >= cfg.eps_phase_steps:
            return cfg.eps_end
        step_rate = (cfg.epsilon - cfg.eps_end) / cfg.eps_phase_steps
        return cfg.epsilon - step_rate * step

    def vec_step(self, training=True, q_override=None):
        cfg, env, E = self.cfg, self.env, self.E
        g = self.vec_steps
        row = g % self.R
        obs = np.zeros((E, self.D), dtype=np.float32)
        for e in range(E):
            if self.needs_reset[e]:
                self.env_state[e] = env.reset(cfg.seed, e, int(self.episode[e]))
                self.episode[e] += 1
                self.step_num[e] = 0
                self.ep_reward[e] = 0.0
                self.needs_reset[e] = 0
            obs[e] = env.obs(self.env_state[e])
        noise = self.noise_fn(nets.NOISE_KIND_ROLLOUT, g) if cfg.noisy else None
        q = nets.np_forward(self.spec, self.mu, self.sigma, noise, obs)
        q_own = q
        if q_override is not None:  # drive the policy with the device's Q so the action indices are comparable bit for bit
            q = np.asarray(q_override, dtype=np.float32)
        actions = np.zeros(E, dtype=np.int32)
        eps = np.float32(self.epsilon_at(g) if training else cfg.epsilon)
        for e in range(E):
            w = philox.words(cfg.seed, philox.STREAM_POLICY, e, g & 0xFFFFFFFF, g >> 32)
            u = philox.u01_f32(w[0])
            if (not cfg.noisy) and (u < eps):
                a = (int(w[1]) * self.A) >> 32
            else:
                a = int(np.argmax(q[e]))
            actions[e] = a
            nst, r, terminated = env.step(self.env_state[e], a, cfg.seed, e, g)
            self.step_num[e] += 1
            truncated = self.step_num[e] >= env.trunc_limit
            if env.trunc_overrides_term:
                term_flag = terminated and not truncated
            else:
                term_flag = terminated
                truncated = truncated and not terminated
            done = bool(terminated or truncated)
            self.ep_reward[e] += r
            rr = (r + cfg.reward_shift) * cfg.reward_scale
            if cfg.enable_reward_clip:
                rr = -1.0 if rr < 0 else (1.0 if rr > 0 else 0.0)
            if training:
                slot = row * E + e
                self.ring_obs[slot] = obs[e]
                self.ring_next_obs[slot] = env.obs(nst)
                self.ring_action[slot] = a
                self.ring_reward[slot] = np.float32(rr)
                self.ring_term[slot] = 1 if term_flag else 0
                self.ring_done[slot] = 1 if done else 0
            self.env_state[e] = nst
            if done:
                self.episode_count += 1
                self.episode_reward_sum += self.ep_reward[e]
                self.episode_len_sum += int(self.step_num[e])
                self.needs_reset[e] = 1
        if training:
            M, R = self.M, self.R
            if cfg.mem_kind == MEM_PROPORTIONAL:
                if M == 1:
                    self._tree_set_const(row * E, E, self.per.max_priority)
                else:
                    self._tree_set_const(row * E, E, 0.0)
                    if g >= M - 1:
                        self._tree_set_const(((g - M + 1) % R) * E, E, self.per.max_priority)
            rows_added = max(0, g + 1 - (M - 1))
            self.mem_size = E * min(rows_added, R - (M - 1))
            self.per.size = self.mem_size
            self.vec_steps += 1
            self.total_step += E
        return dict(q=q_own, actions=actions, obs=obs)

    # ---- window rebuild ----------------------------------------------------------------------------------
    def window(self, slot):
        cfg, E, R, M = self.cfg, self.E, self.R, self.M
        rho, e = divmod(int(slot), E)
        g_last = self.vec_steps - 1
        g_item = g_last - ((g_last - rho) % R)
        states = np.zeros((M + 1, self.D), dtype=np.float32)
        acts = np.zeros(M, dtype=np.int64)
        rews = np.zeros(M, dtype=np.float32)
        terms = np.zeros(M, dtype=np.float32)
        states[0] = self.ring_obs[slot]
        ended, g_end = False, 0
        for k in range(M):
            if not ended:
                sk = ((rho + k) % R) * E + e
                acts[k] = self.ring_action[sk]
                rews[k] = self.ring_reward[sk]
                terms[k] = self.ring_term[sk]
                states[k + 1] = self.ring_next_obs[sk]
                if self.ring_done[sk]:
                    ended, g_end = True, g_item + k
            else:
                gp = g_item + k  # virtual step of the padded record (rainbow.py:358-371)
                w = philox.words(cfg.seed, philox.STREAM_PAD_ACTION, e, gp & 0xFFFFFFFF, gp >> 32)
                acts[k] = (int(w[0]) * self.A) >> 32
                rews[k] = 0.0
                terms[k] = 1.0
                states[k + 1] = states[k]
        return states, acts, rews, terms

    def window_invalid(self, slot):
        """bool [M, A]: invalid actions of the window's M next states; padded records carry none (rainbow.py:366)"""
        E, R, M = self.E, self.R, self.M
        rho, e = divmod(int(slot), E)
        out = np.zeros((M, self.A), dtype=bool)
        for k in range(M):
            sk = ((rho + k) % R) * E + e
            out[k] = [(int(self.ring_invalid[sk]) >> a) & 1 for a in range(self.A)]
            if self.ring_done[sk]:
                break
        return out

    def ext_step(self, obs, next_obs, action, reward, term, done, next_invalid=None):
        """One row of E records from a host loop (csrc/rollout.cu: srlx_ext_step_masked) + the replay add of vec_step."""
        cfg, E = self.cfg, self.E
        g = self.vec_steps
        row = g % self.R
        sl = slice(row * E, (row + 1) * E)
        self.ring_obs[sl], self.ring_next_obs[sl] = obs, next_obs
        self.ring_action[sl], self.ring_reward[sl] = action, np.asarray(reward, np.float32)
        self.ring_term[sl], self.ring_done[sl] = term, done
        if self.ring_invalid is not None:
            self.ring_invalid[sl] = 0 if next_invalid is None else next_invalid
        M, R = self.M, self.R
        if cfg.mem_kind == MEM_PROPORTIONAL:
            if M == 1:
                self._tree_set_const(row * E, E, self.per.max_priority)
            else:
                self._tree_set_const(row * E, E, 0.0)
                if g >= M - 1:
                    self._tree_set_const(((g - M + 1) % R) * E, E, self.per.max_priority)
        self.mem_size = E * min(max(0, g + 1 - (M - 1)), R - (M - 1))
        self.per.size = self.mem_size
        self.vec_steps += 1
        self.total_step += E

    def _valid_range(self):
        g_next, R, M = self.vec_steps, self.R, self.M
        g_lo = max(0, g_next - R)
        n_g = g_next - M + 1 - g_lo
        return g_lo, max(0, n_g)

    # ---- trainer updates ---------------------------------------------------------------------------------
    def _sample(self, tc):
        """(idx, slots, weights) of the batch of update `tc`, drawn from the memory as it is NOW."""
        cfg, B = self.cfg, self.cfg.batch_size
        if cfg.mem_kind == MEM_PROPORTIONAL:
            beta_step = max(tc - 1, 0)  # PriorityReplayBuffer.step is the PREVIOUS update's train_count
            idx, w, pri, retries = self.per.sample(B, beta_step, sumtree.philox_uniforms(cfg.seed, tc))
            self.sample_retries += retries
            return idx, idx - (self.cap - 1), w.astype(np.float32)
        g_lo, n_g = self._valid_range()
        pick = sumtree.uniform_sample_distinct(n_g * self.E, B, cfg.seed, tc)
        slots = ((g_lo + pick // self.E) % self.R) * self.E + pick % self.E
        return slots.copy(), slots, np.ones(B, dtype=np.float32)

    def learn(self, n_updates=1, launch_updates=256):
        """n_updates x Trainer.train().  presample=False: every batch is drawn after the previous update's priorities are in the
        tree (the reference's sequential loop).  presample=True: inside one launch (<= launch_updates updates, the device's chunk)
        batch t+1 is drawn BEFORE update t is applied -- one update of staleness, the order the reference's memory process
        produces in distributed mode (srl/base/run/play_mp_memory.py:253-350); the first batch of a launch is drawn fresh."""
        cfg = self.cfg
        out = []
        done = 0
        while done < n_updates:
            k = min(launch_updates, n_updates - done) if cfg.presample else n_updates - done
            pending = None
            for j in range(k):
                if self.mem_size < cfg.warmup_size:
                    continue
                tc = self.train_count
                idx, slots, weights = pending if pending is not None else self._sample(tc)
                pending = None
                wins = [self.window(s) for s in slots]
                states = np.stack([w_[0] for w_ in wins])
                acts = np.stack([w_[1] for w_ in wins])
                rews = np.stack([w_[2] for w_ in wins])
                terms = np.stack([w_[3] for w_ in wins])
                inv = None if self.ring_invalid is None else np.stack([self.window_invalid(s) for s in slots])
                noise = (None, None, None)
                if cfg.noisy:
                    noise = tuple(self.noise_fn(nets.NOISE_KIND_TRAIN, tc * 3 + p) for p in range(3))
                res = nets.train_update(
                    self.spec, self.adam, self.tgt_mu, self.tgt_sigma, algo=cfg.algo, states=states, actions=acts,
                    rewards=rews, dones=terms, weights=weights, discount=cfg.discount, multisteps=self.M,
                    retrace_h=cfg.retrace_h, enable_double_dqn=cfg.enable_double_dqn, enable_rescale=cfg.enable_rescale,
                    noise=noise, sigma_mask=self.sigma_mask, huber_delta=cfg.huber_delta, next_invalid=inv)
                if cfg.presample and j + 1 < k:
                    pending = self._sample(tc + 1)  # before this update's priorities reach the tree
                if cfg.mem_kind == MEM_PROPORTIONAL:
                    self.per.update(idx, res["priorities"])
                if tc % cfg.target_update_interval == 0:
                    self.tgt_mu = self.mu.copy()
                    if self.tgt_sigma is not None:
                        self.tgt_sigma = self.sigma.copy()
                    self.sync_count += 1
                self.train_count += 1
                self.last_loss = res["loss"]
                res.update(idx=idx, slots=slots, weights=weights, states=states, actions=acts, rewards=rews, terms=terms)
                out.append(res)
            done += k
        return out

    def run(self, n_steps, updates_per_step, training=True):
        for _ in range(n_steps):
            self.vec_step(training)
            if training:
                self.learn(updates_per_step)
