"""VecRunner: the srl.Runner-shaped facade over the device engine (train / rollout / train_only / evaluate).

Mirrors the reference's run loop contract for E vectorised env copies:
  core_play.play            srl/base/run/core_play.py:15-238   (stop checks :117-133, train every train_interval
                                                                 steps x train_repeat :187-194, callbacks :164-214)
  play_trainer_only         srl/base/run/core_train_only.py:61-86
  RunContext stop fields    srl/base/context.py:27-109         (max_episodes, timeout, max_steps, max_train_count,
                                                                 max_memory, train_interval, train_repeat)
  RunState counters         srl/base/context.py:297-343        (total_step, train_count, episode_count,
                                                                 episode_rewards_list, last_episode_*, end_reason)
  Runner.evaluate           srl/runner/runner.py:724-773       (returns the per-episode rewards)
  metric definition         srl/runner/callbacks/print_progress.py:224-237 (st/s = d total_step/dt, tr/s = d train_count/dt)

One host iteration = `steps_per_call` vector steps (each E env steps) with their trainer updates enqueued back to back,
then ONE 128-byte device->host read of the counters (srlx_state) from which the stop conditions, RunState and the
callbacks are fed.  Callbacks are duck-typed against srl's RunCallback (on_start / on_step_end / on_episode_end /
on_end, srl/base/run/callback.py:11-78) and are called at batch granularity: on_step_end once per host iteration,
on_episode_end once per host iteration in which at least one episode finished (state.last_episode_rewards then holds the
MEAN reward of the episodes that finished in that iteration).
"""
import time
from dataclasses import dataclass, field
from typing import Any, List, Optional

import numpy as np

from .engine import DeviceEngine, EngineConfig


@dataclass
class VecRunContext:
    """The RunContext fields the hot path reads (srl/base/context.py:27-109)."""
    max_episodes: int = 0
    timeout: float = 0
    max_steps: int = 0
    max_train_count: int = 0
    max_memory: int = 0
    train_interval: int = 1
    train_repeat: int = 1
    training: bool = False
    rollout: bool = False
    train_only: bool = False
    disable_trainer: bool = False
    steps_per_call: int = 1
    callbacks: List[Any] = field(default_factory=list)


@dataclass
class VecRunState:
    """RunState (srl/base/context.py:297-343) as maintained from the device counters."""
    elapsed_t0: float = 0
    episode_rewards_list: List[List[float]] = field(default_factory=list)
    episode_count: int = 0
    total_step: int = 0
    end_reason: str = ""
    train_count: int = 0
    is_step_trained: bool = False
    last_episode_step: float = 0
    last_episode_time: float = 0
    last_episode_rewards: List[float] = field(default_factory=list)
    memory_size: int = 0
    loss: float = 0.0
    sync: int = 0
    shared_vars: dict = field(default_factory=dict)


def _call(callbacks, name, **kw):
    stop = False
    for c in callbacks:
        fn = getattr(c, name, None)
        if fn is not None:
            stop = bool(fn(**kw)) or stop
    return stop


class VecRunner:
    def __init__(self, cfg: EngineConfig, device="cuda:0", params=None, debug=False):
        self.cfg = cfg
        self.engine = DeviceEngine(cfg, device=device, params=params, debug=debug)
        self.state = VecRunState()
        self.context = VecRunContext()
        self._owed = 0.0

    # ---- the loop -------------------------------------------------------------------------------------
    def _play(self, ctx: VecRunContext) -> VecRunState:
        assert ctx.max_episodes > 0 or ctx.timeout > 0 or ctx.max_steps > 0 or ctx.max_train_count > 0 or ctx.max_memory > 0, \
            "Please specify 'max_episodes', 'timeout' , 'max_steps' or 'max_train_count' or 'max_memory'."
        eng, E = self.engine, self.engine.E
        st0 = eng.read_state()
        base_step, base_train, base_ep = st0.total_step, st0.train_count, st0.episode_count
        prev_ep, prev_rsum, prev_lsum = st0.episode_count, st0.episode_reward_sum, st0.episode_len_sum
        state = self.state = VecRunState(elapsed_t0=time.time(), memory_size=int(st0.mem_size))
        self.context = ctx
        cbs = ctx.callbacks
        training = ctx.training and not ctx.disable_trainer
        _call(cbs, "on_start", context=ctx, state=state)
        _call(cbs, "on_episodes_begin", context=ctx, state=state)
        t_last_ep = time.time()
        while True:
            # ---- stop checks (core_play.py:117-133 / core_train_only.py:63-73)
            if ctx.timeout > 0 and (time.time() - state.elapsed_t0) >= ctx.timeout:
                state.end_reason = "timeout."
                break
            if ctx.max_steps > 0 and state.total_step >= ctx.max_steps:
                state.end_reason = "max_steps over."
                break
            if ctx.max_train_count > 0 and state.train_count >= ctx.max_train_count:
                state.end_reason = "max_train_count over."
                break
            if ctx.max_memory > 0 and state.memory_size >= ctx.max_memory:
                state.end_reason = "max_memory over."
                break
            if ctx.max_episodes > 0 and state.episode_count >= ctx.max_episodes:
                state.end_reason = "episode_count over."
                break
            # ---- enqueue steps_per_call x (vector step + its updates)
            n_calls = 1 if ctx.max_train_count > 0 else max(1, int(ctx.steps_per_call))  # exact max_train_count stop
            if ctx.train_only:
                n_upd = E * n_calls
                if ctx.max_train_count > 0:
                    n_upd = min(n_upd, ctx.max_train_count - state.train_count)
                eng.learn(n_upd)
            else:
                if ctx.max_steps > 0:  # never overshoot max_steps by more than one vector step
                    n_calls = max(1, min(n_calls, -(-(ctx.max_steps - state.total_step) // E)))
                for _ in range(n_calls):
                    eng.vec_step(training=ctx.training or ctx.rollout)
                    if training:
                        # one update per train_interval env steps, x train_repeat (core_play.py:187-194)
                        self._owed += E * ctx.train_repeat / max(1, ctx.train_interval)
                        n_upd = int(self._owed)
                        self._owed -= n_upd
                        if ctx.max_train_count > 0:
                            n_upd = min(n_upd, max(0, ctx.max_train_count - state.train_count))
                        if n_upd:
                            eng.learn(n_upd)
            # ---- ONE device->host read of the counters
            s = eng.read_state()
            trained = s.train_count - base_train
            state.is_step_trained = trained > state.train_count
            if ctx.train_only and trained == state.train_count:
                state.end_reason = "memory warmup (nothing trained)."
                state.total_step, state.train_count = int(s.total_step - base_step), int(trained)
                break
            state.total_step = int(s.total_step - base_step)
            state.train_count = int(trained)
            state.memory_size = int(s.mem_size)
            state.loss, state.sync = float(s.last_loss), int(s.sync_count)
            n_ep = s.episode_count - prev_ep
            if n_ep > 0:
                state.episode_count = int(s.episode_count - base_ep)
                mean_r = (s.episode_reward_sum - prev_rsum) / n_ep
                state.last_episode_rewards = [float(mean_r)]
                state.last_episode_step = float((s.episode_len_sum - prev_lsum) / n_ep)
                now = time.time()
                state.last_episode_time, t_last_ep = now - t_last_ep, now
                prev_ep, prev_rsum, prev_lsum = s.episode_count, s.episode_reward_sum, s.episode_len_sum
                _call(cbs, "on_episode_end", context=ctx, state=state)
            if _call(cbs, "on_step_end", context=ctx, state=state):
                state.end_reason = "callback.on_step_end"
                break
        _call(cbs, "on_episodes_end", context=ctx, state=state)
        _call(cbs, "on_end", context=ctx, state=state)
        return state

    # ---- Runner facade (srl/runner/runner.py:95,185,254,724) -----------------------------------------------
    def train(self, max_episodes=0, timeout=0, max_steps=0, max_train_count=0, max_memory=0, train_interval=1,
              train_repeat=1, steps_per_call=1, callbacks=None) -> VecRunState:
        return self._play(VecRunContext(max_episodes, timeout, max_steps, max_train_count, max_memory, train_interval,
                                        train_repeat, training=True, steps_per_call=steps_per_call,
                                        callbacks=list(callbacks or [])))

    def rollout(self, max_episodes=0, timeout=0, max_steps=0, max_memory=0, steps_per_call=1, callbacks=None) -> VecRunState:
        return self._play(VecRunContext(max_episodes, timeout, max_steps, 0, max_memory, training=True, rollout=True,
                                        disable_trainer=True, steps_per_call=steps_per_call, callbacks=list(callbacks or [])))

    def train_only(self, timeout=0, max_train_count=0, callbacks=None) -> VecRunState:
        return self._play(VecRunContext(timeout=timeout, max_train_count=max_train_count, training=True, train_only=True,
                                        callbacks=list(callbacks or [])))

    def evaluate(self, max_episodes=10, test_epsilon=0.0, max_vec_steps=100_000) -> List[float]:
        """Runner.evaluate (runner.py:724-773): greedy (test_epsilon) episodes on fresh env copies, nothing stored;
        returns the reward of the first episode each of `max_episodes` env copies finishes."""
        cfg = EngineConfig(**{**self.cfg.__dict__, "n_envs": int(max_episodes), "ring_rows": max(1, self.cfg.multisteps),
                              "epsilon": float(test_epsilon), "seed": self.cfg.seed + 0x5EED})
        ev = DeviceEngine(cfg, device=self.engine.device, params=self.engine.get_params(), track_episodes=True)
        for i in range(max_vec_steps):
            ev.vec_step(training=False)
            if i % 8 == 7 and bool((ev.t["env_last_ep_len"] > 0).all().item()):
                break
        first = ev.t["env_first_ep_reward"].cpu().numpy().astype(float)
        self.state.episode_rewards_list = [[float(r)] for r in first]
        return [float(r) for r in first]

    def evaluate_compare_to_baseline_single_player(self, episode: int = -1, baseline: Optional[float] = None,
                                                   eval_kwargs: Optional[dict] = None, enable_backup_restore: bool = True) -> bool:
        """Runner.evaluate_compare_to_baseline_single_player (runner.py:1357-1392), the reference's acceptance gate: mean
        reward of `episode` greedy episodes >= the env's reward_baseline (Grid 0.65 over 100, Pendulum-v1 -500 over 10)."""
        rb = self.engine.env.reward_baseline or {}
        if episode <= 0:
            episode = int(rb.get("episode", 0)) or 100
        if baseline is None:
            baseline = rb.get("baseline", None)
        assert baseline is not None, "Please specify a 'baseline'."
        if enable_backup_restore:  # the reference round-trips the parameters through backup()/restore() first
            self.engine.load_state_dict(self.engine.state_dict())
        rewards = self.evaluate(max_episodes=episode, **(eval_kwargs or {}))
        return bool(float(np.mean(rewards)) >= float(baseline))

    # ---- parameters (RLParameter.call_backup / call_restore) ----------------------------------------------
    def state_dict(self):
        return self.engine.state_dict()

    def load_state_dict(self, sd):
        self.engine.load_state_dict(sd)

    # ---- files the reference reads and writes (RunnerBase.save_parameter / load_parameter / save_memory / load_memory,
    #      srl/runner/runner_base.py:141-165; formats in checkpoint.py) ------------------------------------------
    def save_parameter(self, path: str, compress: bool = True) -> None:
        from . import checkpoint

        mu, sigma = self.engine.get_params()
        checkpoint.save_file(path, checkpoint.parameter_backup(self.engine.spec, mu, sigma), compress)

    def load_parameter(self, path: str) -> None:
        from . import checkpoint

        mu, sigma = checkpoint.parameter_restore(self.engine.spec, checkpoint.load_file(path))
        self.engine.set_params(mu, sigma, also_target=True)  # call_restore loads q_online and q_target (model_torch.py:47-49)

    def memory_backup(self, item_compress: bool = False) -> list:
        from . import checkpoint

        eng = self.engine
        seed, A = int(eng.cfg.seed) & 0xFFFFFFFFFFFFFFFF, eng.A
        return checkpoint.memory_backup(eng.ring_view(), bool(eng.per), compress=item_compress,
                                        pad_action=lambda e, g: checkpoint.philox_pad_action(seed, e, g, A))

    def memory_restore(self, data: list) -> None:
        from . import checkpoint

        eng = self.engine
        eng.load_ring(checkpoint.memory_restore(data, eng.E, eng.R, eng.M, eng.A, eng.D, bool(eng.per)))

    def save_memory(self, path: str, compress: bool = True, item_compress: bool = False) -> None:
        """item_compress = the reference's memory.compress (items stored as zlib(pickle(item)), default True there)."""
        from . import checkpoint

        checkpoint.save_file(path, self.memory_backup(item_compress), compress)

    def load_memory(self, path: str) -> None:
        from . import checkpoint

        self.memory_restore(checkpoint.load_file(path))
