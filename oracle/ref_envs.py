"""TEST / BASELINE INFRASTRUCTURE (not part of the product): CPU `EnvBase` classes for the reference's own Runner.

gymnasium is not installed next to the reference, so `srl.Runner("CartPole-v1", ...)` cannot run there.  These classes
wrap the oracle's restatements (oracle/envs.py, parity vs gymnasium 1.2.0 unpinned) behind the reference's env plug-in
interface (srl/base/env/base.py:60-137) and register them under the gymnasium ids -- a registered id shadows gymnasium
(srl/base/env/registration.py:53-62,116-136) -- so the UNMODIFIED reference loop (core_play.play) can be timed and probed
on the same environments the device steps.  Needs the reference importable (`srl`).
"""
import numpy as np

from srl.base.env.base import EnvBase  # the reference
from srl.base.env.registration import register
from srl.base.spaces.box import BoxSpace
from srl.base.spaces.discrete import DiscreteSpace

from . import envs as oenvs


class _Restated(EnvBase):
    spec_name = ""

    def __init__(self, **kw):
        super().__init__()
        self.spec = oenvs.make_spec(self.spec_name, **kw)
        self._rng = np.random.default_rng(0)
        self._st = None
        self._t = 0

    @property
    def player_num(self) -> int:
        return 1

    @property
    def max_episode_steps(self) -> int:
        return int(self.spec.trunc_limit)

    def _uniforms(self, n):
        return self._rng.random(n)

    def backup(self):
        return [np.array(self._st, copy=True), self._t]

    def restore(self, data):
        self._st, self._t = np.array(data[0], copy=True), data[1]


class CartPoleRestated(_Restated):
    """gymnasium CartPole-v1 as restated in oracle/envs.py::CartPoleSpec (obs Box(4,) float32, Discrete(2), 500 steps)."""
    spec_name = "CartPole-v1"

    @property
    def action_space(self):
        return DiscreteSpace(2)

    @property
    def observation_space(self):
        high = np.array([4.8, np.finfo(np.float32).max, 0.41887903, np.finfo(np.float32).max], dtype=np.float32)
        return BoxSpace((4,), -high, high, np.float32)

    def reset(self, *, seed=None, **kwargs):
        if seed is not None:
            self._rng = np.random.default_rng(seed)
        self._st = np.float64(-0.05) + np.float64(0.1) * self._uniforms(4)
        self._t = 0
        return self.spec.obs(self._st)

    def step(self, action):
        self._st, r, terminated = self.spec.step(self._st, int(action))
        self._t += 1
        truncated = self._t >= self.spec.trunc_limit  # gymnasium TimeLimit
        return self.spec.obs(self._st), float(r), bool(terminated), bool(truncated)


class PendulumRestated(_Restated):
    """gymnasium Pendulum-v1 as restated in oracle/envs.py::PendulumSpec; the action is the continuous Box(1,) in [-2, 2]
    (the value-based workers of the reference discretise it themselves: RLConfig.action_division_num)."""
    spec_name = "Pendulum-v1"

    @property
    def action_space(self):
        return BoxSpace((1,), -2.0, 2.0, np.float32)

    @property
    def observation_space(self):
        high = np.array([1.0, 1.0, 8.0], dtype=np.float32)
        return BoxSpace((3,), -high, high, np.float32)

    @property
    def reward_baseline(self):
        return {"episode": 10, "baseline": -500}

    def reset(self, *, seed=None, **kwargs):
        if seed is not None:
            self._rng = np.random.default_rng(seed)
        u = self._uniforms(2)
        self._st = np.array([-np.pi + 2 * np.pi * u[0], -1.0 + 2.0 * u[1], 0.0, 0.0], dtype=np.float64)
        self._t = 0
        return self.spec.obs(self._st)

    def step(self, action):
        u = float(np.clip(np.asarray(action, dtype=np.float64).reshape(-1)[0], -2.0, 2.0))
        # the oracle spec steps a table index; here the continuous value comes from the reference's own decode
        tbl = self.spec.action_table
        self.spec.action_table = np.array([u], dtype=np.float64)
        try:
            self._st, r, _ = self.spec.step(self._st, 0)
        finally:
            self.spec.action_table = tbl
        self._t += 1
        return self.spec.obs(self._st), float(r), False, bool(self._t >= self.spec.trunc_limit)


def register_restated_envs():
    register("CartPole-v1", "oracle.ref_envs:CartPoleRestated", check_duplicate=False)
    register("Pendulum-v1", "oracle.ref_envs:PendulumRestated", check_duplicate=False)
