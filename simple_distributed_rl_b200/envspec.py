"""Tables of the closed-form environments the device steps (filled into `srlx_engine`).

Grid      srl/envs/grid.py:21-30 (registration kwargs), :88-161 (field, slip table, max_episode_steps=50)
CartPole  gymnasium==1.2.0 CartPole-v1 restated (obs Box(4,) float32, Discrete(2), TimeLimit 500);
          srl/base/env/gymnasium_wrapper.py:290-374 is the reference-side wrapper.
Pendulum  gymnasium==1.2.0 Pendulum-v1 restated (obs Box(3,) float32, action Box(1,) in [-2, 2], TimeLimit 200); value-based
          algorithms see the reference's discretisation of the action Box: RLConfig.action_division_num (default 10,
          srl/base/rl/config.py:55) evenly spaced float32 values, BoxSpace.create_division_tbl (srl/base/spaces/box.py:317-366).
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import _lib

GRID_FIELD = [
    [9, 9, 9, 9, 9, 9],
    [9, 0, 0, 0, 1, 9],
    [9, 0, 9, 0, -1, 9],
    [9, 2, 0, 0, 0, 9],
    [9, 9, 9, 9, 9, 9],
]
LEFT, DOWN, RIGHT, UP = 0, 1, 2, 3


@dataclass
class EnvSpec:
    name: str
    env_id: int
    obs_dim: int
    n_actions: int
    trunc_limit: int
    trunc_overrides_term: int
    max_episode_steps: int
    reward_baseline: dict
    # grid only
    field: List[List[int]] = field(default_factory=list)
    move_prob: float = 0.8
    move_reward: float = -0.04
    goal_reward: float = 1.0
    hole_reward: float = -1.0
    obs_low: tuple = ()
    obs_high: tuple = ()
    action_table: tuple = ()  # continuous value of each discrete action index (Pendulum)

    def fill(self, eng: "_lib.SrlxEngine"):
        eng.env_id, eng.obs_dim, eng.n_actions = self.env_id, self.obs_dim, self.n_actions
        eng.trunc_limit, eng.trunc_overrides_term = self.trunc_limit, self.trunc_overrides_term
        for i, v in enumerate(self.action_table):
            eng.act_tbl[i] = float(v)
        if self.env_id == _lib.ENV_GRID:
            f = np.array(self.field, dtype=np.int8)
            h, w = f.shape
            if h * w > 64:
                raise ValueError("grid field larger than 64 cells")
            eng.grid_w, eng.grid_h = w, h
            for i, v in enumerate(f.reshape(-1)):
                eng.grid_field[i] = int(v)
            starts = [(x, y) for y in range(h) for x in range(w) if f[y, x] == 2]
            if not starts or len(starts) > 16:
                raise ValueError("grid needs 1..16 start cells (value 2)")
            eng.grid_n_starts = len(starts)
            for i, (x, y) in enumerate(starts):
                eng.grid_starts[i] = x | (y << 8)
            side = (1 - self.move_prob) / 2
            table = {UP: [self.move_prob, 0, side, side], DOWN: [0, self.move_prob, side, side],
                     RIGHT: [side, side, self.move_prob, 0], LEFT: [side, side, 0, self.move_prob]}
            for a in range(4):
                c = np.array(table[a], dtype=np.float64).cumsum()
                c /= c[-1]  # np.random.choice normalises the cdf
                for j in range(4):
                    eng.grid_slip_cdf[a * 4 + j] = float(c[j])
            for j, a in enumerate([UP, DOWN, RIGHT, LEFT]):  # dict order of grid.py:121-146
                eng.grid_slip_action[j] = a
            eng.grid_move_reward, eng.grid_goal_reward, eng.grid_hole_reward = self.move_reward, self.goal_reward, self.hole_reward


def make_env_spec(name: str, **kw) -> EnvSpec:
    if name in ("Grid", "EasyGrid"):
        base = dict(move_reward=-0.04, move_prob=0.8, reward_baseline={"episode": 100, "baseline": 0.65})
        if name == "EasyGrid":
            base = dict(move_reward=0.0, move_prob=1.0, reward_baseline={"episode": 100, "baseline": 0.9})
        base.update(kw)
        fld = base.pop("field", GRID_FIELD)
        h, w = len(fld), len(fld[0])
        return EnvSpec(name, _lib.ENV_GRID, 2, 4, 51, 0, 50, field=fld, obs_low=(0, 0), obs_high=(w - 1, h - 1), **base)
    if name == "CartPole-v1":
        return EnvSpec(name, _lib.ENV_CARTPOLE, 4, 2, 500, 1, 500, reward_baseline={"episode": 10, "baseline": 0},
                       obs_low=(-4.8, -np.inf, -0.41887903, -np.inf), obs_high=(4.8, np.inf, 0.41887903, np.inf))
    if name == "Pendulum-v1":
        n = int(kw.get("action_division_num", 10))
        if not 2 <= n <= 16:
            raise ValueError("Pendulum-v1: action_division_num must be in [2, 16]")
        return EnvSpec(name, _lib.ENV_PENDULUM, 3, n, 200, 1, 200, reward_baseline={"episode": 10, "baseline": -500},
                       obs_low=(-1.0, -1.0, -8.0), obs_high=(1.0, 1.0, 8.0), action_table=tuple(division_table(-2.0, 2.0, n)))
    if name == "external":
        # stepped by a HOST loop (the reference's core_play.play through the plug-in classes, srl_classes.py): only the shapes matter
        D, A = int(kw["obs_dim"]), int(kw["n_actions"])
        if not (1 <= D <= 16 and 1 <= A <= 16):  # SRLX_MAX_OBS / SRLX_MAX_ACTIONS; more than 4 floats: the generic learner
            raise NotImplementedError(f"the device learners take <= 16 observation floats and <= 16 discrete actions (got {D}, {A})")
        return EnvSpec(name, _lib.ENV_EXTERNAL, D, A, 2**31 - 1, 0, 0, reward_baseline={})
    raise ValueError(f"environment {name!r} is not available on device (supported: Grid, EasyGrid, CartPole-v1, Pendulum-v1)")


def division_table(low: float, high: float, n: int):
    """BoxSpace.create_division_tbl for a 1-D float32 Box (srl/base/spaces/box.py:340-365): low + diff * j in float32."""
    lo, hi = np.float32(low), np.float32(high)
    diff = (hi - lo) / (n - 1)
    return [np.float32(lo + diff * j) for j in range(n)]
