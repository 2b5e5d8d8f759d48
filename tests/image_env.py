"""A small image environment for the reference's Runner (test helper; needs `srl` importable): the agent is a bright block on a
W x H uint8 RGB frame, actions move it left / right / up / down, the goal is the top-right corner (reward +1), every step costs
0.01, episodes end at the goal or after 40 steps.  Registered as "PixelGrid-b200"."""
import numpy as np

from srl.base.define import SpaceTypes
from srl.base.env import registration
from srl.base.env.base import EnvBase
from srl.base.spaces.box import BoxSpace
from srl.base.spaces.discrete import DiscreteSpace

CELL = 6


class PixelGrid(EnvBase):
    def __init__(self, nx: int = 5, ny: int = 4, **kwargs):
        super().__init__()
        self.nx, self.ny = nx, ny
        self.pos = (0, ny - 1)

    @property
    def action_space(self):
        return DiscreteSpace(4)

    @property
    def observation_space(self):
        return BoxSpace((self.ny * CELL, self.nx * CELL, 3), 0, 255, np.uint8, SpaceTypes.RGB)

    @property
    def max_episode_steps(self) -> int:
        return 40

    @property
    def player_num(self) -> int:
        return 1

    def _frame(self):
        f = np.zeros((self.ny * CELL, self.nx * CELL, 3), np.uint8)
        f[..., 2] = 40
        f[0:CELL, (self.nx - 1) * CELL:, 1] = 200  # the goal cell
        x, y = self.pos
        f[y * CELL:(y + 1) * CELL, x * CELL:(x + 1) * CELL, 0] = 255
        f[y * CELL:(y + 1) * CELL, x * CELL:(x + 1) * CELL, 1] = 128
        return f

    def reset(self, *, seed=None, **kwargs):
        self.pos = (0, self.ny - 1)
        return self._frame()

    def step(self, action):
        x, y = self.pos
        dx, dy = [(-1, 0), (1, 0), (0, -1), (0, 1)][int(action)]
        self.pos = (min(max(x + dx, 0), self.nx - 1), min(max(y + dy, 0), self.ny - 1))
        done = self.pos == (self.nx - 1, 0)
        return self._frame(), (1.0 if done else -0.01), done, False

    def backup(self):
        return self.pos

    def restore(self, data):
        self.pos = tuple(data)


def register():
    registration.register("PixelGrid-b200", entry_point=__name__ + ":PixelGrid", check_duplicate=False)
