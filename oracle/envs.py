"""CPU restatement of the closed-form environments on the hot path (TEST INFRASTRUCTURE).

Grid      follows srl/envs/grid.py:88-208,340-378 of the reference (field :100-108, slip table :121-146,
          step :188-208, _move :340-364, reward_done_func :366-378) and the EnvRun truncation rule
          srl/base/env/env_run.py:360-362 (``step_num > max_episode_steps`` with max_episode_steps=50, grid.py:160-161).
          Pinned against the reference env itself: tests/golden/grid_transitions.npz (see make_golden.py).
CartPole  restates third-party gymnasium==1.2.0 ``envs/classic_control/cartpole.py`` (dockers/latest_requirements.txt:10
          pins the version; the package is neither vendored nor installed) plus its TimeLimit(500) wrapper and the SRL
          wrapper semantics srl/base/env/gymnasium_wrapper.py:334-374.  **Parity vs gymnasium is UNPINNED**: no copy of
          gymnasium exists on this machine and no reference test pins a CartPole transition
          (tests/quick/base/env/test_gymnasium_wrapper.py:30-43 only pins the spaces).
          sin/cos use the fixed polynomial below (plain IEEE mul/add, no fma) so that the CUDA kernel
          (csrc/envs.cuh, __dmul_rn/__dadd_rn) reproduces the float64 state BIT-EXACTLY.

All randomness is drawn from Philox counters (oracle/philox.py) keyed exactly as the device does.
"""
import numpy as np

from . import philox

# --------------------------------------------------------------------------------------------------------
# Grid
# --------------------------------------------------------------------------------------------------------
GRID_FIELD = [
    [9, 9, 9, 9, 9, 9],
    [9, 0, 0, 0, 1, 9],
    [9, 0, 9, 0, -1, 9],
    [9, 2, 0, 0, 0, 9],
    [9, 9, 9, 9, 9, 9],
]
# Action enum, grid.py:81-85
LEFT, DOWN, RIGHT, UP = 0, 1, 2, 3
# np.random.choice sees the executed-action candidates in dict order UP, DOWN, RIGHT, LEFT (grid.py:121-146, :200-203)
GRID_SLIP_ORDER = [UP, DOWN, RIGHT, LEFT]


def grid_slip_cdf(move_prob=0.8):
    """cdf rows per chosen action, built the way np.random.choice does (cumsum, then /= cdf[-1])."""
    side = (1 - move_prob) / 2
    table = {
        UP: [move_prob, 0, side, side],
        DOWN: [0, move_prob, side, side],
        RIGHT: [side, side, move_prob, 0],
        LEFT: [side, side, 0, move_prob],
    }
    cdf = np.zeros((4, 4), dtype=np.float64)
    for a in range(4):
        p = np.array(table[a], dtype=np.float64)
        c = p.cumsum()
        c /= c[-1]
        cdf[a] = c
    return cdf


class GridSpec:
    env_id = 0
    obs_dim = 2
    n_actions = 4
    trunc_limit = 51  # EnvRun: step_num > 50
    trunc_overrides_term = 0

    def __init__(self, move_reward=-0.04, move_prob=0.8, goal_reward=1.0, hole_reward=-1.0, field=None):
        self.field = np.array(GRID_FIELD if field is None else field, dtype=np.int8)
        self.h, self.w = self.field.shape
        self.move_reward = float(move_reward)
        self.goal_reward = float(goal_reward)
        self.hole_reward = float(hole_reward)
        self.cdf = grid_slip_cdf(move_prob)
        self.starts = [(x, y) for y in range(self.h) for x in range(self.w) if self.field[y, x] == 2]

    # grid.py:340-364
    def move(self, x, y, a):
        nx, ny = x, y
        if a == UP:
            ny -= 1
        elif a == DOWN:
            ny += 1
        elif a == LEFT:
            nx -= 1
        elif a == RIGHT:
            nx += 1
        if not (0 <= nx < self.w):
            nx, ny = x, y
        if not (0 <= ny < self.h):
            nx, ny = x, y
        if self.field[ny, nx] == 9:
            nx, ny = x, y
        return nx, ny

    # grid.py:366-378
    def reward_done(self, x, y):
        attr = self.field[y, x]
        if attr == 1:
            return self.goal_reward, True
        if attr == -1:
            return self.hole_reward, True
        return self.move_reward, False

    def slip(self, action, u):
        """executed action for uniform u in [0,1): cdf.searchsorted(u, side='right') (np.random.choice)."""
        k = int(np.searchsorted(self.cdf[action], u, side="right"))
        k = min(k, 3)
        return GRID_SLIP_ORDER[k]

    # ---- vector-engine hooks -----------------------------------------------------------------------
    def reset(self, seed, e, episode):
        w0, _, _, _ = philox.words(seed, philox.STREAM_ENV_RESET, e, episode, 0)
        k = (int(w0) * len(self.starts)) >> 32
        x, y = self.starts[k]
        return np.array([x, y, 0, 0], dtype=np.float64)

    def obs(self, st):
        return np.array([st[0], st[1]], dtype=np.float32)

    def step(self, st, action, seed, e, g):
        w0, w1, _, _ = philox.words(seed, philox.STREAM_ENV_STEP, e, g & 0xFFFFFFFF, g >> 32)
        u = float(philox.u01_f64(w0, w1))
        a = self.slip(action, u)
        nx, ny = self.move(int(st[0]), int(st[1]), a)
        r, done = self.reward_done(nx, ny)
        return np.array([nx, ny, 0, 0], dtype=np.float64), r, done


# --------------------------------------------------------------------------------------------------------
# CartPole-v1 (gymnasium 1.2.0 restatement; parity vs gymnasium unpinned)
# --------------------------------------------------------------------------------------------------------
_SIN_C = [  # Taylor coefficients of sin(x)/x in x^2, highest first: x^16 ... x^0
    1.0 / 355687428096000.0,
    -1.0 / 1307674368000.0,
    1.0 / 6227020800.0,
    -1.0 / 39916800.0,
    1.0 / 362880.0,
    -1.0 / 5040.0,
    1.0 / 120.0,
    -1.0 / 6.0,
    1.0,
]
_COS_C = [  # Taylor coefficients of cos(x) in x^2, highest first: x^16 ... x^0
    1.0 / 20922789888000.0,
    -1.0 / 87178291200.0,
    1.0 / 479001600.0,
    -1.0 / 3628800.0,
    1.0 / 40320.0,
    -1.0 / 720.0,
    1.0 / 24.0,
    -1.0 / 2.0,
    1.0,
]


def poly_sin(x):
    """sin(x) for |x| <~ 1 (CartPole keeps |theta| < 0.3): Horner in x^2, plain mul/add (no fma)."""
    x = np.float64(x)
    z = x * x
    p = np.float64(_SIN_C[0])
    for c in _SIN_C[1:]:
        p = p * z + np.float64(c)
    return p * x


def poly_cos(x):
    x = np.float64(x)
    z = x * x
    p = np.float64(_COS_C[0])
    for c in _COS_C[1:]:
        p = p * z + np.float64(c)
    return p


class CartPoleSpec:
    env_id = 1
    obs_dim = 4
    n_actions = 2
    trunc_limit = 500  # gymnasium TimeLimit(max_episode_steps=500): truncated = elapsed_steps >= 500
    trunc_overrides_term = 1  # EnvRun._step1: `if truncated: TRUNCATED elif terminated: TERMINATED` (env_run.py:327-332)

    gravity = 9.8
    masscart = 1.0
    masspole = 0.1
    total_mass = masspole + masscart
    length = 0.5
    polemass_length = masspole * length
    force_mag = 10.0
    tau = 0.02
    theta_threshold = 12 * 2 * np.pi / 360
    x_threshold = 2.4

    def reset(self, seed, e, episode):
        w = philox.words(seed, philox.STREAM_ENV_RESET, e, episode, 0)
        v = philox.words(seed, philox.STREAM_ENV_RESET, e, episode, 1)
        u = np.array(
            [philox.u01_f64(w[0], w[1]), philox.u01_f64(w[2], w[3]), philox.u01_f64(v[0], v[1]), philox.u01_f64(v[2], v[3])],
            dtype=np.float64,
        )
        # np_random.uniform(low=-0.05, high=0.05): low + (high-low)*u
        return np.float64(-0.05) + np.float64(0.1) * u

    def obs(self, st):
        return st.astype(np.float32)

    def step(self, st, action, seed=None, e=None, g=None):
        x, x_dot, theta, theta_dot = [np.float64(v) for v in st]
        force = np.float64(self.force_mag if action == 1 else -self.force_mag)
        costheta = poly_cos(theta)
        sintheta = poly_sin(theta)
        pml = np.float64(self.polemass_length)
        tm = np.float64(self.total_mass)
        temp = (force + (pml * (theta_dot * theta_dot)) * sintheta) / tm
        denom = np.float64(self.length) * (np.float64(4.0 / 3.0) - (np.float64(self.masspole) * (costheta * costheta)) / tm)
        thetaacc = (np.float64(self.gravity) * sintheta - costheta * temp) / denom
        xacc = temp - ((pml * thetaacc) * costheta) / tm
        tau = np.float64(self.tau)
        x = x + tau * x_dot
        x_dot = x_dot + tau * xacc
        theta = theta + tau * theta_dot
        theta_dot = theta_dot + tau * thetaacc
        terminated = bool(x < -self.x_threshold or x > self.x_threshold or theta < -self.theta_threshold or theta > self.theta_threshold)
        return np.array([x, x_dot, theta, theta_dot], dtype=np.float64), 1.0, terminated


# --------------------------------------------------------------------------------------------------------
# Pendulum-v1 (gymnasium 1.2.0 restatement; parity vs gymnasium unpinned) with the reference's discretised actions
# --------------------------------------------------------------------------------------------------------
_PSIN_C = [-1.0 / 25852016738884976640000.0, 1.0 / 51090942171709440000.0, -1.0 / 121645100408832000.0,
           1.0 / 355687428096000.0, -1.0 / 1307674368000.0, 1.0 / 6227020800.0, -1.0 / 39916800.0, 1.0 / 362880.0,
           -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0, 1.0]
_PCOS_C = [-1.0 / 1124000727777607680000.0, 1.0 / 2432902008176640000.0, -1.0 / 6402373705728000.0,
           1.0 / 20922789888000.0, -1.0 / 87178291200.0, 1.0 / 479001600.0, -1.0 / 3628800.0, 1.0 / 40320.0,
           -1.0 / 720.0, 1.0 / 24.0, -1.0 / 2.0, 1.0]
_PI, _TWO_PI, _INV_TWO_PI, _HALF_PI = np.float64(3.141592653589793), np.float64(6.283185307179586), np.float64(0.15915494309189535), np.float64(1.5707963267948966)


def pend_angle_normalize(x):
    """((x + pi) mod 2 pi) - pi as y - floor(y / 2pi) * 2pi - pi (device twin: csrc/envs.cuh::pend_angle_normalize)."""
    y = np.float64(x) + _PI
    k = np.floor(y * _INV_TWO_PI)
    return (y - k * _TWO_PI) - _PI


def pend_sincos(an):
    an = np.float64(an)
    r, csign = an, np.float64(1.0)
    if an > _HALF_PI:
        r, csign = _PI - an, np.float64(-1.0)
    elif an < -_HALF_PI:
        r, csign = -_PI - an, np.float64(-1.0)
    z = r * r
    ps = np.float64(_PSIN_C[0])
    for c in _PSIN_C[1:]:
        ps = ps * z + np.float64(c)
    pc = np.float64(_PCOS_C[0])
    for c in _PCOS_C[1:]:
        pc = pc * z + np.float64(c)
    return ps * r, csign * pc


def division_table(low, high, n):
    """BoxSpace.create_division_tbl, 1-D float32 Box (srl/base/spaces/box.py:340-365)."""
    lo, hi = np.float32(low), np.float32(high)
    diff = (hi - lo) / (n - 1)
    return np.array([np.float32(lo + diff * j) for j in range(n)], dtype=np.float32)


class PendulumSpec:
    env_id = 2
    obs_dim = 3
    trunc_limit = 200  # gymnasium TimeLimit(max_episode_steps=200)
    trunc_overrides_term = 1
    max_speed, max_torque, dt = 8.0, 2.0, 0.05

    def __init__(self, action_division_num=10):
        self.n_actions = int(action_division_num)
        self.action_table = division_table(-self.max_torque, self.max_torque, self.n_actions).astype(np.float64)

    def reset(self, seed, e, episode):
        w = philox.words(seed, philox.STREAM_ENV_RESET, e, episode, 0)
        th = -_PI + _TWO_PI * np.float64(philox.u01_f64(w[0], w[1]))
        thdot = np.float64(-1.0) + np.float64(2.0) * np.float64(philox.u01_f64(w[2], w[3]))
        return np.array([th, thdot, 0.0, 0.0], dtype=np.float64)

    def obs(self, st):
        sn, cs = pend_sincos(pend_angle_normalize(st[0]))
        return np.array([cs, sn, st[1]], dtype=np.float64).astype(np.float32)

    def step(self, st, action, seed=None, e=None, g=None):
        return self.step_torque(st, np.float64(self.action_table[action]))

    def step_torque(self, st, u):
        """One step with a continuous torque u (already inside [-max_torque, max_torque]): what the policy-gradient algorithms send."""
        th, thdot = np.float64(st[0]), np.float64(st[1])
        u = np.float64(u)
        an = pend_angle_normalize(th)
        sn, _ = pend_sincos(an)
        costs = (an * an + np.float64(0.1) * (thdot * thdot)) + np.float64(0.001) * (u * u)
        newthdot = thdot + (np.float64(15.0) * sn + np.float64(3.0) * u) * np.float64(self.dt)
        newthdot = np.float64(min(max(newthdot, -self.max_speed), self.max_speed))
        newth = th + newthdot * np.float64(self.dt)
        return np.array([newth, newthdot, 0.0, 0.0], dtype=np.float64), float(-costs), False


class ExternalSpec:
    """An env stepped by a host loop (the reference's core_play.play through the plug-in classes): only the shapes matter."""

    trunc_limit, trunc_overrides_term = 2**31 - 1, 0

    def __init__(self, obs_dim, n_actions):
        self.obs_dim, self.n_actions = int(obs_dim), int(n_actions)


def make_spec(env_id, **kw):
    if env_id in (3, "external"):
        return ExternalSpec(**kw)
    if env_id in (0, "Grid", "grid"):
        return GridSpec(**kw)
    if env_id in (1, "CartPole-v1", "cartpole"):
        return CartPoleSpec()
    if env_id in (2, "Pendulum-v1", "pendulum"):
        return PendulumSpec(**kw)
    raise ValueError(env_id)
