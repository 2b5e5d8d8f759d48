// envs.cuh -- closed-form environments stepped on device.  CPU twin: oracle/envs.py (bit-exact, see its header).
//   Grid      srl/envs/grid.py:88-208,340-378 (+ EnvRun truncation srl/base/env/env_run.py:360-362)
//   CartPole  gymnasium==1.2.0 classic_control/cartpole.py restated (third-party; parity vs gymnasium unpinned)
//   Pendulum  gymnasium==1.2.0 classic_control/pendulum.py restated (third-party; parity vs gymnasium unpinned), driven by
//             the reference's discretised action table (BoxSpace.create_division_tbl, srl/base/spaces/box.py:317-366)
#pragma once
#include "philox.cuh"

namespace srlx {

// ---- Grid ------------------------------------------------------------------------------------------------------
// Action enum grid.py:81-85: LEFT=0, DOWN=1, RIGHT=2, UP=3
__device__ inline void grid_move(const srlx_engine& eng, int x, int y, int a, int& nx, int& ny) {
  nx = x;
  ny = y;
  if (a == 3) ny -= 1;
  else if (a == 1) ny += 1;
  else if (a == 0) nx -= 1;
  else if (a == 2) nx += 1;
  if (!(0 <= nx && nx < eng.grid_w)) { nx = x; ny = y; }
  if (!(0 <= ny && ny < eng.grid_h)) { nx = x; ny = y; }
  if (eng.grid_field[ny * eng.grid_w + nx] == 9) { nx = x; ny = y; }
}

__device__ inline void grid_reset(const srlx_engine& eng, uint32_t e, uint32_t episode, double* st) {
  uint4 w = philox(eng.seed, STREAM_ENV_RESET, e, episode, 0);
  int k = (int)u_below(w.x, (uint32_t)eng.grid_n_starts);
  int s = eng.grid_starts[k];
  st[0] = (double)(s & 0xFF);
  st[1] = (double)((s >> 8) & 0xFF);
  st[2] = 0.0;
  st[3] = 0.0;
}

// returns reward; sets terminated
__device__ inline double grid_step(const srlx_engine& eng, uint32_t e, uint64_t g, int action, double* st, bool& terminated) {
  uint4 w = philox(eng.seed, STREAM_ENV_STEP, e, (uint32_t)g, (uint32_t)(g >> 32));
  double u = u01_f64(w.x, w.y);
  // np.random.choice(4, p): cdf.searchsorted(u, side='right') == number of cdf entries <= u  (grid.py:200-203)
  const double* cdf = &eng.grid_slip_cdf[action * 4];
  int k = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) k += (cdf[j] <= u) ? 1 : 0;
  if (k > 3) k = 3;
  int a = eng.grid_slip_action[k];
  int nx, ny;
  grid_move(eng, (int)st[0], (int)st[1], a, nx, ny);
  st[0] = (double)nx;
  st[1] = (double)ny;
  int attr = eng.grid_field[ny * eng.grid_w + nx];
  terminated = false;
  double r = eng.grid_move_reward;
  if (attr == 1) { r = eng.grid_goal_reward; terminated = true; }
  else if (attr == -1) { r = eng.grid_hole_reward; terminated = true; }
  return r;
}

// ---- CartPole-v1 -------------------------------------------------------------------------------------------------
// Every operation is an explicitly rounded IEEE double op (no fma contraction) in the exact order of oracle/envs.py.
__device__ inline double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ inline double da(double a, double b) { return __dadd_rn(a, b); }
__device__ inline double ds(double a, double b) { return __dsub_rn(a, b); }
__device__ inline double dd(double a, double b) { return __ddiv_rn(a, b); }

__device__ inline double poly_sin(double x) {
  const double z = dm(x, x);
  double p = 1.0 / 355687428096000.0;
  p = da(dm(p, z), -1.0 / 1307674368000.0);
  p = da(dm(p, z), 1.0 / 6227020800.0);
  p = da(dm(p, z), -1.0 / 39916800.0);
  p = da(dm(p, z), 1.0 / 362880.0);
  p = da(dm(p, z), -1.0 / 5040.0);
  p = da(dm(p, z), 1.0 / 120.0);
  p = da(dm(p, z), -1.0 / 6.0);
  p = da(dm(p, z), 1.0);
  return dm(p, x);
}
__device__ inline double poly_cos(double x) {
  const double z = dm(x, x);
  double p = 1.0 / 20922789888000.0;
  p = da(dm(p, z), -1.0 / 87178291200.0);
  p = da(dm(p, z), 1.0 / 479001600.0);
  p = da(dm(p, z), -1.0 / 3628800.0);
  p = da(dm(p, z), 1.0 / 40320.0);
  p = da(dm(p, z), -1.0 / 720.0);
  p = da(dm(p, z), 1.0 / 24.0);
  p = da(dm(p, z), -1.0 / 2.0);
  p = da(dm(p, z), 1.0);
  return p;
}

__device__ inline void cartpole_reset(const srlx_engine& eng, uint32_t e, uint32_t episode, double* st) {
  uint4 w = philox(eng.seed, STREAM_ENV_RESET, e, episode, 0);
  uint4 v = philox(eng.seed, STREAM_ENV_RESET, e, episode, 1);
  st[0] = da(-0.05, dm(0.1, u01_f64(w.x, w.y)));
  st[1] = da(-0.05, dm(0.1, u01_f64(w.z, w.w)));
  st[2] = da(-0.05, dm(0.1, u01_f64(v.x, v.y)));
  st[3] = da(-0.05, dm(0.1, u01_f64(v.z, v.w)));
}

__device__ inline double cartpole_step(int action, double* st, bool& terminated) {
  const double gravity = 9.8, masspole = 0.1, total_mass = 1.1, length = 0.5, polemass_length = 0.05, tau = 0.02;
  const double x_threshold = 2.4, theta_threshold = 0.20943951023931953;  // 12 * 2 * pi / 360
  double x = st[0], x_dot = st[1], theta = st[2], theta_dot = st[3];
  const double force = (action == 1) ? 10.0 : -10.0;
  const double costheta = poly_cos(theta);
  const double sintheta = poly_sin(theta);
  const double temp = dd(da(force, dm(dm(polemass_length, dm(theta_dot, theta_dot)), sintheta)), total_mass);
  const double denom = dm(length, ds(4.0 / 3.0, dd(dm(masspole, dm(costheta, costheta)), total_mass)));
  const double thetaacc = dd(ds(dm(gravity, sintheta), dm(costheta, temp)), denom);
  const double xacc = ds(temp, dd(dm(dm(polemass_length, thetaacc), costheta), total_mass));
  x = da(x, dm(tau, x_dot));
  x_dot = da(x_dot, dm(tau, xacc));
  theta = da(theta, dm(tau, theta_dot));
  theta_dot = da(theta_dot, dm(tau, thetaacc));
  st[0] = x; st[1] = x_dot; st[2] = theta; st[3] = theta_dot;
  terminated = (x < -x_threshold) || (x > x_threshold) || (theta < -theta_threshold) || (theta > theta_threshold);
  return 1.0;
}

// ---- Pendulum-v1 -------------------------------------------------------------------------------------------------
// state (theta, theta_dot) in fp64, theta unwrapped as in gymnasium; every operation an explicitly rounded IEEE double op
// in the order of oracle/envs.py::PendulumSpec.  angle_normalize(x) = ((x + pi) mod 2 pi) - pi is evaluated as
// y - floor(y / 2pi) * 2pi - pi with y = x + pi; sin / cos of the normalised angle fold to [-pi/2, pi/2] and use Taylor
// polynomials to x^23 / x^22 (error < 1e-17 there).
__device__ inline double pend_angle_normalize(double x) {
  const double pi = 3.141592653589793, two_pi = 6.283185307179586, inv_two_pi = 0.15915494309189535;
  const double y = da(x, pi);
  const double k = floor(dm(y, inv_two_pi));
  return ds(ds(y, dm(k, two_pi)), pi);
}
__device__ inline void pend_sincos(double an, double& s_out, double& c_out) {  // an in [-pi, pi]
  const double pi = 3.141592653589793, half_pi = 1.5707963267948966;
  double r = an, csign = 1.0;
  if (an > half_pi) { r = ds(pi, an); csign = -1.0; }
  else if (an < -half_pi) { r = ds(-pi, an); csign = -1.0; }
  const double z = dm(r, r);
  double ps = -1.0 / 25852016738884976640000.0;  // -1/23!
  ps = da(dm(ps, z), 1.0 / 51090942171709440000.0);   // 1/21!
  ps = da(dm(ps, z), -1.0 / 121645100408832000.0);    // -1/19!
  ps = da(dm(ps, z), 1.0 / 355687428096000.0);
  ps = da(dm(ps, z), -1.0 / 1307674368000.0);
  ps = da(dm(ps, z), 1.0 / 6227020800.0);
  ps = da(dm(ps, z), -1.0 / 39916800.0);
  ps = da(dm(ps, z), 1.0 / 362880.0);
  ps = da(dm(ps, z), -1.0 / 5040.0);
  ps = da(dm(ps, z), 1.0 / 120.0);
  ps = da(dm(ps, z), -1.0 / 6.0);
  ps = da(dm(ps, z), 1.0);
  s_out = dm(ps, r);
  double pc = -1.0 / 1124000727777607680000.0;        // -1/22!
  pc = da(dm(pc, z), 1.0 / 2432902008176640000.0);     // 1/20!
  pc = da(dm(pc, z), -1.0 / 6402373705728000.0);       // -1/18!
  pc = da(dm(pc, z), 1.0 / 20922789888000.0);
  pc = da(dm(pc, z), -1.0 / 87178291200.0);
  pc = da(dm(pc, z), 1.0 / 479001600.0);
  pc = da(dm(pc, z), -1.0 / 3628800.0);
  pc = da(dm(pc, z), 1.0 / 40320.0);
  pc = da(dm(pc, z), -1.0 / 720.0);
  pc = da(dm(pc, z), 1.0 / 24.0);
  pc = da(dm(pc, z), -1.0 / 2.0);
  pc = da(dm(pc, z), 1.0);
  c_out = dm(csign, pc);
}
__device__ inline void pendulum_reset(const srlx_engine& eng, uint32_t e, uint32_t episode, double* st) {
  uint4 w = philox(eng.seed, STREAM_ENV_RESET, e, episode, 0);
  // np_random.uniform(low=[-pi, -1], high=[pi, 1]): low + (high - low) * u
  st[0] = da(-3.141592653589793, dm(6.283185307179586, u01_f64(w.x, w.y)));
  st[1] = da(-1.0, dm(2.0, u01_f64(w.z, w.w)));
  st[2] = 0.0;
  st[3] = 0.0;
}
// continuous torque u (already inside [-max_torque, max_torque]): the policy-gradient algorithms act on the Box directly
__device__ inline double pendulum_step_torque(const double u, double* st, bool& terminated) {
  const double dt = 0.05, max_speed = 8.0;
  const double th = st[0], thdot = st[1];
  const double an = pend_angle_normalize(th);
  double sn, cs;
  pend_sincos(an, sn, cs);
  const double costs = da(da(dm(an, an), dm(0.1, dm(thdot, thdot))), dm(0.001, dm(u, u)));
  double newthdot = da(thdot, dm(da(dm(15.0, sn), dm(3.0, u)), dt));  // 3g/(2l) = 15, 3/(m l^2) = 3
  newthdot = newthdot < -max_speed ? -max_speed : (newthdot > max_speed ? max_speed : newthdot);
  st[0] = da(th, dm(newthdot, dt));
  st[1] = newthdot;
  terminated = false;  // Pendulum only ever ends by the 200-step TimeLimit
  return -costs;
}
__device__ inline double pendulum_step(const srlx_engine& eng, int action, double* st, bool& terminated) {
  return pendulum_step_torque(eng.act_tbl[action], st, terminated);  // the reference's discretised torques (value-based algorithms)
}

// ---- dispatch ---------------------------------------------------------------------------------------------------
__device__ inline void env_reset(const srlx_engine& eng, uint32_t e, uint32_t episode, double* st) {
  if (eng.env_id == SRLX_ENV_GRID) grid_reset(eng, e, episode, st);
  else if (eng.env_id == SRLX_ENV_PENDULUM) pendulum_reset(eng, e, episode, st);
  else cartpole_reset(eng, e, episode, st);
}
__device__ inline double env_step(const srlx_engine& eng, uint32_t e, uint64_t g, int action, double* st, bool& terminated) {
  if (eng.env_id == SRLX_ENV_GRID) return grid_step(eng, e, g, action, st, terminated);
  if (eng.env_id == SRLX_ENV_PENDULUM) return pendulum_step(eng, action, st, terminated);
  return cartpole_step(action, st, terminated);
}
// observation as the RL side sees it: float32 cast (BoxSpace encode srl/base/spaces/box.py:585-598)
__device__ inline void env_obs(const srlx_engine& eng, const double* st, float* obs) {
  if (eng.env_id == SRLX_ENV_PENDULUM) {  // [cos(theta), sin(theta), theta_dot]
    double sn, cs;
    pend_sincos(pend_angle_normalize(st[0]), sn, cs);
    obs[0] = (float)cs;
    obs[1] = (float)sn;
    obs[2] = (float)st[1];
    return;
  }
  for (int d = 0; d < eng.obs_dim; ++d) obs[d] = (float)st[d];
}

}  // namespace srlx
