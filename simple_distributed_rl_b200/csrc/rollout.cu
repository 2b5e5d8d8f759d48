// rollout.cu -- one vector step of E environment copies entirely on device, i.e. the body of the reference's host loop
// srl/base/run/core_play.py:115-214 for E independent (env, worker) pairs:
//   reset-if-done (:138-159) -> observation encode (srl/base/rl/worker_run.py:310-358, srl/base/spaces/box.py:585-598)
//   -> policy: epsilon-greedy / noisy argmax over Q(s)  (srl/algorithms/dqn/dqn.py:192-211, rainbow.py:301-331)
//   -> env.step + done typing + truncation             (srl/base/env/env_run.py:254-366)
//   -> reward shift/scale/clip + record                 (worker_run.py:348, dqn.py:213-246, rainbow.py:333-356)
//   -> coalesced write of (s, a, r, s', terminated, done) into the ring row of this step
// followed by post_step_kernel: replay "add" (srl/rl/memories/priority_replay_buffer.py:205-217,
// proportional_memory.py:120-129) for the whole row + the RunState counters (srl/base/context.py:297-343).
// CPU twin: oracle/engine.py::OracleEngine.vec_step.
#include "envs.cuh"
#include "net.cuh"

namespace srlx {

constexpr int kRolloutThreads = 256;

struct RolloutSmem {
  size_t weff, acts, q, total;
};
__host__ __device__ inline RolloutSmem rollout_smem(const srlx_engine& eng, const NetPlan& pl) {
  RolloutSmem s;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) / 16 * 16; return o; };
  s.weff = take((size_t)pl.weff_floats * 4);
  s.acts = take((size_t)pl.act_floats * 4);
  s.q = take((size_t)kRowTile * eng.n_actions * 4);
  s.total = off;
  return s;
}

__global__ void __launch_bounds__(kRolloutThreads)
rollout_kernel(const __grid_constant__ srlx_engine eng, const int envs_per_cta, const int training) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned long long s_episodes, s_eplen;
  __shared__ double s_epreward;
  const srlx_net& net = eng.net;
  const NetPlan pl = make_plan(net);
  const RolloutSmem so = rollout_smem(eng, pl);
  float* weff = reinterpret_cast<float*>(smem_raw + so.weff);
  float* acts = reinterpret_cast<float*>(smem_raw + so.acts);
  float* q = reinterpret_cast<float*>(smem_raw + so.q);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int E = eng.n_envs, D = eng.obs_dim, A = eng.n_actions, R = eng.ring_rows;
  const uint64_t g = eng.state->vec_steps;
  const int row = (int)(g % (uint64_t)R);
  const bool noisy = net.noisy != 0;
  // Linear.update(step_in_training).to_float() (schedulers/linear.py:16-21); evaluation passes test_epsilon in eng.epsilon
  double eps_d = eng.epsilon;
  if (training && eng.eps_table && eng.eps_table_len)
    eps_d = eng.eps_table[g < eng.eps_table_len ? g : eng.eps_table_len - 1];  // any scheduler, tabulated by the host
  else if (training && eng.eps_phase_steps)
    eps_d = g >= eng.eps_phase_steps ? eng.eps_end : eng.epsilon - ((eng.epsilon - eng.eps_end) / (double)eng.eps_phase_steps) * (double)g;
  const float eps = (float)eps_d;

  if (tid == 0) { s_episodes = 0; s_eplen = 0; s_epreward = 0.0; }
  zero_floats(weff, pl.weff_floats);
  zero_floats(acts, pl.act_floats);
  __syncthreads();
  // one NoisyLinear draw per forward CALL (noisy_linear.py:35-52): the vector step is one call with batch E
  build_weff(net, pl, eng.params, eng.params_sigma, noisy, eng.seed, NOISE_KIND_ROLLOUT, g, weff);
  __syncthreads();

  const int e_begin = blockIdx.x * envs_per_cta;
  const int e_end = min(E, e_begin + envs_per_cta);
  for (int e0 = e_begin; e0 < e_end; e0 += kRowTile) {
    const int Rr = min(kRowTile, e_end - e0);
    // ---- reset-if-done + observation
    if (tid < Rr) {
      const int e = e0 + tid;
      double* st = eng.env_state + (size_t)e * 4;
      if (eng.env_needs_reset[e]) {
        const uint32_t ep = eng.env_episode[e];
        env_reset(eng, (uint32_t)e, ep, st);
        eng.env_episode[e] = ep + 1;
        eng.env_step_num[e] = 0;
        eng.env_ep_reward[e] = 0.0;
        eng.env_needs_reset[e] = 0;
      }
      env_obs(eng, st, acts + pl.x_s[0] + tid * pl.ldx[0]);
    }
    __syncthreads();
    net_forward_tile(net, pl, weff, acts, Rr, q, A);
    // ---- policy, env step, record
    if (tid < Rr) {
      const int e = e0 + tid;
      const float* qe = q + tid * A;
      const uint4 w = philox(eng.seed, STREAM_POLICY, (uint32_t)e, (uint32_t)g, (uint32_t)(g >> 32));
      int action;
      if (!noisy && u01_f32(w.x) < eps) {
        action = (int)u_below(w.y, (uint32_t)A);  // random.choice over the valid actions (dqn.py:200-202)
      } else {
        action = 0;
        float best = qe[0];
        for (int a = 1; a < A; ++a)
          if (qe[a] > best) { best = qe[a]; action = a; }  // np.argmax: first max wins
      }
      double* st = eng.env_state + (size_t)e * 4;
      float obs[SRLX_MAX_OBS], nobs[SRLX_MAX_OBS];
      for (int d = 0; d < D; ++d) obs[d] = acts[pl.x_s[0] + tid * pl.ldx[0] + d];
      bool terminated = false;
      const double r = env_step(eng, (uint32_t)e, g, action, st, terminated);
      const int step_num = eng.env_step_num[e] + 1;
      eng.env_step_num[e] = step_num;
      bool truncated = step_num >= eng.trunc_limit;
      bool term_flag;
      if (eng.trunc_overrides_term) {
        term_flag = terminated && !truncated;  // env_run.py:327-332: `if truncated ... elif terminated`
      } else {
        term_flag = terminated;
        truncated = truncated && !terminated;  // env_run.py:360-362 only applies while done == NONE
      }
      const bool done = terminated || truncated;
      const double ep_reward = eng.env_ep_reward[e] + r;
      eng.env_ep_reward[e] = ep_reward;
      double rr = (r + eng.reward_shift) * eng.reward_scale;  // worker_run.py:348
      if (eng.enable_reward_clip) rr = (rr < 0.0) ? -1.0 : ((rr > 0.0) ? 1.0 : 0.0);
      env_obs(eng, st, nobs);
      if (training) {
        const size_t slot = (size_t)row * E + e;
        for (int d = 0; d < D; ++d) {
          eng.ring_obs[slot * D + d] = obs[d];
          eng.ring_next_obs[slot * D + d] = nobs[d];
        }
        eng.ring_action[slot] = action;
        eng.ring_reward[slot] = (float)rr;
        eng.ring_term[slot] = term_flag ? 1 : 0;
        eng.ring_done[slot] = done ? 1 : 0;
      }
      if (done) {
        eng.env_needs_reset[e] = 1;
        if (eng.env_last_ep_len) {
          if (eng.env_first_ep_reward && eng.env_last_ep_len[e] == 0) eng.env_first_ep_reward[e] = ep_reward;
          eng.env_last_ep_len[e] = step_num;
        }
        atomicAdd(&s_episodes, 1ull);
        atomicAdd(&s_eplen, (unsigned long long)step_num);
        atomicAdd(&s_epreward, ep_reward);
      }
      if (eng.dbg_q)
        for (int a = 0; a < A; ++a) eng.dbg_q[(size_t)e * A + a] = qe[a];
      if (eng.dbg_action) eng.dbg_action[e] = action;
    }
    __syncthreads();
  }
  if (tid == 0 && s_episodes) {
    atomicAdd((unsigned long long*)&eng.state->episode_count, s_episodes);
    atomicAdd((unsigned long long*)&eng.state->episode_len_sum, s_eplen);
    atomicAdd(&eng.state->episode_reward_sum, s_epreward);
  }
}

// leaves [leaf_lo, leaf_lo + n) <- value; every ancestor += the pairwise-summed change of its two children
// (left + right), level by level.  `scratch` holds 2 * (n + 2) doubles.  One thread block.
__device__ inline void tree_set_row(double* __restrict__ tree, int64_t cap, int64_t leaf_lo, int n, double value,
                                    double* __restrict__ scratch) {
  const int tid = threadIdx.x, nt = blockDim.x;
  int64_t a = leaf_lo + cap - 1, b = a + n - 1;
  double* cur = scratch;
  double* nxt = scratch + (n + 2);
  for (int i = tid; i < n; i += nt) {
    const double old = __ldcg(tree + a + i);
    __stcg(cur + i, value - old);
    __stcg(tree + a + i, value);
  }
  __syncthreads();
  // A row of a non-power-of-two tree can straddle the two leaf depths, so [a, b] may span two levels: keep climbing until
  // the whole range has collapsed into the root (b == 0), not just its left end; the root itself has no parent.
  while (b > 0) {
    const int64_t lo = a > 0 ? a : 1;
    const int64_t pa = (lo - 1) / 2, pb = (b - 1) / 2;
    const int np = (int)(pb - pa + 1);
    for (int i = tid; i < np; i += nt) {
      const int64_t p = pa + i;
      const int64_t l = 2 * p + 1, r = 2 * p + 2;
      double c = 0.0;
      if (l >= lo && l <= b) c += __ldcg(cur + (l - a));
      if (r >= lo && r <= b) c += __ldcg(cur + (r - a));
      __stcg(nxt + i, c);
      __stcg(tree + p, __ldcg(tree + p) + c);
    }
    __syncthreads();
    double* t = cur; cur = nxt; nxt = t;
    a = pa;
    b = pb;
  }
}

// ---- replay add of a whole row when E is a power of two (every BASELINE configuration): the E leaves of ring row `row` are the
// leaves of ONE complete subtree, rooted at node row + R - 1 whatever R is (heap layout: the descendants of node r at relative depth k
// are [(r + 1) 2^k - 1, (r + 1) 2^k - 1 + 2^k)).  Setting them all to the same value v therefore needs no read-modify-write below that
// root: the node h levels above the leaves is exactly v * 2^h.  All CTAs fill the 2E - 2 nodes below the root with plain stores; one
// thread rewrites the root (new = v * E) and adds new - old to its <= log2(2R) ancestors.  57 us (one CTA, a block barrier and an L2
// round trip per level, twice for n-step windows) -> a few us.  CPU twin: oracle/engine.py::_tree_set_const.
__device__ __forceinline__ void subtree_fill(double* __restrict__ tree, int64_t root, int logE, double v, int64_t t0, int64_t stride) {
  // nodes at relative depth k = 1 .. logE below `root`, 2^k each: global thread index walks the 2E - 2 of them
  const int64_t total = ((int64_t)2 << logE) - 2;
  for (int64_t i = t0; i < total; i += stride) {
    const int k = 63 - __clzll(i + 2);              // i + 2 in [2^k, 2^(k+1))
    const int64_t j = i + 2 - ((int64_t)1 << k);
    __stcg(tree + ((root + 1) << k) - 1 + j, v * (double)((int64_t)1 << (logE - k)));
  }
}
__device__ inline void subtree_root_and_path(double* __restrict__ tree, int64_t root, int logE, double v) {
  const double nv = v * (double)((int64_t)1 << logE);
  const double delta = nv - __ldcg(tree + root);
  __stcg(tree + root, nv);
  int64_t p = root;
  while (p > 0) {
    p = (p - 1) / 2;
    __stcg(tree + p, __ldcg(tree + p) + delta);
  }
}
__global__ void __launch_bounds__(1024)
post_step_pow2_kernel(const __grid_constant__ srlx_engine eng, const int logE) {
  srlx_state* st = eng.state;
  const uint64_t g = st->vec_steps;
  const int E = eng.n_envs, R = eng.ring_rows, M = eng.multisteps;
  const int row = (int)(g % (uint64_t)R);
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  const double maxp = st->max_priority;
  const bool add_row = M == 1 || g >= (uint64_t)(M - 1);
  const int arow = M == 1 ? row : (int)((g - (uint64_t)(M - 1)) % (uint64_t)R);
  if (eng.mem_kind == SRLX_MEM_PROPORTIONAL) {
    if (M > 1) subtree_fill(eng.tree, (int64_t)row + R - 1, logE, 0.0, t0, stride);   // the row being overwritten leaves the sampleable set
    if (add_row) subtree_fill(eng.tree, (int64_t)arow + R - 1, logE, maxp, t0, stride);  // its M-step windows are complete
  }
  if (t0 == 0) {
    if (eng.mem_kind == SRLX_MEM_PROPORTIONAL) {
      if (M > 1) subtree_root_and_path(eng.tree, (int64_t)row + R - 1, logE, 0.0);
      if (add_row) subtree_root_and_path(eng.tree, (int64_t)arow + R - 1, logE, maxp);
    }
    const uint64_t rows_added = (g + 1 >= (uint64_t)(M - 1)) ? (g + 1 - (uint64_t)(M - 1)) : 0;
    const uint64_t rows_cap = (uint64_t)(R - (M - 1));
    st->mem_size = (uint64_t)E * (rows_added < rows_cap ? rows_added : rows_cap);
    st->vec_steps = g + 1;
    st->total_step += (uint64_t)E;
  }
}

__global__ void __launch_bounds__(1024)
post_step_kernel(const __grid_constant__ srlx_engine eng, double* __restrict__ scratch) {
  srlx_state* st = eng.state;
  const uint64_t g = st->vec_steps;
  const int E = eng.n_envs, R = eng.ring_rows, M = eng.multisteps;
  const int64_t cap = (int64_t)R * E;
  const int row = (int)(g % (uint64_t)R);
  if (eng.mem_kind == SRLX_MEM_PROPORTIONAL) {
    const double maxp = st->max_priority;
    if (M == 1) {
      tree_set_row(eng.tree, cap, (int64_t)row * E, E, maxp, scratch);
    } else {
      // the row being overwritten leaves the sampleable set; row g-M+1 enters it (its M-step windows are complete)
      tree_set_row(eng.tree, cap, (int64_t)row * E, E, 0.0, scratch);
      if (g >= (uint64_t)(M - 1)) {
        const int arow = (int)((g - (uint64_t)(M - 1)) % (uint64_t)R);
        tree_set_row(eng.tree, cap, (int64_t)arow * E, E, maxp, scratch);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint64_t rows_added = (g + 1 >= (uint64_t)(M - 1)) ? (g + 1 - (uint64_t)(M - 1)) : 0;
    const uint64_t rows_cap = (uint64_t)(R - (M - 1));
    st->mem_size = (uint64_t)E * (rows_added < rows_cap ? rows_added : rows_cap);
    st->vec_steps = g + 1;
    st->total_step += (uint64_t)E;
  }
}

// evaluation rollouts advance the step counter only
__global__ void eval_post_step_kernel(srlx_state* st, int E) {
  st->vec_steps += 1;
  st->total_step += (uint64_t)E;
}

// ---- external actor: transitions produced by a HOST loop (the reference's own core_play.play driving the plug-in Worker,
//      simple_distributed_rl_b200/srl_classes.py) enter the ring as one row of E records; post_step_kernel then does the replay add
__global__ void ext_row_write_kernel(const __grid_constant__ srlx_engine eng, const float* __restrict__ obs, const float* __restrict__ next_obs,
                                     const int32_t* __restrict__ action, const float* __restrict__ reward,
                                     const uint8_t* __restrict__ term, const uint8_t* __restrict__ done,
                                     const uint32_t* __restrict__ next_invalid) {
  const int E = eng.n_envs, D = eng.obs_dim;
  const int row = (int)(eng.state->vec_steps % (uint64_t)eng.ring_rows);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const size_t slot = (size_t)row * E + e;
    for (int d = 0; d < D; ++d) {
      eng.ring_obs[slot * D + d] = obs[(size_t)e * D + d];
      eng.ring_next_obs[slot * D + d] = next_obs[(size_t)e * D + d];
    }
    eng.ring_action[slot] = action[e];
    eng.ring_reward[slot] = reward[e];
    eng.ring_term[slot] = term[e] ? 1 : 0;
    eng.ring_done[slot] = done[e] ? 1 : 0;
    if (eng.ring_invalid) eng.ring_invalid[slot] = next_invalid ? next_invalid[e] : 0u;
  }
}

// ---- device-backed EnvBase (srl/base/env/base.py:60-137): reset() and step(action) of the closed-form envs for a host loop.
// reset-if-needed + observation of every env copy
__global__ void env_reset_obs_kernel(const __grid_constant__ srlx_engine eng, float* __restrict__ out_obs, int force) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < eng.n_envs; e += gridDim.x * blockDim.x) {
    double* st = eng.env_state + (size_t)e * 4;
    if (force || eng.env_needs_reset[e]) {
      const uint32_t ep = eng.env_episode[e];
      env_reset(eng, (uint32_t)e, ep, st);
      eng.env_episode[e] = ep + 1;
      eng.env_step_num[e] = 0;
      eng.env_ep_reward[e] = 0.0;
      eng.env_needs_reset[e] = 0;
    }
    float obs[SRLX_MAX_OBS];
    env_obs(eng, st, obs);
    for (int d = 0; d < eng.obs_dim; ++d) out_obs[(size_t)e * eng.obs_dim + d] = obs[d];
  }
}
// one env.step(action) per env copy with CALLER-supplied actions: (next observation, raw reward, terminated, truncated)
__global__ void env_step_actions_kernel(const __grid_constant__ srlx_engine eng, const int32_t* __restrict__ actions, float* __restrict__ out_obs,
                                        double* __restrict__ out_reward, uint8_t* __restrict__ out_term, uint8_t* __restrict__ out_trunc) {
  const uint64_t g = eng.state->vec_steps;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < eng.n_envs; e += gridDim.x * blockDim.x) {
    double* st = eng.env_state + (size_t)e * 4;
    int a = actions[e];
    a = a < 0 ? 0 : (a >= eng.n_actions ? eng.n_actions - 1 : a);
    bool terminated = false;
    const double r = env_step(eng, (uint32_t)e, g, a, st, terminated);
    const int step_num = eng.env_step_num[e] + 1;
    eng.env_step_num[e] = step_num;
    bool truncated = step_num >= eng.trunc_limit;
    if (eng.trunc_overrides_term) terminated = terminated && !truncated;
    else truncated = truncated && !terminated;
    eng.env_ep_reward[e] += r;
    if (terminated || truncated) eng.env_needs_reset[e] = 1;
    float obs[SRLX_MAX_OBS];
    env_obs(eng, st, obs);
    for (int d = 0; d < eng.obs_dim; ++d) out_obs[(size_t)e * eng.obs_dim + d] = obs[d];
    out_reward[e] = r;
    out_term[e] = terminated ? 1 : 0;
    out_trunc[e] = truncated ? 1 : 0;
  }
}

__global__ void engine_reset_kernel(const __grid_constant__ srlx_engine eng) {
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  const size_t cap = (size_t)eng.ring_rows * eng.n_envs;
  for (size_t i = i0; i < (size_t)eng.n_envs; i += stride) {
    eng.env_needs_reset[i] = 1;
    eng.env_episode[i] = 0;
    eng.env_step_num[i] = 0;
    eng.env_ep_reward[i] = 0.0;
    if (eng.env_last_ep_len) eng.env_last_ep_len[i] = 0;
    if (eng.env_first_ep_reward) eng.env_first_ep_reward[i] = 0.0;
    for (int d = 0; d < 4; ++d) eng.env_state[i * 4 + d] = 0.0;
  }
  for (size_t i = i0; i < cap; i += stride) {
    if (eng.ring_done) eng.ring_done[i] = 0;
    if (eng.ring_term) eng.ring_term[i] = 0;
  }
  if (eng.tree)
    for (size_t i = i0; i < 2 * cap - 1; i += stride) eng.tree[i] = 0.0;
  if (i0 == 0) {
    srlx_state z = {};
    z.max_priority = 1.0;  // ProportionalMemory.clear (proportional_memory.py:114-117)
    *eng.state = z;
  }
}

}  // namespace srlx

// ---- host side ---------------------------------------------------------------------------------------------------
namespace srlx {
// replay add of the row just written + counters: the complete-subtree kernel when E is a power of two (and the ring holds more than
// one row, so that the two subtrees of an n-step add are distinct from the root), else the generic level-by-level kernel
static void launch_post_step(const srlx_engine* eng, cudaStream_t st) {
  const int E = eng->n_envs;
  const bool pow2 = E >= 2 && (E & (E - 1)) == 0 && eng->ring_rows >= 2;
  if (pow2) {
    int logE = 0;
    while ((1 << logE) < E) ++logE;
    const long long nodes = 2ll * ((2ll << logE) - 2);
    int grid = (int)((nodes + 1023) / 1024);
    grid = grid < 1 ? 1 : (grid > 148 ? 148 : grid);
    post_step_pow2_kernel<<<grid, 1024, 0, st>>>(*eng, logE);
  } else {
    post_step_kernel<<<1, 1024, 0, st>>>(*eng, eng->tree_scratch);
  }
}
static int check_engine(const srlx_engine* eng) {
  SRLX_REQUIRE(eng != nullptr, "engine is NULL");
  SRLX_REQUIRE(eng->n_envs >= 1, "n_envs must be >= 1");
  // the device envs have <= 4 observation floats; an external env (host loop) may hand over up to SRLX_MAX_OBS (stacked states)
  SRLX_REQUIRE(eng->obs_dim >= 1 && eng->obs_dim <= (eng->env_id == SRLX_ENV_EXTERNAL ? SRLX_MAX_OBS : 4), "obs_dim %d unsupported", eng->obs_dim);
  SRLX_REQUIRE(eng->n_actions >= 1 && eng->n_actions <= SRLX_MAX_ACTIONS, "n_actions %d out of range", eng->n_actions);
  SRLX_REQUIRE(eng->net.n_layers >= 1 && eng->net.n_layers <= SRLX_MAX_LAYERS, "n_layers %d out of range", eng->net.n_layers);
  SRLX_REQUIRE(eng->net.in_dim == eng->obs_dim, "net.in_dim != obs_dim");
  SRLX_REQUIRE(eng->ring_rows >= eng->multisteps, "ring_rows (%d) must be >= multisteps (%d)", eng->ring_rows, eng->multisteps);
  SRLX_REQUIRE(eng->env_id == SRLX_ENV_GRID || eng->env_id == SRLX_ENV_CARTPOLE || eng->env_id == SRLX_ENV_PENDULUM ||
                   eng->env_id == SRLX_ENV_EXTERNAL,
               "unknown env_id %d", eng->env_id);
  SRLX_REQUIRE(eng->state && eng->env_state && eng->env_step_num && eng->env_episode && eng->env_ep_reward &&
                   eng->env_needs_reset && eng->params,
               "engine buffer pointer is NULL");
  return 0;
}
}  // namespace srlx

extern "C" int srlx_engine_reset(const srlx_engine* eng, uintptr_t cuda_stream) {
  using namespace srlx;
  if (int rc = check_engine(eng)) return rc;
  engine_reset_kernel<<<296, 256, 0, (cudaStream_t)cuda_stream>>>(*eng);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int srlx_ext_step(const srlx_engine* eng, const float* obs_dev, const float* next_obs_dev, const int32_t* action_dev,
                             const float* reward_dev, const unsigned char* term_dev, const unsigned char* done_dev, uintptr_t cuda_stream) {
  return srlx_ext_step_masked(eng, obs_dev, next_obs_dev, action_dev, reward_dev, term_dev, done_dev, nullptr, cuda_stream);
}

extern "C" int srlx_ext_step_masked(const srlx_engine* eng, const float* obs_dev, const float* next_obs_dev, const int32_t* action_dev,
                                    const float* reward_dev, const unsigned char* term_dev, const unsigned char* done_dev,
                                    const uint32_t* next_invalid_dev, uintptr_t cuda_stream) {
  using namespace srlx;
  SRLX_REQUIRE(eng != nullptr, "engine is NULL");
  SRLX_REQUIRE(next_invalid_dev == nullptr || eng->ring_invalid != nullptr, "invalid-action masks need the ring_invalid buffer");
  SRLX_REQUIRE(eng != nullptr, "engine is NULL");
  SRLX_REQUIRE(eng->n_envs >= 1 && eng->obs_dim >= 1 && eng->obs_dim <= SRLX_MAX_OBS, "n_envs / obs_dim out of range");
  SRLX_REQUIRE(eng->ring_rows >= eng->multisteps && eng->multisteps >= 1, "ring_rows (%d) must be >= multisteps (%d)", eng->ring_rows, eng->multisteps);
  SRLX_REQUIRE(eng->state && eng->ring_obs && eng->ring_next_obs && eng->ring_action && eng->ring_reward && eng->ring_term && eng->ring_done,
               "ring buffer pointer is NULL");
  SRLX_REQUIRE(eng->mem_kind == SRLX_MEM_UNIFORM || (eng->tree && eng->tree_scratch), "proportional memory needs tree + tree_scratch");
  SRLX_REQUIRE(obs_dev && next_obs_dev && action_dev && reward_dev && term_dev && done_dev, "record pointer is NULL");
  const int grid = (eng->n_envs + 255) / 256 < 296 ? (eng->n_envs + 255) / 256 : 296;
  ext_row_write_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(*eng, obs_dev, next_obs_dev, action_dev, reward_dev, term_dev, done_dev,
                                                                       next_invalid_dev);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  launch_post_step(eng, (cudaStream_t)cuda_stream);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int check_env_only(const srlx_engine* eng) {
  using namespace srlx;
  SRLX_REQUIRE(eng != nullptr, "engine is NULL");
  SRLX_REQUIRE(eng->env_id == SRLX_ENV_GRID || eng->env_id == SRLX_ENV_CARTPOLE || eng->env_id == SRLX_ENV_PENDULUM,
               "env_id %d has no device implementation", eng->env_id);
  SRLX_REQUIRE(eng->n_envs >= 1 && eng->obs_dim >= 1 && eng->obs_dim <= 4, "n_envs / obs_dim out of range");
  SRLX_REQUIRE(eng->state && eng->env_state && eng->env_step_num && eng->env_episode && eng->env_ep_reward && eng->env_needs_reset,
               "env buffer pointer is NULL");
  return 0;
}

extern "C" int srlx_env_reset_obs(const srlx_engine* eng, int force, float* out_obs_dev, uintptr_t cuda_stream) {
  using namespace srlx;
  if (int rc = check_env_only(eng)) return rc;
  SRLX_REQUIRE(out_obs_dev != nullptr, "out_obs is NULL");
  const int grid = (eng->n_envs + 255) / 256 < 296 ? (eng->n_envs + 255) / 256 : 296;
  env_reset_obs_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(*eng, out_obs_dev, force);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int srlx_env_step_actions(const srlx_engine* eng, const int32_t* actions_dev, float* out_obs_dev, double* out_reward_dev,
                                     unsigned char* out_term_dev, unsigned char* out_trunc_dev, uintptr_t cuda_stream) {
  using namespace srlx;
  if (int rc = check_env_only(eng)) return rc;
  SRLX_REQUIRE(actions_dev && out_obs_dev && out_reward_dev && out_term_dev && out_trunc_dev, "argument pointer is NULL");
  const int grid = (eng->n_envs + 255) / 256 < 296 ? (eng->n_envs + 255) / 256 : 296;
  env_step_actions_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(*eng, actions_dev, out_obs_dev, out_reward_dev, out_term_dev, out_trunc_dev);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  eval_post_step_kernel<<<1, 1, 0, (cudaStream_t)cuda_stream>>>(eng->state, eng->n_envs);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int srlx_vec_step(const srlx_engine* eng, int training, uintptr_t cuda_stream) {
  using namespace srlx;
  if (int rc = check_engine(eng)) return rc;
  SRLX_REQUIRE(eng->env_id != SRLX_ENV_EXTERNAL, "srlx_vec_step: the engine's env is external (host loop); use srlx_ext_step");
  if (training) {
    SRLX_REQUIRE(eng->ring_obs && eng->ring_next_obs && eng->ring_action && eng->ring_reward && eng->ring_term && eng->ring_done,
                 "ring buffer pointer is NULL");
    SRLX_REQUIRE(eng->mem_kind == SRLX_MEM_UNIFORM || (eng->tree && eng->tree_scratch), "proportional memory needs tree + tree_scratch");
  }
  const NetPlan pl = make_plan(eng->net);
  const RolloutSmem so = rollout_smem(*eng, pl);
  int dev = 0, max_smem = 0, n_sm = 0;
  SRLX_CHECK_CUDA(cudaGetDevice(&dev));
  SRLX_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  SRLX_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  SRLX_REQUIRE((int)so.total + 1024 <= max_smem, "network too large for the rollout kernel: needs %zu bytes of shared memory", so.total);
  // one CTA per SM when there is enough work; each CTA walks its envs in tiles of kRowTile
  int per = (eng->n_envs + n_sm - 1) / n_sm;
  per = round_up(per < kRowTile ? kRowTile : per, kRowTile);
  const int grid = (eng->n_envs + per - 1) / per;
  SRLX_CHECK_CUDA(cudaFuncSetAttribute(rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)so.total));
  rollout_kernel<<<grid, kRolloutThreads, so.total, (cudaStream_t)cuda_stream>>>(*eng, per, training);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  if (training) {
    launch_post_step(eng, (cudaStream_t)cuda_stream);
  } else {
    eval_post_step_kernel<<<1, 1, 0, (cudaStream_t)cuda_stream>>>(eng->state, eng->n_envs);
  }
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int srlx_engine_run(const srlx_engine* eng, uint32_t n_steps, uint32_t updates_per_step, int training,
                               uintptr_t cuda_stream) {
  for (uint32_t s = 0; s < n_steps; ++s) {
    if (int rc = srlx_vec_step(eng, training, cuda_stream)) return rc;
    if (training && updates_per_step)
      if (int rc = srlx_learn(eng, updates_per_step, cuda_stream)) return rc;
  }
  return 0;
}
