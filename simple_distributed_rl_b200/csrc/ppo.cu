// ppo.cu -- PPO on device (BASELINE configs[4]: Pendulum-v1 continuous, thousands of env copies, GAE, on-policy
// rollout buffer, no replay).  Restates srl/algorithms/ppo/ppo.py for E vectorised env copies:
//   ppo_rollout_kernel   Worker.policy (:307-356): actor-critic forward, action ~ Normal(loc, exp(log_scale)) (log_scale clipped to the
//                        stable-gradients range, srl/rl/tf/distributions/normal_dist_block.py:140-155) or ~ Categorical(logits)
//                        (categorical_dist_block.py), log_prob floored at log(1e-6); env action = clip(rescale_from(a)) (np_array.py:64-95);
//                        env.step + done typing (env_run.py:254-366); one row of the time-major rollout buffer [T][E]
//   ppo_value_kernel     Worker.on_step at episode end (:389-392): V(s) of every stored state with the CURRENT parameters; the GAE /
//                        Monte-Carlo accumulation itself is srlx_returns_scan (csrc/returns.cu, bit-exact against the reference)
//   ppo_update_kernel    Trainer._train (:208-291) + ActorCriticNetwork.compute_train_loss (:103-169): minibatch of B distinct samples
//                        (ReplayBuffer.sample), baseline, clipped surrogate, clipped value loss, the reference's entropy term, gradient
//                        clipping by global norm, Adam as Keras applies it, staircase exponential LR decay; n dependent updates per launch
// The network is the reference's ActorCriticNetwork (:55-101): trunk MLP -> {value MLP -> 1, policy MLP -> loc / log_scale or logits}.
// It is held as TWO dense stacks over ONE flat parameter buffer (value stack = trunk + value block, policy stack = trunk + policy
// block, the trunk layers carrying the same offsets in both), so the tile forward / backward of net.cuh serve both and the trunk's
// gradient is the sum of what the two backward passes accumulate.  CPU twin: oracle/ppo.py (a torch restatement: TensorFlow is not
// available where this was built, so parity with the reference's TF code is by restatement, not by execution).
#include "envs.cuh"
#include "net.cuh"

namespace srlx {

constexpr int kPpoThreads = 256, kPpoUpdThreads = 512;

struct PpoSmem {
  size_t wv, wp, av, ap, qv, qp, total;
};
__host__ __device__ inline PpoSmem ppo_smem(const srlx_ppo& p, const NetPlan& pv, const NetPlan& pp) {
  PpoSmem s;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) / 16 * 16; return o; };
  s.wv = take((size_t)pv.weff_floats * 4);
  s.wp = take((size_t)pp.weff_floats * 4);
  s.av = take((size_t)pv.act_floats * 4);
  s.ap = take((size_t)pp.act_floats * 4);
  s.qv = take((size_t)kRowTile * 4);
  s.qp = take((size_t)kRowTile * SRLX_MAX_ACTIONS * 4);
  s.total = off;
  return s;
}

// one stack's weights from the flat parameter buffer into its shared-memory plan (padding stays zero)
__device__ inline void ppo_load_weights(const srlx_net& net, const NetPlan& pl, const float* __restrict__ params, float* weff) {
  for (int l = 0; l < net.n_layers; ++l) {
    const int U = net.out_dim[l], K = net.k_dim[l];
    const int n = U * K, nt = blockDim.x;
    for (int i0 = threadIdx.x; i0 < n; i0 += 4 * nt) {  // four independent L2 loads in flight per thread
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * nt;
        v[j] = i < n ? __ldcg(params + net.w_off[l] + i) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * nt;
        if (i < n) {
          const int u = i / K, k = i - u * K;
          weff[pl.w_s[l] + u * pl.ldw[l] + k] = v[j];
        }
      }
    }
    for (int u = threadIdx.x; u < U; u += blockDim.x) weff[pl.b_s[l] + u] = __ldcg(params + net.b_off[l] + u);
  }
}

__device__ __forceinline__ float ppo_normal_logprob(float x, float loc, float ls) {
  // -0.5 log(2 pi) - log_scale - 0.5 ((x - loc) / exp(log_scale))^2   (normal_dist_block.py:14-21)
  const float z = (x - loc) / expf(ls);
  return -0.9189385332046727f - ls - 0.5f * (z * z);
}

__global__ void __launch_bounds__(kPpoThreads)
ppo_rollout_kernel(const __grid_constant__ srlx_ppo ppo, const int envs_per_cta, const int training) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned long long s_episodes, s_eplen;
  __shared__ double s_epreward;
  const srlx_engine& eng = ppo.env;
  const NetPlan pv = make_plan(ppo.net_v), pp = make_plan(ppo.net_p);
  const PpoSmem so = ppo_smem(ppo, pv, pp);
  float* wv = reinterpret_cast<float*>(smem_raw + so.wv);
  float* wp = reinterpret_cast<float*>(smem_raw + so.wp);
  float* av = reinterpret_cast<float*>(smem_raw + so.av);
  float* ap = reinterpret_cast<float*>(smem_raw + so.ap);
  float* qv = reinterpret_cast<float*>(smem_raw + so.qv);
  float* qp = reinterpret_cast<float*>(smem_raw + so.qp);
  const int tid = threadIdx.x;
  const int E = eng.n_envs, D = eng.obs_dim, T = ppo.horizon, nout = ppo.net_p.n_actions;
  const uint64_t g = eng.state->vec_steps;
  const int row = (int)(g % (uint64_t)T);
  if (tid == 0) { s_episodes = 0; s_eplen = 0; s_epreward = 0.0; }
  zero_floats(wv, pv.weff_floats);
  zero_floats(wp, pp.weff_floats);
  zero_floats(av, pv.act_floats);
  zero_floats(ap, pp.act_floats);
  __syncthreads();
  ppo_load_weights(ppo.net_v, pv, ppo.params, wv);
  ppo_load_weights(ppo.net_p, pp, ppo.params, wp);
  __syncthreads();
  const int e_begin = blockIdx.x * envs_per_cta, e_end = min(E, e_begin + envs_per_cta);
  for (int e0 = e_begin; e0 < e_end; e0 += kRowTile) {
    const int Rr = min(kRowTile, e_end - e0);
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
      float obs[SRLX_MAX_OBS];
      env_obs(eng, st, obs);
      for (int d = 0; d < D; ++d) {
        av[pv.x_s[0] + tid * pv.ldx[0] + d] = obs[d];
        ap[pp.x_s[0] + tid * pp.ldx[0] + d] = obs[d];
      }
    }
    __syncthreads();
    net_forward_tile(ppo.net_v, pv, wv, av, Rr, qv, 1);
    net_forward_tile(ppo.net_p, pp, wp, ap, Rr, qp, nout);
    if (tid < Rr) {
      const int e = e0 + tid;
      const float v = qv[tid];
      const float* po = qp + tid * nout;
      const uint4 w = philox(eng.seed, STREAM_POLICY, (uint32_t)e, (uint32_t)g, (uint32_t)(g >> 32));
      float act_f, logp;
      double* st = eng.env_state + (size_t)e * 4;
      bool terminated = false;
      double r;
      if (ppo.continuous) {
        const float loc = po[0];
        const float ls = fminf(fmaxf(po[1], (float)ppo.log_scale_lo), (float)ppo.log_scale_hi);
        float a = loc;  // evaluation: the mean (ppo.py:333-336)
        if (training) {
          const float u1 = ((float)(w.x >> 8) + 1.0f) * (1.0f / 16777216.0f), u2 = (float)(w.y >> 8) * (1.0f / 16777216.0f);
          float sn, cs;
          sincospif(2.0f * u2, &sn, &cs);
          a = fmaf(expf(ls), sqrtf(-2.0f * logf(u1)) * cs, loc);
        }
        logp = ppo_normal_logprob(a, loc, ls);
        act_f = a;
        // env action: rescale_from [-1, 1] to the env's range, then sanitize = clip (np_array.py:64-67, 93-95)
        double ua = ((double)a + 1.0) * 0.5 * (ppo.action_high - ppo.action_low) + ppo.action_low;
        ua = ua < ppo.action_low ? ppo.action_low : (ua > ppo.action_high ? ppo.action_high : ua);
        r = pendulum_step_torque(ua, st, terminated);
      } else {
        // Categorical(logits).sample(): inverse cdf of softmax(logits) on one uniform; log_prob = log_softmax(logits)[a]
        float mx = po[0];
        for (int a = 1; a < nout; ++a) mx = fmaxf(mx, po[a]);
        float sum = 0.f;
        for (int a = 0; a < nout; ++a) sum += expf(po[a] - mx);
        const float u = u01_f32(w.x) * sum;
        int action = nout - 1;
        float run = 0.f;
        for (int a = 0; a < nout; ++a) {
          run += expf(po[a] - mx);
          if (u < run) { action = a; break; }
        }
        logp = (po[action] - mx) - logf(sum);
        act_f = (float)action;
        r = env_step(eng, (uint32_t)e, g, action, st, terminated);
      }
      logp = fmaxf(logp, -13.815510557964274f);  // np.maximum(log_prob, math.log(1e-6)) (ppo.py:326,341)
      const int step_num = eng.env_step_num[e] + 1;
      eng.env_step_num[e] = step_num;
      bool truncated = step_num >= eng.trunc_limit;
      if (eng.trunc_overrides_term) terminated = terminated && !truncated;
      else truncated = truncated && !terminated;
      const bool done = terminated || truncated;
      const double ep_reward = eng.env_ep_reward[e] + r;
      eng.env_ep_reward[e] = ep_reward;
      if (training) {
        const size_t slot = (size_t)row * E + e;
        for (int d = 0; d < D; ++d) ppo.buf_obs[slot * D + d] = av[pv.x_s[0] + tid * pv.ldx[0] + d];
        ppo.buf_action[slot] = act_f;
        ppo.buf_v[slot] = v;
        ppo.buf_logp[slot] = logp;
        ppo.buf_reward[slot] = (float)((r + eng.reward_shift) * eng.reward_scale);  // worker_run.py:348
        ppo.buf_done[slot] = done ? 1 : 0;
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
    }
    __syncthreads();
  }
  if (tid == 0 && s_episodes) {
    atomicAdd((unsigned long long*)&eng.state->episode_count, s_episodes);
    atomicAdd((unsigned long long*)&eng.state->episode_len_sum, s_eplen);
    atomicAdd(&eng.state->episode_reward_sum, s_epreward);
  }
}

__global__ void ppo_step_count_kernel(srlx_state* st, int E) {
  st->vec_steps += 1;
  st->total_step += (uint64_t)E;
}

// V(s) of n stored states with the current parameters (value stack only)
__global__ void __launch_bounds__(kPpoThreads)
ppo_value_kernel(const __grid_constant__ srlx_ppo ppo, const float* __restrict__ obs, const long long n, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const NetPlan pv = make_plan(ppo.net_v), pp = make_plan(ppo.net_p);
  const PpoSmem so = ppo_smem(ppo, pv, pp);
  float* wv = reinterpret_cast<float*>(smem_raw + so.wv);
  float* av = reinterpret_cast<float*>(smem_raw + so.av);
  float* qv = reinterpret_cast<float*>(smem_raw + so.qv);
  const int tid = threadIdx.x, D = ppo.env.obs_dim;
  zero_floats(wv, pv.weff_floats);
  zero_floats(av, pv.act_floats);
  __syncthreads();
  ppo_load_weights(ppo.net_v, pv, ppo.params, wv);
  __syncthreads();
  const long long n_tiles = (n + kRowTile - 1) / kRowTile;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long r0 = t * kRowTile;
    const int Rr = (int)min((long long)kRowTile, n - r0);
    for (int i = tid; i < Rr * D; i += blockDim.x) {
      const int r = i / D, d = i - r * D;
      av[pv.x_s[0] + r * pv.ldx[0] + d] = __ldg(obs + (size_t)(r0 + r) * D + d);
    }
    __syncthreads();
    net_forward_tile(ppo.net_v, pv, wv, av, Rr, qv, 1);
    if (tid < Rr) out[r0 + tid] = qv[tid];
    __syncthreads();
  }
}

// ---- trainer ---------------------------------------------------------------------------------------------------------------
struct PpoUpdSmem {
  size_t wv, wp, av, ap, qv, qp, G, rows, red, total;
};
__host__ __device__ inline PpoUpdSmem ppo_upd_smem(const srlx_ppo& p, const NetPlan& pv, const NetPlan& pp, bool g_in_smem) {
  PpoUpdSmem s;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) / 16 * 16; return o; };
  s.wv = take((size_t)pv.weff_floats * 4);
  s.wp = take((size_t)pp.weff_floats * 4);
  s.av = take((size_t)pv.act_floats * 4);
  s.ap = take((size_t)pp.act_floats * 4);
  s.qv = take((size_t)kRowTile * 4);
  s.qp = take((size_t)kRowTile * SRLX_MAX_ACTIONS * 4);
  s.G = take(g_in_smem ? (size_t)p.n_params * 4 : 0);
  s.rows = take((size_t)kRowTile * 8 * 4);  // per row: idx, action, old_v, old_logp, ret, adv, dv, -
  s.red = take(64 * 8);
  s.total = off;
  return s;
}

// g_ext: the gradient accumulator in global memory (wide networks: two weight copies + activations already fill the shared memory);
// NULL -> shared memory
__global__ void __launch_bounds__(kPpoUpdThreads)
ppo_update_kernel(const __grid_constant__ srlx_ppo ppo, const uint32_t n_updates, float* __restrict__ g_ext) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const NetPlan pv = make_plan(ppo.net_v), pp = make_plan(ppo.net_p);
  const PpoUpdSmem so = ppo_upd_smem(ppo, pv, pp, g_ext == nullptr);
  float* wv = reinterpret_cast<float*>(smem_raw + so.wv);
  float* wp = reinterpret_cast<float*>(smem_raw + so.wp);
  float* av = reinterpret_cast<float*>(smem_raw + so.av);
  float* ap = reinterpret_cast<float*>(smem_raw + so.ap);
  float* qv = reinterpret_cast<float*>(smem_raw + so.qv);
  float* qp = reinterpret_cast<float*>(smem_raw + so.qp);
  float* G = g_ext ? g_ext : reinterpret_cast<float*>(smem_raw + so.G);
  float* rows = reinterpret_cast<float*>(smem_raw + so.rows);
  double* red = reinterpret_cast<double*>(smem_raw + so.red);
  int* ridx = reinterpret_cast<int*>(rows);  // [B] sample index
  float* r_act = rows + kRowTile, *r_oldv = rows + 2 * kRowTile, *r_oldlp = rows + 3 * kRowTile, *r_ret = rows + 4 * kRowTile,
        *r_adv = rows + 5 * kRowTile, *r_dv = rows + 6 * kRowTile;
  const srlx_engine& eng = ppo.env;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x;
  const int B = ppo.batch_size, D = eng.obs_dim, E = eng.n_envs, T = ppo.horizon, P = ppo.n_params, nout = ppo.net_p.n_actions;
  const uint32_t n_items = (uint32_t)((uint64_t)T * E);
  srlx_ppo_state* ps = ppo.pstate;
  const uint64_t tc0 = ps->train_count;
  zero_floats(wv, pv.weff_floats);
  zero_floats(wp, pp.weff_floats);
  zero_floats(av, pv.act_floats);
  zero_floats(ap, pp.act_floats);
  __syncthreads();
  for (uint32_t upd = 0; upd < n_updates; ++upd) {
    const uint64_t tc = tc0 + upd;
    // ---- ReplayBuffer.sample: B distinct valid samples (first attempts in parallel, rejections resolved in sample order)
    if (warp == 0) {
      int pick = -1 - lane;
      if (lane < B) {
        const uint4 w = philox(eng.seed, STREAM_UNIFORM_SAMPLE, (uint32_t)lane, (uint32_t)tc, (uint32_t)(tc >> 32));
        pick = (int)u_below(w.x, n_items);
      }
      bool bad = lane < B && !ppo.buf_valid[pick];
      for (int j = 0; j < B; ++j) {
        const int pj = __shfl_sync(0xffffffffu, pick, j);
        bad |= (j < lane) & (pj == pick);
      }
      if (lane < B) ridx[lane] = pick;
      __syncwarp();
      if (__any_sync(0xffffffffu, bad)) {
        if (lane == 0) {
          for (int i = 0; i < B; ++i) {
            int k = 0;
            while (true) {
              bool b2 = !ppo.buf_valid[ridx[i]];
              for (int j = 0; j < i; ++j) b2 |= (ridx[j] == ridx[i]);
              if (!b2 || ++k >= 65536) break;
              const uint4 w = philox(eng.seed, STREAM_UNIFORM_SAMPLE, (uint32_t)i | ((uint32_t)k << 16), (uint32_t)tc, (uint32_t)(tc >> 32));
              ridx[i] = (int)u_below(w.x, n_items);
            }
          }
        }
        __syncwarp();
      }
      if (lane < B) {
        const int s = ridx[lane];
        r_act[lane] = ppo.buf_action[s];
        r_oldv[lane] = ppo.buf_v[s];
        r_oldlp[lane] = ppo.buf_logp[s];
        r_ret[lane] = ppo.buf_ret[s];
        for (int d = 0; d < D; ++d) {
          const float x = ppo.buf_obs[(size_t)s * D + d];
          av[pv.x_s[0] + lane * pv.ldx[0] + d] = x;
          ap[pp.x_s[0] + lane * pp.ldx[0] + d] = x;
        }
        if (ppo.dbg_idx) ppo.dbg_idx[lane] = s;
      }
    }
    ppo_load_weights(ppo.net_v, pv, ppo.params, wv);
    ppo_load_weights(ppo.net_p, pp, ppo.params, wp);
    for (int i = tid; i < P; i += nt) G[i] = 0.f;
    __syncthreads();
    if (ppo.state_normalized && tid < D) {  // (states - mean) / (std + 1e-8) over the minibatch (ppo.py:217-218)
      float m = 0.f;
      for (int b = 0; b < B; ++b) m += av[pv.x_s[0] + b * pv.ldx[0] + tid];
      m /= (float)B;
      float var = 0.f;
      for (int b = 0; b < B; ++b) { const float d = av[pv.x_s[0] + b * pv.ldx[0] + tid] - m; var += d * d; }
      const float sd = sqrtf(var / (float)B) + 1e-8f;
      for (int b = 0; b < B; ++b) {
        const float x = (av[pv.x_s[0] + b * pv.ldx[0] + tid] - m) / sd;
        av[pv.x_s[0] + b * pv.ldx[0] + tid] = x;
        ap[pp.x_s[0] + b * pp.ldx[0] + tid] = x;
      }
    }
    __syncthreads();
    net_forward_tile(ppo.net_v, pv, wv, av, B, qv, 1);
    net_forward_tile(ppo.net_p, pp, wp, ap, B, qp, nout);
    // ---- baseline on the minibatch (ppo.py:220-232)
    if (tid == 0) {
      float mean = 0.f, sd = 1.f;
      if (ppo.baseline_type >= 1 && ppo.baseline_type <= 3) {
        for (int b = 0; b < B; ++b) mean += r_ret[b];
        mean /= (float)B;
        float var = 0.f;
        for (int b = 0; b < B; ++b) { const float d = r_ret[b] - mean; var += d * d; }
        sd = sqrtf(var / (float)B) + 1e-8f;
      }
      for (int b = 0; b < B; ++b) {
        float a = r_ret[b];
        if (ppo.baseline_type == 1) a -= mean;
        else if (ppo.baseline_type == 2) a /= sd;
        else if (ppo.baseline_type == 3) a = (a - mean) / sd;
        r_adv[b] = a;
      }
    }
    __syncthreads();
    // ---- compute_train_loss (:103-169): per-row loss terms and their gradients wrt v and the policy outputs
    float l_pol = 0.f, l_val = 0.f, l_ent = 0.f;
    if (tid < B) {
      const int b = tid;
      const float v = qv[b], vt = r_ret[b];
      float adv = r_adv[b];
      if (ppo.baseline_type == 4) adv -= v;  // advantage - stop_gradient(v)
      float* po = qp + b * nout;
      float logp, dlp[SRLX_MAX_ACTIONS];
      if (ppo.continuous) {
        const float loc = po[0], ls_raw = po[1];
        const float ls = fminf(fmaxf(ls_raw, (float)ppo.log_scale_lo), (float)ppo.log_scale_hi);
        const float z = (r_act[b] - loc) / expf(ls);
        logp = -0.9189385332046727f - ls - 0.5f * z * z;
        dlp[0] = z / expf(ls);                                                      // d logp / d loc
        dlp[1] = (ls_raw >= (float)ppo.log_scale_lo && ls_raw <= (float)ppo.log_scale_hi) ? (z * z - 1.0f) : 0.f;  // through the clip
      } else {
        float mx = po[0];
        for (int a = 1; a < nout; ++a) mx = fmaxf(mx, po[a]);
        float sum = 0.f;
        for (int a = 0; a < nout; ++a) sum += expf(po[a] - mx);
        const int act = (int)r_act[b];
        logp = (po[act] - mx) - logf(sum);
        for (int a = 0; a < nout; ++a) dlp[a] = ((a == act) ? 1.f : 0.f) - expf(po[a] - mx) / sum;
      }
      const float ratio = expf(logp - r_oldlp[b]);
      float dL_dlogp = 0.f;  // d(total loss) / d logp of this row
      const float invB = 1.0f / (float)B;
      if (ppo.surrogate_clip) {
        const float rc = fminf(fmaxf(ratio, 1.0f - (float)ppo.policy_clip_range), 1.0f + (float)ppo.policy_clip_range);
        const float lu = ratio * adv, lc = rc * adv;
        l_pol = -fminf(lu, lc);
        // tf.minimum sends the gradient to the first argument on ties; the clipped branch has none where the clip binds
        if (lu <= lc) dL_dlogp += -adv * ratio * invB;
        else if (rc == ratio) dL_dlogp += -adv * ratio * invB;
      } else {
        l_pol = -ratio * adv;
        dL_dlogp += -adv * ratio * invB;
      }
      // "entropy": sum(-exp(logp) * logp) of the TAKEN action, weighted and negated (:164-167)
      const float pi = expf(logp);
      l_ent = (float)ppo.entropy_weight * (pi * logp);
      dL_dlogp += (float)ppo.entropy_weight * invB * pi * (logp + 1.0f);
      // value loss (:155-161)
      float dv;
      if (ppo.enable_value_clip) {
        const float lo = r_oldv[b] - (float)ppo.value_clip_range, hi = r_oldv[b] + (float)ppo.value_clip_range;
        const float vc = fminf(fmaxf(v, lo), hi);
        const float a1 = (v - vt) * (v - vt), a2 = (vc - vt) * (vc - vt);
        l_val = fmaxf(a1, a2);
        if (a1 >= a2) dv = 2.0f * (v - vt);                       // tf.maximum: the first argument on ties
        else dv = (v >= lo && v <= hi) ? 2.0f * (vc - vt) : 0.f;
      } else {
        l_val = (v - vt) * (v - vt);
        dv = 2.0f * (v - vt);
      }
      l_val *= (float)ppo.value_loss_weight;
      r_dv[b] = dv * (float)ppo.value_loss_weight * invB;
      for (int a = 0; a < nout; ++a) po[a] = dL_dlogp * dlp[a];  // qp now holds d loss / d policy outputs
    }
    // loss means for the info fields (warp 0 holds the rows)
    if (warp == 0) {
      for (int s = 16; s > 0; s >>= 1) {
        l_pol += __shfl_xor_sync(0xffffffffu, l_pol, s);
        l_val += __shfl_xor_sync(0xffffffffu, l_val, s);
        l_ent += __shfl_xor_sync(0xffffffffu, l_ent, s);
      }
      if (lane == 0) { red[8] = l_pol / B; red[9] = l_val / B; red[10] = l_ent / B; }
    }
    __syncthreads();
    net_backward_tile(ppo.net_v, pv, wv, av, B, r_dv, 1, G);
    __syncthreads();
    net_backward_tile(ppo.net_p, pp, wp, ap, B, qp, nout, G);
    __syncthreads();
    // ---- tf.clip_by_global_norm (:268-269): g * clip / max(norm, clip)
    {
      double acc = 0.0;
      for (int i = tid; i < P; i += nt) acc += (double)G[i] * (double)G[i];
      for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
      if (lane == 0) red[16 + warp] = acc;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < nt / 32; ++w) t += red[16 + w];
        const double norm = sqrt(t);
        red[0] = (ppo.grad_clip_norm > 0.0) ? ppo.grad_clip_norm / fmax(norm, ppo.grad_clip_norm) : 1.0;
        red[1] = norm;
      }
      __syncthreads();
    }
    // ---- Adam as Keras applies it: alpha = lr_t sqrt(1 - b2^t) / (1 - b1^t); p -= alpha m / (sqrt(v) + eps); staircase decay of lr
    {
      const float scale = (float)red[0];
      const uint64_t step = ps->adam_step + upd;  // optimizer.iterations before this apply
      double lr = ppo.lr;
      if (ppo.lr_decay_steps) lr = ppo.lr * pow(ppo.lr_decay_rate, (double)(step / ppo.lr_decay_steps));
      const double t1 = (double)(step + 1);
      const float alpha = (float)(lr * sqrt(1.0 - pow(ppo.adam_beta2, t1)) / (1.0 - pow(ppo.adam_beta1, t1)));
      const float b1 = (float)ppo.adam_beta1, b2 = (float)ppo.adam_beta2, eps = (float)ppo.adam_eps;
      for (int i0 = tid; i0 < P; i0 += 4 * nt) {  // four parameters per thread and pass: their loads are in flight together
        float gg[4], mm[4], vv[4], pp4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + j * nt;
          const bool ok = i < P;
          gg[j] = ok ? G[i] * scale : 0.f;
          mm[j] = ok ? __ldcg(ppo.adam_m + i) : 0.f;
          vv[j] = ok ? __ldcg(ppo.adam_v + i) : 0.f;
          pp4[j] = ok ? __ldcg(ppo.params + i) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + j * nt;
          if (i < P) {
            const float g = gg[j];
            const float m = mm[j] + (g - mm[j]) * (1.0f - b1);
            const float v = vv[j] + (g * g - vv[j]) * (1.0f - b2);
            ppo.adam_m[i] = m;
            ppo.adam_v[i] = v;
            ppo.params[i] = pp4[j] - alpha * m / (sqrtf(v) + eps);
            if (ppo.dbg_grads) ppo.dbg_grads[i] = g;
          }
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      ps->policy_loss = red[8];
      ps->value_loss = red[9];
      ps->entropy_loss = red[10];
      ps->grad_norm = red[1];
    }
    __threadfence();
    __syncthreads();
  }
  if (tid == 0) {
    ps->train_count = tc0 + n_updates;
    ps->adam_step += n_updates;
  }
}

static int ppo_check(const srlx_ppo* p) {
  SRLX_REQUIRE(p != nullptr, "ppo is NULL");
  const srlx_engine& e = p->env;
  SRLX_REQUIRE(e.n_envs >= 1 && e.obs_dim >= 1 && e.obs_dim <= 4, "n_envs / obs_dim out of range");
  SRLX_REQUIRE(e.env_id == SRLX_ENV_GRID || e.env_id == SRLX_ENV_CARTPOLE || e.env_id == SRLX_ENV_PENDULUM, "unknown env_id %d", e.env_id);
  SRLX_REQUIRE(!p->continuous || e.env_id == SRLX_ENV_PENDULUM, "a continuous policy needs a continuous-action env (Pendulum-v1)");
  SRLX_REQUIRE(p->horizon >= 1 && p->batch_size >= 1 && p->batch_size <= kRowTile, "horizon / batch_size (<= %d) out of range", kRowTile);
  SRLX_REQUIRE(p->net_v.n_layers >= 1 && p->net_v.n_layers <= SRLX_MAX_LAYERS && p->net_p.n_layers >= 1 && p->net_p.n_layers <= SRLX_MAX_LAYERS,
               "too many layers");
  SRLX_REQUIRE(p->net_v.n_actions == 1 && p->net_p.n_actions >= 1 && p->net_p.n_actions <= SRLX_MAX_ACTIONS, "bad output widths");
  SRLX_REQUIRE(!p->continuous || p->net_p.n_actions == 2, "a continuous policy has two outputs (loc, log_scale)");
  SRLX_REQUIRE(e.state && e.env_state && e.env_step_num && e.env_episode && e.env_ep_reward && e.env_needs_reset, "env buffer pointer is NULL");
  SRLX_REQUIRE(p->params && p->pstate, "params / pstate is NULL");
  return 0;
}

}  // namespace srlx

extern "C" size_t srlx_sizeof_ppo(void) { return sizeof(srlx_ppo); }
extern "C" size_t srlx_sizeof_ppo_state(void) { return sizeof(srlx_ppo_state); }

// one vector step of all E env copies under the current policy; training != 0: the step is stored in row (vec_steps % horizon)
extern "C" int srlx_ppo_vec_step(const srlx_ppo* ppo, int training, uintptr_t cuda_stream) {
  using namespace srlx;
  if (int rc = ppo_check(ppo)) return rc;
  if (training)
    SRLX_REQUIRE(ppo->buf_obs && ppo->buf_action && ppo->buf_v && ppo->buf_logp && ppo->buf_reward && ppo->buf_done, "rollout buffer pointer is NULL");
  const NetPlan pv = make_plan(ppo->net_v), pp = make_plan(ppo->net_p);
  const PpoSmem so = ppo_smem(*ppo, pv, pp);
  int dev = 0, max_smem = 0, n_sm = 0;
  SRLX_CHECK_CUDA(cudaGetDevice(&dev));
  SRLX_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  SRLX_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  SRLX_REQUIRE((int)so.total + 1024 <= max_smem, "network too large for the PPO rollout kernel: needs %zu bytes of shared memory", so.total);
  int per = (ppo->env.n_envs + n_sm - 1) / n_sm;
  per = round_up(per < kRowTile ? kRowTile : per, kRowTile);
  const int grid = (ppo->env.n_envs + per - 1) / per;
  SRLX_CHECK_CUDA(cudaFuncSetAttribute(ppo_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)so.total));
  ppo_rollout_kernel<<<grid, kPpoThreads, so.total, (cudaStream_t)cuda_stream>>>(*ppo, per, training);
  ppo_step_count_kernel<<<1, 1, 0, (cudaStream_t)cuda_stream>>>(ppo->env.state, ppo->env.n_envs);
  count_launch(2);
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// V(s) of n states (row-major [n][D]) with the current parameters
extern "C" int srlx_ppo_values(const srlx_ppo* ppo, const float* obs_dev, uint64_t n, float* out_dev, uintptr_t cuda_stream) {
  using namespace srlx;
  if (int rc = ppo_check(ppo)) return rc;
  SRLX_REQUIRE(obs_dev && out_dev, "srlx_ppo_values: NULL buffer");
  if (n == 0) return 0;
  const NetPlan pv = make_plan(ppo->net_v), pp = make_plan(ppo->net_p);
  const PpoSmem so = ppo_smem(*ppo, pv, pp);
  SRLX_CHECK_CUDA(cudaFuncSetAttribute(ppo_value_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)so.total));
  const uint64_t tiles = (n + kRowTile - 1) / kRowTile;
  const unsigned grid = (unsigned)(tiles < 592 ? tiles : 592);
  ppo_value_kernel<<<grid, kPpoThreads, so.total, (cudaStream_t)cuda_stream>>>(*ppo, obs_dev, (long long)n, out_dev);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// the worker's end-of-episode work for the whole rollout buffer: V(s) with the current parameters, then GAE / Monte-Carlo returns
// (srlx_returns_scan) into buf_ret, buf_valid (steps of episodes still running at the end of the buffer are not emitted)
extern "C" int srlx_ppo_finish_rollout(const srlx_ppo* ppo, uintptr_t cuda_stream) {
  using namespace srlx;
  if (int rc = ppo_check(ppo)) return rc;
  SRLX_REQUIRE(ppo->buf_vnew && ppo->buf_ret && ppo->buf_valid, "rollout buffer pointer is NULL");
  const uint64_t n = (uint64_t)ppo->horizon * ppo->env.n_envs;
  if (ppo->method == SRLX_RETURNS_GAE) {
    if (int rc = srlx_ppo_values(ppo, ppo->buf_obs, n, ppo->buf_vnew, cuda_stream)) return rc;
  }
  return srlx_returns_scan(ppo->buf_reward, nullptr, ppo->buf_vnew, ppo->buf_vnew + ppo->env.n_envs, ppo->buf_done, ppo->buf_ret, ppo->buf_valid,
                           (uint32_t)ppo->horizon, (uint32_t)ppo->env.n_envs, ppo->discount, ppo->gae_discount, ppo->method, 0,
                           ppo->reward_clip_enable, ppo->reward_clip_lo, ppo->reward_clip_hi, cuda_stream);
}

// n_updates x Trainer._train on the finished rollout buffer
extern "C" int srlx_ppo_learn(const srlx_ppo* ppo, uint32_t n_updates, uintptr_t cuda_stream) {
  using namespace srlx;
  if (int rc = ppo_check(ppo)) return rc;
  SRLX_REQUIRE(ppo->adam_m && ppo->adam_v && ppo->buf_ret && ppo->buf_valid && ppo->buf_obs, "srlx_ppo_learn: buffer pointer is NULL");
  if (n_updates == 0) return 0;
  const NetPlan pv = make_plan(ppo->net_v), pp = make_plan(ppo->net_p);
  int dev = 0, max_smem = 0;
  SRLX_CHECK_CUDA(cudaGetDevice(&dev));
  SRLX_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  PpoUpdSmem so = ppo_upd_smem(*ppo, pv, pp, true);
  float* g_ext = nullptr;
  if ((long long)so.total + 1024 > max_smem) {  // the gradient accumulator moves to global memory (L2-resident)
    SRLX_REQUIRE(ppo->grad_scratch != nullptr, "srlx_ppo_learn: this network needs grad_scratch [n_params]");
    g_ext = ppo->grad_scratch;
    so = ppo_upd_smem(*ppo, pv, pp, false);
  }
  SRLX_REQUIRE((long long)so.total + 1024 <= max_smem, "network too large for the PPO update kernel: needs %zu bytes of shared memory", so.total);
  SRLX_CHECK_CUDA(cudaFuncSetAttribute(ppo_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)so.total));
  ppo_update_kernel<<<1, kPpoUpdThreads, so.total, (cudaStream_t)cuda_stream>>>(*ppo, n_updates, g_ext);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
