// learner.cu -- the trainer inner step as ONE persistent kernel: n_updates consecutive Trainer.train() calls without
// returning to the host (consecutive updates are data-dependent: weights_t -> weights_{t+1}, so the limiter is the
// dependent-step latency, not a roofline).
//
// Per update (reference lines in brackets):
//   1. PER sample: beta, B descents, IS weights  [priority_replay_buffer.py:228-244, proportional_memory.py:131-169]
//      or uniform distinct sample                 [priority_memories/replay_buffer.py:34-36]
//   2. gather the M+1 windows from the ring        [rainbow.py:373-400 / dqn.py:229-246 records]
//   3. target-net and online-net forwards on s'    [dqn.py:144-176, rainbow.py:185-287, rainbow_nomultisteps.py:10-43]
//   4. n-step / Retrace / double-DQN target        [rainbow.py:232-285]
//   5. online forward on s, Huber(target*w, q*w)   [dqn/model_torch.py:113-115, rainbow/model_torch.py:103-105]
//   6. backward, Adam                              [model_torch.py:117-119; torch.optim.Adam defaults]
//   7. priorities |target-q| -> tree update        [model_torch.py:122-123, proportional_memory.py:171-177]
//   8. hard target sync when train_count % interval == 0, train_count += 1  [model_torch.py:126-132]
// CPU twin: oracle/engine.py::OracleEngine.learn.
#include "net.cuh"
#include "tree.cuh"

namespace srlx {

constexpr int kLearnerThreads = 512;

struct LearnerSmem {
  // byte offsets into dynamic shared memory
  size_t s_idx, s_pri, s_chg, s_prinew, weff, acts, G, qtrain, dq, qon, qtg, s_w, slot, src_slot, w_act, w_rew, w_term,
      tq, qsa, total;
};

__host__ __device__ inline LearnerSmem learner_smem(const srlx_engine& eng, const NetPlan& pl) {
  LearnerSmem s;
  const size_t B = eng.batch_size, M = eng.multisteps, A = eng.n_actions;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) / 16 * 16; return o; };
  s.s_idx = take(B * 8);
  s.s_pri = take(B * 8);
  s.s_chg = take(B * 8);
  s.s_prinew = take(B * 8);
  s.weff = take((size_t)pl.weff_floats * 4);
  s.acts = take((size_t)pl.act_floats * 4);
  s.G = take((size_t)eng.net.n_params * 4);
  s.qtrain = take((size_t)kRowTile * A * 4);
  s.dq = take((size_t)kRowTile * A * 4);
  s.qon = take(B * M * A * 4);
  s.qtg = take(B * M * A * 4);
  s.s_w = take(B * 4);
  s.slot = take(B * 4);
  s.src_slot = take(B * M * 4);
  s.w_act = take(B * M * 4);
  s.w_rew = take(B * M * 4);
  s.w_term = take(B * M * 4);
  s.tq = take(B * 4);
  s.qsa = take(B * 4);
  s.total = off;
  return s;
}

struct LearnerScalars {
  uint64_t tc, mem_size, vec_steps, adam_step, retries;
  double total, max_priority, beta, loss;
  float step_size, bc2_sqrt;
  int go;
};

__device__ inline void adam_apply(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, float g, float b1,
                                  float b2, float eps, float step_size, float bc2_sqrt) {
  // torch/optim/adam.py _single_tensor_adam: lerp, mul_/addcmul_, sqrt/div/add_, addcdiv_
  float mm = __ldcg(m), vv = __ldcg(v), pp = __ldcg(p);
  mm = mm + (g - mm) * (1.0f - b1);
  vv = vv * b2 + (1.0f - b2) * g * g;
  const float denom = sqrtf(vv) / bc2_sqrt + eps;
  pp = pp - step_size * (mm / denom);
  __stcg(m, mm);
  __stcg(v, vv);
  __stcg(p, pp);
}

__global__ void __launch_bounds__(kLearnerThreads, 1)
learner_kernel(const __grid_constant__ srlx_engine eng, const uint32_t n_updates) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ LearnerScalars sc;
  const srlx_net& net = eng.net;
  const NetPlan pl = make_plan(net);
  const LearnerSmem so = learner_smem(eng, pl);
  int64_t* s_idx = reinterpret_cast<int64_t*>(smem_raw + so.s_idx);
  double* s_pri = reinterpret_cast<double*>(smem_raw + so.s_pri);
  double* s_chg = reinterpret_cast<double*>(smem_raw + so.s_chg);
  double* s_prinew = reinterpret_cast<double*>(smem_raw + so.s_prinew);
  float* weff = reinterpret_cast<float*>(smem_raw + so.weff);
  float* acts = reinterpret_cast<float*>(smem_raw + so.acts);
  float* G = reinterpret_cast<float*>(smem_raw + so.G);
  float* qtrain = reinterpret_cast<float*>(smem_raw + so.qtrain);
  float* dQ = reinterpret_cast<float*>(smem_raw + so.dq);
  float* qon = reinterpret_cast<float*>(smem_raw + so.qon);
  float* qtg = reinterpret_cast<float*>(smem_raw + so.qtg);
  float* s_w = reinterpret_cast<float*>(smem_raw + so.s_w);
  int* slot = reinterpret_cast<int*>(smem_raw + so.slot);
  int* src_slot = reinterpret_cast<int*>(smem_raw + so.src_slot);
  int* w_act = reinterpret_cast<int*>(smem_raw + so.w_act);
  float* w_rew = reinterpret_cast<float*>(smem_raw + so.w_rew);
  float* w_term = reinterpret_cast<float*>(smem_raw + so.w_term);
  float* tq = reinterpret_cast<float*>(smem_raw + so.tq);
  float* qsa = reinterpret_cast<float*>(smem_raw + so.qsa);

  const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
  const int B = eng.batch_size, M = eng.multisteps, A = eng.n_actions, D = eng.obs_dim, E = eng.n_envs, R = eng.ring_rows;
  const int64_t cap = (int64_t)R * E, n_nodes = 2 * cap - 1;
  const bool per = eng.mem_kind == SRLX_MEM_PROPORTIONAL;
  const bool noisy = net.noisy != 0;
  const bool need_online_next = eng.enable_double_dqn || M > 1;
  srlx_state* st = eng.state;

  zero_floats(weff, pl.weff_floats);
  zero_floats(acts, pl.act_floats);
  __syncthreads();

  for (uint32_t upd = 0; upd < n_updates; ++upd) {
    // ---------------------------------------------------------------- 0. scalars
    if (tid == 0) {
      sc.tc = st->train_count;
      sc.mem_size = st->mem_size;
      sc.vec_steps = st->vec_steps;
      sc.adam_step = st->adam_step;
      sc.max_priority = st->max_priority;
      sc.retries = 0;
      sc.go = (sc.mem_size >= eng.warmup_size && sc.mem_size >= (uint64_t)B) ? 1 : 0;
      if (per) {
        sc.total = __ldcg(eng.tree);
        // PriorityReplayBuffer.step is the train_count of the PREVIOUS update() call (priority_replay_buffer.py:232,250)
        const double step = (sc.tc > 0) ? (double)(sc.tc - 1) : 0.0;
        double beta = eng.per_beta_initial + (1.0 - eng.per_beta_initial) * step / eng.per_beta_steps;
        sc.beta = beta > 1.0 ? 1.0 : beta;
      }
      const double t = (double)(sc.adam_step + 1);
      const double bc1 = 1.0 - pow(eng.adam_beta1, t), bc2 = 1.0 - pow(eng.adam_beta2, t);
      sc.step_size = (float)(eng.lr / bc1);
      sc.bc2_sqrt = (float)sqrt(bc2);
    }
    __syncthreads();
    if (!sc.go) break;  // still warming up: train() returns without incrementing train_count
    const uint64_t tc = sc.tc;

    // ---------------------------------------------------------------- 1. sample
    if (per) {
      per_sample_block(eng.tree, n_nodes, sc.total, B, eng.seed, tc, nullptr, 9999, eng.has_duplicate, s_idx, s_pri, s_chg,
                       (unsigned long long*)&sc.retries);
      per_weights_block(sc.total, (double)sc.mem_size, sc.beta, B, s_pri, s_chg, s_w);
      for (int i = tid; i < B; i += nt) slot[i] = (int)(s_idx[i] - (cap - 1));
    } else {
      if (tid == 0) {
        const uint64_t g_next = sc.vec_steps;
        const uint64_t g_lo = g_next > (uint64_t)R ? g_next - R : 0;
        const uint64_t n_g = g_next - (uint64_t)(M - 1) - g_lo;
        const uint32_t n_valid = (uint32_t)(n_g * E);
        for (int i = 0; i < B; ++i) {
          uint32_t pick = 0;
          for (int k = 0; k < 65536; ++k) {
            const uint4 w = philox(eng.seed, STREAM_UNIFORM_SAMPLE, (uint32_t)i | ((uint32_t)k << 16), (uint32_t)tc, (uint32_t)(tc >> 32));
            pick = u_below(w.x, n_valid);
            bool dup = false;
            for (int j = 0; j < i; ++j) dup |= (s_idx[j] == (int64_t)pick);
            if (!dup) break;
          }
          s_idx[i] = pick;
        }
        for (int i = 0; i < B; ++i) {
          const uint64_t pick = (uint64_t)s_idx[i];
          const uint64_t g = g_lo + pick / E;
          slot[i] = (int)((g % R) * E + pick % E);
          s_idx[i] = slot[i];
          s_w[i] = 1.0f;
        }
      }
    }
    __syncthreads();

    // ---------------------------------------------------------------- 2. gather windows (records only; states are
    // pulled straight into the activation buffers tile by tile)
    for (int i = tid; i < B; i += nt) {
      const int s0 = slot[i];
      const int rho = s0 / E, e = s0 - rho * E;
      const uint64_t g_last = sc.vec_steps - 1;
      const uint64_t g_item = g_last - ((g_last + (uint64_t)R - (uint64_t)rho) % (uint64_t)R);
      bool ended = false;
      int last_slot = s0;
      for (int k = 0; k < M; ++k) {
        if (!ended) {
          const int sk = ((rho + k) % R) * E + e;
          w_act[i * M + k] = __ldcg(eng.ring_action + sk);
          w_rew[i * M + k] = __ldcg(eng.ring_reward + sk);
          w_term[i * M + k] = (float)__ldcg(eng.ring_term + sk);
          src_slot[i * M + k] = sk;
          last_slot = sk;
          if (__ldcg(eng.ring_done + sk)) ended = true;
        } else {
          // padded tail record: random action, reward 0, terminated 1, state = last next_state (rainbow.py:358-371)
          const uint64_t gp = g_item + (uint64_t)k;
          const uint4 w = philox(eng.seed, STREAM_PAD_ACTION, (uint32_t)e, (uint32_t)gp, (uint32_t)(gp >> 32));
          w_act[i * M + k] = (int)u_below(w.x, (uint32_t)A);
          w_rew[i * M + k] = 0.f;
          w_term[i * M + k] = 1.f;
          src_slot[i * M + k] = last_slot;
        }
      }
    }
    __syncthreads();

    // ---------------------------------------------------------------- 3. forwards on the next states
    const int n_next = B * M;
    for (int pass = 2; pass >= 1; --pass) {  // 2: target net, 1: online net
      if (pass == 1 && !need_online_next) continue;
      const float* mu = (pass == 2) ? eng.target : eng.params;
      const float* sg = (pass == 2) ? eng.target_sigma : eng.params_sigma;
      build_weff(net, pl, mu, sg, noisy, eng.seed, NOISE_KIND_TRAIN, tc * 3 + pass, weff);
      float* qdst = (pass == 2) ? qtg : qon;
      for (int r0 = 0; r0 < n_next; r0 += kRowTile) {
        const int Rr = min(kRowTile, n_next - r0);
        for (int w = tid; w < Rr * D; w += nt) {
          const int r = w / D, d = w - r * D;
          acts[pl.x_s[0] + r * pl.ldx[0] + d] = __ldcg(eng.ring_next_obs + (size_t)src_slot[r0 + r] * D + d);
        }
        __syncthreads();
        net_forward_tile(net, pl, weff, acts, Rr, qdst + (size_t)r0 * A, A);
      }
    }

    // ---------------------------------------------------------------- 4. targets (thread per sample)
    for (int i = tid; i < B; i += nt) {
      const float gamma = (float)eng.discount;
      float target = 0.f, retrace = 1.f;
      int greedy_next = 0;  // n_act_idx[k]
      for (int k = 0; k < M; ++k) {
        const float* qo = qon + (size_t)(i * M + k) * A;
        const float* qt = qtg + (size_t)(i * M + k) * A;
        const float* qsel = eng.enable_double_dqn ? qo : qt;
        int am = 0;
        float best = qsel[0];
        for (int a = 1; a < A; ++a)
          if (qsel[a] > best) { best = qsel[a]; am = a; }  // np.argmax: first max wins
        if (k >= 1) {
          // Retrace with the reference's index shift (rainbow.py:267): action taken at s_k vs greedy action at s_{k+1}
          retrace = retrace * ((float)eng.retrace_h * ((w_act[i * M + k] == am) ? 1.f : 0.f));
        }
        greedy_next = am;
        float maxq = qt[greedy_next];
        if (eng.enable_rescale) maxq = inverse_rescaling_f(maxq);
        float gain = w_rew[i * M + k] + ((1.0f - w_term[i * M + k]) * gamma) * maxq;
        if (eng.enable_rescale) gain = rescaling_f(gain);
        float qk = 0.f;  // the first step is learnt by the trainer itself (rainbow.py:232-234)
        if (k >= 1) qk = qon[(size_t)(i * M + k - 1) * A + w_act[i * M + k]];
        const float td = gain - qk;
        target += (td * (float)pow(eng.discount, (double)k)) * retrace;
      }
      tq[i] = target;
    }
    __syncthreads();

    // ---------------------------------------------------------------- 5./6. online forward on s, loss, backward
    build_weff(net, pl, eng.params, eng.params_sigma, noisy, eng.seed, NOISE_KIND_TRAIN, tc * 3 + 0, weff);
    zero_floats(G, net.n_params);
    if (tid == 0) sc.loss = 0.0;
    __syncthreads();
    for (int r0 = 0; r0 < B; r0 += kRowTile) {
      const int Rr = min(kRowTile, B - r0);
      for (int w = tid; w < Rr * D; w += nt) {
        const int r = w / D, d = w - r * D;
        acts[pl.x_s[0] + r * pl.ldx[0] + d] = __ldcg(eng.ring_obs + (size_t)slot[r0 + r] * D + d);
      }
      __syncthreads();
      net_forward_tile(net, pl, weff, acts, Rr, qtrain, A);
      if (tid < 32) {
        float lsum = 0.f;
        for (int r = lane; r < Rr; r += 32) {
          const int i = r0 + r;
          const int a0 = w_act[i * M + 0];
          const float q = qtrain[r * A + a0];
          qsa[i] = q;
          const float w = s_w[i];
          const float d = q * w - tq[i] * w;
          const float ad = fabsf(d);
          const float delta = (float)eng.huber_delta;
          lsum += (ad <= delta) ? 0.5f * d * d : delta * (ad - 0.5f * delta);
          const float dq = fminf(fmaxf(d, -delta), delta) * w / (float)B;
          for (int a = 0; a < A; ++a) dQ[r * A + a] = (a == a0) ? dq : 0.f;
        }
        for (int s = 16; s > 0; s >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, s);
        if (lane == 0) sc.loss += (double)lsum;
      }
      __syncthreads();
      net_backward_tile(net, pl, weff, acts, Rr, dQ, A, G);
    }

    // ---------------------------------------------------------------- Adam (mu, then sigma = dW_eff * eps)
    {
      const float b1 = (float)eng.adam_beta1, b2 = (float)eng.adam_beta2, eps = (float)eng.adam_eps;
      const int nblk = (net.n_params + 3) >> 2;
      for (int blk = tid; blk < nblk; blk += nt) {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (noisy) {
          const float4 n4 = noise4(eng.seed, NOISE_KIND_TRAIN, tc * 3 + 0, (uint32_t)blk);
          z[0] = n4.x; z[1] = n4.y; z[2] = n4.z; z[3] = n4.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int p = 4 * blk + j;
          if (p >= net.n_params) break;
          const float g = G[p];
          adam_apply(eng.params + p, eng.adam_m + p, eng.adam_v + p, g, b1, b2, eps, sc.step_size, sc.bc2_sqrt);
          if (eng.dbg_grads) eng.dbg_grads[p] = g;
          if (noisy) {
            const int l = layer_of_param(net, p);
            const float gs = net.layer_noisy[l] ? g * z[j] : 0.f;
            if (net.layer_noisy[l])
              adam_apply(eng.params_sigma + p, eng.adam_m + net.n_params + p, eng.adam_v + net.n_params + p, gs, b1, b2,
                         eps, sc.step_size, sc.bc2_sqrt);
            if (eng.dbg_grads) eng.dbg_grads[net.n_params + p] = gs;
          }
        }
      }
    }

    // ---------------------------------------------------------------- 7. priorities -> tree
    if (per) {
      for (int i = tid; i < B; i += nt) {
        const double pr = pow(fabs((double)fabsf(tq[i] - qsa[i])) + eng.per_epsilon, eng.per_alpha);
        s_prinew[i] = pr;
      }
      __syncthreads();
      tree_update_batch(eng.tree, s_idx, s_prinew, s_chg, B);
      if (tid == 0) {
        double mp = sc.max_priority;
        for (int i = 0; i < B; ++i) mp = (mp < s_prinew[i]) ? s_prinew[i] : mp;
        sc.max_priority = mp;
      }
    }
    // debug taps
    if (eng.dbg_sample_idx)
      for (int i = tid; i < B; i += nt) eng.dbg_sample_idx[i] = s_idx[i];
    if (eng.dbg_weights)
      for (int i = tid; i < B; i += nt) eng.dbg_weights[i] = s_w[i];
    if (eng.dbg_target_q)
      for (int i = tid; i < B; i += nt) eng.dbg_target_q[i] = tq[i];
    if (eng.dbg_q_sa)
      for (int i = tid; i < B; i += nt) eng.dbg_q_sa[i] = qsa[i];
    if (eng.dbg_windows) {
      float* dw = eng.dbg_windows;
      const size_t n_states = (size_t)B * (M + 1) * D;
      for (int w = tid; w < B * (M + 1) * D; w += nt) {
        const int i = w / ((M + 1) * D), rem = w - i * (M + 1) * D, k = rem / D, d = rem - k * D;
        dw[w] = (k == 0) ? __ldcg(eng.ring_obs + (size_t)slot[i] * D + d)
                         : __ldcg(eng.ring_next_obs + (size_t)src_slot[i * M + k - 1] * D + d);
      }
      for (int w = tid; w < B * M; w += nt) {
        dw[n_states + w] = (float)w_act[w];
        dw[n_states + B * M + w] = w_rew[w];
        dw[n_states + 2 * B * M + w] = w_term[w];
      }
    }
    __syncthreads();  // Adam writes visible to the whole block before the sync copy / next build_weff

    // ---------------------------------------------------------------- 8. target sync, counters
    const bool sync = (tc % (uint64_t)eng.target_update_interval) == 0;
    if (sync) {
      for (int p = tid; p < net.n_params; p += nt) {
        __stcg(eng.target + p, __ldcg(eng.params + p));
        if (noisy) __stcg(eng.target_sigma + p, __ldcg(eng.params_sigma + p));
      }
    }
    if (tid == 0) {
      const double loss = sc.loss / (double)B;
      st->train_count = tc + 1;
      st->adam_step = sc.adam_step + 1;
      st->max_priority = sc.max_priority;
      st->sample_retries += sc.retries;
      st->last_loss = loss;
      st->loss_sum += loss;
      if (sync) st->sync_count += 1;
      __threadfence();
    }
    __syncthreads();
  }
}

}  // namespace srlx

// ---- host side ---------------------------------------------------------------------------------------------------
extern "C" int srlx_learn(const srlx_engine* eng, uint32_t n_updates, uintptr_t cuda_stream) {
  using namespace srlx;
  SRLX_REQUIRE(eng != nullptr, "srlx_learn: eng is NULL");
  SRLX_REQUIRE(eng->batch_size >= 1 && eng->batch_size <= SRLX_MAX_BATCH, "batch_size %d out of range [1,%d]", eng->batch_size, SRLX_MAX_BATCH);
  SRLX_REQUIRE(eng->multisteps >= 1 && eng->multisteps <= SRLX_MAX_MULTISTEPS, "multisteps %d out of range", eng->multisteps);
  SRLX_REQUIRE(eng->n_actions >= 1 && eng->n_actions <= SRLX_MAX_ACTIONS, "n_actions %d out of range", eng->n_actions);
  SRLX_REQUIRE(eng->mem_kind == SRLX_MEM_UNIFORM || eng->tree != nullptr, "proportional memory needs a tree buffer");
  SRLX_REQUIRE(!eng->net.noisy || (eng->params_sigma && eng->target_sigma), "noisy net needs sigma buffers");
  if (n_updates == 0) return 0;
  const NetPlan pl = make_plan(eng->net);
  const LearnerSmem so = learner_smem(*eng, pl);
  int dev = 0, max_smem = 0;
  SRLX_CHECK_CUDA(cudaGetDevice(&dev));
  SRLX_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  SRLX_REQUIRE((int)so.total + 1024 <= max_smem,
               "network / batch too large for the fused learner: needs %zu bytes of shared memory, device allows %d",
               so.total + 1024, max_smem);
  SRLX_CHECK_CUDA(cudaFuncSetAttribute(learner_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)so.total));
  learner_kernel<<<1, kLearnerThreads, so.total, (cudaStream_t)cuda_stream>>>(*eng, n_updates);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
