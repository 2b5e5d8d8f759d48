// learner_small.cu -- Trainer.train() for small plain MLPs on uniform replay, ONE thread block, everything in shared memory.
//
// BASELINE configs[1] (DQN, MLP[64,64], uniform replay 1M) has ~4.6k parameters: the cluster kernels (learner.cu,
// learner_fast.cu) shard a wide layer over 8..16 SMs and pay a DSMEM exchange per layer boundary, which for a net this
// small is all latency and no work (37 us per update on the generic kernel).  Here the online weights, the target
// weights, both Adam moments and the activations of every row tile stay in the shared memory of one SM for the whole
// launch; an update is a short chain of block-wide phases:
//
//   1. sample      B distinct uniform picks (replay_buffer.py:34-36): attempt 0 of all B draws in parallel, duplicates
//                  resolved in sample order (the result equals the sequential rejection loop of oracle/sumtree.py)
//   2. gather      the (M+1)-state windows from the ring, padded tails rebuilt (rainbow.py:358-371)
//   3. forward     online(s), online(s'), target(s'): all row tiles of a layer run concurrently (net.cuh warp tiles)
//   4. targets     double-DQN / n-step Retrace target, Huber gradient (thread per sample; same code as learner.cu)
//   5. backward    net_backward_tile on the s rows
//   6. Adam        torch _single_tensor_adam arithmetic (adam_apply), hard target sync when train_count % interval == 0
//
// Reference path: srl/algorithms/dqn/model_torch.py:90-132, rainbow/model_torch.py:85-122, rainbow/rainbow.py:185-287.
// CPU twin: oracle/engine.py::learn.  Applies to: uniform replay, no NoisyNet, any depth / dueling head that fits.
#include "cluster.cuh"
#include "net.cuh"

namespace srlx {

constexpr int kSmThreads = 512;

struct SPlan {
  NetPlan np;
  int B, M, A, D, BM, P, P4;
  int n_on_rows, n_on_tiles, n_tg_tiles, n_tiles, n_s_tiles;
  size_t off_weff, off_wefft, off_m, off_v, off_g, off_slot, off_acts, off_q, off_dq, off_pick, off_win, off_tq, off_red,
      off_scal, total;
};

struct SScal {
  double loss_sum, last_loss;
  float step_size, bc2_sqrt;
  unsigned int sync_count;
};

__host__ __device__ inline SPlan make_splan(const srlx_engine& eng) {
  SPlan p;
  p.np = make_plan(eng.net);
  p.B = eng.batch_size;
  p.M = eng.multisteps;
  p.A = eng.n_actions;
  p.D = eng.obs_dim;
  p.BM = p.B * p.M;
  p.P = eng.net.n_params;
  p.P4 = round_up(p.P, 4);
  const bool need_online_next = eng.enable_double_dqn || p.M > 1;
  p.n_on_rows = p.B + (need_online_next ? p.BM : 0);  // online rows: [0,B) = s, [B,B+BM) = s'_k
  p.n_on_tiles = (p.n_on_rows + kRowTile - 1) / kRowTile;
  p.n_tg_tiles = (p.BM + kRowTile - 1) / kRowTile;   // target rows: s'_k
  p.n_tiles = p.n_on_tiles + p.n_tg_tiles;
  p.n_s_tiles = (p.B + kRowTile - 1) / kRowTile;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) / 16 * 16; return r; };
  p.off_weff = take((size_t)p.np.weff_floats * 4);
  p.off_wefft = take((size_t)p.np.weff_floats * 4);
  p.off_m = take((size_t)p.P4 * 4);
  p.off_v = take((size_t)p.P4 * 4);
  p.off_g = take((size_t)p.P4 * 4);
  p.off_slot = take((size_t)p.P4 * 2);
  p.off_acts = take((size_t)p.n_tiles * p.np.act_floats * 4);
  p.off_q = take((size_t)(p.B + 2 * p.BM) * p.A * 4);  // Q(s) [B][A], online Q(s') [BM][A], target Q(s') [BM][A]
  p.off_dq = take((size_t)p.B * p.A * 4);
  p.off_pick = take((size_t)p.B * 4 * 2);              // picks, slots
  p.off_win = take((size_t)p.BM * 4 * 4);              // action, reward, term, done of every window step
  p.off_tq = take((size_t)p.B * 4 * 2);                // target, q(s,a)
  p.off_red = take(64 * 4);
  p.off_scal = take(64);
  p.total = o;
  return p;
}

__host__ inline bool small_shape_ok(const srlx_engine& eng) {
  return eng.mem_kind == SRLX_MEM_UNIFORM && !eng.net.noisy && eng.net.n_layers >= 2 && eng.net.n_params < 65536 &&
         eng.obs_dim <= SRLX_MAX_OBS;
}

__device__ __forceinline__ void adam_apply(float& pp, float& mm, float& vv, float g, float b1, float b2, float eps,
                                           float step_size, float bc2_sqrt) {
  // torch/optim/adam.py _single_tensor_adam: lerp, mul_/addcmul_, sqrt/div/add_, addcdiv_
  mm = mm + (g - mm) * (1.0f - b1);
  vv = vv * b2 + (1.0f - b2) * g * g;
  const float denom = sqrtf(vv) / bc2_sqrt + eps;
  pp = pp - step_size * (mm / denom);
}

__global__ void __launch_bounds__(kSmThreads, 1)
learner_small_kernel(const __grid_constant__ srlx_engine eng, const uint32_t n_updates) {
  extern __shared__ __align__(16) unsigned char smem[];
  const srlx_net& net = eng.net;
  const SPlan pl = make_splan(eng);
  const NetPlan& np = pl.np;
  float* weff = reinterpret_cast<float*>(smem + pl.off_weff);
  float* wefft = reinterpret_cast<float*>(smem + pl.off_wefft);
  float* am = reinterpret_cast<float*>(smem + pl.off_m);
  float* av = reinterpret_cast<float*>(smem + pl.off_v);
  float* G = reinterpret_cast<float*>(smem + pl.off_g);
  unsigned short* pslot = reinterpret_cast<unsigned short*>(smem + pl.off_slot);
  float* acts = reinterpret_cast<float*>(smem + pl.off_acts);
  float* Q = reinterpret_cast<float*>(smem + pl.off_q);
  float* dQ = reinterpret_cast<float*>(smem + pl.off_dq);
  int* pick = reinterpret_cast<int*>(smem + pl.off_pick);
  int* slot = pick + pl.B;
  int* w_act = reinterpret_cast<int*>(smem + pl.off_win);
  float* w_rew = reinterpret_cast<float*>(smem + pl.off_win) + pl.BM;
  float* w_term = w_rew + pl.BM;
  int* w_done = reinterpret_cast<int*>(w_term + pl.BM);
  float* tq = reinterpret_cast<float*>(smem + pl.off_tq);
  float* qsa = tq + pl.B;
  float* red = reinterpret_cast<float*>(smem + pl.off_red);
  SScal* sc = reinterpret_cast<SScal*>(smem + pl.off_scal);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kSmThreads >> 5;
  const int B = pl.B, M = pl.M, A = pl.A, D = pl.D, BM = pl.BM, E = eng.n_envs, R = eng.ring_rows, P = pl.P;
  const int L = net.n_layers;
  const bool need_online_next = eng.enable_double_dqn || M > 1;
  srlx_state* st = eng.state;
  const uint64_t tc0 = st->train_count, mem_size = st->mem_size, vec_steps = st->vec_steps, adam0 = st->adam_step;
  if (!(mem_size >= eng.warmup_size && mem_size >= (uint64_t)B)) return;  // warming up: train() returns, no count

  // ---- one-time setup: parameters, moments and target into shared memory --------------------------------------------
  for (int i = tid; i < np.weff_floats; i += kSmThreads) { weff[i] = 0.f; wefft[i] = 0.f; }
  for (int i = tid; i < pl.n_tiles * np.act_floats; i += kSmThreads) acts[i] = 0.f;
  for (int i = tid; i < pl.P4; i += kSmThreads) G[i] = 0.f;
  if (tid == 0) { sc->loss_sum = 0.0; sc->last_loss = 0.0; sc->sync_count = 0; }
  __syncthreads();
  for (int p = tid; p < P; p += kSmThreads) {
    const int s = weff_slot(net, np, p, layer_of_param(net, p));
    pslot[p] = (unsigned short)s;
    weff[s] = __ldcg(eng.params + p);
    wefft[s] = __ldcg(eng.target + p);
    am[p] = __ldcg(eng.adam_m + p);
    av[p] = __ldcg(eng.adam_v + p);
  }
  __syncthreads();

  // row r of the online set / target set -> its place in the tile activation areas
  auto x_row = [&](int set_tile0, int r) -> float* {
    return acts + (size_t)(set_tile0 + r / kRowTile) * np.act_floats + np.x_s[0] + (r % kRowTile) * np.ldx[0];
  };
  const uint64_t g_next = vec_steps;
  const uint64_t g_lo = g_next > (uint64_t)R ? g_next - R : 0;
  const uint32_t n_valid = (uint32_t)((g_next - (uint64_t)(M - 1) - g_lo) * E);
  const float b1 = (float)eng.adam_beta1, b2 = (float)eng.adam_beta2, aeps = (float)eng.adam_eps;

  for (uint32_t upd = 0; upd < n_updates; ++upd) {
    const uint64_t tc = tc0 + upd;
    // ---------------------------------------------------------------- 1. sample
    if (warp == 0) {
      for (int i0 = 0; i0 < B; i0 += 32) {
        const int i = i0 + lane;
        if (i < B) {
          const uint4 w = philox(eng.seed, STREAM_UNIFORM_SAMPLE, (uint32_t)i, (uint32_t)tc, (uint32_t)(tc >> 32));
          pick[i] = (int)u_below(w.x, n_valid);
        }
      }
      __syncwarp();
      if (lane == 0) {
        for (int i = 1; i < B; ++i) {  // a pick that repeats an earlier one is redrawn (attempt k = 1, 2, ...), in order
          int k = 0;
          while (true) {
            bool dup = false;
            for (int j = 0; j < i; ++j) dup |= (pick[j] == pick[i]);
            if (!dup || ++k >= 65536) break;
            const uint4 w = philox(eng.seed, STREAM_UNIFORM_SAMPLE, (uint32_t)i | ((uint32_t)k << 16), (uint32_t)tc, (uint32_t)(tc >> 32));
            pick[i] = (int)u_below(w.x, n_valid);
          }
        }
      }
      __syncwarp();
      for (int i = lane; i < B; i += 32) {
        const uint64_t pk = (uint64_t)pick[i];
        const uint64_t g = g_lo + pk / E;
        slot[i] = (int)((g % R) * E + pk % E);
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- 2. gather (all loads of a record before its stores)
    for (int w = tid; w < BM; w += kSmThreads) {
      const int i = w / M, k = w - i * M;
      const int s0 = slot[i];
      const int rho = s0 / E, e = s0 - rho * E;
      const int sk = ((rho + k) % R) * E + e;
      const int a = __ldcg(eng.ring_action + sk);
      const float rw = __ldcg(eng.ring_reward + sk);
      const unsigned char tm = __ldcg(eng.ring_term + sk), dn = __ldcg(eng.ring_done + sk);
      float xv[SRLX_MAX_OBS];
#pragma unroll
      for (int d = 0; d < SRLX_MAX_OBS; ++d) xv[d] = d < D ? __ldcg(eng.ring_next_obs + (size_t)sk * D + d) : 0.f;
      w_act[w] = a;
      w_rew[w] = rw;
      w_term[w] = (float)tm;
      w_done[w] = (int)dn;
      float* xt = x_row(pl.n_on_tiles, w);
#pragma unroll
      for (int d = 0; d < SRLX_MAX_OBS; ++d)
        if (d < D) xt[d] = xv[d];
    }
    for (int w = tid; w < B * D; w += kSmThreads) {
      const int i = w / D, d = w - i * D;
      x_row(0, i)[d] = __ldcg(eng.ring_obs + (size_t)slot[i] * D + d);
    }
    __syncthreads();
    // padded tails (rainbow.py:358-371), then the online copy of the next states
    for (int i = tid; i < B; i += kSmThreads) {
      const int s0 = slot[i];
      const int rho = s0 / E, e = s0 - rho * E;
      const uint64_t g_last = vec_steps - 1;
      const uint64_t g_item = g_last - ((g_last + (uint64_t)R - (uint64_t)rho) % (uint64_t)R);
      bool ended = false;
      int last_k = 0;
      for (int k = 0; k < M; ++k) {
        const int w = i * M + k;
        if (!ended) {
          last_k = k;
          if (w_done[w]) ended = true;
        } else {
          const uint64_t gp = g_item + (uint64_t)k;
          const uint4 pw = philox(eng.seed, STREAM_PAD_ACTION, (uint32_t)e, (uint32_t)gp, (uint32_t)(gp >> 32));
          w_act[w] = (int)u_below(pw.x, (uint32_t)A);
          w_rew[w] = 0.f;
          w_term[w] = 1.f;
          const float* src = x_row(pl.n_on_tiles, i * M + last_k);
          float* dst = x_row(pl.n_on_tiles, w);
          for (int d = 0; d < D; ++d) dst[d] = src[d];
        }
      }
    }
    __syncthreads();
    if (need_online_next)
      for (int w = tid; w < BM * D; w += kSmThreads) {
        const int r = w / D, d = w - r * D;
        x_row(0, B + r)[d] = x_row(pl.n_on_tiles, r)[d];
      }
    if (eng.dbg_sample_idx)
      for (int i = tid; i < B; i += kSmThreads) eng.dbg_sample_idx[i] = (int64_t)slot[i];
    if (eng.dbg_weights)
      for (int i = tid; i < B; i += kSmThreads) eng.dbg_weights[i] = 1.0f;
    if (eng.dbg_windows) {
      float* dw = eng.dbg_windows;
      const int n_states = B * (M + 1) * D;
      for (int w = tid; w < n_states; w += kSmThreads) {
        const int i = w / ((M + 1) * D), rem = w - i * (M + 1) * D, k = rem / D, d = rem - k * D;
        dw[w] = (k == 0) ? x_row(0, i)[d] : x_row(pl.n_on_tiles, i * M + k - 1)[d];
      }
      for (int w = tid; w < BM; w += kSmThreads) {
        dw[n_states + w] = (float)w_act[w];
        dw[n_states + BM + w] = w_rew[w];
        dw[n_states + 2 * BM + w] = w_term[w];
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- 3. forward, every tile of a layer concurrently
    auto tile_rows = [&](int t) -> int {
      const int rows = t < pl.n_on_tiles ? pl.n_on_rows - t * kRowTile : BM - (t - pl.n_on_tiles) * kRowTile;
      return rows < kRowTile ? rows : kRowTile;
    };
    for (int l = 0; l < L - 1; ++l) {
      const int U = net.out_dim[l], K = net.k_dim[l];
      const int n_ut = (U + 63) >> 6, n_rt = kRowTile / 4;
      for (int t = warp; t < pl.n_tiles * n_rt * n_ut; t += nwarps) {
        const int tile = t / (n_rt * n_ut), rem = t - tile * n_rt * n_ut, rt = rem / n_ut, ut = rem - rt * n_ut;
        const int Rt = tile_rows(tile);
        if (rt * 4 >= Rt) continue;
        const float* wset = tile < pl.n_on_tiles ? weff : wefft;
        float* a = acts + (size_t)tile * np.act_floats;
        dense_relu_task(a + np.x_s[l], np.ldx[l], Rt, K, wset + np.w_s[l], np.ldw[l], wset + np.b_s[l], U, a + np.x_s[l + 1],
                        np.ldx[l + 1], rt, ut);
      }
      __syncthreads();
    }
    for (int tile = 0; tile < pl.n_tiles; ++tile) {
      const float* wset = tile < pl.n_on_tiles ? weff : wefft;
      float* a = acts + (size_t)tile * np.act_floats;
      // Q rows: online set first ([0, n_on_rows)), target rows at B + BM
      float* q = tile < pl.n_on_tiles ? Q + (size_t)tile * kRowTile * A : Q + (size_t)(B + BM + (tile - pl.n_on_tiles) * kRowTile) * A;
      out_layer_fwd(net, a + np.x_s[L - 1], np.ldx[L - 1], tile_rows(tile), wset + np.w_s[L - 1], np.ldw[L - 1], wset + np.b_s[L - 1],
                    a + np.x_s[L], np.ldx[L], q, A);
    }
    __syncthreads();
    // ---------------------------------------------------------------- 4. targets, Huber gradient (thread per sample)
    {
      const float* qon = Q + (size_t)B * A;         // online(s')  [BM][A]
      const float* qtg = Q + (size_t)(B + BM) * A;  // target(s')  [BM][A]
      float lsum = 0.f;
      for (int i = tid; i < B; i += kSmThreads) {
        const float gamma = (float)eng.discount;
        float target = 0.f, retrace = 1.f;
        for (int k = 0; k < M; ++k) {
          const float* qo = qon + (size_t)(i * M + k) * A;
          const float* qt = qtg + (size_t)(i * M + k) * A;
          const float* qsel = eng.enable_double_dqn ? qo : qt;
          int amx = 0;
          float best = qsel[0];
          for (int a = 1; a < A; ++a)
            if (qsel[a] > best) { best = qsel[a]; amx = a; }  // np.argmax: first max wins
          // Retrace with the reference's index shift (rainbow.py:267): action taken at s_k vs greedy action at s_{k+1}
          if (k >= 1) retrace = retrace * ((float)eng.retrace_h * ((w_act[i * M + k] == amx) ? 1.f : 0.f));
          float maxq = qt[amx];
          if (eng.enable_rescale) maxq = inverse_rescaling_f(maxq);
          float gain = w_rew[i * M + k] + ((1.0f - w_term[i * M + k]) * gamma) * maxq;
          if (eng.enable_rescale) gain = rescaling_f(gain);
          float qk = 0.f;  // the first step is learnt by the trainer itself (rainbow.py:232-234)
          if (k >= 1) qk = qon[(size_t)(i * M + k - 1) * A + w_act[i * M + k]];
          const float td = gain - qk;
          target += (td * (float)pow(eng.discount, (double)k)) * retrace;
        }
        tq[i] = target;
        const int a0 = w_act[i * M + 0];
        const float q = Q[i * A + a0];
        qsa[i] = q;
        const float d = q - target;  // IS weight 1 on uniform replay (replay_buffer.py:37)
        const float ad = fabsf(d);
        const float delta = (float)eng.huber_delta;
        lsum += (ad <= delta) ? 0.5f * d * d : delta * (ad - 0.5f * delta);
        const float dq = fminf(fmaxf(d, -delta), delta) / (float)B;
        for (int a = 0; a < A; ++a) dQ[i * A + a] = (a == a0) ? dq : 0.f;
      }
      for (int s = 16; s > 0; s >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, s);
      if (lane == 0) red[warp] = lsum;
    }
    __syncthreads();
    if (tid == 0) {
      double l = 0.0;
      for (int w = 0; w < nwarps; ++w) l += (double)red[w];
      l /= (double)B;
      sc->last_loss = l;
      sc->loss_sum += l;
      const double t = (double)(adam0 + upd + 1);
      sc->step_size = (float)(eng.lr / (1.0 - pow(eng.adam_beta1, t)));
      sc->bc2_sqrt = (float)sqrt(1.0 - pow(eng.adam_beta2, t));
      if ((tc % (uint64_t)eng.target_update_interval) == 0) sc->sync_count += 1;
    }
    if (eng.dbg_target_q)
      for (int i = tid; i < B; i += kSmThreads) eng.dbg_target_q[i] = tq[i];
    if (eng.dbg_q_sa)
      for (int i = tid; i < B; i += kSmThreads) eng.dbg_q_sa[i] = qsa[i];
    // ---------------------------------------------------------------- 5. backward on the s rows (G was zeroed by Adam)
    for (int tile = 0; tile < pl.n_s_tiles; ++tile) {
      const int Rt = min(kRowTile, B - tile * kRowTile);
      net_backward_tile(net, np, weff, acts + (size_t)tile * np.act_floats, Rt, dQ + (size_t)tile * kRowTile * A, A, G);
    }
    __syncthreads();
    // ---------------------------------------------------------------- 6. Adam, target sync
    {
      const float step_size = sc->step_size, bc2_sqrt = sc->bc2_sqrt;
      const bool do_sync = (tc % (uint64_t)eng.target_update_interval) == 0;
      for (int p = tid; p < P; p += kSmThreads) {
        const int s = pslot[p];
        const float g = G[p];
        if (eng.dbg_grads) eng.dbg_grads[p] = g;
        float mu = weff[s], m = am[p], v = av[p];
        adam_apply(mu, m, v, g, b1, b2, aeps, step_size, bc2_sqrt);
        weff[s] = mu;
        am[p] = m;
        av[p] = v;
        G[p] = 0.f;
        if (do_sync) wefft[s] = mu;  // hard sync after the step, before train_count += 1 (model_torch.py:126-132)
      }
    }
    __syncthreads();
  }

  // ---- write the state back ----------------------------------------------------------------------------------------
  for (int p = tid; p < P; p += kSmThreads) {
    const int s = pslot[p];
    __stcg(eng.params + p, weff[s]);
    __stcg(eng.target + p, wefft[s]);
    __stcg(eng.adam_m + p, am[p]);
    __stcg(eng.adam_v + p, av[p]);
  }
  if (tid == 0) {
    st->train_count = tc0 + n_updates;
    st->adam_step = adam0 + n_updates;
    st->last_loss = sc->last_loss;
    st->loss_sum += sc->loss_sum;
    st->sync_count += sc->sync_count;
  }
}

// 1 when the single-block kernel applies to this engine and fits the shared memory of one SM
int small_choose(const srlx_engine* eng, size_t* smem_out) {
  if (!small_shape_ok(*eng)) return 0;
  int dev = 0, max_smem = 0;
  SRLX_CHECK_CUDA(cudaGetDevice(&dev));
  SRLX_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const SPlan pl = make_splan(*eng);
  if ((long long)pl.total + 1024 > max_smem) return 0;
  if (smem_out) *smem_out = pl.total;
  return 1;
}

int learn_small(const srlx_engine* eng, uint32_t n_updates, uintptr_t cuda_stream) {
  const SPlan pl = make_splan(*eng);
  SRLX_CHECK_CUDA(cudaFuncSetAttribute(learner_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.total));
  learner_small_kernel<<<1, kSmThreads, pl.total, (cudaStream_t)cuda_stream>>>(*eng, n_updates);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace srlx
