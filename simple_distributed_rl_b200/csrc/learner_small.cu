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
//   5. backward    on the s rows: net.cuh's backward with register-tiled dW / dX loops (small_backward_tile)
//   6. Adam        torch _single_tensor_adam arithmetic (adam_apply), hard target sync when train_count % interval == 0
//
// Reference path: srl/algorithms/dqn/model_torch.py:90-132, rainbow/model_torch.py:85-122, rainbow/rainbow.py:185-287.
// CPU twin: oracle/engine.py::learn.  Applies to: uniform replay, no NoisyNet, any depth / dueling head that fits.
#include "cluster.cuh"
#include "net.cuh"

namespace srlx {

constexpr int kSmThreads = 512;

// clock64() of thread 0 at the phase boundaries of the second-to-last update of a launch (tools/phase_clocks.py)
#define SRLX_SMSTAMP(slot)                                                                          \
  do {                                                                                              \
    if (eng.dbg_clock && tid == 0 && upd + 2 == n_updates) eng.dbg_clock[slot] = clock64();         \
  } while (0)

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


// ---- register-tiled pieces of the backward pass (same per-output summation order as net.cuh::net_backward_tile: rows /
//      units ascending, one fmaf per term -> bit-identical gradients, ~6x fewer shared-memory loads per FMA) -------------
// dW[u][k] += sum_r dY[r][u] * X[r][k],  db[u] += sum_r dY[r][u]        thread tile 4 units x 4 inputs
__device__ __forceinline__ void small_dw(const float* __restrict__ dY, int ldy, const float* __restrict__ X, int ldx, int R, int U,
                                         int K, float* __restrict__ Gw, float* __restrict__ Gb) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n_kt = (K + 3) >> 2, n_ut = (U + 3) >> 2;
  for (int item = tid; item < n_ut * n_kt; item += nt) {
    const int ut = item / n_kt, kt = item - ut * n_kt;
    const int u0 = ut * 4, k0 = kt * 4;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 4
    for (int r = 0; r < R; ++r) {
      const float4 dy = *reinterpret_cast<const float4*>(dY + r * ldy + u0);
      const float4 x = *reinterpret_cast<const float4*>(X + r * ldx + k0);
      const float dv[4] = {dy.x, dy.y, dy.z, dy.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(dv[a], xv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (u0 + a < U && k0 + b < K) Gw[(u0 + a) * K + k0 + b] += acc[a][b];
  }
  for (int u = tid; u < U; u += nt) {
    float acc = 0.f;
    for (int r = 0; r < R; ++r) acc += dY[r * ldy + u];
    Gb[u] += acc;
  }
}

// X[r][k] <- (X[r][k] > 0) ? sum_u dY[r][u] * W[u][k] : 0                thread tile 2 rows x 4 inputs, 4 units per step
__device__ __forceinline__ void small_dx(const float* __restrict__ dY, int ldy, const float* __restrict__ W, int ldw,
                                         float* __restrict__ X, int ldx, int R, int U, int K) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n_kt = (K + 3) >> 2, n_rt = (R + 1) >> 1;
  const int U4 = U & ~3;
  for (int item = tid; item < n_rt * n_kt; item += nt) {
    const int rt = item / n_kt, kt = item - rt * n_kt;
    const int r0 = rt * 2, r1 = min(r0 + 1, R - 1), k0 = kt * 4;
    float acc[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const float* dy0 = dY + r0 * ldy;
    const float* dy1 = dY + r1 * ldy;
    for (int u = 0; u < U4; u += 4) {
      const float4 d0 = *reinterpret_cast<const float4*>(dy0 + u);
      const float4 d1 = *reinterpret_cast<const float4*>(dy1 + u);
      const float d0v[4] = {d0.x, d0.y, d0.z, d0.w}, d1v[4] = {d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 w = *reinterpret_cast<const float4*>(W + (u + c) * ldw + k0);
        const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          acc[0][b] = fmaf(d0v[c], wv[b], acc[0][b]);
          acc[1][b] = fmaf(d1v[c], wv[b], acc[1][b]);
        }
      }
    }
    for (int u = U4; u < U; ++u) {
      const float4 w = *reinterpret_cast<const float4*>(W + u * ldw + k0);
      const float wv[4] = {w.x, w.y, w.z, w.w};
      const float a0 = dy0[u], a1 = dy1[u];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        acc[0][b] = fmaf(a0, wv[b], acc[0][b]);
        acc[1][b] = fmaf(a1, wv[b], acc[1][b]);
      }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (k0 + b < K) {
        float* x0 = X + r0 * ldx + k0 + b;
        *x0 = (*x0 > 0.f) ? acc[0][b] : 0.f;
        if (r0 + 1 < R) {
          float* x1 = X + (r0 + 1) * ldx + k0 + b;
          *x1 = (*x1 > 0.f) ? acc[1][b] : 0.f;
        }
      }
    }
  }
}

// net.cuh::net_backward_tile with the two hidden-layer loops above; the output layer / dueling head (<= 17 rows) is unchanged
__device__ inline void small_backward_tile(const srlx_net& net, const NetPlan& pl, const float* weff, float* acts, int R,
                                           const float* dQ, int lddq, float* G) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int L = net.n_layers, A = net.n_actions;
  const int nout = net.out_dim[L - 1], Ko = net.k_dim[L - 1];
  float* raw = acts + pl.x_s[L];
  const int ldr = pl.ldx[L];
  for (int r = tid; r < R; r += nt) {  // dueling combine backward (dueling_network.py:51-58) -> d raw
    if (net.dueling == SRLX_DUEL_NONE) {
      for (int a = 0; a < A; ++a) raw[r * ldr + a] = dQ[r * lddq + a];
    } else {
      float sum = 0.f;
      for (int a = 0; a < A; ++a) sum += dQ[r * lddq + a];
      int amax = 0;
      if (net.dueling == SRLX_DUEL_MAX) {
        float best = raw[r * ldr + 1];
        for (int a = 1; a < A; ++a)
          if (raw[r * ldr + 1 + a] > best) { best = raw[r * ldr + 1 + a]; amax = a; }
      }
      for (int a = 0; a < A; ++a) {
        float d = dQ[r * lddq + a];
        if (net.dueling == SRLX_DUEL_AVERAGE) d -= sum / (float)A;
        else if (net.dueling == SRLX_DUEL_MAX && a == amax) d -= sum;
        raw[r * ldr + 1 + a] = d;
      }
      raw[r * ldr + 0] = sum;
    }
  }
  __syncthreads();
  {  // output layer: dW, db
    const float* X = acts + pl.x_s[L - 1];
    const int ldx = pl.ldx[L - 1];
    for (int w = tid; w < nout * Ko; w += nt) {
      const int o = w / Ko, k = w - o * Ko;
      const int koff = (net.dueling != SRLX_DUEL_NONE && o > 0) ? Ko : 0;
      float acc = 0.f;
      for (int r = 0; r < R; ++r) acc = fmaf(raw[r * ldr + o], X[r * ldx + koff + k], acc);
      G[net.w_off[L - 1] + w] += acc;
    }
    for (int o = tid; o < nout; o += nt) {
      float acc = 0.f;
      for (int r = 0; r < R; ++r) acc += raw[r * ldr + o];
      G[net.b_off[L - 1] + o] += acc;
    }
  }
  __syncthreads();
  if (L > 1) {  // d hidden (input of the output layer), ReLU-masked, in place
    float* X = acts + pl.x_s[L - 1];
    const int ldx = pl.ldx[L - 1], width = pl.xw[L - 1];
    const float* W = weff + pl.w_s[L - 1];
    const int ldw = pl.ldw[L - 1];
    for (int w = tid; w < R * width; w += nt) {
      const int r = w / width, kk = w - r * width;
      float d = 0.f;
      if (net.dueling == SRLX_DUEL_NONE) {
        for (int o = 0; o < nout; ++o) d = fmaf(raw[r * ldr + o], W[o * ldw + kk], d);
      } else if (kk < Ko) {
        d = raw[r * ldr + 0] * W[kk];
      } else {
        for (int o = 1; o < nout; ++o) d = fmaf(raw[r * ldr + o], W[o * ldw + (kk - Ko)], d);
      }
      X[r * ldx + kk] = (X[r * ldx + kk] > 0.f) ? d : 0.f;
    }
  }
  __syncthreads();
  for (int l = L - 2; l >= 0; --l) {  // hidden layers, top down
    const float* dY = acts + pl.x_s[l + 1];
    float* X = acts + pl.x_s[l];
    small_dw(dY, pl.ldx[l + 1], X, pl.ldx[l], R, net.out_dim[l], net.k_dim[l], G + net.w_off[l], G + net.b_off[l]);
    __syncthreads();
    if (l > 0) {
      small_dx(dY, pl.ldx[l + 1], weff + pl.w_s[l], pl.ldw[l], X, pl.ldx[l], R, net.out_dim[l], net.k_dim[l]);
      __syncthreads();
    }
  }
}

// Output layer of a row tile, thread per (row, output): float4 dot products instead of a warp per row
__device__ inline void small_out_layer(const srlx_net& net, const float* __restrict__ X, int ldx, int R, const float* __restrict__ W,
                                       int ldw, const float* __restrict__ b, float* __restrict__ raw, int ldr) {
  const int L = net.n_layers, nout = net.out_dim[L - 1], K = net.k_dim[L - 1];
  const int K4 = round_up(K, 4);
  for (int w = threadIdx.x; w < R * nout; w += blockDim.x) {
    const int o = w / R, r = w - o * R;  // rows fastest: a warp reads one weight row (broadcast) and 32 activation rows
    const int koff = (net.dueling != SRLX_DUEL_NONE && o > 0) ? K : 0;
    const float* x = X + r * ldx + koff;
    const float* wr = W + o * ldw;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int k = 0; k < K4; k += 4) {
      const float4 xv = *reinterpret_cast<const float4*>(x + k);
      const float4 wv = *reinterpret_cast<const float4*>(wr + k);
      a0 = fmaf(xv.x, wv.x, a0);
      a1 = fmaf(xv.y, wv.y, a1);
      a2 = fmaf(xv.z, wv.z, a2);
      a3 = fmaf(xv.w, wv.w, a3);
    }
    raw[r * ldr + o] = ((a0 + a1) + (a2 + a3)) + b[o];
  }
}

// raw -> Q (plain | V + A - mean(A) | V + A - max(A) | V + A), thread per row
__device__ inline void small_dueling(const srlx_net& net, const float* __restrict__ raw, int ldr, int R, float* __restrict__ Q, int ldq) {
  const int A = net.n_actions;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const float* rr = raw + r * ldr;
    if (net.dueling == SRLX_DUEL_NONE) {
      for (int a = 0; a < A; ++a) Q[r * ldq + a] = rr[a];
    } else {
      const float v = rr[0];
      float red = 0.f;
      if (net.dueling == SRLX_DUEL_AVERAGE) {
        for (int a = 0; a < A; ++a) red += rr[1 + a];
        red = red / (float)A;
      } else if (net.dueling == SRLX_DUEL_MAX) {
        red = rr[1];
        for (int a = 1; a < A; ++a) red = fmaxf(red, rr[1 + a]);
      }
      for (int a = 0; a < A; ++a) Q[r * ldq + a] = v + rr[1 + a] - red;
    }
  }
}

__global__ void __launch_bounds__(kSmThreads, 1)
learner_small_kernel(const __grid_constant__ srlx_engine eng, const uint32_t n_updates) {
  extern __shared__ __align__(16) unsigned char smem[];
  const srlx_net& net = eng.net;
  __shared__ SPlan s_plan;  // the plan lives in shared memory: its per-layer tables are indexed at run time
  if (threadIdx.x == 0) s_plan = make_splan(eng);
  __syncthreads();
  const SPlan& pl = s_plan;
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
  float* disc_pow = red + 32;  // multi_discounts (rainbow.py:175): float32 of discount ** k
  if (tid < M) disc_pow[tid] = (float)pow(eng.discount, (double)tid);
  __syncthreads();

  for (uint32_t upd = 0; upd < n_updates; ++upd) {
    const uint64_t tc = tc0 + upd;
    // ---------------------------------------------------------------- 1. sample
    SRLX_SMSTAMP(0);
    if (warp == 0) {
      for (int i = lane; i < B; i += 32) {
        const uint4 w = philox(eng.seed, STREAM_UNIFORM_SAMPLE, (uint32_t)i, (uint32_t)tc, (uint32_t)(tc >> 32));
        pick[i] = (int)u_below(w.x, n_valid);
      }
      __syncwarp();
      bool dup = false;  // does any pick repeat an earlier one?  (rare: B^2 / (2 n_valid))
      for (int i = lane; i < B; i += 32) {
        const int pi = pick[i];
        for (int j = 0; j < i; ++j) dup |= (pick[j] == pi);
      }
      if (__any_sync(0xffffffffu, dup)) {
        if (lane == 0) {
          for (int i = 1; i < B; ++i) {  // a repeated pick is redrawn (attempt k = 1, 2, ...), in sample order
            int k = 0;
            while (true) {
              bool d2 = false;
              for (int j = 0; j < i; ++j) d2 |= (pick[j] == pick[i]);
              if (!d2 || ++k >= 65536) break;
              const uint4 w = philox(eng.seed, STREAM_UNIFORM_SAMPLE, (uint32_t)i | ((uint32_t)k << 16), (uint32_t)tc, (uint32_t)(tc >> 32));
              pick[i] = (int)u_below(w.x, n_valid);
            }
          }
        }
        __syncwarp();
      }
      for (int i = lane; i < B; i += 32) {
        const uint64_t pk = (uint64_t)pick[i];
        const uint64_t g = g_lo + pk / E;
        slot[i] = (int)((g % R) * E + pk % E);
      }
    } else if (tid == 32) {
      // Adam bias corrections of this update (torch/optim/adam.py: 1 - beta ** step), off the critical path
      const double t = (double)(adam0 + upd + 1);
      sc->step_size = (float)(eng.lr / (1.0 - pow(eng.adam_beta1, t)));
      sc->bc2_sqrt = (float)sqrt(1.0 - pow(eng.adam_beta2, t));
    }
    __syncthreads();
    SRLX_SMSTAMP(1);
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
    SRLX_SMSTAMP(2);
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
    SRLX_SMSTAMP(3);
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
    SRLX_SMSTAMP(4);
    for (int tile = 0; tile < pl.n_tiles; ++tile) {
      const float* wset = tile < pl.n_on_tiles ? weff : wefft;
      float* a = acts + (size_t)tile * np.act_floats;
      small_out_layer(net, a + np.x_s[L - 1], np.ldx[L - 1], tile_rows(tile), wset + np.w_s[L - 1], np.ldw[L - 1], wset + np.b_s[L - 1],
                      a + np.x_s[L], np.ldx[L]);
    }
    __syncthreads();
    for (int tile = 0; tile < pl.n_tiles; ++tile) {
      float* a = acts + (size_t)tile * np.act_floats;
      // Q rows: online set first ([0, n_on_rows)), target rows at B + BM
      float* q = tile < pl.n_on_tiles ? Q + (size_t)tile * kRowTile * A : Q + (size_t)(B + BM + (tile - pl.n_on_tiles) * kRowTile) * A;
      small_dueling(net, a + np.x_s[L], np.ldx[L], tile_rows(tile), q, A);
    }
    __syncthreads();
    SRLX_SMSTAMP(5);
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
          target += (td * disc_pow[k]) * retrace;
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
      if ((tc % (uint64_t)eng.target_update_interval) == 0) sc->sync_count += 1;
    }
    if (eng.dbg_target_q)
      for (int i = tid; i < B; i += kSmThreads) eng.dbg_target_q[i] = tq[i];
    if (eng.dbg_q_sa)
      for (int i = tid; i < B; i += kSmThreads) eng.dbg_q_sa[i] = qsa[i];
    SRLX_SMSTAMP(6);
    // ---------------------------------------------------------------- 5. backward on the s rows (G was zeroed by Adam)
    for (int tile = 0; tile < pl.n_s_tiles; ++tile) {
      const int Rt = min(kRowTile, B - tile * kRowTile);
      small_backward_tile(net, np, weff, acts + (size_t)tile * np.act_floats, Rt, dQ + (size_t)tile * kRowTile * A, A, G);
    }
    __syncthreads();
    SRLX_SMSTAMP(7);
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
    SRLX_SMSTAMP(8);
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
