// learner_small.cu -- Trainer.train() for small plain MLPs on uniform replay: the batch rows are split over a small
// thread-block cluster, every CTA holds the whole network in shared memory.
//
// BASELINE configs[1] (DQN, MLP[64,64], uniform replay 1M) has ~4.6k parameters.  The unit-sharded cluster kernels
// (learner.cu, learner_fast.cu) pay a DSMEM exchange per layer boundary, which for a net this small is all latency and no
// work (37 us per update on the generic kernel); one SM alone is bound by its shared-memory bandwidth (25 us).  Here the
// cluster is data parallel over the BATCH: CTA c owns B/C of the sampled items end to end -- gather, the three forwards,
// targets, backward -- with a full copy of the online and target weights in its shared memory, and owns 1/C of the
// parameters for the optimiser step.  An update is
//
//   1. sample      B distinct uniform picks (replay_buffer.py:34-36), every CTA redundantly: attempt 0 of all draws in
//                  parallel, duplicates resolved in sample order (== the sequential rejection loop of oracle/sumtree.py);
//                  done one update ahead by an idle warp while warp 0 evaluates the targets
//   2. gather      the (M+1)-state windows of the CTA's items from the ring, padded tails rebuilt (rainbow.py:358-371)
//   3. forward     online(s), online(s'), target(s') on the CTA's rows (register-tiled warp tasks, weights in smem)
//   4. targets     double-DQN / n-step Retrace target, Huber gradient (thread per item; same code as learner.cu)
//   5. backward    on the CTA's rows: a delta chain (one barrier per layer), then every layer's dW / db in one sweep
//                  -> a partial gradient of every parameter
//   6. REDUCE-SCATTER the partial gradients to the CTA that owns the parameter slice (one bulk DSMEM copy per peer, mbarrier
//      transaction count), summed there in CTA order; Adam on the slice (torch _single_tensor_adam arithmetic)
//   7. ALL-GATHER the updated slice into every CTA's weight copy; hard target sync when train_count % interval == 0
//
// two DSMEM hops per update instead of two per layer.  Reference path: srl/algorithms/dqn/model_torch.py:90-132,
// rainbow/model_torch.py:85-122, rainbow/rainbow.py:185-287.  CPU twin: oracle/engine.py::learn.
// Proportional replay adds a REPLAY CTA to the cluster: it receives the |TD| of the batch from the compute CTAs, applies
// ProportionalMemory.update in the reference's order, samples the next batch and sends (slot, IS weight) back, all in the shadow
// of the compute CTAs' backward / exchanges.
// Applies to: no NoisyNet, any depth / dueling head whose plan fits the shared memory of an SM.
#include "cluster.cuh"
#include "net.cuh"
#include "tree.cuh"

namespace srlx {

constexpr int kSmThreads = 512;
constexpr int kSmMaxCluster = 8;
constexpr int kSmTreeCache = 4095;  // replay CTA: the top 12 levels of the SumTree in shared memory

// clock64() of thread 0 of CTA 0 at the phase boundaries of the second-to-last update of a launch (tools/phase_clocks.py)
#define SRLX_SMSTAMP(slot)                                                                                     \
  do {                                                                                                         \
    if (eng.dbg_clock && rank == 0 && tid == 0 && upd + 2 == n_updates) eng.dbg_clock[slot] = clock64();       \
  } while (0)

struct SPlan {
  NetPlan np;
  int C, B, Bc, M, A, D, BcM, P, S;  // Bc = items per CTA, S = floats per parameter slice (multiple of 4), C * S >= P
  int n_on_rows, n_on_tiles, n_tg_tiles, n_tiles, n_s_tiles;
  size_t off_mbar, off_weff, off_wefft, off_m, off_v, off_g, off_recv, off_wflat, off_wnew, off_slot, off_acts, off_dacts, off_q, off_pick,
      off_win, off_tq, off_red, off_loss, off_scal, off_per, off_mem, total;
};

struct SScal {
  double loss_sum, last_loss;
  float step_size, bc2_sqrt;
  unsigned int sync_count;
};

__host__ __device__ inline SPlan make_splan(const srlx_engine& eng, int C) {
  SPlan p;
  p.np = make_plan(eng.net);
  p.C = C;
  p.B = eng.batch_size;
  p.Bc = (p.B + C - 1) / C;
  p.M = eng.multisteps;
  p.A = eng.n_actions;
  p.D = eng.obs_dim;
  p.BcM = p.Bc * p.M;
  p.P = eng.net.n_params;
  p.S = round_up((p.P + C - 1) / C, 4);
  const bool need_online_next = eng.enable_double_dqn || p.M > 1;
  p.n_on_rows = p.Bc + (need_online_next ? p.BcM : 0);  // online rows of a CTA: [0,Bc) = s, [Bc,Bc+BcM) = s'_k
  p.n_on_tiles = (p.n_on_rows + kRowTile - 1) / kRowTile;
  p.n_tg_tiles = (p.BcM + kRowTile - 1) / kRowTile;     // target rows: s'_k
  p.n_tiles = p.n_on_tiles + p.n_tg_tiles;
  p.n_s_tiles = (p.Bc + kRowTile - 1) / kRowTile;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) / 16 * 16; return r; };
  p.off_mbar = take(64);
  p.off_weff = take((size_t)p.np.weff_floats * 4);
  p.off_wefft = take((size_t)p.np.weff_floats * 4);
  p.off_m = take((size_t)p.S * 4);
  p.off_v = take((size_t)p.S * 4);
  p.off_g = take((size_t)C * p.S * 4);       // this CTA's partial gradient, flat parameter order (padded to C * S)
  p.off_recv = take((size_t)C * p.S * 4);    // partial gradients of the own slice, one row per source CTA
  p.off_wflat = take((size_t)C * p.S * 4);   // all-gathered new parameters, flat order
  p.off_wnew = take((size_t)p.S * 4);        // the own slice after Adam (source of the all-gather copies)
  p.off_slot = take((size_t)C * p.S * 2);
  p.off_acts = take((size_t)p.n_tiles * p.np.act_floats * 4);
  p.off_dacts = take((size_t)p.n_s_tiles * p.np.act_floats * 4);  // d loss / d activation of the s rows (same layout as acts)
  p.off_q = take((size_t)(p.Bc + 2 * p.BcM) * p.A * 4);  // Q(s) [Bc][A], online Q(s') [BcM][A], target Q(s') [BcM][A]
  p.off_pick = take((size_t)p.B * 4 * 3);                // picks, slots of update t and t+1 (all B items, every CTA)
  p.off_win = take((size_t)p.BcM * 4 * 4);               // action, reward, term, done of every local window step
  p.off_tq = take((size_t)p.Bc * 4 * 2);                 // target, q(s,a)
  p.off_red = take(64 * 4);
  p.off_loss = take((size_t)kSmMaxCluster * 8);
  p.off_scal = take(64);
  p.off_per = take((size_t)2 * p.B * 8);  // proportional replay: (slot, IS weight) of every item, two updates deep
  p.total = o;
  // the replay CTA (proportional replay only) overlays its own layout on the same allocation: hash scratch, batch arrays, TDs
  p.off_mem = 128;
  const size_t mem_total = p.off_mem + sizeof(TreeHashScratch) + (size_t)p.B * (8 + 8 + 8 + 4 + 8) + 64 + (size_t)kSmTreeCache * 8;
  if (eng.mem_kind == SRLX_MEM_PROPORTIONAL && mem_total > p.total) p.total = mem_total;
  return p;
}

__host__ inline bool small_shape_ok(const srlx_engine& eng) {
  const bool per_ok = eng.mem_kind == SRLX_MEM_UNIFORM ||
                      (eng.tree != nullptr && (long long)eng.ring_rows * eng.n_envs < (1ll << 30) && eng.batch_size <= 1024);
  return per_ok && !eng.net.noisy && eng.net.n_layers >= 2 && eng.net.n_params < 60000 && eng.obs_dim <= SRLX_MAX_OBS;
}

// ---- PTX: DSMEM stores that complete a transaction count on the destination CTA's mbarrier -----------------------------
__device__ __forceinline__ uint32_t sm_mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void sm_st_async_f4(uint32_t raddr, float4 v, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(rmbar)
               : "memory");
}
__device__ __forceinline__ void sm_st_async_b32(uint32_t raddr, uint32_t v, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(v), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void sm_st_async_f2(uint32_t raddr, float a, float b, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(raddr), "f"(a),
               "f"(b), "r"(rmbar)
               : "memory");
}
// one bulk copy of `bytes` (multiple of 16) from this CTA's shared memory into a peer's, completing on the peer's mbarrier
__device__ __forceinline__ void sm_bulk_s2c(uint32_t dst_cluster_addr, const void* src_smem, uint32_t bytes, uint32_t rmbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster_addr),
               "r"(smem_u32(src_smem)), "r"(bytes), "r"(rmbar)
               : "memory");
}
__device__ __forceinline__ void sm_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sm_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// one warp tile of a hidden layer with ONE unit per lane (4 rows x 32 units): half the dependent chain of net.cuh's
// 64-unit tile, for the few rows a CTA owns
__device__ __forceinline__ void dense_relu_task32(const float* __restrict__ X, int ldx, int R, int K, const float* __restrict__ W,
                                                  int ldw, const float* __restrict__ b, int U, float* __restrict__ Y, int ldy,
                                                  int rt, int ut) {
  const int lane = threadIdx.x & 31;
  const int K4 = round_up(K, 4);
  const int r0 = rt * 4;
  const int u0 = ut * 32 + lane;
  const bool v0 = u0 < U;
  const float* w0p = W + (v0 ? u0 : 0) * ldw;
  const float b0 = v0 ? b[u0] : 0.f;
  float acc[4];
  const float* xr[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc[i] = b0;
    xr[i] = X + min(r0 + i, R - 1) * ldx;
  }
  for (int k = 0; k < K4; k += 4) {
    const float4 wa = *reinterpret_cast<const float4*>(w0p + k);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 x = *reinterpret_cast<const float4*>(xr[i] + k);
      acc[i] = fmaf(x.x, wa.x, acc[i]);
      acc[i] = fmaf(x.y, wa.y, acc[i]);
      acc[i] = fmaf(x.z, wa.z, acc[i]);
      acc[i] = fmaf(x.w, wa.w, acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (r0 + i < R && v0) Y[(r0 + i) * ldy + u0] = fmaxf(acc[i], 0.f);
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
                                         int K, float* __restrict__ Gw, float* __restrict__ Gb, bool accum, int rot) {
  // rot: the thread that takes item 0 -- the sweeps of different layers start on different warps, so they run side by side
  const int nt = blockDim.x;
  const int tid = ((int)threadIdx.x - rot + nt) % nt;
  const int n_kt = (K + 3) >> 2, n_ut = (U + 3) >> 2;
  for (int item = tid; item < n_ut * n_kt; item += nt) {
    const int ut = item / n_kt, kt = item - ut * n_kt;
    const int u0 = ut * 4, k0 = kt * 4;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 1
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
        if (u0 + a < U && k0 + b < K) {
          float* gp = Gw + (u0 + a) * K + k0 + b;
          *gp = accum ? *gp + acc[a][b] : acc[a][b];  // first row tile of an update: plain store (no zeroing pass needed)
        }
  }
  for (int u = tid; u < U; u += nt) {
    float acc = 0.f;
    for (int r = 0; r < R; ++r) acc += dY[r * ldy + u];
    Gb[u] = accum ? Gb[u] + acc : acc;
  }
}

// dX[r][k] = (X[r][k] > 0) ? sum_u dY[r][u] * W[u][k] : 0                thread tile 2 rows x 4 inputs, 4 units per step
__device__ __forceinline__ void small_dx(const float* __restrict__ dY, int ldy, const float* __restrict__ W, int ldw,
                                         const float* __restrict__ X, float* __restrict__ dX, int ldx, int R, int U, int K) {
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
        const int o0 = r0 * ldx + k0 + b;
        dX[o0] = (X[o0] > 0.f) ? acc[0][b] : 0.f;
        if (r0 + 1 < R) {
          const int o1 = (r0 + 1) * ldx + k0 + b;
          dX[o1] = (X[o1] > 0.f) ? acc[1][b] : 0.f;
        }
      }
    }
  }
}

// Backward of a row tile in two stages (same arithmetic as net.cuh::net_backward_tile, fewer block barriers):
//   small_delta_chain: d loss / d (pre-activation) of every hidden layer, top down, into `dacts` (layout of `acts`); the
//                      output layer's d raw was written there by the target step
//   small_all_dw:      every layer's dW / db from (dacts, acts) in one barrier-free sweep
__device__ inline void small_delta_chain(const srlx_net& net, const NetPlan& pl, const float* weff, const float* acts, float* dacts, int R,
                                         long long* clk) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int L = net.n_layers;
  const int nout = net.out_dim[L - 1], Ko = net.k_dim[L - 1];
  const float* raw = dacts + pl.x_s[L];
  const int ldr = pl.ldx[L];
  {  // d hidden (input of the output layer), ReLU-masked
    const float* X = acts + pl.x_s[L - 1];
    float* dX = dacts + pl.x_s[L - 1];
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
      dX[r * ldx + kk] = (X[r * ldx + kk] > 0.f) ? d : 0.f;
    }
  }
  __syncthreads();
  if (clk) clk[10] = clock64();
#pragma unroll 1
  for (int l = L - 2; l >= 1; --l) {  // delta of layer l-1's output from delta of layer l's output
    small_dx(dacts + pl.x_s[l + 1], pl.ldx[l + 1], weff + pl.w_s[l], pl.ldw[l], acts + pl.x_s[l], dacts + pl.x_s[l], pl.ldx[l], R,
             net.out_dim[l], net.k_dim[l]);
    __syncthreads();
  }
  if (clk) clk[11] = clock64();
}

__device__ inline void small_all_dw(const srlx_net& net, const NetPlan& pl, const float* acts, const float* dacts, int R, float* G,
                                    bool accum, long long* clk) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int L = net.n_layers;
  const int nout = net.out_dim[L - 1], Ko = net.k_dim[L - 1];
  const float* raw = dacts + pl.x_s[L];
  const int ldr = pl.ldx[L];
  {  // output layer
    const float* X = acts + pl.x_s[L - 1];
    const int ldx = pl.ldx[L - 1];
    for (int w = tid; w < nout * Ko; w += nt) {
      const int o = w / Ko, k = w - o * Ko;
      const int koff = (net.dueling != SRLX_DUEL_NONE && o > 0) ? Ko : 0;
      float acc = 0.f;
      for (int r = 0; r < R; ++r) acc = fmaf(raw[r * ldr + o], X[r * ldx + koff + k], acc);
      float* gp = G + net.w_off[L - 1] + w;
      *gp = accum ? *gp + acc : acc;
    }
    for (int o = tid; o < nout; o += nt) {
      float acc = 0.f;
      for (int r = 0; r < R; ++r) acc += raw[r * ldr + o];
      float* gp = G + net.b_off[L - 1] + o;
      *gp = accum ? *gp + acc : acc;
    }
  }
  if (clk) clk[12] = clock64();
#pragma unroll 1
  for (int l = L - 2; l >= 0; --l)
    small_dw(dacts + pl.x_s[l + 1], pl.ldx[l + 1], acts + pl.x_s[l], pl.ldx[l], R, net.out_dim[l], net.k_dim[l], G + net.w_off[l],
             G + net.b_off[l], accum, ((nt >> 1) + (L - 2 - l) * 96) % nt);
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

// Narrow heads (nout * K <= 512, e.g. DQN's 64 -> 2): one thread per row computes all raw outputs and the row's Q values,
// saving the barrier between small_out_layer and small_dueling
__device__ inline void small_out_q_row(const srlx_net& net, const float* __restrict__ xrow, const float* __restrict__ W, int ldw,
                                       const float* __restrict__ b, float* __restrict__ rr, float* __restrict__ q) {
  const int L = net.n_layers, nout = net.out_dim[L - 1], K = net.k_dim[L - 1], A = net.n_actions;
  const int K4 = round_up(K, 4);
#pragma unroll 1
  for (int o = 0; o < nout; ++o) {
    const int koff = (net.dueling != SRLX_DUEL_NONE && o > 0) ? K : 0;
    const float* x = xrow + koff;
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
    rr[o] = ((a0 + a1) + (a2 + a3)) + b[o];
  }
  if (net.dueling == SRLX_DUEL_NONE) {
    for (int a = 0; a < A; ++a) q[a] = rr[a];
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
    for (int a = 0; a < A; ++a) q[a] = v + rr[1 + a] - red;
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

// B distinct uniform picks of update `tc` -> ring slots (one warp).  Attempt 0 of every draw in parallel; a pick that repeats
// an earlier one is redrawn (attempt k = 1, 2, ...) in sample order, which is what the sequential rejection loop produces.
__device__ __noinline__ void small_sample(const srlx_engine& eng, uint64_t tc, int B, uint32_t n_valid, uint32_t g_lo_mod, int* pick,
                                          int* slot) {
  const int lane = threadIdx.x & 31;
  const uint32_t E = (uint32_t)eng.n_envs, R = (uint32_t)eng.ring_rows;
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
      for (int i = 1; i < B; ++i) {
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
    const uint32_t pk = (uint32_t)pick[i];
    const uint32_t q = pk / E;                       // row offset inside the valid range (g = g_lo + q)
    slot[i] = (int)(((g_lo_mod + q) % R) * E + (pk - q * E));
  }
}

// Adam bias corrections of optimiser step t (torch/optim/adam.py: 1 - beta ** step)
__device__ __noinline__ void small_adam_scalars(const srlx_engine& eng, double t, SScal* sc) {
  sc->step_size = (float)(eng.lr / (1.0 - pow(eng.adam_beta1, t)));
  sc->bc2_sqrt = (float)sqrt(1.0 - pow(eng.adam_beta2, t));
}

__global__ void __launch_bounds__(kSmThreads, 1)
learner_small_kernel(const __grid_constant__ srlx_engine eng, const uint32_t n_updates) {
  extern __shared__ __align__(16) unsigned char smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const bool per = eng.mem_kind == SRLX_MEM_PROPORTIONAL;
  // proportional replay: the last CTA of the cluster is the replay CTA (SumTree update + next sample), the others compute
  const int C = (int)cluster.num_blocks() - (per ? 1 : 0);
  const int rank = (int)cluster.block_rank();
  const bool is_replay = per && rank == C;
  const srlx_net& net = eng.net;
  __shared__ SPlan s_plan;  // the plan lives in shared memory: its per-layer tables are indexed at run time
  if (threadIdx.x == 0) s_plan = make_splan(eng, C);
  __syncthreads();
  const SPlan& pl = s_plan;
  const NetPlan& np = pl.np;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + pl.off_mbar);  // [0] reduce-scatter, [1] all-gather
  float* weff = reinterpret_cast<float*>(smem + pl.off_weff);
  float* wefft = reinterpret_cast<float*>(smem + pl.off_wefft);
  float* am = reinterpret_cast<float*>(smem + pl.off_m);
  float* av = reinterpret_cast<float*>(smem + pl.off_v);
  float* G = reinterpret_cast<float*>(smem + pl.off_g);
  float* recv = reinterpret_cast<float*>(smem + pl.off_recv);
  float* wflat = reinterpret_cast<float*>(smem + pl.off_wflat);
  unsigned short* pslot = reinterpret_cast<unsigned short*>(smem + pl.off_slot);
  float* acts = reinterpret_cast<float*>(smem + pl.off_acts);
  float* dacts = reinterpret_cast<float*>(smem + pl.off_dacts);
  float* Q = reinterpret_cast<float*>(smem + pl.off_q);
  int* pick = reinterpret_cast<int*>(smem + pl.off_pick);
  int* slot = pick + pl.B;
  int* w_act = reinterpret_cast<int*>(smem + pl.off_win);
  float* w_rew = reinterpret_cast<float*>(smem + pl.off_win) + pl.BcM;
  float* w_term = w_rew + pl.BcM;
  int* w_done = reinterpret_cast<int*>(w_term + pl.BcM);
  float* tq = reinterpret_cast<float*>(smem + pl.off_tq);
  float* qsa = tq + pl.Bc;
  float* red = reinterpret_cast<float*>(smem + pl.off_red);
  float* lossbuf = reinterpret_cast<float*>(smem + pl.off_loss);  // [C][2] partial Huber sums (CTA 0)
  SScal* sc = reinterpret_cast<SScal*>(smem + pl.off_scal);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kSmThreads >> 5;
  const int B = pl.B, Bc = pl.Bc, M = pl.M, A = pl.A, D = pl.D, BcM = pl.BcM, E = eng.n_envs, R = eng.ring_rows, P = pl.P, S = pl.S;
  const int L = net.n_layers;
  const int i0 = rank * Bc;                           // first batch item of this CTA
  const int nb = max(0, min(Bc, B - i0));             // its item count
  const int p0 = rank * S;                            // first parameter of its slice
  const bool need_online_next = eng.enable_double_dqn || M > 1;
  srlx_state* st = eng.state;
  const uint64_t tc0 = st->train_count, mem_size = st->mem_size, vec_steps = st->vec_steps, adam0 = st->adam_step;
  if (!(mem_size >= eng.warmup_size && mem_size >= (uint64_t)B)) return;  // warming up (uniform across the cluster)

  // ---- one-time setup: the whole network into every CTA, the moments of the own slice ------------------------------------
  for (int i = tid; i < np.weff_floats; i += kSmThreads) { weff[i] = 0.f; wefft[i] = 0.f; }
  for (int i = tid; i < pl.n_tiles * np.act_floats; i += kSmThreads) acts[i] = 0.f;
  for (int i = tid; i < C * S; i += kSmThreads) { G[i] = 0.f; wflat[i] = 0.f; pslot[i] = 0; }
  for (int i = tid; i < pl.n_s_tiles * np.act_floats; i += kSmThreads) dacts[i] = 0.f;
  for (int i = tid; i < S; i += kSmThreads) { am[i] = 0.f; av[i] = 0.f; }
  if (tid == 0) {
    sc->loss_sum = 0.0; sc->last_loss = 0.0; sc->sync_count = 0;
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_init(&mbar[2], 1);  // compute CTAs: (slot, IS weight) of the next batch from the replay CTA
    mbar_init(&mbar[3], 1);  // replay CTA: |TD| of the current batch from the compute CTAs
    mbar_init(&mbar[4], 1);  // compute CTAs: IS weights of the next batch (sent after the slots: only the target step needs them)
    fence_mbar_init();
    if (!is_replay) {
      sm_expect_tx(&mbar[0], (uint32_t)((C - 1) * S * 4 + (rank == 0 ? C * 8 : 0)));
      sm_expect_tx(&mbar[1], (uint32_t)((C - 1) * S * 4));
      if (per) {
        sm_expect_tx(&mbar[2], (uint32_t)B * 4);
        sm_expect_tx(&mbar[4], (uint32_t)B * 4);
      }
    } else {
      sm_expect_tx(&mbar[3], (uint32_t)B * 8);
    }
  }
  __syncthreads();
  for (int p = tid; p < P; p += kSmThreads) {
    const int s = weff_slot(net, np, p, layer_of_param(net, p));
    pslot[p] = (unsigned short)s;
    weff[s] = __ldcg(eng.params + p);
    wefft[s] = __ldcg(eng.target + p);
    if (p >= p0 && p < p0 + S) {
      am[p - p0] = __ldcg(eng.adam_m + p);
      av[p - p0] = __ldcg(eng.adam_v + p);
    }
  }
  cluster.sync();  // mbarriers initialised and every CTA's shared memory ready before any remote store
  int* per_slot = reinterpret_cast<int*>(smem + pl.off_per);          // [2][B] ring slots of the batches of update t, t+1
  float* per_w = reinterpret_cast<float*>(smem + pl.off_per) + 2 * B;  // [2][B] their IS weights

  if (is_replay) {
    // ============================ REPLAY CTA (proportional replay): SumTree update(t), then sample(t+1) ============================
    // ProportionalMemory.update / .sample (proportional_memory.py:131-177) with the block-wide routines of tree.cuh: the update
    // applies the items in the reference's order (bit-identical node sums), the sampler makes the reference's decisions on the
    // same node values.  Both run while the compute CTAs are in backward / exchange / Adam of update t.
    unsigned char* mb = smem + pl.off_mem;
    TreeHashScratch* hs = reinterpret_cast<TreeHashScratch*>(mb);
    int64_t* s_idx = reinterpret_cast<int64_t*>(mb + sizeof(TreeHashScratch));
    double* s_pri = reinterpret_cast<double*>(s_idx + B);
    double* s_tmp = s_pri + B;
    float* tdbuf = reinterpret_cast<float*>(s_tmp + B);   // [B][2] |td| of the current batch (from the compute CTAs)
    float* s_w = tdbuf + 2 * B;
    double* cache = reinterpret_cast<double*>(mb + sizeof(TreeHashScratch) + (((size_t)B * 36 + 15) / 16) * 16);
    __shared__ unsigned long long s_retries;
    __shared__ double s_maxp;
    const int64_t cap = (int64_t)R * E, n_nodes = 2 * cap - 1;
    // whole top levels that fit the cache and the tree's inner nodes
    int n_cache = 1;
    while (2 * n_cache + 1 <= kSmTreeCache && 2 * n_cache + 1 <= (int)min((int64_t)kSmTreeCache, cap - 1)) n_cache = 2 * n_cache + 1;
    for (int i = tid; i < n_cache; i += kSmThreads) cache[i] = __ldcg(eng.tree + i);
    if (tid == 0) { s_retries = 0; s_maxp = st->max_priority; }
    __syncthreads();
    auto sample_and_send = [&](uint64_t tc, uint32_t parb) {
      const double total = cache[0];
      // PriorityReplayBuffer.step is the train_count of the PREVIOUS update() call (priority_replay_buffer.py:232,250)
      const double stepd = (tc > 0) ? (double)(tc - 1) : 0.0;
      double beta = eng.per_beta_initial + (1.0 - eng.per_beta_initial) * stepd / eng.per_beta_steps;
      beta = beta > 1.0 ? 1.0 : beta;
      per_sample_block(eng.tree, n_nodes, total, B, eng.seed, tc, nullptr, 9999, eng.has_duplicate, s_idx, s_pri, s_tmp, &s_retries, cache,
                       n_cache);
      if (eng.dbg_clock && tid == 0 && tc + 2 == tc0 + n_updates) eng.dbg_clock[23] = clock64();
      // the slots go out first: the compute CTAs start their gather and forwards while the IS weights are formed
      for (int w = tid; w < C * B; w += kSmThreads) {
        const int c = w / B, i = w - c * B;
        sm_st_async_b32(sm_mapa(smem_u32(per_slot + (size_t)parb * B + i), (uint32_t)c), (uint32_t)(int)(s_idx[i] - (cap - 1)),
                        sm_mapa(smem_u32(&mbar[2]), (uint32_t)c));
      }
      per_weights_block(total, (double)mem_size, beta, B, s_pri, s_tmp, s_w);
      if (eng.dbg_clock && tid == 0 && tc + 2 == tc0 + n_updates) eng.dbg_clock[24] = clock64();
      for (int w = tid; w < C * B; w += kSmThreads) {
        const int c = w / B, i = w - c * B;
        sm_st_async_b32(sm_mapa(smem_u32(per_w + (size_t)parb * B + i), (uint32_t)c), __float_as_uint(s_w[i]),
                        sm_mapa(smem_u32(&mbar[4]), (uint32_t)c));
      }
      for (int i = tid; i < B; i += kSmThreads) {
        if (eng.dbg_sample_idx) eng.dbg_sample_idx[i] = s_idx[i];
        if (eng.dbg_weights) eng.dbg_weights[i] = s_w[i];
      }
      __syncthreads();
      // the first half of this batch's priority update needs only its tree indices: claim the hash slots and fetch the nodes
      // now, while the compute CTAs gather / forward / form the targets (nothing else writes the tree meanwhile)
      if (B <= kTreeHashChunk) tree_update_prepare(eng.tree, s_idx, B, hs);
    };
    sample_and_send(tc0, 0);
    for (uint32_t upd = 0; upd < n_updates; ++upd) {
      const uint32_t par = upd & 1;
      if (warp == 0) mbar_wait_sleep(&mbar[3], par);  // all |td| of update `upd` have arrived
      __syncthreads();
      if (eng.dbg_clock && tid == 0 && upd + 3 == n_updates) eng.dbg_clock[20] = clock64();
      if (tid == 0) sm_expect_tx(&mbar[3], (uint32_t)B * 8);
      for (int i = tid; i < B; i += kSmThreads) s_pri[i] = pow(fabs((double)tdbuf[2 * i]) + eng.per_epsilon, eng.per_alpha);
      __syncthreads();
      if (eng.dbg_clock && tid == 0 && upd + 3 == n_updates) eng.dbg_clock[21] = clock64();
      if (B <= kTreeHashChunk) tree_update_apply(eng.tree, s_pri, B, hs, cache, n_cache);
      else tree_update_batch(eng.tree, s_idx, s_pri, B, hs, cache, n_cache);
      if (eng.dbg_clock && tid == 0 && upd + 3 == n_updates) eng.dbg_clock[22] = clock64();
      if (tid == 0) {
        double mp = s_maxp;
        for (int i = 0; i < B; ++i) mp = (mp < s_pri[i]) ? s_pri[i] : mp;
        s_maxp = mp;
      }
      __syncthreads();
      if (upd + 1 < n_updates) sample_and_send(tc0 + upd + 1, par ^ 1);
    }
    if (tid == 0) {
      st->max_priority = s_maxp;
      st->sample_retries += s_retries;
    }
    cluster.sync();
    return;
  }

  // local row r of the online set / target set -> its place in the tile activation areas
  auto x_row = [&](int set_tile0, int r) -> float* {
    return acts + (size_t)(set_tile0 + r / kRowTile) * np.act_floats + np.x_s[0] + (r % kRowTile) * np.ldx[0];
  };
  const uint64_t g_next = vec_steps;
  const uint64_t g_lo = g_next > (uint64_t)R ? g_next - R : 0;
  const uint32_t n_valid = (uint32_t)((g_next - (uint64_t)(M - 1) - g_lo) * E);
  const uint32_t g_lo_mod = (uint32_t)(g_lo % (uint64_t)R), g_last_mod = (uint32_t)((vec_steps - 1) % (uint64_t)R);
  const float b1 = (float)eng.adam_beta1, b2 = (float)eng.adam_beta2, aeps = (float)eng.adam_eps;
  float* disc_pow = red + 32;  // multi_discounts (rainbow.py:175): float32 of discount ** k
  if (tid < M) disc_pow[tid] = (float)pow(eng.discount, (double)tid);
  if (warp == 1 && !per) small_sample(eng, tc0, B, n_valid, g_lo_mod, pick, slot);  // slots of the first update
  __syncthreads();
  const uint32_t mb_rs = smem_u32(&mbar[0]), mb_ag = smem_u32(&mbar[1]);
  // where the replay CTA keeps the |td| of the batch (its own layout, see above), as an address in this CTA's window for mapa
  float* tdbuf_remote = reinterpret_cast<float*>(smem + pl.off_mem + sizeof(TreeHashScratch) + (size_t)B * 24);
  const int nbM = nb * M;
  const bool one_warp = nbM <= 32;  // the CTA's window steps fit one warp: gather / padding / copy need no block barrier
  // uniform replay samples update t+1 during the target step of update t: the last warp then issues that batch's ring loads before the
  // backward pass and parks them in registers until the pass no longer needs the current rows
  const bool can_prefetch = !per && one_warp && D <= 4;
  float* wnew = reinterpret_cast<float*>(smem + pl.off_wnew);

  for (uint32_t upd = 0; upd < n_updates; ++upd) {
    const uint64_t tc = tc0 + upd;
    const uint32_t par = upd & 1;
    const int* slot_t = slot + par * B;  // sampled during the previous update's target phase
    if (per) {  // proportional replay: (slot, IS weight) of this batch come from the replay CTA
      if (warp == 0) mbar_wait_sleep(&mbar[2], par);
      __syncthreads();
      if (tid == 0) sm_expect_tx(&mbar[2], (uint32_t)B * 4);
      for (int i = tid; i < B; i += kSmThreads) slot[par * B + i] = per_slot[(size_t)par * B + i];
      __syncthreads();
    }
    SRLX_SMSTAMP(0);
    // ---------------------------------------------------------------- 2. gather the CTA's items (loads before stores)
    // (uniform replay knows its next batch an update ahead: from the second update on the rows were fetched across the
    // previous update's backward pass, see "prefetch" below)
    if ((!one_warp || warp == 0) && !(can_prefetch && upd > 0)) {
      for (int w = tid; w < nbM; w += kSmThreads) {
        const int il = w / M, k = w - il * M;
        const int s0 = slot_t[i0 + il];
        const int rho = s0 / E, e = s0 - rho * E;
        const int sk = ((rho + k) % R) * E + e;
        const int a = __ldcg(eng.ring_action + sk);
        const float rw = __ldcg(eng.ring_reward + sk);
        const unsigned char tm = __ldcg(eng.ring_term + sk), dn = __ldcg(eng.ring_done + sk);
        float* xt = x_row(pl.n_on_tiles, w);
        float* xs = x_row(0, il);
        if (D <= 4) {
          float xv[4], sv[4];
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            xv[d] = d < D ? __ldcg(eng.ring_next_obs + (size_t)sk * D + d) : 0.f;
            sv[d] = (d < D && k == 0) ? __ldcg(eng.ring_obs + (size_t)s0 * D + d) : 0.f;
          }
#pragma unroll
          for (int d = 0; d < 4; ++d)
            if (d < D) {
              xt[d] = xv[d];
              if (k == 0) xs[d] = sv[d];
            }
        } else {
#pragma unroll 1
          for (int d = 0; d < D; ++d) {
            xt[d] = __ldcg(eng.ring_next_obs + (size_t)sk * D + d);
            if (k == 0) xs[d] = __ldcg(eng.ring_obs + (size_t)s0 * D + d);
          }
        }
        w_act[w] = a;
        w_rew[w] = rw;
        w_term[w] = (float)tm;
        w_done[w] = (int)dn;
      }
    }
    if (one_warp) __syncwarp(); else __syncthreads();
    // padded tails (rainbow.py:358-371)
    if (M > 1 && (!one_warp || warp == 0)) {
      for (int il = tid; il < nb; il += kSmThreads) {
        const int s0 = slot_t[i0 + il];
        const int rho = s0 / E, e = s0 - rho * E;
        const uint64_t g_item = (vec_steps - 1) - (uint64_t)((g_last_mod + (uint32_t)R - (uint32_t)rho) % (uint32_t)R);
        bool ended = false;
        int last_k = 0;
#pragma unroll 1
        for (int k = 0; k < M; ++k) {
          const int w = il * M + k;
          if (!ended) {
            last_k = k;
            if (w_done[w]) ended = true;
          } else {
            const uint64_t gp = g_item + (uint64_t)k;
            const uint4 pw = philox(eng.seed, STREAM_PAD_ACTION, (uint32_t)e, (uint32_t)gp, (uint32_t)(gp >> 32));
            w_act[w] = (int)u_below(pw.x, (uint32_t)A);
            w_rew[w] = 0.f;
            w_term[w] = 1.f;
            const float* src = x_row(pl.n_on_tiles, il * M + last_k);
            float* dst = x_row(pl.n_on_tiles, w);
            for (int d = 0; d < D; ++d) dst[d] = src[d];
          }
        }
      }
      if (one_warp) __syncwarp(); else __syncthreads();
    }
    if (need_online_next && (!one_warp || warp == 0))
      for (int w = tid; w < nbM * D; w += (one_warp ? 32 : kSmThreads)) {
        const int r = w / D, d = w - r * D;
        x_row(0, Bc + r)[d] = x_row(pl.n_on_tiles, r)[d];
      }
    __syncthreads();
    if (eng.dbg_windows) {  // debug taps (tests): sampled slots, weights, rebuilt windows
      float* dw = eng.dbg_windows;
      const int n_states = B * (M + 1) * D, BM = B * M;
      for (int w = tid; w < nb * (M + 1) * D; w += kSmThreads) {
        const int il = w / ((M + 1) * D), rem = w - il * (M + 1) * D, k = rem / D, d = rem - k * D;
        dw[(size_t)i0 * (M + 1) * D + w] = (k == 0) ? x_row(0, il)[d] : x_row(pl.n_on_tiles, il * M + k - 1)[d];
      }
      for (int w = tid; w < nbM; w += kSmThreads) {
        dw[n_states + i0 * M + w] = (float)w_act[w];
        dw[n_states + BM + i0 * M + w] = w_rew[w];
        dw[n_states + 2 * BM + i0 * M + w] = w_term[w];
      }
      if (!per)  // (proportional replay: the replay CTA reports the tree indices and IS weights)
        for (int il = tid; il < nb; il += kSmThreads) {
          if (eng.dbg_sample_idx) eng.dbg_sample_idx[i0 + il] = (int64_t)slot_t[i0 + il];
          if (eng.dbg_weights) eng.dbg_weights[i0 + il] = 1.0f;
        }
    }
    SRLX_SMSTAMP(3);
    // ---------------------------------------------------------------- 3. forward: the row tiles of a layer spread over the warps
    auto tile_rows = [&](int t) -> int {
      const int rows = t < pl.n_on_tiles ? pl.n_on_rows - t * kRowTile : BcM - (t - pl.n_on_tiles) * kRowTile;
      return rows < kRowTile ? rows : kRowTile;
    };
#pragma unroll 1
    for (int l = 0; l < L - 1; ++l) {
      const int U = net.out_dim[l], K = net.k_dim[l];
      const int n_ut = (U + 31) >> 5;
      int base = 0;  // tasks (tile, 4-row group, 32-unit group) are dealt round-robin across the tiles
#pragma unroll 1
      for (int tile = 0; tile < pl.n_tiles; ++tile) {
        const int Rt = tile_rows(tile);
        const int n_task = ((Rt + 3) >> 2) * n_ut;
        const float* wset = tile < pl.n_on_tiles ? weff : wefft;
        float* a = acts + (size_t)tile * np.act_floats;
        int t = warp - (base & (nwarps - 1));
        if (t < 0) t += nwarps;
#pragma unroll 1
        for (; t < n_task; t += nwarps) {
          const int rt = t / n_ut, ut = t - rt * n_ut;
          dense_relu_task32(a + np.x_s[l], np.ldx[l], Rt, K, wset + np.w_s[l], np.ldw[l], wset + np.b_s[l], U, a + np.x_s[l + 1],
                            np.ldx[l + 1], rt, ut);
        }
        base += n_task;
      }
      __syncthreads();
    }
    SRLX_SMSTAMP(4);
    if (net.out_dim[L - 1] * net.k_dim[L - 1] <= 512) {
      // all tiles' rows in one sweep: thread = (tile, row)
      for (int w = tid; w < pl.n_tiles * kRowTile; w += kSmThreads) {
        const int tile = w / kRowTile, r = w - tile * kRowTile;
        if (r >= tile_rows(tile)) continue;
        const float* wset = tile < pl.n_on_tiles ? weff : wefft;
        float* a = acts + (size_t)tile * np.act_floats;
        float* q = tile < pl.n_on_tiles ? Q + (size_t)tile * kRowTile * A : Q + (size_t)(Bc + BcM + (tile - pl.n_on_tiles) * kRowTile) * A;
        small_out_q_row(net, a + np.x_s[L - 1] + r * np.ldx[L - 1], wset + np.w_s[L - 1], np.ldw[L - 1], wset + np.b_s[L - 1],
                        a + np.x_s[L] + r * np.ldx[L], q + r * A);
      }
      __syncthreads();
    } else {
#pragma unroll 1
      for (int tile = 0; tile < pl.n_tiles; ++tile) {
        const float* wset = tile < pl.n_on_tiles ? weff : wefft;
        float* a = acts + (size_t)tile * np.act_floats;
        small_out_layer(net, a + np.x_s[L - 1], np.ldx[L - 1], tile_rows(tile), wset + np.w_s[L - 1], np.ldw[L - 1], wset + np.b_s[L - 1],
                        a + np.x_s[L], np.ldx[L]);
      }
      __syncthreads();
#pragma unroll 1
      for (int tile = 0; tile < pl.n_tiles; ++tile) {
        float* a = acts + (size_t)tile * np.act_floats;
        // Q rows: online set first ([0, n_on_rows)), target rows at Bc + BcM
        float* q = tile < pl.n_on_tiles ? Q + (size_t)tile * kRowTile * A : Q + (size_t)(Bc + BcM + (tile - pl.n_on_tiles) * kRowTile) * A;
        small_dueling(net, a + np.x_s[L], np.ldx[L], tile_rows(tile), q, A);
      }
      __syncthreads();
    }
    SRLX_SMSTAMP(5);
    // ---------------------------------------------------------------- 4. targets, Huber gradient (thread per local item) | next sample
    if (warp == 0) {
      if (per) {  // the IS weights of this batch (the replay CTA sent them after the slots)
        mbar_wait_sleep(&mbar[4], par);
        if (lane == 0) sm_expect_tx(&mbar[4], (uint32_t)B * 4);
        __syncwarp();
      }
      const float* qon = Q + (size_t)Bc * A;          // online(s')  [BcM][A]
      const float* qtg = Q + (size_t)(Bc + BcM) * A;  // target(s')  [BcM][A]
      float lsum = 0.f;
      for (int il = lane; il < nb; il += 32) {
        const float gamma = (float)eng.discount;
        float target = 0.f, retrace = 1.f;
#pragma unroll 1
        for (int k = 0; k < M; ++k) {
          const float* qo = qon + (size_t)(il * M + k) * A;
          const float* qt = qtg + (size_t)(il * M + k) * A;
          const float* qsel = eng.enable_double_dqn ? qo : qt;
          int amx = 0;
          float best = qsel[0];
          for (int a = 1; a < A; ++a)
            if (qsel[a] > best) { best = qsel[a]; amx = a; }  // np.argmax: first max wins
          // Retrace with the reference's index shift (rainbow.py:267): action taken at s_k vs greedy action at s_{k+1}
          if (k >= 1) retrace = retrace * ((float)eng.retrace_h * ((w_act[il * M + k] == amx) ? 1.f : 0.f));
          float maxq = qt[amx];
          if (eng.enable_rescale) maxq = inverse_rescaling_f(maxq);
          float gain = w_rew[il * M + k] + ((1.0f - w_term[il * M + k]) * gamma) * maxq;
          if (eng.enable_rescale) gain = rescaling_f(gain);
          float qk = 0.f;  // the first step is learnt by the trainer itself (rainbow.py:232-234)
          if (k >= 1) qk = qon[(size_t)(il * M + k - 1) * A + w_act[il * M + k]];
          const float td = gain - qk;
          target += (td * disc_pow[k]) * retrace;
        }
        tq[il] = target;
        const int a0 = w_act[il * M + 0];
        const float q = Q[il * A + a0];
        qsa[il] = q;
        // Huber(target * w, q * w): the IS weight sits inside the loss argument (model_torch.py:115); 1 on uniform replay
        const float wgt = per ? per_w[(size_t)par * B + i0 + il] : 1.f;
        const float d = q * wgt - target * wgt;
        const float ad = fabsf(d);
        const float delta = (float)eng.huber_delta;
        lsum += (ad <= delta) ? 0.5f * d * d : delta * (ad - 0.5f * delta);
        const float dq = fminf(fmaxf(d, -delta), delta) * wgt / (float)B;
        if (per)  // priorities = abs(target - q), unweighted (model_torch.py:123) -> the replay CTA
          sm_st_async_f2(sm_mapa(smem_u32(tdbuf_remote + 2 * (i0 + il)), (uint32_t)C), fabsf(target - q), 0.f,
                         sm_mapa(smem_u32(&mbar[3]), (uint32_t)C));
        {  // dueling combine backward (dueling_network.py:51-58) -> d raw of the s row, where the backward pass picks it up
          const size_t trow = (size_t)(il / kRowTile) * np.act_floats + np.x_s[L] + (il % kRowTile) * np.ldx[L];
          float* dr = dacts + trow;
          if (net.dueling == SRLX_DUEL_NONE) {
            for (int a = 0; a < A; ++a) dr[a] = (a == a0) ? dq : 0.f;
          } else {
            const float* fr = acts + trow;  // forward raw outputs of the row
            int amax = 0;
            if (net.dueling == SRLX_DUEL_MAX) {
              float bestr = fr[1];
              for (int a = 1; a < A; ++a)
                if (fr[1 + a] > bestr) { bestr = fr[1 + a]; amax = a; }
            }
            for (int a = 0; a < A; ++a) {
              float dd = (a == a0) ? dq : 0.f;
              if (net.dueling == SRLX_DUEL_AVERAGE) dd -= dq / (float)A;
              else if (net.dueling == SRLX_DUEL_MAX && a == amax) dd -= dq;
              dr[1 + a] = dd;
            }
            dr[0] = dq;
          }
        }
        if (eng.dbg_target_q) eng.dbg_target_q[i0 + il] = target;
        if (eng.dbg_q_sa) eng.dbg_q_sa[i0 + il] = q;
      }
      for (int s = 16; s > 0; s >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, s);
      // the CTA's partial Huber sum -> CTA 0 (rides on the reduce-scatter barrier)
      if (lane == 0) sm_st_async_f2(sm_mapa(smem_u32(lossbuf + 2 * rank), 0u), lsum, 0.f, sm_mapa(mb_rs, 0u));
    } else if (warp == 1) {
      if (upd + 1 < n_updates && !per) small_sample(eng, tc + 1, B, n_valid, g_lo_mod, pick, slot + (par ^ 1) * B);
    } else if (tid == 64) {
      small_adam_scalars(eng, (double)(adam0 + upd + 1), sc);
    }
    __syncthreads();
    SRLX_SMSTAMP(6);
    // ---------------------------------------------------------------- prefetch: the next batch's ring rows -> registers (last warp)
    int pf_a = 0;
    float pf_rw = 0.f, pf_xv[4] = {0.f, 0.f, 0.f, 0.f}, pf_sv[4] = {0.f, 0.f, 0.f, 0.f};
    int pf_tm = 0, pf_dn = 0;
    const bool do_prefetch = can_prefetch && upd + 1 < n_updates && warp == nwarps - 1 && lane < nbM;  // the last warp: least work in the backward stages
    if (do_prefetch) {
      const int* slot_n = slot + (par ^ 1) * B;
      const int il = lane / M, k = lane - il * M;
      const int s0 = slot_n[i0 + il];
      const int rho = s0 / E, e = s0 - rho * E;
      const int sk = ((rho + k) % R) * E + e;
      pf_a = __ldcg(eng.ring_action + sk);
      pf_rw = __ldcg(eng.ring_reward + sk);
      pf_tm = (int)__ldcg(eng.ring_term + sk);
      pf_dn = (int)__ldcg(eng.ring_done + sk);
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        pf_xv[d] = d < D ? __ldcg(eng.ring_next_obs + (size_t)sk * D + d) : 0.f;
        pf_sv[d] = (d < D && k == 0) ? __ldcg(eng.ring_obs + (size_t)s0 * D + d) : 0.f;
      }
    }
    // ---------------------------------------------------------------- 5. backward on the CTA's s rows -> G (stored, not accumulated)
#pragma unroll 1
    for (int tile = 0; tile < pl.n_s_tiles; ++tile) {
      const int Rt = min(kRowTile, Bc - tile * kRowTile);
      long long* clk = (eng.dbg_clock && rank == 0 && tid == 0 && upd + 2 == n_updates) ? eng.dbg_clock : nullptr;
      small_delta_chain(net, np, weff, acts + (size_t)tile * np.act_floats, dacts + (size_t)tile * np.act_floats, Rt, clk);
      small_all_dw(net, np, acts + (size_t)tile * np.act_floats, dacts + (size_t)tile * np.act_floats, Rt, G, tile > 0, clk);
      if (clk) clk[13] = clock64();
    }
    sm_fence_proxy_async();  // this thread's gradient stores -> visible to the bulk-copy (async) proxy
    __syncthreads();
    if (do_prefetch) {  // the backward pass is done with the current rows: the next batch's rows take their place
      // (first use of the prefetched registers: the empty asm keeps the compiler from converting / consuming them any earlier,
      // which would make warp 2 wait for the loads at the start of the backward pass instead of here)
      asm volatile("" : "+r"(pf_a), "+r"(pf_tm), "+r"(pf_dn), "+f"(pf_rw), "+f"(pf_xv[0]), "+f"(pf_xv[1]), "+f"(pf_xv[2]), "+f"(pf_xv[3]),
                        "+f"(pf_sv[0]), "+f"(pf_sv[1]), "+f"(pf_sv[2]), "+f"(pf_sv[3]));
      const int il = lane / M, k = lane - il * M;
      float* xt = x_row(pl.n_on_tiles, lane);
      float* xs = x_row(0, il);
#pragma unroll
      for (int d = 0; d < 4; ++d)
        if (d < D) {
          xt[d] = pf_xv[d];
          if (k == 0) xs[d] = pf_sv[d];
        }
      w_act[lane] = pf_a;
      w_rew[lane] = pf_rw;
      w_term[lane] = (float)pf_tm;
      w_done[lane] = (int)pf_dn;
    }
    SRLX_SMSTAMP(7);
    // ---------------------------------------------------------------- 6. reduce-scatter the partial gradients, Adam on the slice
    {
      // slice c of the local gradient -> row `rank` of owner c's receive buffer: one bulk DSMEM copy per peer, plain stores
      // for the own slice
      for (int q = tid; q < (S >> 2); q += kSmThreads)
        reinterpret_cast<float4*>(recv + (size_t)rank * S)[q] = reinterpret_cast<const float4*>(G + (size_t)rank * S)[q];
      if (tid == 0) {
        for (int c = 0; c < C; ++c)
          if (c != rank)
            sm_bulk_s2c(sm_mapa(smem_u32(recv + (size_t)rank * S), (uint32_t)c), G + (size_t)c * S, (uint32_t)S * 4, sm_mapa(mb_rs, (uint32_t)c));
      }
      if (warp == 0) mbar_wait_sleep(&mbar[0], par);
      __syncthreads();
      if (tid == 0) {
        if (rank == 0) {  // the partial losses are read BEFORE this CTA's all-gather copies leave: no peer can be a phase ahead
          double l = 0.0;
          for (int c = 0; c < C; ++c) l += (double)lossbuf[2 * c];
          l /= (double)B;
          sc->last_loss = l;
          sc->loss_sum += l;
          if ((tc % (uint64_t)eng.target_update_interval) == 0) sc->sync_count += 1;
        }
        sm_expect_tx(&mbar[0], (uint32_t)((C - 1) * S * 4 + (rank == 0 ? C * 8 : 0)));  // arm the next phase
      }
      const float step_size = sc->step_size, bc2_sqrt = sc->bc2_sqrt;
      for (int j = tid; j < S; j += kSmThreads) {
        float g = 0.f;
        for (int c = 0; c < C; ++c) g += recv[(size_t)c * S + j];  // CTA order: deterministic
        const int p = p0 + j;
        float w = p < P ? weff[pslot[p]] : 0.f, m = am[j], v = av[j];
        adam_apply(w, m, v, g, b1, b2, aeps, step_size, bc2_sqrt);
        am[j] = m;
        av[j] = v;
        wnew[j] = w;
        wflat[p0 + j] = w;  // the own slice of the flat copy; peers get it by bulk copy below
        if (eng.dbg_grads && p < P) eng.dbg_grads[p] = g;
      }
      sm_fence_proxy_async();
      __syncthreads();
      if (tid == 0) {  // all-gather: the updated slice into every peer's flat parameter copy
        for (int c = 0; c < C; ++c)
          if (c != rank) sm_bulk_s2c(sm_mapa(smem_u32(wflat + (size_t)p0), (uint32_t)c), wnew, (uint32_t)S * 4, sm_mapa(mb_ag, (uint32_t)c));
      }
    }
    SRLX_SMSTAMP(8);
    // ---------------------------------------------------------------- 7. new parameters into the weight copies, target sync
    {
      if (warp == 0) mbar_wait_sleep(&mbar[1], par);
      __syncthreads();
      if (tid == 0) sm_expect_tx(&mbar[1], (uint32_t)((C - 1) * S * 4));
      const bool do_sync = (tc % (uint64_t)eng.target_update_interval) == 0;
      for (int p = tid; p < P; p += kSmThreads) {
        const int s = pslot[p];
        const float w = wflat[p];
        weff[s] = w;
        if (do_sync) wefft[s] = w;  // hard sync after the step, before train_count += 1 (model_torch.py:126-132)
      }
      // (every peer has finished its Adam -- its all-gather copy arrived -- so it has consumed this CTA's gradient slices: the
      // next backward may overwrite G; its first row tile stores, later tiles accumulate, the padding past P stays zero)
    }
    __syncthreads();
    SRLX_SMSTAMP(9);
  }

  // ---- write the state back: every CTA its slice -----------------------------------------------------------------------
  for (int j = tid; j < S; j += kSmThreads) {
    const int p = p0 + j;
    if (p < P) {
      const int s = pslot[p];
      __stcg(eng.params + p, weff[s]);
      __stcg(eng.target + p, wefft[s]);
      __stcg(eng.adam_m + p, am[j]);
      __stcg(eng.adam_v + p, av[j]);
    }
  }
  if (rank == 0 && tid == 0) {
    st->train_count = tc0 + n_updates;
    st->adam_step = adam0 + n_updates;
    st->last_loss = sc->last_loss;
    st->loss_sum += sc->loss_sum;
    st->sync_count += sc->sync_count;
  }
  cluster.sync();  // no CTA may exit while a peer can still address its shared memory
}

// Cluster size for the row-split kernel: SRLX_SMALL_CLUSTER or 8, halved until the batch gives every CTA an item and the
// plan fits one SM's shared memory.  Returns 0 when the kernel does not apply.
int small_choose(const srlx_engine* eng, size_t* smem_out, int* C_out) {
  if (!small_shape_ok(*eng)) return 0;
  int dev = 0, max_smem = 0;
  SRLX_CHECK_CUDA(cudaGetDevice(&dev));
  SRLX_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int want = 8;
  if (const char* e = getenv("SRLX_SMALL_CLUSTER")) want = atoi(e);
  if (want < 1) want = 1;
  if (want > kSmMaxCluster) want = kSmMaxCluster;
  for (int C = want; C >= 1; --C) {
    if ((C & (C - 1)) != 0 || C > eng->batch_size) continue;
    const SPlan pl = make_splan(*eng, C);
    if ((long long)pl.total + 1024 > max_smem || pl.np.weff_floats > 65535) continue;
    if (smem_out) *smem_out = pl.total;
    if (C_out) *C_out = C;
    return 1;
  }
  return 0;
}

int learn_small(const srlx_engine* eng, uint32_t n_updates, uintptr_t cuda_stream) {
  size_t total = 0;
  int C = 0;
  SRLX_REQUIRE(small_choose(eng, &total, &C) == 1, "learn_small: the row-split learner does not apply to this engine");
  SRLX_CHECK_CUDA(cudaFuncSetAttribute(learner_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));
  cudaLaunchConfig_t cfg = {};
  const unsigned n_cta = (unsigned)C + (eng->mem_kind == SRLX_MEM_PROPORTIONAL ? 1u : 0u);  // + the replay CTA
  if (n_cta > 8) SRLX_CHECK_CUDA(cudaFuncSetAttribute(learner_small_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cfg.gridDim = dim3(n_cta, 1, 1);
  cfg.blockDim = dim3(kSmThreads, 1, 1);
  cfg.dynamicSmemBytes = total;
  cfg.stream = (cudaStream_t)cuda_stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = n_cta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SRLX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, learner_small_kernel, *eng, n_updates));
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace srlx
