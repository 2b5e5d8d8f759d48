// learner.cu -- the trainer inner step as ONE persistent thread-block-CLUSTER kernel: n_updates consecutive
// Trainer.train() calls without returning to the host.  Consecutive updates are data-dependent (weights_t ->
// weights_t+1, priorities_t -> sample_t+1), so the limiter is the dependent-step latency, not a roofline; the design
// therefore shortens the critical path instead of widening it:
//
//   * a cluster of C CTAs (C <= 8, one SM each) shares one update.  The LAST HIDDEN layer is sharded by units
//     ("wide" layer: each CTA owns Us = U/C units, their slice of the output layer's columns and the matching Adam
//     state); every layer below it (the "trunk") is replicated.  A CTA's parameters, Adam moments, target copy and the
//     three effective-weight sets of an update live in ITS shared memory for the whole launch: HBM sees the weights
//     once at launch start and once at the end.
//   * per update the CTAs exchange only the partial output sums (rows x (1+A) floats) through distributed shared
//     memory + an mbarrier (and, if there is a trunk, the partial d(trunk output)); every CTA then holds identical Q
//     values and derives the identical targets / loss gradient, so backward and Adam need no further exchange.
//   * inside each CTA the warps are specialised: 12 "compute" warps (effective weights incl. NoisyNet draws, forward,
//     backward, Adam) and 4 "aux" warps (gather of the sampled windows, partial-sum reduction, n-step/Retrace targets,
//     Huber gradient; in CTA 0 also the SumTree priority update and the NEXT update's PER sample, broadcast to the
//     other CTAs).  The sample/gather chain of update t+1 overlaps backward + Adam of update t.
//   * CTA 0 keeps the top levels of the SumTree in shared memory (write-through), so a descent costs a few dependent
//     shared-memory reads plus 2 L2 round trips instead of 21.
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
#include <stdlib.h>

#include "cluster.cuh"
#include "net.cuh"
#include "tree.cuh"

namespace srlx {

constexpr int kLearnThreads = 512;
constexpr int kAuxWarps = 4, kCmpWarps = 12;
constexpr int NA = kAuxWarps * 32, NCP = kCmpWarps * 32;
constexpr int NSUB = 4;          // unit sub-chunks of the wide layer per row tile (work items = tiles x NSUB)
constexpr int kMaxSeg = 40;
constexpr int kSampleGroup = 8;  // descents a sampler warp keeps in flight
constexpr int kMaxCacheLevels = 14;
enum { BAR_X = 1, BAR_D = 2, BAR_CMP = 3, BAR_AUX = 4 };
enum { SEG_W = 0, SEG_LIN = 1, SEG_OUT = 2 };

// A contiguous run of the flat (global) parameter vector held by this CTA.
struct Seg {
  int g0, n;       // global flat offset, length
  int l0;          // offset in the CTA-local parameter arrays
  int kind;        // SEG_W: rows of K columns -> padded rows; SEG_LIN: contiguous; SEG_OUT: output-layer row -> WoT column
  int base, K, ld; // destination inside an effective-weight set (SEG_OUT: ld = output row o, base = first local unit)
  int noisy;       // the layer draws NoisyNet noise
  int replicated;  // every CTA holds (and identically updates) it; only rank 0 writes it back
};

struct LPlan {
  int C, L, lw, lo, Kw, Uh, Us, nout, NOP, Ko, ldw, ldh, ldx0, B, M, A, D, BM, NX, NRq, nSt, nNt, nTiles;
  int t_ldw[SRLX_MAX_LAYERS], t_ws[SRLX_MAX_LAYERS], t_bs[SRLX_MAX_LAYERS];  // trunk weights inside an effective-weight set
  int a_ld[SRLX_MAX_LAYERS + 1], a_s[SRLX_MAX_LAYERS + 1], act_floats;        // trunk activations of one row tile (l = 1..lw)
  int o_w, o_b, o_wot, o_bo, WS;  // wide W, wide b, transposed output weights, output bias; floats per set
  int Pl;                         // upper bound of CTA-local parameters
  int n_cache;                    // SumTree nodes cached in shared memory (2^levels - 1)
  size_t off_mbar, off_seg, off_scal, off_par, off_weff, off_xin, off_acts, off_hS, off_dH, off_slots, off_part, off_gpart,
      off_q, off_raw, off_samp_slot, off_samp_w, off_g, off_win, off_tq, off_draw, off_sidx, off_sdbl, off_cache, total;
};

__host__ __device__ inline int pick_cluster(const srlx_net& net, int want) {
  if (net.n_layers < 2) return 1;
  const int Uh = net.out_dim[net.n_layers - 2];
  int c = 8;
  if (want > 0) c = want;
  while (c > 1 && (Uh % c != 0 || Uh / c < 4)) c >>= 1;
  return c;
}

__host__ __device__ inline LPlan make_lplan(const srlx_engine& eng, int C, int max_smem, long long n_tree_nodes) {
  LPlan p;
  const srlx_net& net = eng.net;
  p.C = C;
  p.L = net.n_layers;
  p.lw = net.n_layers - 2;
  p.lo = net.n_layers - 1;
  p.Kw = net.k_dim[p.lw];
  p.Uh = net.out_dim[p.lw];
  p.Us = p.Uh / C;
  p.nout = net.out_dim[p.lo];
  p.NOP = round_up(p.nout, 4);
  p.Ko = net.k_dim[p.lo];
  p.ldw = padded_ld(p.Kw);
  p.ldh = padded_ld(p.Us);
  p.B = eng.batch_size;
  p.M = eng.multisteps;
  p.A = eng.n_actions;
  p.D = eng.obs_dim;
  p.ldx0 = padded_ld(p.D);
  p.BM = p.B * p.M;
  p.NX = p.B + p.BM;
  p.NRq = p.B + 2 * p.BM;
  p.nSt = (p.B + 31) / 32;
  p.nNt = (p.BM + 31) / 32;
  p.nTiles = p.nSt + 2 * p.nNt;
  int off = 0, ptrunk = 0;
  for (int l = 0; l < p.lw; ++l) {
    p.t_ldw[l] = padded_ld(net.k_dim[l]);
    p.t_ws[l] = off;
    off += net.out_dim[l] * p.t_ldw[l];
    p.t_bs[l] = off;
    off += round_up(net.out_dim[l], 4);
    ptrunk += net.out_dim[l] * (net.k_dim[l] + 1);
  }
  p.o_w = off;  off += p.Us * p.ldw;
  p.o_b = off;  off += round_up(p.Us, 4);
  p.o_wot = off; off += p.Us * p.NOP;
  p.o_bo = off; off += p.NOP;
  p.WS = round_up(off, 4);
  int aoff = 0;
  p.a_ld[0] = p.ldx0;
  p.a_s[0] = 0;
  for (int l = 1; l <= p.lw; ++l) {
    p.a_ld[l] = padded_ld(net.out_dim[l - 1]);
    p.a_s[l] = aoff;
    aoff += 32 * p.a_ld[l];
  }
  p.act_floats = round_up(aoff, 4);
  p.Pl = round_up(ptrunk + p.Us * p.Kw + p.Us + p.nout * p.Us + p.nout, 4);
  const int n_arrays = 10;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) / 16 * 16; return r; };
  p.off_mbar = take(64);
  p.off_seg = take(sizeof(Seg) * kMaxSeg);
  p.off_scal = take(256);
  p.off_par = take((size_t)n_arrays * p.Pl * 4);
  p.off_weff = take((size_t)3 * p.WS * 4);
  p.off_xin = take((size_t)2 * p.NX * p.ldx0 * 4);
  p.off_acts = take((size_t)p.nTiles * p.act_floats * 4);
  p.off_hS = take((size_t)p.nSt * 32 * p.ldh * 4);
  p.off_dH = take((size_t)p.nSt * 32 * p.ldh * 4);
  p.off_slots = take((size_t)NSUB * p.NRq * p.NOP * 4);
  p.off_part = take((size_t)2 * C * p.NRq * p.nout * 4);
  p.off_gpart = take(p.lw > 0 ? (size_t)C * p.B * p.Kw * 4 : 0);
  p.off_q = take((size_t)p.NRq * p.A * 4);
  p.off_raw = take((size_t)p.B * p.nout * 4);
  p.off_samp_slot = take((size_t)2 * p.B * 4);
  p.off_samp_w = take((size_t)2 * p.B * 4);
  p.off_g = take((size_t)4 * p.BM * 4);    // gathered action / reward / term / done
  p.off_win = take((size_t)4 * p.BM * 4);  // window action / reward / term / next-state invalid-action mask after padding
  p.off_tq = take((size_t)2 * p.B * 4);    // target_q, q(s,a)
  p.off_draw = take((size_t)p.B * p.NOP * 4);
  p.off_sidx = take((size_t)p.B * 8);
  p.off_sdbl = take((size_t)5 * p.B * 8);  // s_pri, s_new, s_chg, s_val, s_tmp
  // SumTree top levels: whatever shared memory is left (CTA 0 uses it; the layout is uniform across the cluster)
  p.n_cache = 0;
  if (eng.mem_kind == SRLX_MEM_PROPORTIONAL) {
    const long long avail = ((long long)max_smem - 1024 - (long long)o) / 8;
    int lev = 0;
    while (lev < kMaxCacheLevels && ((1ll << (lev + 1)) - 1) <= avail && ((1ll << (lev + 1)) - 1) <= n_tree_nodes) ++lev;
    p.n_cache = (int)((1ll << lev) - 1);
  }
  p.off_cache = take((size_t)p.n_cache * 8);
  p.total = o;
  return p;
}

#define SRLX_STAMP(cond, slot)                                                        \
  do {                                                                                \
    if (eng.dbg_clock && rank == 0 && (cond) && upd + 2 == n_updates) eng.dbg_clock[slot] = clock64(); \
  } while (0)

struct LScal {
  double total, beta, max_priority, loss_sum, last_loss;
  unsigned long long retries;
  float step_size, bc2_sqrt;
  unsigned int sync_count;
};

__device__ __forceinline__ void adam_apply_s(float& pp, float& mm, float& vv, float g, float b1, float b2, float eps,
                                             float step_size, float bc2_sqrt) {
  // torch/optim/adam.py _single_tensor_adam: lerp, mul_/addcmul_, sqrt/div/add_, addcdiv_
  mm = mm + (g - mm) * (1.0f - b1);
  vv = vv * b2 + (1.0f - b2) * g * g;
  const float denom = sqrtf(vv) / bc2_sqrt + eps;
  pp = pp - step_size * (mm / denom);
}

// slot of element j of segment s inside an effective-weight set
__device__ __forceinline__ int seg_slot(const Seg& s, int j, int NOP) {
  if (s.kind == SEG_W) {
    const int u = j / s.K, k = j - u * s.K;
    return s.base + u * s.ld + k;
  }
  if (s.kind == SEG_LIN) return s.base + j;
  return s.base + j * NOP + s.ld;  // SEG_OUT: base = o_wot + first_local_unit * NOP, ld = output row
}

// ---- lane-per-row dense layer (trunk): Y[u] = relu(b[u] + sum_k x[k] W[u][k]) for this lane's row ---------------------
__device__ __forceinline__ void lane_dense_relu(const float* __restrict__ x, int K, const float* __restrict__ W, int ldw,
                                                const float* __restrict__ b, int U, float* __restrict__ y) {
  const int K4 = round_up(K, 4);
  int u = 0;
  for (; u + 4 <= U; u += 4) {
    float a0 = b[u], a1 = b[u + 1], a2 = b[u + 2], a3 = b[u + 3];
    const float* w0 = W + u * ldw;
    for (int k = 0; k < K4; k += 4) {
      const float4 xv = *reinterpret_cast<const float4*>(x + k);
      const float4 wa = *reinterpret_cast<const float4*>(w0 + k);
      const float4 wb = *reinterpret_cast<const float4*>(w0 + ldw + k);
      const float4 wc = *reinterpret_cast<const float4*>(w0 + 2 * ldw + k);
      const float4 wd = *reinterpret_cast<const float4*>(w0 + 3 * ldw + k);
      a0 = fmaf(xv.x, wa.x, a0); a0 = fmaf(xv.y, wa.y, a0); a0 = fmaf(xv.z, wa.z, a0); a0 = fmaf(xv.w, wa.w, a0);
      a1 = fmaf(xv.x, wb.x, a1); a1 = fmaf(xv.y, wb.y, a1); a1 = fmaf(xv.z, wb.z, a1); a1 = fmaf(xv.w, wb.w, a1);
      a2 = fmaf(xv.x, wc.x, a2); a2 = fmaf(xv.y, wc.y, a2); a2 = fmaf(xv.z, wc.z, a2); a2 = fmaf(xv.w, wc.w, a2);
      a3 = fmaf(xv.x, wd.x, a3); a3 = fmaf(xv.y, wd.y, a3); a3 = fmaf(xv.z, wd.z, a3); a3 = fmaf(xv.w, wd.w, a3);
    }
    y[u] = fmaxf(a0, 0.f); y[u + 1] = fmaxf(a1, 0.f); y[u + 2] = fmaxf(a2, 0.f); y[u + 3] = fmaxf(a3, 0.f);
  }
  for (; u < U; ++u) {
    float a0 = b[u];
    const float* w0 = W + u * ldw;
    for (int k = 0; k < K4; k += 4) {
      const float4 xv = *reinterpret_cast<const float4*>(x + k);
      const float4 wa = *reinterpret_cast<const float4*>(w0 + k);
      a0 = fmaf(xv.x, wa.x, a0); a0 = fmaf(xv.y, wa.y, a0); a0 = fmaf(xv.z, wa.z, a0); a0 = fmaf(xv.w, wa.w, a0);
    }
    y[u] = fmaxf(a0, 0.f);
  }
}

// ---- one work item of the wide layer: rows of one tile (lane = row) x units [ub, ue) of this CTA's slice ---------------
//   h[u]     = relu(b[u] + sum_k x[k] W[u][k])
//   out[o]  += h[u] * WoT[u][o]            (partial output sums over this item's units)
// NOPT = padded output count (compile time), K4ONE: K <= 4 (x stays in registers).
template <int NOPT, bool K4ONE>
__device__ __forceinline__ void wide_item(const float* __restrict__ x, int K, const float* __restrict__ W, int ldw,
                                          const float* __restrict__ b, const float* __restrict__ WoT, int ub, int ue,
                                          float* __restrict__ h_row, float* __restrict__ out_row, bool row_valid) {
  float acc[NOPT];
#pragma unroll
  for (int o = 0; o < NOPT; ++o) acc[o] = 0.f;
  const int K4 = K4ONE ? 4 : round_up(K, 4);
  float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (K4ONE) x0 = *reinterpret_cast<const float4*>(x);
#pragma unroll 2
  for (int u = ub; u < ue; ++u) {
    float a = b[u];
    const float* w = W + u * ldw;
    if (K4ONE) {
      const float4 wv = *reinterpret_cast<const float4*>(w);
      a = fmaf(x0.x, wv.x, a); a = fmaf(x0.y, wv.y, a); a = fmaf(x0.z, wv.z, a); a = fmaf(x0.w, wv.w, a);
    } else {
      float a1 = 0.f;
      for (int k = 0; k < K4; k += 8) {
        const float4 xv = *reinterpret_cast<const float4*>(x + k);
        const float4 wv = *reinterpret_cast<const float4*>(w + k);
        a = fmaf(xv.x, wv.x, a); a = fmaf(xv.y, wv.y, a); a = fmaf(xv.z, wv.z, a); a = fmaf(xv.w, wv.w, a);
        if (k + 4 < K4) {
          const float4 xw = *reinterpret_cast<const float4*>(x + k + 4);
          const float4 ww = *reinterpret_cast<const float4*>(w + k + 4);
          a1 = fmaf(xw.x, ww.x, a1); a1 = fmaf(xw.y, ww.y, a1); a1 = fmaf(xw.z, ww.z, a1); a1 = fmaf(xw.w, ww.w, a1);
        }
      }
      a += a1;
    }
    a = fmaxf(a, 0.f);
    if (h_row) h_row[u] = a;
    const float* wo = WoT + u * NOPT;
#pragma unroll
    for (int o = 0; o < NOPT; o += 4) {
      const float4 wv = *reinterpret_cast<const float4*>(wo + o);
      acc[o] = fmaf(a, wv.x, acc[o]);
      acc[o + 1] = fmaf(a, wv.y, acc[o + 1]);
      acc[o + 2] = fmaf(a, wv.z, acc[o + 2]);
      acc[o + 3] = fmaf(a, wv.w, acc[o + 3]);
    }
  }
  if (row_valid) {
#pragma unroll
    for (int o = 0; o < NOPT; o += 4) *reinterpret_cast<float4*>(out_row + o) = make_float4(acc[o], acc[o + 1], acc[o + 2], acc[o + 3]);
  }
}

template <int NOPT>
__device__ __forceinline__ void wide_item_k(const float* x, int K, const float* W, int ldw, const float* b, const float* WoT,
                                            int ub, int ue, float* h_row, float* out_row, bool row_valid) {
  if (K <= 4) wide_item<NOPT, true>(x, K, W, ldw, b, WoT, ub, ue, h_row, out_row, row_valid);
  else wide_item<NOPT, false>(x, K, W, ldw, b, WoT, ub, ue, h_row, out_row, row_valid);
}

// =====================================================================================================================
__global__ void __launch_bounds__(kLearnThreads, 1)
learner_kernel(const __grid_constant__ srlx_engine eng, const uint32_t n_updates, const int max_smem) {
  extern __shared__ __align__(16) unsigned char smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const srlx_net& net = eng.net;
  const int64_t cap = (int64_t)eng.ring_rows * eng.n_envs, n_nodes = 2 * cap - 1;
  const LPlan pl = make_lplan(eng, C, max_smem, n_nodes);

  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + pl.off_mbar);  // [0] partial-Q exchange, [1] trunk-grad exchange, [2] sample
  Seg* segs = reinterpret_cast<Seg*>(smem + pl.off_seg);
  LScal* sc = reinterpret_cast<LScal*>(smem + pl.off_scal);
  int* n_seg_p = reinterpret_cast<int*>(smem + pl.off_scal + 128);
  float* par = reinterpret_cast<float*>(smem + pl.off_par);
  float *p_mu = par, *p_sg = par + pl.Pl, *p_m1 = par + 2 * pl.Pl, *p_v1 = par + 3 * pl.Pl, *p_m2 = par + 4 * pl.Pl,
        *p_v2 = par + 5 * pl.Pl, *p_tmu = par + 6 * pl.Pl, *p_tsg = par + 7 * pl.Pl, *p_eps = par + 8 * pl.Pl,
        *G = par + 9 * pl.Pl;
  float* weff = reinterpret_cast<float*>(smem + pl.off_weff);  // [3][WS]: 0 = online(s), 1 = online(s'), 2 = target(s')
  float* xin = reinterpret_cast<float*>(smem + pl.off_xin);    // [2][NX][ldx0]
  float* acts = reinterpret_cast<float*>(smem + pl.off_acts);  // [nTiles][act_floats]
  float* hS = reinterpret_cast<float*>(smem + pl.off_hS);
  float* dH = reinterpret_cast<float*>(smem + pl.off_dH);
  float* slots = reinterpret_cast<float*>(smem + pl.off_slots);  // [NSUB][NRq][NOP]
  float* part = reinterpret_cast<float*>(smem + pl.off_part);    // [2][C][NRq*nout]
  float* gpart = reinterpret_cast<float*>(smem + pl.off_gpart);  // [C][B*Kw]
  float* Q = reinterpret_cast<float*>(smem + pl.off_q);          // [NRq][A]
  float* rawS = reinterpret_cast<float*>(smem + pl.off_raw);     // [B][nout]
  int* samp_slot = reinterpret_cast<int*>(smem + pl.off_samp_slot);
  float* samp_w = reinterpret_cast<float*>(smem + pl.off_samp_w);
  int* g_act = reinterpret_cast<int*>(smem + pl.off_g);
  float* g_rew = reinterpret_cast<float*>(smem + pl.off_g) + pl.BM;
  float* g_term = g_rew + pl.BM;
  int* g_done = reinterpret_cast<int*>(g_term + pl.BM);
  int* w_act = reinterpret_cast<int*>(smem + pl.off_win);
  float* w_rew = reinterpret_cast<float*>(smem + pl.off_win) + pl.BM;
  float* w_term = w_rew + pl.BM;
  uint32_t* w_inv = reinterpret_cast<uint32_t*>(w_term + pl.BM);
  float* tq = reinterpret_cast<float*>(smem + pl.off_tq);
  float* qsa = tq + pl.B;
  float* dRaw = reinterpret_cast<float*>(smem + pl.off_draw);  // [B][NOP]
  int64_t* s_idx = reinterpret_cast<int64_t*>(smem + pl.off_sidx);
  double* s_pri = reinterpret_cast<double*>(smem + pl.off_sdbl);
  double *s_new = s_pri + pl.B, *s_chg = s_pri + 2 * pl.B, *s_val = s_pri + 3 * pl.B, *s_tmp = s_pri + 4 * pl.B;
  double* cache = reinterpret_cast<double*>(smem + pl.off_cache);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_aux = warp >= kCmpWarps;
  const int ct = tid, cw = warp;                             // compute-group thread / warp index
  const int at = tid - NCP, aw = warp - kCmpWarps;           // aux-group thread / warp index
  const int B = pl.B, M = pl.M, A = pl.A, D = pl.D, E = eng.n_envs, R = eng.ring_rows, BM = pl.BM;
  const int Us = pl.Us, nout = pl.nout, NOP = pl.NOP, lw = pl.lw;
  const int u0 = rank * Us;  // first unit of this CTA's slice of the wide layer
  const bool per = eng.mem_kind == SRLX_MEM_PROPORTIONAL;
  const bool noisy = net.noisy != 0;
  const bool need_online_next = eng.enable_double_dqn || M > 1;
  srlx_state* st = eng.state;

  // ---- launch-constant scalars (the rollout never runs concurrently with the learner) ---------------------------------
  const uint64_t tc0 = st->train_count, mem_size = st->mem_size, vec_steps = st->vec_steps, adam0 = st->adam_step;
  const bool go = mem_size >= eng.warmup_size && mem_size >= (uint64_t)B;
  if (!go) return;  // still warming up: train() returns without incrementing train_count (uniform across the cluster)

  // ---- one-time setup ---------------------------------------------------------------------------------------------
  if (tid == 0) {
    mbar_init(&mbar[0], C);
    mbar_init(&mbar[1], C);
    mbar_init(&mbar[2], 1);
    fence_mbar_init();
    int ns = 0, l0 = 0;
    auto add = [&](int g0, int n, int kind, int base, int K, int ld, int nz, int rep) {
      if (n <= 0) return;
      Seg s; s.g0 = g0; s.n = n; s.l0 = l0; s.kind = kind; s.base = base; s.K = K; s.ld = ld; s.noisy = nz; s.replicated = rep;
      segs[ns++] = s;
      l0 += n;
    };
    for (int l = 0; l < lw; ++l) {
      add(net.w_off[l], net.out_dim[l] * net.k_dim[l], SEG_W, pl.t_ws[l], net.k_dim[l], pl.t_ldw[l], net.layer_noisy[l], 1);
      add(net.b_off[l], net.out_dim[l], SEG_LIN, pl.t_bs[l], 1, 0, net.layer_noisy[l], 1);
    }
    add(net.w_off[lw] + u0 * pl.Kw, Us * pl.Kw, SEG_W, pl.o_w, pl.Kw, pl.ldw, net.layer_noisy[lw], 0);
    add(net.b_off[lw] + u0, Us, SEG_LIN, pl.o_b, 1, 0, net.layer_noisy[lw], 0);
    for (int o = 0; o < nout; ++o) {
      // output row o reads wide-layer units [koff, koff + Ko)
      const int koff = (net.dueling != SRLX_DUEL_NONE && o > 0) ? pl.Ko : 0;
      const int lo_u = max(u0, koff), hi_u = min(u0 + Us, koff + pl.Ko);
      add(net.w_off[pl.lo] + o * pl.Ko + (lo_u - koff), hi_u - lo_u, SEG_OUT, pl.o_wot + (lo_u - u0) * NOP, 1, o,
          net.layer_noisy[pl.lo], 0);
    }
    add(net.b_off[pl.lo], nout, SEG_LIN, pl.o_bo, 1, 0, net.layer_noisy[pl.lo], 1);
    *n_seg_p = ns;
    sc->loss_sum = 0.0;
    sc->last_loss = 0.0;
    sc->retries = 0;
    sc->sync_count = 0;
    sc->max_priority = st->max_priority;
  }
  for (int i = tid; i < 3 * pl.WS; i += kLearnThreads) weff[i] = 0.f;
  for (int i = tid; i < 2 * pl.NX * pl.ldx0; i += kLearnThreads) xin[i] = 0.f;
  for (int i = tid; i < pl.nTiles * pl.act_floats; i += kLearnThreads) acts[i] = 0.f;
  for (int i = tid; i < pl.nSt * 32 * pl.ldh; i += kLearnThreads) { hS[i] = 0.f; dH[i] = 0.f; }
  for (int i = tid; i < B * NOP; i += kLearnThreads) dRaw[i] = 0.f;
  for (int i = tid; i < 10 * pl.Pl; i += kLearnThreads) par[i] = 0.f;
  if (rank == 0)
    for (int i = tid; i < pl.n_cache; i += kLearnThreads) cache[i] = __ldcg(eng.tree + i);
  __syncthreads();
  const int n_seg = *n_seg_p;
  // load this CTA's parameters, Adam moments and target copy
  for (int s = 0; s < n_seg; ++s) {
    const Seg sg = segs[s];
    for (int j = tid; j < sg.n; j += kLearnThreads) {
      const int p = sg.g0 + j, i = sg.l0 + j;
      p_mu[i] = __ldcg(eng.params + p);
      p_tmu[i] = __ldcg(eng.target + p);
      p_m1[i] = __ldcg(eng.adam_m + p);
      p_v1[i] = __ldcg(eng.adam_v + p);
      if (noisy) {
        p_sg[i] = __ldcg(eng.params_sigma + p);
        p_tsg[i] = __ldcg(eng.target_sigma + p);
        p_m2[i] = __ldcg(eng.adam_m + net.n_params + p);
        p_v2[i] = __ldcg(eng.adam_v + net.n_params + p);
      }
    }
  }
  cluster.sync();  // mbarriers initialised + everyone's shared memory ready before any remote access

  // ==================================================================================================================
  if (!is_aux) {
    // ================================================ COMPUTE WARPS ================================================
    // Adam (update `upd`, if do_adam) + target sync + effective weights of update upd+1, one pass over the CTA's parameters
    auto finish = [&](bool do_adam, uint64_t tc_done, uint64_t tc_next) {
      const float b1 = (float)eng.adam_beta1, b2 = (float)eng.adam_beta2, aeps = (float)eng.adam_eps;
      const float step_size = sc->step_size, bc2_sqrt = sc->bc2_sqrt;
      const bool do_sync = do_adam && (tc_done % (uint64_t)eng.target_update_interval) == 0;
      for (int s = 0; s < n_seg; ++s) {
        const Seg sg = segs[s];
        const int blk0 = sg.g0 >> 2, blk1 = (sg.g0 + sg.n - 1) >> 2;
        const bool nz = noisy && sg.noisy;
        for (int blk = blk0 + ct; blk <= blk1; blk += NCP) {
          float zS[4] = {0.f, 0.f, 0.f, 0.f}, zN[4] = {0.f, 0.f, 0.f, 0.f}, zT[4] = {0.f, 0.f, 0.f, 0.f};
          if (nz) {
            const float4 a = noise4(eng.seed, NOISE_KIND_TRAIN, tc_next * 3 + 0, (uint32_t)blk);
            const float4 b = noise4(eng.seed, NOISE_KIND_TRAIN, tc_next * 3 + 1, (uint32_t)blk);
            const float4 c = noise4(eng.seed, NOISE_KIND_TRAIN, tc_next * 3 + 2, (uint32_t)blk);
            zS[0] = a.x; zS[1] = a.y; zS[2] = a.z; zS[3] = a.w;
            zN[0] = b.x; zN[1] = b.y; zN[2] = b.z; zN[3] = b.w;
            zT[0] = c.x; zT[1] = c.y; zT[2] = c.z; zT[3] = c.w;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int p = 4 * blk + q;
            if (p < sg.g0 || p >= sg.g0 + sg.n) continue;
            const int j = p - sg.g0, i = sg.l0 + j;
            float mu = p_mu[i], sgm = nz ? p_sg[i] : 0.f;
            if (do_adam) {
              const float g = G[i];
              float m = p_m1[i], v = p_v1[i];
              adam_apply_s(mu, m, v, g, b1, b2, aeps, step_size, bc2_sqrt);
              p_mu[i] = mu; p_m1[i] = m; p_v1[i] = v;
              const bool wr = eng.dbg_grads && (!sg.replicated || rank == 0);
              if (wr) eng.dbg_grads[p] = g;
              if (noisy) {
                float gs = 0.f;
                if (nz) {
                  gs = g * p_eps[i];
                  float m2 = p_m2[i], v2 = p_v2[i];
                  adam_apply_s(sgm, m2, v2, gs, b1, b2, aeps, step_size, bc2_sqrt);
                  p_sg[i] = sgm; p_m2[i] = m2; p_v2[i] = v2;
                }
                if (wr) eng.dbg_grads[net.n_params + p] = gs;
              }
              if (do_sync) { p_tmu[i] = mu; if (nz) p_tsg[i] = sgm; }
            }
            const int slot = seg_slot(sg, j, NOP);
            const float tmu = p_tmu[i], tsg = nz ? p_tsg[i] : 0.f;
            weff[slot] = fmaf(sgm, zS[q], mu);
            weff[pl.WS + slot] = fmaf(sgm, zN[q], mu);
            weff[2 * pl.WS + slot] = fmaf(tsg, zT[q], tmu);
            if (nz) p_eps[i] = zS[q];
          }
        }
      }
    };

    if (ct == 0) {
      const double t = (double)(adam0 + 1);
      sc->step_size = (float)(eng.lr / (1.0 - pow(eng.adam_beta1, t)));
      sc->bc2_sqrt = (float)sqrt(1.0 - pow(eng.adam_beta2, t));
    }
    named_bar_sync(BAR_CMP, NCP);
    finish(false, 0, tc0);
    named_bar_sync(BAR_CMP, NCP);

    for (uint32_t upd = 0; upd < n_updates; ++upd) {
      const uint64_t tc = tc0 + upd;
      const int parb = upd & 1;
      const float* x_cur = xin + (size_t)parb * pl.NX * pl.ldx0;
      named_bar_sync(BAR_X, kLearnThreads);  // x(t) gathered by the aux warps
      SRLX_STAMP(ct == 0, 0);

      // ---------------------------------------------------------------- trunk forward: warp per row tile, lane per row
      if (lw > 0) {
        for (int t = cw; t < pl.nTiles; t += kCmpWarps) {
          int set, row0, nrows;
          if (t < pl.nSt) { set = 0; row0 = t * 32; nrows = min(32, B - row0); }
          else if (t < pl.nSt + pl.nNt) { set = 1; row0 = B + (t - pl.nSt) * 32; nrows = min(32, B + BM - row0); }
          else { set = 2; row0 = B + (t - pl.nSt - pl.nNt) * 32; nrows = min(32, B + BM - row0); }
          if (set == 1 && !need_online_next) continue;
          const float* ws = weff + set * pl.WS;
          float* ta = acts + (size_t)t * pl.act_floats;
          const int r = min(lane, nrows - 1);
          const float* xrow = x_cur + (size_t)(row0 + r) * pl.ldx0;
          for (int l = 0; l < lw; ++l) {
            float* yrow = ta + pl.a_s[l + 1] + lane * pl.a_ld[l + 1];
            lane_dense_relu(xrow, net.k_dim[l], ws + pl.t_ws[l], pl.t_ldw[l], ws + pl.t_bs[l], net.out_dim[l], yrow);
            xrow = yrow;
          }
        }
        named_bar_sync(BAR_CMP, NCP);
      }
      SRLX_STAMP(ct == 0, 1);
      // ---------------------------------------------------------------- wide layer + partial outputs: (tile, unit chunk)
      {
        const int usub = (Us + NSUB - 1) / NSUB;
        for (int it = cw; it < pl.nTiles * NSUB; it += kCmpWarps) {
          const int t = it / NSUB, sub = it - t * NSUB;
          int set, row0, nrows, xrow0;
          if (t < pl.nSt) { set = 0; row0 = t * 32; nrows = min(32, B - row0); xrow0 = row0; }
          else if (t < pl.nSt + pl.nNt) { set = 1; const int j0 = (t - pl.nSt) * 32; row0 = B + j0; nrows = min(32, BM - j0); xrow0 = B + j0; }
          else { set = 2; const int j0 = (t - pl.nSt - pl.nNt) * 32; row0 = B + BM + j0; nrows = min(32, BM - j0); xrow0 = B + j0; }
          if (set == 1 && !need_online_next) continue;
          const float* ws = weff + set * pl.WS;
          const int r = min(lane, nrows - 1);
          const float* x;
          if (lw > 0) x = acts + (size_t)t * pl.act_floats + pl.a_s[lw] + r * pl.a_ld[lw];
          else x = x_cur + (size_t)(xrow0 + r) * pl.ldx0;
          const int ub = sub * usub, ue = min(Us, ub + usub);
          float* h_row = (set == 0) ? hS + (size_t)(row0 + r) * pl.ldh : nullptr;
          float* out_row = slots + ((size_t)sub * pl.NRq + row0 + r) * NOP;
          const bool valid = lane < nrows;
          if (NOP == 4) wide_item_k<4>(x, pl.Kw, ws + pl.o_w, pl.ldw, ws + pl.o_b, ws + pl.o_wot, ub, ue, h_row, out_row, valid);
          else if (NOP == 8) wide_item_k<8>(x, pl.Kw, ws + pl.o_w, pl.ldw, ws + pl.o_b, ws + pl.o_wot, ub, ue, h_row, out_row, valid);
          else if (NOP == 12) wide_item_k<12>(x, pl.Kw, ws + pl.o_w, pl.ldw, ws + pl.o_b, ws + pl.o_wot, ub, ue, h_row, out_row, valid);
          else if (NOP == 16) wide_item_k<16>(x, pl.Kw, ws + pl.o_w, pl.ldw, ws + pl.o_b, ws + pl.o_wot, ub, ue, h_row, out_row, valid);
          else wide_item_k<20>(x, pl.Kw, ws + pl.o_w, pl.ldw, ws + pl.o_b, ws + pl.o_wot, ub, ue, h_row, out_row, valid);
        }
      }
      named_bar_sync(BAR_CMP, NCP);
      SRLX_STAMP(ct == 0, 2);
      // ---------------------------------------------------------------- push this CTA's partial sums to every CTA
      {
        const int row_lo = need_online_next ? 0 : 0;
        const int n_vals = pl.NRq * nout;
        float* my_part = part + ((size_t)parb * C + rank) * n_vals;
        for (int w = ct; w < n_vals; w += NCP) {
          const int row = w / nout, o = w - row * nout;
          if (!need_online_next && row >= B && row < B + BM) continue;
          float v = 0.f;
#pragma unroll
          for (int sub = 0; sub < NSUB; ++sub) v += slots[((size_t)sub * pl.NRq + row) * NOP + o];
          for (int c = 0; c < C; ++c) map_rank(my_part, c)[w] = v;
        }
        (void)row_lo;
      }
      named_bar_sync(BAR_CMP, NCP);
      if (cw == 0 && lane < C) {
        fence_cluster();
        mbar_arrive_remote(&mbar[0], lane);
      }
      SRLX_STAMP(ct == 0, 3);
      // ---------------------------------------------------------------- wait for d(raw outputs) from the aux warps
      named_bar_sync(BAR_D, kLearnThreads);
      SRLX_STAMP(ct == 0, 4);
      // ---------------------------------------------------------------- backward of this CTA's slice
      for (int i = ct; i < pl.Pl; i += NCP) G[i] = 0.f;
      const float* wS = weff;  // the online(s) set
      // dH[r][u] = relu'(h) * sum_o dRaw[r][o] * WoT[u][o]
      for (int w = ct; w < B * Us; w += NCP) {
        const int r = w / Us, u = w - r * Us;
        const float* dr = dRaw + r * NOP;
        const float* wo = wS + pl.o_wot + u * NOP;
        float d = 0.f;
        for (int o = 0; o < nout; ++o) d = fmaf(dr[o], wo[o], d);
        dH[r * pl.ldh + u] = (hS[r * pl.ldh + u] > 0.f) ? d : 0.f;
      }
      named_bar_sync(BAR_CMP, NCP);
      {
        // segment-wise gradients
        const float* Xw_base;  // input of the wide layer for S rows
        for (int s = 0; s < n_seg; ++s) {
          const Seg sg = segs[s];
          if (sg.replicated && !(sg.kind == SEG_LIN && sg.base == pl.o_bo)) continue;  // trunk: after the exchange
          if (sg.kind == SEG_OUT) {  // dWout[o][u] = sum_r dRaw[r][o] * h[r][u]
            const int o = sg.ld, ul0 = (sg.base - pl.o_wot) / NOP;
            for (int j = ct; j < sg.n; j += NCP) {
              float a = 0.f;
              for (int r = 0; r < B; ++r) a = fmaf(dRaw[r * NOP + o], hS[r * pl.ldh + ul0 + j], a);
              G[sg.l0 + j] = a;
            }
          } else if (sg.kind == SEG_LIN && sg.base == pl.o_bo) {  // output bias
            for (int j = ct; j < sg.n; j += NCP) {
              float a = 0.f;
              for (int r = 0; r < B; ++r) a += dRaw[r * NOP + j];
              G[sg.l0 + j] = a;
            }
          } else if (sg.kind == SEG_LIN) {  // wide bias
            for (int j = ct; j < sg.n; j += NCP) {
              float a = 0.f;
              for (int r = 0; r < B; ++r) a += dH[r * pl.ldh + j];
              G[sg.l0 + j] = a;
            }
          } else {  // wide W: dW[u][k] = sum_r dH[r][u] * x[r][k]; item = (k4, u), u fastest
            const int K = pl.Kw, K4n = (K + 3) >> 2;
            for (int w = ct; w < Us * K4n; w += NCP) {
              const int k4 = w / Us, u = w - k4 * Us, k = k4 * 4;
              float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
              for (int r = 0; r < B; ++r) {
                const float dy = dH[r * pl.ldh + u];
                const float* xr = (lw > 0) ? acts + (size_t)(r >> 5) * pl.act_floats + pl.a_s[lw] + (r & 31) * pl.a_ld[lw]
                                           : x_cur + (size_t)r * pl.ldx0;
                const float4 xv = *reinterpret_cast<const float4*>(xr + k);
                a0 = fmaf(dy, xv.x, a0); a1 = fmaf(dy, xv.y, a1); a2 = fmaf(dy, xv.z, a2); a3 = fmaf(dy, xv.w, a3);
              }
              float* g = G + sg.l0 + u * K + k;
              g[0] = a0;
              if (k + 1 < K) g[1] = a1;
              if (k + 2 < K) g[2] = a2;
              if (k + 3 < K) g[3] = a3;
            }
          }
        }
        (void)Xw_base;
      }
      // ---------------------------------------------------------------- trunk: exchange d(trunk output), replicated backward
      if (lw > 0) {
        const int Kw = pl.Kw, n_vals = B * Kw;
        float* my_g = gpart + (size_t)rank * n_vals;
        for (int w = ct; w < n_vals; w += NCP) {
          const int r = w / Kw, k = w - r * Kw;
          float d = 0.f;
          for (int u = 0; u < Us; ++u) d = fmaf(dH[r * pl.ldh + u], wS[pl.o_w + u * pl.ldw + k], d);
          for (int c = 0; c < C; ++c) map_rank(my_g, c)[w] = d;
        }
        named_bar_sync(BAR_CMP, NCP);
        if (cw == 0 && lane < C) {
          fence_cluster();
          mbar_arrive_remote(&mbar[1], lane);
        }
        if (cw == 0) mbar_wait_sleep(&mbar[1], upd & 1);  // one polling warp; the others block on the named barrier
        named_bar_sync(BAR_CMP, NCP);
        // dA_lw[r][k] (in place in the stored activations), ReLU-masked, summed over the CTAs in rank order
        for (int w = ct; w < n_vals; w += NCP) {
          const int r = w / Kw, k = w - r * Kw;
          float d = 0.f;
          for (int c = 0; c < C; ++c) d += gpart[(size_t)c * n_vals + w];
          float* a = acts + (size_t)(r >> 5) * pl.act_floats + pl.a_s[lw] + (r & 31) * pl.a_ld[lw] + k;
          *a = (*a > 0.f) ? d : 0.f;
        }
        named_bar_sync(BAR_CMP, NCP);
        for (int l = lw - 1; l >= 0; --l) {
          const int U = net.out_dim[l], K = net.k_dim[l], K4n = (K + 3) >> 2;
          const Seg sW = segs[2 * l], sB = segs[2 * l + 1];
          auto dYp = [&](int r) { return acts + (size_t)(r >> 5) * pl.act_floats + pl.a_s[l + 1] + (r & 31) * pl.a_ld[l + 1]; };
          auto Xp = [&](int r) {
            return (l > 0) ? acts + (size_t)(r >> 5) * pl.act_floats + pl.a_s[l] + (r & 31) * pl.a_ld[l]
                           : const_cast<float*>(x_cur) + (size_t)r * pl.ldx0;
          };
          for (int w = ct; w < U * K4n; w += NCP) {
            const int k4 = w / U, u = w - k4 * U, k = k4 * 4;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            for (int r = 0; r < B; ++r) {
              const float dy = dYp(r)[u];
              const float4 xv = *reinterpret_cast<const float4*>(Xp(r) + k);
              a0 = fmaf(dy, xv.x, a0); a1 = fmaf(dy, xv.y, a1); a2 = fmaf(dy, xv.z, a2); a3 = fmaf(dy, xv.w, a3);
            }
            float* g = G + sW.l0 + u * K + k;
            g[0] = a0;
            if (k + 1 < K) g[1] = a1;
            if (k + 2 < K) g[2] = a2;
            if (k + 3 < K) g[3] = a3;
          }
          for (int u = ct; u < U; u += NCP) {
            float a = 0.f;
            for (int r = 0; r < B; ++r) a += dYp(r)[u];
            G[sB.l0 + u] = a;
          }
          named_bar_sync(BAR_CMP, NCP);
          if (l > 0) {
            const float* W = wS + pl.t_ws[l];
            const int ldw = pl.t_ldw[l];
            for (int w = ct; w < B * K; w += NCP) {
              const int r = w / K, k = w - r * K;
              float d = 0.f;
              const float* dy = dYp(r);
              for (int u = 0; u < U; ++u) d = fmaf(dy[u], W[u * ldw + k], d);
              float* x = Xp(r) + k;
              *x = (*x > 0.f) ? d : 0.f;
            }
            named_bar_sync(BAR_CMP, NCP);
          }
        }
      }
      SRLX_STAMP(ct == 0, 5);
      // ---------------------------------------------------------------- Adam + target sync + next effective weights
      if (ct == 0) {
        const double t = (double)(adam0 + upd + 1);
        sc->step_size = (float)(eng.lr / (1.0 - pow(eng.adam_beta1, t)));
        sc->bc2_sqrt = (float)sqrt(1.0 - pow(eng.adam_beta2, t));
      }
      named_bar_sync(BAR_CMP, NCP);
      finish(true, tc, tc + 1);
      if (ct == 0 && (tc % (uint64_t)eng.target_update_interval) == 0) sc->sync_count += 1;
      named_bar_sync(BAR_CMP, NCP);
      SRLX_STAMP(ct == 0, 6);
    }
  } else {
    // ================================================== AUX WARPS ==================================================
    const int64_t cap1 = cap - 1;
    // ---- PER sample of update `tc` (CTA 0): writes s_idx / s_pri, then slot + IS weight into buffer `parb` of every CTA
    auto sample_and_broadcast = [&](uint64_t tc, int parb) {
      int* slot_dst = samp_slot + parb * B;
      float* w_dst = samp_w + parb * B;
      if (per) {
        const double total = cache && pl.n_cache > 0 ? cache[0] : __ldcg(eng.tree);
        // PriorityReplayBuffer.step is the train_count of the PREVIOUS update() call (priority_replay_buffer.py:232,250)
        const double stepd = (tc > 0) ? (double)(tc - 1) : 0.0;
        double beta = eng.per_beta_initial + (1.0 - eng.per_beta_initial) * stepd / eng.per_beta_steps;
        beta = beta > 1.0 ? 1.0 : beta;
        auto draw = [&](int i, int k) -> double {
          const uint4 w = philox(eng.seed, STREAM_SAMPLE, (uint32_t)i | ((uint32_t)k << 16), (uint32_t)tc, (uint32_t)(tc >> 32));
          return u01_f64(w.x, w.y);
        };
        // (a) lane per sample: walk the cached top levels out of shared memory
        for (int i = at; i < B; i += NA) {
          double val = draw(i, 0) * total;
          int64_t idx = 0;
          while (2 * idx + 1 < (int64_t)pl.n_cache) {
            const double tl = cache[2 * idx + 1];
            if (val <= tl) idx = 2 * idx + 1;
            else { val -= tl; idx = 2 * idx + 2; }
          }
          s_idx[i] = idx;
          s_val[i] = val;
        }
        named_bar_sync(BAR_AUX, NA);
        // (b) warp per group of kSampleGroup samples: the remaining levels, 5 per L2 round trip, all descents in flight
        for (int g0 = aw * kSampleGroup; g0 < B; g0 += kAuxWarps * kSampleGroup) {
          int64_t idx[kSampleGroup];
          double val[kSampleGroup];
#pragma unroll
          for (int g = 0; g < kSampleGroup; ++g) {
            const int i = min(g0 + g, B - 1);
            idx[g] = s_idx[i];
            val[g] = s_val[i];
          }
          const int k_l = 32 - __clz(lane + 1);            // level (1..5) this lane fetches; lane 31 idles
          const int q_l = lane + 1 - (1 << (k_l - 1));     // it fetches the LEFT child at position 2*q_l of that level
          bool any = true;
          while (any) {
            double v[kSampleGroup];
            // unconditional loads from clamped addresses (a predicated load + select makes every load wait for the previous
            // one); values of nodes past the end of the tree are never looked at
#pragma unroll
            for (int g = 0; g < kSampleGroup; ++g) {
              int64_t node = ((idx[g] + 1) << k_l) - 1 + 2 * q_l;
              node = node < n_nodes ? node : n_nodes - 1;
              v[g] = __ldcg(eng.tree + node);
            }
            any = false;
#pragma unroll
            for (int g = 0; g < kSampleGroup; ++g) {
              int rel = 0;
#pragma unroll
              for (int k = 1; k <= 5; ++k) {
                const int64_t left = 2 * idx[g] + 1;
                const double tl = __shfl_sync(0xffffffffu, v[g], (1 << (k - 1)) - 1 + rel);
                if (left < n_nodes) {
                  if (val[g] <= tl) { idx[g] = left; rel = 2 * rel; }
                  else { idx[g] = left + 1; val[g] -= tl; rel = 2 * rel + 1; }
                }
              }
              any |= (2 * idx[g] + 1 < n_nodes);
            }
          }
          // leaf priorities; a zero-priority leaf is re-drawn (proportional_memory.py:150-152), sequentially.  Every lane of the warp read
          // s_idx / s_val of this group above; lane g overwrites them below: order the two (compute-sanitizer racecheck reported the
          // intra-warp read -> write pair, profiles/sanitizer/r1_o_racecheck.tail.txt)
          __syncwarp();
#pragma unroll
          for (int g = 0; g < kSampleGroup; ++g) {
            if (lane == g && g0 + g < B) {
              const int i = g0 + g;
              int64_t li = idx[g];
              double p = __ldcg(eng.tree + li);
              int k = 0;
              while (p == 0.0 && k + 1 < 9999) {
                ++k;
                li = tree_retrieve_seq(eng.tree, n_nodes, draw(i, k) * total);
                p = __ldcg(eng.tree + li);
              }
              s_idx[i] = li;
              s_pri[i] = p;
              s_tmp[i] = (double)k;
              if (k) atomicAdd(&sc->retries, (unsigned long long)k);
            }
          }
        }
        named_bar_sync(BAR_AUX, NA);
        if (!eng.has_duplicate) {
          if (at == 0) {
            for (int i = 1; i < B; ++i) {
              int k = (int)s_tmp[i];
              while (k < 9999) {
                bool dup = false;
                for (int j = 0; j < i; ++j) dup |= (s_idx[j] == s_idx[i]);
                if (!dup && s_pri[i] != 0.0) break;
                ++k;
                sc->retries += 1;
                if (k >= 9999) break;
                s_idx[i] = tree_retrieve_seq(eng.tree, n_nodes, draw(i, k) * total);
                s_pri[i] = __ldcg(eng.tree + s_idx[i]);
              }
            }
          }
          named_bar_sync(BAR_AUX, NA);
        }
        // (c) IS weights (proportional_memory.py:159-167)
        for (int i = at; i < B; i += NA) s_tmp[i] = pow((double)mem_size * (s_pri[i] / total), -beta);
        named_bar_sync(BAR_AUX, NA);
        if (aw == 0) {
          double mx = 0.0;
          for (int i = lane; i < B; i += 32) mx = fmax(mx, s_tmp[i]);
          for (int s = 16; s > 0; s >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, s));
          for (int i = lane; i < B; i += 32) s_val[i] = s_tmp[i] / mx;
        }
        named_bar_sync(BAR_AUX, NA);
      } else {
        // uniform replay: B distinct items (replay_buffer.py:34-36), sequential rejection
        if (at == 0) {
          const uint64_t g_next = vec_steps;
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
            s_idx[i] = (int64_t)((g % R) * E + pick % E) + cap1;  // stored as if it were a tree index
            s_val[i] = 1.0;
          }
        }
        named_bar_sync(BAR_AUX, NA);
      }
      // broadcast slot + weight to every CTA of the cluster, then signal their sample barrier
      for (int w = at; w < B * C; w += NA) {
        const int c = w / B, i = w - c * B;
        map_rank(slot_dst, c)[i] = (int)(s_idx[i] - cap1);
        map_rank(w_dst, c)[i] = (float)s_val[i];
      }
      named_bar_sync(BAR_AUX, NA);
      if (aw == 0 && lane < C) {
        fence_cluster();
        mbar_arrive_remote(&mbar[2], lane);
      }
    };

    if (rank == 0) sample_and_broadcast(tc0, 0);

    for (uint32_t upd = 0; upd < n_updates; ++upd) {
      const uint64_t tc = tc0 + upd;
      const int parb = upd & 1;
      float* x_cur = xin + (size_t)parb * pl.NX * pl.ldx0;
      const int* slot = samp_slot + parb * B;
      const float* s_w = samp_w + parb * B;
      if (aw == 0) mbar_wait_sleep(&mbar[2], upd & 1);  // sample(t) has arrived from CTA 0 (one polling warp)
      named_bar_sync(BAR_AUX, NA);
      SRLX_STAMP(at == 0, 16);
      // ---------------------------------------------------------------- gather the windows (one L2 round trip)
      for (int w = at; w < BM; w += NA) {
        const int i = w / M, k = w - i * M;
        const int s0 = slot[i];
        const int rho = s0 / E, e = s0 - rho * E;
        const int sk = ((rho + k) % R) * E + e;
        // all loads of this record are issued before the first dependent store
        const int a = __ldcg(eng.ring_action + sk);
        const float rw = __ldcg(eng.ring_reward + sk);
        const unsigned char tm = __ldcg(eng.ring_term + sk), dn = __ldcg(eng.ring_done + sk);
        float xv[SRLX_MAX_OBS];
#pragma unroll
        for (int d = 0; d < SRLX_MAX_OBS; ++d) xv[d] = d < D ? __ldcg(eng.ring_next_obs + (size_t)sk * D + d) : 0.f;
        g_act[w] = a;
        g_rew[w] = rw;
        g_term[w] = (float)tm;
        g_done[w] = (int)dn;
        w_inv[w] = eng.ring_invalid ? __ldcg(eng.ring_invalid + sk) : 0u;
        float* xr = x_cur + (size_t)(B + w) * pl.ldx0;
#pragma unroll
        for (int d = 0; d < SRLX_MAX_OBS; ++d)
          if (d < D) xr[d] = xv[d];
      }
      for (int w = at; w < B * D; w += NA) {
        const int i = w / D, d = w - i * D;
        x_cur[(size_t)i * pl.ldx0 + d] = __ldcg(eng.ring_obs + (size_t)slot[i] * D + d);
      }
      named_bar_sync(BAR_AUX, NA);
      for (int i = at; i < B; i += NA) {
        const int s0 = slot[i];
        const int rho = s0 / E, e = s0 - rho * E;
        const uint64_t g_last = vec_steps - 1;
        const uint64_t g_item = g_last - ((g_last + (uint64_t)R - (uint64_t)rho) % (uint64_t)R);
        bool ended = false;
        int last_k = 0;
        for (int k = 0; k < M; ++k) {
          const int w = i * M + k;
          if (!ended) {
            w_act[w] = g_act[w];
            w_rew[w] = g_rew[w];
            w_term[w] = g_term[w];
            last_k = k;
            if (g_done[w]) ended = true;
          } else {
            // padded tail record: random action, reward 0, terminated 1, state = last next_state (rainbow.py:358-371)
            const uint64_t gp = g_item + (uint64_t)k;
            const uint4 pw = philox(eng.seed, STREAM_PAD_ACTION, (uint32_t)e, (uint32_t)gp, (uint32_t)(gp >> 32));
            w_act[w] = (int)u_below(pw.x, (uint32_t)A);
            w_rew[w] = 0.f;
            w_term[w] = 1.f;
            w_inv[w] = 0u;  // padded records carry no invalid actions (rainbow.py:366)
            const float* src = x_cur + (size_t)(B + i * M + last_k) * pl.ldx0;
            float* dst = x_cur + (size_t)(B + w) * pl.ldx0;
            for (int d = 0; d < D; ++d) dst[d] = src[d];
          }
        }
      }
      named_bar_sync(BAR_AUX, NA);
      named_bar_arrive(BAR_X, kLearnThreads);  // x(t) ready for the compute warps
      SRLX_STAMP(at == 0, 17);

      // ---------------------------------------------------------------- reduce the partial sums, dueling combine
      if (aw == 0) mbar_wait_sleep(&mbar[0], upd & 1);
      named_bar_sync(BAR_AUX, NA);
      SRLX_STAMP(at == 0, 18);
      {
        const int n_vals = pl.NRq * nout;
        const float* pb = part + (size_t)parb * C * n_vals;
        for (int row = at; row < pl.NRq; row += NA) {
          const int set = row < B ? 0 : (row < B + BM ? 1 : 2);
          if (set == 1 && !need_online_next) continue;
          const float* bo = weff + set * pl.WS + pl.o_bo;
          float raw[SRLX_MAX_ACTIONS + 1];
          for (int o = 0; o < nout; ++o) {
            float v = 0.f;
            for (int c = 0; c < C; ++c) v += pb[(size_t)c * n_vals + row * nout + o];
            raw[o] = v + bo[o];
          }
          float* q = Q + row * A;
          if (net.dueling == SRLX_DUEL_NONE) {
            for (int a = 0; a < A; ++a) q[a] = raw[a];
          } else {
            const float v = raw[0];
            float red = 0.f;
            if (net.dueling == SRLX_DUEL_AVERAGE) {
              for (int a = 0; a < A; ++a) red += raw[1 + a];
              red = red / (float)A;
            } else if (net.dueling == SRLX_DUEL_MAX) {
              red = raw[1];
              for (int a = 1; a < A; ++a) red = fmaxf(red, raw[1 + a]);
            }
            for (int a = 0; a < A; ++a) q[a] = v + raw[1 + a] - red;
          }
          if (set == 0)
            for (int o = 0; o < nout; ++o) rawS[row * nout + o] = raw[o];
        }
      }
      named_bar_sync(BAR_AUX, NA);
      SRLX_STAMP(at == 0, 19);
      // ---------------------------------------------------------------- targets, Huber gradient (thread per sample)
      {
        const float* qon = Q + B * A;          // online(s')  [BM][A]
        const float* qtg = Q + (B + BM) * A;   // target(s')  [BM][A]
        // invalid actions, one-step targets (dqn.py:156-165, rainbow_nomultisteps.py:19-31): the masked entries take the MINIMUM OF THE
        // WHOLE BATCH'S Q matrix (online with double DQN, else target), not -inf
        float gmin = 0.f;
        if (eng.ring_invalid && M == 1) {
          __shared__ float s_gmin[kLearnThreads / 32];
          const float* qm = eng.enable_double_dqn ? qon : qtg;
          float m = INFINITY;
          for (int w = at; w < B * A; w += NA) m = fminf(m, qm[w]);
          for (int sft = 16; sft > 0; sft >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, sft));
          if ((at & 31) == 0) s_gmin[at >> 5] = m;
          named_bar_sync(BAR_AUX, NA);
          gmin = s_gmin[0];
          for (int w = 1; w < NA / 32; ++w) gmin = fminf(gmin, s_gmin[w]);
          named_bar_sync(BAR_AUX, NA);
        }
        float lsum = 0.f;
        for (int i = at; i < B; i += NA) {
          const float gamma = (float)eng.discount;
          float target = 0.f, retrace = 1.f;
          for (int k = 0; k < M; ++k) {
            const float* qo = qon + (size_t)(i * M + k) * A;
            const float* qt = qtg + (size_t)(i * M + k) * A;
            const float* qsel = eng.enable_double_dqn ? qo : qt;
            const uint32_t inv = w_inv[i * M + k];
            // masked value: the batch minimum (one-step) or -inf (n-step, rainbow.py:244-249)
            const float fill = M == 1 ? gmin : -INFINITY;
            int am = 0;
            float best = (inv & 1u) ? fill : qsel[0];
            for (int a = 1; a < A; ++a) {
              const float v = ((inv >> a) & 1u) ? fill : qsel[a];
              if (v > best) { best = v; am = a; }  // np.argmax: first max wins
            }
            // Retrace with the reference's index shift (rainbow.py:267): action taken at s_k vs greedy action at s_{k+1}
            if (k >= 1) retrace = retrace * ((float)eng.retrace_h * ((w_act[i * M + k] == am) ? 1.f : 0.f));
            // the value comes from the TARGET net; without double DQN the target matrix itself was overwritten at the masked entries
            float maxq = (!eng.enable_double_dqn && ((inv >> am) & 1u)) ? fill : qt[am];
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
          const float w = s_w[i];
          const float d = q * w - target * w;
          const float ad = fabsf(d);
          const float delta = (float)eng.huber_delta;
          lsum += (ad <= delta) ? 0.5f * d * d : delta * (ad - 0.5f * delta);
          const float dq = fminf(fmaxf(d, -delta), delta) * w / (float)B;
          // dueling combine backward (dueling_network.py:51-58) -> d raw
          float* dr = dRaw + i * NOP;
          if (net.dueling == SRLX_DUEL_NONE) {
            for (int a = 0; a < A; ++a) dr[a] = (a == a0) ? dq : 0.f;
          } else {
            int amax = 0;
            if (net.dueling == SRLX_DUEL_MAX) {
              float bestr = rawS[i * nout + 1];
              for (int a = 1; a < A; ++a)
                if (rawS[i * nout + 1 + a] > bestr) { bestr = rawS[i * nout + 1 + a]; amax = a; }
            }
            for (int a = 0; a < A; ++a) {
              float dd = (a == a0) ? dq : 0.f;
              if (net.dueling == SRLX_DUEL_AVERAGE) dd -= dq / (float)A;
              else if (net.dueling == SRLX_DUEL_MAX && a == amax) dd -= dq;
              dr[1 + a] = dd;
            }
            dr[0] = dq;
          }
          if (per && rank == 0) s_new[i] = pow(fabs((double)fabsf(target - q)) + eng.per_epsilon, eng.per_alpha);
        }
        // loss = mean over the batch (every aux warp holds a partial; combine through shared scratch)
        for (int s = 16; s > 0; s >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, s);
        if (lane == 0) s_tmp[aw] = (double)lsum;  // s_tmp is free here (weights already broadcast)
      }
      named_bar_sync(BAR_AUX, NA);
      named_bar_arrive(BAR_D, kLearnThreads);  // d(raw) ready: the compute warps start backward
      SRLX_STAMP(at == 0, 20);
      if (rank == 0 && at == 0) {
        double l = 0.0;
        for (int w = 0; w < kAuxWarps; ++w) l += s_tmp[w];
        l /= (double)B;
        sc->last_loss = l;
        sc->loss_sum += l;
      }
      // debug taps (CTA 0)
      if (rank == 0) {
        if (eng.dbg_sample_idx)
          for (int i = at; i < B; i += NA) eng.dbg_sample_idx[i] = per ? s_idx[i] : (s_idx[i] - cap1);
        if (eng.dbg_weights)
          for (int i = at; i < B; i += NA) eng.dbg_weights[i] = s_w[i];
        if (eng.dbg_target_q)
          for (int i = at; i < B; i += NA) eng.dbg_target_q[i] = tq[i];
        if (eng.dbg_q_sa)
          for (int i = at; i < B; i += NA) eng.dbg_q_sa[i] = qsa[i];
        if (eng.dbg_windows) {
          float* dw = eng.dbg_windows;
          const int n_states = B * (M + 1) * D;
          for (int w = at; w < n_states; w += NA) {
            const int i = w / ((M + 1) * D), rem = w - i * (M + 1) * D, k = rem / D, d = rem - k * D;
            dw[w] = (k == 0) ? x_cur[(size_t)i * pl.ldx0 + d] : x_cur[(size_t)(B + i * M + k - 1) * pl.ldx0 + d];
          }
          for (int w = at; w < BM; w += NA) {
            dw[n_states + w] = (float)w_act[w];
            dw[n_states + BM + w] = w_rew[w];
            dw[n_states + 2 * BM + w] = w_term[w];
          }
        }
      }
      named_bar_sync(BAR_AUX, NA);  // s_tmp / loss consumed before the sampler reuses the scratch
      SRLX_STAMP(at == 0, 21);
      // ---------------------------------------------------------------- CTA 0: priorities -> tree, then sample t+1
      if (rank == 0) {
        if (per) {
          // ProportionalMemory.update (proportional_memory.py:171-177), bit-identical to the sequential loop:
          // per-item change in item order for duplicate leaves ...
          for (int i = at; i < B; i += NA) {
            const int64_t li = s_idx[i];
            double prev = s_pri[i];
            for (int j = i - 1; j >= 0; --j)
              if (s_idx[j] == li) { prev = s_new[j]; break; }
            s_chg[i] = s_new[i] - prev;
            bool last = true;
            for (int j = i + 1; j < B; ++j)
              if (s_idx[j] == li) { last = false; break; }
            if (last) {
              __stcg(eng.tree + li, s_new[i]);
              if (li < (int64_t)pl.n_cache) cache[li] = s_new[i];
            }
          }
          named_bar_sync(BAR_AUX, NA);
          // ... and every ancestor receives the changes of the items below it in item order: warp per tree level,
          // lane per item, the first lane of each group of equal nodes applies the whole group.
          const int dmax = 63 - __clzll((long long)n_nodes);  // depth of the deepest leaf
          for (int a = aw; a < dmax; a += kAuxWarps) {
            for (int c0 = 0; c0 < B; c0 += 32) {
              const int i = c0 + lane;
              const bool valid = i < B;
              const long long ip1 = valid ? (long long)s_idx[i] + 1 : 1;
              const int d = 63 - __clzll(ip1);
              const bool has = valid && d > a;
              const long long node = has ? (ip1 >> (d - a)) - 1 : -1 - (long long)lane;
              const unsigned mask = __match_any_sync(0xffffffffu, node);
              if (has && lane == __ffs(mask) - 1) {
                double v = __ldcg(eng.tree + node);  // write-through keeps the global tree current
                unsigned m = mask;
                while (m) {
                  const int j = __ffs(m) - 1;
                  m &= m - 1;
                  v += s_chg[c0 + j];
                }
                __stcg(eng.tree + node, v);
                if (node < (long long)pl.n_cache) cache[node] = v;
              }
              __syncwarp();
            }
          }
          if (at == 0) {
            double mp = sc->max_priority;
            for (int i = 0; i < B; ++i) mp = (mp < s_new[i]) ? s_new[i] : mp;
            sc->max_priority = mp;
          }
          named_bar_sync(BAR_AUX, NA);
        }
        SRLX_STAMP(at == 0, 22);
        if (upd + 1 < n_updates) sample_and_broadcast(tc + 1, parb ^ 1);
        SRLX_STAMP(at == 0, 23);
      }
    }
  }

  // ---- write back: parameters, Adam moments, target copy, counters ----------------------------------------------------
  __syncthreads();
  for (int s = 0; s < n_seg; ++s) {
    const Seg sg = segs[s];
    if (sg.replicated && rank != 0) continue;
    for (int j = tid; j < sg.n; j += kLearnThreads) {
      const int p = sg.g0 + j, i = sg.l0 + j;
      __stcg(eng.params + p, p_mu[i]);
      __stcg(eng.target + p, p_tmu[i]);
      __stcg(eng.adam_m + p, p_m1[i]);
      __stcg(eng.adam_v + p, p_v1[i]);
      if (noisy) {
        __stcg(eng.params_sigma + p, p_sg[i]);
        __stcg(eng.target_sigma + p, p_tsg[i]);
        __stcg(eng.adam_m + net.n_params + p, p_m2[i]);
        __stcg(eng.adam_v + net.n_params + p, p_v2[i]);
      }
    }
  }
  if (rank == 0 && tid == 0) {
    st->train_count = tc0 + n_updates;
    st->adam_step = adam0 + n_updates;
    st->max_priority = sc->max_priority;
    st->sample_retries += sc->retries;
    st->last_loss = sc->last_loss;
    st->loss_sum += sc->loss_sum;
    st->sync_count += sc->sync_count;
  }
  cluster.sync();  // no CTA may exit while a peer can still address its shared memory
}

int learn_generic(const srlx_engine* eng, uint32_t n_updates, uintptr_t cuda_stream);

}  // namespace srlx

// ---- host side ---------------------------------------------------------------------------------------------------
// generic kernel (any MLP depth); srlx_learn (learner_fast.cu) dispatches here when the short-critical-path kernel does
// not apply
int srlx::learn_generic(const srlx_engine* eng, uint32_t n_updates, uintptr_t cuda_stream) {
  using namespace srlx;
  SRLX_REQUIRE(eng != nullptr, "srlx_learn: eng is NULL");
  SRLX_REQUIRE(eng->batch_size >= 1 && eng->batch_size <= SRLX_MAX_BATCH, "batch_size %d out of range [1,%d]", eng->batch_size, SRLX_MAX_BATCH);
  SRLX_REQUIRE(eng->multisteps >= 1 && eng->multisteps <= SRLX_MAX_MULTISTEPS, "multisteps %d out of range", eng->multisteps);
  SRLX_REQUIRE(eng->n_actions >= 1 && eng->n_actions <= SRLX_MAX_ACTIONS, "n_actions %d out of range", eng->n_actions);
  SRLX_REQUIRE(eng->net.n_layers >= 2 && eng->net.n_layers <= SRLX_MAX_LAYERS,
               "the fused learner needs at least one hidden layer (n_layers = %d)", eng->net.n_layers);
  SRLX_REQUIRE(eng->mem_kind == SRLX_MEM_UNIFORM || eng->tree != nullptr, "proportional memory needs a tree buffer");
  SRLX_REQUIRE(!eng->net.noisy || (eng->params_sigma && eng->target_sigma), "noisy net needs sigma buffers");
  SRLX_REQUIRE(eng->state && eng->params && eng->target && eng->adam_m && eng->adam_v, "srlx_learn: parameter buffer is NULL");
  SRLX_REQUIRE(eng->ring_obs && eng->ring_next_obs && eng->ring_action && eng->ring_reward && eng->ring_term && eng->ring_done,
               "srlx_learn: ring buffer pointer is NULL");
  if (n_updates == 0) return 0;
  int dev = 0, max_smem = 0;
  SRLX_CHECK_CUDA(cudaGetDevice(&dev));
  SRLX_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int want = 0;
  if (const char* e = getenv("SRLX_CLUSTER")) want = atoi(e);
  const long long n_nodes = 2ll * eng->ring_rows * eng->n_envs - 1;
  int C = pick_cluster(eng->net, want);
  LPlan pl = make_lplan(*eng, C, max_smem, n_nodes);
  // the exchange buffers grow with the cluster size: fall back to a narrower cluster when the widest does not fit
  while ((long long)pl.total + 1024 > max_smem && C > 1) {
    C = pick_cluster(eng->net, C / 2);
    pl = make_lplan(*eng, C, max_smem, n_nodes);
  }
  SRLX_REQUIRE((long long)pl.total + 1024 <= max_smem,
               "network / batch too large for the fused learner: needs %zu bytes of shared memory per CTA, device allows %d",
               pl.total + 1024, max_smem);
  SRLX_CHECK_CUDA(cudaFuncSetAttribute(learner_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.total));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)C, 1, 1);
  cfg.blockDim = dim3(kLearnThreads, 1, 1);
  cfg.dynamicSmemBytes = pl.total;
  cfg.stream = (cudaStream_t)cuda_stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SRLX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, learner_kernel, *eng, n_updates, max_smem));
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
