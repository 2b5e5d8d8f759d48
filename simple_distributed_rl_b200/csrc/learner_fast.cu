// learner_fast.cu -- the trainer inner step for SINGLE-HIDDEN-LAYER Q-networks (the reference's Rainbow default: dueling
// (512,) head on a <= 4-float observation; srl/algorithms/rainbow/rainbow.py:57-108, dueling_network.py:12-14), as one
// persistent thread-block-cluster kernel whose critical path per update is a handful of dependent hops:
//
//   forward (all rows, units sharded over the C CTAs, weights in registers, packed FFMA2)
//     -> reduce-scatter of the partial output sums to the CTA that OWNS the batch item   (st.async + mbarrier tx, DSMEM)
//     -> owner: dueling combine, n-step/Retrace target, Huber gradient                    (B/C items per CTA)
//     -> all-gather of d(raw outputs) + (target, q)                                        (st.async + mbarrier tx)
//     -> backward + Adam + next effective weights (every CTA, its slice)      ||   CTA 0: priorities -> SumTree update
//                                                                                   -> PER sample(t+1) -> slots to all
//                                                                                   CTAs -> every CTA gathers x(t+1)
//
// What is NOT on the critical path any more (compared with learner.cu, which stays as the generic kernel for deeper nets):
//   * NoisyNet draws: noise_precompute_kernel fills HBM with the N(0,1) tensors of a whole chunk of updates using all
//     148 SMs (3 x P Gaussians per update; Philox + Box-Muller is ~150 instructions per 4 values), laid out per CTA in
//     CTA-local parameter order; the learner streams its slice with one cp.async.bulk (TMA, mbarrier complete_tx) per
//     update, two updates ahead, into a 3-deep shared-memory ring.
//   * Adam bias-correction pow()s: tabulated per launch.
//   * IS weights: sent after the sampled slots (only the Huber step needs them).
// 16 compute warps + 4 memory warps per CTA; the memory warps run one update ahead (sample/gather of t+1 overlaps
// backward + Adam of t).  Same arithmetic, reference line citations and CPU twin (oracle/engine.py::OracleEngine.learn)
// as learner.cu.
#include <stdlib.h>

#include "cluster.cuh"
#include "net.cuh"
#include "tree.cuh"

namespace srlx {

int learn_generic(const srlx_engine* eng, uint32_t n_updates, uintptr_t cuda_stream);  // learner.cu

constexpr int kFCmpWarps = 16, kFMemWarps = 4;
constexpr int FNC = kFCmpWarps * 32, FNM = kFMemWarps * 32, FNT = FNC + FNM;
constexpr int kFMaxC = 16, kFMaxSeg = 8, kFCacheLevels = 12, kFMaxChunk = 256, kFSubLd = 66;
enum { FBAR_CMP = 1, FBAR_MEM = 2, FBAR_MA = 3, FBAR_MB = 4 };
enum { FSEG_W = 0, FSEG_B = 1, FSEG_O = 2, FSEG_OB = 3 };
enum { MB_RS = 0, MB_AG, MB_S, MB_WT, MB_XR, MB_NZ0, MB_NZ1, MB_NZ2, MB_TQ, MB_COUNT };

// A contiguous run of the flat (global) parameter vector held by one CTA.
struct FSeg {
  int g0, n, l0;  // global flat offset, length, offset in the CTA-local parameter arrays
  int kind, o;    // FSEG_*; output row for FSEG_O
  int ulo;        // FSEG_O: first local unit the run covers
  int noisy, replicated;
};

struct FPlan {
  int C, B, M, A, D, K, nout, Uh, Us, nCh, BM, NX, NRq, ipc, rpi, nOwn, Pl, WS, nRG, rpg, B4, n_cache;
  size_t off_mbar, off_seg, off_scal, off_adam, off_gpow, off_gidx, off_slot, off_par, off_nz, off_weff, off_xin, off_meta,
      off_part, off_rs, off_ag, off_tq, off_qown, off_rawown, off_samp_slot, off_samp_w, off_sdbl, off_plan, off_sub, off_cache, total;
};

__host__ __device__ inline int fast_pick_cluster(const srlx_net& net, int want) {
  const int Uh = net.out_dim[0];
  int c = want > 0 ? want : 16;
  if (c > kFMaxC) c = kFMaxC;
  while (c > 1 && (Uh % c != 0 || (Uh / c) % 4 != 0)) c >>= 1;
  if (Uh % c != 0 || (Uh / c) % 4 != 0) return 0;
  return c;
}

// the runs of CTA `rank`; returns their number
__host__ __device__ inline int fast_segs(const srlx_net& net, int C, int rank, FSeg* out) {
  const int K = net.k_dim[0], Uh = net.out_dim[0], Us = Uh / C, u0 = rank * Us, nout = net.out_dim[1], Ko = net.k_dim[1];
  int ns = 0, l0 = 0;
  auto add = [&](int g0, int n, int kind, int o, int ulo, int nz, int rep) {
    if (n <= 0) return;
    FSeg s;
    s.g0 = g0; s.n = n; s.l0 = l0; s.kind = kind; s.o = o; s.ulo = ulo; s.noisy = nz; s.replicated = rep;
    out[ns++] = s;
    l0 += n;
  };
  add(net.w_off[0] + u0 * K, Us * K, FSEG_W, 0, 0, net.layer_noisy[0], 0);
  add(net.b_off[0] + u0, Us, FSEG_B, 0, 0, net.layer_noisy[0], 0);
  for (int o = 0; o < nout; ++o) {
    const int koff = (net.dueling != SRLX_DUEL_NONE && o > 0) ? Ko : 0;  // output row o reads wide units [koff, koff+Ko)
    const int lo_u = u0 > koff ? u0 : koff, hi_u = (u0 + Us) < (koff + Ko) ? (u0 + Us) : (koff + Ko);
    add(net.w_off[1] + o * Ko + (lo_u - koff), hi_u - lo_u, FSEG_O, o, lo_u - u0, net.layer_noisy[1], 0);
  }
  add(net.b_off[1], nout, FSEG_OB, 0, 0, net.layer_noisy[1], 1);
  return ns;
}

__host__ __device__ inline bool fast_shape_ok(const srlx_engine& eng) {
  const srlx_net& net = eng.net;
  return net.n_layers == 2 && eng.obs_dim >= 1 && eng.obs_dim <= 4 && net.k_dim[0] == eng.obs_dim && net.out_dim[1] <= 4 &&
         eng.n_actions <= 4 && eng.batch_size >= 1 && eng.batch_size <= 32 && eng.multisteps >= 1 &&
         eng.multisteps <= SRLX_MAX_MULTISTEPS && (2ll * eng.ring_rows * eng.n_envs) < (1ll << 27);
}

__host__ __device__ inline FPlan make_fplan(const srlx_engine& eng, int C, long long n_tree_nodes, int max_cache_levels) {
  FPlan p;
  const srlx_net& net = eng.net;
  p.C = C;
  p.B = eng.batch_size;
  p.M = eng.multisteps;
  p.A = eng.n_actions;
  p.D = eng.obs_dim;
  p.K = net.k_dim[0];
  p.nout = net.out_dim[1];
  p.Uh = net.out_dim[0];
  p.Us = p.Uh / C;
  p.nCh = p.Us / 4;
  p.BM = p.B * p.M;
  p.NX = p.B + p.BM;
  p.NRq = p.B + 2 * p.BM;
  p.ipc = (p.B + C - 1) / C;
  p.rpi = 1 + 2 * p.M;
  p.nOwn = p.ipc * p.rpi;
  p.Pl = round_up(p.Us * (p.K + 1) + p.nout * p.Us + p.nout, 4);
  p.WS = p.nCh * 36 + 4;
  p.nRG = 1;
  while (p.nRG * 2 * p.Us <= FNC && p.nRG * 2 <= 32) p.nRG *= 2;
  p.rpg = (p.B + p.nRG - 1) / p.nRG;
  p.B4 = round_up(p.B, 4);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) / 16 * 16; return r; };
  p.off_mbar = take(8 * MB_COUNT);
  p.off_seg = take(sizeof(FSeg) * kFMaxSeg);
  p.off_scal = take(256);
  p.off_adam = take((size_t)2 * kFMaxChunk * 4);
  p.off_gpow = take((size_t)SRLX_MAX_MULTISTEPS * 4);
  p.off_gidx = take((size_t)p.Pl * 4);
  p.off_slot = take((size_t)p.Pl * 4);
  p.off_par = take((size_t)8 * p.Pl * 4);
  p.off_nz = take(net.noisy ? (size_t)9 * p.Pl * 4 : 0);
  p.off_weff = take((size_t)3 * p.WS * 4);
  p.off_xin = take((size_t)2 * p.NX * 16);
  p.off_meta = take((size_t)2 * 3 * p.BM * 4);  // [2][act, rew, term][BM]
  {
    const size_t a = (size_t)kFCmpWarps * p.NRq * 16, b = (size_t)p.nRG * p.Pl * 4;
    p.off_part = take(a > b ? a : b);
  }
  p.off_rs = take((size_t)2 * C * p.nOwn * 16);
  p.off_ag = take((size_t)2 * p.B * 32);
  p.off_tq = take((size_t)2 * p.B * 8);  // CTA 0: (target, q) of every item, sent ahead of the all-gather
  p.off_qown = take((size_t)p.nOwn * 16);
  p.off_rawown = take((size_t)p.ipc * 16);
  p.off_samp_slot = take((size_t)2 * p.B4 * 4);
  p.off_samp_w = take((size_t)2 * p.B4 * 4);
  p.off_sdbl = take(2048);  // s_idx, s_att, sperm[4][32] (int), then s_pri, s_tmp (double)
  p.off_plan = take(eng.mem_kind == SRLX_MEM_PROPORTIONAL ? (size_t)FNM * (5 * 8 + 2 * 9 * 4 + 5 * 4 + 4 + 8) : 0);  // old[5], node[9], end[9], blocked address[5], pad, old leaf
  p.n_cache = 0;
  if (eng.mem_kind == SRLX_MEM_PROPORTIONAL) {
    p.off_sub = take((size_t)kFMemWarps * 8 * kFSubLd * 8);
    int lev = 0;
    while (lev < max_cache_levels && lev < kFCacheLevels && ((1ll << (lev + 1)) - 1) <= n_tree_nodes) ++lev;
    p.n_cache = (int)((1ll << lev) - 1);
  } else {
    p.off_sub = take(0);
  }
  p.off_cache = take((size_t)p.n_cache * 8);
  p.total = o;
  return p;
}

// ---- blocked copy of the deep SumTree levels ---------------------------------------------------------------------------
// Below the `clev` levels cached in shared memory the sampler descends five levels per memory round trip.  In the flat
// (BFS) layout the 62 nodes of a 5-level subtree sit in five separate runs (6-8 cache lines, two loads per lane); the
// blocked copy stores every such subtree contiguously -- block = [level 1: 2][level 2: 4]...[level 5: 32] doubles, 512-byte
// stride -- so lane l's 16-byte load at offset 16*l fetches both children of the subtree's l-th node: one coalesced load
// per lane, 4 lines per subtree.  Tier r holds the subtrees rooted at level (clev-1) + 5r.
struct BlkPlan {
  int n_tiers, total;   // total blocks
  int first[4], off[4]; // first node index of the tier's root level, block offset of the tier
};
__host__ __device__ inline BlkPlan make_blk_plan(long long n_nodes, int clev) {
  BlkPlan p;
  p.n_tiers = 0;
  p.total = 0;
  for (int r = 0; r < 4; ++r) { p.first[r] = 0; p.off[r] = 0; }
  if (clev < 1) return p;
  for (int r = 0; r < 4; ++r) {
    const int L = clev - 1 + 5 * r;
    if (L > 28) break;
    const long long first = (1ll << L) - 1;
    if (first >= n_nodes || 2 * first + 1 >= n_nodes) break;
    const long long nb = (1ll << L) < (n_nodes - first) ? (1ll << L) : (n_nodes - first);
    p.first[r] = (int)first;
    p.off[r] = p.total;
    p.total += (int)nb;
    p.n_tiers = r + 1;
  }
  return p;
}
// element index (in doubles) of tree node `node` (level >= clev) inside the blocked copy
__device__ __forceinline__ int blk_addr(const BlkPlan& bp, int clev, int node) {
  const int a = 31 - __clz(node + 1);  // level of the node
  const int r = (a - clev) / 5, k = (a - clev) - 5 * r + 1;
  const int root = ((node + 1) >> k) - 1, j = (node + 1) - ((root + 1) << k);
  const int first = r == 0 ? bp.first[0] : (r == 1 ? bp.first[1] : (r == 2 ? bp.first[2] : bp.first[3]));
  const int off = r == 0 ? bp.off[0] : (r == 1 ? bp.off[1] : (r == 2 ? bp.off[2] : bp.off[3]));
  return (off + root - first) * 64 + (1 << k) - 2 + j;
}
__global__ void __launch_bounds__(256) tree_blk_build_kernel(const double* __restrict__ tree, const int n_nodes, const int clev,
                                                             double* __restrict__ blk) {
  const BlkPlan bp = make_blk_plan(n_nodes, clev);
  const int b = blockIdx.x * 4 + (threadIdx.x >> 6), e = threadIdx.x & 63;
  if (b >= bp.total) return;
  int r = 0;
  for (int t = 1; t < bp.n_tiers; ++t)
    if (b >= bp.off[t]) r = t;
  const int root = bp.first[r] + (b - bp.off[r]);
  double v = 0.0;
  if (e < 62) {
    const int k = 31 - __clz(e + 2), j = e + 2 - (1 << k);
    const long long node = (((long long)root + 1) << k) - 1 + j;
    if (node < n_nodes) v = __ldcg(tree + node);
  }
  blk[(size_t)b * 64 + e] = v;
}

// ---- PTX helpers: DSMEM stores that complete a transaction count on the destination CTA's mbarrier, TMA bulk copy ------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_f4(uint32_t raddr, float4 v, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(rmbar)
               : "memory");
}
__device__ __forceinline__ void st_async_f2(uint32_t raddr, float a, float b, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(raddr), "f"(a),
               "f"(b), "r"(rmbar)
               : "memory");
}
__device__ __forceinline__ void st_async_i4(uint32_t raddr, int4 v, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.s32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rmbar)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// ---- data-parallel learner: exchange buffer of one rank (peers write into it over NVLink) -------------------------
// Every 64-bit word carries its own flag: {payload: 32 bits, update number (train_count + 1): 32 bits}.  A sender just stores
// the words into the peer's buffer (no fence, no separate flag: an aligned 8-byte store lands whole); a receiver polls each word
// until its tag is the update it waits for -- one NVLink store latency per exchange instead of store + fence + flag.  Two
// parities, so a rank one update ahead never overwrites words its peer has not read yet.
//   [0, 1024)   uint64 wsc[2 parities][8 src ranks][8]: the halves of {shard total, shard size, min priority of the batch, max_priority}
//   [2048, ...) uint64 grad[2 parities][8 src ranks][16 CTAs][PlPad]: a CTA's gradient slice
constexpr int kDpMaxWorld = 8, kDpHeader = 2048;
__host__ __device__ inline size_t dp_bytes_for(int PlPad) { return (size_t)kDpHeader + (size_t)2 * kDpMaxWorld * kFMaxC * PlPad * 8; }
__device__ __forceinline__ unsigned long long* dp_wsc(void* base, int par, int src) {
  return reinterpret_cast<unsigned long long*>(base) + (par * kDpMaxWorld + src) * 8;
}
__device__ __forceinline__ unsigned long long* dp_grad(void* base, int par, int src, int cta, int PlPad) {
  return reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(base) + kDpHeader) + ((size_t)(par * kDpMaxWorld + src) * kFMaxC + cta) * PlPad;
}
__device__ __forceinline__ void dp_store(unsigned long long* p, uint32_t payload, uint32_t tag) {
  const unsigned long long v = ((unsigned long long)tag << 32) | payload;
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// poll one word until it carries `tag`; gives up after ~1 s (a peer that died must not hang this GPU): *dead is then set
__device__ __noinline__ uint32_t dp_load(const unsigned long long* p, uint32_t tag, volatile int* dead) {
  unsigned long long v;
  long long t0 = 0;
  for (int it = 0;; ++it) {
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if ((uint32_t)(v >> 32) == tag || *dead) break;
    __nanosleep(100);  // thousands of threads poll at once: back off so the replay warps keep their L2 bandwidth
    if (it == 256) t0 = clock64();
    if (it > 256 && clock64() - t0 > 2000000000ll) { *dead = 1; break; }
  }
  return (uint32_t)v;
}
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// x^a on the dependent chain td -> priority -> SumTree update -> next sample, in ~70 fp64 instructions instead of the
// ~300 of exp(a*log(x)) / pow(): what the chain pays for is instruction count (tools/ubench: a cold instruction stream
// runs at ~7 cycles/instruction), not flops.  ln x = e ln2 + 2 atanh((m-1)/(m+1)) with m in [sqrt(1/2), sqrt(2))
// (odd series to s^23), a*ln x carried as hi + lo, exp by Cody-Waite reduction + degree-13 Taylor.  Max relative
// difference to libm pow over x in [1e-14, 1e5], a in [-1, 1]: 3.6e-15 (tests/test_oracle_golden.py restates the check).
// Out of line on purpose: the per-update code of a CTA should stay inside the SM's instruction cache.
__constant__ double kPowAtanh[11] = {1.0 / 23.0, 1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0,
                                     1.0 / 11.0, 1.0 / 9.0,  1.0 / 7.0,  1.0 / 5.0,  1.0 / 3.0};
__constant__ double kPowExp[14] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
                                   1.0 / 40320.0,      1.0 / 5040.0,      1.0 / 720.0,      1.0 / 120.0,     1.0 / 24.0,
                                   1.0 / 6.0,          0.5,               1.0,              1.0};
__device__ __noinline__ double pow_chain(double x, double a) {
  bool ok = (x >= 1e-290 && x <= 1e290) && (fabs(a) <= 8.0);
  double res = 0.0;
  if (ok) {
    const long long bits = __double_as_longlong(x);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    double m = __longlong_as_double((bits & 0x000fffffffffffffll) | 0x3ff0000000000000ll);
    if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
    const double s = (m - 1.0) / (m + 1.0), s2 = s * s;
    double p = kPowAtanh[0];
#pragma unroll
    for (int i = 1; i < 11; ++i) p = fma(p, s2, kPowAtanh[i]);
    const double lnm = 2.0 * fma(p * s2, s, s);
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double lx_hi = fma((double)e, LN2_HI, lnm), lx_lo = (double)e * LN2_LO;
    const double y = a * lx_hi, y_lo = fma(a, lx_hi, -y) + a * lx_lo;
    ok = fabs(y) <= 690.0;
    const double k = rint(y * 1.44269504088896338700e+00);
    double r = fma(-k, LN2_HI, y);
    r = fma(-k, LN2_LO, r) + y_lo;
    double q = kPowExp[0];
#pragma unroll
    for (int i = 1; i < 14; ++i) q = fma(q, r, kPowExp[i]);
    res = __longlong_as_double(__double_as_longlong(q) + ((long long)k << 52));
  }
  if (!ok) res = pow(x, a);  // zeros, denormals, huge exponents: the library routine
  return res;
}
__device__ __noinline__ uint4 philox_ni(uint64_t seed, uint32_t stream, uint32_t a, uint32_t b, uint32_t c) {
  return philox(seed, stream, a, b, c);
}

struct FScal {
  double max_priority, loss_sum, last_loss;
  unsigned long long retries;
  unsigned int sync_count;
};

// phase clocks (tools/phase_clocks.py): compiled in only with -DSRLX_STAMPS (libsrlx_stamps.so), they cost code space
#ifdef SRLX_STAMPS
#define SRLX_FSTAMP(cond, slot)                                                                                    \
  do {                                                                                                             \
    if (eng.dbg_clock && rank == 0 && (cond) && upd + 2 == n_updates) eng.dbg_clock[slot] = clock64();             \
  } while (0)
#else
#define SRLX_FSTAMP(cond, slot) do { } while (0)
#endif

// =====================================================================================================================
// N(0,1) tensors of `n_updates` consecutive updates, CTA-local order: out[((u * C + rank) * 3 + set) * Pl + i].
// set 0 = online(s), 1 = online(s'), 2 = target(s'): NoisyLinear draws per forward call (noisy_linear.py:35-52), call ids
// as in learner.cu (train_count * 3 + set).
__global__ void __launch_bounds__(128) noise_precompute_kernel(const __grid_constant__ srlx_engine eng, const uint32_t n_updates,
                                                               const int C, const int Pl, float* __restrict__ out) {
  __shared__ FSeg segs[kFMaxSeg];
  __shared__ int n_seg_s;
  const int u = blockIdx.x / C, rank = blockIdx.x - u * C;
  if (threadIdx.x == 0) n_seg_s = fast_segs(eng.net, C, rank, segs);
  __syncthreads();
  const uint64_t tc = eng.state->train_count + (uint64_t)u;
  float* dst = out + (size_t)blockIdx.x * 3 * Pl;
  for (int s = 0; s < n_seg_s; ++s) {
    const FSeg sg = segs[s];
    const int blk0 = sg.g0 >> 2, blk1 = (sg.g0 + sg.n - 1) >> 2;
    for (int blk = blk0 + (int)threadIdx.x; blk <= blk1; blk += (int)blockDim.x) {
      float z[3][4] = {};
      if (sg.noisy) {
#pragma unroll
        for (int set = 0; set < 3; ++set) {
          const float4 a = noise4(eng.learner_seed ? eng.learner_seed : eng.seed, NOISE_KIND_TRAIN, tc * 3 + set, (uint32_t)blk);
          z[set][0] = a.x; z[set][1] = a.y; z[set][2] = a.z; z[set][3] = a.w;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int p = 4 * blk + q;
        if (p < sg.g0 || p >= sg.g0 + sg.n) continue;
        const int i = sg.l0 + (p - sg.g0);
#pragma unroll
        for (int set = 0; set < 3; ++set) dst[(size_t)set * Pl + i] = z[set][q];
      }
    }
  }
}

// =====================================================================================================================
// TC = 8 / 16: the reference's Rainbow default shape (batch 32, 3-step, 2 actions, 4 observation floats, dueling-average
// head of 2 x 512 units) on a cluster of TC CTAs, every loop bound a compile-time constant; TC = 0: same code, run-time bounds.
// DP: the data-parallel variant (gradient / replay-scalar exchange with the peer ranks); the single-GPU instantiations carry none of
// its code (the per-update instruction stream is what bounds this kernel).
template <int TC, bool DP>
__global__ void __launch_bounds__(FNT, 1)
learner_fast_kernel(const __grid_constant__ srlx_engine eng, const uint32_t n_updates, const float* __restrict__ noise,
                    const int max_cache_levels) {
  extern __shared__ __align__(16) unsigned char smem[];
  cg::cluster_group cluster = cg::this_cluster();
  constexpr bool FLAG = TC > 0;
  const int C = FLAG ? TC : (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const srlx_net& net = eng.net;
  const int cap = eng.ring_rows * eng.n_envs, n_nodes = 2 * cap - 1, cap1 = cap - 1;
  const FPlan pl = make_fplan(eng, C, n_nodes, max_cache_levels);

  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + pl.off_mbar);
  FSeg* segs = reinterpret_cast<FSeg*>(smem + pl.off_seg);
  FScal* sc = reinterpret_cast<FScal*>(smem + pl.off_scal);
  int* n_seg_p = reinterpret_cast<int*>(smem + pl.off_scal + 128);
  int* n_used_p = n_seg_p + 1;
  volatile int* dp_dead = n_seg_p + 2;  // a peer rank stopped answering: no further waits in this launch
  float* adam_ss = reinterpret_cast<float*>(smem + pl.off_adam);  // step_size per update of this launch
  float* adam_bc = adam_ss + kFMaxChunk;                          // sqrt(bias_correction2)
  float* gpow = reinterpret_cast<float*>(smem + pl.off_gpow);
  int* gidx = reinterpret_cast<int*>(smem + pl.off_gidx);
  int* slot_t = reinterpret_cast<int*>(smem + pl.off_slot);
  float* par = reinterpret_cast<float*>(smem + pl.off_par);
  const int Pl = FLAG ? (8 * (1024 / (FLAG ? TC : 1)) + 4) : pl.Pl;
  float *p_mu = par, *p_sg = par + Pl, *p_m1 = par + 2 * Pl, *p_v1 = par + 3 * Pl, *p_m2 = par + 4 * Pl,
        *p_v2 = par + 5 * Pl, *p_tmu = par + 6 * Pl, *p_tsg = par + 7 * Pl;
  float* nzr = reinterpret_cast<float*>(smem + pl.off_nz);      // [3 ring][3 set][Pl]
  float* weff = reinterpret_cast<float*>(smem + pl.off_weff);   // [3][WS]
  float* xin = reinterpret_cast<float*>(smem + pl.off_xin);     // [2][NX][4]
  float* meta = reinterpret_cast<float*>(smem + pl.off_meta);   // [2][3][BM]
  float* part = reinterpret_cast<float*>(smem + pl.off_part);   // [16][NRq][4]  (forward)  /  [nRG][Pl] (backward)
  float* rs = reinterpret_cast<float*>(smem + pl.off_rs);       // [2][C][nOwn][4]
  float* ag = reinterpret_cast<float*>(smem + pl.off_ag);       // [2][B][8]: d(raw)[4], target, q, loss term, -
  float* tqb = reinterpret_cast<float*>(smem + pl.off_tq);      // [2][B][2]
  float* qown = reinterpret_cast<float*>(smem + pl.off_qown);   // [nOwn][4]
  float* rawown = reinterpret_cast<float*>(smem + pl.off_rawown);
  int* samp_slot = reinterpret_cast<int*>(smem + pl.off_samp_slot);
  float* samp_w = reinterpret_cast<float*>(smem + pl.off_samp_w);
  int* s_idx = reinterpret_cast<int*>(smem + pl.off_sdbl);  // tree index of each sampled item (slot + cap - 1)
  int* s_att = s_idx + 32;
  int* sperm = s_idx + 64;                                  // [4 warps][32]
  double* s_pri = reinterpret_cast<double*>(smem + pl.off_sdbl + 1024);
  double* s_tmp = s_pri + 32;
  double* s_wval = s_tmp + 32;                              // hand-over of the top-level walk: remaining value,
  int* s_widx = reinterpret_cast<int*>(s_wval + 32);        // node reached
  double* s_tot = reinterpret_cast<double*>(s_widx + 32);   // tree total the current batch was sampled under
  double* sub = reinterpret_cast<double*>(smem + pl.off_sub);
  double* plan_old = reinterpret_cast<double*>(smem + pl.off_plan);
  int* plan_node = reinterpret_cast<int*>(smem + pl.off_plan + (size_t)FNM * 5 * 8);
  int* plan_end = plan_node + 9 * FNM;
  int* plan_baddr = plan_end + 9 * FNM;
  double* plan_oldleaf = reinterpret_cast<double*>(plan_baddr + 6 * FNM);  // [FNM] the leaf's value before the update
  double* cache = reinterpret_cast<double*>(smem + pl.off_cache);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = FLAG ? 32 : pl.B, M = FLAG ? 3 : pl.M, A = FLAG ? 2 : pl.A, D = FLAG ? 4 : pl.D, K = D;
  const int E = eng.n_envs, R = eng.ring_rows, BM = B * M, NRq = B + 2 * BM, NX = B + BM, B4 = FLAG ? 32 : pl.B4;
  const int Us = FLAG ? 1024 / (FLAG ? TC : 1) : pl.Us, nout = FLAG ? 3 : pl.nout, nCh = Us / 4, WS = nCh * 36 + 4, rpi = 1 + 2 * M;
  const int ipc = FLAG ? 32 / (FLAG ? TC : 1) : pl.ipc, nOwn = ipc * rpi, nRG = FLAG ? (TC == 8 ? 4 : 8) : pl.nRG,
            rpg = FLAG ? (TC == 8 ? 8 : 4) : pl.rpg;
  const int dueling = FLAG ? (int)SRLX_DUEL_AVERAGE : net.dueling;
  const int n_cache = pl.n_cache;
  const int u0 = rank * Us;
  const bool per = eng.mem_kind == SRLX_MEM_PROPORTIONAL;
  const bool presample = eng.presample != 0;
  const bool noisy = net.noisy != 0;
  const bool need_online_next = FLAG ? true : (eng.enable_double_dqn || M > 1);
  // data-parallel learner (dp_world ranks, one engine each): gradients summed over the ranks every update
  constexpr bool dp_on = DP;
  const int G = DP ? eng.dp_world : 1, dp_rank = DP ? eng.dp_rank : 0;
  const int PlPad = round_up(Pl, 4);
  void* const dp_own = dp_on ? eng.dp_peer[dp_rank] : nullptr;
  const int rows_sent_per_item = 1 + M + (need_online_next ? M : 0);
  const int nIt = max(0, min(ipc, B - rank * ipc));  // batch items this CTA owns: [rank*ipc, rank*ipc + nIt)
  srlx_state* st = eng.state;

  const uint64_t tc0 = st->train_count, mem_size = st->mem_size, vec_steps = st->vec_steps, adam0 = st->adam_step;
  const bool go = mem_size >= eng.warmup_size && mem_size >= (uint64_t)B;
  if (!go) return;  // still warming up (uniform across the cluster)

  const size_t nz_bytes = (size_t)3 * Pl * 4;
  auto nz_src = [&](uint32_t u) { return noise + ((size_t)u * C + rank) * 3 * Pl; };

  // ---- one-time setup ---------------------------------------------------------------------------------------------
  if (tid == 0) {
    mbar_init(&mbar[MB_RS], 1);
    mbar_init(&mbar[MB_AG], 1);
    mbar_init(&mbar[MB_S], 1);
    mbar_init(&mbar[MB_WT], 1);
    mbar_init(&mbar[MB_XR], 1);
    mbar_init(&mbar[MB_NZ0], 1);
    mbar_init(&mbar[MB_NZ1], 1);
    mbar_init(&mbar[MB_NZ2], 1);
    mbar_init(&mbar[MB_TQ], 1);
    fence_mbar_init();
    mbar_expect_tx(&mbar[MB_S], (uint32_t)B4 * 4);
    mbar_expect_tx(&mbar[MB_WT], (uint32_t)B4 * 4);
    if (rank == 0) mbar_expect_tx(&mbar[MB_TQ], (uint32_t)B * 8);
    if (noisy) {
      mbar_expect_tx(&mbar[MB_NZ0], (uint32_t)nz_bytes);
      bulk_g2s(nzr, nz_src(0), (uint32_t)nz_bytes, &mbar[MB_NZ0]);
      if (n_updates > 1) {
        mbar_expect_tx(&mbar[MB_NZ1], (uint32_t)nz_bytes);
        bulk_g2s(nzr + 3 * Pl, nz_src(1), (uint32_t)nz_bytes, &mbar[MB_NZ1]);
      }
    }
    const int ns = fast_segs(net, C, rank, segs);
    *n_seg_p = ns;
    *n_used_p = segs[ns - 1].l0 + segs[ns - 1].n;
    sc->loss_sum = 0.0;
    sc->last_loss = 0.0;
    sc->retries = 0;
    sc->sync_count = 0;
    sc->max_priority = st->max_priority;
    *dp_dead = 0;
  }
  for (int i = tid; i < 3 * WS; i += FNT) weff[i] = 0.f;
  for (int i = tid; i < 2 * NX * 4; i += FNT) xin[i] = 0.f;
  for (int i = tid; i < 8 * Pl; i += FNT) par[i] = 0.f;
  for (int i = tid; i < (int)n_updates && i < kFMaxChunk; i += FNT) {
    const double t = (double)(adam0 + (uint64_t)i + 1);
    adam_ss[i] = (float)(eng.lr / (1.0 - pow(eng.adam_beta1, t)));
    adam_bc[i] = (float)sqrt(1.0 - pow(eng.adam_beta2, t));
  }
  if (tid < M) gpow[tid] = (float)pow(eng.discount, (double)tid);
  if (rank == 0)
    for (int i = tid; i < n_cache; i += FNT) cache[i] = __ldcg(eng.tree + i);
  __syncthreads();
  const int n_seg = *n_seg_p, n_used = *n_used_p;
  for (int s = 0; s < n_seg; ++s) {
    const FSeg sg = segs[s];
    for (int j = tid; j < sg.n; j += FNT) {
      const int p = sg.g0 + j, i = sg.l0 + j;
      int sl;
      if (sg.kind == FSEG_W) {
        const int u = j / K, k = j - u * K;
        sl = (u >> 2) * 16 + k * 4 + (u & 3);
      } else if (sg.kind == FSEG_B) {
        sl = nCh * 16 + j;
      } else if (sg.kind == FSEG_O) {
        const int u = sg.ulo + j;
        sl = nCh * 20 + (u >> 2) * 16 + sg.o * 4 + (u & 3);
      } else {
        sl = nCh * 36 + j;
      }
      gidx[i] = p;
      slot_t[i] = sl | ((noisy && sg.noisy) ? (1 << 29) : 0) | (sg.replicated ? (1 << 30) : 0);
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

  const uint32_t mb_rs = smem_u32(&mbar[MB_RS]), mb_ag = smem_u32(&mbar[MB_AG]), mb_s = smem_u32(&mbar[MB_S]),
                 mb_wt = smem_u32(&mbar[MB_WT]);
  const unsigned FULL = 0xffffffffu;

  // ==================================================================================================================
  if (warp < kFCmpWarps) {
    // ================================================ COMPUTE WARPS ================================================
    const int ct = tid, cw = warp;
    const float b1 = (float)eng.adam_beta1, b2 = (float)eng.adam_beta2, aeps = (float)eng.adam_eps;
    const float gamma = (float)eng.discount, rh = (float)eng.retrace_h, delta = (float)eng.huber_delta;
    const bool ddqn = eng.enable_double_dqn != 0, rescale = eng.enable_rescale != 0;

    // Adam (update `upd`, if do_adam) + target sync + effective weights of the next update: one pass over the CTA's parameters
    auto finish = [&](bool do_adam, uint32_t upd, bool build_next) {
      const uint64_t tc_done = tc0 + upd;
      const float step_size = do_adam ? adam_ss[upd] : 0.f, bc2_sqrt = do_adam ? adam_bc[upd] : 1.f;
      const bool do_sync = do_adam && (tc_done % (uint64_t)eng.target_update_interval) == 0;
      const uint32_t un = do_adam ? upd + 1 : 0;  // update whose weights are built
      const float* z_cur = nzr + (size_t)(upd % 3) * 3 * Pl;   // noise of update `upd` (set 0 = the eps of online(s))
      const float* z_nxt = nzr + (size_t)(un % 3) * 3 * Pl;
      for (int i = ct; i < n_used; i += FNC) {
        const int sl = slot_t[i];
        const bool nz = (sl >> 29) & 1;
        const int slot = sl & 0x1fffffff;
        float mu = p_mu[i], sgm = nz ? p_sg[i] : 0.f;
        if (do_adam) {
          float g = 0.f;
          if (!dp_on) {
            for (int rg = 0; rg < nRG; ++rg) g += part[(size_t)rg * Pl + i];
          } else {
            // mean over the global batch: the ranks' sums added in rank order (identical bits on every rank), own one from smem,
            // the peers' from this rank's mailboxes as they land
            const int tpar = (int)(tc_done & 1);
            const uint32_t tag = (uint32_t)(tc_done + 1);
            // all peers' words polled together (one L2 round trip per attempt whatever the number of ranks)
            float gv[kDpMaxWorld];
            unsigned pend = ((1u << G) - 1u) & ~(1u << dp_rank);
            long long t0 = 0;
            for (int it = 0; pend; ++it) {
              unsigned long long w[kDpMaxWorld];
#pragma unroll
              for (int r = 0; r < kDpMaxWorld; ++r) {
                w[r] = 0;
                if ((pend >> r) & 1u)
                  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w[r]) : "l"(dp_grad(dp_own, tpar, r, rank, PlPad) + i) : "memory");
              }
#pragma unroll
              for (int r = 0; r < kDpMaxWorld; ++r)
                if (((pend >> r) & 1u) && (uint32_t)(w[r] >> 32) == tag) { gv[r] = __uint_as_float((uint32_t)w[r]); pend &= ~(1u << r); }
              if (pend) {
                if (*dp_dead) break;
                __nanosleep(100);
                if (it == 256) t0 = clock64();
                if (it > 256 && clock64() - t0 > 2000000000ll) { *dp_dead = 1; break; }
              }
            }
#pragma unroll
            for (int r = 0; r < kDpMaxWorld; ++r)
              if (r < G) g += (r == dp_rank) ? part[i] : ((pend >> r) & 1u ? 0.f : gv[r]);
            g *= 1.0f / (float)G;
          }
          float m = p_m1[i], v = p_v1[i];
          // torch/optim/adam.py _single_tensor_adam: lerp, mul_/addcmul_, sqrt/div/add_, addcdiv_
          m = m + (g - m) * (1.0f - b1);
          v = v * b2 + (1.0f - b2) * g * g;
          mu = mu - step_size * (m / (sqrtf(v) / bc2_sqrt + aeps));
          p_mu[i] = mu; p_m1[i] = m; p_v1[i] = v;
          const bool wr = eng.dbg_grads && (!((sl >> 30) & 1) || rank == 0);
          if (wr) eng.dbg_grads[gidx[i]] = g;
          if (noisy) {
            float gs = 0.f;
            if (nz) {
              gs = g * z_cur[i];
              float m2 = p_m2[i], v2 = p_v2[i];
              m2 = m2 + (gs - m2) * (1.0f - b1);
              v2 = v2 * b2 + (1.0f - b2) * gs * gs;
              sgm = sgm - step_size * (m2 / (sqrtf(v2) / bc2_sqrt + aeps));
              p_sg[i] = sgm; p_m2[i] = m2; p_v2[i] = v2;
            }
            if (wr) eng.dbg_grads[net.n_params + gidx[i]] = gs;
          }
          if (do_sync) { p_tmu[i] = mu; if (nz) p_tsg[i] = sgm; }
        }
        if (build_next) {
          const float tmu = p_tmu[i];
          float zS = 0.f, zN = 0.f, zT = 0.f, tsg = 0.f;
          if (nz) { zS = z_nxt[i]; zN = z_nxt[Pl + i]; zT = z_nxt[2 * Pl + i]; tsg = p_tsg[i]; }
          weff[slot] = fmaf(sgm, zS, mu);
          weff[WS + slot] = fmaf(sgm, zN, mu);
          weff[2 * WS + slot] = fmaf(tsg, zT, tmu);
        }
      }
    };

    // Only warp 0 of a group ever polls an mbarrier; the rest of the group blocks on a named barrier (hardware-blocked,
    // no polling), so waiting warps do not compete with working warps for the MIO queue.
    if (noisy && cw == 0) mbar_wait_sleep(&mbar[MB_NZ0], 0);
    named_bar_sync(FBAR_CMP, FNC);
    finish(false, 0, true);
    if (cw == 0) mbar_wait_sleep(&mbar[MB_XR], 0);  // x(0) gathered by the memory warps
    named_bar_sync(FBAR_CMP, FNC);

    const int nActive = min(kFCmpWarps, nCh);
    for (uint32_t upd = 0; upd < n_updates; ++upd) {
      const int parb = upd & 1;
      const float* x_cur = xin + (size_t)parb * NX * 4;
      const int* w_act = reinterpret_cast<const int*>(meta + (size_t)parb * 3 * BM);
      const float* w_rew = meta + (size_t)parb * 3 * BM + BM;
      const float* w_term = w_rew + BM;
      if (ct == 0) {
        mbar_expect_tx(&mbar[MB_RS], (uint32_t)(C * nIt * rows_sent_per_item * 16));
        mbar_expect_tx(&mbar[MB_AG], (uint32_t)(B * 32));
      }
      SRLX_FSTAMP(ct == 0, 0);

      // ---------------------------------------------------------------- forward: warp = 4-unit chunk(s), lane = row
      if (FLAG) {
        // compile-time shape: 1 tile of s rows + 3 tiles of s' rows for each of the two s' forwards; the warp's chunk(s)
        // accumulate into registers (no read-modify-write of the partial buffer), tiles are independent FMA chains.
        // A chunk lies entirely in the value half (output 0 only) or in the advantage half (outputs 1, 2) of the head.
        constexpr int NCHW = TC == 8 ? 2 : 1;
        float4 acc[7];
#pragma unroll
        for (int t = 0; t < 7; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int cp = 0; cp < NCHW; ++cp) {
          const int ch = cw + kFCmpWarps * cp;
          const bool vhalf = (u0 + ch * 4) < 512;
#pragma unroll
          for (int set = 0; set < 3; ++set) {
            const float* ws = weff + set * WS;
            const float4 W0 = ld4(ws + ch * 16), W1 = ld4(ws + ch * 16 + 4), W2 = ld4(ws + ch * 16 + 8), W3 = ld4(ws + ch * 16 + 12);
            const float4 Bv = ld4(ws + nCh * 16 + ch * 4);
            const float* wo = ws + nCh * 20 + ch * 16;
            const float4 Oa = ld4(wo + (vhalf ? 0 : 4)), Ob = ld4(wo + 8);  // value row, or the two advantage rows
#pragma unroll
            for (int tl = 0; tl < (set == 0 ? 1 : 3); ++tl) {
              const int t = set == 0 ? 0 : (set == 1 ? 1 + tl : 4 + tl);
              const float4 x = ld4(x_cur + (size_t)((set == 0 ? 0 : 32) + tl * 32 + lane) * 4);
              float2 a01 = f2(Bv.x, Bv.y), a23 = f2(Bv.z, Bv.w);
              a01 = __ffma2_rn(f2(W0.x, W0.y), f2(x.x, x.x), a01); a23 = __ffma2_rn(f2(W0.z, W0.w), f2(x.x, x.x), a23);
              a01 = __ffma2_rn(f2(W1.x, W1.y), f2(x.y, x.y), a01); a23 = __ffma2_rn(f2(W1.z, W1.w), f2(x.y, x.y), a23);
              a01 = __ffma2_rn(f2(W2.x, W2.y), f2(x.z, x.z), a01); a23 = __ffma2_rn(f2(W2.z, W2.w), f2(x.z, x.z), a23);
              a01 = __ffma2_rn(f2(W3.x, W3.y), f2(x.w, x.w), a01); a23 = __ffma2_rn(f2(W3.z, W3.w), f2(x.w, x.w), a23);
              const float2 h01 = f2(fmaxf(a01.x, 0.f), fmaxf(a01.y, 0.f)), h23 = f2(fmaxf(a23.x, 0.f), fmaxf(a23.y, 0.f));
              float2 ta = __ffma2_rn(h01, f2(Oa.x, Oa.y), f2(0.f, 0.f));
              ta = __ffma2_rn(h23, f2(Oa.z, Oa.w), ta);
              if (vhalf) {
                acc[t].x += ta.x + ta.y;
              } else {
                float2 tb = __ffma2_rn(h01, f2(Ob.x, Ob.y), f2(0.f, 0.f));
                tb = __ffma2_rn(h23, f2(Ob.z, Ob.w), tb);
                acc[t].y += ta.x + ta.y;
                acc[t].z += tb.x + tb.y;
              }
            }
          }
        }
        float* mypart = part + (size_t)cw * NRq * 4;
#pragma unroll
        for (int t = 0; t < 7; ++t) *reinterpret_cast<float4*>(mypart + (size_t)(t * 32 + lane) * 4) = acc[t];
      } else
      for (int ch = cw; ch < nCh; ch += kFCmpWarps) {
        const bool first = ch < kFCmpWarps;
        float* mypart = part + (size_t)cw * NRq * 4;
        unsigned omask = 0;
        {
          const int ua = u0 + ch * 4, ub = ua + 4;
          for (int o = 0; o < nout; ++o) {
            const int koff = (dueling != SRLX_DUEL_NONE && o > 0) ? net.k_dim[1] : 0;
            if (ua < koff + net.k_dim[1] && ub > koff) omask |= 1u << o;
          }
        }
        for (int set = 0; set < 3; ++set) {
          if (set == 1 && !need_online_next) continue;
          const float* ws = weff + set * WS;
          const float4 W0 = ld4(ws + ch * 16), W1 = ld4(ws + ch * 16 + 4), W2 = ld4(ws + ch * 16 + 8), W3 = ld4(ws + ch * 16 + 12);
          const float4 Bv = ld4(ws + nCh * 16 + ch * 4);
          const float* wo = ws + nCh * 20 + ch * 16;
          const float4 O0 = ld4(wo), O1 = ld4(wo + 4), O2 = ld4(wo + 8), O3 = ld4(wo + 12);
          const int nrows = set == 0 ? B : BM, row_base = set == 0 ? 0 : (set == 1 ? B : B + BM), xbase = set == 0 ? 0 : B;
          for (int j0 = 0; j0 < nrows; j0 += 32) {
            const int r = j0 + lane;
            const bool valid = r < nrows;
            const float4 x = ld4(x_cur + (size_t)(xbase + min(r, nrows - 1)) * 4);
            float2 a01 = f2(Bv.x, Bv.y), a23 = f2(Bv.z, Bv.w);
            a01 = __ffma2_rn(f2(W0.x, W0.y), f2(x.x, x.x), a01); a23 = __ffma2_rn(f2(W0.z, W0.w), f2(x.x, x.x), a23);
            a01 = __ffma2_rn(f2(W1.x, W1.y), f2(x.y, x.y), a01); a23 = __ffma2_rn(f2(W1.z, W1.w), f2(x.y, x.y), a23);
            a01 = __ffma2_rn(f2(W2.x, W2.y), f2(x.z, x.z), a01); a23 = __ffma2_rn(f2(W2.z, W2.w), f2(x.z, x.z), a23);
            a01 = __ffma2_rn(f2(W3.x, W3.y), f2(x.w, x.w), a01); a23 = __ffma2_rn(f2(W3.z, W3.w), f2(x.w, x.w), a23);
            const float2 h01 = f2(fmaxf(a01.x, 0.f), fmaxf(a01.y, 0.f)), h23 = f2(fmaxf(a23.x, 0.f), fmaxf(a23.y, 0.f));
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (omask & 1u) { float2 t = __ffma2_rn(h01, f2(O0.x, O0.y), f2(0.f, 0.f)); t = __ffma2_rn(h23, f2(O0.z, O0.w), t); acc.x = t.x + t.y; }
            if (omask & 2u) { float2 t = __ffma2_rn(h01, f2(O1.x, O1.y), f2(0.f, 0.f)); t = __ffma2_rn(h23, f2(O1.z, O1.w), t); acc.y = t.x + t.y; }
            if (omask & 4u) { float2 t = __ffma2_rn(h01, f2(O2.x, O2.y), f2(0.f, 0.f)); t = __ffma2_rn(h23, f2(O2.z, O2.w), t); acc.z = t.x + t.y; }
            if (omask & 8u) { float2 t = __ffma2_rn(h01, f2(O3.x, O3.y), f2(0.f, 0.f)); t = __ffma2_rn(h23, f2(O3.z, O3.w), t); acc.w = t.x + t.y; }
            if (valid) {
              float4* dst = reinterpret_cast<float4*>(mypart + (size_t)(row_base + r) * 4);
              if (!first) { const float4 o = *dst; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
              *dst = acc;
            }
          }
        }
      }
      named_bar_sync(FBAR_CMP, FNC);
      SRLX_FSTAMP(ct == 0, 1);
      // ---------------------------------------------------------------- reduce over warps, scatter to the owning CTA
      // two threads per row (each sums half of the warp copies), combined with one shuffle
      for (int wb = cw * 32; wb < 2 * NRq; wb += FNC) {  // whole warps iterate (shuffles below); NRq <= 288 -> one pass
        const int w2 = wb + lane;
        const int row = min(w2 >> 1, NRq - 1), half = w2 & 1;  // an (even, odd) lane pair shares a row
        int item, j;
        bool send = w2 < 2 * NRq;
        if (row < B) { item = row; j = 0; }
        else if (row < B + BM) {
          send = send && need_online_next;
          const int w = row - B; item = w / M; j = 1 + (w - item * M);
        } else { const int w = row - B - BM; item = w / M; j = 1 + M + (w - item * M); }
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int w = half; w < nActive; w += 2) {
          const float4 o = ld4(part + ((size_t)w * NRq + row) * 4);
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        v.x += __shfl_xor_sync(FULL, v.x, 1); v.y += __shfl_xor_sync(FULL, v.y, 1);
        v.z += __shfl_xor_sync(FULL, v.z, 1); v.w += __shfl_xor_sync(FULL, v.w, 1);
        if (half == 0 && send) {
          const int c = item / ipc, ii = item - c * ipc;
          const float* dst = rs + (((size_t)parb * C + rank) * nOwn + ii * rpi + j) * 4;
          st_async_f4(mapa_u32(smem_u32(dst), (uint32_t)c), v, mapa_u32(mb_rs, (uint32_t)c));
        }
      }
      if (cw == 0) {  // the owner work is one warp's worth: warp 0 does it, the other warps go straight to the AG barrier
        mbar_wait_sleep(&mbar[MB_RS], parb);
        SRLX_FSTAMP(ct == 0, 2);
        // -------------------------------------------------------------- owner: sum the C partials, dueling combine
        for (int w = lane; w < nIt * rpi; w += 32) {
          const int ii = w / rpi, j = w - ii * rpi;
          const int set = j == 0 ? 0 : (j <= M ? 1 : 2);
          if (set == 1 && !need_online_next) continue;
          float4 v = ld4(weff + set * WS + nCh * 36);  // effective output bias of this forward call
#pragma unroll 4
          for (int c = 0; c < C; ++c) {
            const float4 o = ld4(rs + (((size_t)parb * C + c) * nOwn + w) * 4);
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          }
          // (fixed-bound unrolled loops keep raw/q in registers: A <= 4, nout <= 4)
          const float raw[4] = {v.x, v.y, v.z, v.w};
          float q[4] = {0.f, 0.f, 0.f, 0.f};
          if (dueling == SRLX_DUEL_NONE) {
#pragma unroll
            for (int a = 0; a < 4; ++a) q[a] = raw[a];
          } else {
            float red = 0.f;
            if (dueling == SRLX_DUEL_AVERAGE) {
#pragma unroll
              for (int a = 0; a < 3; ++a)
                if (a < A) red += raw[1 + a];
              red = red / (float)A;
            } else if (dueling == SRLX_DUEL_MAX) {
              red = raw[1];
#pragma unroll
              for (int a = 1; a < 3; ++a)
                if (a < A) red = fmaxf(red, raw[1 + a]);
            }
#pragma unroll
            for (int a = 0; a < 3; ++a)
              if (a < A) q[a] = raw[0] + raw[1 + a] - red;
          }
          *reinterpret_cast<float4*>(qown + (size_t)w * 4) = make_float4(q[0], q[1], q[2], q[3]);
          if (set == 0) *reinterpret_cast<float4*>(rawown + (size_t)ii * 4) = v;
        }
        __syncwarp();
        SRLX_FSTAMP(ct == 0, 6);
        // -------------------------------------------------------------- owner: targets, Huber gradient (lane per item)
        bool wt_seen = false;
        for (int ii0 = 0; ii0 < nIt; ii0 += 32) {
          const int ii = ii0 + lane;
          float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
          if (ii < nIt) {
            const int i = rank * ipc + ii;
            const float* qs = qown + (size_t)(ii * rpi) * 4;
            const float* qon = qs + 4;             // online(s')  [M][4]
            const float* qtg = qs + 4 * (1 + M);   // target(s')  [M][4]
            float target = 0.f, retrace = 1.f;
#pragma unroll 1
            for (int k = 0; k < M; ++k) {
              const float* qo = qon + k * 4;
              const float* qt = qtg + k * 4;
              const float* qsel = ddqn ? qo : qt;
              int am = 0;
              float best = qsel[0];
              for (int a = 1; a < A; ++a)
                if (qsel[a] > best) { best = qsel[a]; am = a; }  // np.argmax: first max wins
              // Retrace with the reference's index shift (rainbow.py:267): action taken at s_k vs greedy action at s_{k+1}
              if (k >= 1) retrace = retrace * (rh * ((w_act[i * M + k] == am) ? 1.f : 0.f));
              float maxq = qt[am];
              if (rescale) maxq = inverse_rescaling_f(maxq);
              float gain = w_rew[i * M + k] + ((1.0f - w_term[i * M + k]) * gamma) * maxq;
              if (rescale) gain = rescaling_f(gain);
              float qk = 0.f;  // the first step is learnt by the trainer itself (rainbow.py:232-234)
              if (k >= 1) qk = qon[(k - 1) * 4 + w_act[i * M + k]];
              const float td = gain - qk;
              target += (td * gpow[k]) * retrace;
            }
            const int a0 = w_act[i * M + 0];
            const float q = qs[a0];
            // the SumTree chain (priority -> update -> next sample) needs only |target - q|: CTA 0 gets it ahead of the
            // all-gather, before the Huber gradient is even computed
            st_async_f2(mapa_u32(smem_u32(tqb + ((size_t)parb * B + i) * 2), 0u), target, q, mapa_u32(smem_u32(&mbar[MB_TQ]), 0u));
            if (!wt_seen) {
              mbar_wait_sleep(&mbar[MB_WT], parb);  // IS weights of this batch (sent after the slots)
              wt_seen = true;
            }
            const float w = samp_w[parb * B4 + i];
            const float d = q * w - target * w;
            const float ad = fabsf(d);
            const float lterm = (ad <= delta) ? 0.5f * d * d : delta * (ad - 0.5f * delta);
            const float dq = fminf(fmaxf(d, -delta), delta) * w / (float)B;
            float dr[4] = {0.f, 0.f, 0.f, 0.f};
            if (dueling == SRLX_DUEL_NONE) {
#pragma unroll
              for (int a = 0; a < 4; ++a) dr[a] = (a == a0) ? dq : 0.f;
            } else {  // dueling combine backward (dueling_network.py:51-58)
              int amax = 0;
              if (dueling == SRLX_DUEL_MAX) {
                const float* rw = rawown + ii * 4;
                float bestr = rw[1];
                for (int a = 1; a < A; ++a)
                  if (rw[1 + a] > bestr) { bestr = rw[1 + a]; amax = a; }
              }
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                if (a < A) {
                  float dd = (a == a0) ? dq : 0.f;
                  if (dueling == SRLX_DUEL_AVERAGE) dd -= dq / (float)A;
                  else if (dueling == SRLX_DUEL_MAX && a == amax) dd -= dq;
                  dr[1 + a] = dd;
                }
              }
              dr[0] = dq;
            }
            v0 = make_float4(dr[0], dr[1], dr[2], dr[3]);
            v1 = make_float4(target, q, lterm, 0.f);
          }
          // all-gather: lane l sends the result of item (l % n) to CTAs l / n, l / n + 32 / n, ... (n items this pass)
          const int n = min(32, nIt - ii0);
          if (n > 0) {
            const int src = lane % n;
            float4 s0, s1;
            s0.x = __shfl_sync(FULL, v0.x, src); s0.y = __shfl_sync(FULL, v0.y, src); s0.z = __shfl_sync(FULL, v0.z, src);
            s0.w = __shfl_sync(FULL, v0.w, src); s1.x = __shfl_sync(FULL, v1.x, src); s1.y = __shfl_sync(FULL, v1.y, src);
            s1.z = __shfl_sync(FULL, v1.z, src); s1.w = __shfl_sync(FULL, v1.w, src);
            const int i = rank * ipc + ii0 + src;
            const uint32_t dst = smem_u32(ag + ((size_t)parb * B + i) * 8);
            const int cstep = 32 / n;  // n <= 32
            if (lane < n * cstep) {
              for (int c = lane / n; c < C; c += cstep) {
                const uint32_t rb = mapa_u32(mb_ag, (uint32_t)c);
                st_async_f4(mapa_u32(dst, (uint32_t)c), s0, rb);
                st_async_f4(mapa_u32(dst + 16, (uint32_t)c), s1, rb);
              }
            }
          }
        }
        SRLX_FSTAMP(ct == 0, 8);
        mbar_wait_sleep(&mbar[MB_WT], parb);  // (a CTA without items still consumes the phase)
        if (ct == 0 && upd + 1 < n_updates) mbar_expect_tx(&mbar[MB_WT], (uint32_t)B4 * 4);
        mbar_wait_sleep(&mbar[MB_AG], parb);
      }
      named_bar_sync(FBAR_CMP, FNC);
      SRLX_FSTAMP(ct == 0, 3);
      // ---------------------------------------------------------------- backward of this CTA's slice: thread = (unit, row group)
      {
        const float* agb = ag + (size_t)parb * B * 8;
        const float* wS = weff;  // online(s)
        for (int it = ct; it < Us * nRG; it += FNC) {
          const int rg = it / Us, u = it - rg * Us;
          const float* wo = wS + nCh * 20 + (u >> 2) * 16 + (u & 3);
          const float wo0 = wo[0], wo1 = wo[4], wo2 = wo[8], wo3 = wo[12];
          // the hidden activation is recomputed (4 FMAs) instead of being staged through shared memory by the forward
          const float* wi = wS + (u >> 2) * 16 + (u & 3);
          const float wi0 = wi[0], wi1 = wi[4], wi2 = wi[8], wi3 = wi[12], bi = wS[nCh * 16 + u];
          float gW0 = 0.f, gW1 = 0.f, gW2 = 0.f, gW3 = 0.f, gb = 0.f, gO0 = 0.f, gO1 = 0.f, gO2 = 0.f, gO3 = 0.f;
          const int r0 = rg * rpg, r1 = min(B, r0 + rpg);
#pragma unroll 2
          for (int r = r0; r < r1; ++r) {
            const float4 dr = ld4(agb + (size_t)r * 8);
            const float4 x = ld4(x_cur + (size_t)r * 4);
            float h = fmaf(wi0, x.x, bi);  // same operation order as the forward (FFMA2 lanes are plain fp32 FMAs)
            h = fmaf(wi1, x.y, h); h = fmaf(wi2, x.z, h); h = fmaf(wi3, x.w, h);
            h = fmaxf(h, 0.f);
            float d = dr.x * wo0;
            d = fmaf(dr.y, wo1, d); d = fmaf(dr.z, wo2, d); d = fmaf(dr.w, wo3, d);
            const float dh = h > 0.f ? d : 0.f;
            gO0 = fmaf(dr.x, h, gO0); gO1 = fmaf(dr.y, h, gO1); gO2 = fmaf(dr.z, h, gO2); gO3 = fmaf(dr.w, h, gO3);
            gb += dh;
            gW0 = fmaf(dh, x.x, gW0); gW1 = fmaf(dh, x.y, gW1); gW2 = fmaf(dh, x.z, gW2); gW3 = fmaf(dh, x.w, gW3);
          }
          float* g = part + (size_t)rg * Pl;
          const float gW[4] = {gW0, gW1, gW2, gW3};
          if (K == 4) *reinterpret_cast<float4*>(g + u * 4) = make_float4(gW0, gW1, gW2, gW3);
          else
            for (int k = 0; k < K; ++k) g[u * K + k] = gW[k];
          g[Us * K + u] = gb;
          for (int s = 2; s < n_seg; ++s) {
            const FSeg& sg = segs[s];
            if (sg.kind == FSEG_O && u >= sg.ulo && u < sg.ulo + sg.n)
              g[sg.l0 + (u - sg.ulo)] = sg.o == 0 ? gO0 : (sg.o == 1 ? gO1 : (sg.o == 2 ? gO2 : gO3));
          }
        }
        // output bias (replicated parameter, identical on every CTA): row group 0 holds the sum, the others zero
        if (ct < nout * nRG) {
          const int rg = ct / nout, o = ct - rg * nout;
          float a = 0.f;
          if (rg == 0)
            for (int r = 0; r < B; ++r) a += agb[(size_t)r * 8 + o];
          part[(size_t)rg * Pl + segs[n_seg - 1].l0 + o] = a;
        }
      }
      const bool more = upd + 1 < n_updates;
      if (noisy && more && cw == 0) mbar_wait_sleep(&mbar[MB_NZ0 + (upd + 1) % 3], ((upd + 1) / 3) & 1);
      named_bar_sync(FBAR_CMP, FNC);
      SRLX_FSTAMP(ct == 0, 4);
      if (dp_on) {
        // ---------------------------------------------------------------- gradient all-reduce over the ranks (NVLink peer stores):
        // this CTA's slice, summed over its row groups, goes word by word (value + update tag) into every peer's mailbox
        // [parity][this rank][this CTA]; the Adam pass below picks the peers' words up as they land
        const uint32_t tag = (uint32_t)(tc0 + upd + 1);
        const int tpar = (int)((tc0 + upd) & 1);
        for (int i = ct; i < n_used; i += FNC) {
          float g = 0.f;
          for (int rg = 0; rg < nRG; ++rg) g += part[(size_t)rg * Pl + i];
          part[i] = g;  // row group 0's slot: only this thread touches index i, here and in the Adam pass
          for (int r = 0; r < G; ++r)
            if (r != dp_rank) dp_store(dp_grad(eng.dp_peer[r], tpar, dp_rank, rank, PlPad) + i, __float_as_uint(g), tag);
        }
      }
      // ---------------------------------------------------------------- Adam + target sync + next effective weights
      finish(true, upd, more);
      SRLX_FSTAMP(ct == 0, 5);
      if (more && cw == 0) mbar_wait_sleep(&mbar[MB_XR], parb ^ 1);  // x(t+1) gathered by the memory warps
      named_bar_sync(FBAR_CMP, FNC);
    }
  } else {
    // ================================================= MEMORY WARPS =================================================
    // warp 0 of every CTA: gather of x(t) as soon as the slots arrive, TMA prefetch of the noise ring.
    // CTA 0, all four warps: the replay memory -- IS weights, the SumTree update of batch t, the PER sample of batch t+1.
    const int mt = tid - FNC, mw = warp - kFCmpWarps;
    if (rank != 0 && mw != 0) goto done_roles;  // nothing to do for memory warps 1..3 outside CTA 0
    {
      const uint32_t glR = (uint32_t)((vec_steps - 1) % (uint64_t)R);  // ring row of the last vector step
      const int dmax = 31 - __clz(n_nodes);                            // depth of the deepest leaf
      // ---- sampler state (CTA 0): lane g < 8 of warp mw owns sample mw*8+g ------------------------------------------
      const int own_i = mw * 8 + lane;
      const bool own = lane < 8 && own_i < B;
      double u_next = 0.0;  // warp 0, lane = sample: pre-drawn uniform of the next batch
      auto draw = [&](uint64_t tc, int i, int k) -> double {
        const uint4 w = philox_ni(eng.seed, STREAM_SAMPLE, (uint32_t)i | ((uint32_t)k << 16), (uint32_t)tc, (uint32_t)(tc >> 32));
        return u01_f64(w.x, w.y);
      };
      // ---- update plan of the current batch (per warp: its tree levels mw, mw+4, ...), filled before the targets arrive
      int s_item = 0, s_li = 0x7fffffff;  // sorted lane r: item, its tree index (sorted by root-to-leaf path)
      bool s_valid = false;
      // levels are dealt to memory warps 1..3 (warp 0 keeps the gather / IS weights / polling): slot q < kLc is level
      // (mw-1) + 3q of the shared-memory-cached top of the tree, slot kLc + q is level clev + (mw-1) + 3q below it
      constexpr int kLc = 4, kLu = 5, kLv = kLc + kLu;  // 12 cached levels / 3 warps; up to 15 deeper levels / 3 warps
      const int clev = 31 - __clz(n_cache + 1);         // number of cached levels
      const BlkPlan bp = make_blk_plan(n_nodes, clev);
      const bool use_blk = per && eng.tree_blk != nullptr && bp.n_tiers > 0 && eng.tree_blk_bytes >= (uint64_t)bp.total * 512;
      int s_leaf_b = -1;  // sorted lane: element of the item's leaf inside the blocked copy (-1: none)
      // per-thread plan, parked in shared memory ([slot][thread], conflict-free) so the level loops stay rolled:
      int* p_node = plan_node + mt;   // [kLv][FNM] node a leader lane writes at slot q, -1 otherwise
      int* p_end = plan_end + mt;     // [kLv][FNM] last lane of the leader's run
      int* p_baddr = plan_baddr + mt; // [kLu][FNM] where the uncached node lives in the blocked copy
      double* p_old = plan_old + mt;  // [kLu][FNM] old values of the uncached nodes, fetched while the forward pass runs
      double* p_oldleaf = plan_oldleaf + mt;
      auto level_of = [&](int q) -> int { return q < kLc ? (mw - 1) + 3 * q : clev + (mw - 1) + 3 * (q - kLc); };
      auto level_ok = [&](int q, int a) -> bool { return q < kLc ? (a < clev && a < dmax) : (a < dmax); };

      // PER sample of update `tc` (CTA 0, all memory threads): tree walk -> s_idx / s_pri -> slots to every CTA
#ifdef SRLX_STAMPS
#define SRLX_SSTAMP(slot)                                                                                  \
  do {                                                                                                     \
    if (eng.dbg_clock && mt == ((slot) >= 40 ? 32 : 0) && stamp) eng.dbg_clock[slot] = clock64();           \
  } while (0)
#else
#define SRLX_SSTAMP(slot) do { (void)stamp; } while (0)
#endif
      // Walk of the shared-memory-cached top levels for ALL samples of update `tc` (warp 0, lane = sample; branch-free,
      // the cached top is a complete binary tree): leaves (node, remaining value) in s_widx / s_wval.
      auto walk_top = [&](double u) {
        const double total = cache[0];
        int idx = 0;
        double val = u * total;
#pragma unroll 1
        for (int l = 1; l < clev; ++l) {
          const double tl = cache[2 * idx + 1];
          const bool right = !(val <= tl);
          const double vr = val - tl;
          val = right ? vr : val;
          idx = 2 * idx + 1 + (right ? 1 : 0);
        }
        if (lane < B) { s_widx[lane] = idx; s_wval[lane] = val; }
      };
      auto sample_slots = [&](uint64_t tc, int pb, bool stamp) {
        if (per) {
          const double total = cache[0];
          if (mt == 0) *s_tot = total;  // the IS weights of this batch use the total it was drawn under
          // (a) the cached top levels were walked by warp 0 for all B samples (walk_top) while warps 1..3 finished the
          //     deep levels of the update; each owner lane picks its sample's state up from shared memory
          int idx = 0;
          double val = 0.0, pcur = 0.0;
          if (own) { idx = s_widx[own_i]; val = s_wval[own_i]; pcur = cache[idx]; }
          SRLX_SSTAMP(22);
          bool done = !own || (2 * idx + 1 >= n_nodes);
          bool pc_ok = true;
          // (b) the remaining levels, five per L2 round trip: 31 lanes fetch both children of every node of the 5-level
          //     subtree below each of the warp's 8 samples, the owner lanes replay the "val <= tree[left]" walk out of smem
          double* wsub = sub + (size_t)mw * 8 * kFSubLd;
          const int k_l = 32 - __clz(lane + 1);         // level (1..5) whose nodes this lane fetches; lane 31 idles
          const int q_l = lane + 1 - (1 << (k_l - 1));  // parent position inside level k_l - 1
          const int pos_l = (1 << k_l) - 2 + 2 * q_l;
          const unsigned c_l = (1u << k_l) - 1u + 2u * (unsigned)q_l;
          const unsigned last_pair = n_nodes > 1 ? (unsigned)n_nodes - 2u : 0u;
          int rnd = 0;
          while (__any_sync(FULL, !done)) {
            double v0[8], v1[8];
            if (use_blk) {
              // blocked copy: the 5-level subtree below each sample's node is one 512-byte block; lane l < 31 fetches
              // both children of the subtree's l-th node with one 16-byte load (finished samples read block 0, unused)
              const int tfirst = rnd == 0 ? bp.first[0] : (rnd == 1 ? bp.first[1] : (rnd == 2 ? bp.first[2] : bp.first[3]));
              const int toff = rnd == 0 ? bp.off[0] : (rnd == 1 ? bp.off[1] : (rnd == 2 ? bp.off[2] : bp.off[3]));
              const double2* bbase = reinterpret_cast<const double2*>(eng.tree_blk) + (lane < 31 ? lane : 0);
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const int ig = __shfl_sync(FULL, idx, g);
                const int dg = __shfl_sync(FULL, (int)done, g);
                const int b = dg ? 0 : toff + (ig - tfirst);
                const double2 c2 = __ldcg(bbase + (size_t)b * 32);
                v0[g] = c2.x;
                v1[g] = c2.y;
              }
            } else {
            // unconditional loads from clamped addresses (a predicated load + select makes ptxas wait for every load in
            // turn): nodes past the end of the tree or below finished samples are fetched but never looked at.  Left
            // children drive the walk; a right child's value is only ever needed as the priority of the leaf the walk
            // ends on, so only the lanes of the deepest fetched level load it (fewer L1 wavefronts on the critical chain).
            unsigned nodes[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const unsigned ig = (unsigned)__shfl_sync(FULL, idx, g);
              unsigned node = (ig << k_l) + c_l;  // < 2^32: fast_shape_ok bounds the tree at 2^27 nodes
              nodes[g] = node < last_pair ? node : last_pair;
              v0[g] = __ldcg(eng.tree + nodes[g]);
              v1[g] = 0.0;
            }
            if (k_l == 5) {
#pragma unroll
              for (int g = 0; g < 8; ++g) v1[g] = __ldcg(eng.tree + nodes[g] + (n_nodes > 1 ? 1 : 0));
            }
            }
            SRLX_SSTAMP(32 + rnd * 3);
            if (lane < 31) {
#pragma unroll
              for (int g = 0; g < 8; ++g) *reinterpret_cast<double2*>(wsub + g * kFSubLd + pos_l) = make_double2(v0[g], v1[g]);
            }
            __syncwarp();
            SRLX_SSTAMP(33 + rnd * 3);
            if (!done) {
              const double* ms = wsub + lane * kFSubLd;
              int rel = 0;
#pragma unroll 1
              for (int k = 1; k <= 5; ++k) {  // branch-free; a lane that reached its leaf keeps its state
                const int left = 2 * idx + 1;
                const bool act = left < n_nodes;
                const int base = (1 << k) - 2 + 2 * rel;
                const double2 ch = *reinterpret_cast<const double2*>(ms + base);
                const bool right = !(val <= ch.x);
                const double vr = val - ch.x;
                val = (act && right) ? vr : val;
                pcur = act ? (right ? ch.y : ch.x) : pcur;
                pc_ok = act ? (!right || k == 5 || use_blk) : pc_ok;
                idx = act ? left + (right ? 1 : 0) : idx;
                rel = act ? 2 * rel + (right ? 1 : 0) : 0;
              }
              done = 2 * idx + 1 >= n_nodes;
            }
            __syncwarp();
            SRLX_SSTAMP(34 + rnd * 3);
            ++rnd;
          }
          SRLX_SSTAMP(23);
          int s_li_final = cap1;
          if (own) {  // a zero-priority leaf is re-drawn (proportional_memory.py:150-152), sequentially
            int li = idx;
            if (!pc_ok) pcur = __ldcg(eng.tree + idx);  // ragged tree: the walk ended on a right child above the fetched bottom
            double p = pcur;
            int k = 0;
            while (p == 0.0 && k + 1 < 9999) {
              ++k;
              li = (int)tree_retrieve_seq(eng.tree, n_nodes, draw(tc, own_i, k) * total);
              p = __ldcg(eng.tree + li);
            }
            s_idx[own_i] = li;
            s_pri[own_i] = p;
            s_att[own_i] = k;
            s_li_final = li;
            if (k) atomicAdd(&sc->retries, (unsigned long long)k);
          }
          {
            // The ring rows of the sampled windows are read by every CTA ~1000 cycles from now: start their DRAM fetch
            // (L2 prefetch).  Lane = (sample lane & 7, part lane >> 3): part 0 the observation, parts 1.. one window step
            // each (next observation, action, reward, flags) -- the idle lanes share the address arithmetic.
            const int g = lane & 7, part_j = lane >> 3;
            const int lg = __shfl_sync(FULL, s_li_final, g);
            if (mw * 8 + g < B && part_j <= M) {
              const int s0 = lg - cap1, rho = s0 / E, e = s0 - rho * E;
              if (part_j == 0) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(eng.ring_obs + (size_t)s0 * D));
              } else {
                const int sk = ((rho + part_j - 1) % R) * E + e;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(eng.ring_next_obs + (size_t)sk * D));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(eng.ring_action + sk));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(eng.ring_reward + sk));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(eng.ring_term + sk));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(eng.ring_done + sk));
              }
            }
          }
          named_bar_sync(FBAR_MEM, FNM);
          if (!eng.has_duplicate) {
            if (mt == 0) {
              for (int i2 = 1; i2 < B; ++i2) {
                int k = s_att[i2];
                while (k < 9999) {
                  bool dup = false;
                  for (int j = 0; j < i2; ++j) dup |= (s_idx[j] == s_idx[i2]);
                  if (!dup && s_pri[i2] != 0.0) break;
                  ++k;
                  sc->retries += 1;
                  if (k >= 9999) break;
                  s_idx[i2] = (int)tree_retrieve_seq(eng.tree, n_nodes, draw(tc, i2, k) * total);
                  s_pri[i2] = __ldcg(eng.tree + s_idx[i2]);
                }
              }
            }
            named_bar_sync(FBAR_MEM, FNM);
          }
        } else {
          // uniform replay: B distinct items (replay_buffer.py:34-36).  Attempt 0 of every draw in parallel (B <= 32: one lane
          // each); a pick that repeats an earlier one is redrawn (attempt k = 1, 2, ...) in sample order, which is exactly what
          // the sequential rejection loop (oracle/sumtree.py::uniform_sample_distinct) produces
          if (mt < 32) {
            const uint64_t g_next = vec_steps;
            const uint64_t g_lo = g_next > (uint64_t)R ? g_next - R : 0;
            const uint32_t n_g = (uint32_t)(g_next - (uint64_t)(M - 1) - g_lo);
            const uint32_t n_valid = n_g * (uint32_t)E;
            const uint32_t g_lo_mod = (uint32_t)(g_lo % (uint64_t)R);
            int pick = -1 - mt;  // idle lanes hold distinct negative values
            if (mt < B) {
              const uint4 w = philox_ni(eng.seed, STREAM_UNIFORM_SAMPLE, (uint32_t)mt, (uint32_t)tc, (uint32_t)(tc >> 32));
              pick = (int)u_below(w.x, n_valid);
            }
            bool dup = false;
#pragma unroll 1
            for (int j = 0; j < B; ++j) {
              const int pj = __shfl_sync(FULL, pick, j);
              dup |= (j < mt) & (pj == pick);
            }
            if (__any_sync(FULL, dup)) {  // rare (B^2 / (2 n_valid)): resolve in sample order on one lane
              if (mt < B) s_idx[mt] = pick;
              __syncwarp();
              if (mt == 0) {
                for (int i = 1; i < B; ++i) {
                  int k = 0;
                  while (true) {
                    bool d2 = false;
                    for (int j = 0; j < i; ++j) d2 |= (s_idx[j] == s_idx[i]);
                    if (!d2 || ++k >= 65536) break;
                    const uint4 w = philox_ni(eng.seed, STREAM_UNIFORM_SAMPLE, (uint32_t)i | ((uint32_t)k << 16), (uint32_t)tc, (uint32_t)(tc >> 32));
                    s_idx[i] = (int)u_below(w.x, n_valid);
                  }
                }
              }
              __syncwarp();
              if (mt < B) pick = s_idx[mt];
            }
            if (mt < B) {
              const uint32_t pk = (uint32_t)pick, q = pk / (uint32_t)E;
              s_idx[mt] = (int)(((g_lo_mod + q) % (uint32_t)R) * (uint32_t)E + (pk - q * (uint32_t)E)) + cap1;  // stored as if it were a tree index
            }
          }
          named_bar_sync(FBAR_MEM, FNM);
        }
        SRLX_SSTAMP(24);
        // slots to every CTA (4 per store)
        {
          const int nq = B4 / 4;
          const uint32_t dst = smem_u32(samp_slot + pb * B4);
          for (int w = mt; w < nq * C; w += FNM) {
            const int c = w / nq, q = w - c * nq;
            int v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (4 * q + e < B) ? (s_idx[4 * q + e] - cap1) : 0;
            st_async_i4(mapa_u32(dst + 16 * q, (uint32_t)c), make_int4(v[0], v[1], v[2], v[3]), mapa_u32(mb_s, (uint32_t)c));
          }
        }
        SRLX_SSTAMP(25);
        if (dp_on && per && mt < 32) {
          // data-parallel learner: this shard's replay scalars of batch `tc` go to the peers NOW (the IS weights that need the
          // peers' scalars are formed only after the gather, one NVLink latency from here)
          double pmin = mt < B ? s_pri[mt] : 1e300;
          for (int s = 16; s > 0; s >>= 1) pmin = fmin(pmin, __shfl_xor_sync(FULL, pmin, s));
          if (mt < G && mt != dp_rank) {
            const double v[4] = {cache[0], (double)mem_size, pmin, sc->max_priority};
            unsigned long long* dst = dp_wsc(eng.dp_peer[mt], (int)(tc & 1), dp_rank);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const unsigned long long bits = (unsigned long long)__double_as_longlong(v[k]);
              dp_store(dst + 2 * k, (uint32_t)bits, (uint32_t)(tc + 1));
              dp_store(dst + 2 * k + 1, (uint32_t)(bits >> 32), (uint32_t)(tc + 1));
            }
          }
        }
        if (eng.dbg_sample_idx)
          for (int i = mt; i < B; i += FNM) eng.dbg_sample_idx[i] = per ? (int64_t)s_idx[i] : (int64_t)(s_idx[i] - cap1);
      };

      // IS weights of the batch sampled for update `tc` (proportional_memory.py:159-167), one warp; only the Huber step
      // needs them, so they are computed after the slots have gone out
      auto send_weights = [&](uint64_t tc, int pb) {
        double wv = 1.0;
        if (per) {
          const double total = *s_tot;  // the total the batch was drawn under (== cache[0] unless the batch was pre-sampled)
          // PriorityReplayBuffer.step is the train_count of the PREVIOUS update() call (priority_replay_buffer.py:232,250)
          const double stepd = (tc > 0) ? (double)(tc - 1) : 0.0;
          double beta = eng.per_beta_initial + (1.0 - eng.per_beta_initial) * stepd / eng.per_beta_steps;
          beta = beta > 1.0 ? 1.0 : beta;
          if (!dp_on) {
            const double w = lane < B ? pow_chain((double)mem_size * (s_pri[lane] / total), -beta) : 0.0;
            double mx = w;
            for (int s = 16; s > 0; s >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, s));
            wv = w / mx;
          } else {
            // one memory sharded over the ranks: N, total and the weight maximum are those of ALL shards
            // (proportional_memory.py:138-167).  w = (N p / total)^-beta falls with p, so the global maximum is the weight
            // of the smallest priority any rank sampled: each rank publishes {total, size, min p of its batch, max_priority}
            double pmin = lane < B ? s_pri[lane] : 1e300;
            for (int s = 16; s > 0; s >>= 1) pmin = fmin(pmin, __shfl_xor_sync(FULL, pmin, s));
            const uint32_t tag = (uint32_t)(tc + 1);
            const int tpar = (int)(tc & 1);
            double v[4] = {total, (double)mem_size, pmin, sc->max_priority};  // (what sample_slots sent to the peers)
            if (lane < G && lane != dp_rank) {
              const unsigned long long* src = dp_wsc(dp_own, tpar, lane);
              unsigned long long w[8];
              long long t0 = 0;
              for (int it = 0;; ++it) {  // the eight words polled together: one L2 round trip per attempt
                bool ok = true;
#pragma unroll
                for (int k = 0; k < 8; ++k) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w[k]) : "l"(src + k) : "memory");
#pragma unroll
                for (int k = 0; k < 8; ++k) ok = ok && ((uint32_t)(w[k] >> 32) == tag);
                if (ok || *dp_dead) break;
                if (it == 64) t0 = clock64();
                if (it > 64 && clock64() - t0 > 2000000000ll) { *dp_dead = 1; break; }
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) v[k] = __longlong_as_double((long long)((w[2 * k + 1] << 32) | (w[2 * k] & 0xffffffffull)));
            }
            const double v_tot = v[0], v_n = v[1], v_pm = v[2], v_mx = v[3];
            double g_tot = 0.0, g_n = 0.0, g_pm = 1e300, g_mx = 0.0;
            for (int r = 0; r < G; ++r) {  // rank order: the same sums on every rank
              g_tot += __shfl_sync(FULL, v_tot, r);
              g_n += __shfl_sync(FULL, v_n, r);
              g_pm = fmin(g_pm, __shfl_sync(FULL, v_pm, r));
              g_mx = fmax(g_mx, __shfl_sync(FULL, v_mx, r));
            }
            if (lane == 0 && g_mx > sc->max_priority) sc->max_priority = g_mx;  // one memory: one max_priority
            const double w = lane < B ? pow_chain(g_n * (s_pri[lane] / g_tot), -beta) : 0.0;
            const double mx = pow_chain(g_n * (g_pm / g_tot), -beta);
            wv = w / mx;
          }
        }
        if (lane < B) s_tmp[lane] = wv;
        __syncwarp();
        const int nq = B4 / 4;
        const uint32_t dst = smem_u32(samp_w + pb * B4);
        for (int w = lane; w < nq * C; w += 32) {
          const int c = w / nq, q = w - c * nq;
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = (4 * q + e < B) ? (float)s_tmp[4 * q + e] : 0.f;
          st_async_f4(mapa_u32(dst + 16 * q, (uint32_t)c), make_float4(v[0], v[1], v[2], v[3]), mapa_u32(mb_wt, (uint32_t)c));
          if (c == 0 && eng.dbg_weights) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (4 * q + e < B) eng.dbg_weights[4 * q + e] = v[e];
          }
        }
        __syncwarp();
      };

      // Plan of the SumTree update of the current batch (every warp for its own levels): items sorted by root-to-leaf
      // path, so the items below any node are consecutive lanes; per level the first lane of a run ("leader") knows the
      // node, the last lane of its run, and has the node's old value in a register before the new priorities exist.
      auto plan_update = [&](bool stamp) {
        SRLX_SSTAMP(46);
        const bool v = lane < B;
        const int li = v ? s_idx[lane] : 0x7ffffffe;
        const unsigned ip1 = (unsigned)li + 1u;
        const int d = 31 - __clz(ip1);
        const unsigned key = v ? (ip1 << (31 - d)) : 0xffffffffu;
        int rank_l = 0;
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
          const unsigned kj = __shfl_sync(FULL, key, j);
          rank_l += (int)(kj < key) + ((int)(kj == key) & (int)(j < lane));  // bitwise: no divergent short-circuit branches
        }
        SRLX_SSTAMP(47);
        int* pw = sperm + mw * 32;
        pw[rank_l] = lane;
        __syncwarp();
        s_item = pw[lane];
        __syncwarp();
        s_valid = s_item < B;
        s_li = s_valid ? s_idx[s_item] : 0x7ffffffe;
        // pre-sampled batches were drawn before the previous update: their leaves may have changed since, read them now
        *p_oldleaf = s_valid ? (presample ? __ldcg(eng.tree + s_li) : s_pri[s_item]) : 0.0;
        const unsigned sip1 = (unsigned)s_li + 1u;
        const int sd = 31 - __clz(sip1);
#pragma unroll 1
        for (int q = 0; q < kLv; ++q) {
          const int a = level_of(q);
          int pn = -1, pe = lane;
          if (level_ok(q, a)) {
            const bool has = (int)s_valid & (int)(sd > a);
            const int node = has ? (int)(sip1 >> (sd - a)) - 1 : -1;
            const int prevn = __shfl_up_sync(FULL, node, 1), nextn = __shfl_down_sync(FULL, node, 1);
            const bool lead = (int)has & ((int)(lane == 0) | (int)(prevn != node));
            const bool last = (int)has & ((int)(lane == 31) | (int)(nextn != node));
            const unsigned bl = __ballot_sync(FULL, last);
            pe = (lane + __ffs(bl >> lane) - 1) & 31;  // first "last" flag at or after this lane (leaders find one)
            pn = lead ? node : -1;
          }
          p_node[q * FNM] = pn;
          p_end[q * FNM] = pe;
        }
        SRLX_SSTAMP(48);
        // uncached levels: unconditional loads (node 0 for non-leaders), all in flight at once, then parked in shared memory
        double ov[kLu];
#pragma unroll
        for (int q = 0; q < kLu; ++q) {
          const int pn = p_node[(kLc + q) * FNM];
          ov[q] = __ldcg(eng.tree + (pn >= 0 ? pn : 0));
        }
        if (use_blk) {
#pragma unroll 1
          for (int q = 0; q < kLu; ++q) {
            const int pn = p_node[(kLc + q) * FNM];
            p_baddr[q * FNM] = pn >= 0 ? blk_addr(bp, clev, pn) : -1;
          }
          s_leaf_b = (s_valid && sd >= clev) ? blk_addr(bp, clev, s_li) : -1;
        }
#pragma unroll
        for (int q = 0; q < kLu; ++q) p_old[q * FNM] = ov[q];
        SRLX_SSTAMP(49);
      };

      // ProportionalMemory.update of the current batch (proportional_memory.py:171-177).  The new priorities, their
      // changes and the running sum are computed by every warp (no cross-warp hand-off); each warp then writes the nodes
      // of its levels: new = old + (sum of the changes of the items below the node).  The reference adds the changes one
      // item at a time in batch order; here a node's changes are summed in path order as a difference of running sums
      // (exact for the common single-item run) -- a different association of the same fp64 terms, see DESIGN.md.
      auto apply_update = [&](const float* tq, bool stamp) {
        double pnew = 0.0;
        if (s_valid) pnew = pow_chain(fabs((double)fabsf(tq[s_item * 2] - tq[s_item * 2 + 1])) + eng.per_epsilon, eng.per_alpha);
        SRLX_SSTAMP(40);
        const int prevli = __shfl_up_sync(FULL, s_li, 1), nextli = __shfl_down_sync(FULL, s_li, 1);
        const double prevp = __shfl_up_sync(FULL, pnew, 1);
        const bool dupprev = lane > 0 && prevli == s_li;  // duplicates of a leaf are consecutive, in batch order
        const double chg = s_valid ? pnew - (dupprev ? prevp : *p_oldleaf) : 0.0;
        if (mw == 1) {
          if (s_valid && (lane == 31 || nextli != s_li)) {  // the last item touching a leaf wins
            __stcg(eng.tree + s_li, pnew);
            if (s_li < n_cache) cache[s_li] = pnew;
            if (use_blk && s_leaf_b >= 0) __stcg(eng.tree_blk + s_leaf_b, pnew);
          }
        }
        SRLX_SSTAMP(41);
        double P = chg;  // inclusive running sum over the sorted lanes
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
          const double t = __shfl_up_sync(FULL, P, s);
          if (lane >= s) P += t;
        }
        double Pex = __shfl_up_sync(FULL, P, 1);
        if (lane == 0) Pex = 0.0;
        SRLX_SSTAMP(42);
        // cached levels first: the sampler's top-level walk (warp 0) needs only these
        auto level_slot = [&](int q) {
          const int pn = p_node[q * FNM], e = p_end[q * FNM];
          const double Pe = __shfl_sync(FULL, P, e);
          if (pn >= 0) {
            const double sum = (e == lane) ? chg : (Pe - Pex);
            const double old = q < kLc ? cache[pn] : p_old[(q < kLc ? 0 : q - kLc) * FNM];
            const double nv = old + sum;
            __stcg(eng.tree + pn, nv);
            if (q < kLc) cache[pn] = nv;
            else if (use_blk) __stcg(eng.tree_blk + p_baddr[(q - kLc) * FNM], nv);
          }
        };
#pragma unroll 2
        for (int q = 0; q < kLc; ++q) level_slot(q);
        named_bar_arrive(FBAR_MB, FNM);  // the cached top is final: warp 0 starts the next batch's top-level walk
#pragma unroll 1
        for (int q = kLc; q < kLv; ++q) level_slot(q);
        SRLX_SSTAMP(43);
      };

      // loss / counters / debug taps of update `u` (off the critical path)
      auto bookkeeping = [&](uint32_t u) {
        const float* agb = ag + (size_t)(u & 1) * B * 8;
        if (lane == 0) {
          float l = 0.f;
          for (int i = 0; i < B; ++i) l += agb[i * 8 + 6];
          const double ld = (double)l / (double)B;
          sc->last_loss = ld;
          sc->loss_sum += ld;
          if (((tc0 + u) % (uint64_t)eng.target_update_interval) == 0) sc->sync_count += 1;
        }
        if (eng.dbg_target_q)
          for (int i = lane; i < B; i += 32) eng.dbg_target_q[i] = agb[i * 8 + 4];
        if (eng.dbg_q_sa)
          for (int i = lane; i < B; i += 32) eng.dbg_q_sa[i] = agb[i * 8 + 5];
      };

      if (rank == 0) {
        if (per) {
          if (mw == 0) walk_top(lane < B ? draw(tc0, lane, 0) : 0.0);
          named_bar_sync(FBAR_MEM, FNM);
        }
        sample_slots(tc0, 0, false);
      }

      for (uint32_t upd = 0; upd < n_updates; ++upd) {
        const uint64_t tc = tc0 + upd;
        const int parb = upd & 1;
        float* x_cur = xin + (size_t)parb * NX * 4;
        int* w_act = reinterpret_cast<int*>(meta + (size_t)parb * 3 * BM);
        float* w_rew = meta + (size_t)parb * 3 * BM + BM;
        float* w_term = w_rew + BM;
        const int* slot = samp_slot + parb * B4;
        if (mw == 0) {
          mbar_wait_sleep(&mbar[MB_S], parb);  // slots of update t have arrived from CTA 0
          SRLX_FSTAMP(mt == 0, 16);
          // -------------------------------------------------------------- gather the windows: lane = batch item, every load
          // of the window in flight at once (one memory round trip), the padding rule applied in registers
          const bool v = lane < B;
          const int s0 = v ? slot[lane] : 0;
          const int rho = s0 / E, e = s0 - rho * E;
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          auto ldx = [&](const float* base, int sk) -> float4 {
            if (D == 4) return __ldcg(reinterpret_cast<const float4*>(base + (size_t)sk * 4));
            float4 xv = z4;
            const float* src = base + (size_t)sk * D;
            xv.x = __ldcg(src);
            if (D > 1) xv.y = __ldcg(src + 1);
            if (D > 2) xv.z = __ldcg(src + 2);
            return xv;
          };
          const float4 x0 = ldx(eng.ring_obs, s0);
          // g_item: the vector step that wrote ring row rho (the ring holds the last R steps)
          const uint32_t back = glR >= (uint32_t)rho ? glR - (uint32_t)rho : glR + (uint32_t)R - (uint32_t)rho;
          const uint64_t g_item = (vec_steps - 1) - (uint64_t)back;
          if (FLAG) {
            int a[3]; float rw[3]; unsigned char tm[3], dn[3]; float4 xv[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              int rk = rho + k; rk = rk >= R ? rk - R : rk;
              const int sk = rk * E + e;
              a[k] = __ldcg(eng.ring_action + sk); rw[k] = __ldcg(eng.ring_reward + sk);
              tm[k] = __ldcg(eng.ring_term + sk); dn[k] = __ldcg(eng.ring_done + sk);
              xv[k] = ldx(eng.ring_next_obs, sk);
            }
            bool ended = false;
            int last_k = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              if (ended) {  // padded tail record: random action, reward 0, terminated 1, state = last next_state (rainbow.py:358-371)
                const uint64_t gp = g_item + (uint64_t)k;
                const uint4 pw = philox_ni(eng.seed, STREAM_PAD_ACTION, (uint32_t)e, (uint32_t)gp, (uint32_t)(gp >> 32));
                a[k] = (int)u_below(pw.x, (uint32_t)A); rw[k] = 0.f; tm[k] = 1;
                xv[k] = last_k == 0 ? xv[0] : xv[1];
              } else {
                last_k = k;
                if (dn[k]) ended = true;
              }
            }
            if (v) {
              *reinterpret_cast<float4*>(x_cur + (size_t)lane * 4) = x0;
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                const int w = lane * 3 + k;
                w_act[w] = a[k]; w_rew[w] = rw[k]; w_term[w] = (float)tm[k];
                *reinterpret_cast<float4*>(x_cur + (size_t)(B + w) * 4) = xv[k];
              }
            }
          } else {
            if (v) *reinterpret_cast<float4*>(x_cur + (size_t)lane * 4) = x0;
            bool ended = false;
            float4 xlast = z4;
            for (int k = 0; k < M; ++k) {
              const int w = lane * M + k;
              int a; float rw, tm; float4 xv;
              if (!ended) {
                const int sk = ((rho + k) % R) * E + e;
                a = __ldcg(eng.ring_action + sk); rw = __ldcg(eng.ring_reward + sk);
                tm = (float)__ldcg(eng.ring_term + sk);
                const int dn = (int)__ldcg(eng.ring_done + sk);
                xv = ldx(eng.ring_next_obs, sk);
                xlast = xv;
                if (dn) ended = true;
              } else {
                const uint64_t gp = g_item + (uint64_t)k;
                const uint4 pw = philox_ni(eng.seed, STREAM_PAD_ACTION, (uint32_t)e, (uint32_t)gp, (uint32_t)(gp >> 32));
                a = (int)u_below(pw.x, (uint32_t)A); rw = 0.f; tm = 1.f; xv = xlast;
              }
              if (v) {
                w_act[w] = a; w_rew[w] = rw; w_term[w] = tm;
                *reinterpret_cast<float4*>(x_cur + (size_t)(B + w) * 4) = xv;
              }
            }
          }
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_local(&mbar[MB_XR]);  // x(t) ready for the compute warps
            if (upd + 1 < n_updates) mbar_expect_tx(&mbar[MB_S], (uint32_t)B4 * 4);
          }
          SRLX_FSTAMP(mt == 0, 17);
          if (rank == 0 && eng.dbg_windows) {
            float* dw = eng.dbg_windows;
            const int n_states = B * (M + 1) * D;
            for (int w = lane; w < n_states; w += 32) {
              const int i = w / ((M + 1) * D), rem = w - i * (M + 1) * D, k = rem / D, d = rem - k * D;
              dw[w] = (k == 0) ? x_cur[(size_t)i * 4 + d] : x_cur[(size_t)(B + i * M + k - 1) * 4 + d];
            }
            for (int w = lane; w < BM; w += 32) {
              dw[n_states + w] = (float)w_act[w];
              dw[n_states + BM + w] = w_rew[w];
              dw[n_states + 2 * BM + w] = w_term[w];
            }
          }
        }
        if (rank == 0) {
          // ---- off the critical path: after the compute warps have issued forward(t), while they wait for the targets --
          if (mw == 0) {
            send_weights(tc, parb);
            if (upd > 0) bookkeeping(upd - 1);
          } else if (per) {
            plan_update(upd + 2 == n_updates);
          }
          if (per && mw == 0 && lane < B && upd + 1 < n_updates) u_next = draw(tc + 1, lane, 0);
          SRLX_FSTAMP(mt == 0, 26);
          SRLX_FSTAMP(mt == 32, 27);
          if (presample && upd + 1 < n_updates) {
            // pre-sampled mode: batch t+1 is drawn NOW, from the tree as updates <= t-1 left it; update t lands afterwards
            if (per) {
              if (mw == 0) walk_top(u_next);
              named_bar_sync(FBAR_MEM, FNM);
            }
            sample_slots(tc + 1, parb ^ 1, upd + 2 == n_updates);
          }
        }
        if (mw == 0) {
          if (rank == 0) {
            mbar_wait_sleep(&mbar[MB_TQ], parb);  // (target, q) of every item, ahead of the all-gather
            if (per) named_bar_arrive(FBAR_MA, FNM);  // release the update warps first
            if (lane == 0 && upd + 1 < n_updates) mbar_expect_tx(&mbar[MB_TQ], (uint32_t)B * 8);
          }
          mbar_wait_sleep(&mbar[MB_AG], parb);  // update t's targets are known everywhere: forward(t) is over in every CTA
          if (lane == 0 && noisy && upd + 2 < n_updates) {  // ring slot (t+2)%3 was last read by Adam(t-1)
            uint64_t* nb = &mbar[MB_NZ0 + (upd + 2) % 3];
            mbar_expect_tx(nb, (uint32_t)nz_bytes);
            bulk_g2s(nzr + (size_t)((upd + 2) % 3) * 3 * Pl, nz_src(upd + 2), (uint32_t)nz_bytes, nb);
          }
        }
        if (rank == 0) {
          SRLX_FSTAMP(mt == 0, 18);
          if (per) {
            if (mw != 0) {
              named_bar_sync(FBAR_MA, FNM);  // targets of update t have arrived (warp 0 saw the AG barrier complete)
              apply_update(tqb + (size_t)parb * B * 2, upd + 2 == n_updates);
            } else {
              // max_priority (proportional_memory.py:176): the priority is monotone in |td|, so one evaluation at max |td|
              const float* tq = tqb + (size_t)parb * B * 2;
              float m = lane < B ? fabsf(tq[lane * 2] - tq[lane * 2 + 1]) : 0.f;
              for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, s));
              if (lane == 0) {
                const double pm = pow_chain(fabs((double)m) + eng.per_epsilon, eng.per_alpha);
                if (pm > sc->max_priority) sc->max_priority = pm;
              }
              named_bar_sync(FBAR_MB, FNM);  // cached levels of the tree are updated
              SRLX_FSTAMP(mt == 0, 28);
              if (upd + 1 < n_updates && !presample) walk_top(u_next);
            }
            named_bar_sync(FBAR_MEM, FNM);  // deep levels written + top-level walk handed over
          } else {
            named_bar_sync(FBAR_MEM, FNM);
          }
          SRLX_FSTAMP(mt == 0, 20);
          if (upd + 1 < n_updates && !presample) sample_slots(tc + 1, parb ^ 1, upd + 2 == n_updates);
          SRLX_FSTAMP(mt == 0, 21);
        }
      }
      if (rank == 0 && mw == 0) bookkeeping(n_updates - 1);
    }
  done_roles:;
  }

  // ---- write back: parameters, Adam moments, target copy, counters ----------------------------------------------------
  __syncthreads();
  for (int s = 0; s < n_seg; ++s) {
    const FSeg sg = segs[s];
    if (sg.replicated && rank != 0) continue;
    for (int j = tid; j < sg.n; j += FNT) {
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
  if (DP && tid == 0 && *dp_dead) st->reserved[0] = 1;  // a peer rank stopped answering: the host raises (check_dp_alive)
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

// ---- host side ---------------------------------------------------------------------------------------------------
static int learn_fast(const srlx_engine* eng, uint32_t n_updates, uintptr_t cuda_stream, int C, int cache_levels) {
  const long long n_nodes = 2ll * eng->ring_rows * eng->n_envs - 1;
  const FPlan pl = make_fplan(*eng, C, n_nodes, cache_levels);
  // the reference's Rainbow default shape gets the instantiations with compile-time loop bounds
  const bool flag = eng->batch_size == 32 && eng->multisteps == 3 && eng->n_actions == 2 && eng->obs_dim == 4 &&
                    eng->net.out_dim[1] == 3 && eng->net.dueling == SRLX_DUEL_AVERAGE && eng->net.out_dim[0] == 1024;
  const bool dp = eng->dp_world > 1;
  auto kern = dp ? ((flag && C == 16) ? learner_fast_kernel<16, true> : learner_fast_kernel<0, true>)
                 : ((flag && C == 8) ? learner_fast_kernel<8, false> : (flag && C == 16) ? learner_fast_kernel<16, false> : learner_fast_kernel<0, false>);
  SRLX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.total));
  if (C > 8) SRLX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const size_t per_update = (size_t)C * 3 * pl.Pl * 4;
  uint32_t chunk = kFMaxChunk;
  if (eng->net.noisy) {
    const uint64_t fit = eng->noise_scratch_bytes / per_update;
    if (fit < chunk) chunk = (uint32_t)fit;
  }
  if (const char* e = getenv("SRLX_CHUNK")) {
    const int v = atoi(e);
    if (v >= 1 && (uint32_t)v < chunk) chunk = (uint32_t)v;
  }
  if (eng->dp_world > 1) {
    SRLX_REQUIRE(eng->dp_world <= kDpMaxWorld && eng->dp_rank >= 0 && eng->dp_rank < eng->dp_world, "dp_rank %d / dp_world %d out of range",
                 eng->dp_rank, eng->dp_world);
    for (int r = 0; r < eng->dp_world; ++r) SRLX_REQUIRE(eng->dp_peer[r] != nullptr, "dp_peer[%d] is NULL", r);
    SRLX_REQUIRE(eng->dp_bytes >= dp_bytes_for(round_up(pl.Pl, 4)), "dp exchange buffer too small: %llu < %zu bytes",
                 (unsigned long long)eng->dp_bytes, dp_bytes_for(round_up(pl.Pl, 4)));
  }
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  // blocked copy of the deep tree levels: the rollout writes only the flat tree, so rebuild once per call
  const void* l2_base = eng->tree;
  size_t l2_want = eng->mem_kind == SRLX_MEM_PROPORTIONAL ? (size_t)n_nodes * 8 : 0;
  if (eng->mem_kind == SRLX_MEM_PROPORTIONAL && eng->tree_blk) {
    int clev = 0;
    while (clev < cache_levels && clev < kFCacheLevels && ((1ll << (clev + 1)) - 1) <= n_nodes) ++clev;
    const BlkPlan bp = make_blk_plan(n_nodes, clev);
    if (bp.n_tiers > 0 && eng->tree_blk_bytes >= (uint64_t)bp.total * 512) {
      tree_blk_build_kernel<<<(unsigned)((bp.total + 3) / 4), 256, 0, stream>>>(eng->tree, (int)n_nodes, clev, eng->tree_blk);
      count_launch();
      SRLX_CHECK_CUDA(cudaGetLastError());
      l2_base = eng->tree_blk;  // what the sampler re-reads at random every update
      l2_want = (size_t)bp.total * 512;
    }
  }
  // L2 set-aside (persisting access-policy window, a launch attribute of the learner) for the structure the sampler reads:
  // its deep levels are touched once per ~4000 updates each and would otherwise be evicted by the ring / noise traffic.
  // The device-wide set-aside size only ever grows; SRLX_L2_PERSIST=0 disables.
  size_t l2_window_bytes = 0;
  {
    const char* e = getenv("SRLX_L2_PERSIST");
    if (l2_want > 0 && !(e && e[0] == '0')) {
      static size_t persist_limit[64] = {};  // current cudaLimitPersistingL2CacheSize we set, per device
      static int persist_state[64] = {};     // 0 = unknown, 1 = available, -1 = unavailable
      static size_t persist_max[64] = {}, window_max[64] = {};
      int dev = 0;
      SRLX_CHECK_CUDA(cudaGetDevice(&dev));
      if (dev >= 0 && dev < 64) {
        if (persist_state[dev] == 0) {
          int max_persist = 0, max_window = 0;
          cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
          cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
          persist_max[dev] = (size_t)(max_persist > 0 ? max_persist : 0);
          window_max[dev] = (size_t)(max_window > 0 ? max_window : 0);
          persist_state[dev] = (max_persist > 0 && max_window > 0) ? 1 : -1;
        }
        if (persist_state[dev] == 1) {
          const size_t want = l2_want < persist_max[dev] ? l2_want : persist_max[dev];
          if (want > persist_limit[dev]) {
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) persist_limit[dev] = want;
            else cudaGetLastError();
          }
          if (persist_limit[dev] > 0) l2_window_bytes = l2_want < window_max[dev] ? l2_want : window_max[dev];
        }
      }
    }
  }
  for (uint32_t done = 0; done < n_updates;) {
    const uint32_t n = (n_updates - done) < chunk ? (n_updates - done) : chunk;
    if (eng->net.noisy) {
      noise_precompute_kernel<<<n * C, 128, 0, stream>>>(*eng, n, C, pl.Pl, eng->noise_scratch);
      count_launch();
      SRLX_CHECK_CUDA(cudaGetLastError());
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)C, 1, 1);
    cfg.blockDim = dim3(FNT, 1, 1);
    cfg.dynamicSmemBytes = pl.total;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (l2_window_bytes > 0) {  // keep the SumTree resident in L2: its deep levels are re-read at random every update
      attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
      attr[1].val.accessPolicyWindow.base_ptr = const_cast<void*>(l2_base);
      attr[1].val.accessPolicyWindow.num_bytes = l2_window_bytes;
      attr[1].val.accessPolicyWindow.hitRatio = 1.0f;
      attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cfg.numAttrs = 2;
    }
    SRLX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, *eng, n, (const float*)eng->noise_scratch, cache_levels));
    count_launch();
    SRLX_CHECK_CUDA(cudaGetLastError());
    done += n;
  }
  return 0;
}

__global__ void pow_chain_kernel(const double* __restrict__ x, double a, double* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = pow_chain(x[i], a);
}

}  // namespace srlx

// bytes of the blocked deep-tree copy for `capacity` leaves: the largest layout over the cache depths srlx_learn may pick
extern "C" size_t srlx_tree_blk_bytes(uint64_t capacity) {
  using namespace srlx;
  const long long n_nodes = 2ll * (long long)capacity - 1;
  size_t best = 0;
  for (int clev = 1; clev <= kFCacheLevels; ++clev) {
    if (((1ll << clev) - 1) > n_nodes) break;
    const BlkPlan bp = make_blk_plan(n_nodes, clev);
    if ((size_t)bp.total * 512 > best) best = (size_t)bp.total * 512;
  }
  return best;
}

// out[i] = the learner's priority power x[i]^a (test tap: parity of pow_chain with libm pow)
extern "C" int srlx_dbg_pow(const double* x_dev, double a, double* out_dev, size_t n, uintptr_t cuda_stream) {
  using namespace srlx;
  SRLX_REQUIRE(x_dev && out_dev, "srlx_dbg_pow: NULL buffer");
  if (n == 0) return 0;
  pow_chain_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(x_dev, a, out_dev, n);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Pick the cluster size and the number of SumTree levels cached in shared memory: the widest cluster (16, then 8, ...)
// whose plan fits the per-CTA shared memory with at least 8 cached levels.  Returns 1 if the fast kernel applies.
static int fast_choose(const srlx_engine* eng, int* C_out, int* lev_out, size_t* smem_out) {
  using namespace srlx;
  const char* force = getenv("SRLX_LEARNER");
  if ((force && force[0] == 'g') || !fast_shape_ok(*eng) || eng->ring_invalid) return 0;  // masks: the generic learner
  int want = 0;
  if (const char* e = getenv("SRLX_CLUSTER")) want = atoi(e);
  int dev = 0, max_smem = 0;
  SRLX_CHECK_CUDA(cudaGetDevice(&dev));
  SRLX_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const long long n_nodes = 2ll * eng->ring_rows * eng->n_envs - 1;
  for (int c0 = want > 0 ? want : kFMaxC; c0 >= 1; c0 >>= 1) {
    const int C = fast_pick_cluster(eng->net, c0);
    if (C < 1) return 0;
    for (int lev = kFCacheLevels; lev >= 8; --lev) {
      const FPlan pl = make_fplan(*eng, C, n_nodes, lev);
      const bool noise_ok = !eng->net.noisy || (eng->noise_scratch && eng->noise_scratch_bytes >= (uint64_t)C * 3 * pl.Pl * 4);
      if (!noise_ok) return 0;
      if ((long long)pl.total + 1024 <= max_smem) {
        *C_out = C;
        *lev_out = lev;
        *smem_out = pl.total;
        return 1;
      }
      if (eng->mem_kind != SRLX_MEM_PROPORTIONAL) break;  // no cache to shrink
    }
    if (C == 1) break;
    c0 = C;  // next: half of the cluster just tried
  }
  return 0;
}

namespace srlx {
int small_choose(const srlx_engine* eng, size_t* smem_out, int* C_out);                 // learner_small.cu
int learn_small(const srlx_engine* eng, uint32_t n_updates, uintptr_t cuda_stream);
}  // namespace srlx

// 2 when the row-split kernel (learner_small.cu: uniform replay, no NoisyNet, batch rows over a small cluster) is the one to run
static int small_pick(const srlx_engine* eng, size_t* smem_out, int* C_out) {
  const char* force = getenv("SRLX_LEARNER");
  if ((force && force[0] == 'g') || eng->ring_invalid) return 0;
  return srlx::small_choose(eng, smem_out, C_out) == 1 ? 2 : 0;
}

// Which kernel srlx_learn will run for this engine: 1 = learner_fast_kernel (single hidden layer, <= 4 observation floats,
// <= 4 outputs, batch <= 32), 2 = learner_small_kernel (uniform replay, plain weights, one thread block), 0 = the generic
// learner_kernel; negative = error.
extern "C" int srlx_learner_info(const srlx_engine* eng, int* cluster_size, size_t* smem_bytes) {
  using namespace srlx;
  SRLX_REQUIRE(eng != nullptr, "srlx_learner_info: eng is NULL");
  int C = 0, lev = 0;
  size_t sm = 0;
  int rc = fast_choose(eng, &C, &lev, &sm);
  if (rc == 0) {
    rc = small_pick(eng, &sm, &C);
    if (rc != 2) C = 0;
  }
  if (cluster_size) *cluster_size = rc >= 1 ? C : 0;
  if (smem_bytes) *smem_bytes = rc >= 1 ? sm : 0;
  return rc;
}

// bytes of one rank's exchange buffer for the data-parallel learner (0: this engine does not take the cluster kernel)
extern "C" size_t srlx_dp_bytes(const srlx_engine* eng) {
  using namespace srlx;
  if (eng == nullptr) return 0;
  int C = 0, lev = 0;
  size_t sm = 0;
  if (fast_choose(eng, &C, &lev, &sm) != 1) return 0;
  const FPlan pl = make_fplan(*eng, C, 2ll * eng->ring_rows * eng->n_envs - 1, lev);
  return dp_bytes_for(round_up(pl.Pl, 4));
}

extern "C" int srlx_learn(const srlx_engine* eng, uint32_t n_updates, uintptr_t cuda_stream) {
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
  int C = 0, lev = 0;
  size_t smem_bytes = 0;
  const int rc = fast_choose(eng, &C, &lev, &smem_bytes);
  if (rc < 0) return rc;
  if (rc == 1) return learn_fast(eng, n_updates, cuda_stream, C, lev);
  SRLX_REQUIRE(eng->dp_world <= 1, "the data-parallel learner (dp_world = %d) exists for the single-hidden-layer cluster kernel only", eng->dp_world);
  SRLX_REQUIRE(!eng->presample, "the pre-sampled mode exists for the single-hidden-layer cluster kernel only");
  if (small_pick(eng, &smem_bytes, &C) == 2) return learn_small(eng, n_updates, cuda_stream);
  return learn_generic(eng, n_updates, cuda_stream);
}
