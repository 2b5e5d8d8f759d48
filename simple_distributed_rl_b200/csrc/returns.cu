// returns.cu -- the return accumulation of the reference's PPO worker (srl/algorithms/ppo/ppo.py:357-406) over a time-major
// rollout buffer [T][E]: GAE advantages or Monte-Carlo returns, one backward scan per env column, episode ends where done == 1.
//
//   GAE  delta_t = r_t - v_t                       at the last step of an episode (:392-393, no bootstrap)
//        delta_t = r_t + discount * nv_t - v_t     otherwise (:395)
//        gae_t   = delta_t + (discount * gae_discount) * gae_{t+1}      float32, two roundings per step, never fused (:396)
//   MC   mc_t    = r_t + discount * mc_{t+1}       in float64, stored as float32 (:376-383)
//
// The recurrence is sequential in t and has to keep the reference's rounding order, so time cannot be split across threads;
// the bytes can.  A CTA owns 32 adjacent env columns (one 128-byte line per array row) and walks the buffer backwards in chunks
// of kChunk time steps through two shared-memory buffers: seven warps stream the next chunk with independent coalesced loads
// (delta and the episode-end flag are formed on the fly, thread per element) while one warp runs the 32 column recurrences of
// the current chunk in place; then all warps store the finished rows.  Several CTAs per SM keep the memory system busy
// (algorithmic bytes: 13 B read + 5 B written per env step for GAE).  Measured on B200 at 16384 x 200 (profiles/r1_j_*):
// 13.4 us = 4.4 TB/s, 67 % of the measured copy peak (5.1 TB/s = 78 % at 65536 x 200 or 16384 x 1000: the 59 MB job is
// short enough for launch ramp and tail to show), with 16-byte accesses (a thread owns 4 adjacent columns); the scalar
// path, used when n_envs % 4 != 0 or a buffer is not 16-byte aligned, reaches 3.2 TB/s -- it is issue-bound (107 instructions
// per warp row), which the ncu capture showed before the vector path existed.  A cp.async (LDGSTS.32) three-stage variant was
// slower (26.6 us) and is not kept; 2-D TMA tiles over 256-byte column groups are the next step.  CPU twin: oracle/gae.py, pinned by tests/golden/ppo_returns.npz.
#include "common.cuh"

namespace srlx {

constexpr int kRetThreads = 256;
constexpr int kRetCols = 32;
constexpr int kChunk = 56;  // time steps staged per pass (7 loader warps x 8 rows: one memory round trip per chunk); two buffers

template <bool MC, bool VEC>
__global__ void __launch_bounds__(kRetThreads, 4)
returns_scan_kernel(const float* __restrict__ reward, const double* __restrict__ reward64, const float* __restrict__ value,
                    const float* __restrict__ next_value, const unsigned char* __restrict__ done, float* __restrict__ out,
                    unsigned char* __restrict__ valid, const int T, const int E, const double discount, const double gae_discount,
                    const int tail_is_end, const int clip_enable, const double clip_lo, const double clip_hi) {
  using acc_t = typename std::conditional<MC, double, float>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // two staging buffers: [2][kChunk][32] values (delta in, result out) and [2][kChunk][32] flags (episode end in, valid out)
  acc_t* s_val = reinterpret_cast<acc_t*>(smem_raw);
  unsigned char* s_flag = smem_raw + (size_t)2 * kChunk * kRetCols * sizeof(acc_t);
  const int col0 = blockIdx.x * kRetCols;
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int col = col0 + lane;
  const bool col_ok = col < E;
  const float g = (float)discount, c = (float)(discount * gae_discount);
  const int n_chunks = (T + kChunk - 1) / kChunk;
  // chunk k (k = 0 is the LAST one in time) covers [t_lo(k), t_hi(k)); the first (earliest) chunk is the ragged one
  auto t_hi = [&](int k) { return T - k * kChunk; };
  auto t_lo = [&](int k) { const int lo = T - (k + 1) * kChunk; return lo > 0 ? lo : 0; };

  // stream chunk k into buffer b: thread per element, rows dealt to `nw` warps starting at warp w0, every load independent
  auto load_chunk = [&](int k, int b, int w0, int nw) {
    const int lo = t_lo(k), n_t = t_hi(k) - lo;
    acc_t* sv = s_val + (size_t)b * kChunk * kRetCols;
    unsigned char* sf = s_flag + (size_t)b * kChunk * kRetCols;
    if constexpr (VEC) {
      // 16-byte path (n_envs % 4 == 0, 16-byte aligned buffers): a thread owns 4 adjacent columns of a row, a warp covers 4
      // rows per instruction; kV quads per thread are in flight before the first delta is formed
      constexpr int kV = 2;
      const int n_q = n_t * (kRetCols / 4);
      for (int i0 = (wrp - w0) * 32 + lane; i0 < n_q; i0 += nw * 32 * kV) {
        float4 r4[kV], v4[kV], n4[kV];
        double2 ra[kV], rb[kV];
        uchar4 d4[kV];
#pragma unroll
        for (int u = 0; u < kV; ++u) {
          const int i = i0 + u * nw * 32;
          const int tt = i >> 3, q = i & 7;
          const bool ok = i < n_q && col0 + 4 * q < E;
          const size_t o = ok ? (size_t)(lo + tt) * E + col0 + 4 * q : 0;
          d4[u] = __ldcs(reinterpret_cast<const uchar4*>(done + o));
          if (MC) {
            if (reward64) {
              ra[u] = __ldcs(reinterpret_cast<const double2*>(reward64 + o));
              rb[u] = __ldcs(reinterpret_cast<const double2*>(reward64 + o + 2));
            } else {
              r4[u] = __ldcs(reinterpret_cast<const float4*>(reward + o));
            }
          } else {
            r4[u] = __ldcs(reinterpret_cast<const float4*>(reward + o));
            v4[u] = __ldcs(reinterpret_cast<const float4*>(value + o));
            n4[u] = __ldcs(reinterpret_cast<const float4*>(next_value + o));
          }
        }
#pragma unroll
        for (int u = 0; u < kV; ++u) {
          const int i = i0 + u * nw * 32;
          if (i >= n_q) continue;
          const int tt = i >> 3, q = i & 7;
          const bool last = tail_is_end && lo + tt == T - 1;
          const unsigned char dd[4] = {d4[u].x, d4[u].y, d4[u].z, d4[u].w};
          unsigned char en[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) en[e] = (dd[e] != 0 || last) ? 1 : 0;
          acc_t* dst = sv + tt * kRetCols + 4 * q;
          if (MC) {
            double rr[4];
            if (reward64) { rr[0] = ra[u].x; rr[1] = ra[u].y; rr[2] = rb[u].x; rr[3] = rb[u].y; }
            else { rr[0] = r4[u].x; rr[1] = r4[u].y; rr[2] = r4[u].z; rr[3] = r4[u].w; }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (clip_enable) rr[e] = rr[e] < clip_lo ? clip_lo : (rr[e] > clip_hi ? clip_hi : rr[e]);
              dst[e] = (acc_t)rr[e];
            }
          } else {
            float rr[4] = {r4[u].x, r4[u].y, r4[u].z, r4[u].w};
            const float vv4[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w}, nn4[4] = {n4[u].x, n4[u].y, n4[u].z, n4[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (clip_enable) rr[e] = rr[e] < (float)clip_lo ? (float)clip_lo : (rr[e] > (float)clip_hi ? (float)clip_hi : rr[e]);
              dst[e] = (acc_t)(en[e] ? __fsub_rn(rr[e], vv4[e]) : __fsub_rn(__fadd_rn(rr[e], __fmul_rn(g, nn4[e])), vv4[e]));
            }
          }
          *reinterpret_cast<uchar4*>(sf + tt * kRetCols + 4 * q) = make_uchar4(en[0], en[1], en[2], en[3]);
        }
      }
      return;
    }
    // rows in groups of kU: all loads of a group are issued before the first result is formed, so a thread keeps up to
    // 4 * kU loads in flight (a row-at-a-time loop serialises on the shared-memory stores between its loads)
    constexpr int kU = 8;
    for (int tt0 = wrp - w0; tt0 < n_t; tt0 += nw * kU) {
      float r32[kU], vv[kU], nvv[kU];
      double r64[kU];
      unsigned char dn[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int tt = tt0 + u * nw;
        const bool ok = col_ok && tt < n_t;
        const size_t o = ok ? (size_t)(lo + tt) * E + col : 0;  // clamped address: the load is unconditional, the value unused
        dn[u] = __ldcs(done + o);
        if (MC) {
          if (reward64) r64[u] = __ldcs(reward64 + o);
          else r64[u] = (double)__ldcs(reward + o);
        } else {
          r32[u] = __ldcs(reward + o);
          vv[u] = __ldcs(value + o);
          nvv[u] = __ldcs(next_value + o);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int tt = tt0 + u * nw;
        if (tt >= n_t) continue;
        const int t = lo + tt;
        acc_t d = 0;
        unsigned char end = 0;
        if (col_ok) {
          end = dn[u] != 0 || (tail_is_end && t == T - 1);
          if (MC) {
            double r = r64[u];
            if (clip_enable) r = r < clip_lo ? clip_lo : (r > clip_hi ? clip_hi : r);
            d = (acc_t)r;
          } else {
            float r = r32[u];
            if (clip_enable) r = r < (float)clip_lo ? (float)clip_lo : (r > (float)clip_hi ? (float)clip_hi : r);
            d = (acc_t)(end ? __fsub_rn(r, vv[u]) : __fsub_rn(__fadd_rn(r, __fmul_rn(g, nvv[u])), vv[u]));
          }
        }
        sv[tt * kRetCols + lane] = d;
        sf[tt * kRetCols + lane] = end;
      }
    }
  };

  acc_t acc = 0;      // carried by warp 0 across chunks (backwards in time)
  bool live = false;  // an episode end has been seen later in time -> the steps before it are emitted
  load_chunk(0, 0, 0, kRetThreads / 32);
  __syncthreads();
  for (int k = 0; k < n_chunks; ++k) {
    const int b = k & 1;
    acc_t* sv = s_val + (size_t)b * kChunk * kRetCols;
    unsigned char* sf = s_flag + (size_t)b * kChunk * kRetCols;
    const int lo = t_lo(k), n_t = t_hi(k) - lo;
    if (wrp == 0) {
      // the recurrences of the 32 columns, backwards, in place: value <- result, flag <- valid
      // eight steps at a time through registers: the shared-memory loads of a group are independent of the in-place stores of
      // the previous one only if they are issued first, which the explicit staging guarantees
      constexpr int kS = 8;
#pragma unroll 1
      for (int hi = n_t; hi > 0; hi -= kS) {
        acc_t dv[kS];
        bool ev[kS];
#pragma unroll
        for (int u = 0; u < kS; ++u) {
          const int tt = hi - 1 - u;
          const int ix = (tt >= 0 ? tt : 0) * kRetCols + lane;
          dv[u] = sv[ix];
          ev[u] = sf[ix] != 0;
        }
        acc_t rv[kS];
        bool lv[kS];
#pragma unroll
        for (int u = 0; u < kS; ++u) {
          if (hi - 1 - u >= 0) {
            live = live || ev[u];
            const acc_t prev = ev[u] ? (acc_t)0 : acc;
            if (MC) acc = (acc_t)__dadd_rn((double)dv[u], __dmul_rn(discount, (double)prev));
            else acc = (acc_t)__fadd_rn((float)dv[u], __fmul_rn(c, (float)prev));
          }
          rv[u] = live ? acc : (acc_t)0;
          lv[u] = live;
        }
#pragma unroll
        for (int u = 0; u < kS; ++u) {
          const int tt = hi - 1 - u;
          if (tt >= 0) {
            sv[tt * kRetCols + lane] = rv[u];
            sf[tt * kRetCols + lane] = lv[u] ? 1 : 0;
          }
        }
      }
    } else if (k + 1 < n_chunks) {
      load_chunk(k + 1, b ^ 1, 1, kRetThreads / 32 - 1);  // the other seven warps stream the next (earlier) chunk meanwhile
    }
    __syncthreads();
    // results of chunk k -> global, coalesced rows, all warps
    if constexpr (VEC) {
      for (int i = tid; i < n_t * (kRetCols / 4); i += kRetThreads) {
        const int tt = i >> 3, q = i & 7;
        if (col0 + 4 * q < E) {
          const size_t o = (size_t)(lo + tt) * E + col0 + 4 * q;
          const acc_t* src = sv + tt * kRetCols + 4 * q;
          __stcs(reinterpret_cast<float4*>(out + o), make_float4((float)src[0], (float)src[1], (float)src[2], (float)src[3]));
          if (valid) *reinterpret_cast<uchar4*>(valid + o) = *reinterpret_cast<const uchar4*>(sf + tt * kRetCols + 4 * q);
        }
      }
    } else if (col_ok) {
      for (int tt = wrp; tt < n_t; tt += kRetThreads / 32) {
        const size_t o = (size_t)(lo + tt) * E + col;
        __stcs(out + o, (float)sv[tt * kRetCols + lane]);
        if (valid) valid[o] = sf[tt * kRetCols + lane];
      }
    }
    __syncthreads();  // buffer b is free for chunk k + 2
  }
}

}  // namespace srlx

extern "C" int srlx_returns_scan(const float* reward_dev, const double* reward_f64_dev, const float* value_dev,
                                 const float* next_value_dev, const unsigned char* done_dev, float* out_dev, unsigned char* valid_dev,
                                 uint32_t n_steps, uint32_t n_envs, double discount, double gae_discount, int method,
                                 int tail_is_episode_end, int clip_enable, double clip_lo, double clip_hi, uintptr_t cuda_stream) {
  using namespace srlx;
  SRLX_REQUIRE(method == SRLX_RETURNS_GAE || method == SRLX_RETURNS_MC, "srlx_returns_scan: unknown method %d", method);
  SRLX_REQUIRE(done_dev && out_dev, "srlx_returns_scan: done / out buffer is NULL");
  SRLX_REQUIRE(reward_dev || (method == SRLX_RETURNS_MC && reward_f64_dev), "srlx_returns_scan: reward buffer is NULL");
  SRLX_REQUIRE(method == SRLX_RETURNS_MC || (value_dev && next_value_dev), "srlx_returns_scan: GAE needs value and next_value");
  SRLX_REQUIRE((uint64_t)n_steps * n_envs < (1ull << 40), "srlx_returns_scan: buffer too large");
  if (n_steps == 0 || n_envs == 0) return 0;
  const unsigned grid = (n_envs + kRetCols - 1) / kRetCols;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (method == SRLX_RETURNS_MC) {
    const size_t sm = (size_t)2 * kChunk * kRetCols * (sizeof(double) + 1);
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool vec = n_envs % 4 == 0 && al16(reward_f64_dev ? (const void*)reward_f64_dev : (const void*)reward_dev) && al16(out_dev) &&
                     (reinterpret_cast<uintptr_t>(done_dev) & 3) == 0 && (!valid_dev || (reinterpret_cast<uintptr_t>(valid_dev) & 3) == 0) &&
                     !getenv("SRLX_RETURNS_SCALAR");
    if (vec)
      returns_scan_kernel<true, true><<<grid, kRetThreads, sm, st>>>(reward_dev, reward_f64_dev, value_dev, next_value_dev, done_dev, out_dev,
                                                                     valid_dev, (int)n_steps, (int)n_envs, discount, gae_discount,
                                                                     tail_is_episode_end, clip_enable, clip_lo, clip_hi);
    else
      returns_scan_kernel<true, false><<<grid, kRetThreads, sm, st>>>(reward_dev, reward_f64_dev, value_dev, next_value_dev, done_dev, out_dev,
                                                                      valid_dev, (int)n_steps, (int)n_envs, discount, gae_discount,
                                                                      tail_is_episode_end, clip_enable, clip_lo, clip_hi);
  } else {
    const size_t sm = (size_t)2 * kChunk * kRetCols * (sizeof(float) + 1);
    // 16-byte accesses when every row of every buffer starts on a 16-byte boundary
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool vec = n_envs % 4 == 0 && al16(reward_dev) && al16(value_dev) && al16(next_value_dev) && al16(out_dev) &&
                     (reinterpret_cast<uintptr_t>(done_dev) & 3) == 0 && (!valid_dev || (reinterpret_cast<uintptr_t>(valid_dev) & 3) == 0) &&
                     !getenv("SRLX_RETURNS_SCALAR");
    if (vec)
      returns_scan_kernel<false, true><<<grid, kRetThreads, sm, st>>>(reward_dev, reward_f64_dev, value_dev, next_value_dev, done_dev,
                                                                      out_dev, valid_dev, (int)n_steps, (int)n_envs, discount, gae_discount,
                                                                      tail_is_episode_end, clip_enable, clip_lo, clip_hi);
    else
      returns_scan_kernel<false, false><<<grid, kRetThreads, sm, st>>>(reward_dev, reward_f64_dev, value_dev, next_value_dev, done_dev,
                                                                       out_dev, valid_dev, (int)n_steps, (int)n_envs, discount, gae_discount,
                                                                       tail_is_episode_end, clip_enable, clip_lo, clip_hi);
  }
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
