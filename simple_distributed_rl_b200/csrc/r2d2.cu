// r2d2.cu -- R2D2 on device (BASELINE configs[3]): srl/algorithms/r2d2/r2d2.py restated for E vectorised env copies.
//   Worker.on_reset / policy / on_step (:221-318)  -> r2d2_pre_kernel, r2d2_hidden_kernel, lstm_fwd_kernel, sgemm (head), r2d2_act_kernel,
//                                                     r2d2_add_kernel (proportional memory)
//   Trainer.train / _train_on_batches (:90-215)    -> r2d2_sample_*_kernel, r2d2_gather_kernel, lstm_fwd_kernel x (burnin + seq_len + 1)
//                                                     for both networks, head GEMMs, srlx_sequence_targets (the reference's own target loop,
//                                                     bit-exact, csrc/sequence_targets.cu), r2d2_loss_kernel (keras Huber on target * w, q * w),
//                                                     head backward GEMMs, lstm_bwd_kernel x seq_len (BPTT, burn-in outside the tape),
//                                                     weight-gradient GEMMs, r2d2_adam_kernel (keras Adam), priority update, target sync
// The reference's R2D2 is TensorFlow-only (r2d2.py:5); the CPU oracle (oracle/r2d2.py) is a torch RESTATEMENT of the keras model and says so.
// Every dense map is one strided fp32 GEMM (sgemm_kernel) whose bias rides as the last weight column against a trailing 1 in the
// activations, so forward, input gradient and weight gradient (bias included) are the same kernel with different strides.  The LSTM
// step fuses the gate GEMM with the cell update (a thread owns the four gates of a unit); the BPTT step fuses dh = dgates . Wh^T with
// the gate derivatives.  One launch per time step: consecutive steps are data dependent; the whole update is launch-ordered on one
// stream and capturable in a CUDA graph (no host synchronisation, no allocation; the warm-up gate is evaluated on device).
#include <stdlib.h>

#include "envs.cuh"
#include "gemm.cuh"
#include "tree.cuh"

namespace srlx {

constexpr uint32_t R2D2_PAD_TAIL = 0, R2D2_PAD_PRE = 1;

// ---- LSTM step, forward: gates = [x_t | h_{t-1} | 1] . W^T, keras gate order (i, f, c~, o); c_t = f c_{t-1} + i c~; h_t = o tanh(c_t) -------
struct LstmFwdP {
  const float* xh; long long z_xh;   // [z][M][K]
  const float* W0; const float* W1;  // per z: [4u][K]
  const float* c_in; float* c_out; long long z_c;  // [z][M][u]
  float* h_out; long long ld_h, z_h;               // h_t -> row m at h_out + z*z_h + m*ld_h
  float* gates;                       // z == 0 only: [M][4u] activated gates (BPTT), or NULL
  int M, u, K;
  Gate gate;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int BM, int TM>
__global__ void __launch_bounds__((BM / TM) * 8) lstm_fwd_kernel(const LstmFwdP p) {
  if (p.gate.closed()) return;
  constexpr int BN = 32, TN = 4;
  const int z = blockIdx.z;
  const float* W = z ? p.W1 : p.W0;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[TM][TN];
  gemm_mainloop<BM, BN, TM, TN>(p.xh + (long long)z * p.z_xh, p.K, 1, W, 1, p.K, p.M, 4 * p.u, p.K, m0, n0, acc);
  const int tx = threadIdx.x % (BN / TN), ty = threadIdx.x / (BN / TN);
  const int unit = (n0 >> 2) + tx;
  if (unit >= p.u) return;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
    const float ig = sigmoidf_(acc[i][0]), fg = sigmoidf_(acc[i][1]), gg = tanhf(acc[i][2]), og = sigmoidf_(acc[i][3]);
    const long long ci = (long long)z * p.z_c + (long long)m * p.u + unit;
    const float c = fmaf(fg, p.c_in[ci], ig * gg);
    p.c_out[ci] = c;
    p.h_out[(long long)z * p.z_h + (long long)m * p.ld_h + unit] = og * tanhf(c);
    if (p.gates && z == 0) *reinterpret_cast<float4*>(p.gates + ((long long)m * p.u + unit) * 4) = make_float4(ig, fg, gg, og);
  }
}

// cell update on pre-activations a tiled GEMM left in `pre` [M][4u] (the rollout's E-row step: the GEMM runs on the tensor-core tiles)
__global__ void lstm_pointwise_kernel(const float* __restrict__ pre, float* __restrict__ c, float* __restrict__ h_out, long long ld_h, int M, int u) {
  const long long n = (long long)M * u;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / u;
    const int unit = (int)(i - m * u);
    const float4 g4 = *reinterpret_cast<const float4*>(pre + i * 4);
    const float ig = sigmoidf_(g4.x), fg = sigmoidf_(g4.y), gg = tanhf(g4.z), og = sigmoidf_(g4.w);
    const float cn = fmaf(fg, c[i], ig * gg);
    c[i] = cn;
    h_out[m * ld_h + unit] = og * tanhf(cn);
  }
}

// ---- LSTM step, backward (BPTT): dh_t = dgates_{t+1} . Wh + dL/dh_t (head); gate derivatives; dc carried in place ---------------------
struct LstmBwdP {
  const float* dg_next;  // [M][4u] or NULL (last step of the tape)
  const float* W; int in; int K;
  const float* dh_head;  // [M][u]
  const float* gates;    // [M][4u] of step t
  const float* c_prev;   // [M][u] c_{t-1}
  const float* c_cur;    // [M][u] c_t
  float* dc;             // [M][u] in: dL/dc_t from step t+1, out: dL/dc_{t-1}
  float* dg;             // [M][4u] out: gradient wrt the pre-activation gates of step t
  int M, u, first;       // first != 0: dc starts at 0
  Gate gate;
};

template <int BM, int TM>
__global__ void __launch_bounds__((BM / TM) * 8) lstm_bwd_kernel(const LstmBwdP p) {
  if (p.gate.closed()) return;
  constexpr int BN = 32, TN = 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[TM][TN];
  if (p.dg_next) {
    gemm_mainloop<BM, BN, TM, TN>(p.dg_next, 4LL * p.u, 1, p.W + p.in, p.K, 1, p.M, p.u, 4 * p.u, m0, n0, acc);
  } else {
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  }
  const int tx = threadIdx.x % (BN / TN), ty = threadIdx.x / (BN / TN);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int unit = n0 + tx * TN + j;
      if (unit >= p.u) continue;
      const long long ui = (long long)m * p.u + unit;
      const float4 g4 = *reinterpret_cast<const float4*>(p.gates + ui * 4);
      const float ig = g4.x, fg = g4.y, gg = g4.z, og = g4.w;
      const float tc = tanhf(p.c_cur[ui]);
      const float dh = acc[i][j] + p.dh_head[ui];
      const float dcv = (p.first ? 0.f : p.dc[ui]) + dh * og * (1.f - tc * tc);
      float4 d;
      d.x = dcv * gg * ig * (1.f - ig);
      d.y = dcv * p.c_prev[ui] * fg * (1.f - fg);
      d.z = dcv * ig * (1.f - gg * gg);
      d.w = dh * tc * og * (1.f - og);
      *reinterpret_cast<float4*>(p.dg + ui * 4) = d;
      p.dc[ui] = dcv * fg;
    }
  }
}

// ---- persistent unroll: all time steps of both networks in ONE cooperative launch -------------------------------------------------------
// A recurrent step is a [B x K] . [K x 4u] map with B <= 64: far too small to fill the GPU per launch, and consecutive steps are data
// dependent.  So the grid stays resident for the whole sequence: CTA (z, c) owns 8 LSTM units (32 gate rows) of network z and keeps
// their weights in REGISTERS for all steps (lane = gate row, warp = one of 8 k slices: 65 floats per thread at K = 517); per step it
// stages [x_t | h_{t-1} | 1] (B x K floats, L2 hits) into shared memory, every warp accumulates its k slice for all rows out of
// warp-wide BROADCAST float4 reads (one shared-memory wavefront per 4 k values, no bank conflicts by construction), the 8 slices are
// summed in slice order, the thread that owns (row, unit) applies the gates, keeps c in a register, writes h_t into the next step's
// input row and the activated gates for BPTT, and the CTAs of one network meet at a counting barrier in global memory (release
// fence + atomic add, acquire spin).  Weights are read from L2 once per launch instead of once per step.
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_arrive(unsigned* bar) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(bar, 1u);
}
// a CTA that waits longer than ~2 s gives up (the results are then garbage, bar[3] records it) instead of hanging the GPU: the launch
// is cooperative, so every CTA is resident and this cannot happen unless a CTA died
__device__ __forceinline__ void grid_wait(unsigned* bar_base, const unsigned* bar, unsigned target) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (ld_acquire_u32(bar) < target) {
      __nanosleep(20);
      if (clock64() - t0 > 4000000000ll) { atomicExch(bar_base + 3, 0xDEADu); break; }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

struct SeqFwdP {
  const float* W0; const float* W1;  // LSTM weights of the online / target network: [4u][K]
  float* xh; long long z_xh;         // [z][T + 1][B][K]
  float* cbuf; long long z_c;        // [z][T + 1][B][u]
  float* gates;                      // [T][B][4u] activated gates of network 0
  unsigned* bar;                     // [4] arrival counters (forward: [z]; backward: [2]; [3] = time-out flag), zero at launch
  int B, u, D, K, T, cpn;            // T steps; cpn CTAs per network
  Gate gate;
};

// Thread layout of a forward CTA: lane = unit * 4 + ksub, slice = warp * 4 + ksub: 32 slices of 16 h columns (u <= 512); slice s also
// takes "extra" column s of the [x | 1] part (D + 1 <= 32 columns).  A thread owns the FOUR gates of its unit over its slice: 64 + 4
// weights in registers, and every h value it reads from shared memory feeds 4 FMAs (shared-memory wavefronts, not FMA issue, bound this
// loop: a float4 read is served one quarter-warp at a time, and the slice offsets are skewed so that the 4 addresses of a quarter-warp
// fall into different banks).  Rows go in passes of 16: partial sums -> shared [slice][row][gate row], summed in slice order by the
// two threads that own (row, unit), which keep c in a register across all steps.
#ifndef SRLX_FWD_RP
#define SRLX_FWD_RP 16
#endif
constexpr int kFwdRP = SRLX_FWD_RP;  // rows per pass of the forward unroll
constexpr int kFwdXS = 608, kFwdExtra = 576, kFwdRedStride = 16 * 32 + 8;
__device__ __forceinline__ int fwd_hoff(int s) { return s * 16 + (s >> 1) * 4; }

template <int ROWS>
__global__ void __launch_bounds__(256, 1) lstm_seq_fwd_kernel(const SeqFwdP p) {
  if (p.gate.closed()) return;
  extern __shared__ __align__(16) float seq_smem[];
  constexpr int XS = kFwdXS, RP = kFwdRP, NPASS = ROWS / RP, RSTR = kFwdRedStride;
  float* xs = seq_smem;              // [ROWS][XS]
  float* red = seq_smem + ROWS * XS; // [32 slices][RSTR]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, j = lane >> 2, ksub = lane & 3, sl = warp * 4 + ksub;
  const int z = blockIdx.x / p.cpn, unit0 = (blockIdx.x % p.cpn) * 8;
  const int B = p.B, u = p.u, D = p.D, K = p.K;
  const float* __restrict__ W = z ? p.W1 : p.W0;
  const int unit = unit0 + j;
  float wh[4][16], we[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float* wr = W + (size_t)(unit * 4 + q) * K;
#pragma unroll
    for (int i = 0; i < 16; ++i) wh[q][i] = (unit < u && sl * 16 + i < u) ? __ldg(wr + D + sl * 16 + i) : 0.f;
    we[q] = (unit < u && sl <= D) ? __ldg(wr + (sl < D ? sl : K - 1)) : 0.f;
  }
  for (int i = tid; i < ROWS * XS; i += 256) xs[i] = 0.f;
  __syncthreads();
  for (int m = tid; m < B; m += 256) xs[m * XS + kFwdExtra + D] = 1.f;  // the bias column's partner
  float* xh = p.xh + (size_t)z * p.z_xh;
  float* cb = p.cbuf + (size_t)z * p.z_c;
  // finalising thread: half = which 16 slices it sums, (fm, fj) = row within the pass / unit
  const int half = (tid >> 4) & 1, pair = (tid & 15) + (tid >> 5) * 16, fm = pair >> 3, fj = pair & 7, funit = unit0 + fj;
  float c_reg[NPASS];
#pragma unroll
  for (int ps = 0; ps < NPASS; ++ps) {
    const int m = ps * RP + fm;
    c_reg[ps] = (m < B && funit < u) ? __ldcg(cb + (size_t)m * u + funit) : 0.f;
  }
  __syncthreads();
  for (int t = 0; t < p.T; ++t) {
    if (t > 0) grid_wait(p.bar, p.bar + z, (unsigned)t * (unsigned)p.cpn);
    const float* src = xh + (size_t)t * B * K;
    // The rows are K = D + u + 1 floats apart (odd: the bias column), so the copies are 4 bytes wide.  A load -> register -> store loop
    // leaves the step bound by L2 latency (the weights take the registers a deep software pipeline would need), so the copies are
    // cp.async: ~130 per thread, all in flight at once.  4-byte cp.async exists only as .ca (through L1): safe when no 128-byte line
    // straddles two time rows (B * K * 4 bytes a multiple of 128, i.e. B a multiple of 32) -- every line is then fetched once per
    // launch, after the step that wrote it.  Other batch sizes take plain ld.cg loads.
    if ((B & 31) == 0) {
      for (int m = warp; m < B; m += 8) {
        const float* row = src + (size_t)m * K;
        float* dst = xs + m * XS;
        for (int k = lane; k < u; k += 32) cp_async4(dst + fwd_hoff(k >> 4) + (k & 15), row + D + k);
        if (lane < D) cp_async4(dst + kFwdExtra + lane, row + lane);
      }
      cp_async_commit();
      cp_async_wait<0>();
    } else {
      for (int m = warp; m < B; m += 8) {
        const float* row = src + (size_t)m * K;
        float* dst = xs + m * XS;
        for (int k = lane; k < u; k += 32) dst[fwd_hoff(k >> 4) + (k & 15)] = __ldcg(row + D + k);
        if (lane < D) dst[kFwdExtra + lane] = __ldcg(row + lane);
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int ps = 0; ps < NPASS; ++ps) {
      const int r0 = ps * RP;
      if (r0 < B) {
        float acc[RP][4];
#pragma unroll
        for (int m = 0; m < RP; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
        const float* xrow = xs + r0 * XS + fwd_hoff(sl);
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
#pragma unroll
          for (int m = 0; m < RP; ++m) {
            const float4 a = *reinterpret_cast<const float4*>(xrow + m * XS + i);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              acc[m][q] = fmaf(a.x, wh[q][i], acc[m][q]);
              acc[m][q] = fmaf(a.y, wh[q][i + 1], acc[m][q]);
              acc[m][q] = fmaf(a.z, wh[q][i + 2], acc[m][q]);
              acc[m][q] = fmaf(a.w, wh[q][i + 3], acc[m][q]);
            }
          }
        }
#pragma unroll
        for (int m = 0; m < RP; ++m) {
          const float a = xs[(r0 + m) * XS + kFwdExtra + sl];
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[m][q] = fmaf(a, we[q], acc[m][q]);
          *reinterpret_cast<float4*>(red + sl * RSTR + m * 32 + j * 4) = make_float4(acc[m][0], acc[m][1], acc[m][2], acc[m][3]);
        }
      }
      __syncthreads();
      if (r0 < B && pair < RP * 8) {
        float4 sg = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s2 = 0; s2 < 16; ++s2) {
          const float4 r = *reinterpret_cast<const float4*>(red + (half * 16 + s2) * RSTR + fm * 32 + fj * 4);
          sg.x += r.x; sg.y += r.y; sg.z += r.z; sg.w += r.w;
        }
        // slices 0..15 + slices 16..31, in that order on both threads of the pair
        const float ox = __shfl_xor_sync(0xffffffffu, sg.x, 16), oy = __shfl_xor_sync(0xffffffffu, sg.y, 16);
        const float oz = __shfl_xor_sync(0xffffffffu, sg.z, 16), ow = __shfl_xor_sync(0xffffffffu, sg.w, 16);
        if (half == 0) { sg.x += ox; sg.y += oy; sg.z += oz; sg.w += ow; }
        else { sg.x = ox + sg.x; sg.y = oy + sg.y; sg.z = oz + sg.z; sg.w = ow + sg.w; }
        const int m = r0 + fm;
        if (m < B && funit < u) {
          const float ig = sigmoidf_(sg.x), fg = sigmoidf_(sg.y), gg = tanhf(sg.z), og = sigmoidf_(sg.w);
          const float c = fmaf(fg, c_reg[ps], ig * gg);
          c_reg[ps] = c;
          if (half == 0) {
            __stcg(cb + ((size_t)(t + 1) * B + m) * u + funit, c);
            __stcg(xh + ((size_t)(t + 1) * B + m) * K + D + funit, og * tanhf(c));
            if (z == 0 && p.gates) __stcg(reinterpret_cast<float4*>(p.gates + (((size_t)t * B + m) * u + funit) * 4), make_float4(ig, fg, gg, og));
          }
        }
      }
      __syncthreads();
    }
    if (t + 1 < p.T) grid_arrive(p.bar + z);
  }
}

// BPTT, persistent: CTA c owns 4 units; dh_t = dgates_{t+1} . Wh needs every gate row of the step before, so the CTA's 4u x 4 slice of
// Wh sits in registers and the [B x 4u] gate gradients are streamed through shared memory in windows of 256 gate rows (cp.async,
// double-buffered, so the L2 reads of window v + 1 overlap the FMAs of window v).  lane = gs * 4 + rg: the thread handles rows
// rg, rg + 4, ... and, in every window, 4 gate rows (slice warp * 8 + gs) for all 4 units: each float4 it reads feeds 16 FMAs; row
// stride 264 puts the 8 addresses of a quarter-warp into 8 different bank groups.  The 64 slices are summed in slice order by the
// thread that owns (row, unit), which keeps dc in a register across all steps.
struct SeqBwdP {
  const float* W; int K, D;   // online LSTM weights [4u][K]
  const float* dh_head;       // [S][B][u]
  const float* gates;         // [S][B][4u] activated gates of the trained steps
  const float* cbuf;          // [S + 1][B][u]: c before trained step 0, then after every trained step
  float* dgates;              // [S][B][4u]
  unsigned* bar;              // the same [4] words
  int B, u, S, n_cta;
  Gate gate;
};

constexpr int kBwdRow = 264, kBwdRedStride = 272;
template <int ROWS>
__global__ void __launch_bounds__(256, 1) lstm_seq_bwd_kernel(const SeqBwdP p) {
  if (p.gate.closed()) return;
  extern __shared__ __align__(16) float seq_smem[];
  constexpr int RW = kBwdRow, NR = ROWS / 4, RSTR = kBwdRedStride, NWMAX = 8;
  float* ds = seq_smem;                 // [2][ROWS][RW]
  float* red = seq_smem + 2 * ROWS * RW;  // [64 slices][RSTR]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, rg = lane & 3, gs = lane >> 2, sl = warp * 8 + gs;
  const int B = p.B, u = p.u, G = 4 * u, unit0 = blockIdx.x * 4;
  const int nw = (G + 255) / 256;
  float w[NWMAX][4][4];  // [window][gate row in the slice][unit]
#pragma unroll
  for (int v = 0; v < NWMAX; ++v)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int g = v * 256 + sl * 4 + i;
        w[v][i][jj] = (g < G && unit0 + jj < u) ? __ldg(p.W + (size_t)g * p.K + p.D + unit0 + jj) : 0.f;
      }
  for (int i = tid; i < 2 * ROWS * RW; i += 256) ds[i] = 0.f;
  const int fm = tid >> 2, fj = tid & 3, funit = unit0 + fj;  // the (row, unit) this thread finalises (ROWS * 4 <= 256)
  const bool fok = fm < B && funit < u && fm < ROWS;
  float dc_reg = 0.f;
  __syncthreads();
  auto stage = [&](const float* dgn, int v, int buf) {
    float* dst = ds + buf * ROWS * RW;
#pragma unroll
    for (int it = 0; it < ROWS * 64 / 256; ++it) {
      const int idx = tid + it * 256, m = idx >> 6, g = v * 256 + (idx & 63) * 4;
      if (m < B && g < G) cp_async16(dst + m * RW + (idx & 63) * 4, dgn + (size_t)m * G + g);
    }
    cp_async_commit();
  };
  for (int t = p.S - 1; t >= 0; --t) {
    const bool have_next = t < p.S - 1;
    if (have_next) {
      grid_wait(p.bar, p.bar + 2, (unsigned)(p.S - 1 - t) * (unsigned)p.n_cta);
      const float* dgn = p.dgates + (size_t)(t + 1) * B * G;
      float acc[NR][4];
#pragma unroll
      for (int i = 0; i < NR; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      stage(dgn, 0, 0);
#pragma unroll
      for (int v = 0; v < NWMAX; ++v) {
        if (v < nw) {
          if (v + 1 < nw) {
            stage(dgn, v + 1, (v + 1) & 1);
            cp_async_wait<1>();
          } else {
            cp_async_wait<0>();
          }
          __syncthreads();
          const float* base = ds + (v & 1) * ROWS * RW + rg * RW + sl * 4;
#pragma unroll
          for (int i = 0; i < NR; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(base + i * 4 * RW);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              acc[i][jj] = fmaf(a.x, w[v][0][jj], acc[i][jj]);
              acc[i][jj] = fmaf(a.y, w[v][1][jj], acc[i][jj]);
              acc[i][jj] = fmaf(a.z, w[v][2][jj], acc[i][jj]);
              acc[i][jj] = fmaf(a.w, w[v][3][jj], acc[i][jj]);
            }
          }
          __syncthreads();
        }
      }
#pragma unroll
      for (int i = 0; i < NR; ++i)
        *reinterpret_cast<float4*>(red + sl * RSTR + (i * 4 + rg) * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      __syncthreads();
    }
    if (fok) {
      float dh = p.dh_head[((size_t)t * B + fm) * u + funit];
      if (have_next) {
        float sum = 0.f;
#pragma unroll 8
        for (int s2 = 0; s2 < 64; ++s2) sum += red[s2 * RSTR + fm * 4 + fj];
        dh += sum;
      }
      const size_t ui = ((size_t)t * B + fm) * u + funit;
      const float4 g4 = *reinterpret_cast<const float4*>(p.gates + ui * 4);
      const float ig = g4.x, fg = g4.y, gg = g4.z, og = g4.w;
      const float c_prev = p.cbuf[ui], tc = tanhf(p.cbuf[ui + (size_t)B * u]);
      const float dcv = dc_reg + dh * og * (1.f - tc * tc);
      float4 d;
      d.x = dcv * gg * ig * (1.f - ig);
      d.y = dcv * c_prev * fg * (1.f - fg);
      d.z = dcv * ig * (1.f - gg * gg);
      d.w = dh * tc * og * (1.f - og);
      __stcg(reinterpret_cast<float4*>(p.dgates + ui * 4), d);
      dc_reg = dcv * fg;
    }
    if (t > 0) grid_arrive(p.bar + 2);
  }
}

// ---- ring geometry -------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long ring_slot(const srlx_r2d2& r, long long pos, int e) {
  return (pos % r.env.ring_rows) * (long long)r.env.n_envs + e;
}
// SumTree leaf of ring row (pos % R) of column e: ENV-major, so that the rows one env writes in a step (up to seq_len at an episode's
// end) are adjacent leaves and share their ancestors -- the bulk add rebuilds ~2 nodes per row instead of one per row and level
__device__ __forceinline__ long long ring_leaf(const srlx_r2d2& r, long long pos, int e) {
  return (long long)e * r.env.ring_rows + (pos % r.env.ring_rows);
}
// position held by ring row `row` of column e, or -1 if the row has never been written
__device__ __forceinline__ long long ring_pos_of_row(const srlx_r2d2& r, int row, long long cur) {
  const int R = r.env.ring_rows;
  if (cur <= R) return row < cur ? row : -1;
  const long long last = cur - 1;
  long long back = (last - row) % R;
  if (back < 0) back += R;
  return last - back;
}
// an anchor is sampleable while the burnin + seq_len rows in front of it have not been overwritten
__device__ __forceinline__ bool ring_anchor_valid(const srlx_r2d2& r, long long pos, long long cur) {
  if (pos < 0) return false;
  const int R = r.env.ring_rows, W = r.burnin + r.seq_len;
  return cur <= R || pos - (W - 1) >= cur - R;
}

// ---- rollout ---------------------------------------------------------------------------------------------------------------------------
// Worker.on_reset + the state half of on_step: reset-if-needed, observation -> network input and ring row
__global__ void r2d2_pre_kernel(const __grid_constant__ srlx_r2d2 r, const int training) {
  const srlx_engine& eng = r.env;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= eng.n_envs) return;
  const int D = eng.obs_dim, K = D + r.lstm_units + 1;
  double* st = eng.env_state + (size_t)e * 4;
  unsigned char was_reset = 0;
  if (eng.env_needs_reset[e]) {
    const uint32_t ep = eng.env_episode[e];
    env_reset(eng, (uint32_t)e, ep, st);
    eng.env_episode[e] = ep + 1;
    eng.env_step_num[e] = 0;
    eng.env_ep_reward[e] = 0.0;
    eng.env_needs_reset[e] = 0;
    was_reset = 1;
  }
  r.roll_reset[e] = was_reset;
  float obs[SRLX_MAX_OBS];
  env_obs(eng, st, obs);
  float* x = r.roll_xh + (size_t)e * K;
  for (int d = 0; d < D; ++d) x[d] = obs[d];
  x[K - 1] = 1.f;
  if (training) {
    const long long slot = ring_slot(r, r.cursor[e], e);
    for (int d = 0; d < D; ++d) r.ring_obs[slot * D + d] = obs[d];
    r.ring_tstep[slot] = eng.env_step_num[e];
  }
}

// LSTM state: zero at an episode's first step (get_initial_state, r2d2.py:243), else the h_t / c_t the previous step left in roll_h /
// roll_c; stored with the row BEFORE the step consumes the state (recent_hidden_states[i] belongs to _recent_states[i], :232-283)
__global__ void r2d2_hidden_kernel(const __grid_constant__ srlx_r2d2 r, const int training) {
  const srlx_engine& eng = r.env;
  const int u = r.lstm_units, D = eng.obs_dim, K = D + u + 1;
  const long long n = (long long)eng.n_envs * u;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i / u), j = (int)(i - (long long)e * u);
    float h = r.roll_h[(size_t)e * (u + 1) + j], c = r.roll_c[i];
    if (r.roll_reset[e]) { h = 0.f; c = 0.f; r.roll_c[i] = 0.f; }
    r.roll_xh[(size_t)e * K + D + j] = h;
    if (j == 0) r.roll_h[(size_t)e * (u + 1) + u] = 1.f;
    if (training) {
      const long long slot = ring_slot(r, r.cursor[e], e);
      r.ring_h[slot * u + j] = h;
      r.ring_c[slot * u + j] = c;
    }
  }
}

__device__ __forceinline__ void dueling_combine(const float* o, int A, int dueling, float* q) {
  if (dueling == SRLX_DUEL_NONE) {
    for (int a = 0; a < A; ++a) q[a] = o[a];
    return;
  }
  const float v = o[0];
  float red = 0.f;
  if (dueling == SRLX_DUEL_AVERAGE) {
    for (int a = 0; a < A; ++a) red += o[1 + a];
    red /= (float)A;
  } else if (dueling == SRLX_DUEL_MAX) {
    red = o[1];
    for (int a = 1; a < A; ++a) red = fmaxf(red, o[1 + a]);
  }
  for (int a = 0; a < A; ++a) q[a] = v + o[1 + a] - red;
}

// Worker.policy (epsilon-greedy probabilities, funcs.calc_epsilon_greedy_probs / random_choice_by_probs, srl/rl/functions.py:188-214) +
// env.step + the transition half of Worker.on_step, the padded tail at an episode's end included (r2d2.py:285-303)
__global__ void r2d2_act_kernel(const __grid_constant__ srlx_r2d2 r, const int training) {
  const srlx_engine& eng = r.env;
  __shared__ unsigned long long s_episodes, s_eplen, s_rows;
  __shared__ double s_epreward;
  if (threadIdx.x == 0) { s_episodes = 0; s_eplen = 0; s_rows = 0; s_epreward = 0.0; }
  __syncthreads();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < eng.n_envs) {
    const int A = eng.n_actions, D = eng.obs_dim, R = eng.ring_rows;
    const int n_out = r.head_out[r.n_head - 1];
    const uint64_t g = eng.state->vec_steps;
    float q[SRLX_MAX_ACTIONS];
    dueling_combine(r.roll_act[r.n_head - 1] + (size_t)e * n_out, A, r.dueling, q);
    if (eng.dbg_q) for (int a = 0; a < A; ++a) eng.dbg_q[(size_t)e * A + a] = q[a];
    const double eps = training ? eng.epsilon : r.test_epsilon;
    float qmax = q[0];
    for (int a = 1; a < A; ++a) qmax = fmaxf(qmax, q[a]);
    int nmax = 0;
    for (int a = 0; a < A; ++a) nmax += (q[a] == qmax);
    double probs[SRLX_MAX_ACTIONS], total = 0.0;
    for (int a = 0; a < A; ++a) {
      double pr = __ddiv_rn(eps, (double)A);
      if (q[a] == qmax) pr = __dadd_rn(pr, __ddiv_rn(__dsub_rn(1.0, eps), (double)nmax));
      probs[a] = pr;
      total = __dadd_rn(total, pr);
    }
    const uint4 w = philox(eng.seed, STREAM_POLICY, (uint32_t)e, (uint32_t)g, (uint32_t)(g >> 32));
    const double rr = __dmul_rn(u01_f64(w.x, w.y), total);
    int action = A - 1;
    double num = 0.0;
    for (int a = 0; a < A; ++a) {
      num = __dadd_rn(num, probs[a]);
      if (rr <= num) { action = a; break; }
    }
    if (eng.dbg_action) eng.dbg_action[e] = action;
    double* st = eng.env_state + (size_t)e * 4;
    bool terminated = false;
    const double rew = env_step(eng, (uint32_t)e, g, action, st, terminated);
    const int step_num = eng.env_step_num[e] + 1;
    eng.env_step_num[e] = step_num;
    bool truncated = step_num >= eng.trunc_limit;
    if (eng.trunc_overrides_term) terminated = terminated && !truncated;
    else truncated = truncated && !terminated;
    const bool done = terminated || truncated;
    const double ep_reward = eng.env_ep_reward[e] + rew;
    eng.env_ep_reward[e] = ep_reward;
    if (training) {
      const uint32_t c0 = r.cursor[e];
      float nobs[SRLX_MAX_OBS];
      env_obs(eng, st, nobs);
      long long slot = ring_slot(r, c0, e);
      for (int d = 0; d < D; ++d) r.ring_next_obs[slot * D + d] = nobs[d];
      r.ring_action[slot] = action;
      r.ring_prob[slot] = probs[action];
      r.ring_reward[slot] = (rew + eng.reward_shift) * eng.reward_scale;  // worker_run.py:348
      r.ring_done[slot] = terminated ? 1 : 0;
      uint32_t n_new = 1;
      if (done) {
        for (int m = 1; m < r.seq_len; ++m) {
          slot = ring_slot(r, (long long)c0 + m, e);
          for (int d = 0; d < D; ++d) {
            r.ring_obs[slot * D + d] = (m == 1) ? nobs[d] : 0.f;
            r.ring_next_obs[slot * D + d] = 0.f;
          }
          const uint4 pw = philox(eng.seed, STREAM_PAD_ACTION, (uint32_t)e, c0 + (uint32_t)m, R2D2_PAD_TAIL);
          r.ring_action[slot] = (int)u_below(pw.x, (uint32_t)A);
          r.ring_prob[slot] = 1.0 / (double)A;
          r.ring_reward[slot] = 0.0;
          r.ring_done[slot] = 1;
          r.ring_tstep[slot] = step_num - 1 + m;
        }
        n_new = (uint32_t)r.seq_len;
      }
      r.cursor[e] = c0 + n_new;
      r.new_c0[e] = c0;
      r.new_n[e] = n_new;
      const unsigned long long before = c0 < (uint32_t)R ? c0 : (uint32_t)R, after = (c0 + n_new) < (uint32_t)R ? (c0 + n_new) : (uint32_t)R;
      if (after > before) atomicAdd(&s_rows, after - before);
      atomicMax((unsigned long long*)&eng.state->reserved[0], (unsigned long long)(c0 + n_new));
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
  if (threadIdx.x == 0) {
    if (s_episodes) {
      atomicAdd((unsigned long long*)&eng.state->episode_count, s_episodes);
      atomicAdd((unsigned long long*)&eng.state->episode_len_sum, s_eplen);
      atomicAdd(&eng.state->episode_reward_sum, s_epreward);
    }
    if (s_rows) atomicAdd((unsigned long long*)&eng.state->mem_size, s_rows);
  }
}

__global__ void r2d2_step_count_kernel(srlx_state* st, int E) {
  st->vec_steps += 1;
  st->total_step += (uint64_t)E;
}

// ProportionalMemory.add for the rows of one vector step (proportional_memory.py:120-129): new leaves take max_priority; the anchor
// whose window the new row cuts (pos - R + W - 1; at a column's first wrap all of 1 .. W - 1) drops to 0 so that the sampler's
// zero-priority rejection skips it.  A vector step touches a few leaves per env copy -- thousands per step -- so they are set in
// bulk: every touched leaf is written, then the touched paths are rebuilt bottom-up, one level per block-wide barrier, every
// ancestor as left child + right child (the value the reference's node holds up to the rounding of its `+= change` chain; the trainer's
// priority updates keep the reference's sequential association, tree_update_batch).  The result is a pure function of the leaves:
// deterministic, no atomics.  (In a tree whose leaves sit on two depths a node can be rebuilt once before its deeper child is final;
// the deeper path rebuilds it again one level later.)
__global__ void __launch_bounds__(1024) r2d2_add_kernel(const __grid_constant__ srlx_r2d2 r) {
  // cooperative launch over a few CTAs: CTA 0 builds the entry list and writes the leaves, then all CTAs rebuild the touched paths,
  // one level per grid barrier (bar[4] = arrival counter, bar[5] = number of entries; zeroed by the host before the launch)
  __shared__ int s_scan[1024];
  __shared__ int s_base;
  const srlx_engine& eng = r.env;
  const int E = eng.n_envs, R = eng.ring_rows, W = r.burnin + r.seq_len, tid = threadIdx.x;
  const long long cap = (long long)R * E;
  double* tree = eng.tree;
  unsigned* bar = r.bar + 4;
  volatile unsigned* n_entries = r.bar + 5;
  if (blockIdx.x == 0) {
    const double maxp = eng.state->max_priority;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int e0 = 0; e0 < E; e0 += 1024) {
      const int e = e0 + tid;
      int cnt = 0;
      uint32_t c0 = 0, n = 0;
      if (e < E) {
        c0 = r.new_c0[e];
        n = r.new_n[e];
        for (uint32_t j = 0; j < n; ++j) {
          const long long pos = (long long)c0 + j;
          cnt += 1 + (pos >= R ? (W - 1) - (pos == R ? 1 : W - 1) + 1 : 0);
        }
      }
      s_scan[tid] = cnt;
      __syncthreads();
      for (int off = 1; off < 1024; off <<= 1) {  // inclusive Hillis-Steele scan
        const int v = tid >= off ? s_scan[tid - off] : 0;
        __syncthreads();
        s_scan[tid] += v;
        __syncthreads();
      }
      int o = s_base + s_scan[tid] - cnt;
      for (uint32_t j = 0; j < n; ++j) {
        const long long pos = (long long)c0 + j;
        if (pos >= R) {  // the column's first wrap cuts every anchor that still reaches back to row 0; later rows cut one anchor each
          for (int a = (pos == R ? 1 : W - 1); a <= W - 1; ++a) {
            const long long idx = ring_leaf(r, pos - R + a, e) + cap - 1;
            r.add_idx[o++] = idx;
            __stcg(tree + idx, 0.0);
          }
        }
        const long long idx = ring_leaf(r, pos, e) + cap - 1;
        r.add_idx[o++] = idx;
        __stcg(tree + idx, maxp);
      }
      __syncthreads();
      if (tid == 1023) s_base += s_scan[1023];
      __syncthreads();
    }
    if (tid == 0) *n_entries = (unsigned)s_base;
  }
  unsigned phase = 1;
  grid_arrive(bar);
  grid_wait(r.bar, bar, phase * gridDim.x);
  const int total = (int)*n_entries;
  int depth = 0;
  while (((2 * cap - 1) >> (depth + 1)) > 0) ++depth;  // levels above the deepest leaf
  // entries of one env are neighbours in the list and (env-major leaves) in the tree: an entry whose ancestor at this level equals its
  // list predecessor's leaves the node to that entry; the survivors go four at a time so that their loads overlap
  const int nthr = gridDim.x * 1024, gtid = blockIdx.x * 1024 + tid;
  for (int lv = 1; lv <= depth; ++lv) {
    for (int i0 = gtid; i0 < total; i0 += 4 * nthr) {
      uint64_t node[4];
      double a[4], b[4];
      bool on[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * nthr;
        on[j] = false;
        if (i < total) {
          const uint64_t ip1 = (uint64_t)__ldcg(r.add_idx + i) + 1;
          if ((ip1 >> lv) != 0) {
            node[j] = (ip1 >> lv) - 1;
            on[j] = i == 0 || (((uint64_t)__ldcg(r.add_idx + i - 1) + 1) >> lv) != (ip1 >> lv);
          }
        }
        if (on[j]) { a[j] = __ldcg(tree + 2 * node[j] + 1); b[j] = __ldcg(tree + 2 * node[j] + 2); }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (on[j]) __stcg(tree + node[j], a[j] + b[j]);
    }
    ++phase;
    grid_arrive(bar);
    grid_wait(r.bar, bar, phase * gridDim.x);
  }
}

// ---- trainer: sample --------------------------------------------------------------------------------------------------------------------
// ReplayBuffer.sample (replay_buffer.py:34-36: random.sample, distinct items): uniform (row, env) draws, rejected when the row is not a
// valid anchor or was already picked, accepted in attempt order
__global__ void __launch_bounds__(256) r2d2_sample_uniform_kernel(const __grid_constant__ srlx_r2d2 r, const Gate gate) {
  if (gate.closed()) return;
  const srlx_engine& eng = r.env;
  __shared__ long long s_cand[256];
  __shared__ int s_n;
  const int B = eng.batch_size, E = eng.n_envs, R = eng.ring_rows, tid = threadIdx.x;
  const uint64_t step = eng.state->train_count;
  const unsigned long long maxc = eng.state->reserved[0];
  const uint32_t rows = (uint32_t)(maxc < (unsigned long long)R ? maxc : (unsigned long long)R);
  if (tid == 0) s_n = 0;
  __syncthreads();
  for (int round = 0; round < 256 && s_n < B; ++round) {
    const uint32_t a = (uint32_t)round * 256u + (uint32_t)tid;
    const uint4 w = philox(eng.seed, STREAM_UNIFORM_SAMPLE, a, (uint32_t)step, (uint32_t)(step >> 32));
    const int row = (int)u_below(w.x, rows), e = (int)u_below(w.y, (uint32_t)E);
    const long long cur = r.cursor[e];
    const long long pos = ring_pos_of_row(r, row, cur);
    s_cand[tid] = ring_anchor_valid(r, pos, cur) ? (long long)row * E + e : -1;
    __syncthreads();
    if (tid == 0) {
      int n = s_n;
      for (int i = 0; i < 256 && n < B; ++i) {
        const long long c = s_cand[i];
        if (c < 0) continue;
        bool dup = false;
        for (int j = 0; j < n; ++j) dup |= (r.sel[j] == c);
        if (!dup) { r.sel[n] = c; r.weights[n] = 1.0f; ++n; }
      }
      s_n = n;
    }
    __syncthreads();
  }
}

// ProportionalMemory.sample (proportional_memory.py:131-169) on the tree over all R*E rows
__global__ void __launch_bounds__(1024) r2d2_sample_per_kernel(const __grid_constant__ srlx_r2d2 r, const Gate gate) {
  if (gate.closed()) return;
  const srlx_engine& eng = r.env;
  __shared__ int64_t s_idx[SRLX_MAX_BATCH];
  __shared__ double s_pri[SRLX_MAX_BATCH];
  __shared__ double s_tmp[SRLX_MAX_BATCH];
  __shared__ float s_w[SRLX_MAX_BATCH];
  __shared__ unsigned long long retries;
  if (threadIdx.x == 0) retries = 0;
  __syncthreads();
  const long long cap = (long long)eng.ring_rows * eng.n_envs;
  const uint64_t step = eng.state->train_count;
  const double total = __ldcg(eng.tree);
  // PriorityReplayBuffer.step is the train_count the PREVIOUS update handed to memory.update (priority_replay_buffer.py:228-250)
  const double beta_step = step > 0 ? (double)(step - 1) : 0.0;
  double beta = eng.per_beta_initial + (1.0 - eng.per_beta_initial) * beta_step / eng.per_beta_steps;
  if (beta > 1.0) beta = 1.0;
  per_sample_block(eng.tree, 2 * cap - 1, total, eng.batch_size, eng.seed, step, nullptr, 10000, eng.has_duplicate, s_idx, s_pri, s_tmp, &retries);
  per_weights_block(total, (double)eng.state->mem_size, beta, eng.batch_size, s_pri, s_tmp, s_w);
  for (int i = threadIdx.x; i < eng.batch_size; i += blockDim.x) { r.sel[i] = s_idx[i]; r.weights[i] = s_w[i]; }
  if (threadIdx.x == 0) eng.state->sample_retries += retries;
}

// the batch as _train_on_batches lays it out (r2d2.py:109-133): W + 1 states per item (padding in front of an episode's first step
// rebuilt by index: dummy state 0, random action, probability 1/A, reward 0, not done, zero LSTM state), time-major for the unroll
__global__ void __launch_bounds__(128) r2d2_gather_kernel(const __grid_constant__ srlx_r2d2 r, const Gate gate) {
  if (gate.closed()) return;
  const srlx_engine& eng = r.env;
  const int b = blockIdx.x, B = eng.batch_size, E = eng.n_envs, D = eng.obs_dim, u = r.lstm_units, A = eng.n_actions;
  const int W = r.burnin + r.seq_len, S = r.seq_len, K = D + u + 1, tid = threadIdx.x;
  const long long cap = (long long)eng.ring_rows * E;
  long long slot = r.sel[b];
  if (eng.mem_kind == SRLX_MEM_PROPORTIONAL) {  // tree leaves are env-major (ring_leaf), ring slots row-major
    const long long leaf = slot - (cap - 1);
    slot = (leaf % eng.ring_rows) * E + leaf / eng.ring_rows;
  }
  const int e = (int)(slot % E), row = (int)(slot / E);
  const long long cur = r.cursor[e];
  const long long p = ring_pos_of_row(r, row, cur);
  const long long ep_start = p - r.ring_tstep[slot];
  const size_t zx = (size_t)(W + 2) * B * K, zc = (size_t)(W + 2) * B * u;
  for (int w = tid; w < (W + 1) * D; w += blockDim.x) {
    const int i = w / D, d = w - i * D;
    float v;
    if (i == W) v = r.ring_next_obs[slot * D + d];
    else {
      const long long pos = p - (W - 1) + i;
      v = pos < ep_start ? 0.f : r.ring_obs[ring_slot(r, pos, e) * D + d];
    }
    const size_t o = ((size_t)i * B + b) * K + d;
    r.xh[o] = v;
    r.xh[zx + o] = v;
  }
  const long long pos0 = p - (W - 1);
  const bool pre0 = pos0 < ep_start;
  const long long s0 = pre0 ? 0 : ring_slot(r, pos0, e);
  for (int j = tid; j < u; j += blockDim.x) {
    const float h = pre0 ? 0.f : r.ring_h[s0 * u + j], c = pre0 ? 0.f : r.ring_c[s0 * u + j];
    r.xh[(size_t)b * K + D + j] = h;
    r.xh[zx + (size_t)b * K + D + j] = h;
    r.cbuf[(size_t)b * u + j] = c;
    r.cbuf[zc + (size_t)b * u + j] = c;
  }
  for (int k = tid; k < S; k += blockDim.x) {
    const long long pos = p - (S - 1) + k;
    const size_t o = (size_t)b * S + k;
    if (pos < ep_start) {
      const uint4 pw = philox(eng.seed, STREAM_PAD_ACTION, (uint32_t)e, (uint32_t)pos, R2D2_PAD_PRE);
      r.b_actions[o] = (int)u_below(pw.x, (uint32_t)A);
      r.b_mu[o] = 1.0 / (double)A;
      r.b_rewards[o] = 0.0;
      r.b_dones[o] = 0;
    } else {
      const long long s = ring_slot(r, pos, e);
      r.b_actions[o] = r.ring_action[s];
      r.b_mu[o] = r.ring_prob[s];
      r.b_rewards[o] = r.ring_reward[s];
      r.b_dones[o] = r.ring_done[s];
    }
  }
}

// head output rows (t, b) -> Q [z][b][t][a] for the reference's per-sequence loop
__global__ void r2d2_combine_kernel(const __grid_constant__ srlx_r2d2 r, const Gate gate) {
  if (gate.closed()) return;
  const int B = r.env.batch_size, S1 = r.seq_len + 1, A = r.env.n_actions, n_out = r.head_out[r.n_head - 1];
  const int rows = S1 * B;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * rows) return;
  const int z = i / rows, row = i - z * rows, t = row / B, b = row - t * B;
  float q[SRLX_MAX_ACTIONS];
  dueling_combine(r.act[r.n_head - 1] + ((size_t)z * rows + row) * n_out, A, r.dueling, q);
  float* out = r.q + (((size_t)z * B + b) * S1 + t) * A;
  for (int a = 0; a < A; ++a) out[a] = q[a];
}

// keras.losses.Huber()(target * w, q_onehot * w) (r2d2.py:206-209): mean over (B, seq_len); gradient wrt the head outputs
__global__ void __launch_bounds__(256) r2d2_loss_kernel(const __grid_constant__ srlx_r2d2 r, const Gate gate) {
  if (gate.closed()) return;
  __shared__ double s_sum[256];
  const int B = r.env.batch_size, S = r.seq_len, S1 = S + 1, A = r.env.n_actions, n_out = r.head_out[r.n_head - 1], tid = threadIdx.x;
  const float inv_n = 1.0f / (float)(B * S);
  double part = 0.0;
  for (int i = tid; i < B * S; i += blockDim.x) {
    const int t = i / B, b = i - t * B;  // row (t, b) of the head
    const int a = r.b_actions[(size_t)b * S + t];
    const float w = r.weights[b];
    const float* qrow = r.q + ((size_t)b * S1 + t) * A;
    const float y_pred = qrow[a] * w;
    const float y_true = (float)(r.b_target[(size_t)b * S + t] * (double)w);
    const float err = y_pred - y_true, ae = fabsf(err);
    part += (ae <= 1.0f) ? 0.5 * (double)err * (double)err : (double)ae - 0.5;
    const float gq = w * fminf(fmaxf(err, -1.0f), 1.0f) * inv_n;
    float* d = r.dact[r.n_head - 1] + (size_t)i * n_out;
    if (r.dueling == SRLX_DUEL_NONE) {
      for (int j = 0; j < A; ++j) d[j] = (j == a) ? gq : 0.f;
    } else {
      d[0] = gq;
      int amax = 0;
      if (r.dueling == SRLX_DUEL_MAX) {
        const float* o = r.act[r.n_head - 1] + (size_t)i * n_out;
        for (int j = 1; j < A; ++j) if (o[1 + j] > o[1 + amax]) amax = j;
      }
      for (int j = 0; j < A; ++j) {
        float v = (j == a) ? gq : 0.f;
        if (r.dueling == SRLX_DUEL_AVERAGE) v -= gq / (float)A;
        else if (r.dueling == SRLX_DUEL_MAX && j == amax) v -= gq;
        d[1 + j] = v;
      }
    }
  }
  s_sum[tid] = part;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) s_sum[tid] += s_sum[tid + s];
    __syncthreads();
  }
  if (tid == 0) {
    const double loss = s_sum[0] / (double)(B * S);
    r.env.state->last_loss = loss;
    r.env.state->loss_sum += loss;
  }
}

// keras Adam (optimizers.Adam defaults: beta 0.9 / 0.999, epsilon 1e-7 outside the root): alpha = lr sqrt(1 - b2^t) / (1 - b1^t)
__global__ void r2d2_adam_kernel(const __grid_constant__ srlx_r2d2 r, const Gate gate) {
  if (gate.closed()) return;
  const srlx_engine& eng = r.env;
  const double t1 = (double)(eng.state->adam_step + 1);
  const float alpha = (float)(eng.lr * sqrt(1.0 - pow(eng.adam_beta2, t1)) / (1.0 - pow(eng.adam_beta1, t1)));
  const float b1 = (float)eng.adam_beta1, b2 = (float)eng.adam_beta2, eps = (float)eng.adam_eps;
  const int lo = r.head_off[r.n_head - 1], ld = r.head_k[r.n_head - 1] + 1, H = r.duel_hidden;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < r.n_params; i += gridDim.x * blockDim.x) {
    if (H > 0 && i >= lo) {  // structural zeros of the dueling output layer: V reads [0, H), the advantages [H, 2H)
      const int row = (i - lo) / ld, col = (i - lo) - row * ld;
      if (row == 0 ? (col >= H && col < 2 * H) : col < H) continue;
    }
    const float g = r.grads[i];
    const float m = r.adam_m[i] + (1.f - b1) * (g - r.adam_m[i]);
    const float v = r.adam_v[i] + (1.f - b2) * (g * g - r.adam_v[i]);
    r.adam_m[i] = m;
    r.adam_v[i] = v;
    r.params[i] -= alpha * m / (sqrtf(v) + eps);
  }
}

// memory.update(update_args, td_errors, train_count) (r2d2.py:97; proportional_memory.py:171-177 on the mean TD error of a sequence)
__global__ void __launch_bounds__(1024) r2d2_priority_kernel(const __grid_constant__ srlx_r2d2 r, const Gate gate) {
  if (gate.closed()) return;
  extern __shared__ __align__(16) unsigned char tree_smem[];
  TreeHashScratch* hs = reinterpret_cast<TreeHashScratch*>(tree_smem);
  __shared__ int64_t s_idx[SRLX_MAX_BATCH];
  __shared__ double s_pri[SRLX_MAX_BATCH];
  const srlx_engine& eng = r.env;
  const int B = eng.batch_size;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    s_idx[i] = r.sel[i];
    s_pri[i] = pow(fabs(r.b_tdmean[i]) + eng.per_epsilon, eng.per_alpha);
  }
  __syncthreads();
  tree_update_batch(eng.tree, s_idx, s_pri, B, hs);
  if (threadIdx.x == 0) {
    double mp = eng.state->max_priority;
    for (int i = 0; i < B; ++i) mp = (mp < s_pri[i]) ? s_pri[i] : mp;
    eng.state->max_priority = mp;
  }
}

// hard target sync when train_count % interval == 0, evaluated BEFORE the increment (r2d2.py:100-104)
__global__ void r2d2_sync_kernel(const __grid_constant__ srlx_r2d2 r, const Gate gate) {
  if (gate.closed()) return;
  if (r.env.state->train_count % (uint64_t)r.env.target_update_interval != 0) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < r.n_params; i += gridDim.x * blockDim.x) r.target[i] = r.params[i];
}

__global__ void r2d2_finish_kernel(const __grid_constant__ srlx_r2d2 r, const Gate gate) {
  if (gate.closed()) return;
  srlx_state* st = r.env.state;
  if (st->train_count % (uint64_t)r.env.target_update_interval == 0) st->sync_count += 1;
  st->train_count += 1;
  st->adam_step += 1;
}

// copy rows [n][w] between strided buffers (srlx_r2d2_forward staging)
__global__ void r2d2_copy_rows_kernel(const float* __restrict__ src, long long ld_s, float* __restrict__ dst, long long ld_d, long long n, int w) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * w; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / w;
    const int j = (int)(i - row * w);
    dst[row * ld_d + j] = src[row * ld_s + j];
  }
}
__global__ void r2d2_combine_rows_kernel(const float* __restrict__ o, int n_out, int A, int dueling, long long n, float* __restrict__ q) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float qq[SRLX_MAX_ACTIONS];
  dueling_combine(o + i * n_out, A, dueling, qq);
  for (int a = 0; a < A; ++a) q[i * A + a] = qq[a];
}


// ---- host side ------------------------------------------------------------------------------------------------------------------------------
static int r2d2_check(const srlx_r2d2* r) {
  SRLX_REQUIRE(r != nullptr, "r2d2 is NULL");
  const srlx_engine& e = r->env;
  SRLX_REQUIRE(e.n_envs >= 1 && e.obs_dim >= 1 && e.obs_dim <= 4, "n_envs / obs_dim out of range");
  SRLX_REQUIRE(e.env_id == SRLX_ENV_GRID || e.env_id == SRLX_ENV_CARTPOLE || e.env_id == SRLX_ENV_PENDULUM, "unknown env_id %d", e.env_id);
  SRLX_REQUIRE(e.n_actions >= 1 && e.n_actions <= SRLX_MAX_ACTIONS, "n_actions out of range");
  SRLX_REQUIRE(r->lstm_units >= 1 && r->burnin >= 0 && r->seq_len >= 1 && r->seq_len <= 128, "lstm_units / burnin / seq_len (<= 128) out of range");
  SRLX_REQUIRE(r->n_head >= 1 && r->n_head <= SRLX_MAX_LAYERS, "head layers out of range");
  SRLX_REQUIRE(r->head_k[0] == r->lstm_units, "the first head layer reads the LSTM output");
  SRLX_REQUIRE(r->head_out[r->n_head - 1] == (r->dueling == SRLX_DUEL_NONE ? e.n_actions : 1 + e.n_actions), "output layer width does not match");
  SRLX_REQUIRE(e.state && e.env_state && e.env_step_num && e.env_episode && e.env_ep_reward && e.env_needs_reset, "env buffer pointer is NULL");
  SRLX_REQUIRE(r->params && r->target, "params / target pointer is NULL");
  return 0;
}

// one LSTM step + head on n rows: xh_in [n][K] -> h into h_out (row stride ld_h, which must be followed by a 1 at column u: the head
// reads [h | 1]), c in place, head activations -> acts[l]
static int r2d2_net_step(const srlx_r2d2* r, const float* W, const float* xh_in, float* h_out, long long ld_h, float* c, float* const* acts,
                         int n, cudaStream_t s, float* pre = nullptr) {
  const int D = r->env.obs_dim, u = r->lstm_units, K = D + u + 1;
  if (n > 64 && pre != nullptr) {  // many rows: gate pre-activations on the tiled GEMM, then the cell update
    GemmP g{};
    g.A = xh_in; g.sa_m = K; g.sa_k = 1;
    g.B = W + r->lstm_off; g.sb_k = 1; g.sb_n = K;
    g.C = pre; g.ldc = 4 * u; g.M = n; g.N = 4 * u; g.K = K; g.gate = Gate{nullptr, 0};
    launch_gemm(g, 1, s);
    const long long nb = ((long long)n * u + 255) / 256;
    lstm_pointwise_kernel<<<(unsigned)(nb < 2368 ? nb : 2368), 256, 0, s>>>(pre, c, h_out, ld_h, n, u);
    count_launch();
  } else {
    LstmFwdP lp{};
    lp.xh = xh_in; lp.z_xh = 0; lp.W0 = W + r->lstm_off; lp.W1 = nullptr; lp.c_in = c; lp.c_out = c; lp.z_c = 0;
    lp.h_out = h_out; lp.ld_h = ld_h; lp.z_h = 0; lp.gates = nullptr; lp.M = n; lp.u = u; lp.K = K; lp.gate = Gate{nullptr, 0};
    if (n <= 32) lstm_fwd_kernel<32, 2><<<dim3((4 * u + 31) / 32, (n + 31) / 32, 1), 128, 0, s>>>(lp);
    else lstm_fwd_kernel<64, 4><<<dim3((4 * u + 31) / 32, (n + 63) / 64, 1), 128, 0, s>>>(lp);
    count_launch();
  }
  const float* in = h_out;
  long long ld_in = ld_h;
  for (int l = 0; l < r->n_head; ++l) {
    const bool last = l == r->n_head - 1;
    GemmP g{};
    g.A = in; g.sa_m = ld_in; g.sa_k = 1;
    g.B = W + r->head_off[l]; g.sb_k = 1; g.sb_n = r->head_k[l] + 1;
    g.C = acts[l]; g.ldc = last ? r->head_out[l] : r->head_out[l] + 1;
    g.M = n; g.N = r->head_out[l]; g.K = r->head_k[l] + 1; g.relu = last ? 0 : 1;
    g.gate = Gate{nullptr, 0};
    launch_gemm(g, 1, s);
    in = acts[l];
    ld_in = r->head_out[l] + 1;
  }
  return 0;
}

}  // namespace srlx

using namespace srlx;

extern "C" size_t srlx_sizeof_r2d2(void) { return sizeof(srlx_r2d2); }

extern "C" int srlx_sgemm(const float* a_dev, long long sa_m, long long sa_k, const float* b_dev, long long sb_k, long long sb_n, float* c_dev,
                          long long ldc, int M, int N, int K, int relu, int accumulate, uintptr_t cuda_stream) {
  SRLX_REQUIRE(a_dev && b_dev && c_dev, "srlx_sgemm: NULL buffer");
  SRLX_REQUIRE(M >= 0 && N >= 0 && K >= 0, "srlx_sgemm: negative size");
  GemmP g{};
  g.A = a_dev; g.sa_m = sa_m; g.sa_k = sa_k; g.B = b_dev; g.sb_k = sb_k; g.sb_n = sb_n; g.C = c_dev; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.relu = relu; g.accumulate = accumulate; g.gate = Gate{nullptr, 0};
  launch_gemm(g, 1, (cudaStream_t)cuda_stream);
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int srlx_r2d2_vec_step(const srlx_r2d2* r, int training, uintptr_t cuda_stream) {
  if (int rc = r2d2_check(r)) return rc;
  const srlx_engine& eng = r->env;
  cudaStream_t s = (cudaStream_t)cuda_stream;
  SRLX_REQUIRE(r->roll_xh && r->roll_h && r->roll_c && r->roll_reset, "rollout workspace pointer is NULL");
  for (int l = 0; l < r->n_head; ++l) SRLX_REQUIRE(r->roll_act[l] != nullptr, "roll_act[%d] is NULL", l);
  const bool per = eng.mem_kind == SRLX_MEM_PROPORTIONAL;
  if (training) {
    SRLX_REQUIRE(r->cursor && r->ring_obs && r->ring_next_obs && r->ring_action && r->ring_prob && r->ring_reward && r->ring_done &&
                 r->ring_tstep && r->ring_h && r->ring_c && r->new_c0 && r->new_n, "replay ring pointer is NULL");
    SRLX_REQUIRE(eng.ring_rows >= 2 * (r->burnin + r->seq_len), "ring_rows must be >= 2 * (burnin + seq_len)");
    SRLX_REQUIRE(!per || (eng.tree && r->add_idx), "proportional memory needs tree / add_idx");
  }
  const int E = eng.n_envs, u = r->lstm_units;
  r2d2_pre_kernel<<<(E + 127) / 128, 128, 0, s>>>(*r, training);
  const long long nh = (long long)E * u;
  const long long hb = (nh + 255) / 256;
  r2d2_hidden_kernel<<<(unsigned)(hb < 2368 ? hb : 2368), 256, 0, s>>>(*r, training);
  count_launch(2);
  r2d2_net_step(r, r->params, r->roll_xh, r->roll_h, u + 1, r->roll_c, r->roll_act, E, s, r->roll_gates);
  r2d2_act_kernel<<<(E + 127) / 128, 128, 0, s>>>(*r, training);
  count_launch();
  if (training && per) {
    SRLX_REQUIRE(r->bar != nullptr, "proportional replay add needs the barrier words (bar)");
    SRLX_CHECK_CUDA(cudaMemsetAsync(r->bar + 4, 0, 2 * sizeof(unsigned), s));
    void* args[] = {(void*)r};
    const int n_cta = E >= 1024 ? 32 : (E >= 64 ? 8 : 1);
    SRLX_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)r2d2_add_kernel, dim3(n_cta), dim3(1024), args, 0, s));
    count_launch();
  }
  r2d2_step_count_kernel<<<1, 1, 0, s>>>(eng.state, E);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int srlx_r2d2_forward(const srlx_r2d2* r, int use_target, const float* obs_dev, const float* h_dev, const float* c_dev, uint32_t n,
                                 float* q_out_dev, float* h_out_dev, float* c_out_dev, uintptr_t cuda_stream) {
  if (int rc = r2d2_check(r)) return rc;
  const srlx_engine& eng = r->env;
  SRLX_REQUIRE(obs_dev && h_dev && c_dev && q_out_dev && h_out_dev && c_out_dev, "srlx_r2d2_forward: NULL buffer");
  SRLX_REQUIRE(n >= 1 && n <= (uint32_t)eng.batch_size, "srlx_r2d2_forward: n = %u exceeds batch_size %d (the call runs in the learner workspace)", n, eng.batch_size);
  SRLX_REQUIRE(r->xh && r->cbuf, "learner workspace pointer is NULL");
  for (int l = 0; l < r->n_head; ++l) SRLX_REQUIRE(r->act[l] != nullptr, "act[%d] is NULL", l);
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const int D = eng.obs_dim, u = r->lstm_units, K = D + u + 1, B = eng.batch_size;
  float* x0 = r->xh;                      // time row 0: input
  float* x1 = r->xh + (size_t)B * K;      // time row 1: receives h (followed by the row's trailing 1)
  const unsigned nb = (unsigned)(((long long)n * (u > D ? u : D) + 255) / 256);
  r2d2_copy_rows_kernel<<<nb, 256, 0, s>>>(obs_dev, D, x0, K, n, D);
  r2d2_copy_rows_kernel<<<nb, 256, 0, s>>>(h_dev, u, x0 + D, K, n, u);
  r2d2_copy_rows_kernel<<<nb, 256, 0, s>>>(c_dev, u, r->cbuf, u, n, u);
  count_launch(3);
  r2d2_net_step(r, use_target ? r->target : r->params, x0, x1 + D, K, r->cbuf, r->act, (int)n, s);
  r2d2_copy_rows_kernel<<<nb, 256, 0, s>>>(x1 + D, K, h_out_dev, u, n, u);
  r2d2_copy_rows_kernel<<<nb, 256, 0, s>>>(r->cbuf, u, c_out_dev, u, n, u);
  const int n_out = r->head_out[r->n_head - 1];
  r2d2_combine_rows_kernel<<<(n + 127) / 128, 128, 0, s>>>(r->act[r->n_head - 1], n_out, eng.n_actions, r->dueling, n, q_out_dev);
  count_launch(3);
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int srlx_r2d2_learn(const srlx_r2d2* r, uint32_t n_updates, uintptr_t cuda_stream) {
  return srlx_r2d2_learn_phase(r, n_updates, 3, cuda_stream);
}

extern "C" int srlx_r2d2_learn_phase(const srlx_r2d2* r, uint32_t n_updates, int phases, uintptr_t cuda_stream) {
  if (int rc = r2d2_check(r)) return rc;
  SRLX_REQUIRE(phases >= 1 && phases <= 3, "srlx_r2d2_learn_phase: phases must be 1 (gradients), 2 (apply) or 3 (both)");
  SRLX_REQUIRE(phases == 3 || n_updates == 1, "srlx_r2d2_learn_phase: one update per call when the phases are split");
  const srlx_engine& eng = r->env;
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const bool per = eng.mem_kind == SRLX_MEM_PROPORTIONAL;
  SRLX_REQUIRE(eng.batch_size >= 1 && eng.batch_size <= SRLX_MAX_BATCH, "batch_size out of range");
  SRLX_REQUIRE(eng.target_update_interval >= 1, "target_update_interval must be >= 1");
  SRLX_REQUIRE(r->cursor && r->ring_obs && r->ring_next_obs && r->ring_action && r->ring_prob && r->ring_reward && r->ring_done &&
               r->ring_tstep && r->ring_h && r->ring_c, "replay ring pointer is NULL");
  SRLX_REQUIRE(r->adam_m && r->adam_v && r->grads && r->xh && r->cbuf && r->gates && r->dgates && r->dc && r->dh && r->q && r->sel && r->weights &&
               r->b_actions && r->b_mu && r->b_rewards && r->b_dones && r->b_target && r->b_tdmean && r->b_tdkind, "learner workspace pointer is NULL");
  for (int l = 0; l < r->n_head; ++l) SRLX_REQUIRE(r->act[l] && r->dact[l], "act / dact[%d] is NULL", l);
  SRLX_REQUIRE(!per || eng.tree, "proportional memory needs the tree");
  const int B = eng.batch_size, D = eng.obs_dim, u = r->lstm_units, K = D + u + 1, S = r->seq_len, W = r->burnin + S, A = eng.n_actions;
  const int rows1 = (S + 1) * B, rowsS = S * B;
  const size_t zx = (size_t)(W + 2) * B * K, zc = (size_t)(W + 2) * B * u;
  const Gate gate{eng.state, (unsigned long long)eng.warmup_size};
  // the persistent unroll needs every CTA resident at once (cooperative launch) and the per-thread weight slices in registers
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    SRLX_CHECK_CUDA(cudaGetDevice(&dev));
    SRLX_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  const bool persistent = !r->no_persistent && r->bar != nullptr && B <= 64 && u <= 512 && D + 1 <= 32 && 2 * ((u + 7) / 8) <= n_sm && (u + 3) / 4 <= n_sm;
  if (per) SRLX_CHECK_CUDA(cudaFuncSetAttribute(r2d2_priority_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TreeHashScratch)));
  for (uint32_t it = 0; it < n_updates; ++it) {
    if (phases & 1) {
    // memory.sample (r2d2.py:91) + the batch layout of _train_on_batches (:109-133)
    if (per) r2d2_sample_per_kernel<<<1, 1024, 0, s>>>(*r, gate);
    else r2d2_sample_uniform_kernel<<<1, 256, 0, s>>>(*r, gate);
    r2d2_gather_kernel<<<B, 128, 0, s>>>(*r, gate);
    count_launch(2);
    // burn-in + unroll of both networks (:136-150): step t reads [x_t | h_{t-1} | 1] in time row t and leaves h_t in time row t + 1
    if (persistent) {
      SRLX_CHECK_CUDA(cudaMemsetAsync(r->bar, 0, 4 * sizeof(unsigned), s));
      SeqFwdP fp{};
      fp.W0 = r->params + r->lstm_off; fp.W1 = r->target + r->lstm_off;
      fp.xh = r->xh; fp.z_xh = (long long)zx; fp.cbuf = r->cbuf; fp.z_c = (long long)zc; fp.gates = r->gates; fp.bar = r->bar;
      fp.B = B; fp.u = u; fp.D = D; fp.K = K; fp.T = W + 1; fp.cpn = (u + 7) / 8; fp.gate = gate;
      void* args[] = {&fp};
      const dim3 grid(2 * fp.cpn), block(256);
      const void* fn = B <= 32 ? (const void*)lstm_seq_fwd_kernel<32> : (const void*)lstm_seq_fwd_kernel<64>;
      const size_t smem = ((size_t)(B <= 32 ? 32 : 64) * kFwdXS + (size_t)32 * kFwdRedStride) * sizeof(float);
      SRLX_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      SRLX_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, grid, block, args, smem, s));
      count_launch();
    } else {
      for (int t = 0; t <= W; ++t) {
        LstmFwdP lp{};
        lp.xh = r->xh + (size_t)t * B * K; lp.z_xh = (long long)zx;
        lp.W0 = r->params + r->lstm_off; lp.W1 = r->target + r->lstm_off;
        lp.c_in = r->cbuf + (size_t)t * B * u; lp.c_out = r->cbuf + (size_t)(t + 1) * B * u; lp.z_c = (long long)zc;
        lp.h_out = r->xh + (size_t)(t + 1) * B * K + D; lp.ld_h = K; lp.z_h = (long long)zx;
        lp.gates = r->gates + (size_t)t * B * 4 * u; lp.M = B; lp.u = u; lp.K = K; lp.gate = gate;
        if (B <= 32) lstm_fwd_kernel<32, 2><<<dim3((4 * u + 31) / 32, (B + 31) / 32, 2), 128, 0, s>>>(lp);
        else lstm_fwd_kernel<64, 4><<<dim3((4 * u + 31) / 32, (B + 63) / 64, 2), 128, 0, s>>>(lp);
        count_launch();
      }
    }
    // hidden block on the S + 1 unrolled steps of both networks
    for (int l = 0; l < r->n_head; ++l) {
      const bool last = l == r->n_head - 1;
      const int ldo = last ? r->head_out[l] : r->head_out[l] + 1;
      GemmP g{};
      if (l == 0) { g.A = r->xh + (size_t)(r->burnin + 1) * B * K + D; g.sa_m = K; g.zA = (long long)zx; }
      else { g.A = r->act[l - 1]; g.sa_m = r->head_out[l - 1] + 1; g.zA = (long long)rows1 * (r->head_out[l - 1] + 1); }
      g.sa_k = 1;
      g.B = r->params + r->head_off[l]; g.B1 = r->target + r->head_off[l]; g.sb_k = 1; g.sb_n = r->head_k[l] + 1;
      g.C = r->act[l]; g.ldc = ldo; g.zC = (long long)rows1 * ldo;
      g.M = rows1; g.N = r->head_out[l]; g.K = r->head_k[l] + 1; g.relu = last ? 0 : 1; g.gate = gate;
      launch_gemm(g, 2, s);
    }
    r2d2_combine_kernel<<<(2 * rows1 + 127) / 128, 128, 0, s>>>(*r, gate);
    count_launch();
    // the reference's own per-sequence loop (:158-202), bit-exact (csrc/sequence_targets.cu)
    if (int rc = srlx_sequence_targets(r->q, r->q + (size_t)B * (S + 1) * A, r->b_actions, r->b_mu, r->b_rewards, r->b_dones, r->b_target,
                                       r->b_tdmean, r->b_tdkind, (uint32_t)B, (uint32_t)S, (uint32_t)A, eng.discount, eng.retrace_h,
                                       eng.enable_double_dqn, eng.enable_rescale, r->enable_retrace, cuda_stream)) return rc;
    r2d2_loss_kernel<<<1, 256, 0, s>>>(*r, gate);
    count_launch();
    // hidden block backward over the S trained steps (rows (t, b), t < S, are the first S * B rows of every activation buffer)
    for (int l = r->n_head - 1; l >= 0; --l) {
      const float* in = l == 0 ? r->xh + (size_t)(r->burnin + 1) * B * K + D : r->act[l - 1];
      const long long ld_in = l == 0 ? K : r->head_out[l - 1] + 1;
      GemmP gw{};  // dW_l = dO_l^T . [in_l | 1]
      gw.A = r->dact[l]; gw.sa_m = 1; gw.sa_k = r->head_out[l];
      gw.B = in; gw.sb_k = ld_in; gw.sb_n = 1;
      gw.C = r->grads + r->head_off[l]; gw.ldc = r->head_k[l] + 1;
      gw.M = r->head_out[l]; gw.N = r->head_k[l] + 1; gw.K = rowsS; gw.gate = gate;
      launch_gemm(gw, 1, s, r->gemm_ws, (size_t)r->gemm_ws_floats);
      GemmP gi{};  // d in_l = dO_l . W_l[:, :k_l], through the ReLU of the layer below
      gi.A = r->dact[l]; gi.sa_m = r->head_out[l]; gi.sa_k = 1;
      gi.B = r->params + r->head_off[l]; gi.sb_k = r->head_k[l] + 1; gi.sb_n = 1;
      gi.C = l == 0 ? r->dh : r->dact[l - 1]; gi.ldc = r->head_k[l];
      gi.M = rowsS; gi.N = r->head_k[l]; gi.K = r->head_out[l]; gi.gate = gate;
      if (l > 0) { gi.mask = r->act[l - 1]; gi.ldmask = r->head_out[l - 1] + 1; }
      launch_gemm(gi, 1, s);
    }
    // BPTT over the S trained steps; the burn-in state is a constant of the tape (:136-138)
    if (persistent) {
      SeqBwdP bp{};
      bp.W = r->params + r->lstm_off; bp.K = K; bp.D = D;
      bp.dh_head = r->dh; bp.gates = r->gates + (size_t)r->burnin * B * 4 * u; bp.cbuf = r->cbuf + (size_t)r->burnin * B * u;
      bp.dgates = r->dgates; bp.bar = r->bar; bp.B = B; bp.u = u; bp.S = S; bp.n_cta = (u + 3) / 4; bp.gate = gate;
      void* args[] = {&bp};
      const void* fn = B <= 32 ? (const void*)lstm_seq_bwd_kernel<32> : (const void*)lstm_seq_bwd_kernel<64>;
      const size_t smem = ((size_t)2 * (B <= 32 ? 32 : 64) * kBwdRow + (size_t)64 * kBwdRedStride) * sizeof(float);
      SRLX_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      SRLX_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(bp.n_cta), dim3(256), args, smem, s));
      count_launch();
    } else {
      for (int t = S - 1; t >= 0; --t) {
        LstmBwdP bp{};
        bp.dg_next = t == S - 1 ? nullptr : r->dgates + (size_t)(t + 1) * B * 4 * u;
        bp.W = r->params + r->lstm_off; bp.in = D; bp.K = K;
        bp.dh_head = r->dh + (size_t)t * B * u;
        bp.gates = r->gates + (size_t)(r->burnin + t) * B * 4 * u;
        bp.c_prev = r->cbuf + (size_t)(r->burnin + t) * B * u;
        bp.c_cur = r->cbuf + (size_t)(r->burnin + t + 1) * B * u;
        bp.dc = r->dc; bp.dg = r->dgates + (size_t)t * B * 4 * u; bp.M = B; bp.u = u; bp.first = t == S - 1; bp.gate = gate;
        if (B <= 32) lstm_bwd_kernel<32, 2><<<dim3((u + 31) / 32, (B + 31) / 32, 1), 128, 0, s>>>(bp);
        else lstm_bwd_kernel<64, 4><<<dim3((u + 31) / 32, (B + 63) / 64, 1), 128, 0, s>>>(bp);
        count_launch();
      }
    }
    {
      GemmP gw{};  // dW_lstm = sum_t dgates_t^T . [x_t | h_{t-1} | 1]
      gw.A = r->dgates; gw.sa_m = 1; gw.sa_k = 4LL * u;
      gw.B = r->xh + (size_t)r->burnin * B * K; gw.sb_k = K; gw.sb_n = 1;
      gw.C = r->grads + r->lstm_off; gw.ldc = K;
      gw.M = 4 * u; gw.N = K; gw.K = rowsS; gw.gate = gate;
      launch_gemm(gw, 1, s);
    }
    }  // phase 1: r->grads holds the gradient of this rank's batch
    if (!(phases & 2)) continue;
    r2d2_adam_kernel<<<(r->n_params + 255) / 256 < 1184 ? (r->n_params + 255) / 256 : 1184, 256, 0, s>>>(*r, gate);
    count_launch();
    if (per) {
      r2d2_priority_kernel<<<1, 1024, sizeof(TreeHashScratch), s>>>(*r, gate);
      count_launch();
    }
    r2d2_sync_kernel<<<(r->n_params + 255) / 256 < 1184 ? (r->n_params + 255) / 256 : 1184, 256, 0, s>>>(*r, gate);
    r2d2_finish_kernel<<<1, 1, 0, s>>>(*r, gate);
    count_launch(2);
  }
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
