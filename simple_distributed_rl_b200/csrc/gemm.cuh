// gemm.cuh -- the strided fp32 GEMM family every dense / LSTM / convolution map of the sequence and image networks runs on (csrc/r2d2.cu,
// csrc/imageq.cu): FMA tiles, register-double-buffered tiles, 3 x TF32 tensor-core tiles (mma.sync m16n8k8, hi / lo split, fp32
// accumulate: fp32 accuracy), a warp-per-row kernel for narrow outputs, split-K with an ordered reduction for narrow weight gradients.
// Arbitrary operand strides, so forward, input-gradient and weight-gradient maps are the same kernels; a bias rides as the last weight
// column against a trailing 1 in the activations.
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace srlx {

struct Gate {  // kernels of an update return at once while the memory is below warmup_size (Trainer.train: `if batches is None: return`)
  const srlx_state* st;
  unsigned long long warmup;
  __device__ __forceinline__ bool closed() const { return st != nullptr && st->mem_size < warmup; }
};

// ---- strided fp32 GEMM --------------------------------------------------------------------------------------------------------------
struct GemmP {
  const float* A; long long sa_m, sa_k;
  const float* B; long long sb_k, sb_n;
  float* C; long long ldc;
  int M, N, K;
  int relu, accumulate;
  const float* mask; long long ldmask;  // C = mask > 0 ? C : 0 (ReLU backward against the layer's stored output)
  long long zA, zB, zC;                 // blockIdx.z strides (floats)
  const float* B1;                      // if set: blockIdx.z == 1 reads B1 instead of B + zB (online / target parameter buffers)
  Gate gate;
  float* ws; int ksplit, klen;          // split-K (narrow outputs over a long reduction): slice z of k -> ws[z][M][N], summed in slice order
};

template <int BM, int BN, int TM, int TN>
__device__ __forceinline__ void gemm_mainloop(const float* __restrict__ A, long long sa_m, long long sa_k, const float* __restrict__ B,
                                              long long sb_k, long long sb_n, int M, int N, int K, int m0, int n0, float (&acc)[TM][TN]) {
  constexpr int BK = 16, NT = (BM / TM) * (BN / TN);
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid % (BN / TN), ty = tid / (BN / TN);
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int i = tid; i < BM * BK; i += NT) {
      int mm, kk;
      if (sa_k == 1) { kk = i % BK; mm = i / BK; } else { mm = i % BM; kk = i / BM; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? __ldg(A + (long long)gm * sa_m + (long long)gk * sa_k) : 0.f;
    }
    for (int i = tid; i < BN * BK; i += NT) {
      int nn, kk;
      if (sb_n == 1) { nn = i % BN; kk = i / BN; } else { kk = i % BK; nn = i / BK; }
      const int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < K) ? __ldcg(B + (long long)gk * sb_k + (long long)gn * sb_n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
}

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) sgemm_kernel(const GemmP p) {
  if (p.gate.closed()) return;
  const int z = blockIdx.z;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = threadIdx.x % (BN / TN), ty = threadIdx.x / (BN / TN);
  float acc[TM][TN];
  if (p.ksplit > 1) {
    const int kb = z * p.klen, kl = min(p.klen, p.K - kb);
    gemm_mainloop<BM, BN, TM, TN>(p.A + (long long)kb * p.sa_k, p.sa_m, p.sa_k, p.B + (long long)kb * p.sb_k, p.sb_k, p.sb_n, p.M, p.N, kl, m0, n0, acc);
    float* out = p.ws + (size_t)z * p.M * p.N;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int m = m0 + ty * TM + i, n = n0 + tx * TN + j;
        if (m < p.M && n < p.N) out[(size_t)m * p.N + n] = acc[i][j];
      }
    return;
  }
  const float* A = p.A + (long long)z * p.zA;
  const float* B = (z == 1 && p.B1) ? p.B1 : p.B + (long long)z * p.zB;
  float* C = p.C + (long long)z * p.zC;
  gemm_mainloop<BM, BN, TM, TN>(A, p.sa_m, p.sa_k, B, p.sb_k, p.sb_n, p.M, p.N, p.K, m0, n0, acc);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      float* c = C + (long long)m * p.ldc + n;
      if (p.accumulate) v += *c;
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.mask && !(p.mask[(long long)m * p.ldmask + n] > 0.f)) v = 0.f;
      *c = v;
    }
  }
}

// Register-tiled version for the large maps (hidden block over (S + 1) * B rows, weight gradients, the rollout's E-row steps): BM x BN
// tile per CTA, 16-deep k slices double-buffered through registers, every thread an (BM/16) x (BN/16) micro-tile read as float4 from
// k-major shared tiles.  Operand strides are arbitrary (row- or column-major A and B: forward, input-gradient and weight-gradient maps
// are the same kernel); the bias columns make the leading dimensions odd, so global loads are scalar, with the lanes laid along
// whichever dimension is contiguous.
template <int BM, int BN>
__global__ void __launch_bounds__(256) sgemm_tiled_kernel(const GemmP p) {
  if (p.gate.closed()) return;
  constexpr int BK = 16, TM = BM / 16, TN = BN / 16, NA = BM * BK / 256, NB = BN * BK / 256;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int z = blockIdx.z;
  const float* __restrict__ A = p.A + (long long)z * p.zA;
  const float* __restrict__ B = (z == 1 && p.B1) ? p.B1 : p.B + (long long)z * p.zB;
  float* C = p.C + (long long)z * p.zC;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN, tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const bool a_k = p.sa_k == 1, b_n = p.sb_n == 1;
  float ra[NA], rb[NB];
  auto load = [&](int k0) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int idx = tid + i * 256;
      const int kk = a_k ? idx % BK : idx / BM, mm = a_k ? idx / BK : idx % BM;
      const int gm = m0 + mm, gk = k0 + kk;
      ra[i] = (gm < p.M && gk < p.K) ? __ldcg(A + (long long)gm * p.sa_m + (long long)gk * p.sa_k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int idx = tid + i * 256;
      const int nn = b_n ? idx % BN : idx / BK, kk = b_n ? idx / BN : idx % BK;
      const int gn = n0 + nn, gk = k0 + kk;
      rb[i] = (gn < p.N && gk < p.K) ? __ldcg(B + (long long)gk * p.sb_k + (long long)gn * p.sb_n) : 0.f;
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int idx = tid + i * 256;
      const int kk = a_k ? idx % BK : idx / BM, mm = a_k ? idx / BK : idx % BM;
      As[buf][kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int idx = tid + i * 256;
      const int nn = b_n ? idx % BN : idx / BK, kk = b_n ? idx / BN : idx % BK;
      Bs[buf][kk][nn] = rb[i];
    }
  };
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  load(0);
  store(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < p.K; k0 += BK) {
    const bool more = k0 + BK < p.K;
    if (more) load(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) *reinterpret_cast<float4*>(&a[i]) = *reinterpret_cast<const float4*>(&As[buf][kk][(i / 4) * (BM / (TM / 4)) + ty * 4]);
#pragma unroll
      for (int j = 0; j < TN; j += 4) *reinterpret_cast<float4*>(&b[j]) = *reinterpret_cast<const float4*>(&Bs[buf][kk][(j / 4) * (BN / (TN / 4)) + tx * 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (i / 4) * (BM / (TM / 4)) + ty * 4 + (i & 3);
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (j / 4) * (BN / (TN / 4)) + tx * 4 + (j & 3);
      if (n >= p.N) continue;
      float v = acc[i][j];
      float* c = C + (long long)m * p.ldc + n;
      if (p.accumulate) v += *c;
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.mask && !(p.mask[(long long)m * p.ldmask + n] > 0.f)) v = 0.f;
      *c = v;
    }
  }
}

// Tensor-core version of the tiled GEMM at fp32 accuracy: 3 x TF32.  Every operand is split into hi = tf32(x) and lo = tf32(x - hi)
// and a product is formed as hi*hi + hi*lo + lo*hi on mma.sync.m16n8k8 (fp32 accumulate): the dropped lo*lo term is 2^-22 relative, so
// the result sits within a few fp32 ulps of the FMA chain -- the 1e-4 parity bar holds with two orders of magnitude to spare -- at
// several times the SIMT rate.  Same tiles, strides and epilogue as sgemm_tiled_kernel; 8 warps as 2 x 4, k-major shared tiles padded
// to a stride of 8 mod 32 so that the 32 fragment loads of a warp fall into 32 different banks.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float rest = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rest));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int BM, int BN>
__global__ void __launch_bounds__(256) sgemm_mma_kernel(const GemmP p) {
  if (p.gate.closed()) return;
  constexpr int BK = 16, NA = BM * BK / 256, NB = BN * BK / 256, WM = BM / 2, WN = BN / 4, MT = WM / 16, NT = WN / 8;
  constexpr int LDA = BM + 8, LDB = BN + 8;
  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][LDB];
  const int z = blockIdx.z;
  const float* __restrict__ A = p.A + (long long)z * p.zA;
  const float* __restrict__ B = (z == 1 && p.B1) ? p.B1 : p.B + (long long)z * p.zB;
  float* C = p.C + (long long)z * p.zC;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 2) * WM, wn = (warp & 3) * WN, g = lane >> 2, t = lane & 3;
  const bool a_k = p.sa_k == 1, b_n = p.sb_n == 1;
  float ra[NA], rb[NB];
  auto load = [&](int k0) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int idx = tid + i * 256;
      const int kk = a_k ? idx % BK : idx / BM, mm = a_k ? idx / BK : idx % BM;
      const int gm = m0 + mm, gk = k0 + kk;
      ra[i] = (gm < p.M && gk < p.K) ? __ldcg(A + (long long)gm * p.sa_m + (long long)gk * p.sa_k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int idx = tid + i * 256;
      const int nn = b_n ? idx % BN : idx / BK, kk = b_n ? idx / BN : idx % BK;
      const int gn = n0 + nn, gk = k0 + kk;
      rb[i] = (gn < p.N && gk < p.K) ? __ldcg(B + (long long)gk * p.sb_k + (long long)gn * p.sb_n) : 0.f;
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int idx = tid + i * 256;
      const int kk = a_k ? idx % BK : idx / BM, mm = a_k ? idx / BK : idx % BM;
      As[buf][kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int idx = tid + i * 256;
      const int nn = b_n ? idx % BN : idx / BK, kk = b_n ? idx / BN : idx % BK;
      Bs[buf][kk][nn] = rb[i];
    }
  };
  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
  load(0);
  store(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < p.K; k0 += BK) {
    const bool more = k0 + BK < p.K;
    if (more) load(k0 + BK);
#pragma unroll
    for (int ks = 0; ks < BK; ks += 8) {
      uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        split_tf32(Bs[buf][ks + t][wn + j * 8 + g], bh[j][0], bl[j][0]);
        split_tf32(Bs[buf][ks + t + 4][wn + j * 8 + g], bh[j][1], bl[j][1]);
      }
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        uint32_t ah[4], al[4];
        split_tf32(As[buf][ks + t][wm + i * 16 + g], ah[0], al[0]);
        split_tf32(As[buf][ks + t][wm + i * 16 + g + 8], ah[1], al[1]);
        split_tf32(As[buf][ks + t + 4][wm + i * 16 + g], ah[2], al[2]);
        split_tf32(As[buf][ks + t + 4][wm + i * 16 + g + 8], ah[3], al[3]);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          mma_tf32(acc[i][j], al, bh[j]);  // small terms first
          mma_tf32(acc[i][j], ah, bl[j]);
          mma_tf32(acc[i][j], ah, bh[j]);
        }
      }
    }
    if (more) {
      store(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int m = m0 + wm + i * 16 + g + (q >> 1) * 8, n = n0 + wn + j * 8 + 2 * t + (q & 1);
        if (m >= p.M || n >= p.N) continue;
        float v = acc[i][j][q];
        float* c = C + (long long)m * p.ldc + n;
        if (p.accumulate) v += *c;
        if (p.relu) v = fmaxf(v, 0.f);
        if (p.mask && !(p.mask[(long long)m * p.ldmask + n] > 0.f)) v = 0.f;
        *c = v;
      }
}

// Narrow outputs (the Q head: N = A or 1 + A columns over a K of several hundred): a tile kernel would run one column of CTAs with a
// long serial k loop.  Here a warp owns a row: the lanes stride over k (coalesced in A and, for a row-major weight, in B), keep N
// partial sums each and meet in a shuffle tree -- one pass over A at memory speed.
template <int NMAX>
__global__ void __launch_bounds__(256) sgemm_narrow_kernel(const GemmP p) {
  if (p.gate.closed()) return;
  const int z = blockIdx.z;
  const float* __restrict__ A = p.A + (long long)z * p.zA;
  const float* __restrict__ B = (z == 1 && p.B1) ? p.B1 : p.B + (long long)z * p.zB;
  float* C = p.C + (long long)z * p.zC;
  const int lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= p.M) return;
  float acc[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
  const float* a = A + m * p.sa_m;
  for (int k = lane; k < p.K; k += 32) {
    const float av = __ldcg(a + (long long)k * p.sa_k);
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < p.N) acc[n] = fmaf(av, __ldg(B + (long long)k * p.sb_k + (long long)n * p.sb_n), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < NMAX; ++n)
    for (int sft = 16; sft > 0; sft >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], sft);
  if (lane == 0) {
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
      if (n >= p.N) break;
      float v = acc[n];
      float* c = C + m * p.ldc + n;
      if (p.accumulate) v += *c;
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.mask && !(p.mask[m * p.ldmask + n] > 0.f)) v = 0.f;
      *c = v;
    }
  }
}

static __global__ void splitk_reduce_kernel(const GemmP p) {
  if (p.gate.closed()) return;
  const long long n_out = (long long)p.M * p.N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / p.N), n = (int)(i - (long long)m * p.N);
    float v = 0.f;
    for (int z = 0; z < p.ksplit; ++z) v += p.ws[(size_t)z * n_out + i];
    float* c = p.C + (long long)m * p.ldc + n;
    if (p.accumulate) v += *c;
    if (p.relu) v = fmaxf(v, 0.f);
    if (p.mask && !(p.mask[(long long)m * p.ldmask + n] > 0.f)) v = 0.f;
    *c = v;
  }
}

static const bool g_gemm_simt = getenv("SRLX_GEMM_SIMT") != nullptr;  // diagnostic: FMA tiles instead of 3 x TF32 tensor-core tiles

// wide_split: split-K whenever the output has fewer 64 x 64 tiles than the GPU has SMs (the convolution weight gradients: a few tiles
// over a reduction of batch x positions rows)
static int launch_gemm(const GemmP& p_in, int nz, cudaStream_t s, float* ws = nullptr, size_t ws_floats = 0, bool wide_split = false) {
  GemmP p = p_in;
  if (p.M <= 0 || p.N <= 0) return 0;
  const bool few_tiles = wide_split && (long long)((p.M + 63) / 64) * ((p.N + 63) / 64) < 148;
  if (nz == 1 && ws && (p.M <= 32 || p.N <= 32 || few_tiles) && p.K >= 1024) {
    int splits = p.K / 256;
    if (splits > 32) splits = 32;
    while (splits > 1 && (size_t)splits * p.M * p.N > ws_floats) --splits;
    if (splits > 1) {
      p.klen = ((p.K + splits - 1) / splits + 15) / 16 * 16;
      p.ksplit = (p.K + p.klen - 1) / p.klen;
      p.ws = ws;
      dim3 grid((p.N + 31) / 32, (p.M + 31) / 32, p.ksplit);
      sgemm_kernel<32, 32, 2, 4><<<grid, 128, 0, s>>>(p);
      const long long n_out = (long long)p.M * p.N;
      splitk_reduce_kernel<<<(unsigned)((n_out + 255) / 256 < 592 ? (n_out + 255) / 256 : 592), 256, 0, s>>>(p);
      count_launch(2);
      return 0;
    }
  }
  const long long t128 = (long long)((p.M + 127) / 128) * ((p.N + 127) / 128) * nz, t64 = (long long)((p.M + 63) / 64) * ((p.N + 63) / 64) * nz;
  if (p.N <= 8 && p.M >= 256 && p.K >= 64) {  // a narrow head over many rows: warp per row
    dim3 grid((unsigned)((p.M + 7) / 8), 1, nz);
    sgemm_narrow_kernel<8><<<grid, 256, 0, s>>>(p);
  } else if (p.M <= 32 || p.N <= 32) {
    dim3 grid((p.N + 31) / 32, (p.M + 31) / 32, nz);
    sgemm_kernel<32, 32, 2, 4><<<grid, 128, 0, s>>>(p);
  } else if (t128 >= 120) {  // enough 128 x 128 tiles for the 148 SMs
    dim3 grid((p.N + 127) / 128, (p.M + 127) / 128, nz);
    if (g_gemm_simt) sgemm_tiled_kernel<128, 128><<<grid, 256, 0, s>>>(p);
    else sgemm_mma_kernel<128, 128><<<grid, 256, 0, s>>>(p);
  } else if (t64 >= 32) {
    dim3 grid((p.N + 63) / 64, (p.M + 63) / 64, nz);
    if (g_gemm_simt) sgemm_tiled_kernel<64, 64><<<grid, 256, 0, s>>>(p);
    else sgemm_mma_kernel<64, 64><<<grid, 256, 0, s>>>(p);
  } else {
    dim3 grid((p.N + 63) / 64, (p.M + 63) / 64, nz);
    sgemm_kernel<64, 64, 4, 4><<<grid, 256, 0, s>>>(p);
  }
  count_launch();
  return 0;
}


}  // namespace srlx
