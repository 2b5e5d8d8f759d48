// imageq.cu -- the image configs of the reference on device:
//   srlx_image_process  ImageProcessor.remap_observation (srl/rl/processors/image_processor.py:104-154) for a batch of uint8 frames:
//                       gray / colour conversion, trimming, cv2.resize INTER_LINEAR and the normalisation in one pass
//   srlx_imageq_*       InputImageBlock + DQNImageBlock + hidden block + Linear(A) (srl/rl/torch_/blocks/dqn_image_block.py:10-62,
//                       srl/algorithms/dqn/model_torch.py:17-29) and Trainer.train (model_torch.py:75-131) with calc_target_q
//                       (srl/algorithms/dqn/dqn.py:143-173)
// Convolutions are im2col (replicate padding = clamped index, the bias column's 1 written by the same kernel) + the strided GEMM
// family of gemm.cuh; the input gradient of a convolution is the GEMM dOut x W followed by a GATHER col2im (every input pixel sums the
// window slots that read it, in a fixed order: deterministic, no atomics).  Activations are NHWC, so a conv layer's output is the next
// layer's im2col source and the last one IS the flattened input of the first dense layer.
#include <algorithm>
#include <mutex>
#include <type_traits>

#include "gemm_tc3.cuh"
#include "net.cuh"

namespace srlx {

// ---- image processor ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int img_px(const unsigned char* __restrict__ f, const srlx_image_proc& p, int y, int x, int c) {
  const unsigned char* s = f + ((size_t)(y + p.top) * p.src_w + (x + p.left)) * p.src_c;
  if (p.src_c == 3 && p.out_c == 1) return (s[0] * 9798 + s[1] * 19235 + s[2] * 3735 + 16384) >> 15;  // cv2 RGB2GRAY, 15-bit coefficients
  return p.src_c == 1 ? s[0] : s[c];
}

template <typename OutT>
__global__ void __launch_bounds__(256) image_process_kernel(const __grid_constant__ srlx_image_proc p, const unsigned char* __restrict__ src,
                                                            const uint32_t n, OutT* __restrict__ out, const uint64_t out_stride) {
  const long long per = (long long)p.out_h * p.out_w * p.out_c, total = per * n;
  const size_t src_frame = (size_t)p.src_h * p.src_w * p.src_c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long f = i / per;
    int r = (int)(i - f * per);
    const int c = r % p.out_c;
    r /= p.out_c;
    const int ox = r % p.out_w, oy = r / p.out_w;
    const unsigned char* fr = src + (size_t)f * src_frame;
    int v;
    if (p.resize) {
      const int x0 = p.x_idx[ox], x1 = min(x0 + 1, p.trim_w - 1), a0 = p.x_coef[2 * ox], a1 = p.x_coef[2 * ox + 1];
      const int yr = p.y_idx[oy], y0 = min(max(yr, 0), p.trim_h - 1), y1 = min(max(yr + 1, 0), p.trim_h - 1);
      const int b0 = p.y_coef[2 * oy], b1 = p.y_coef[2 * oy + 1];
      const int r0 = img_px(fr, p, y0, x0, c) * a0 + img_px(fr, p, y0, x1, c) * a1;
      const int r1 = img_px(fr, p, y1, x0, c) * a0 + img_px(fr, p, y1, x1, c) * a1;
      v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;  // VResizeLinear, FixedPtCast<int, uchar, 22>
      v = min(max(v, 0), 255);
    } else {
      v = img_px(fr, p, oy, ox, c);
    }
    const size_t o = (size_t)f * out_stride + (size_t)(i - f * per);
    if constexpr (sizeof(OutT) == 1) {
      out[o] = (OutT)v;
    } else {
      const float x = (float)v;  // image_processor.py:141-146, float32 arithmetic
      out[o] = p.normalize == 1 ? __fdiv_rn(x, p.max_val) : __fsub_rn(__fdiv_rn(__fmul_rn(x, 2.f), p.max_val), 1.f);
    }
  }
}

// Staged version (the usual case: the frame, as gray bytes when the output is gray, fits in shared memory and its size is a multiple of
// 16 bytes): one CTA per frame.  Phase 1 streams the frame in with 16-byte loads -- 48 bytes = 16 RGB pixels per step, converted to 16
// gray bytes and stored as one 16-byte word -- so every source byte crosses the memory system exactly once and no lane issues byte
// loads to global memory; phase 2 forms the outputs from shared memory (4 taps each) and writes them coalesced.  The resize tables sit
// in shared memory too.
// OC = output channels, SC = channels of the staged frame (1: gray bytes or a gray source; 3: RGB kept), RESIZE: through the tables --
// compile-time so that the per-pixel loop has no branches and the channel loop unrolls (37.8 k -> 20.6 k warp instructions per Atari
// frame together with the dp2a gray conversion below)
template <typename OutT, int OC, int SC, bool RESIZE>
__global__ void __launch_bounds__(256) image_process_staged_kernel(const __grid_constant__ srlx_image_proc p, const unsigned char* __restrict__ src,
                                                                   const uint32_t n, OutT* __restrict__ out, const uint64_t out_stride,
                                                                   const int to_gray) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  int* xi = reinterpret_cast<int*>(sm_raw);
  int* xc = xi + p.out_w;
  int* yi = xc + 2 * p.out_w;
  int* yc = yi + p.out_h;
  unsigned char* fr = sm_raw + (((size_t)(3 * p.out_w + 3 * p.out_h) * 4 + 15) / 16) * 16;
  const int tid = threadIdx.x;
  if (RESIZE) {  // x: both tap offsets in one word and both coefficients in one word (offsets < 2^16, coefficients <= 2048)
    for (int i = tid; i < p.out_w; i += 256) {
      const int x0 = p.x_idx[i], x1 = min(x0 + 1, p.trim_w - 1);
      xi[i] = ((x0 + p.left) * SC) | (((x1 + p.left) * SC) << 16);
      xc[i] = p.x_coef[2 * i] | (p.x_coef[2 * i + 1] << 16);
    }
    for (int i = tid; i < p.out_h; i += 256) { yi[i] = p.y_idx[i]; yc[2 * i] = p.y_coef[2 * i]; yc[2 * i + 1] = p.y_coef[2 * i + 1]; }
  }
  const size_t frame_bytes = (size_t)p.src_h * p.src_w * p.src_c;
  constexpr int sc = SC;  // channels of the staged frame
  __shared__ float lut[256];  // image_processor.py:141-146 for every byte value, float32 arithmetic
  {
    const float x = (float)tid;
    lut[tid] = p.normalize == 1 ? __fdiv_rn(x, p.max_val) : __fsub_rn(__fdiv_rn(__fmul_rn(x, 2.f), p.max_val), 1.f);
  }
  __syncthreads();
  for (uint32_t f = blockIdx.x; f < n; f += gridDim.x) {
    const uint4* g = reinterpret_cast<const uint4*>(src + (size_t)f * frame_bytes);
    if (to_gray) {
      const int groups = p.src_h * p.src_w / 16;
#pragma unroll 3
      for (int i = tid; i < groups; i += 256) {  // unrolled: 9 independent 16-byte loads in flight per thread
        uint32_t w[12];
        const uint4 a = __ldcs(g + 3 * i), b = __ldcs(g + 3 * i + 1), c = __ldcs(g + 3 * i + 2);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w; w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
        uint32_t o[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < 16; ++j) {  // pixel j = bytes 3j .. 3j + 2: one byte permute, two 16 x 8-bit dot products, one shift
          const int k = 3 * j, sft = k & 3;
          const uint32_t px = __byte_perm(w[k >> 2], w[(k + 3) >> 2 < 12 ? (k + 3) >> 2 : 11], (uint32_t)(sft | ((sft + 1) << 4) | ((sft + 2) << 8) | ((sft + 2) << 12)));
          const uint32_t y = __dp2a_hi(3735u, px, __dp2a_lo(9798u | (19235u << 16), px, 16384u)) >> 15;  // R * 9798 + G * 19235 + 16384, + B * 3735
          o[j >> 2] |= y << (8 * (j & 3));
        }
        reinterpret_cast<uint4*>(fr)[i] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    } else {
      const int words = (int)(frame_bytes / 16);
      for (int i = tid; i < words; i += 256) reinterpret_cast<uint4*>(fr)[i] = __ldcs(g + i);
    }
    __syncthreads();
    OutT* of = out + (size_t)f * out_stride;
    // lanes own output COLUMNS (3 per lane and block of 96: their tap offsets and coefficients stay in registers), warps own rows: a
    // row's two source rows and vertical coefficients are warp-uniform, a pixel costs four byte taps and the fixed-point blend; the
    // normalisation is a 256-entry table
    const int warp = tid >> 5, lane = tid & 31;
    for (int xb = 0; xb < p.out_w; xb += 96) {
      int o0[3], o1[3], a0[3], a1[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int ox = min(xb + lane + 32 * k, p.out_w - 1);
        o0[k] = o1[k] = (ox + p.left) * SC; a0[k] = a1[k] = 0;
        if (RESIZE) {
          const uint32_t xo = (uint32_t)xi[ox], xa = (uint32_t)xc[ox];
          o0[k] = xo & 0xffff; o1[k] = xo >> 16; a0[k] = xa & 0xffff; a1[k] = xa >> 16;
        }
      }
      for (int oy = warp; oy < p.out_h; oy += 8) {
        int r0 = (oy + p.top) * p.src_w * SC, r1 = r0, b0 = 0, b1 = 0;
        if (RESIZE) {
          const int yr = yi[oy];
          r0 = (min(max(yr, 0), p.trim_h - 1) + p.top) * p.src_w * SC;
          r1 = (min(max(yr + 1, 0), p.trim_h - 1) + p.top) * p.src_w * SC;
          b0 = yc[2 * oy]; b1 = yc[2 * oy + 1];
        }
        OutT* orow = of + (size_t)oy * p.out_w * OC;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int ox = xb + lane + 32 * k;
          if (ox >= p.out_w) break;
#pragma unroll
          for (int c = 0; c < OC; ++c) {
            const int cs = SC == 1 ? 0 : c;
            int v;
            if (RESIZE) {
              const int t0 = fr[r0 + o0[k] + cs] * a0[k] + fr[r0 + o1[k] + cs] * a1[k];
              const int t1 = fr[r1 + o0[k] + cs] * a0[k] + fr[r1 + o1[k] + cs] * a1[k];
              v = (((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2;
              v = min(max(v, 0), 255);
            } else {
              v = fr[r0 + o0[k] + cs];
            }
            if constexpr (sizeof(OutT) == 1) orow[ox * OC + c] = (OutT)v;
            else __stcs(orow + ox * OC + c, lut[v]);
          }
        }
      }
    }
    __syncthreads();
  }
}

// ---- im2col / col2im ----------------------------------------------------------------------------------------------------------------
// dIn[b][ih][iw][c] = (act > 0) * sum over the window slots (oh, kh, ow, kw) whose clamped source is (ih, iw) of dcol[(b, oh, ow)][(kh, kw, c)]
__global__ void __launch_bounds__(256) col2im_kernel(const ConvG g, const float* __restrict__ dcol, const float* __restrict__ act,
                                                     float* __restrict__ din, const long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % g.C);
    const int iw = (int)((i / g.C) % g.W), ih = (int)((i / ((long long)g.C * g.W)) % g.H);
    const long long b = i / ((long long)g.C * g.W * g.H);
    float sum = 0.f;
    if (act[i] > 0.f) {
      const int oh_lo = ih == 0 ? 0 : max(0, (ih + g.p - g.k + 1 + g.s - 1) / g.s), oh_hi = ih == g.H - 1 ? g.OH - 1 : min(g.OH - 1, (ih + g.p) / g.s);
      const int ow_lo = iw == 0 ? 0 : max(0, (iw + g.p - g.k + 1 + g.s - 1) / g.s), ow_hi = iw == g.W - 1 ? g.OW - 1 : min(g.OW - 1, (iw + g.p) / g.s);
      const bool in_h = ih > 0 && ih < g.H - 1, in_w = iw > 0 && iw < g.W - 1;  // an interior pixel is read by exactly ONE kh (kw) per oh (ow)
      for (int oh = oh_lo; oh <= oh_hi; ++oh) {
        const int kh0 = in_h ? ih + g.p - oh * g.s : 0, kh1 = in_h ? kh0 : g.k - 1;
        for (int kh = kh0; kh <= kh1; ++kh) {
          if (!in_h && min(max(oh * g.s - g.p + kh, 0), g.H - 1) != ih) continue;
          for (int ow = ow_lo; ow <= ow_hi; ++ow) {
            const float* row = dcol + ((b * g.OH + oh) * g.OW + ow) * (long long)g.K;
            const int kw0 = in_w ? iw + g.p - ow * g.s : 0, kw1 = in_w ? kw0 : g.k - 1;
            for (int kw = kw0; kw <= kw1; ++kw) {
              if (!in_w && min(max(ow * g.s - g.p + kw, 0), g.W - 1) != iw) continue;
              const int j = g.c_fast ? (kh * g.k + kw) * g.C + c : (c * g.k + kh) * g.k + kw;
              sum += row[j];
            }
          }
        }
      }
    }
    din[i] = sum;
  }
}

// ---- implicit-GEMM tiles ------------------------------------------------------------------------------------------------------------
// The convolution maps without a materialised im2col matrix: the tile loader GATHERS the operand from the NHWC / NCHW source (replicate
// padding = clamped index, the bias column = 1), as the A operand of a forward map (rows = output positions) or as the B operand of a
// weight-gradient map (reduction over output positions).  Same 3 x TF32 mma.sync tiles, register double buffering and epilogue as
// sgemm_mma_kernel; plus split-K over blockIdx.z (partials to ws[z][M][N], summed in slice order by splitk_reduce_kernel) so that a map
// with few output tiles still fills the 148 SMs.  Plain strided operands go through the same kernel (gather = 0).
struct IGemmP {
  GemmP g;
  ConvG cv;
  const float* src;  // gather source (fp32: uint8 states are converted once per pass, u8_to_f32_kernel)
  const float* one;  // a 1.0f in global memory: the source of the bias column
  int gather;        // 0: none; 1: A = im2col(src) [M = positions][K = cv.K + 1]; 2: B = im2col(src) [K = positions][N = cv.K + 1]
};

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, int bytes) {  // bytes = 0: zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__global__ void u8_to_f32_kernel(const unsigned char* __restrict__ in, const long long n, const float div, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = __fdiv_rn((float)in[i], div);
}

// The k loop is a STAGES-deep cp.async pipeline: every thread copies its 4-byte tile slots of slice it + STAGES - 1 straight into shared
// memory (zero fill outside the operand, the bias column from a constant 1) while the tensor cores work on slice it -- no register
// staging, one __syncthreads per slice, global latency covered by STAGES - 1 slices in flight.
template <int BM, int BN>
__global__ void __launch_bounds__(256) igemm_kernel(const IGemmP q) {
  const GemmP& p = q.g;
  const ConvG& cv = q.cv;
  constexpr int BK = 16, STAGES = 4, NA = BM * BK / 256, NB = BN * BK / 256, WM = BM / 2, WN = BN / 4, MT = WM / 16, NT = WN / 8;
  constexpr int LDA = BM + 8, LDB = BN + 8;
  __shared__ __align__(16) float As[STAGES][BK][LDA];
  __shared__ __align__(16) float Bs[STAGES][BK][LDB];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 2) * WM, wn = (warp & 3) * WN, g = lane >> 2, t = lane & 3;
  const int kb = p.ksplit > 1 ? blockIdx.z * p.klen : 0, ke = p.ksplit > 1 ? min(p.K, kb + p.klen) : p.K;
  const bool a_k = q.gather == 1 || p.sa_k == 1, b_n = q.gather == 2 || p.sb_n == 1;
  // Everything about a tile slot that does not change along k is decoded ONCE (the loop is issue-bound on index arithmetic otherwise):
  // per A slot its shared-memory offset and either (row base, ih0, iw0) of a gathered row or the element offset of a strided one; per B
  // slot likewise with (c, kh, kw) of a gathered column.  Along k the offsets advance by additions: 16 columns of the im2col matrix are
  // walked as (c, kw, kh) counters, 16 positions as (ow, oh, image) counters.  Offsets are 32-bit (the launcher checks the extents).
  int sa[NA], sb[NB];                 // shared-memory offsets
  int a0[NA], a1[NA], a2[NA];         // gather A: row base (-1: no row), ih0, iw0   | strided A: element offset (-1: no row), k index, -
  int b0[NB], b1[NB], b2[NB];         // gather B: c (-1 bias, -2 none), kh, kw      | strided B: element offset (-1: no column), k index, -
  int kc = 0, kkw = 0, kkh = 0, kg = 0;     // gather A: this thread's column kg = (kkh, kkw, kc)
  int rw[NB], rh[NB], rbase[NB], rrow[NB];  // gather B: this slot's position (ow, oh, image base offset, row)
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const int idx = tid + i * 256;
    const int kk = a_k ? idx % BK : idx / BM, mm = a_k ? idx / BK : idx % BM;
    sa[i] = kk * LDA + mm;
    const long long row = m0 + mm;
    if (q.gather == 1) {
      a0[i] = -1; a1[i] = 0; a2[i] = 0;
      if (row < p.M) {
        const int ow = (int)(row % cv.OW), oh = (int)((row / cv.OW) % cv.OH);
        a0[i] = (int)((row / ((long long)cv.OW * cv.OH)) * cv.sb);
        a1[i] = oh * cv.s - cv.p;
        a2[i] = ow * cv.s - cv.p;
      }
    } else {
      a0[i] = row < p.M ? (int)(row * p.sa_m + (long long)(kb + kk) * p.sa_k) : -1;
      a1[i] = kb + kk; a2[i] = 0;
    }
  }
  if (q.gather == 1) {
    kg = kb + (tid & (BK - 1));
    if (kg < cv.K) col_split(cv, kg, kc, kkh, kkw);
  }
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int idx = tid + i * 256;
    const int nn = b_n ? idx % BN : idx / BK, kk = b_n ? idx / BN : idx % BK;
    sb[i] = kk * LDB + nn;
    const int j = n0 + nn;
    b1[i] = 0; b2[i] = 0; rw[i] = 0; rh[i] = 0; rbase[i] = 0; rrow[i] = 0;
    if (q.gather == 2) {
      b0[i] = j < cv.K ? 0 : (j == cv.K ? -1 : -2);
      if (j < cv.K) col_split(cv, j, b0[i], b1[i], b2[i]);
      const long long row = kb + kk;
      rrow[i] = (int)row;
      rw[i] = (int)(row % cv.OW); rh[i] = (int)((row / cv.OW) % cv.OH);
      rbase[i] = (int)((row / ((long long)cv.OW * cv.OH)) * cv.sb);
    } else {
      b0[i] = j < p.N ? (int)((long long)(kb + kk) * p.sb_k + (long long)j * p.sb_n) : -1;
      b1[i] = kb + kk;
    }
  }
  const int a_step = (int)(BK * p.sa_k), b_step = (int)(BK * p.sb_k), img = (int)cv.sb;
  const int csc = (int)cv.sc, csh = (int)cv.sh, csw = (int)cv.sw;
  auto issue = [&](int buf) {  // copies the NEXT k slice of this thread's slots into stage `buf` and advances the counters
    float* as = &As[buf][0][0];
    float* bs = &Bs[buf][0][0];
    if (q.gather == 1) {
      const bool in_k = kg < ke, bias = kg >= cv.K;
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        const float* src = q.one;
        int bytes = 0;
        if (a0[i] >= 0 && in_k) {
          bytes = 4;
          if (!bias) {
            const int ih = min(max(a1[i] + kkh, 0), cv.H - 1), iw = min(max(a2[i] + kkw, 0), cv.W - 1);
            src = q.src + (a0[i] + kc * csc + ih * csh + iw * csw);
          }
        }
        cp_async4(as + sa[i], src, bytes);
      }
      kg += BK;
      if (cv.c_fast) {
        kc += BK;
        while (kc >= cv.C) { kc -= cv.C; if (++kkw == cv.k) { kkw = 0; ++kkh; } }
      } else {
        kkw += BK;
        while (kkw >= cv.k) { kkw -= cv.k; if (++kkh == cv.k) { kkh = 0; ++kc; } }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        const bool ok = a0[i] >= 0 && a1[i] < ke;
        cp_async4(as + sa[i], ok ? p.A + a0[i] : p.A, ok ? 4 : 0);
        a0[i] += a0[i] >= 0 ? a_step : 0;
        a1[i] += BK;
      }
    }
    if (q.gather == 2) {
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const float* src = q.one;
        int bytes = 0;
        if (rrow[i] < ke && b0[i] != -2) {
          bytes = 4;
          if (b0[i] != -1) {
            const int ih = min(max(rh[i] * cv.s - cv.p + b1[i], 0), cv.H - 1), iw = min(max(rw[i] * cv.s - cv.p + b2[i], 0), cv.W - 1);
            src = q.src + (rbase[i] + b0[i] * csc + ih * csh + iw * csw);
          }
        }
        cp_async4(bs + sb[i], src, bytes);
        rrow[i] += BK;
        rw[i] += BK;
        while (rw[i] >= cv.OW) { rw[i] -= cv.OW; if (++rh[i] == cv.OH) { rh[i] = 0; rbase[i] += img; } }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const bool ok = b0[i] >= 0 && b1[i] < ke;
        cp_async4(bs + sb[i], ok ? p.B + b0[i] : p.B, ok ? 4 : 0);
        b0[i] += b0[i] >= 0 ? b_step : 0;
        b1[i] += BK;
      }
    }
  };
  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
  const int nk = (ke - kb + BK - 1) / BK;
#pragma unroll
  for (int st = 0; st < STAGES - 1; ++st) {
    if (st < nk) issue(st);
    cp_async_commit();
  }
  for (int it = 0; it < nk; ++it) {
    cp_async_wait<STAGES - 2>();  // slice `it` has landed (for this thread; the barrier makes it so for all)
    __syncthreads();              // ... and everybody is done with slice it - 1, whose stage the next copy overwrites
    if (it + STAGES - 1 < nk) issue((it + STAGES - 1) % STAGES);
    cp_async_commit();
    const int buf = it % STAGES;
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
          mma_tf32(acc[i][j], al, bh[j]);
          mma_tf32(acc[i][j], ah, bl[j]);
          mma_tf32(acc[i][j], ah, bh[j]);
        }
      }
    }
  }
  float* part = p.ksplit > 1 ? p.ws + (size_t)blockIdx.z * p.M * p.N : nullptr;
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int m = m0 + wm + i * 16 + g + (r >> 1) * 8, n = n0 + wn + j * 8 + 2 * t + (r & 1);
        if (m >= p.M || n >= p.N) continue;
        float v = acc[i][j][r];
        if (part) { part[(size_t)m * p.N + n] = v; continue; }
        float* c = p.C + (long long)m * p.ldc + n;
        if (p.accumulate) v += *c;
        if (p.relu) v = fmaxf(v, 0.f);
        if (p.mask && !(p.mask[(long long)m * p.ldmask + n] > 0.f)) v = 0.f;
        *c = v;
      }
}

template <int BN>
static int t3_launch_bn(const T3P& p, dim3 grid, cudaStream_t s) {
  static bool attr_set[64] = {};  // function attributes are per device: one process may drive several (tests/test_multi_gpu.py)
  int dev = 0;
  SRLX_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    SRLX_CHECK_CUDA(cudaFuncSetAttribute(t3_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t3_smem_bytes<BN>()));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  t3_gemm_kernel<BN><<<grid, T3_THREADS, t3_smem_bytes<BN>(), s>>>(p);
  count_launch();
  return 0;
}

// C[M][N] = A[M][K] x B[K][N] on the tcgen05 tiles of gemm_tc3.cuh: the larger of (M, N) takes the 128-row side, split-K when the
// tiles do not fill the SMs (one CTA per SM: the 3-stage hi / lo ring takes 120 - 198 KB of shared memory)
static int launch_t3(const IGemmP& q, cudaStream_t s, float* ws, size_t ws_floats) {
  const GemmP& g = q.g;
  if (g.M <= 0 || g.N <= 0) return 0;
  const long long lim = (1LL << 31) - 1;
  SRLX_REQUIRE((q.gather == 1 || (long long)g.M * llabs(g.sa_m) + (long long)g.K * llabs(g.sa_k) < lim) &&
                   (q.gather == 2 || (long long)g.K * llabs(g.sb_k) + (long long)g.N * llabs(g.sb_n) < lim) &&
                   (q.gather == 0 || ((long long)(q.gather == 1 ? g.M : g.K) / ((long long)q.cv.OH * q.cv.OW) + 1) * q.cv.sb < lim),
               "imageq: an operand of a map exceeds 2^31 elements (batch too large for one launch)");
  T3Op x{}, y{};  // x: rows = the M index, y: rows = the N index
  x.rows = g.M; y.rows = g.N;
  if (q.gather == 1) { x.mode = 1; x.ptr = q.src; } else { x.mode = 0; x.ptr = g.A; x.s_row = (int)g.sa_m; x.s_k = (int)g.sa_k; }
  if (q.gather == 2) { y.mode = 2; y.ptr = q.src; } else { y.mode = 0; y.ptr = g.B; y.s_row = (int)g.sb_n; y.s_k = (int)g.sb_k; }
  const bool swap = g.N > g.M;
  T3P p{};
  p.a = swap ? y : x; p.b = swap ? x : y;
  p.cv = q.cv; p.one = q.one; p.K = g.K;
  p.C = g.C; p.c_a = swap ? 1 : g.ldc; p.c_b = swap ? g.ldc : 1;
  p.mask = g.mask; p.m_a = swap ? 1 : g.ldmask; p.m_b = swap ? g.ldmask : 1;
  p.w_a = swap ? 1 : g.N; p.w_b = swap ? g.N : 1; p.w_slice = (long long)g.M * g.N;
  p.relu = g.relu; p.accumulate = g.accumulate;
  const int BN = p.b.rows <= 32 ? 32 : 64;
  const long long tiles = (long long)((p.a.rows + T3_BM - 1) / T3_BM) * ((p.b.rows + BN - 1) / BN);
  int splits = 1;
  if (tiles < 148 && g.K >= 128 && ws) {
    splits = (int)((148 + tiles - 1) / tiles);
    if (splits > g.K / 64) splits = g.K / 64;
    if (splits > 64) splits = 64;
    while (splits > 1 && (size_t)splits * g.M * g.N > ws_floats) --splits;
  }
  p.ksplit = 1;
  if (splits > 1) {
    p.klen = ((g.K + splits - 1) / splits + T3_BK - 1) / T3_BK * T3_BK;
    p.ksplit = (g.K + p.klen - 1) / p.klen;
    p.ws = ws;
  }
  dim3 grid((p.b.rows + BN - 1) / BN, (p.a.rows + T3_BM - 1) / T3_BM, p.ksplit);
  int rc = BN == 32 ? t3_launch_bn<32>(p, grid, s) : t3_launch_bn<64>(p, grid, s);
  if (rc) return rc;
  if (p.ksplit > 1) {
    GemmP r = g;
    r.gate = Gate{nullptr, 0}; r.ws = ws; r.ksplit = p.ksplit;
    const long long n_out = (long long)g.M * g.N;
    splitk_reduce_kernel<<<(unsigned)((n_out + 255) / 256 < 592 ? (n_out + 255) / 256 : 592), 256, 0, s>>>(r);
    count_launch();
  }
  return 0;
}

// Which tile engine runs a map.  Measured per map at the Atari setting (profiles/r3e_* mma.sync, r3m_* tcgen05): the tcgen05 tiles win
// where a gathered forward map has enough 128-row tiles to put two CTAs on every SM (batch 256: conv 2 / 3 forward 83 vs 117 us,
// conv 1 forward 114 vs 139 us) and lose on small grids and on the weight-gradient maps (row-fast 4-byte gathers: 353 vs 230 us), so the
// default is a HYBRID: tcgen05 for forward convolutions with >= 200 tiles, the cp.async + mma.sync tiles for everything else.
// SRLX_IMAGE_TC3=1 forces every map onto tcgen05 (tests run the whole suite of parity checks that way), SRLX_IMAGE_MMA_SYNC=1 onto
// mma.sync.
static const bool g_image_tc3 = getenv("SRLX_IMAGE_TC3") != nullptr;
static const bool g_image_mma_sync = getenv("SRLX_IMAGE_MMA_SYNC") != nullptr;

static int launch_igemm(IGemmP q, cudaStream_t s, float* ws, size_t ws_floats) {
  if (g_image_tc3) return launch_t3(q, s, ws, ws_floats);
  if (!g_image_mma_sync && q.gather == 1 && (long long)((q.g.M + T3_BM - 1) / T3_BM) * ((q.g.N + 63) / 64) >= 200) return launch_t3(q, s, ws, ws_floats);
  GemmP& p = q.g;
  if (p.M <= 0 || p.N <= 0) return 0;
  p.gate = Gate{nullptr, 0};
  const long long lim = (1LL << 31) - 1;  // 32-bit element offsets inside the kernel
  SRLX_REQUIRE((q.gather == 1 || (long long)p.M * llabs(p.sa_m) + (long long)p.K * llabs(p.sa_k) < lim) &&
                   (q.gather == 2 || (long long)p.K * llabs(p.sb_k) + (long long)p.N * llabs(p.sb_n) < lim) &&
                   (q.gather == 0 || (long long)(q.gather == 1 ? p.M : p.K) / ((long long)q.cv.OH * q.cv.OW) * q.cv.sb < lim),
               "imageq: an operand of a map exceeds 2^31 elements (batch too large for one launch)");
  const bool n32 = p.N <= 32, m32 = !n32 && p.M <= 32;
  const int BM = m32 ? 32 : 64, BN = n32 ? 32 : 64;
  const long long tiles = (long long)((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  if (q.gather == 0 && tiles >= 4 * 148 && p.M >= 128 && p.N >= 128) return launch_gemm(p, 1, s);  // large plain maps: the 128 x 128 tiles
  int splits = 1;
  if (tiles < 148 && p.K >= 128 && ws) {
    splits = (int)((2 * 148 + tiles - 1) / tiles);
    if (splits > p.K / 64) splits = p.K / 64;
    if (splits > 64) splits = 64;
    while (splits > 1 && (size_t)splits * p.M * p.N > ws_floats) --splits;
  }
  p.ksplit = 1;
  if (splits > 1) {
    p.klen = ((p.K + splits - 1) / splits + 15) / 16 * 16;
    p.ksplit = (p.K + p.klen - 1) / p.klen;
    p.ws = ws;
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.ksplit);
  if (n32) igemm_kernel<64, 32><<<grid, 256, 0, s>>>(q);
  else if (m32) igemm_kernel<32, 64><<<grid, 256, 0, s>>>(q);
  else igemm_kernel<64, 64><<<grid, 256, 0, s>>>(q);
  count_launch();
  if (p.ksplit > 1) {
    const long long n_out = (long long)p.M * p.N;
    splitk_reduce_kernel<<<(unsigned)((n_out + 255) / 256 < 592 ? (n_out + 255) / 256 : 592), 256, 0, s>>>(p);
    count_launch();
  }
  return 0;
}

// ---- loss, Adam ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rescaling_d(double x) {  // srl/rl/functions.py:10-12 on a float64 array
  const double s = (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : 0.0);
  return s * (sqrt(fabs(x) + 1.0) - 1.0) + 0.001 * x;
}

// one CTA: calc_target_q (dqn.py:143-173, without invalid actions), q = sum(Q * onehot), nn.HuberLoss()(target * w, q * w) (mean), its
// gradient wrt Q, priorities = |target - q|
__global__ void __launch_bounds__(256) imageq_loss_kernel(const srlx_imageq q, const float* __restrict__ q0, const float* __restrict__ qn_online,
                                                          const float* __restrict__ qn_target, const int32_t* __restrict__ action,
                                                          const float* __restrict__ reward, const float* __restrict__ undone,
                                                          const float* __restrict__ weights, const int B, float* __restrict__ dq,
                                                          float* __restrict__ pri, float* __restrict__ loss_out, float* __restrict__ tq_buf) {
  __shared__ float red[256];
  const int A = q.n_actions;
  float part = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* nt = qn_target + (size_t)b * A;
    float maxq;
    if (q.enable_double_dqn) {
      const float* no = qn_online + (size_t)b * A;
      int best = 0;
      for (int a = 1; a < A; ++a)
        if (no[a] > no[best]) best = a;  // np.argmax: first maximum
      maxq = nt[best];
    } else {
      maxq = nt[0];
      for (int a = 1; a < A; ++a) maxq = fmaxf(maxq, nt[a]);
    }
    if (q.enable_rescale) maxq = inverse_rescaling_f(maxq);
    float tq;
    if (q.target_f32) {  // rainbow_nomultisteps.py:36-41: every array is float32, the python-float discount is a weak scalar
      tq = __fadd_rn(reward[b], __fmul_rn(__fmul_rn(undone[b], (float)q.discount), maxq));
      if (q.enable_rescale) tq = rescaling_f(tq);
    } else {  // dqn.py:166-171: reward (f32) + undone (int array) * discount (python float) * maxq (f32) -> float64, then .astype(float32)
      double t = (double)reward[b] + ((double)undone[b] * q.discount) * (double)maxq;
      if (q.enable_rescale) t = rescaling_d(t);
      tq = (float)t;
    }
    const int a_sel = action[b];
    const float qv = q0[(size_t)b * A + a_sel], w = weights[b];
    const float x = __fsub_rn(__fmul_rn(qv, w), __fmul_rn(tq, w));
    const float ax = fabsf(x);
    part += ax <= 1.f ? 0.5f * x * x : ax - 0.5f;
    const float gx = fminf(fmaxf(x, -1.f), 1.f);
    for (int a = 0; a < A; ++a) dq[(size_t)b * A + a] = a == a_sel ? w * gx / (float)B : 0.f;
    pri[b] = fabsf(tq - qv);
    if (tq_buf) tq_buf[b] = tq;
  }
  red[threadIdx.x] = part;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_out = red[0] / (float)B;
}

// torch.optim.Adam (_single_tensor_adam) + the target sync of model_torch.py:124-127 (the check runs BEFORE train_count += 1)
__global__ void __launch_bounds__(256) imageq_adam_kernel(const srlx_imageq q) {
  const uint64_t tc = q.counters[0];
  const double t = (double)(q.counters[1] + 1);
  const float step_size = (float)(q.lr / (1.0 - pow(q.adam_beta1, t))), bc2_sqrt = (float)sqrt(1.0 - pow(q.adam_beta2, t));
  const float b1 = (float)q.adam_beta1, b2 = (float)q.adam_beta2, eps = (float)q.adam_eps;
  const bool sync = tc % (uint64_t)q.target_update_interval == 0;
  const int h_off = q.dense_off[q.n_dense - 1], H = q.duel_hidden, h_ld = 2 * H + 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < q.n_params; i += gridDim.x * blockDim.x) {
    if (q.dueling && i >= h_off) {  // the off-branch blocks of the dueling output layer are not parameters
      const int r = (i - h_off) / h_ld, c = (i - h_off) - r * h_ld;
      if (r == 0 ? (c >= H && c < 2 * H) : c < H) continue;
    }
    const float g = q.grads[i];
    float p = q.params[i], m = q.adam_m[i], v = q.adam_v[i];
    m = m + (g - m) * (1.0f - b1);
    v = v * b2 + (1.0f - b2) * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - step_size * (m / denom);
    q.params[i] = p;
    q.adam_m[i] = m;
    q.adam_v[i] = v;
    if (sync) q.target[i] = p;
  }
}
__global__ void imageq_count_kernel(const srlx_imageq q) {
  if (q.counters[0] % (uint64_t)q.target_update_interval == 0) q.counters[2] += 1;
  q.counters[0] += 1;
  q.counters[1] += 1;
}

// dueling_network.py:51-58: y = [V, Adv_0..Adv_{A-1}] -> Q = V + Adv - mean(Adv) | - max(Adv) | nothing
__global__ void duel_combine_kernel(const float* __restrict__ y, const int n, const int A, const int kind, float* __restrict__ qout) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
    const float* yb = y + (size_t)b * (1 + A);
    float sub = 0.f;
    if (kind == SRLX_DUEL_AVERAGE) {
      for (int a = 0; a < A; ++a) sub += yb[1 + a];
      sub /= (float)A;
    } else if (kind == SRLX_DUEL_MAX) {
      sub = yb[1];
      for (int a = 1; a < A; ++a) sub = fmaxf(sub, yb[1 + a]);
    }
    for (int a = 0; a < A; ++a) qout[(size_t)b * A + a] = yb[0] + yb[1 + a] - sub;
  }
}
// its backward: dV = sum_a dQ_a; dAdv_a = dQ_a - mean_a dQ (average) | dQ_a - [a == argmax Adv] sum_a dQ (max) | dQ_a (naive)
__global__ void duel_backward_kernel(const float* __restrict__ dq, const float* __restrict__ y, const int n, const int A, const int kind,
                                     float* __restrict__ dy) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
    const float* d = dq + (size_t)b * A;
    const float* yb = y + (size_t)b * (1 + A);
    float sum = 0.f;
    for (int a = 0; a < A; ++a) sum += d[a];
    int arg = 0;
    for (int a = 1; a < A; ++a)
      if (yb[1 + a] > yb[1 + arg]) arg = a;  // torch.max: the first maximum takes the gradient
    float* o = dy + (size_t)b * (1 + A);
    o[0] = sum;
    for (int a = 0; a < A; ++a)
      o[1 + a] = kind == SRLX_DUEL_AVERAGE ? d[a] - sum / (float)A : (kind == SRLX_DUEL_MAX ? d[a] - (a == arg ? sum : 0.f) : d[a]);
  }
}

__global__ void fill_kernel(float* p, const long long n, const long long stride, const float v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i * stride] = v;
}

// ---- workspace plan -----------------------------------------------------------------------------------------------------------------
struct ImageQPlan {
  ConvG g[SRLX_MAX_CONV];
  long long rows[SRLX_MAX_CONV];          // per sample: OH * OW
  size_t cact[SRLX_MAX_CONV], dcact[SRLX_MAX_CONV], dcol;
  size_t act[SRLX_MAX_LAYERS], dact[SRLX_MAX_LAYERS];
  size_t ones, one4, qbuf, dq, tq, split, in_f32, y, dy;
  size_t cact2[SRLX_MAX_CONV], act2[SRLX_MAX_LAYERS], y2, split2, in_f32b;  // second and third forward sets: the two passes over n_state run
  size_t cact3[SRLX_MAX_CONV], act3[SRLX_MAX_LAYERS], y3, split3;           // beside the online pass over state
  size_t split_floats, total;
  int flat;                               // inputs of dense 0
};

static int imageq_plan(const srlx_imageq* q, ImageQPlan& pl) {
  SRLX_REQUIRE(q->n_conv >= 1 && q->n_conv <= SRLX_MAX_CONV && q->n_dense >= 1 && q->n_dense <= SRLX_MAX_LAYERS, "imageq: 1..%d conv layers, 1..%d dense layers", SRLX_MAX_CONV, SRLX_MAX_LAYERS);
  const int duel = q->dueling != SRLX_DUEL_NONE;
  SRLX_REQUIRE(q->batch_cap >= 1 && q->n_actions >= 1 && q->dense_out[q->n_dense - 1] == q->n_actions + duel,
               "imageq: the last dense layer has n_actions rows (1 + n_actions with a dueling head)");
  SRLX_REQUIRE(!duel || (q->dueling >= SRLX_DUEL_AVERAGE && q->dueling <= SRLX_DUEL_NAIVE && q->n_dense >= 2 && q->duel_hidden >= 1 &&
                         q->dense_out[q->n_dense - 2] == 2 * q->duel_hidden),
               "imageq: a dueling head is the last two dense layers, [2 * duel_hidden] then [1 + n_actions]");
  const long long B = q->batch_cap;
  size_t off = 0;
  auto take = [&](size_t n) { const size_t o = off; off += (n + 3) / 4 * 4; return o; };
  int C = q->in_c, H = q->in_h, W = q->in_w, off_p = 0;
  size_t dcol_max = 0, split_max = 0;
  for (int l = 0; l < q->n_conv; ++l) {
    ConvG& g = pl.g[l];
    g.C = C; g.H = H; g.W = W; g.k = q->conv_k[l]; g.s = q->conv_s[l]; g.p = q->conv_p[l];
    g.OH = (H + 2 * g.p - g.k) / g.s + 1; g.OW = (W + 2 * g.p - g.k) / g.s + 1; g.K = C * g.k * g.k;
    SRLX_REQUIRE(g.k >= 1 && g.s >= 1 && g.p >= 0 && H + 2 * g.p >= g.k && W + 2 * g.p >= g.k, "imageq: conv layer %d has an empty output", l);
    SRLX_REQUIRE(q->conv_oh[l] == g.OH && q->conv_ow[l] == g.OW && q->conv_off[l] == off_p, "imageq: conv layer %d: conv_oh / conv_ow / conv_off do not follow from the shapes (%d x %d at %d expected)", l, g.OH, g.OW, off_p);
    if (l == 0) { g.sb = q->in_sb; g.sc = q->in_sc; g.sh = q->in_sh; g.sw = q->in_sw; }
    else { g.sc = 1; g.sw = C; g.sh = (long long)C * W; g.sb = (long long)C * W * H; }
    g.c_fast = g.sc == 1;
    pl.rows[l] = (long long)g.OH * g.OW;
    const int F = q->conv_f[l];
    pl.cact[l] = take((size_t)B * pl.rows[l] * F);
    pl.cact2[l] = take((size_t)B * pl.rows[l] * F);
    pl.cact3[l] = take((size_t)B * pl.rows[l] * F);
    pl.dcact[l] = take((size_t)B * pl.rows[l] * F);
    if (l > 0 && (size_t)B * pl.rows[l] * g.K > dcol_max) dcol_max = (size_t)B * pl.rows[l] * g.K;
    if ((size_t)F * (g.K + 1) > split_max) split_max = (size_t)F * (g.K + 1);
    off_p += F * (g.K + 1);
    C = F; H = g.OH; W = g.OW;
  }
  pl.flat = C * H * W;
  if ((size_t)B * pl.flat > split_max && (size_t)B * pl.flat <= ((size_t)1 << 21)) split_max = (size_t)B * pl.flat;
  pl.dcol = take(dcol_max);
  int k = pl.flat;
  for (int l = 0; l < q->n_dense; ++l) {
    SRLX_REQUIRE(q->dense_k[l] == k && q->dense_off[l] == off_p, "imageq: dense layer %d: dense_k / dense_off do not follow from the shapes (%d at %d expected)", l, k, off_p);
    const int out = q->dense_out[l], last = l == q->n_dense - 1;
    pl.act[l] = take((size_t)B * (out + (last ? 0 : 1)));
    pl.act2[l] = take((size_t)B * (out + (last ? 0 : 1)));
    pl.act3[l] = take((size_t)B * (out + (last ? 0 : 1)));
    pl.dact[l] = take((size_t)B * out);
    if ((size_t)B * out > split_max) split_max = (size_t)B * out;  // forward / input-gradient maps of a small batch over a long reduction
    if (l > 0 && (size_t)out * (k + 1) > split_max) split_max = (size_t)out * (k + 1);
    off_p += out * (k + 1);
    k = out;
  }
  SRLX_REQUIRE(q->n_params == off_p, "imageq: n_params = %d, the layers hold %d", q->n_params, off_p);
  pl.ones = take((size_t)B);
  pl.one4 = take(4);  // {1, 0, 0, 0}: the source of the bias column's 16-byte chunk
  pl.qbuf = take((size_t)3 * B * q->n_actions);
  pl.dq = take((size_t)B * q->n_actions);
  pl.tq = take((size_t)B + 4);
  pl.split_floats = std::min<size_t>(64 * split_max, (size_t)16 << 20);  // launch_igemm takes fewer slices when they do not fit
  pl.split = take(pl.split_floats);
  pl.split2 = take(pl.split_floats);
  pl.split3 = take(pl.split_floats);
  pl.in_f32 = take(q->in_u8 ? (size_t)B * q->in_sb : 0);
  pl.in_f32b = take(q->in_u8 ? (size_t)B * q->in_sb : 0);
  pl.y = take(duel ? (size_t)B * (1 + q->n_actions) : 0);
  pl.y2 = take(duel ? (size_t)B * (1 + q->n_actions) : 0);
  pl.y3 = take(duel ? (size_t)B * (1 + q->n_actions) : 0);
  pl.dy = take(duel ? (size_t)B * (1 + q->n_actions) : 0);
  pl.total = off;
  return 0;
}

static unsigned grid_for(long long n) { const long long g = (n + 255) / 256; return (unsigned)(g < 148 * 8 ? (g < 1 ? 1 : g) : 148 * 8); }

// uint8 frames -> "0to1" floats once per pass (the tile loaders copy 4-byte words straight into shared memory); float states pass through
static const float* imageq_input(const srlx_imageq* q, const ImageQPlan& pl, const void* state, int n, int which, cudaStream_t s) {
  if (!q->in_u8) return (const float*)state;
  float* dst = q->ws + (which ? pl.in_f32b : pl.in_f32);
  const long long n_in = (long long)n * q->in_sb;
  u8_to_f32_kernel<<<grid_for(n_in), 256, 0, s>>>((const unsigned char*)state, n_in, q->in_max_val, dst);
  count_launch();
  return dst;
}

// forward of n samples (float states `in`) with parameter buffer P on stream s, in forward-buffer set `set`; leaves cact / act / y of the
// pass in that set, Q in qout [n][A]
static int imageq_forward(const srlx_imageq* q, const ImageQPlan& pl, const float* P, const float* in, int n, float* qout, cudaStream_t s, int set) {
  float* ws = q->ws;
  float* sws = ws + (set == 0 ? pl.split : (set == 1 ? pl.split2 : pl.split3));
  const size_t* cact = set == 0 ? pl.cact : (set == 1 ? pl.cact2 : pl.cact3);
  const size_t* act = set == 0 ? pl.act : (set == 1 ? pl.act2 : pl.act3);
  const size_t ybuf = set == 0 ? pl.y : (set == 1 ? pl.y2 : pl.y3);
  for (int l = 0; l < q->n_conv; ++l) {
    const ConvG& g = pl.g[l];
    IGemmP c{};
    c.cv = g; c.gather = 1; c.one = ws + pl.one4;
    c.src = l == 0 ? in : ws + cact[l - 1];
    GemmP& p = c.g;
    p.B = P + q->conv_off[l]; p.sb_k = 1; p.sb_n = g.K + 1;
    p.C = ws + cact[l]; p.ldc = q->conv_f[l];
    p.M = (int)((long long)n * pl.rows[l]); p.N = q->conv_f[l]; p.K = g.K + 1; p.relu = 1;
    if (launch_igemm(c, s, sws, pl.split_floats)) return -1;
  }
  for (int l = 0; l < q->n_dense; ++l) {
    const int k = q->dense_k[l], out = q->dense_out[l], last = l == q->n_dense - 1;
    const float* Wl = P + q->dense_off[l];
    IGemmP d{};
    GemmP& p = d.g;
    p.B = Wl; p.sb_k = 1; p.sb_n = k + 1;
    p.C = last ? (q->dueling ? ws + ybuf : qout) : ws + act[l]; p.ldc = last ? out : out + 1;
    p.M = n; p.N = out;
    if (l == 0) {  // the flattened conv output has no constant column: bias as a second, K = 1 map against the ones vector
      p.A = ws + cact[q->n_conv - 1]; p.sa_m = k; p.sa_k = 1; p.K = k;
      if (launch_igemm(d, s, sws, pl.split_floats)) return -1;
      IGemmP b = d;
      b.g.A = ws + pl.ones; b.g.sa_m = 1; b.g.sa_k = 1; b.g.B = Wl + k; b.g.K = 1; b.g.accumulate = 1; b.g.relu = !last;
      if (launch_igemm(b, s, sws, pl.split_floats)) return -1;
    } else {
      p.A = ws + act[l - 1]; p.sa_m = k + 1; p.sa_k = 1; p.K = k + 1; p.relu = !last;
      if (launch_igemm(d, s, sws, pl.split_floats)) return -1;
    }
  }
  if (q->dueling) {
    duel_combine_kernel<<<grid_for(n), 256, 0, s>>>(ws + ybuf, n, q->n_actions, q->dueling, qout);
    count_launch();
  }
  return 0;
}

// side stream + fork / join events, one set per device (created on first use, kept for the life of the process)
struct ImageAux { cudaStream_t side, side2; cudaEvent_t fork, join, join2, evb[SRLX_MAX_CONV + SRLX_MAX_LAYERS]; bool ok; };
static const bool g_image_one_stream = getenv("SRLX_IMAGE_ONE_STREAM") != nullptr;
static ImageAux* image_aux() {
  static ImageAux aux[64] = {};
  static std::mutex mu;  // first use from two host threads at once
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  ImageAux& a = aux[dev];
  if (!a.ok) {
    if (cudaStreamCreateWithFlags(&a.side, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&a.side2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&a.join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&a.join2, cudaEventDisableTiming) != cudaSuccess)
      return nullptr;
    for (int i = 0; i < SRLX_MAX_CONV + SRLX_MAX_LAYERS; ++i)
      if (cudaEventCreateWithFlags(&a.evb[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    a.ok = true;
  }
  return &a;
}

static int imageq_check(const srlx_imageq* q, ImageQPlan& pl) {
  SRLX_REQUIRE(q != nullptr, "imageq: null handle");
  if (int rc = imageq_plan(q, pl)) return rc;
  SRLX_REQUIRE(q->params && q->target && q->ws && q->ws_floats >= pl.total, "imageq: params / target / ws missing or ws too small (%llu floats needed)", (unsigned long long)pl.total);
  return 0;
}

}  // namespace srlx

using namespace srlx;

extern "C" {

int srlx_sgemm_tc3(const float* a, long long sa_m, long long sa_k, const float* b, long long sb_k, long long sb_n, float* c, long long ldc, int M,
                   int N, int K, int relu, int accumulate, float* ws, uint64_t ws_floats, uintptr_t stream) {
  SRLX_REQUIRE(a && b && c, "srlx_sgemm_tc3: NULL buffer");
  SRLX_REQUIRE(M >= 0 && N >= 0 && K >= 0, "srlx_sgemm_tc3: negative size");
  IGemmP q{};
  GemmP& g = q.g;
  g.A = a; g.sa_m = sa_m; g.sa_k = sa_k; g.B = b; g.sb_k = sb_k; g.sb_n = sb_n; g.C = c; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.relu = relu; g.accumulate = accumulate; g.gate = Gate{nullptr, 0};
  q.one = a;
  if (int rc = launch_t3(q, (cudaStream_t)stream, ws, ws_floats)) return rc;
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

size_t srlx_sizeof_imageq(void) { return sizeof(srlx_imageq); }

int srlx_image_linear_table(int32_t dst, int32_t src, int border_reset, int32_t* idx, int32_t* coef) {
  SRLX_REQUIRE(dst > 0 && src > 0 && idx && coef, "image table: dst, src > 0 and two output arrays");
  const double scale = 1.0 / ((double)dst / (double)src);  // resize.cpp: inv_scale = dsize / ssize, scale = 1 / inv_scale
  for (int d = 0; d < dst; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (border_reset) {
      if (s < 0) { s = 0; f = 0.f; }
      if (s >= src - 1) { s = src - 1; f = 0.f; }
    }
    idx[d] = s;
    coef[2 * d] = (int32_t)lrintf((1.f - f) * 2048.f);  // saturate_cast<short>(float) = cvRound
    coef[2 * d + 1] = (int32_t)lrintf(f * 2048.f);
  }
  return 0;
}

int srlx_image_process(const srlx_image_proc* p, const unsigned char* src, uint32_t n, void* out, uint64_t out_stride, uintptr_t stream) {
  SRLX_REQUIRE(p && src && out, "image_process: null argument");
  SRLX_REQUIRE((p->src_c == 1 || p->src_c == 3) && (p->out_c == 1 || p->out_c == 3), "image_process: 1 or 3 channels (got %d -> %d)", p->src_c, p->out_c);
  SRLX_REQUIRE(p->top >= 0 && p->left >= 0 && p->trim_h >= 1 && p->trim_w >= 1 && p->top + p->trim_h <= p->src_h && p->left + p->trim_w <= p->src_w,
               "image_process: trimming window outside the frame");
  SRLX_REQUIRE(p->resize ? (p->x_idx && p->x_coef && p->y_idx && p->y_coef) : (p->out_h == p->trim_h && p->out_w == p->trim_w),
               "image_process: resize needs the four tables; without resize the output is the trimming window");
  SRLX_REQUIRE(p->normalize >= 0 && p->normalize <= 2 && out_stride >= (uint64_t)p->out_h * p->out_w * p->out_c, "image_process: normalize in 0..2, frame stride >= frame size");
  if (n == 0) return 0;
  const long long total = (long long)n * p->out_h * p->out_w * p->out_c;
  cudaStream_t s = (cudaStream_t)stream;
  // staged path: whole frame (gray bytes when the output is gray) in shared memory, 16-byte loads
  const int to_gray = p->src_c == 3 && p->out_c == 1;
  const size_t frame_bytes = (size_t)p->src_h * p->src_w * p->src_c, staged = to_gray ? frame_bytes / 3 : frame_bytes;
  const size_t smem = (((size_t)(3 * p->out_w + 3 * p->out_h) * 4 + 15) / 16) * 16 + staged;
  static const bool no_staged = getenv("SRLX_IMAGE_SCALAR") != nullptr;  // diagnostic: the one-thread-per-output kernel for every shape
  const bool can_stage = !no_staged && frame_bytes % 16 == 0 && (!to_gray || ((size_t)p->src_h * p->src_w) % 16 == 0) && smem <= 200 * 1024 &&
                         ((uintptr_t)src & 15) == 0;
  if (can_stage) {
    const int sc = to_gray ? 1 : p->src_c;
    auto launch = [&](auto kern, auto* o) -> int {
      int per_sm = 0;
      SRLX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      SRLX_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
      const unsigned grid = (unsigned)std::min<long long>(n, 148LL * std::max(per_sm, 1));
      kern<<<grid, 256, smem, s>>>(*p, src, n, o, out_stride, to_gray);
      return 0;
    };
    auto pick = [&](auto* o) -> int {
      using T = std::remove_pointer_t<decltype(o)>;
      const bool rz = p->resize != 0;
      if (p->out_c == 1) return rz ? launch(image_process_staged_kernel<T, 1, 1, true>, o) : launch(image_process_staged_kernel<T, 1, 1, false>, o);
      if (sc == 1) return rz ? launch(image_process_staged_kernel<T, 3, 1, true>, o) : launch(image_process_staged_kernel<T, 3, 1, false>, o);
      return rz ? launch(image_process_staged_kernel<T, 3, 3, true>, o) : launch(image_process_staged_kernel<T, 3, 3, false>, o);
    };
    if (int rc = p->normalize == 0 ? pick((unsigned char*)out) : pick((float*)out)) return rc;
  } else if (p->normalize == 0) image_process_kernel<unsigned char><<<grid_for(total), 256, 0, s>>>(*p, src, n, (unsigned char*)out, out_stride);
  else image_process_kernel<float><<<grid_for(total), 256, 0, s>>>(*p, src, n, (float*)out, out_stride);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

uint64_t srlx_imageq_ws_floats(const srlx_imageq* q) {
  ImageQPlan pl;
  if (!q || imageq_plan(q, pl)) return 0;
  return pl.total;
}

int srlx_imageq_init(const srlx_imageq* q, uintptr_t stream) {
  ImageQPlan pl;
  if (int rc = imageq_check(q, pl)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const long long B = q->batch_cap;
  fill_kernel<<<grid_for(B), 256, 0, s>>>(q->ws + pl.ones, B, 1, 1.f);
  fill_kernel<<<1, 32, 0, s>>>(q->ws + pl.one4, 4, 1, 0.f);
  fill_kernel<<<1, 32, 0, s>>>(q->ws + pl.one4, 1, 1, 1.f);
  for (int l = 0; l + 1 < q->n_dense; ++l) {
    fill_kernel<<<grid_for(B), 256, 0, s>>>(q->ws + pl.act[l] + q->dense_out[l], B, q->dense_out[l] + 1, 1.f);
    fill_kernel<<<grid_for(B), 256, 0, s>>>(q->ws + pl.act2[l] + q->dense_out[l], B, q->dense_out[l] + 1, 1.f);
    fill_kernel<<<grid_for(B), 256, 0, s>>>(q->ws + pl.act3[l] + q->dense_out[l], B, q->dense_out[l] + 1, 1.f);
  }
  count_launch(q->n_dense);
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int srlx_imageq_forward(const srlx_imageq* q, int use_target, const void* state, uint32_t n, float* q_out, uintptr_t stream) {
  ImageQPlan pl;
  if (int rc = imageq_check(q, pl)) return rc;
  SRLX_REQUIRE(state && q_out && n >= 1 && n <= (uint32_t)q->batch_cap, "imageq_forward: 1 <= n <= batch_cap (%d), got %u", q->batch_cap, n);
  if (imageq_forward(q, pl, use_target ? q->target : q->params, imageq_input(q, pl, state, (int)n, 0, (cudaStream_t)stream), (int)n, q_out, (cudaStream_t)stream, 0)) return -1;
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int srlx_imageq_train(const srlx_imageq* q, const void* state, const void* n_state, const int32_t* action, const float* reward,
                      const float* undone, const float* weights, uint32_t batch, float* pri_out, float* loss_out, float* tq_out, int phases,
                      uintptr_t stream) {
  ImageQPlan pl;
  if (int rc = imageq_check(q, pl)) return rc;
  SRLX_REQUIRE(q->adam_m && q->adam_v && q->grads && q->counters, "imageq_train: adam_m / adam_v / grads / counters missing");
  SRLX_REQUIRE(batch >= 1 && batch <= (uint32_t)q->batch_cap, "imageq_train: 1 <= batch <= batch_cap (%d), got %u", q->batch_cap, batch);
  SRLX_REQUIRE(phases >= 1 && phases <= 3, "imageq_train: phases in 1..3");
  SRLX_REQUIRE(q->target_update_interval >= 1, "imageq_train: target_update_interval >= 1");
  cudaStream_t s = (cudaStream_t)stream;
  float* ws = q->ws;
  const int B = (int)batch, A = q->n_actions;
  if (phases & 1) {
    SRLX_REQUIRE(state && n_state && action && reward && undone && weights && pri_out && loss_out, "imageq_train: null batch array");
    float* qb = ws + pl.qbuf;
    // calc_target_q: pred_target_q(n_state), pred_q(n_state) for double DQN (dqn.py:154-162), then the online pass whose activations the
    // backward pass reads
    // The three forward passes are independent until the loss: the two over n_state run on side streams in their own forward-buffer
    // sets beside the online pass over state (at batch 32 no map fills the GPU, so the passes overlap); SRLX_IMAGE_ONE_STREAM=1 keeps
    // one stream.
    const float* in_n = imageq_input(q, pl, n_state, B, 1, s);
    const float* in_s = imageq_input(q, pl, state, B, 0, s);
    ImageAux* ax = g_image_one_stream ? nullptr : image_aux();
    if (ax) {
      SRLX_CHECK_CUDA(cudaEventRecord(ax->fork, s));
      SRLX_CHECK_CUDA(cudaStreamWaitEvent(ax->side, ax->fork, 0));
      if (imageq_forward(q, pl, q->target, in_n, B, qb + (size_t)2 * B * A, ax->side, 1)) return -1;
      SRLX_CHECK_CUDA(cudaEventRecord(ax->join, ax->side));
      if (q->enable_double_dqn) {
        SRLX_CHECK_CUDA(cudaStreamWaitEvent(ax->side2, ax->fork, 0));
        if (imageq_forward(q, pl, q->params, in_n, B, qb + (size_t)B * A, ax->side2, 2)) return -1;
        SRLX_CHECK_CUDA(cudaEventRecord(ax->join2, ax->side2));
      }
    } else {
      if (imageq_forward(q, pl, q->target, in_n, B, qb + (size_t)2 * B * A, s, 1)) return -1;
      if (q->enable_double_dqn && imageq_forward(q, pl, q->params, in_n, B, qb + (size_t)B * A, s, 2)) return -1;
    }
    if (imageq_forward(q, pl, q->params, in_s, B, qb, s, 0)) return -1;
    if (ax) {
      SRLX_CHECK_CUDA(cudaStreamWaitEvent(s, ax->join, 0));
      if (q->enable_double_dqn) SRLX_CHECK_CUDA(cudaStreamWaitEvent(s, ax->join2, 0));
    }
    imageq_loss_kernel<<<1, 256, 0, s>>>(*q, qb, qb + (size_t)B * A, qb + (size_t)2 * B * A, action, reward, undone, weights, B, ws + pl.dq, pri_out,
                                         loss_out, tq_out);
    count_launch();
    // dense layers, last to first.  A layer's weight gradient (dW = dOut^T x X) is off the critical path -- only the input gradient feeds
    // the next layer down -- so every dW map goes to the side stream (own split-K workspace), forked after its dOut is complete and joined
    // before the optimiser step.
    float* sws = ws + pl.split;
    float* sws_w = ws + (ax ? pl.split2 : pl.split);
    int n_fork = 0;
    auto dw_stream = [&]() -> cudaStream_t {
      if (!ax) return s;
      cudaEventRecord(ax->evb[n_fork], s);
      cudaStreamWaitEvent(ax->side, ax->evb[n_fork], 0);
      ++n_fork;
      return ax->side;
    };
    const float* dout = ws + pl.dq;
    int ld_dout = A;
    if (q->dueling) {  // pl.y still holds [V, Adv] of the online pass on `state` (the last forward)
      duel_backward_kernel<<<grid_for(B), 256, 0, s>>>(ws + pl.dq, ws + pl.y, B, A, q->dueling, ws + pl.dy);
      count_launch();
      dout = ws + pl.dy;
      ld_dout = 1 + A;
    }
    for (int l = q->n_dense - 1; l >= 0; --l) {
      const int k = q->dense_k[l], out = q->dense_out[l];
      const float* X = l == 0 ? ws + pl.cact[q->n_conv - 1] : ws + pl.act[l - 1];
      const int ldx = l == 0 ? k : k + 1;
      IGemmP wq{};  // dW[out][k (+1)] = dOut^T x X
      GemmP& w = wq.g;
      w.A = dout; w.sa_m = 1; w.sa_k = ld_dout;
      w.B = X; w.sb_k = ldx; w.sb_n = 1;
      w.C = q->grads + q->dense_off[l]; w.ldc = k + 1;
      w.M = out; w.N = l == 0 ? k : k + 1; w.K = B;
      const cudaStream_t sw = dw_stream();
      if (launch_igemm(wq, sw, sws_w, pl.split_floats)) return -1;
      if (l == 0) {
        IGemmP bq = wq;
        bq.g.B = ws + pl.ones; bq.g.sb_k = 1; bq.g.sb_n = 1; bq.g.C = q->grads + q->dense_off[l] + k; bq.g.N = 1;
        if (launch_igemm(bq, sw, sws_w, pl.split_floats)) return -1;
      }
      IGemmP xq{};  // dX[B][k] = (X > 0) * dOut x W[:, :k]
      GemmP& x = xq.g;
      x.A = dout; x.sa_m = ld_dout; x.sa_k = 1;
      x.B = q->params + q->dense_off[l]; x.sb_k = k + 1; x.sb_n = 1;
      x.C = l == 0 ? ws + pl.dcact[q->n_conv - 1] : ws + pl.dact[l - 1]; x.ldc = k;
      x.M = B; x.N = k; x.K = out; x.mask = X; x.ldmask = ldx;
      if (launch_igemm(xq, s, sws, pl.split_floats)) return -1;
      dout = x.C;
      ld_dout = k;
    }
    // conv layers, last to first
    for (int l = q->n_conv - 1; l >= 0; --l) {
      const ConvG& g = pl.g[l];
      const int F = q->conv_f[l];
      const long long rows = (long long)B * pl.rows[l];
      IGemmP wq{};  // dW[F][K+1] = dOut^T x im2col(input of the layer), the im2col matrix gathered by the tile loader
      wq.cv = g; wq.gather = 2;
      wq.src = l == 0 ? in_s : ws + pl.cact[l - 1]; wq.one = ws + pl.one4;
      GemmP& w = wq.g;
      w.A = ws + pl.dcact[l]; w.sa_m = 1; w.sa_k = F;
      w.C = q->grads + q->conv_off[l]; w.ldc = g.K + 1;
      w.M = F; w.N = g.K + 1; w.K = (int)rows;
      if (launch_igemm(wq, dw_stream(), sws_w, pl.split_floats)) return -1;
      if (l == 0) break;
      IGemmP xq{};  // dcol[rows][K] = dOut x W[:, :K]
      GemmP& x = xq.g;
      x.A = ws + pl.dcact[l]; x.sa_m = F; x.sa_k = 1;
      x.B = q->params + q->conv_off[l]; x.sb_k = g.K + 1; x.sb_n = 1;
      x.C = ws + pl.dcol; x.ldc = g.K;
      x.M = (int)rows; x.N = g.K; x.K = F;
      if (launch_igemm(xq, s, sws, pl.split_floats)) return -1;
      const long long n_in = (long long)B * g.H * g.W * g.C;
      col2im_kernel<<<grid_for(n_in), 256, 0, s>>>(g, ws + pl.dcol, ws + pl.cact[l - 1], ws + pl.dcact[l - 1], n_in);
      count_launch();
    }
    if (ax) {  // the gradient is complete once the side stream has drained
      SRLX_CHECK_CUDA(cudaEventRecord(ax->join, ax->side));
      SRLX_CHECK_CUDA(cudaStreamWaitEvent(s, ax->join, 0));
    }
  }
  if (phases & 2) {
    imageq_adam_kernel<<<grid_for(q->n_params), 256, 0, s>>>(*q);
    imageq_count_kernel<<<1, 1, 0, s>>>(*q);
    count_launch(2);
  }
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
