// gemm_tc3.cuh -- fp32-accurate GEMM tiles on the 5th-generation tensor cores: 3 x TF32 through tcgen05.mma.kind::tf32 with the
// accumulator in tensor memory (sm_100a).
//
// Every operand element x is split into hi = tf32(x) and lo = tf32(x - hi) and a product is formed as lo*hi + hi*lo + hi*hi (fp32
// accumulate): the dropped lo*lo term is 2^-22 relative, so the result sits within a few fp32 ulps of an FMA chain -- the 1e-4 parity
// bar of the trainers holds with two orders of magnitude to spare.  The legacy mma.sync path runs TF32 at about the FMA rate on this
// chip (measured: profiles/r3e_*); tcgen05 is the only way to the tensor cores' real throughput.
//
// One CTA = one 128 x BN output tile (BN = 32 / 64), 256 threads, all of them PRODUCERS: operands are not plain matrices (strided views,
// and the im2col matrix of a convolution gathered on the fly with replicate padding), so instead of TMA the threads copy their fixed
// tile slots with cp.async (16-byte chunks where 4 consecutive k are contiguous, else 4-byte words), one k slice (32 columns) ahead,
// straight into the canonical K-major 128-byte-swizzle layout (row r, 16-byte chunk c at r * 128 + ((c ^ (r & 7)) << 4)); when
// a slice has landed each thread splits its own slots in place (hi) and into a second tile (lo).  fence.proxy.async + barrier, then
// ONE thread issues 4 k-steps x 3 tcgen05.mma (M = 128, N = BN, K = 8) and commits the stage's "empty" mbarrier.  Two stages of
// (hi + lo) tiles = 97 KB and 256 TMEM columns per CTA, so TWO CTAs share an SM and cover each other's copy latency (measured: 1.00 ->
// 0.91 ms per update against one CTA with a 4-stage ring, batch-256 forward 0.56 -> 0.40 ms).  Epilogue: tcgen05.ld (32 lanes x 32 columns) of every k chunk's
// accumulator (see t3_gemm_kernel), fp32 sum, then the same accumulate / ReLU / mask epilogue as gemm.cuh, or split-K partials
// (ws[z][M][N], summed in slice order by splitk_reduce_kernel).
// The larger of (M, N) rides on the 128-row side; element addresses of C / mask / partials are (row_a * c_a + row_b * c_b).
// Descriptor encodings follow cute/arch/mma_sm100_desc.hpp as in qnet_tc.cu; every mbarrier wait is bounded (trap, never hang).
#pragma once
#include "gemm.cuh"

namespace srlx {

struct ConvG {
  int C, H, W, k, s, p, OH, OW, K;       // K = C * k * k
  long long sb, sc, sh, sw;              // element strides of the source
  int c_fast;                            // column order (kh, kw, c) instead of (c, kh, kw)
};

__device__ __forceinline__ void col_split(const ConvG& g, int j, int& c, int& kh, int& kw) {
  if (g.c_fast) { c = j % g.C; j /= g.C; kw = j % g.k; kh = j / g.k; }
  else { kw = j % g.k; j /= g.k; kh = j % g.k; c = j / g.k; }
}

#ifndef T3_STAGES_N
#define T3_STAGES_N 2
#endif
// T3_STAGES / T3_AHEAD: shared-memory ring depth and how many k slices the copies run ahead; T3_TMEM: accumulator columns per CTA
// (2 stages = 97 KB and 256 columns let TWO CTAs share an SM: 16 warps to hide the copy latency instead of 8)
constexpr int T3_BM = 128, T3_BK = 32, T3_STAGES = T3_STAGES_N, T3_AHEAD = T3_STAGES_N > 2 ? 2 : 1, T3_TMEM = T3_STAGES_N > 2 ? 512 : 256, T3_THREADS = 256;

struct T3Op {      // one operand as a source of K-major tiles: element (r, k)
  const float* ptr;
  int mode;        // 0: strided view; 1: im2col, rows = output positions, k = column; 2: im2col, rows = columns, k = output positions
  int s_row, s_k;  // mode 0: element strides (32-bit: the launcher checks the extents)
  int rows;
};
struct T3P {
  T3Op a, b;       // a: the 128-row side, b: the BN-row side
  ConvG cv;
  const float* one;
  int K, ksplit, klen;
  float* C; long long c_a, c_b;
  const float* mask; long long m_a, m_b;
  float* ws; long long w_a, w_b, w_slice;
  int relu, accumulate;
};

__device__ __forceinline__ uint32_t t3_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t3_mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t3_smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void t3_mbar_wait(uint64_t* b, uint32_t parity) {  // bounded: ~1 s, then trap
  const uint32_t a = t3_smem_u32(b);
  const long long t0 = clock64();
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 2000000000ll) __trap();
  }
}
// K-major tile, 128-byte swizzle: start address >> 4, LBO 1 (unused), SBO 1024 >> 4 (8-row groups), version 1, layout SWIZZLE_128B
__device__ __forceinline__ uint64_t t3_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32 instruction descriptor: D = F32 (bit 4), A = B = TF32 (2 at bits 7 and 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t t3_instr_desc(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(T3_BM >> 4) << 24);
}
__device__ __forceinline__ void t3_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void t3_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(t3_smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void t3_cp4(unsigned char* dst, const float* src, int bytes) {  // bytes = 0: zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(t3_smem_u32(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t3_cp16(unsigned char* dst, const float* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(t3_smem_u32(dst)), "l"(src), "r"(bytes) : "memory");
}

// The producer side of one operand.  Raw fp32 words go global -> shared memory with cp.async (two k slices ahead, no register staging),
// straight into their place in the swizzled K-major "hi" tile; when a slice has landed every thread revisits ITS OWN slots, replaces x
// by hi = tf32(x) in place and writes lo = tf32(x - hi) to the "lo" tile (no barrier in between: a thread only touches what it copied).
// Slot maps (ROWS rows x 32 k per slice, 256 threads):
//   vec  (modes 0, 1 when 4 consecutive k are 16 contiguous, aligned bytes): chunk = tid & 7, rows (tid >> 3) + 32 i: 16-byte copies
//   k-fast (modes 0, 1 otherwise): k = lane, rows warp + 8 i: a warp copies 32 consecutive k of one row
//   row-fast (mode 2): row = lane + 32 g, k = warp + 8 g4: a warp copies 32 consecutive columns of the im2col matrix (contiguous
//        channels of one pixel) at one output position
// Everything that does not change along k is decoded once; positions and columns advance as (ow, oh, image) / (c, kw, kh) counters.
template <int ROWS>
struct T3Producer {
  static constexpr int NS = ROWS / 8, RG = ROWS / 32;
  const T3Op& op;
  const ConvG& cv;
  const float* one;  // {1, 0, 0, 0} in global memory, 16-byte aligned
  int ke, lane, warp, row0;
  bool vec;
  int off0, kidx;                              // mode 0: element offset of slot 0, its k
  int kc, kkw, kkh, kg, pw, ph, pimg, prow;    // mode 1: this thread's column kg = (kkh, kkw, kc), its first position; mode 2: position of k = warp
  int jc[RG], jkh[RG], jkw[RG];                // mode 2: the row groups' columns

  __device__ __forceinline__ T3Producer(const T3Op& op_, const ConvG& cv_, const float* one_, int row0_, int kb, int ke_)
      : op(op_), cv(cv_), one(one_), ke(ke_), lane(threadIdx.x & 31), warp(threadIdx.x >> 5), row0(row0_) {
    off0 = kidx = kc = kkw = kkh = kg = pw = ph = pimg = prow = 0;
    vec = false;
    if (op.mode == 0) {
      vec = op.s_k == 1 && (op.s_row & 3) == 0 && (((uintptr_t)op.ptr) & 15) == 0;
      const int r = vec ? (int)(threadIdx.x >> 3) : warp, k = vec ? 4 * (int)(threadIdx.x & 7) : lane;
      off0 = (row0 + r) * op.s_row + (kb + k) * op.s_k;
      kidx = kb + k;
    } else if (op.mode == 1) {
      vec = cv.c_fast && (cv.C & 3) == 0 && cv.sc == 1 && (cv.sw & 3) == 0 && (cv.sh & 3) == 0 && (cv.sb & 3) == 0 && (((uintptr_t)op.ptr) & 15) == 0;
      const int r = vec ? (int)(threadIdx.x >> 3) : warp, k = vec ? 4 * (int)(threadIdx.x & 7) : lane;
      kg = kb + k;
      if (kg < cv.K) col_split(cv, kg, kc, kkh, kkw);
      const long long row = row0 + r;
      pw = (int)(row % cv.OW); ph = (int)((row / cv.OW) % cv.OH);
      pimg = (int)((row / ((long long)cv.OW * cv.OH)) * cv.sb);
    } else {
#pragma unroll
      for (int g = 0; g < RG; ++g) {
        const int j = row0 + lane + 32 * g;
        jc[g] = j < cv.K ? 0 : (j == cv.K ? -1 : -2);
        jkh[g] = jkw[g] = 0;
        if (j < cv.K) col_split(cv, j, jc[g], jkh[g], jkw[g]);
      }
      const long long row = kb + warp;
      prow = (int)row;
      pw = (int)(row % cv.OW); ph = (int)((row / cv.OW) % cv.OH);
      pimg = (int)((row / ((long long)cv.OW * cv.OH)) * cv.sb);
    }
  }
  __device__ __forceinline__ void adv_pos(int& w, int& h, int& img, int n) const {
    w += n;
    while (w >= cv.OW) { w -= cv.OW; if (++h == cv.OH) { h = 0; img += (int)cv.sb; } }
  }
  __device__ __forceinline__ void adv_col(int n) {
    kg += n;
    if (cv.c_fast) {
      kc += n;
      while (kc >= cv.C) { kc -= cv.C; if (++kkw == cv.k) { kkw = 0; ++kkh; } }
    } else {
      kkw += n;
      while (kkw >= cv.k) { kkw -= cv.k; if (++kkh == cv.k) { kkh = 0; ++kc; } }
    }
  }
  // shared-memory offsets of this thread's slots
  __device__ __forceinline__ uint32_t vec_off(int i) const {
    const uint32_t r = threadIdx.x >> 3, ch = threadIdx.x & 7;
    return (r + 32u * i) * 128u + ((ch ^ (r & 7u)) << 4);
  }
  __device__ __forceinline__ uint32_t kfast_off(int i) const {
    return ((uint32_t)warp + 8u * i) * 128u + ((((uint32_t)lane >> 2) ^ ((uint32_t)warp & 7u)) << 4) + (((uint32_t)lane & 3u) << 2);
  }
  __device__ __forceinline__ uint32_t rfast_off(int g4, int g) const {
    const uint32_t kk = (uint32_t)warp + 8u * g4, row = (uint32_t)lane + 32u * g;
    return row * 128u + (((kk >> 2) ^ (row & 7u)) << 4) + ((kk & 3u) << 2);
  }
  // copy the next k slice into the tile at `hi` (asynchronously) and advance
  __device__ __forceinline__ void issue(unsigned char* hi) {
    const int csc = (int)cv.sc, csh = (int)cv.sh, csw = (int)cv.sw;
    if (op.mode == 0) {
      const bool in_k = kidx < ke;
      if (vec) {
        const int r = threadIdx.x >> 3;
#pragma unroll
        for (int i = 0; i < RG; ++i) {
          const bool ok = in_k && row0 + r + 32 * i < op.rows;
          // the last chunk of a row may reach past the operand's k extent (never past the allocation's 16-byte granule: rows are 16-byte
          // multiples): the bytes beyond ke are multiplied by the other operand's zero fill
          t3_cp16(hi + vec_off(i), ok ? op.ptr + off0 + 32 * i * op.s_row : op.ptr, ok ? min(16, 4 * (ke - kidx)) : 0);
        }
      } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          const bool ok = in_k && row0 + warp + 8 * i < op.rows;
          t3_cp4(hi + kfast_off(i), ok ? op.ptr + off0 + 8 * i * op.s_row : op.ptr, ok ? 4 : 0);
        }
      }
      off0 += T3_BK * op.s_k;
      kidx += T3_BK;
    } else if (op.mode == 1) {
      const bool in_k = kg < ke, bias = kg >= cv.K;
      int w = pw, h = ph, img = pimg;
      if (vec) {
        const int r = threadIdx.x >> 3;
#pragma unroll
        for (int i = 0; i < RG; ++i) {
          const float* src = one;
          int bytes = 0;
          if (in_k && kg <= cv.K && row0 + r + 32 * i < op.rows) {
            bytes = 16;
            if (!bias) {
              const int ih = min(max(h * cv.s - cv.p + kkh, 0), cv.H - 1), iw = min(max(w * cv.s - cv.p + kkw, 0), cv.W - 1);
              src = op.ptr + img + kc + ih * csh + iw * csw;
            }
          }
          t3_cp16(hi + vec_off(i), src, bytes);
          adv_pos(w, h, img, 32);
        }
      } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          const float* src = one;
          int bytes = 0;
          if (in_k && kg <= cv.K && row0 + warp + 8 * i < op.rows) {
            bytes = 4;
            if (!bias) {
              const int ih = min(max(h * cv.s - cv.p + kkh, 0), cv.H - 1), iw = min(max(w * cv.s - cv.p + kkw, 0), cv.W - 1);
              src = op.ptr + img + kc * csc + ih * csh + iw * csw;
            }
          }
          t3_cp4(hi + kfast_off(i), src, bytes);
          adv_pos(w, h, img, 8);
        }
      }
      adv_col(T3_BK);
    } else {
      int w = pw, h = ph, img = pimg;
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {  // k = warp + 8 g4
        const bool in_k = prow + 8 * g4 < ke;
        const int oh = h * cv.s - cv.p, ow = w * cv.s - cv.p;
#pragma unroll
        for (int g = 0; g < RG; ++g) {
          const float* src = one;
          int bytes = 0;
          if (in_k && jc[g] != -2) {
            bytes = 4;
            if (jc[g] != -1) {
              const int ih = min(max(oh + jkh[g], 0), cv.H - 1), iw = min(max(ow + jkw[g], 0), cv.W - 1);
              src = op.ptr + img + jc[g] * csc + ih * csh + iw * csw;
            }
          }
          t3_cp4(hi + rfast_off(g4, g), src, bytes);
        }
        adv_pos(w, h, img, 8);
      }
      prow += T3_BK;
      adv_pos(pw, ph, pimg, T3_BK);
    }
  }
  // own slots of a landed slice: x -> hi in place, lo to the second tile
  __device__ __forceinline__ void fixup(unsigned char* hi, unsigned char* lo) const {
    if (vec) {
#pragma unroll
      for (int i = 0; i < RG; ++i) {
        const uint32_t off = vec_off(i);
        const float4 x = *reinterpret_cast<const float4*>(hi + off);
        uint4 h, l;
        split_tf32(x.x, h.x, l.x); split_tf32(x.y, h.y, l.y); split_tf32(x.z, h.z, l.z); split_tf32(x.w, h.w, l.w);
        *reinterpret_cast<uint4*>(hi + off) = h;
        *reinterpret_cast<uint4*>(lo + off) = l;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const uint32_t off = op.mode == 2 ? rfast_off(i / RG, i % RG) : kfast_off(i);
        uint32_t h, l;
        split_tf32(*reinterpret_cast<const float*>(hi + off), h, l);
        *reinterpret_cast<uint32_t*>(hi + off) = h;
        *reinterpret_cast<uint32_t*>(lo + off) = l;
      }
    }
  }
};

// The tensor cores add into the fp32 accumulator by truncation: over n tcgen05.mma the error grows like n * 2^-24 of the running sum
// (measured: K = 4096 in one accumulator is 2.7e-6 of sum |a||b|, 40 x an fp32 FMA chain).  So the k range of a CTA is cut into up to
// T3_TMEM / BN CHUNKS, each with its own accumulator in tensor memory (256 columns per CTA: two CTAs share an SM), and the epilogue
// adds the chunks in fp32 with round-to-nearest.
template <int BN>
__global__ void __launch_bounds__(T3_THREADS, T3_STAGES_N > 2 ? 1 : 2) t3_gemm_kernel(const T3P p) {
  constexpr uint32_t TILE_A = T3_BM * 128, TILE_B = BN * 128, STAGE = 2 * TILE_A + 2 * TILE_B;
  constexpr int NCH = T3_TMEM / BN;
  extern __shared__ unsigned char t3_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)t3_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* empty = reinterpret_cast<uint64_t*>(smem + T3_STAGES * STAGE);
  uint64_t* tmem_full = empty + T3_STAGES;
  uint32_t* tmem_base_p = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int a0 = blockIdx.y * T3_BM, b0 = blockIdx.x * BN;
  const int kb = p.ksplit > 1 ? blockIdx.z * p.klen : 0, ke = p.ksplit > 1 ? min(p.K, kb + p.klen) : p.K;
  const int nk = (ke - kb + T3_BK - 1) / T3_BK;
  const int spc = (nk + NCH - 1) / NCH > 0 ? (nk + NCH - 1) / NCH : 1;  // k slices per chunk
  const int n_chunks = (nk + spc - 1) / spc;

  if (tid == 0) {
    for (int s = 0; s < T3_STAGES; ++s) t3_mbar_init(&empty[s], 1);
    t3_mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(t3_smem_u32(tmem_base_p)), "n"(T3_TMEM) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_p;

  T3Producer<T3_BM> pa(p.a, p.cv, p.one, a0, kb, ke);
  T3Producer<BN> pb(p.b, p.cv, p.one, b0, kb, ke);
  const uint32_t idesc = t3_instr_desc(BN);
  auto stage_ptr = [&](int it) { return smem + (it % T3_STAGES) * STAGE; };
#pragma unroll
  for (int it = 0; it < T3_AHEAD; ++it) {  // T3_AHEAD slices ahead
    if (it < nk) { pa.issue(stage_ptr(it)); pb.issue(stage_ptr(it) + 2 * TILE_A); }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int it = 0; it < nk; ++it) {
    if (it + T3_AHEAD < nk) {
      const int nx = it + T3_AHEAD;
      if (nx >= T3_STAGES) t3_mbar_wait(&empty[nx % T3_STAGES], ((nx / T3_STAGES) - 1) & 1);  // the MMAs that read that stage are done
      pa.issue(stage_ptr(nx));
      pb.issue(stage_ptr(nx) + 2 * TILE_A);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(T3_AHEAD) : "memory");  // slice `it` of THIS thread has landed
    unsigned char* st = stage_ptr(it);
    pa.fixup(st, st + TILE_A);
    pb.fixup(st + 2 * TILE_A, st + 2 * TILE_A + TILE_B);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor cores' async proxy
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ah = t3_smem_u32(st), al = ah + TILE_A, bh = ah + 2 * TILE_A, bl = bh + TILE_B;
      const uint32_t acc = tmem_base + (uint32_t)((it / spc) * BN);
      const bool first = it % spc == 0;
#pragma unroll
      for (int k = 0; k < T3_BK / 8; ++k) {  // K = 8 tf32 = 32 bytes per instruction, inside the 128-byte swizzle atom
        const uint64_t dah = t3_smem_desc(ah + k * 32), dal = t3_smem_desc(al + k * 32), dbh = t3_smem_desc(bh + k * 32),
                       dbl = t3_smem_desc(bl + k * 32);
        t3_mma(acc, dal, dbh, idesc, (first && k == 0) ? 0u : 1u);  // small terms first
        t3_mma(acc, dah, dbl, idesc, 1u);
        t3_mma(acc, dah, dbh, idesc, 1u);
      }
      t3_commit(&empty[it % T3_STAGES]);  // (implies tcgen05.fence::before_thread_sync)
      if (it == nk - 1) t3_commit(tmem_full);
    }
  }
  // ---- epilogue: warp w reads TMEM lane quarter (w & 3); warps w and w + 4 share the quarter's 32-column groups
  if (nk > 0) t3_mbar_wait(tmem_full, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int q = warp & 3;
  const long long ra = a0 + q * 32 + lane;
#pragma unroll 1
  for (int c = warp >> 2; c < BN / 32; c += 2) {
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = 0.f;
#pragma unroll 1
    for (int ch = 0; ch < n_chunks; ++ch) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * BN + c * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
    }
    if (ra < p.a.rows) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const long long rb = b0 + c * 32 + j;
        if (rb >= p.b.rows) break;
        float x = f[j];
        if (p.ksplit > 1) { p.ws[(long long)blockIdx.z * p.w_slice + ra * p.w_a + rb * p.w_b] = x; continue; }
        float* dst = p.C + ra * p.c_a + rb * p.c_b;
        if (p.accumulate) x += *dst;
        if (p.relu) x = fmaxf(x, 0.f);
        if (p.mask && !(p.mask[ra * p.m_a + rb * p.m_b] > 0.f)) x = 0.f;
        *dst = x;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(T3_TMEM) : "memory");
  }
}

template <int BN>
constexpr size_t t3_smem_bytes() { return (size_t)T3_STAGES * (2 * T3_BM * 128 + 2 * BN * 128) + 1024 + 256; }

}  // namespace srlx
