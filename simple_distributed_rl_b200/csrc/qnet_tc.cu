// qnet_tc.cu -- the LARGE-BATCH Q-network inference mode on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// The trainer path (learner*.cu) is fp32 FMA on purpose: the reference's parity bar (1e-4 rel on targets / loss) rules out bf16
// operands, and at batch 32 the layers are latency-bound, not flop-bound.  Where the Q-net really is a GEMM -- RLParameter.pred_q /
// pred_target_q (srl/algorithms/dqn/model_torch.py:58-70) over hundreds of thousands of states: evaluation sweeps, value maps,
// distillation targets -- this file provides an explicit NON-parity mode: bf16 operands, fp32 accumulation in tensor memory.
//
//   dense_bf16_tc_kernel   Y[M][N] = act(X[M][K] . W[N][K]^T + b), one CTA per 128 x 128 output tile, 192 threads:
//       warp 0 (one lane)  TMA producer: 128 x 64 bf16 boxes of X and W (128-byte swizzle) into a 4-stage shared-memory ring,
//                          completion on the stage's "full" mbarrier (complete_tx::bytes)
//       warp 1             allocates 128 TMEM columns; one lane issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = 128, K = 16,
//                          four per stage, shared-memory descriptors advanced 32 bytes inside the swizzle atom) and hands stages back
//                          with tcgen05.commit on the "empty" mbarrier; the last commit signals the epilogue
//       warps 2..5         epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) -> + bias -> ReLU -> bf16 (hidden layers) or
//                          fp32 (last layer) -> 16-byte global stores; warp w reads the TMEM lane quarter (w % 4)
//   qnet_pack_bf16_kernel  effective weights (mu + sigma * eps on NoisyLinear layers) of every layer as padded bf16 matrices; the
//                          dueling output layer becomes one dense [1 + A][2H] matrix (zeros where a row does not read)
//   dueling_combine_kernel Q = V + A - mean(A) / max(A) / nothing (srl/rl/torch_/blocks/dueling_network.py:51-58)
// Descriptor encodings follow cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor / InstrDescriptor) of the CUTLASS tree shipped in
// the image; every mbarrier wait is bounded (a wrong descriptor must trap, not hang the GPU).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "philox.cuh"

namespace srlx {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64, TC_STAGES = 4, TC_THREADS = 192;
constexpr uint32_t TC_STAGE_A = TC_BM * TC_BK * 2, TC_STAGE_B = TC_BN * TC_BK * 2, TC_STAGE = TC_STAGE_A + TC_STAGE_B;
constexpr uint32_t TC_SMEM = TC_STAGES * TC_STAGE + 1024 /* alignment slack */ + 256 /* barriers */;

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(b)), "r"(bytes) : "memory");
}
// bounded wait: ~1 s, then trap (the launch fails with an error instead of hanging the device)
__device__ __forceinline__ void tc_mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t a = tc_smem_u32(b);
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
__device__ __forceinline__ void tc_tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   tc_smem_u32(smem_dst)),
               "l"(map), "r"(tc_smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// K-major operand tile in the canonical 128-byte-swizzle layout (rows of 128 bytes, 8-row groups 1024 bytes apart):
// start address >> 4, leading byte offset 1 (unused for swizzled K-major), stride byte offset 1024 >> 4, version 1, layout SWIZZLE_128B
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, N = 128, M = 128
__device__ __forceinline__ uint32_t tc_instr_desc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}

template <bool OUT_F32>
__global__ void __launch_bounds__(TC_THREADS, 1)
dense_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const float* __restrict__ bias,
                     void* __restrict__ y, const int M, const int N, const int K, const int ldy, const int relu) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* tmem_full = empty + TC_STAGES;
  uint32_t* tmem_base_p = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TC_BN, m0 = blockIdx.y * TC_BM;
  const int n_k = (K + TC_BK - 1) / TC_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { tc_mbar_init(&full[s], 1); tc_mbar_init(&empty[s], 1); }
    tc_mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 1) {  // TMEM: 128 columns of 128 lanes x 32 bit = the fp32 accumulator tile
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_base_p)), "n"(TC_BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_p;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer
      for (int kb = 0; kb < n_k; ++kb) {
        const int s = kb % TC_STAGES;
        if (kb >= TC_STAGES) tc_mbar_wait(&empty[s], ((kb / TC_STAGES) - 1) & 1);
        tc_mbar_expect_tx(&full[s], TC_STAGE);
        tc_tma_load_2d(smem + s * TC_STAGE, &map_x, kb * TC_BK, m0, &full[s]);
        tc_tma_load_2d(smem + s * TC_STAGE + TC_STAGE_A, &map_w, kb * TC_BK, n0, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer
      const uint32_t idesc = tc_instr_desc();
      for (int kb = 0; kb < n_k; ++kb) {
        const int s = kb % TC_STAGES;
        tc_mbar_wait(&full[s], (kb / TC_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = tc_smem_u32(smem + s * TC_STAGE), b0 = a0 + TC_STAGE_A;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k)
          tc_mma(tmem_base, tc_smem_desc(a0 + k * 32), tc_smem_desc(b0 + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
        tc_commit(&empty[s]);  // (implies tcgen05.fence::before_thread_sync) the stage is free once these MMAs have read it
      }
      tc_commit(tmem_full);
    }
  } else {
    // ---- epilogue warps 2..5: TMEM lane quarter (warp % 4), one output row per lane
    tc_mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < TC_BN / 32; ++c) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
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
      const int col0 = n0 + c * 32;
      if (row < M) {
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]);
          if (bias && col0 + j < N) x += __ldg(bias + col0 + j);
          if (relu) x = fmaxf(x, 0.f);
          f[j] = x;
        }
        if (OUT_F32) {
          float* dst = reinterpret_cast<float*>(y) + (size_t)row * ldy + col0;
          if (col0 + 32 <= N && (ldy & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < N) dst[j] = f[j];
          }
        } else {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(y) + (size_t)row * ldy + col0;
          if (col0 + 32 <= N && (ldy & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              __nv_bfloat162 p0 = __floats2bfloat162_rn(f[j], f[j + 1]), p1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
              __nv_bfloat162 p2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]), p3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
              uint4 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
              pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
              *reinterpret_cast<uint4*>(dst + j) = pk;
            }
          } else {
            // ragged edge: the padding columns [N, ldy) are written too (zeros: zero-filled weight rows, no bias), the next layer
            // multiplies them with zero weights and must not meet uninitialised bits
            for (int j = 0; j < 32; ++j)
              if (col0 + j < ldy) dst[j] = __float2bfloat16_rn(f[j]);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_BN) : "memory");
  }
}

// ---- weights of one forward call as bf16 matrices: layer l -> [out_pad_l][k_pad_l], row-major (K contiguous) ------------------------
struct TcLayer {
  int n_out, k_in, k_pad, w_elem_off;  // logical rows / columns, padded columns (multiple of 8), element offset in the bf16 pack
};
struct TcPlan {
  int n_layers;
  TcLayer l[SRLX_MAX_LAYERS];
  size_t w_elems;
};
__host__ __device__ inline TcPlan make_tc_plan(const srlx_net& net) {
  TcPlan p;
  p.n_layers = net.n_layers;
  size_t off = 0;
  for (int i = 0; i < net.n_layers; ++i) {
    const bool duel_out = net.dueling != SRLX_DUEL_NONE && i == net.n_layers - 1;
    p.l[i].n_out = net.out_dim[i];
    p.l[i].k_in = duel_out ? net.out_dim[i - 1] : net.k_dim[i];  // the dueling output layer reads the whole [value ; advantage] hidden vector
    p.l[i].k_pad = round_up(p.l[i].k_in, 8);
    p.l[i].w_elem_off = (int)off;
    off += (size_t)p.l[i].n_out * p.l[i].k_pad;
    off = (off + 7) / 8 * 8;
  }
  p.w_elems = off;
  return p;
}
__global__ void qnet_pack_bf16_kernel(const __grid_constant__ srlx_net net, const float* __restrict__ mu, const float* __restrict__ sigma,
                                      const int use_noise, const uint64_t seed, const uint64_t call_id, __nv_bfloat16* __restrict__ wpack,
                                      float* __restrict__ bias_pack) {
  const TcPlan tp = make_tc_plan(net);
  for (int l = 0; l < net.n_layers; ++l) {
    const TcLayer L = tp.l[l];
    const bool duel_out = net.dueling != SRLX_DUEL_NONE && l == net.n_layers - 1;
    const int H = net.k_dim[l];
    const int total = L.n_out * L.k_pad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int o = i / L.k_pad, k = i - o * L.k_pad;
      float v = 0.f;
      int src = -1;
      if (!duel_out) {
        if (k < L.k_in) src = net.w_off[l] + o * L.k_in + k;
      } else {  // row 0 (V) reads hidden[0:H], rows 1..A read hidden[H:2H]
        const int lo = o == 0 ? 0 : H;
        if (k >= lo && k < lo + H) src = net.w_off[l] + o * H + (k - lo);
      }
      if (src >= 0) {
        v = mu[src];
        if (use_noise && net.layer_noisy[l]) {
          const float4 z = noise4(seed, NOISE_KIND_PRED, call_id, (uint32_t)(src >> 2));
          const float zz = (src & 3) == 0 ? z.x : ((src & 3) == 1 ? z.y : ((src & 3) == 2 ? z.z : z.w));
          v = fmaf(sigma[src], zz, v);
        }
      }
      wpack[L.w_elem_off + i] = __float2bfloat16_rn(v);
    }
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < L.n_out; o += gridDim.x * blockDim.x) {
      const int src = net.b_off[l] + o;
      float v = mu[src];
      if (use_noise && net.layer_noisy[l]) {
        const float4 z = noise4(seed, NOISE_KIND_PRED, call_id, (uint32_t)(src >> 2));
        const float zz = (src & 3) == 0 ? z.x : ((src & 3) == 1 ? z.y : ((src & 3) == 2 ? z.z : z.w));
        v = fmaf(sigma[src], zz, v);
      }
      bias_pack[l * 4096 + o] = v;
    }
  }
}
__global__ void obs_to_bf16_kernel(const float* __restrict__ obs, int n, int D, int k_pad, __nv_bfloat16* __restrict__ out) {
  const size_t total = (size_t)n * k_pad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / k_pad;
    const int k = (int)(i - r * k_pad);
    out[i] = __float2bfloat16_rn(k < D ? obs[r * D + k] : 0.f);
  }
}
__global__ void dueling_combine_kernel(const float* __restrict__ raw, int n, int A, int ld, int dueling, float* __restrict__ q) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const float* x = raw + (size_t)r * ld;
    if (dueling == SRLX_DUEL_NONE) {
      for (int a = 0; a < A; ++a) q[(size_t)r * A + a] = x[a];
    } else {
      float red = 0.f;
      if (dueling == SRLX_DUEL_AVERAGE) {
        for (int a = 0; a < A; ++a) red += x[1 + a];
        red /= (float)A;
      } else if (dueling == SRLX_DUEL_MAX) {
        red = x[1];
        for (int a = 1; a < A; ++a) red = fmaxf(red, x[1 + a]);
      }
      for (int a = 0; a < A; ++a) q[(size_t)r * A + a] = x[0] + x[1 + a] - red;
    }
  }
}

// ---- host: tensor maps through the driver entry point (no link-time dependency on libcuda) ---------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// rows x K bf16 matrix, K contiguous (row stride ld elements): boxes of 64 (K) x 128 (rows), 128-byte swizzle, zero fill out of bounds
static int make_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t K, uint64_t ld) {
  EncodeTiledFn enc = get_encode();
  SRLX_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[2] = {K, rows};
  const cuuint64_t strides[1] = {ld * 2};
  const cuuint32_t box[2] = {TC_BK, TC_BM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SRLX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): rows %llu, K %llu, ld %llu", (int)r, (unsigned long long)rows,
               (unsigned long long)K, (unsigned long long)ld);
  return 0;
}

static int dense_tc(const void* x, int ldx, const void* w, int ldw, const float* bias, void* y, int ldy, int out_f32, int M, int N, int K,
                    int relu, cudaStream_t st) {
  SRLX_REQUIRE(M >= 1 && N >= 1 && K >= 8 && (K % 8) == 0 && (ldx % 8) == 0 && (ldw % 8) == 0, "dense_tc: K, ldx, ldw must be multiples of 8 (16-byte rows)");
  SRLX_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y & 15) == 0, "dense_tc: buffers must be 16-byte aligned");
  CUtensorMap mx, mw;
  if (int rc = make_map(&mx, x, (uint64_t)M, (uint64_t)K, (uint64_t)ldx)) return rc;
  if (int rc = make_map(&mw, w, (uint64_t)N, (uint64_t)K, (uint64_t)ldw)) return rc;
  const dim3 grid((unsigned)((N + TC_BN - 1) / TC_BN), (unsigned)((M + TC_BM - 1) / TC_BM));
  if (out_f32) {
    SRLX_CHECK_CUDA(cudaFuncSetAttribute(dense_bf16_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    dense_bf16_tc_kernel<true><<<grid, TC_THREADS, TC_SMEM, st>>>(mx, mw, bias, y, M, N, K, ldy, relu);
  } else {
    SRLX_CHECK_CUDA(cudaFuncSetAttribute(dense_bf16_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    dense_bf16_tc_kernel<false><<<grid, TC_THREADS, TC_SMEM, st>>>(mx, mw, bias, y, M, N, K, ldy, relu);
  }
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace srlx

// Y = act(X . W^T + b): X [M][K] bf16 (row stride ldx), W [N][K] bf16 (row stride ldw), bias fp32 [N] or NULL, Y [M][N] bf16 or fp32
// (row stride ldy).  K, ldx, ldw multiples of 8; out-of-range rows / columns / K are zero-filled by TMA and masked on store.
extern "C" int srlx_dense_bf16_tc(const void* x_dev, int ldx, const void* w_dev, int ldw, const float* bias_dev, void* y_dev, int ldy,
                                  int out_f32, int M, int N, int K, int relu, uintptr_t cuda_stream) {
  using namespace srlx;
  SRLX_REQUIRE(x_dev && w_dev && y_dev, "srlx_dense_bf16_tc: NULL buffer");
  return dense_tc(x_dev, ldx, w_dev, ldw, bias_dev, y_dev, ldy, out_f32, M, N, K, relu, (cudaStream_t)cuda_stream);
}

// workspace bytes of srlx_qnet_forward_tc for n states
extern "C" size_t srlx_qnet_tc_workspace_bytes(const srlx_engine* eng, uint32_t n) {
  using namespace srlx;
  if (!eng) return 0;
  const TcPlan tp = make_tc_plan(eng->net);
  int wmax = round_up(eng->obs_dim, 8);
  for (int l = 0; l < eng->net.n_layers; ++l) wmax = wmax > round_up(eng->net.out_dim[l], 8) ? wmax : round_up(eng->net.out_dim[l], 8);
  size_t o = 0;
  o += (tp.w_elems * 2 + 255) / 256 * 256;                    // bf16 weights
  o += (size_t)SRLX_MAX_LAYERS * 4096 * 4;                    // biases (fp32, 4096 per layer)
  o += 2 * (((size_t)n * wmax * 2 + 255) / 256 * 256);        // two activation buffers (bf16)
  o += ((size_t)n * 8 * 4 + 255) / 256 * 256 * 4;             // raw outputs (fp32, up to 32 columns)
  return o;
}

// RLParameter.pred_q / pred_target_q for n states on the tensor cores (bf16 operands, fp32 accumulation): q_out [n][A] fp32.
extern "C" int srlx_qnet_forward_tc(const srlx_engine* eng, int use_target, const float* obs_dev, uint32_t n, uint64_t noise_call_id,
                                    float* q_out_dev, void* workspace_dev, size_t workspace_bytes, uintptr_t cuda_stream) {
  using namespace srlx;
  SRLX_REQUIRE(eng && obs_dev && q_out_dev && workspace_dev, "srlx_qnet_forward_tc: bad arguments");
  SRLX_REQUIRE(use_target ? eng->target != nullptr : eng->params != nullptr, "srlx_qnet_forward_tc: parameter buffer is NULL");
  SRLX_REQUIRE(workspace_bytes >= srlx_qnet_tc_workspace_bytes(eng, n), "srlx_qnet_forward_tc: workspace too small");
  if (n == 0) return 0;
  const srlx_net& net = eng->net;
  for (int l = 0; l < net.n_layers; ++l) SRLX_REQUIRE(net.out_dim[l] <= 4096, "layer wider than 4096");
  SRLX_REQUIRE(net.out_dim[net.n_layers - 1] <= 32, "more than 32 outputs");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const TcPlan tp = make_tc_plan(net);
  int wmax = round_up(eng->obs_dim, 8);
  for (int l = 0; l < net.n_layers; ++l) wmax = wmax > round_up(net.out_dim[l], 8) ? wmax : round_up(net.out_dim[l], 8);
  unsigned char* ws = (unsigned char*)workspace_dev;
  __nv_bfloat16* wpack = (__nv_bfloat16*)ws;
  size_t o = (tp.w_elems * 2 + 255) / 256 * 256;
  float* bias = (float*)(ws + o);
  o += (size_t)SRLX_MAX_LAYERS * 4096 * 4;
  const size_t act_bytes = ((size_t)n * wmax * 2 + 255) / 256 * 256;
  __nv_bfloat16* act[2] = {(__nv_bfloat16*)(ws + o), (__nv_bfloat16*)(ws + o + act_bytes)};
  o += 2 * act_bytes;
  float* raw = (float*)(ws + o);
  const float* mu = use_target ? eng->target : eng->params;
  const float* sg = use_target ? eng->target_sigma : eng->params_sigma;
  const int use_noise = net.noisy && sg != nullptr;
  qnet_pack_bf16_kernel<<<296, 256, 0, st>>>(net, mu, sg, use_noise, eng->seed, noise_call_id, wpack, bias);
  const int k0 = round_up(eng->obs_dim, 8);
  obs_to_bf16_kernel<<<1184, 256, 0, st>>>(obs_dev, (int)n, eng->obs_dim, k0, act[0]);
  count_launch(2);
  SRLX_CHECK_CUDA(cudaGetLastError());
  int cur = 0, ld_in = k0;
  const int raw_ld = 32;
  for (int l = 0; l < net.n_layers; ++l) {
    const bool last = l == net.n_layers - 1;
    const TcLayer L = tp.l[l];
    const int ld_out = last ? raw_ld : round_up(L.n_out, 8);
    if (int rc = dense_tc(act[cur], ld_in, wpack + L.w_elem_off, L.k_pad, bias + l * 4096, last ? (void*)raw : (void*)act[cur ^ 1], ld_out,
                          last ? 1 : 0, (int)n, L.n_out, L.k_pad, last ? 0 : 1, st))
      return rc;
    cur ^= 1;
    ld_in = ld_out;
  }
  dueling_combine_kernel<<<(n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184, 256, 0, st>>>(raw, (int)n, eng->n_actions, raw_ld, net.dueling, q_out_dev);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
