// rankbased.cu -- rank-based prioritized replay on device, behind the reference's IPriorityMemory seam:
//   RankBasedMemory.sample  srl/rl/memories/priority_memories/rankbased_memory.py:42-60
//     sorted_indices = np.argsort(-priorities[:N]); probs = (1 / ranks) ** alpha, normalised; np.random.choice(sorted_indices,
//     size=B, p=probs, replace=False); weights = (N * prob) ** (-beta) / max
// The reference sorts all N priorities on the host at every sample (an O(N log N) numpy argsort per trainer step).  Here:
//   * rank_sort: LSD radix sort of the N float32 keys (8 bits per pass, 4 passes) carrying the item index -- three kernels per
//     pass (per-tile digit histogram, exclusive scan over (digit, tile), stable scatter), tiles of 4096 keys per CTA so that the
//     grid covers all 148 SMs from ~600 k items on.  Keys: -priority in ascending order, NaN last (numpy puts NaN at the end; a
//     priority of None is stored as NaN, rankbased_memory.py:40), -0.0 == +0.0; ties keep ascending item order (numpy's
//     introsort leaves tie order unspecified).
//   * rank_cdf: cumulative (1/k)^alpha over the ranks in fp64 (3-kernel scan).  It depends on (N, alpha) only, so it is rebuilt
//     only when N changes -- a full replay memory never rebuilds it.
//   * rank_draw: np.random.choice(..., replace=False) as numpy's legacy RandomState does it (mtrand choice): draw size - n_uniq
//     uniforms, zero the probabilities already found, cdf = cumsum(p) / cdf[-1], searchsorted(x, side='right'), keep the first
//     occurrence of every new index in draw order, repeat until `size` distinct ranks are found.  The zeroed-and-renormalised cdf
//     is evaluated as (cum(k) - sum of found weights <= k) / (total - sum of found weights) instead of being re-accumulated.
// CPU twin: oracle/rankbased.py; pinned by tests/golden/rankbased.npz (the reference class run with a seeded np.random).
#include "common.cuh"
#include "philox.cuh"

namespace srlx {

constexpr int kRsThreads = 256, kRsItems = 16, kRsTile = kRsThreads * kRsItems;  // 4096 keys per CTA

__device__ __forceinline__ uint32_t rank_key(float p) {
  // ascending order of -p; NaN last
  if (p != p) return 0xFFFFFFFFu;
  float q = -p;
  if (q == 0.0f) q = 0.0f;  // -0.0 -> +0.0
  const uint32_t b = __float_as_uint(q);
  const uint32_t k = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return k == 0xFFFFFFFFu ? 0xFFFFFFFEu : k;
}

__global__ void __launch_bounds__(kRsThreads) rs_init_kernel(const float* __restrict__ pri, uint32_t n, uint32_t* __restrict__ keys,
                                                             uint32_t* __restrict__ idx) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    keys[i] = rank_key(pri[i]);
    idx[i] = i;
  }
}

// per-tile histogram of the digit at `shift`: hist[digit * n_tiles + tile]
__global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t n_tiles,
                                                             uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[256];
  const uint32_t tile = blockIdx.x;
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = tile * kRsTile;
#pragma unroll
  for (int j = 0; j < kRsItems; ++j) {
    const uint32_t i = base + j * kRsThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * n_tiles + tile] = h[threadIdx.x];
}

// exclusive scan over the 256 * n_tiles counters (digit-major), one CTA
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t* __restrict__ hist, uint32_t total) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < total; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < total ? hist[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, x, s);
      if (lane >= s) x += t;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = wsum[lane];
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, w, s);
        if (lane >= s) w += t;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    const uint32_t incl = x + (warp > 0 ? wsum[warp - 1] : 0u) + carry;
    if (i < total) hist[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = incl;
    __syncthreads();
  }
}

// stable scatter: warp w owns the contiguous keys [w * 512, (w + 1) * 512) of the tile, 16 rounds of 32 consecutive keys; within a
// round lanes with equal digits are ranked by lane (match_any), across rounds by a per-warp digit counter, across warps by a prefix
__global__ void __launch_bounds__(kRsThreads) rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ idx_in,
                                                                uint32_t n, int shift, uint32_t n_tiles, const uint32_t* __restrict__ offs,
                                                                uint32_t* __restrict__ keys_out, uint32_t* __restrict__ idx_out) {
  __shared__ uint32_t cnt[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tile = blockIdx.x, base = tile * kRsTile + warp * (kRsTile / 8);
  for (int d = threadIdx.x; d < 8 * 256; d += kRsThreads) (&cnt[0][0])[d] = 0;
  __syncthreads();
  uint32_t k[kRsItems], v[kRsItems], lr[kRsItems];
#pragma unroll
  for (int j = 0; j < kRsItems; ++j) {
    const uint32_t i = base + j * 32 + lane;
    const bool ok = i < n;
    k[j] = ok ? keys_in[i] : 0xFFFFFFFFu;
    v[j] = ok ? idx_in[i] : 0u;
    const uint32_t d = (k[j] >> shift) & 255u;
    // lanes past the end vote in their own class (digit 256 + lane never matches a real digit)
    const unsigned peers = __match_any_sync(0xffffffffu, ok ? d : 256u + lane);
    const uint32_t before = __popc(peers & ((1u << lane) - 1u));
    const uint32_t c0 = ok ? cnt[warp][d] : 0u;
    __syncwarp();
    if (ok && before == 0) cnt[warp][d] = c0 + __popc(peers);
    __syncwarp();
    lr[j] = c0 + before;
  }
  __syncthreads();
  // per digit: exclusive prefix over the 8 warps, plus the global offset of (digit, tile)
  {
    const int d = threadIdx.x;
    uint32_t run = offs[(size_t)d * n_tiles + tile];
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const uint32_t c = cnt[w][d];
      cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kRsItems; ++j) {
    const uint32_t i = base + j * 32 + lane;
    if (i < n) {
      const uint32_t d = (k[j] >> shift) & 255u;
      const uint32_t o = cnt[warp][d] + lr[j];
      keys_out[o] = k[j];
      idx_out[o] = v[j];
    }
  }
}

// ---- cumulative rank weights: cum[k] = sum_{j <= k} (1 / (j + 1)) ^ alpha, fp64 -------------------------------------------------
constexpr int kCdfThreads = 256, kCdfItems = 8, kCdfTile = kCdfThreads * kCdfItems;

__device__ __forceinline__ double rank_weight(uint32_t k, double alpha) { return pow(1.0 / (double)(k + 1), alpha); }

__global__ void __launch_bounds__(kCdfThreads) cdf_tile_kernel(uint32_t n, double alpha, double* __restrict__ cum, double* __restrict__ tile_sum) {
  __shared__ double ws[kCdfThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = blockIdx.x * kCdfTile + threadIdx.x * kCdfItems;
  double loc[kCdfItems], run = 0.0;
#pragma unroll
  for (int j = 0; j < kCdfItems; ++j) {
    const uint32_t k = base + j;
    run += k < n ? rank_weight(k, alpha) : 0.0;
    loc[j] = run;
  }
  double x = run;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, x, s);
    if (lane >= s) x += t;
  }
  if (lane == 31) ws[warp] = x;
  __syncthreads();
  double pre = x - run;
  for (int w = 0; w < warp; ++w) pre += ws[w];
#pragma unroll
  for (int j = 0; j < kCdfItems; ++j)
    if (base + j < n) cum[base + j] = pre + loc[j];
  if (threadIdx.x == kCdfThreads - 1) tile_sum[blockIdx.x] = pre + run;
}
__global__ void cdf_tile_scan_kernel(double* __restrict__ tile_sum, uint32_t n_tiles) {  // one thread: n_tiles <= a few thousand
  double run = 0.0;
  for (uint32_t t = 0; t < n_tiles; ++t) {
    const double v = tile_sum[t];
    tile_sum[t] = run;
    run += v;
  }
}
__global__ void __launch_bounds__(kCdfThreads) cdf_add_kernel(uint32_t n, double* __restrict__ cum, const double* __restrict__ tile_pre) {
  const double pre = tile_pre[blockIdx.x];
  const uint32_t base = blockIdx.x * kCdfTile;
  for (uint32_t i = threadIdx.x; i < kCdfTile; i += kCdfThreads)
    if (base + i < n) cum[base + i] += pre;
}

// ---- np.random.choice(sorted_indices, size=B, p=probs, replace=False) + IS weights ------------------------------------------------
// one warp.  u01: the uniform stream in the order RandomState.random_sample would hand it out (tests inject numpy's own stream);
// NULL -> Philox(seed, STREAM_SAMPLE, (draw number, draw_id lo, draw_id hi)).
__global__ void __launch_bounds__(32) rank_draw_kernel(const double* __restrict__ cum, const uint32_t* __restrict__ sorted_idx, uint32_t n,
                                                       double alpha, double beta, uint32_t batch, const double* __restrict__ u01, uint32_t n_u,
                                                       uint64_t seed, uint64_t draw_id, int64_t* __restrict__ out_idx,
                                                       double* __restrict__ out_w, uint32_t* __restrict__ out_ranks, uint32_t* __restrict__ out_used) {
  __shared__ uint32_t found[SRLX_MAX_BATCH];
  __shared__ double found_w[SRLX_MAX_BATCH];
  __shared__ uint32_t cand[SRLX_MAX_BATCH];
  const int lane = threadIdx.x;
  const double total = cum[n - 1];
  uint32_t n_uniq = 0, used = 0;
  for (int iter = 0; iter < 4096 && n_uniq < batch; ++iter) {
    const uint32_t need = batch - n_uniq;
    double removed_all = 0.0;
    for (uint32_t f = 0; f < n_uniq; ++f) removed_all += found_w[f];
    const double rem_total = total - removed_all;
    // searchsorted(x, side='right') on the zeroed, renormalised cdf: smallest k with adj(k) / rem_total > x
    for (uint32_t j = lane; j < need; j += 32) {
      const uint32_t q = used + j;
      double x;
      if (u01) x = q < n_u ? u01[q] : 0.5;
      else {
        const uint4 w = philox(seed, STREAM_SAMPLE, q, (uint32_t)draw_id, (uint32_t)(draw_id >> 32));
        x = u01_f64(w.x, w.y);
      }
      uint32_t lo = 0, hi = n;  // answer in [lo, hi]
      while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        double a = cum[mid];
        for (uint32_t f = 0; f < n_uniq; ++f)
          if (found[f] <= mid) a -= found_w[f];
        if (a / rem_total > x) hi = mid; else lo = mid + 1;
      }
      cand[j] = lo < n ? lo : n - 1;
    }
    __syncwarp();
    used += need;
    // first occurrences, in draw order (np.unique(return_index) + sort of the indices)
    if (lane == 0) {
      const uint32_t n0 = n_uniq;
      for (uint32_t j = 0; j < need; ++j) {
        bool dup = false;
        for (uint32_t f = n0; f < n_uniq; ++f) dup |= (found[f] == cand[j]);
        if (!dup) {
          found[n_uniq] = cand[j];
          found_w[n_uniq] = rank_weight(cand[j], alpha);
          ++n_uniq;
        }
      }
      cand[0] = n_uniq;
    }
    __syncwarp();
    n_uniq = cand[0];
    __syncwarp();
  }
  // weights = (N * prob) ** (-beta) / max  (rankbased_memory.py:57-59), prob = weight of the rank / total
  double wmax = 0.0;
  for (uint32_t j = lane; j < batch; j += 32) {
    const double w = pow((double)n * (found_w[j] / total), -beta);
    out_w[j] = w;
    wmax = fmax(wmax, w);
  }
  for (int s = 16; s > 0; s >>= 1) wmax = fmax(wmax, __shfl_xor_sync(0xffffffffu, wmax, s));
  __syncwarp();
  for (uint32_t j = lane; j < batch; j += 32) {
    out_w[j] = out_w[j] / wmax;
    out_idx[j] = (int64_t)sorted_idx[found[j]];
    if (out_ranks) out_ranks[j] = found[j];
  }
  if (lane == 0 && out_used) *out_used = used;
}

__global__ void rank_set_kernel(float* __restrict__ pri, const int64_t* __restrict__ idx, const float* __restrict__ val, uint32_t n) {
  // RankBasedMemory.update (rankbased_memory.py:62-64): sequential assignment -- for a repeated index the LAST value wins
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (uint32_t i = 0; i < n; ++i) pri[idx[i]] = val[i];
}

struct RankScratch {
  uint32_t *keys[2], *idx[2], *hist;
  double *cum, *tile_sum;
  uint32_t n_tiles_max, n_cdf_tiles_max;
};
static size_t align_up(size_t x) { return (x + 255) / 256 * 256; }
static size_t rank_layout(uint32_t cap, unsigned char* base, RankScratch* s) {
  const uint32_t nt = (cap + kRsTile - 1) / kRsTile, nc = (cap + kCdfTile - 1) / kCdfTile;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes); return r; };
  const size_t k0 = take((size_t)cap * 4), k1 = take((size_t)cap * 4), i0 = take((size_t)cap * 4), i1 = take((size_t)cap * 4);
  const size_t h = take((size_t)256 * nt * 4), c = take((size_t)cap * 8), ts = take((size_t)nc * 8);
  if (s) {
    s->keys[0] = (uint32_t*)(base + k0); s->keys[1] = (uint32_t*)(base + k1);
    s->idx[0] = (uint32_t*)(base + i0); s->idx[1] = (uint32_t*)(base + i1);
    s->hist = (uint32_t*)(base + h); s->cum = (double*)(base + c); s->tile_sum = (double*)(base + ts);
    s->n_tiles_max = nt; s->n_cdf_tiles_max = nc;
  }
  return o;
}

}  // namespace srlx

extern "C" size_t srlx_rank_scratch_bytes(uint64_t capacity) {
  if (capacity == 0 || capacity > (1ull << 27)) return 0;
  return srlx::rank_layout((uint32_t)capacity, nullptr, nullptr);
}

// RankBasedMemory.sample for the first n priorities: sort, (re)build the rank cdf if asked, draw `batch` distinct items, IS weights.
extern "C" int srlx_rank_sample(const float* priorities_dev, uint64_t capacity, uint32_t n, double alpha, double beta, uint32_t batch,
                                const double* u01_dev, uint32_t n_u, uint64_t seed, uint64_t draw_id, int rebuild_cdf, void* scratch_dev,
                                int64_t* out_idx_dev, double* out_weights_dev, uint32_t* out_ranks_dev, uint32_t* out_used_dev,
                                uintptr_t cuda_stream) {
  using namespace srlx;
  SRLX_REQUIRE(priorities_dev && scratch_dev && out_idx_dev && out_weights_dev, "srlx_rank_sample: NULL buffer");
  SRLX_REQUIRE(capacity >= 1 && capacity <= (1ull << 27) && n >= 1 && n <= capacity, "srlx_rank_sample: n = %u outside [1, capacity]", n);
  SRLX_REQUIRE(batch >= 1 && batch <= SRLX_MAX_BATCH && batch <= n, "srlx_rank_sample: batch %u outside [1, min(%d, n)]", batch, SRLX_MAX_BATCH);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  RankScratch s;
  rank_layout((uint32_t)capacity, (unsigned char*)scratch_dev, &s);
  const uint32_t n_tiles = (n + kRsTile - 1) / kRsTile;
  rs_init_kernel<<<n_tiles < 1184 ? n_tiles : 1184, kRsThreads, 0, st>>>(priorities_dev, n, s.keys[0], s.idx[0]);
  count_launch();
  int cur = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 8 * pass;
    rs_hist_kernel<<<n_tiles, kRsThreads, 0, st>>>(s.keys[cur], n, shift, n_tiles, s.hist);
    rs_scan_kernel<<<1, 1024, 0, st>>>(s.hist, 256u * n_tiles);
    rs_scatter_kernel<<<n_tiles, kRsThreads, 0, st>>>(s.keys[cur], s.idx[cur], n, shift, n_tiles, s.hist, s.keys[cur ^ 1], s.idx[cur ^ 1]);
    count_launch(3);
    cur ^= 1;
  }
  SRLX_CHECK_CUDA(cudaGetLastError());
  if (rebuild_cdf) {
    const uint32_t nc = (n + kCdfTile - 1) / kCdfTile;
    cdf_tile_kernel<<<nc, kCdfThreads, 0, st>>>(n, alpha, s.cum, s.tile_sum);
    cdf_tile_scan_kernel<<<1, 1, 0, st>>>(s.tile_sum, nc);
    cdf_add_kernel<<<nc, kCdfThreads, 0, st>>>(n, s.cum, s.tile_sum);
    count_launch(3);
    SRLX_CHECK_CUDA(cudaGetLastError());
  }
  rank_draw_kernel<<<1, 32, 0, st>>>(s.cum, s.idx[cur], n, alpha, beta, batch, u01_dev, n_u, seed, draw_id, out_idx_dev, out_weights_dev,
                                     out_ranks_dev, out_used_dev);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// RankBasedMemory.update: priorities[indices[i]] = values[i] in order
extern "C" int srlx_rank_update(float* priorities_dev, const int64_t* idx_dev, const float* values_dev, uint32_t n, uintptr_t cuda_stream) {
  using namespace srlx;
  SRLX_REQUIRE(priorities_dev && idx_dev && values_dev, "srlx_rank_update: NULL buffer");
  if (n == 0) return 0;
  rank_set_kernel<<<1, 32, 0, (cudaStream_t)cuda_stream>>>(priorities_dev, idx_dev, values_dev, n);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// sorted order only (tests / tools): out_sorted_idx[k] = item with the k-th largest priority
extern "C" int srlx_rank_argsort(const float* priorities_dev, uint64_t capacity, uint32_t n, void* scratch_dev, uint32_t* out_sorted_idx_dev,
                                 uintptr_t cuda_stream) {
  using namespace srlx;
  SRLX_REQUIRE(priorities_dev && scratch_dev && out_sorted_idx_dev, "srlx_rank_argsort: NULL buffer");
  SRLX_REQUIRE(capacity >= 1 && capacity <= (1ull << 27) && n >= 1 && n <= capacity, "srlx_rank_argsort: n = %u outside [1, capacity]", n);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  RankScratch s;
  rank_layout((uint32_t)capacity, (unsigned char*)scratch_dev, &s);
  const uint32_t n_tiles = (n + kRsTile - 1) / kRsTile;
  rs_init_kernel<<<n_tiles < 1184 ? n_tiles : 1184, kRsThreads, 0, st>>>(priorities_dev, n, s.keys[0], s.idx[0]);
  count_launch();
  int cur = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 8 * pass;
    rs_hist_kernel<<<n_tiles, kRsThreads, 0, st>>>(s.keys[cur], n, shift, n_tiles, s.hist);
    rs_scan_kernel<<<1, 1024, 0, st>>>(s.hist, 256u * n_tiles);
    rs_scatter_kernel<<<n_tiles, kRsThreads, 0, st>>>(s.keys[cur], s.idx[cur], n, shift, n_tiles, s.hist, s.keys[cur ^ 1], s.idx[cur ^ 1]);
    count_launch(3);
    cur ^= 1;
  }
  SRLX_CHECK_CUDA(cudaMemcpyAsync(out_sorted_idx_dev, s.idx[cur], (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
