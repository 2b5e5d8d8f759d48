// cabi.cu -- library-level entry points of libsrlx.so (include/srlx.h): error string, RNG taps, the standalone
// SumTree / ProportionalMemory calls (the narrow IPriorityMemory seam) and the pred_q inference seam.
#include <stdarg.h>

#include <atomic>

#include <string.h>
#include "net.cuh"
#include "tree.cuh"

namespace srlx {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n); }

// ---- RNG taps ------------------------------------------------------------------------------------------------------
__global__ void philox_words_kernel(uint64_t seed, uint32_t stream, uint32_t a0, uint32_t b, uint32_t c, uint32_t* out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 w = philox(seed, stream, a0 + (uint32_t)i, b, c);
  out[4 * i + 0] = w.x;
  out[4 * i + 1] = w.y;
  out[4 * i + 2] = w.z;
  out[4 * i + 3] = w.w;
}

__global__ void noise_fill_kernel(uint64_t seed, uint32_t kind, uint64_t call_id, float* out, size_t n_params) {
  const size_t blk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (blk * 4 >= n_params) return;
  const float4 z = noise4(seed, kind, call_id, (uint32_t)blk);
  const float zz[4] = {z.x, z.y, z.z, z.w};
  for (int j = 0; j < 4; ++j)
    if (blk * 4 + j < n_params) out[blk * 4 + j] = zz[j];
}

// ---- standalone tree kernels (one thread block each; batch <= 1024) ------------------------------------------------
constexpr int kTreeThreads = 1024;  // sampler: one warp walks one sample, 32 samples in flight
constexpr int kTreeMaxBatch = 1024;
constexpr int kTreeAddChunk = 1024;  // items staged per pass of a bulk add

__global__ void tree_clear_kernel(double* tree, uint64_t n_nodes, srlx_state* meta) {
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = i0; i < n_nodes; i += stride) tree[i] = 0.0;
  if (i0 == 0) {
    meta->mem_size = 0;
    meta->vec_steps = 0;  // ring `write`
    meta->max_priority = 1.0;
  }
}

__global__ void __launch_bounds__(kTreeThreads)
tree_add_kernel(double* tree, uint64_t capacity, srlx_state* meta, const double* priorities, uint64_t n, double alpha,
                double epsilon, int restore_skip) {
  extern __shared__ __align__(16) unsigned char tree_smem[];
  TreeHashScratch* hs = reinterpret_cast<TreeHashScratch*>(tree_smem);
  __shared__ int64_t s_idx[kTreeAddChunk];
  __shared__ double s_pri[kTreeAddChunk];
  const uint64_t write = meta->vec_steps;
  const double maxp = meta->max_priority;
  for (uint64_t base = 0; base < n; base += kTreeAddChunk) {
    const int m = (int)((n - base < (uint64_t)kTreeAddChunk) ? (n - base) : kTreeAddChunk);
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      s_idx[i] = (int64_t)((write + base + i) % capacity) + (int64_t)capacity - 1;
      double p = maxp;
      if (priorities) {
        p = priorities[base + i];
        if (!restore_skip) p = pow(fabs(p) + epsilon, alpha);
      }
      s_pri[i] = p;
    }
    __syncthreads();
    tree_update_batch(tree, s_idx, s_pri, m, hs);
  }
  if (threadIdx.x == 0) {
    meta->vec_steps = (write + n) % capacity;
    const uint64_t sz = meta->mem_size + n;
    meta->mem_size = sz > capacity ? capacity : sz;
  }
}

__global__ void __launch_bounds__(kTreeThreads)
tree_sample_kernel(const double* tree, uint64_t capacity, srlx_state* meta, uint32_t batch, uint64_t step,
                   double beta_initial, double beta_steps, int has_duplicate, uint64_t seed, const double* u01,
                   uint32_t max_tries, int64_t* out_idx, float* out_w, double* out_pri) {
  __shared__ int64_t s_idx[kTreeMaxBatch];
  __shared__ double s_pri[kTreeMaxBatch];
  __shared__ double s_tmp[kTreeMaxBatch];
  __shared__ float s_w[kTreeMaxBatch];
  __shared__ unsigned long long retries;
  if (threadIdx.x == 0) retries = 0;
  __syncthreads();
  const double total = __ldcg(tree);
  double beta = beta_initial + (1.0 - beta_initial) * (double)step / beta_steps;
  if (beta > 1.0) beta = 1.0;
  per_sample_block(tree, 2 * (int64_t)capacity - 1, total, (int)batch, seed, step, u01, (int)max_tries, has_duplicate, s_idx,
                   s_pri, s_tmp, &retries);
  per_weights_block(total, (double)meta->mem_size, beta, (int)batch, s_pri, s_tmp, s_w);
  for (int i = threadIdx.x; i < (int)batch; i += blockDim.x) {
    out_idx[i] = s_idx[i];
    out_w[i] = s_w[i];
    if (out_pri) out_pri[i] = s_pri[i];
  }
  if (threadIdx.x == 0) meta->sample_retries += retries;
}

__global__ void __launch_bounds__(kTreeThreads)
tree_update_kernel(double* tree, srlx_state* meta, const int64_t* idx, const float* priorities, uint32_t n, double alpha,
                   double epsilon) {
  extern __shared__ __align__(16) unsigned char tree_smem[];
  TreeHashScratch* hs = reinterpret_cast<TreeHashScratch*>(tree_smem);
  __shared__ int64_t s_idx[kTreeMaxBatch];
  __shared__ double s_pri[kTreeMaxBatch];
  for (int i = threadIdx.x; i < (int)n; i += blockDim.x) {
    s_idx[i] = idx[i];
    s_pri[i] = pow(fabs((double)priorities[i]) + epsilon, alpha);
  }
  __syncthreads();
  tree_update_batch(tree, s_idx, s_pri, (int)n, hs);
  if (threadIdx.x == 0) {
    double mp = meta->max_priority;
    for (int i = 0; i < (int)n; ++i) mp = (mp < s_pri[i]) ? s_pri[i] : mp;
    meta->max_priority = mp;
  }
}

// The IPriorityMemory seam in ONE launch per sample(): the add / update calls the host made since the last sample (an op list in
// program order, read straight from mapped pinned host memory) are applied with the reference's sequential association, then the
// batch is drawn and the results go straight back to mapped host memory, followed by a sequence word the host polls -- no stream
// synchronisation, no staging copies.  op idx >= 0: update of that tree index with raw priority val; -1: add with raw priority val;
// -2: add with priority None (takes max_priority AS OF that point of the sequence); -3: add with the stored priority val (restore).
constexpr int64_t kOpAdd = -1, kOpAddNone = -2, kOpAddRaw = -3;
__global__ void __launch_bounds__(kTreeThreads)
tree_seam_kernel(double* tree, uint64_t capacity, srlx_state* meta, const int64_t* __restrict__ ops_idx, const double* __restrict__ ops_val,
                 uint32_t n_ops, double alpha, double epsilon, uint32_t batch, uint64_t step, double beta_initial, double beta_steps,
                 int has_duplicate, uint64_t seed, const double* u01, uint32_t max_tries, int64_t* out_idx, float* out_w,
                 volatile unsigned long long* flag, unsigned long long seq) {
  extern __shared__ __align__(16) unsigned char tree_smem[];
  TreeHashScratch* hs = reinterpret_cast<TreeHashScratch*>(tree_smem);
  __shared__ int64_t s_idx[kTreeMaxBatch];
  __shared__ double s_pri[kTreeMaxBatch];
  __shared__ double s_tmp[kTreeMaxBatch];
  __shared__ float s_w[kTreeMaxBatch];
  __shared__ unsigned long long retries;
  __shared__ unsigned long long s_write, s_size;
  __shared__ double s_maxp;
  if (threadIdx.x == 0) {
    retries = 0;
    s_write = meta->vec_steps;
    s_size = meta->mem_size;
    s_maxp = meta->max_priority;
  }
  __syncthreads();
  for (uint32_t base = 0; base < n_ops; base += kTreeHashChunk) {
    const int m = (int)min((uint32_t)kTreeHashChunk, n_ops - base);
    if ((int)threadIdx.x < m) {
      const int64_t k = ops_idx[base + threadIdx.x];
      const double v = ops_val[base + threadIdx.x];
      s_idx[threadIdx.x] = k;
      s_pri[threadIdx.x] = (k >= 0 || k == kOpAdd) ? pow(fabs(v) + epsilon, alpha) : v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // cursor, size and max_priority advance in program order
      unsigned long long w = s_write, sz = s_size;
      double mp = s_maxp;
      for (int i = 0; i < m; ++i) {
        const int64_t k = s_idx[i];
        if (k >= 0) {
          mp = (mp < s_pri[i]) ? s_pri[i] : mp;
        } else {
          if (k == kOpAddNone) s_pri[i] = mp;
          s_idx[i] = (int64_t)w + (int64_t)capacity - 1;
          w = (w + 1) % capacity;
          sz = sz < capacity ? sz + 1 : capacity;
        }
      }
      s_write = w; s_size = sz; s_maxp = mp;
    }
    __syncthreads();
    if (m <= 4) {  // a lone add between two samples: one warp, lane = level above the leaf, one L2 round trip per item
      if (threadIdx.x < 32) {
        for (int i = 0; i < m; ++i) {
          const uint64_t ip1 = (uint64_t)s_idx[i] + 1;
          const int l = threadIdx.x;
          const bool on = (ip1 >> l) != 0;
          const uint64_t node = on ? (ip1 >> l) - 1 : 0;
          double v = on ? __ldcg(tree + node) : 0.0;
          const double change = s_pri[i] - __shfl_sync(0xffffffffu, v, 0);
          if (on) __stcg(tree + node, l == 0 ? s_pri[i] : v + change);
          __syncwarp();
        }
      }
      __syncthreads();
    } else {
      tree_update_batch(tree, s_idx, s_pri, m, hs);
    }
  }
  if (threadIdx.x == 0 && n_ops) {
    meta->vec_steps = s_write;
    meta->mem_size = s_size;
    meta->max_priority = s_maxp;
  }
  if (batch) {
    __threadfence_block();
    __syncthreads();
    const double total = __ldcg(tree);
    double beta = beta_initial + (1.0 - beta_initial) * (double)step / beta_steps;
    if (beta > 1.0) beta = 1.0;
    per_sample_block(tree, 2 * (int64_t)capacity - 1, total, (int)batch, seed, step, u01, (int)max_tries, has_duplicate, s_idx, s_pri,
                     s_tmp, &retries);
    per_weights_block(total, (double)s_size, beta, (int)batch, s_pri, s_tmp, s_w);
    for (int i = threadIdx.x; i < (int)batch; i += blockDim.x) {
      out_idx[i] = s_idx[i];
      out_w[i] = s_w[i];
    }
    if (threadIdx.x == 0) meta->sample_retries += retries;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0 && flag) {
    *flag = seq;
    __threadfence_system();
  }
}

__global__ void tree_retrieve_kernel(const double* tree, int64_t n_nodes, const double* vals, uint32_t n, int64_t* out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= (int)n) return;
  const int64_t idx = tree_retrieve_warp(tree, n_nodes, vals[warp]);
  if ((threadIdx.x & 31) == 0) out[warp] = idx;
}

// ---- pred_q / pred_target_q ------------------------------------------------------------------------------------------
struct FwdSmem { size_t weff, acts, q, total; };
__host__ __device__ inline FwdSmem fwd_smem(const srlx_engine& eng, const NetPlan& pl) {
  FwdSmem s;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) / 16 * 16; return o; };
  s.weff = take((size_t)pl.weff_floats * 4);
  s.acts = take((size_t)pl.act_floats * 4);
  s.q = take((size_t)kRowTile * eng.n_actions * 4);
  s.total = off;
  return s;
}

__global__ void __launch_bounds__(256)
qnet_forward_kernel(const __grid_constant__ srlx_engine eng, int use_target, const float* __restrict__ obs, uint32_t n,
                    uint64_t noise_call_id, float* __restrict__ q_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const srlx_net& net = eng.net;
  const NetPlan pl = make_plan(net);
  const FwdSmem so = fwd_smem(eng, pl);
  float* weff = reinterpret_cast<float*>(smem_raw + so.weff);
  float* acts = reinterpret_cast<float*>(smem_raw + so.acts);
  float* q = reinterpret_cast<float*>(smem_raw + so.q);
  const int D = eng.obs_dim, A = eng.n_actions;
  zero_floats(weff, pl.weff_floats);
  zero_floats(acts, pl.act_floats);
  __syncthreads();
  build_weff(net, pl, use_target ? eng.target : eng.params, use_target ? eng.target_sigma : eng.params_sigma,
             net.noisy != 0, eng.seed, NOISE_KIND_PRED, noise_call_id, weff);
  __syncthreads();
  for (uint32_t r0 = blockIdx.x * kRowTile; r0 < n; r0 += gridDim.x * kRowTile) {
    const int Rr = (int)min((uint32_t)kRowTile, n - r0);
    for (int w = threadIdx.x; w < Rr * D; w += blockDim.x) {
      const int r = w / D, d = w - r * D;
      acts[pl.x_s[0] + r * pl.ldx[0] + d] = obs[(size_t)(r0 + r) * D + d];
    }
    __syncthreads();
    net_forward_tile(net, pl, weff, acts, Rr, q, A);
    for (int w = threadIdx.x; w < Rr * A; w += blockDim.x) q_out[(size_t)r0 * A + w] = q[w];
    __syncthreads();
  }
}

}  // namespace srlx

// ======================================================================================================================
using namespace srlx;

extern "C" int srlx_version(void) { return SRLX_VERSION; }
extern "C" const char* srlx_last_error(void) { return g_err; }
extern "C" size_t srlx_sizeof_engine(void) { return sizeof(srlx_engine); }
extern "C" size_t srlx_sizeof_state(void) { return sizeof(srlx_state); }
extern "C" size_t srlx_sizeof_net(void) { return sizeof(srlx_net); }
extern "C" uint64_t srlx_launch_count(void) { return g_launches.load(); }

extern "C" int srlx_philox_words(uint64_t seed, uint32_t stream, uint32_t a0, uint32_t b, uint32_t c, uint32_t* out_dev,
                                 size_t n, uintptr_t cuda_stream) {
  SRLX_REQUIRE(out_dev != nullptr, "out_dev is NULL");
  if (n == 0) return 0;
  philox_words_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(seed, stream, a0, b, c, out_dev, n);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int srlx_noise_fill(uint64_t seed, uint32_t kind, uint64_t call_id, float* out_dev, size_t n_params,
                               uintptr_t cuda_stream) {
  SRLX_REQUIRE(out_dev != nullptr, "out_dev is NULL");
  if (n_params == 0) return 0;
  const size_t nblk = (n_params + 3) / 4;
  noise_fill_kernel<<<(unsigned)((nblk + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(seed, kind, call_id, out_dev, n_params);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int srlx_tree_clear(double* tree, uint64_t capacity, srlx_state* meta, uintptr_t cuda_stream) {
  SRLX_REQUIRE(tree && meta && capacity >= 1, "srlx_tree_clear: bad arguments");
  tree_clear_kernel<<<296, 256, 0, (cudaStream_t)cuda_stream>>>(tree, 2 * capacity - 1, meta);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int srlx_tree_add(double* tree, uint64_t capacity, srlx_state* meta, const double* priorities_dev, uint64_t n,
                             double alpha, double epsilon, int restore_skip, uintptr_t cuda_stream) {
  SRLX_REQUIRE(tree && meta && capacity >= 1, "srlx_tree_add: bad arguments");
  SRLX_REQUIRE(n <= capacity, "srlx_tree_add: n (%llu) > capacity (%llu)", (unsigned long long)n, (unsigned long long)capacity);
  if (n == 0) return 0;
  SRLX_REQUIRE(capacity < (1ull << 30), "srlx_tree_add: capacity %llu too large (node ids are 32-bit)", (unsigned long long)capacity);
  SRLX_CHECK_CUDA(cudaFuncSetAttribute(tree_add_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TreeHashScratch)));
  tree_add_kernel<<<1, kTreeThreads, sizeof(TreeHashScratch), (cudaStream_t)cuda_stream>>>(tree, capacity, meta, priorities_dev, n, alpha,
                                                                                          epsilon, restore_skip);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int srlx_tree_sample(const double* tree, uint64_t capacity, srlx_state* meta, uint32_t batch, uint64_t step,
                                double beta_initial, double beta_steps, int has_duplicate, uint64_t seed,
                                const double* u01_dev, uint32_t max_tries, int64_t* out_tree_idx, float* out_weights,
                                double* out_priority, uintptr_t cuda_stream) {
  SRLX_REQUIRE(tree && meta && out_tree_idx && out_weights, "srlx_tree_sample: bad arguments");
  SRLX_REQUIRE(batch >= 1 && batch <= (uint32_t)kTreeMaxBatch, "srlx_tree_sample: batch %u out of range [1,%d]", batch, kTreeMaxBatch);
  SRLX_REQUIRE(max_tries >= 1 && max_tries <= 9999, "srlx_tree_sample: max_tries %u out of range [1,9999]", max_tries);
  tree_sample_kernel<<<1, kTreeThreads, 0, (cudaStream_t)cuda_stream>>>(tree, capacity, meta, batch, step, beta_initial, beta_steps,
                                                                       has_duplicate, seed, u01_dev, max_tries, out_tree_idx,
                                                                       out_weights, out_priority);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int srlx_tree_update(double* tree, uint64_t capacity, srlx_state* meta, const int64_t* tree_idx_dev,
                                const float* priorities_dev, uint32_t n, double alpha, double epsilon, uintptr_t cuda_stream) {
  SRLX_REQUIRE(tree && meta && tree_idx_dev && priorities_dev, "srlx_tree_update: bad arguments");
  SRLX_REQUIRE(n <= (uint32_t)kTreeMaxBatch, "srlx_tree_update: n %u > %d", n, kTreeMaxBatch);
  SRLX_REQUIRE(capacity < (1ull << 30), "srlx_tree_update: capacity %llu too large (node ids are 32-bit)", (unsigned long long)capacity);
  if (n == 0) return 0;
  SRLX_CHECK_CUDA(cudaFuncSetAttribute(tree_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TreeHashScratch)));
  tree_update_kernel<<<1, kTreeThreads, sizeof(TreeHashScratch), (cudaStream_t)cuda_stream>>>(tree, meta, tree_idx_dev, priorities_dev, n,
                                                                                             alpha, epsilon);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int srlx_tree_seam(double* tree, uint64_t capacity, srlx_state* meta, const int64_t* ops_idx, const double* ops_val, uint32_t n_ops,
                              double alpha, double epsilon, uint32_t batch, uint64_t step, double beta_initial, double beta_steps,
                              int has_duplicate, uint64_t seed, const double* u01_dev, uint32_t max_tries, int64_t* out_tree_idx,
                              float* out_weights, unsigned long long* flag, unsigned long long seq, uintptr_t cuda_stream) {
  SRLX_REQUIRE(tree && meta && capacity >= 1, "srlx_tree_seam: bad arguments");
  SRLX_REQUIRE(n_ops == 0 || (ops_idx && ops_val), "srlx_tree_seam: op list is NULL");
  SRLX_REQUIRE(batch <= (uint32_t)kTreeMaxBatch, "srlx_tree_seam: batch %u > %d", batch, kTreeMaxBatch);
  SRLX_REQUIRE(batch == 0 || (out_tree_idx && out_weights), "srlx_tree_seam: output buffer is NULL");
  SRLX_REQUIRE(batch == 0 || (max_tries >= 1 && max_tries <= 9999), "srlx_tree_seam: max_tries %u out of range [1,9999]", max_tries);
  SRLX_REQUIRE(capacity < (1ull << 30), "srlx_tree_seam: capacity %llu too large (node ids are 32-bit)", (unsigned long long)capacity);
  static bool attr_set = false;
  if (!attr_set) {
    SRLX_CHECK_CUDA(cudaFuncSetAttribute(tree_seam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TreeHashScratch)));
    attr_set = true;
  }
  tree_seam_kernel<<<1, kTreeThreads, sizeof(TreeHashScratch), (cudaStream_t)cuda_stream>>>(
      tree, capacity, meta, ops_idx, ops_val, n_ops, alpha, epsilon, batch, step, beta_initial, beta_steps, has_duplicate, seed, u01_dev,
      max_tries, out_tree_idx, out_weights, flag, seq);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int srlx_tree_seam_desc(const srlx_seam* d, uint32_t n_ops, uint32_t batch, uint64_t step, uint64_t seed, const double* u01_dev,
                                   uint32_t max_tries, unsigned long long seq, uintptr_t cuda_stream) {
  SRLX_REQUIRE(d != nullptr, "srlx_tree_seam_desc: descriptor is NULL");
  return srlx_tree_seam(d->tree, d->capacity, d->meta, d->ops_idx, d->ops_val, n_ops, d->alpha, d->epsilon, batch, step, d->beta_initial,
                        d->beta_steps, d->has_duplicate, seed, u01_dev, max_tries, d->out_tree_idx, d->out_weights, d->flag, seq, cuda_stream);
}

extern "C" int srlx_host_alloc(size_t bytes, void** host_ptr, void** dev_ptr) {
  SRLX_REQUIRE(host_ptr && dev_ptr && bytes > 0, "srlx_host_alloc: bad arguments");
  SRLX_CHECK_CUDA(cudaHostAlloc(host_ptr, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
  SRLX_CHECK_CUDA(cudaHostGetDevicePointer(dev_ptr, *host_ptr, 0));
  memset(*host_ptr, 0, bytes);
  return 0;
}

extern "C" int srlx_host_free(void* host_ptr) {
  if (host_ptr) SRLX_CHECK_CUDA(cudaFreeHost(host_ptr));
  return 0;
}

extern "C" int srlx_tree_retrieve(const double* tree, uint64_t capacity, const double* vals_dev, uint32_t n,
                                  int64_t* out_tree_idx, uintptr_t cuda_stream) {
  SRLX_REQUIRE(tree && vals_dev && out_tree_idx, "srlx_tree_retrieve: bad arguments");
  if (n == 0) return 0;
  const unsigned threads = 256, warps_per_block = threads / 32;
  tree_retrieve_kernel<<<(n + warps_per_block - 1) / warps_per_block, threads, 0, (cudaStream_t)cuda_stream>>>(
      tree, 2 * (int64_t)capacity - 1, vals_dev, n, out_tree_idx);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int srlx_qnet_forward(const srlx_engine* eng, int use_target, const float* obs_dev, uint32_t n,
                                 uint64_t noise_call_id, float* q_out_dev, uintptr_t cuda_stream) {
  SRLX_REQUIRE(eng && obs_dev && q_out_dev, "srlx_qnet_forward: bad arguments");
  SRLX_REQUIRE(use_target ? eng->target != nullptr : eng->params != nullptr, "srlx_qnet_forward: parameter buffer is NULL");
  if (n == 0) return 0;
  const NetPlan pl = make_plan(eng->net);
  const FwdSmem so = fwd_smem(*eng, pl);
  int dev = 0, max_smem = 0, n_sm = 0;
  SRLX_CHECK_CUDA(cudaGetDevice(&dev));
  SRLX_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  SRLX_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  SRLX_REQUIRE((int)so.total + 1024 <= max_smem, "network too large: needs %zu bytes of shared memory", so.total);
  SRLX_CHECK_CUDA(cudaFuncSetAttribute(qnet_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)so.total));
  unsigned tiles = (n + kRowTile - 1) / kRowTile;
  unsigned grid = tiles < (unsigned)n_sm ? tiles : (unsigned)n_sm;
  qnet_forward_kernel<<<grid, 256, so.total, (cudaStream_t)cuda_stream>>>(*eng, use_target, obs_dev, n, noise_call_id, q_out_dev);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- exchange buffers of the data-parallel learner (learner_fast.cu): one per rank, written by the peers over NVLink ----------
extern "C" int srlx_dp_alloc(size_t bytes, void** ptr_out, unsigned char handle_out[64]) {
  SRLX_REQUIRE(ptr_out && handle_out && bytes > 0, "srlx_dp_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  SRLX_CHECK_CUDA(cudaMalloc(&p, bytes));
  SRLX_CHECK_CUDA(cudaMemset(p, 0, bytes));
  SRLX_CHECK_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  SRLX_CHECK_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle_out, &h, 64);
  *ptr_out = p;
  return 0;
}
extern "C" int srlx_dp_open(const unsigned char handle[64], void** ptr_out) {
  SRLX_REQUIRE(ptr_out && handle, "srlx_dp_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* p = nullptr;
  SRLX_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = p;
  return 0;
}
extern "C" int srlx_dp_close(void* ptr) {
  if (ptr) SRLX_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}
extern "C" int srlx_dp_free(void* ptr) {
  if (ptr) SRLX_CHECK_CUDA(cudaFree(ptr));
  return 0;
}
extern "C" int srlx_dp_enable_peer(int device_a, int device_b) {
  if (device_a == device_b) return 0;
  int cur = 0, ok = 0;
  SRLX_CHECK_CUDA(cudaGetDevice(&cur));
  for (int k = 0; k < 2; ++k) {
    const int a = k == 0 ? device_a : device_b, b = k == 0 ? device_b : device_a;
    SRLX_CHECK_CUDA(cudaDeviceCanAccessPeer(&ok, a, b));
    SRLX_REQUIRE(ok, "device %d cannot access device %d (no P2P path)", a, b);
    SRLX_CHECK_CUDA(cudaSetDevice(a));
    cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
      cudaSetDevice(cur);
      SRLX_CHECK_CUDA(e);
    }
    cudaGetLastError();
  }
  SRLX_CHECK_CUDA(cudaSetDevice(cur));
  return 0;
}
