// Sampler-like access pattern: 4 warps x 8 samples, 31 lanes fetch both children of every node of a 5-level subtree.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ double ld_cg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ double ld_ca(const double* p) { return *p; }

template <int MODE>
__global__ void tree_round(double* tree, int n_nodes, int level0, int iters, int do_store, long long* cyc, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k_l = 32 - __clz(lane + 1), q_l = lane + 1 - (1 << (k_l - 1));
  const unsigned c_l = (1u << k_l) - 1u + 2u * (unsigned)q_l;
  uint32_t s = 1234567u + warp * 7919u;
  double acc = 0;
  long long tot = 0;
  for (int it = 0; it < iters; ++it) {
    unsigned idx[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) { s = s * 1664525u + 1013904223u; idx[g] = (1u << level0) - 1u + ((s >> 8) & ((1u << level0) - 1u)); }
    if (do_store) {  // emulate the update's write-through stores to random deep nodes, then a barrier
      s = s * 1664525u + 1013904223u;
      __stcg(tree + ((s >> 4) % (unsigned)n_nodes), 1.0);
      __syncthreads();
    }
    const long long t0 = clock64();
    double v0[8], v1[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      unsigned node = (idx[g] << k_l) + c_l;
      node = node < (unsigned)n_nodes - 2u ? node : (unsigned)n_nodes - 2u;
      if (MODE == 0) { v0[g] = ld_cg(tree + node); v1[g] = ld_cg(tree + node + 1); }
      else { v0[g] = ld_ca(tree + node); v1[g] = ld_ca(tree + node + 1); }
    }
    double a = 0;
#pragma unroll
    for (int g = 0; g < 8; ++g) a += v0[g] + v1[g];
    const long long t1 = clock64();
    acc += a;
    tot += t1 - t0;
    __syncthreads();
  }
  if (threadIdx.x == 0) *cyc = tot / iters;
  out[threadIdx.x] = acc;
}

// ring-gather-like: one warp, 16 independent loads per lane from random rows of a 96 MB buffer (DRAM), then use
__global__ void gather_round(const float4* ring, size_t n_rows, int iters, int prefetch, long long* cyc, float* out) {
  uint32_t s = 99991u * (threadIdx.x + 1);
  float acc = 0;
  long long tot = 0;
  for (int it = 0; it < iters; ++it) {
    size_t rows[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { s = s * 1664525u + 1013904223u; rows[k] = (size_t)(s >> 3) % n_rows; }
    if (prefetch) {
#pragma unroll
      for (int k = 0; k < 16; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(ring + rows[k]));
      for (int w = 0; w < 60; ++w) { s = s * 1664525u + 1013904223u; acc += (float)(s & 1); __nanosleep(20); }
    }
    const long long t0 = clock64();
    float4 v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = __ldcg(ring + rows[k]);
    float a = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) a += v[k].x + v[k].w;
    const long long t1 = clock64();
    acc += a;
    tot += t1 - t0;
  }
  if (threadIdx.x == 0) *cyc = tot / iters;
  out[threadIdx.x] = acc;
}

int main() {
  const int n_nodes = (1 << 22) - 1;
  double* tree; long long* cyc; double* out; long long h;
  CK(cudaMalloc(&tree, (size_t)n_nodes * 8)); CK(cudaMemset(tree, 0, (size_t)n_nodes * 8));
  CK(cudaMalloc(&cyc, 8)); CK(cudaMalloc(&out, 4096));
  char* flush; CK(cudaMalloc(&flush, 256 << 20));
  for (int mode = 0; mode < 2; ++mode)
    for (int level0 = 11; level0 <= 16; level0 += 5)
      for (int st = 0; st < 2; ++st)
        for (int cold = 0; cold < 2; ++cold) {
          if (cold) CK(cudaMemset(flush, 1, 256 << 20));
          else { if (mode == 0) tree_round<0><<<1, 128>>>(tree, n_nodes, level0, 2000, st, cyc, out); else tree_round<1><<<1, 128>>>(tree, n_nodes, level0, 2000, st, cyc, out); }
          if (mode == 0) tree_round<0><<<1, 128>>>(tree, n_nodes, level0, cold ? 20 : 200, st, cyc, out); else tree_round<1><<<1, 128>>>(tree, n_nodes, level0, cold ? 20 : 200, st, cyc, out);
          CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
          printf("tree round %s levels %d..%d store=%d %s: %lld cycles\n", mode ? "ld.ca" : "ld.cg", level0 + 1, level0 + 5, st, cold ? "cold(L2 flushed)" : "warm", h);
        }
  float4* ring; float* fo; const size_t n_rows = (size_t)6 << 20;  // 96 MB
  CK(cudaMalloc(&ring, n_rows * 16)); CK(cudaMemset(ring, 0, n_rows * 16)); CK(cudaMalloc(&fo, 4096));
  for (int pf = 0; pf < 2; ++pf) {
    CK(cudaMemset(flush, 1, 256 << 20));
    gather_round<<<1, 32>>>(ring, n_rows, 50, pf, cyc, fo);
    CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("gather 16 rows/lane from 96MB (DRAM) prefetch=%d: %lld cycles\n", pf, h);
  }
  return 0;
}
