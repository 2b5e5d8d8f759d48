// Instruction fetch: (a) first (cold) pass over a straight-line body, (b) alternating between two large bodies so that
// each pass finds its lines evicted from the SM's instruction cache, (c) the same with 4 warps running different bodies.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int N, int SALT>
__device__ __noinline__ int body(int x, int y) {
#pragma unroll
  for (int k = 0; k < N; ++k) x = x * y + (k ^ SALT);
  return x;
}

// one warp; per iteration: run body A (NA instr), then body B (NB instr, the "evictor"); time body A only
template <int NA, int NB>
__global__ void alt_kernel(int* io, int iters, long long* cyc_first, long long* cyc_steady) {
  int x = io[threadIdx.x], y = io[32 + threadIdx.x];
  long long tot = 0, first = 0;
  for (int it = 0; it < iters; ++it) {
    const long long t0 = clock64();
    x = body<NA, 1>(x, y);
    const long long t1 = clock64();
    if (it == 0) first = t1 - t0; else tot += t1 - t0;
    if (NB > 0) x = body<NB, 2>(x, y);
  }
  io[threadIdx.x] = x;
  if (threadIdx.x == 0) { *cyc_first = first; *cyc_steady = tot / (iters - 1); }
}

// 4 warps (one per SMSP), each its own body of N instr, timed by warp 0
template <int N>
__global__ void four_kernel(int* io, int iters, long long* cyc) {
  int x = io[threadIdx.x], y = io[32 + (threadIdx.x & 31)];
  const int w = threadIdx.x >> 5;
  long long tot = 0;
  for (int it = 0; it < iters; ++it) {
    __syncthreads();
    const long long t0 = clock64();
    if (w == 0) x = body<N, 11>(x, y);
    else if (w == 1) x = body<N, 12>(x, y);
    else if (w == 2) x = body<N, 13>(x, y);
    else x = body<N, 14>(x, y);
    const long long t1 = clock64();
    if (it > 0) tot += t1 - t0;
  }
  io[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = tot / (iters - 1);
}

int main() {
  int* io; long long *c1, *c2, h1, h2; CK(cudaMalloc(&io, 1024)); CK(cudaMemset(io, 1, 1024)); CK(cudaMalloc(&c1, 8)); CK(cudaMalloc(&c2, 8));
#define RUN(NA, NB) alt_kernel<NA, NB><<<1, 32>>>(io, 20, c1, c2); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h1, c1, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&h2, c2, 8, cudaMemcpyDeviceToHost)); \
  printf("body %5.1f KB, evictor %5.1f KB: first pass %.2f cyc/instr, steady %.2f cyc/instr\n", NA * 16 / 1024.0, NB * 16 / 1024.0, (double)h1 / NA, (double)h2 / NA);
  RUN(128, 0) RUN(128, 1024) RUN(128, 2048) RUN(128, 4096) RUN(128, 8192)
  RUN(512, 0) RUN(512, 2048) RUN(512, 4096) RUN(512, 8192)
#define RUN4(N) four_kernel<N><<<1, 128>>>(io, 20, c1); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h1, c1, 8, cudaMemcpyDeviceToHost)); \
  printf("4 warps x different bodies of %5.1f KB: %.2f cyc/instr\n", N * 16 / 1024.0, (double)h1 / N);
  RUN4(256) RUN4(512) RUN4(1024) RUN4(2048)
  return 0;
}
