// FP64 throughput per SM: independent DFMA chains, 1..16 warps; and pow_chain-like latency with 1 vs 4 warps
#include <cuda_runtime.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__global__ void dfma_kernel(double* io, int iters, long long* cyc) {
  double a0 = io[threadIdx.x], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double w = 1.0000001, b = 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, w, b); a1 = fma(a1, w, b); a2 = fma(a2, w, b); a3 = fma(a3, w, b);
    a4 = fma(a4, w, b); a5 = fma(a5, w, b); a6 = fma(a6, w, b); a7 = fma(a7, w, b);
  }
  const long long t1 = clock64();
  io[threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void powchain_kernel(double* io, double alpha, int iters, long long* cyc) {
  double x = io[threadIdx.x];
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) x = exp(alpha * log(x * 0.37 + 1e-4)) + 0.5;
  const long long t1 = clock64();
  io[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* io; long long* cyc; long long h; CK(cudaMalloc(&io, 8192)); CK(cudaMemset(io, 0, 8192)); CK(cudaMalloc(&cyc, 8));
  for (int warps = 1; warps <= 16; warps *= 2) {
    dfma_kernel<<<1, warps * 32>>>(io, 1000, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("DFMA %2d warps: %lld cycles for %d warp-instr -> %.2f cycles per warp-DFMA per SM (%.1f DFMA lanes/clk/SM)\n", warps, h, warps * 8000, (double)h / (warps * 8000), warps * 8000.0 * 32 / h);
  }
  for (int warps = 1; warps <= 4; warps *= 2) {
    powchain_kernel<<<1, warps * 32>>>(io, 0.6, 100, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("exp(a*log x) with %d warps concurrently: %.1f cycles per call\n", warps, h / 100.0);
  }
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("persistingL2CacheMaxSize %d MB, accessPolicyMaxWindowSize %d MB, l2CacheSize %d MB\n", p.persistingL2CacheMaxSize >> 20, p.accessPolicyMaxWindowSize >> 20, p.l2CacheSize >> 20);
  return 0;
}
