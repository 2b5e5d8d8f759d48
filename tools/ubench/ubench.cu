// Micro-benchmarks that pin the B200 numbers the learner design depends on (run: tools/ubench/run.sh under gpurun).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void fma_kernel(float* out, int iters, long long* cyc) {
  float2 a0 = make_float2(threadIdx.x * 1e-3f, 1.f), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0, a6 = a0, a7 = a0;
  const float2 w = make_float2(1.0001f, 0.9999f), b = make_float2(1e-4f, 2e-4f);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {  // scalar FFMA, 16 independent chains
      a0.x = fmaf(a0.x, w.x, b.x); a0.y = fmaf(a0.y, w.y, b.y); a1.x = fmaf(a1.x, w.x, b.x); a1.y = fmaf(a1.y, w.y, b.y);
      a2.x = fmaf(a2.x, w.x, b.x); a2.y = fmaf(a2.y, w.y, b.y); a3.x = fmaf(a3.x, w.x, b.x); a3.y = fmaf(a3.y, w.y, b.y);
      a4.x = fmaf(a4.x, w.x, b.x); a4.y = fmaf(a4.y, w.y, b.y); a5.x = fmaf(a5.x, w.x, b.x); a5.y = fmaf(a5.y, w.y, b.y);
      a6.x = fmaf(a6.x, w.x, b.x); a6.y = fmaf(a6.y, w.y, b.y); a7.x = fmaf(a7.x, w.x, b.x); a7.y = fmaf(a7.y, w.y, b.y);
    } else {  // FFMA2, 8 independent chains (same 16 FMAs per iteration)
      a0 = __ffma2_rn(a0, w, b); a1 = __ffma2_rn(a1, w, b); a2 = __ffma2_rn(a2, w, b); a3 = __ffma2_rn(a3, w, b);
      a4 = __ffma2_rn(a4, w, b); a5 = __ffma2_rn(a5, w, b); a6 = __ffma2_rn(a6, w, b); a7 = __ffma2_rn(a7, w, b);
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0.x + a0.y + a1.x + a1.y + a2.x + a2.y + a3.x + a3.y + a4.x + a4.y + a5.x + a5.y + a6.x + a6.y + a7.x + a7.y;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// dependent pointer chase through L2 (ld.global.cg) : latency per load
__global__ void chase_kernel(const int* next, int n, int* out, long long* cyc) {
  int i = 0;
  const long long t0 = clock64();
  for (int k = 0; k < n; ++k) i = __ldcg(next + i);
  const long long t1 = clock64();
  *out = i;
  *cyc = t1 - t0;
}

// 16 independent loads in flight then a dependent use: batch latency
__global__ void batch_kernel(const double* p, int stride, double* out, long long* cyc) {
  double v[16];
  const long long t0 = clock64();
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = __ldcg(p + (size_t)(threadIdx.x + 32 * k) * stride);
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += v[k];
  const long long t1 = clock64();
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

// named barrier round: 16 warps bar.sync repeatedly
__global__ void bar_kernel(int iters, long long* cyc) {
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) asm volatile("bar.sync 1, 512;" ::: "memory");
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

// double pow latency
__global__ void pow_kernel(double* io, int iters, long long* cyc) {
  double x = io[threadIdx.x];
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) x = pow(x + 1e-4, 0.6) + 0.5;
  const long long t1 = clock64();
  io[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* out; long long* cyc; long long h;
  CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&cyc, 8));
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) fma_kernel<0><<<1, warps * 32>>>(out, iters, cyc); else fma_kernel<1><<<1, warps * 32>>>(out, iters, cyc);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
      const double fma_per_clk = (double)warps * 32 * 16 * iters / (double)h;
      printf("fma mode=%s warps=%2d: %lld cycles, %.1f FMA/clk/SM\n", mode ? "FFMA2" : "FFMA ", warps, h, fma_per_clk);
    }
  // pointer chase
  {
    const int N = 1 << 22;  // 16 MB of ints: L2 resident after first pass
    int* hn = (int*)malloc(N * 4);
    uint32_t s = 12345; for (int i = 0; i < N; ++i) { s = s * 1664525u + 1013904223u; hn[i] = (int)(s % N); }
    int* dn; int* dout; CK(cudaMalloc(&dn, N * 4)); CK(cudaMalloc(&dout, 4));
    CK(cudaMemcpy(dn, hn, N * 4, cudaMemcpyHostToDevice));
    for (int rep = 0; rep < 3; ++rep) { chase_kernel<<<1, 1>>>(dn, 2000, dout, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); printf("chase (16MB, rep %d): %.1f cycles/load\n", rep, h / 2000.0); }
  }
  {
    double* dp; double* dout; CK(cudaMalloc(&dp, (size_t)64 << 20)); CK(cudaMalloc(&dout, 4096)); CK(cudaMemset(dp, 0, (size_t)64 << 20));
    for (int rep = 0; rep < 3; ++rep) { batch_kernel<<<1, 32>>>(dp, 977, dout, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); printf("batch of 16 scattered loads/lane (rep %d): %lld cycles\n", rep, h); }
  }
  { bar_kernel<<<1, 512>>>(1000, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); printf("bar.sync 512 threads: %.1f cycles/barrier\n", h / 1000.0); }
  { double* io; CK(cudaMalloc(&io, 256)); CK(cudaMemset(io, 0, 256)); pow_kernel<<<1, 32>>>(io, 100, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); printf("double pow: %.1f cycles/call\n", h / 100.0); }
  return 0;
}
