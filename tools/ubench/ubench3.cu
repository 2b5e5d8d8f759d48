// latencies of the scalar building blocks on the SumTree critical chain
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int OP>
__global__ void lat_kernel(double* io, double alpha, int iters, long long* cyc) {
  double x = io[threadIdx.x], y = io[32 + threadIdx.x];
  float xf = (float)x;
  int xi = (int)threadIdx.x + 3;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) x = pow(x * 0.37 + 1e-4, alpha) + 0.5;              // runtime exponent
    if (OP == 1) x = exp(alpha * log(x * 0.37 + 1e-4)) + 0.5;
    if (OP == 2) x = x + y;                                          // DADD chain
    if (OP == 3) x = fma(x, y, y);                                   // DFMA chain
    if (OP == 4) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;       // double shuffle + DADD
    if (OP == 5) xf = __shfl_xor_sync(0xffffffffu, xf, 1) + 1.0f;    // float shuffle + FADD
    if (OP == 6) x = (x <= y) ? x + 1.0 : x - y;                     // DSETP + select + DADD
    if (OP == 7) x = x / (y + 2.0);                                  // double division
    if (OP == 8) xi = xi / (int)(y + 7.0) + 1000003;                 // int division by a runtime value
    if (OP == 9) xf = fmaf(xf, 1.0001f, 0.5f);                       // FFMA chain
    if (OP == 10) xi = xi * 3 + 1;                                   // IMAD chain
    if (OP == 11) x = exp2((double)(float)(alpha) * log2(x * 0.37 + 1e-4)) + 0.5;
    if (OP == 12) xf = __powf(xf * 0.37f + 1e-4f, (float)alpha) + 0.5f;
  }
  const long long t1 = clock64();
  io[threadIdx.x] = x + xf + xi;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double* io; long long* cyc; long long h;
  CK(cudaMalloc(&io, 1024)); CK(cudaMalloc(&cyc, 8));
  double hv[64]; for (int i = 0; i < 64; ++i) hv[i] = 0.3 + 0.01 * i;
  const char* names[] = {"pow(x, runtime alpha)", "exp(alpha*log x)", "DADD", "DFMA", "shfl(double)+DADD", "shfl(float)+FADD", "DSETP+sel+DADD", "double div", "int div (runtime)", "FFMA", "IMAD", "exp2(a*log2 x) double", "__powf"};
  for (int op = 0; op < 13; ++op) {
    CK(cudaMemcpy(io, hv, 512, cudaMemcpyHostToDevice));
    const int iters = 200;
    switch (op) {
#define L(n) case n: lat_kernel<n><<<1, 32>>>(io, 0.6, iters, cyc); break;
      L(0) L(1) L(2) L(3) L(4) L(5) L(6) L(7) L(8) L(9) L(10) L(11) L(12)
    }
    CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("%-26s %.1f cycles/iter\n", names[op], (double)h / iters);
  }
  return 0;
}
