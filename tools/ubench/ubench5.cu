// Instruction-cache capacity / miss cost for ONE warp running a straight-line loop body of S KB, and cluster sizes.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdint.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int N>  // N instructions (16 B each) of straight-line dependent IMADs per loop iteration
__global__ void body_kernel(int* io, int iters, long long* cyc) {
  int x = io[threadIdx.x], y = io[32 + threadIdx.x];
  long long t0 = 0;
  for (int it = 0; it < iters; ++it) {
    if (it == 2) t0 = clock64();
#pragma unroll
    for (int k = 0; k < N; ++k) x = x * y + k;
  }
  const long long t1 = clock64();
  io[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = (t1 - t0) / (iters - 2);
}

__global__ void cluster_probe(int* out) {
  cg::cluster_group cl = cg::this_cluster();
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = (int)cl.num_blocks();
  cl.sync();
}

template <int N>
int run(int* io, long long* cyc) {
  long long h;
  body_kernel<N><<<1, 32>>>(io, 12, cyc);
  CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("loop body %6.1f KB (%5d instr): %8lld cycles/iter = %.2f cycles/instr\n", N * 16 / 1024.0, N, h, (double)h / N);
  return 0;
}

int main() {
  int* io; long long* cyc; CK(cudaMalloc(&io, 1024)); CK(cudaMemset(io, 1, 1024)); CK(cudaMalloc(&cyc, 8));
  run<256>(io, cyc); run<512>(io, cyc); run<1024>(io, cyc); run<1536>(io, cyc); run<2048>(io, cyc); run<2560>(io, cyc);
  run<3072>(io, cyc); run<4096>(io, cyc); run<6144>(io, cyc); run<8192>(io, cyc);
  int* out; CK(cudaMalloc(&out, 4));
  for (int cs : {9, 10, 12, 15, 16}) {
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(cs); cfg.blockDim = dim3(64);
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaFuncSetAttribute(cluster_probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaError_t e = cudaLaunchKernelEx(&cfg, cluster_probe, out);
    cudaError_t e2 = cudaDeviceSynchronize();
    int h = -1; cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
    printf("cluster size %2d: launch=%s sync=%s reported=%d\n", cs, cudaGetErrorString(e), cudaGetErrorString(e2), h);
    cudaGetLastError();
  }
  return 0;
}
