// DSMEM ping-pong: st.async (+ mbarrier complete_tx) round trip between two CTAs of a cluster
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdint.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void st_async(uint32_t ra, float4 v, uint32_t rb) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(ra), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(rb) : "memory");
}
__device__ __forceinline__ void expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void wait(uint64_t* b, uint32_t par) {
  uint32_t ok = 0;
  do { asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory"); } while (!ok);
}
__global__ void __cluster_dims__(2, 1, 1) pingpong(int iters, long long* cyc) {
  __shared__ __align__(16) float4 buf[4];
  __shared__ __align__(8) uint64_t bar;
  cg::cluster_group cl = cg::this_cluster();
  const int rank = cl.block_rank();
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); expect(&bar, 16); }
  cl.sync();
  if (threadIdx.x == 0) {
    const uint32_t peer_buf = mapa(smem_u32(&buf[0]), rank ^ 1), peer_bar = mapa(smem_u32(&bar), rank ^ 1);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (rank == 0) { st_async(peer_buf, make_float4(i, 0, 0, 0), peer_bar); wait(&bar, i & 1); if (i + 1 < iters) expect(&bar, 16); }
      else { wait(&bar, i & 1); if (i + 1 < iters) expect(&bar, 16); st_async(peer_buf, make_float4(i, 1, 0, 0), peer_bar); }
    }
    long long t1 = clock64();
    if (rank == 0) *cyc = (t1 - t0) / iters;
  }
  cl.sync();
}
// self-send: st.async to own CTA then wait
__global__ void __cluster_dims__(2, 1, 1) selfsend(int iters, long long* cyc) {
  __shared__ __align__(16) float4 buf[4];
  __shared__ __align__(8) uint64_t bar;
  cg::cluster_group cl = cg::this_cluster();
  const int rank = cl.block_rank();
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); expect(&bar, 16); }
  cl.sync();
  if (threadIdx.x == 0) {
    const uint32_t my_buf = mapa(smem_u32(&buf[0]), rank), my_bar = mapa(smem_u32(&bar), rank);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { st_async(my_buf, make_float4(i, 0, 0, 0), my_bar); wait(&bar, i & 1); if (i + 1 < iters) expect(&bar, 16); }
    long long t1 = clock64();
    if (rank == 0) *cyc = (t1 - t0) / iters;
  }
  cl.sync();
}
int main() {
  long long* cyc; long long h; CK(cudaMalloc(&cyc, 8));
  pingpong<<<2, 32>>>(1000, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("st.async ping-pong round trip: %lld cycles (one way ~%lld)\n", h, h / 2);
  selfsend<<<2, 32>>>(1000, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("st.async to own CTA + wait: %lld cycles\n", h);
  return 0;
}
