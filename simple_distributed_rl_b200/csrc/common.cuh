// common.cuh -- shared host/device helpers for libsrlx.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/srlx.h"

namespace srlx {

// thread-local error string + launch counter (defined in cabi.cu)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define SRLX_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      srlx::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -1;                                                                           \
    }                                                                                      \
  } while (0)

#define SRLX_REQUIRE(cond, ...)       \
  do {                                \
    if (!(cond)) {                    \
      srlx::set_error(__VA_ARGS__);   \
      return -2;                      \
    }                                 \
  } while (0)

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Row stride (floats) for a [rows][K] shared-memory matrix that is read one float4 per lane with consecutive lanes on
// consecutive rows: multiple of 4 and (stride/4) odd -> the 8 lanes of a quarter-warp hit 8 distinct 16-byte bank groups.
__host__ __device__ inline int padded_ld(int K) {
  int ld = round_up(K, 4);
  if (((ld >> 2) & 1) == 0) ld += 4;
  return ld;
}

}  // namespace srlx
