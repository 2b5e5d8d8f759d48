// philox.cuh -- Philox4x32-10 counter RNG on device; the CPU twin is oracle/philox.py (same words, same conversions).
// The reference draws from global Mersenne-Twister streams (srl/algorithms/dqn/dqn.py:200-202, srl/envs/grid.py:174,203,
// srl/rl/memories/priority_memories/proportional_memory.py:147); a counter RNG makes every draw a pure function of
// (seed, stream, who, when) so the oracle can replay it.
#pragma once
#include "common.cuh"

namespace srlx {

enum : uint32_t {
  STREAM_ENV_RESET = 1,
  STREAM_ENV_STEP = 2,
  STREAM_POLICY = 3,
  STREAM_NOISE = 4,
  STREAM_SAMPLE = 5,
  STREAM_PAD_ACTION = 6,
  STREAM_UNIFORM_SAMPLE = 7,
};
enum : uint32_t { NOISE_KIND_ROLLOUT = 0, NOISE_KIND_TRAIN = 1, NOISE_KIND_PRED = 3 };

__host__ __device__ inline uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c.x;
    uint64_t p1 = (uint64_t)M1 * c.z;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += W0;
    k1 += W1;
  }
  return c;
}

__host__ __device__ inline uint4 philox(uint64_t seed, uint32_t stream, uint32_t a, uint32_t b, uint32_t c) {
  return philox4x32_10(make_uint4(a, b, c, stream), (uint32_t)seed, (uint32_t)(seed >> 32));
}

// 24-bit uniform in [0,1), exact in fp32
__host__ __device__ inline float u01_f32(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }
// 53-bit uniform in [0,1), exact in fp64
__host__ __device__ inline double u01_f64(uint32_t hi, uint32_t lo) {
  return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}
// uniform integer in [0,n) by multiply-shift
__host__ __device__ inline uint32_t u_below(uint32_t w, uint32_t n) { return (uint32_t)(((uint64_t)w * n) >> 32); }

// Four N(0,1) draws for flat parameter block `blk` (parameters 4*blk .. 4*blk+3) of NoisyLinear forward call
// (kind, call_id): Box-Muller over Philox(seed, STREAM_NOISE, (blk, call_lo, call_hi | kind<<28)).
__device__ inline float4 noise4(uint64_t seed, uint32_t kind, uint64_t call_id, uint32_t blk) {
  uint4 w = philox(seed, STREAM_NOISE, blk, (uint32_t)call_id, ((uint32_t)(call_id >> 32) & 0x0FFFFFFFu) | (kind << 28));
  float4 z;
  {
    float u1 = ((float)(w.x >> 8) + 1.0f) * (1.0f / 16777216.0f);
    float u2 = (float)(w.y >> 8) * (1.0f / 16777216.0f);
    float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    z.x = r * c;
    z.y = r * s;
  }
  {
    float u1 = ((float)(w.z >> 8) + 1.0f) * (1.0f / 16777216.0f);
    float u2 = (float)(w.w >> 8) * (1.0f / 16777216.0f);
    float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    z.z = r * c;
    z.w = r * s;
  }
  return z;
}

}  // namespace srlx
