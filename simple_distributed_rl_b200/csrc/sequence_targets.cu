// sequence_targets.cu -- the per-sequence target loop of the reference's R2D2 trainer (srl/algorithms/r2d2/r2d2.py:150-203):
// walking a stored sequence backwards,
//   gain_t   = r_t  (done)  |  r_t + discount * Q_target(s_{t+1})[argmax Q_sel(s_{t+1})]   (Q_sel = online Q for double DQN;
//              with enable_rescale the bootstrap goes through inverse_rescaling and the gain through rescaling)
//   target_t = gain_t + retrace * td_{t+1},   td_t = gain_t - Q_online(s_t)[a_t],
//   retrace *= discount * retrace_h * min(1, pi(a_t|s_t) / mu_t)      (greedy pi of the online Q, ties share the mass)
// and the sequence's priority input mean_t(td_t) (:204).  One thread per sequence (B <= a few hundred sequences of <= 128
// steps: latency-bound by construction, the value is exactness, not speed).  The reference mixes float32 Q values with python
// floats; under numpy >= 2 promotion that fixes a precision per value ("kind": float32, float64 or python float), which this
// kernel tracks so that every rounding happens where the reference's does -- including numpy's pairwise summation order in the
// mean -- and the results equal tests/golden/r2d2_targets.npz bit for bit.  No FMA contraction anywhere (__f*_rn / __d*_rn).
// One known difference: inverse_rescaling's `n ** 2` on a numpy float32 scalar goes through libm powf, which is not correctly
// rounded; the device forms the exact product, so with enable_rescale about one value in 2000 differs by a float32 ulp.
// CPU twin: oracle/r2d2_targets.py.  The LSTM Q-network, burn-in and sequence replay around this loop are not built.
#include "common.cuh"

namespace srlx {

constexpr int kSeqMaxT = 128;
enum { K32 = 0, K64 = 1, KPY = 2 };

__device__ __forceinline__ float sgn_f(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }
__device__ __forceinline__ double sgn_d(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : 0.0); }
// srl/rl/functions.py:10-17 on a float32 scalar (python-float constants are weak: 0.001, 4*0.001, 2*0.001 rounded to float32)
__device__ __forceinline__ float rescaling_f32(float x) {
  const float a = __fsub_rn(__fsqrt_rn(__fadd_rn(fabsf(x), 1.0f)), 1.0f);
  return __fadd_rn(__fmul_rn(sgn_f(x), a), __fmul_rn(0.001f, x));
}
__device__ __forceinline__ float inverse_rescaling_f32(float x) {
  const float inner = __fadd_rn(__fadd_rn(fabsf(x), 1.0f), 0.001f);
  float n = __fsub_rn(__fsqrt_rn(__fadd_rn(1.0f, __fmul_rn((float)(4.0 * 0.001), inner))), 1.0f);
  n = __fdiv_rn(n, (float)(2.0 * 0.001));
  return __fmul_rn(sgn_f(x), __fsub_rn(__fmul_rn(n, n), 1.0f));
}
__device__ __forceinline__ double rescaling_f64(double x) {
  const double a = __dsub_rn(__dsqrt_rn(__dadd_rn(fabs(x), 1.0)), 1.0);
  return __dadd_rn(__dmul_rn(sgn_d(x), a), __dmul_rn(0.001, x));
}

// np.add.reduce over a contiguous 1-D array of n <= 128 values (numpy's pairwise summation, one block): eight running sums
template <class T>
__device__ inline T np_pairwise_sum(const double* a, int n) {
  auto add = [](T x, T y) -> T {
    if (sizeof(T) == 4) return (T)__fadd_rn((float)x, (float)y);
    return (T)__dadd_rn((double)x, (double)y);
  };
  if (n < 8) {
    T res = (T)0;
    for (int i = 0; i < n; ++i) res = add(res, (T)a[i]);
    return res;
  }
  T r[8];
  for (int j = 0; j < 8; ++j) r[j] = (T)a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] = add(r[j], (T)a[i + j]);
  T res = add(add(add(r[0], r[1]), add(r[2], r[3])), add(add(r[4], r[5]), add(r[6], r[7])));
  for (; i < n; ++i) res = add(res, (T)a[i]);
  return res;
}

__global__ void __launch_bounds__(64)
sequence_targets_kernel(const float* __restrict__ q_on, const float* __restrict__ q_tg, const int* __restrict__ actions,
                        const double* __restrict__ mu, const double* __restrict__ rewards, const unsigned char* __restrict__ dones,
                        double* __restrict__ target, double* __restrict__ td_mean, unsigned char* __restrict__ td_kind, const int B,
                        const int T, const int A, const double discount, const double retrace_h, const int dbl, const int rescale,
                        const int retrace_on) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* qo = q_on + (size_t)b * (T + 1) * A;
  const float* qt = q_tg + (size_t)b * (T + 1) * A;
  double tds[kSeqMaxT];
  bool any64 = false;
  double retrace = 1.0;   // python int 1 until the first Retrace update, np.float64 afterwards (same value semantics)
  double next_td = 0.0;   // python int 0 / float32 / float64: only its value enters (the product is float64 either way)
  bool have_next = false; // false while retrace * next_td is the integer 0 (target = gain, in gain's own kind)
  const float disc32 = (float)discount;
  for (int t = T - 1; t >= 0; --t) {
    const int a = actions[(size_t)b * T + t];
    const double r = rewards[(size_t)b * T + t];
    double gain;  // value of the gain; kind tells how the reference holds it
    int kind;
    if (dones[(size_t)b * T + t]) {
      gain = r;
      kind = KPY;
      if (rescale) { gain = rescaling_f64(r); kind = K64; }
    } else {
      const float* sel = (dbl ? qo : qt) + (size_t)(t + 1) * A;
      int am = 0;
      float best = sel[0];
      for (int j = 1; j < A; ++j)
        if (sel[j] > best) { best = sel[j]; am = j; }  // np.argmax: first maximum
      float maxq = qt[(size_t)(t + 1) * A + am];
      if (rescale) maxq = inverse_rescaling_f32(maxq);
      float g32 = __fadd_rn((float)r, __fmul_rn(disc32, maxq));
      if (rescale) g32 = rescaling_f32(g32);
      gain = (double)g32;
      kind = K32;
    }
    // target_t = gain + retrace * next_td
    target[(size_t)b * T + t] = have_next ? __dadd_rn(gain, __dmul_rn(retrace, next_td)) : gain;
    // td_t = gain - Q_online(s_t)[a_t]
    const float qsa = qo[(size_t)t * A + a];
    double td;
    if (kind == K64) {
      td = __dsub_rn(gain, (double)qsa);
      any64 = true;
    } else {
      td = (double)__fsub_rn((float)gain, qsa);  // a python-float gain is rounded to float32 here (weak scalar)
    }
    tds[T - 1 - t] = td;
    if (retrace_on) {
      next_td = td;
      have_next = true;
      const float* row = qo + (size_t)t * A;
      float qmax = row[0];
      for (int j = 1; j < A; ++j) qmax = fmaxf(qmax, row[j]);
      int cnt = 0;
      for (int j = 0; j < A; ++j) cnt += (row[j] == qmax);
      const double pi = (row[a] == qmax) ? __ddiv_rn(1.0, (double)cnt) : 0.0;
      const double ratio = __ddiv_rn(pi, mu[(size_t)b * T + t]);
      const double rr = __dmul_rn(retrace_h, ratio < 1.0 ? ratio : 1.0);
      retrace = __dmul_rn(retrace, __dmul_rn(discount, rr));
    }
  }
  // np.mean(td_errors): float64 if any entry is float64, else float32; pairwise sum, then a true divide in that dtype
  if (any64) {
    td_mean[b] = __ddiv_rn(np_pairwise_sum<double>(tds, T), (double)T);
    td_kind[b] = K64;
  } else {
    td_mean[b] = (double)__fdiv_rn(np_pairwise_sum<float>(tds, T), (float)T);
    td_kind[b] = K32;
  }
}

}  // namespace srlx

extern "C" int srlx_sequence_targets(const float* q_online_dev, const float* q_target_dev, const int32_t* actions_dev,
                                     const double* mu_dev, const double* rewards_dev, const unsigned char* dones_dev,
                                     double* target_out_dev, double* td_mean_out_dev, unsigned char* td_kind_out_dev, uint32_t n_seq,
                                     uint32_t seq_len, uint32_t n_actions, double discount, double retrace_h, int enable_double_dqn,
                                     int enable_rescale, int enable_retrace, uintptr_t cuda_stream) {
  using namespace srlx;
  SRLX_REQUIRE(q_online_dev && q_target_dev && actions_dev && mu_dev && rewards_dev && dones_dev, "srlx_sequence_targets: input buffer is NULL");
  SRLX_REQUIRE(target_out_dev && td_mean_out_dev && td_kind_out_dev, "srlx_sequence_targets: output buffer is NULL");
  SRLX_REQUIRE(seq_len >= 1 && seq_len <= (uint32_t)kSeqMaxT, "srlx_sequence_targets: sequence length %u out of range [1,%d]", seq_len, kSeqMaxT);
  SRLX_REQUIRE(n_actions >= 1 && n_actions <= SRLX_MAX_ACTIONS, "srlx_sequence_targets: n_actions %u out of range", n_actions);
  if (n_seq == 0) return 0;
  sequence_targets_kernel<<<(n_seq + 63) / 64, 64, 0, (cudaStream_t)cuda_stream>>>(
      q_online_dev, q_target_dev, actions_dev, mu_dev, rewards_dev, dones_dev, target_out_dev, td_mean_out_dev, td_kind_out_dev,
      (int)n_seq, (int)seq_len, (int)n_actions, discount, retrace_h, enable_double_dqn, enable_rescale, enable_retrace);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
