// net.cuh -- Q-network forward / backward on a row tile held in shared memory (fp32 FMA; tolerance 1e-4 rel rules out
// bf16/tf32 tensor-core math for these 4..1024-wide layers).
//   dense stack   MLPBlock            srl/rl/torch_/blocks/mlp_block.py:9-47
//   dueling head  DuelingNetworkBlock srl/rl/torch_/blocks/dueling_network.py:8-59
//   noisy weights NoisyLinear         srl/rl/torch_/modules/noisy_linear.py:8-52 (W = mu + sigma * N(0,1), drawn per call)
// CPU twin: oracle/nets.py::forward / train_update.
#pragma once
#include "philox.cuh"

namespace srlx {

constexpr int kRowTile = 32;  // rows processed per tile (one lane per row in the per-row phases)

// Shared-memory plan, computed identically on host (for the launch size) and device.
struct NetPlan {
  int n_layers;
  int ldw[SRLX_MAX_LAYERS];      // padded row stride of W_l
  int w_s[SRLX_MAX_LAYERS];      // float offset of W_l inside the weight area
  int b_s[SRLX_MAX_LAYERS];      // float offset of b_l
  int weff_floats;
  int xw[SRLX_MAX_LAYERS + 1];   // logical width of activation buffer l (input of layer l; [L] = raw outputs)
  int ldx[SRLX_MAX_LAYERS + 1];  // padded stride
  int x_s[SRLX_MAX_LAYERS + 1];  // float offset inside the activation area (kRowTile rows each)
  int act_floats;
};

__host__ __device__ inline NetPlan make_plan(const srlx_net& net) {
  NetPlan p;
  p.n_layers = net.n_layers;
  int off = 0;
  for (int l = 0; l < net.n_layers; ++l) {
    p.ldw[l] = padded_ld(net.k_dim[l]);
    p.w_s[l] = off;
    off += net.out_dim[l] * p.ldw[l];
    p.b_s[l] = off;
    off += round_up(net.out_dim[l], 4);
  }
  p.weff_floats = round_up(off, 4);
  int aoff = 0;
  for (int l = 0; l <= net.n_layers; ++l) {
    p.xw[l] = (l == 0) ? net.in_dim : net.out_dim[l - 1];
    p.ldx[l] = padded_ld(p.xw[l]);
    p.x_s[l] = aoff;
    aoff += kRowTile * p.ldx[l];
  }
  p.act_floats = round_up(aoff, 4);
  return p;
}

__device__ inline int layer_of_param(const srlx_net& net, int p) {
  int l = 0;
#pragma unroll
  for (int j = 1; j < SRLX_MAX_LAYERS; ++j)
    if (j < net.n_layers && p >= net.w_off[j]) l = j;
  return l;
}

// shared-memory slot of flat parameter p
__device__ inline int weff_slot(const srlx_net& net, const NetPlan& pl, int p, int l) {
  if (p < net.b_off[l]) {
    const int q = p - net.w_off[l];
    const int u = q / net.k_dim[l], k = q - u * net.k_dim[l];
    return pl.w_s[l] + u * pl.ldw[l] + k;
  }
  return pl.b_s[l] + (p - net.b_off[l]);
}

// Effective weights of one forward call into shared memory: W = mu (+ sigma * eps(kind, call_id) on noisy layers).
__device__ inline void build_weff(const srlx_net& net, const NetPlan& pl, const float* __restrict__ mu,
                                  const float* __restrict__ sigma, bool use_noise, uint64_t seed, uint32_t kind,
                                  uint64_t call_id, float* weff) {
  const int nblk = (net.n_params + 3) >> 2;
  for (int blk = threadIdx.x; blk < nblk; blk += blockDim.x) {
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (use_noise) {
      const float4 n4 = noise4(seed, kind, call_id, (uint32_t)blk);
      z[0] = n4.x; z[1] = n4.y; z[2] = n4.z; z[3] = n4.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = 4 * blk + j;
      if (p < net.n_params) {
        const int l = layer_of_param(net, p);
        float v = __ldcg(mu + p);
        if (use_noise && net.layer_noisy[l]) v = fmaf(__ldcg(sigma + p), z[j], v);
        weff[weff_slot(net, pl, p, l)] = v;
      }
    }
  }
}

// zero the whole weight + activation areas once (padding columns must stay 0 so float4 k-loops may over-read them)
__device__ inline void zero_floats(float* p, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0.f;
}

// Hidden layer: Y[r][u] = relu(sum_k X[r][k] W[u][k] + b[u]), r < R (<= kRowTile), u < U.
// Warp tile = 4 rows x 64 units (lane -> units lane, lane+32); X rows are warp-broadcast float4 loads, W rows are
// conflict-free float4 loads (padded_ld) -> 32 FMA per 6 shared loads.  dense_relu_task is one such warp tile
// (row tile rt, unit tile ut); dense_relu_fwd spreads the tiles of one layer over the warps of the block.
__device__ __forceinline__ void dense_relu_task(const float* __restrict__ X, int ldx, int R, int K, const float* __restrict__ W,
                                                int ldw, const float* __restrict__ b, int U, float* __restrict__ Y, int ldy,
                                                int rt, int ut) {
  const int lane = threadIdx.x & 31;
  const int K4 = round_up(K, 4);
  const int r0 = rt * 4;
  const int u0 = ut * 64 + lane, u1 = u0 + 32;
  const bool v0 = u0 < U, v1 = u1 < U;
  const float* w0p = W + (v0 ? u0 : 0) * ldw;
  const float* w1p = W + (v1 ? u1 : 0) * ldw;
  const float b0 = v0 ? b[u0] : 0.f, b1 = v1 ? b[u1] : 0.f;
  float acc0[4], acc1[4];
  const float* xr[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc0[i] = b0;
    acc1[i] = b1;
    xr[i] = X + min(r0 + i, R - 1) * ldx;
  }
  for (int k = 0; k < K4; k += 4) {
    const float4 wa = *reinterpret_cast<const float4*>(w0p + k);
    const float4 wb = *reinterpret_cast<const float4*>(w1p + k);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 x = *reinterpret_cast<const float4*>(xr[i] + k);
      acc0[i] = fmaf(x.x, wa.x, acc0[i]);
      acc0[i] = fmaf(x.y, wa.y, acc0[i]);
      acc0[i] = fmaf(x.z, wa.z, acc0[i]);
      acc0[i] = fmaf(x.w, wa.w, acc0[i]);
      acc1[i] = fmaf(x.x, wb.x, acc1[i]);
      acc1[i] = fmaf(x.y, wb.y, acc1[i]);
      acc1[i] = fmaf(x.z, wb.z, acc1[i]);
      acc1[i] = fmaf(x.w, wb.w, acc1[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (r0 + i < R) {
      if (v0) Y[(r0 + i) * ldy + u0] = fmaxf(acc0[i], 0.f);
      if (v1) Y[(r0 + i) * ldy + u1] = fmaxf(acc1[i], 0.f);
    }
  }
}

__device__ inline void dense_relu_fwd(const float* __restrict__ X, int ldx, int R, int K, const float* __restrict__ W,
                                      int ldw, const float* __restrict__ b, int U, float* __restrict__ Y, int ldy) {
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int n_ut = (U + 63) >> 6, n_rt = (R + 3) >> 2;
  for (int t = warp; t < n_ut * n_rt; t += nwarps) {
    const int rt = t / n_ut, ut = t - rt * n_ut;
    dense_relu_task(X, ldx, R, K, W, ldw, b, U, Y, ldy, rt, ut);
  }
}

// Output layer (+ dueling combine).  One warp per row; lanes stride over k; nout <= 1 + SRLX_MAX_ACTIONS.
//   raw[r][o] = sum_k X[r][koff(o) + k] W[o][k] + b[o];   koff = 0 for plain nets and for the V row, H for advantage rows
//   Q[r][a]   = raw (plain) | V + A_a - mean(A) | V + A_a - max(A) | V + A_a
__device__ inline void out_layer_fwd(const srlx_net& net, const float* __restrict__ X, int ldx, int R,
                                     const float* __restrict__ W, int ldw, const float* __restrict__ b,
                                     float* __restrict__ raw, int ldr, float* __restrict__ Q, int ldq) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int L = net.n_layers, nout = net.out_dim[L - 1], K = net.k_dim[L - 1], A = net.n_actions;
  for (int r = warp; r < R; r += nwarps) {
    const float* x = X + r * ldx;
    for (int o = 0; o < nout; ++o) {
      const int koff = (net.dueling != SRLX_DUEL_NONE && o > 0) ? K : 0;
      const float* w = W + o * ldw;
      float acc = 0.f;
      for (int k = lane; k < K; k += 32) acc = fmaf(x[koff + k], w[k], acc);
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
      if (lane == 0) raw[r * ldr + o] = acc + b[o];
    }
    __syncwarp();
    if (lane == 0) {
      const float* rr = raw + r * ldr;
      if (net.dueling == SRLX_DUEL_NONE) {
        for (int a = 0; a < A; ++a) Q[r * ldq + a] = rr[a];
      } else {
        const float v = rr[0];
        float red = 0.f;
        if (net.dueling == SRLX_DUEL_AVERAGE) {
          for (int a = 0; a < A; ++a) red += rr[1 + a];
          red = red / (float)A;
        } else if (net.dueling == SRLX_DUEL_MAX) {
          red = rr[1];
          for (int a = 1; a < A; ++a) red = fmaxf(red, rr[1 + a]);
        }
        for (int a = 0; a < A; ++a) Q[r * ldq + a] = v + rr[1 + a] - red;
      }
    }
  }
}

// Full forward of a row tile.  acts = activation area (plan offsets), acts[x_s[0]] must hold the inputs.
// Leaves the hidden activations in place (needed by backward) and writes Q[r][0..A).
__device__ inline void net_forward_tile(const srlx_net& net, const NetPlan& pl, const float* weff, float* acts, int R,
                                        float* Q, int ldq) {
  const int L = net.n_layers;
  for (int l = 0; l < L - 1; ++l) {
    dense_relu_fwd(acts + pl.x_s[l], pl.ldx[l], R, net.k_dim[l], weff + pl.w_s[l], pl.ldw[l], weff + pl.b_s[l],
                   net.out_dim[l], acts + pl.x_s[l + 1], pl.ldx[l + 1]);
    __syncthreads();
  }
  out_layer_fwd(net, acts + pl.x_s[L - 1], pl.ldx[L - 1], R, weff + pl.w_s[L - 1], pl.ldw[L - 1], weff + pl.b_s[L - 1],
                acts + pl.x_s[L], pl.ldx[L], Q, ldq);
  __syncthreads();
}

// Backward of a row tile.  dQ[r][a] is the loss gradient wrt Q; G (flat parameter layout) is ACCUMULATED into.
// The hidden activations in `acts` are overwritten by their gradients on the way down.
__device__ inline void net_backward_tile(const srlx_net& net, const NetPlan& pl, const float* weff, float* acts, int R,
                                         const float* dQ, int lddq, float* G) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int L = net.n_layers, A = net.n_actions;
  const int nout = net.out_dim[L - 1], Ko = net.k_dim[L - 1];
  float* raw = acts + pl.x_s[L];  // reuse: raw[r][o] <- d loss / d raw output
  const int ldr = pl.ldx[L];
  // ---- dueling combine backward (dueling_network.py:51-58) -> d raw
  for (int r = tid; r < R; r += nt) {
    if (net.dueling == SRLX_DUEL_NONE) {
      for (int a = 0; a < A; ++a) raw[r * ldr + a] = dQ[r * lddq + a];
    } else {
      float sum = 0.f;
      for (int a = 0; a < A; ++a) sum += dQ[r * lddq + a];
      int amax = 0;
      if (net.dueling == SRLX_DUEL_MAX) {
        float best = raw[r * ldr + 1];
        for (int a = 1; a < A; ++a)
          if (raw[r * ldr + 1 + a] > best) { best = raw[r * ldr + 1 + a]; amax = a; }
      }
      for (int a = 0; a < A; ++a) {
        float d = dQ[r * lddq + a];
        if (net.dueling == SRLX_DUEL_AVERAGE) d -= sum / (float)A;
        else if (net.dueling == SRLX_DUEL_MAX && a == amax) d -= sum;
        raw[r * ldr + 1 + a] = d;
      }
      raw[r * ldr + 0] = sum;
    }
  }
  __syncthreads();
  // ---- output layer: dW, db
  {
    const float* X = acts + pl.x_s[L - 1];
    const int ldx = pl.ldx[L - 1];
    for (int w = tid; w < nout * Ko; w += nt) {
      const int o = w / Ko, k = w - o * Ko;
      const int koff = (net.dueling != SRLX_DUEL_NONE && o > 0) ? Ko : 0;
      float acc = 0.f;
      for (int r = 0; r < R; ++r) acc = fmaf(raw[r * ldr + o], X[r * ldx + koff + k], acc);
      G[net.w_off[L - 1] + w] += acc;
    }
    for (int o = tid; o < nout; o += nt) {
      float acc = 0.f;
      for (int r = 0; r < R; ++r) acc += raw[r * ldr + o];
      G[net.b_off[L - 1] + o] += acc;
    }
  }
  __syncthreads();
  // ---- d hidden (input of the output layer), ReLU-masked, in place
  if (L > 1) {
    float* X = acts + pl.x_s[L - 1];
    const int ldx = pl.ldx[L - 1], width = pl.xw[L - 1];
    const float* W = weff + pl.w_s[L - 1];
    const int ldw = pl.ldw[L - 1];
    for (int w = tid; w < R * width; w += nt) {
      const int r = w / width, kk = w - r * width;
      float d = 0.f;
      if (net.dueling == SRLX_DUEL_NONE) {
        for (int o = 0; o < nout; ++o) d = fmaf(raw[r * ldr + o], W[o * ldw + kk], d);
      } else if (kk < Ko) {
        d = raw[r * ldr + 0] * W[kk];
      } else {
        for (int o = 1; o < nout; ++o) d = fmaf(raw[r * ldr + o], W[o * ldw + (kk - Ko)], d);
      }
      X[r * ldx + kk] = (X[r * ldx + kk] > 0.f) ? d : 0.f;
    }
  }
  __syncthreads();
  // ---- hidden layers, top down
  for (int l = L - 2; l >= 0; --l) {
    const float* dY = acts + pl.x_s[l + 1];
    const int ldy = pl.ldx[l + 1];
    float* X = acts + pl.x_s[l];
    const int ldx = pl.ldx[l];
    const int U = net.out_dim[l], K = net.k_dim[l];
    const int K4n = (K + 3) >> 2;
    // dW[u][k..k+3]: item = (k4, u) with u fastest -> dY conflict-free, X float4 broadcast
    for (int w = tid; w < U * K4n; w += nt) {
      const int k4 = w / U, u = w - k4 * U;
      const int k = k4 * 4;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int r = 0; r < R; ++r) {
        const float dy = dY[r * ldy + u];
        const float4 x = *reinterpret_cast<const float4*>(X + r * ldx + k);
        a0 = fmaf(dy, x.x, a0);
        a1 = fmaf(dy, x.y, a1);
        a2 = fmaf(dy, x.z, a2);
        a3 = fmaf(dy, x.w, a3);
      }
      float* g = G + net.w_off[l] + u * K + k;
      g[0] += a0;
      if (k + 1 < K) g[1] += a1;
      if (k + 2 < K) g[2] += a2;
      if (k + 3 < K) g[3] += a3;
    }
    for (int u = tid; u < U; u += nt) {
      float acc = 0.f;
      for (int r = 0; r < R; ++r) acc += dY[r * ldy + u];
      G[net.b_off[l] + u] += acc;
    }
    __syncthreads();
    if (l > 0) {
      const float* W = weff + pl.w_s[l];
      const int ldw = pl.ldw[l];
      for (int w = tid; w < R * K; w += nt) {
        const int r = w / K, k = w - r * K;
        float d = 0.f;
        for (int u = 0; u < U; ++u) d = fmaf(dY[r * ldy + u], W[u * ldw + k], d);
        X[r * ldx + k] = (X[r * ldx + k] > 0.f) ? d : 0.f;
      }
      __syncthreads();
    }
  }
}

// srl/rl/functions.py:10-17 in fp32 (numpy keeps float32 for python-float scalars)
__device__ inline float rescaling_f(float x) {
  const float s = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
  return s * (sqrtf(fabsf(x) + 1.0f) - 1.0f) + 0.001f * x;
}
__device__ inline float inverse_rescaling_f(float x) {
  const float eps = 0.001f;
  const float s = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
  float n = sqrtf(1.0f + 4.0f * eps * (fabsf(x) + 1.0f + eps)) - 1.0f;
  n = n / (2.0f * eps);
  return s * (n * n - 1.0f);
}

}  // namespace srlx
