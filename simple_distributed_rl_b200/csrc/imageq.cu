// imageq.cu -- the image configs of the reference on device:
//   srlx_image_process  ImageProcessor.remap_observation (srl/rl/processors/image_processor.py:104-154) for a batch of uint8 frames:
//                       gray / colour conversion, trimming, cv2.resize INTER_LINEAR and the normalisation in one pass
//   srlx_imageq_*       InputImageBlock + DQNImageBlock + hidden block + Linear(A) (srl/rl/torch_/blocks/dqn_image_block.py:10-62,
//                       srl/algorithms/dqn/model_torch.py:17-29) and Trainer.train (model_torch.py:75-131) with calc_target_q
//                       (srl/algorithms/dqn/dqn.py:143-173)
// Convolutions are im2col (replicate padding = clamped index, the bias column's 1 written by the same kernel) + the strided GEMM
// family of gemm.cuh; the input gradient of a convolution is the GEMM dOut x W followed by a GATHER col2im (every input pixel sums the
// window slots that read it, in a fixed order: deterministic, no atomics).  Activations are NHWC, so a conv layer's output is the next
// layer's im2col source and the last one IS the flattened input of the first dense layer.
#include "gemm.cuh"
#include "net.cuh"

namespace srlx {

// ---- image processor ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int img_px(const unsigned char* __restrict__ f, const srlx_image_proc& p, int y, int x, int c) {
  const unsigned char* s = f + ((size_t)(y + p.top) * p.src_w + (x + p.left)) * p.src_c;
  if (p.src_c == 3 && p.out_c == 1) return (s[0] * 9798 + s[1] * 19235 + s[2] * 3735 + 16384) >> 15;  // cv2 RGB2GRAY, 15-bit coefficients
  return p.src_c == 1 ? s[0] : s[c];
}

template <typename OutT>
__global__ void __launch_bounds__(256) image_process_kernel(const __grid_constant__ srlx_image_proc p, const unsigned char* __restrict__ src,
                                                            const uint32_t n, OutT* __restrict__ out, const uint64_t out_stride) {
  const long long per = (long long)p.out_h * p.out_w * p.out_c, total = per * n;
  const size_t src_frame = (size_t)p.src_h * p.src_w * p.src_c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long f = i / per;
    int r = (int)(i - f * per);
    const int c = r % p.out_c;
    r /= p.out_c;
    const int ox = r % p.out_w, oy = r / p.out_w;
    const unsigned char* fr = src + (size_t)f * src_frame;
    int v;
    if (p.resize) {
      const int x0 = p.x_idx[ox], x1 = min(x0 + 1, p.trim_w - 1), a0 = p.x_coef[2 * ox], a1 = p.x_coef[2 * ox + 1];
      const int yr = p.y_idx[oy], y0 = min(max(yr, 0), p.trim_h - 1), y1 = min(max(yr + 1, 0), p.trim_h - 1);
      const int b0 = p.y_coef[2 * oy], b1 = p.y_coef[2 * oy + 1];
      const int r0 = img_px(fr, p, y0, x0, c) * a0 + img_px(fr, p, y0, x1, c) * a1;
      const int r1 = img_px(fr, p, y1, x0, c) * a0 + img_px(fr, p, y1, x1, c) * a1;
      v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;  // VResizeLinear, FixedPtCast<int, uchar, 22>
      v = min(max(v, 0), 255);
    } else {
      v = img_px(fr, p, oy, ox, c);
    }
    const size_t o = (size_t)f * out_stride + (size_t)(i - f * per);
    if constexpr (sizeof(OutT) == 1) {
      out[o] = (OutT)v;
    } else {
      const float x = (float)v;  // image_processor.py:141-146, float32 arithmetic
      out[o] = p.normalize == 1 ? __fdiv_rn(x, p.max_val) : __fsub_rn(__fdiv_rn(__fmul_rn(x, 2.f), p.max_val), 1.f);
    }
  }
}

// ---- im2col / col2im ----------------------------------------------------------------------------------------------------------------
struct ConvG {
  int C, H, W, k, s, p, OH, OW, K;       // K = C * k * k
  long long sb, sc, sh, sw;              // element strides of the source
  int c_fast;                            // column order (kh, kw, c) instead of (c, kh, kw)
};

__device__ __forceinline__ void col_split(const ConvG& g, int j, int& c, int& kh, int& kw) {
  if (g.c_fast) { c = j % g.C; j /= g.C; kw = j % g.k; kh = j / g.k; }
  else { kw = j % g.k; j /= g.k; kh = j % g.k; c = j / g.k; }
}

template <typename InT>
__global__ void __launch_bounds__(256) im2col_kernel(const ConvG g, const InT* __restrict__ in, const float inv_div, float* __restrict__ col,
                                                     const long long rows) {
  const int ld = g.K + 1;
  const long long total = rows * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / ld;
    const int j = (int)(i - row * ld);
    float v = 1.f;  // the bias column
    if (j < g.K) {
      int c, kh, kw;
      col_split(g, j, c, kh, kw);
      const int ow = (int)(row % g.OW), oh = (int)((row / g.OW) % g.OH);
      const long long b = row / ((long long)g.OW * g.OH);
      const int ih = min(max(oh * g.s - g.p + kh, 0), g.H - 1), iw = min(max(ow * g.s - g.p + kw, 0), g.W - 1);  // padding_mode="replicate"
      const InT x = in[b * g.sb + c * g.sc + ih * g.sh + iw * g.sw];
      if constexpr (sizeof(InT) == 1) v = __fdiv_rn((float)x, inv_div); else v = x;
    }
    col[i] = v;
  }
}

// dIn[b][ih][iw][c] = (act > 0) * sum over the window slots (oh, kh, ow, kw) whose clamped source is (ih, iw) of dcol[(b, oh, ow)][(kh, kw, c)]
__global__ void __launch_bounds__(256) col2im_kernel(const ConvG g, const float* __restrict__ dcol, const float* __restrict__ act,
                                                     float* __restrict__ din, const long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % g.C);
    const int iw = (int)((i / g.C) % g.W), ih = (int)((i / ((long long)g.C * g.W)) % g.H);
    const long long b = i / ((long long)g.C * g.W * g.H);
    float sum = 0.f;
    if (act[i] > 0.f) {
      const int oh_lo = ih == 0 ? 0 : max(0, (ih + g.p - g.k + 1 + g.s - 1) / g.s), oh_hi = ih == g.H - 1 ? g.OH - 1 : min(g.OH - 1, (ih + g.p) / g.s);
      const int ow_lo = iw == 0 ? 0 : max(0, (iw + g.p - g.k + 1 + g.s - 1) / g.s), ow_hi = iw == g.W - 1 ? g.OW - 1 : min(g.OW - 1, (iw + g.p) / g.s);
      for (int oh = oh_lo; oh <= oh_hi; ++oh)
        for (int kh = 0; kh < g.k; ++kh) {
          if (min(max(oh * g.s - g.p + kh, 0), g.H - 1) != ih) continue;
          for (int ow = ow_lo; ow <= ow_hi; ++ow) {
            const float* row = dcol + ((b * g.OH + oh) * g.OW + ow) * (long long)g.K;
            for (int kw = 0; kw < g.k; ++kw) {
              if (min(max(ow * g.s - g.p + kw, 0), g.W - 1) != iw) continue;
              const int j = g.c_fast ? (kh * g.k + kw) * g.C + c : (c * g.k + kh) * g.k + kw;
              sum += row[j];
            }
          }
        }
    }
    din[i] = sum;
  }
}

// ---- loss, Adam ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rescaling_d(double x) {  // srl/rl/functions.py:10-12 on a float64 array
  const double s = (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : 0.0);
  return s * (sqrt(fabs(x) + 1.0) - 1.0) + 0.001 * x;
}

// one CTA: calc_target_q (dqn.py:143-173, without invalid actions), q = sum(Q * onehot), nn.HuberLoss()(target * w, q * w) (mean), its
// gradient wrt Q, priorities = |target - q|
__global__ void __launch_bounds__(256) imageq_loss_kernel(const srlx_imageq q, const float* __restrict__ q0, const float* __restrict__ qn_online,
                                                          const float* __restrict__ qn_target, const int32_t* __restrict__ action,
                                                          const float* __restrict__ reward, const float* __restrict__ undone,
                                                          const float* __restrict__ weights, const int B, float* __restrict__ dq,
                                                          float* __restrict__ pri, float* __restrict__ loss_out, float* __restrict__ tq_buf) {
  __shared__ float red[256];
  const int A = q.n_actions;
  float part = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* nt = qn_target + (size_t)b * A;
    float maxq;
    if (q.enable_double_dqn) {
      const float* no = qn_online + (size_t)b * A;
      int best = 0;
      for (int a = 1; a < A; ++a)
        if (no[a] > no[best]) best = a;  // np.argmax: first maximum
      maxq = nt[best];
    } else {
      maxq = nt[0];
      for (int a = 1; a < A; ++a) maxq = fmaxf(maxq, nt[a]);
    }
    if (q.enable_rescale) maxq = inverse_rescaling_f(maxq);
    // reward (f32) + undone (int array) * discount (python float) * maxq (f32): numpy promotes to float64, then .astype(float32)
    double t = (double)reward[b] + ((double)undone[b] * q.discount) * (double)maxq;
    if (q.enable_rescale) t = rescaling_d(t);
    const float tq = (float)t;
    const int a_sel = action[b];
    const float qv = q0[(size_t)b * A + a_sel], w = weights[b];
    const float x = __fsub_rn(__fmul_rn(qv, w), __fmul_rn(tq, w));
    const float ax = fabsf(x);
    part += ax <= 1.f ? 0.5f * x * x : ax - 0.5f;
    const float gx = fminf(fmaxf(x, -1.f), 1.f);
    for (int a = 0; a < A; ++a) dq[(size_t)b * A + a] = a == a_sel ? w * gx / (float)B : 0.f;
    pri[b] = fabsf(tq - qv);
    if (tq_buf) tq_buf[b] = tq;
  }
  red[threadIdx.x] = part;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_out = red[0] / (float)B;
}

// torch.optim.Adam (_single_tensor_adam) + the target sync of model_torch.py:124-127 (the check runs BEFORE train_count += 1)
__global__ void __launch_bounds__(256) imageq_adam_kernel(const srlx_imageq q) {
  const uint64_t tc = q.counters[0];
  const double t = (double)(q.counters[1] + 1);
  const float step_size = (float)(q.lr / (1.0 - pow(q.adam_beta1, t))), bc2_sqrt = (float)sqrt(1.0 - pow(q.adam_beta2, t));
  const float b1 = (float)q.adam_beta1, b2 = (float)q.adam_beta2, eps = (float)q.adam_eps;
  const bool sync = tc % (uint64_t)q.target_update_interval == 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < q.n_params; i += gridDim.x * blockDim.x) {
    const float g = q.grads[i];
    float p = q.params[i], m = q.adam_m[i], v = q.adam_v[i];
    m = m + (g - m) * (1.0f - b1);
    v = v * b2 + (1.0f - b2) * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - step_size * (m / denom);
    q.params[i] = p;
    q.adam_m[i] = m;
    q.adam_v[i] = v;
    if (sync) q.target[i] = p;
  }
}
__global__ void imageq_count_kernel(const srlx_imageq q) {
  if (q.counters[0] % (uint64_t)q.target_update_interval == 0) q.counters[2] += 1;
  q.counters[0] += 1;
  q.counters[1] += 1;
}

__global__ void fill_kernel(float* p, const long long n, const long long stride, const float v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i * stride] = v;
}

// ---- workspace plan -----------------------------------------------------------------------------------------------------------------
struct ImageQPlan {
  ConvG g[SRLX_MAX_CONV];
  long long rows[SRLX_MAX_CONV];          // per sample: OH * OW
  size_t col[SRLX_MAX_CONV], cact[SRLX_MAX_CONV], dcact[SRLX_MAX_CONV], dcol;
  size_t act[SRLX_MAX_LAYERS], dact[SRLX_MAX_LAYERS];
  size_t ones, qbuf, dq, tq, split;
  size_t split_floats, total;
  int flat;                               // inputs of dense 0
};

static int imageq_plan(const srlx_imageq* q, ImageQPlan& pl) {
  SRLX_REQUIRE(q->n_conv >= 1 && q->n_conv <= SRLX_MAX_CONV && q->n_dense >= 1 && q->n_dense <= SRLX_MAX_LAYERS, "imageq: 1..%d conv layers, 1..%d dense layers", SRLX_MAX_CONV, SRLX_MAX_LAYERS);
  SRLX_REQUIRE(q->batch_cap >= 1 && q->n_actions >= 1 && q->dense_out[q->n_dense - 1] == q->n_actions, "imageq: the last dense layer has n_actions rows");
  const long long B = q->batch_cap;
  size_t off = 0;
  auto take = [&](size_t n) { const size_t o = off; off += (n + 3) / 4 * 4; return o; };
  int C = q->in_c, H = q->in_h, W = q->in_w, off_p = 0;
  size_t dcol_max = 0, split_max = 0;
  for (int l = 0; l < q->n_conv; ++l) {
    ConvG& g = pl.g[l];
    g.C = C; g.H = H; g.W = W; g.k = q->conv_k[l]; g.s = q->conv_s[l]; g.p = q->conv_p[l];
    g.OH = (H + 2 * g.p - g.k) / g.s + 1; g.OW = (W + 2 * g.p - g.k) / g.s + 1; g.K = C * g.k * g.k;
    SRLX_REQUIRE(g.k >= 1 && g.s >= 1 && g.p >= 0 && H + 2 * g.p >= g.k && W + 2 * g.p >= g.k, "imageq: conv layer %d has an empty output", l);
    SRLX_REQUIRE(q->conv_oh[l] == g.OH && q->conv_ow[l] == g.OW && q->conv_off[l] == off_p, "imageq: conv layer %d: conv_oh / conv_ow / conv_off do not follow from the shapes (%d x %d at %d expected)", l, g.OH, g.OW, off_p);
    if (l == 0) { g.sb = q->in_sb; g.sc = q->in_sc; g.sh = q->in_sh; g.sw = q->in_sw; }
    else { g.sc = 1; g.sw = C; g.sh = (long long)C * W; g.sb = (long long)C * W * H; }
    g.c_fast = g.sc == 1;
    pl.rows[l] = (long long)g.OH * g.OW;
    const int F = q->conv_f[l];
    pl.col[l] = take((size_t)B * pl.rows[l] * (g.K + 1));
    pl.cact[l] = take((size_t)B * pl.rows[l] * F);
    pl.dcact[l] = take((size_t)B * pl.rows[l] * F);
    if (l > 0 && (size_t)B * pl.rows[l] * g.K > dcol_max) dcol_max = (size_t)B * pl.rows[l] * g.K;
    if ((size_t)F * (g.K + 1) > split_max) split_max = (size_t)F * (g.K + 1);
    off_p += F * (g.K + 1);
    C = F; H = g.OH; W = g.OW;
  }
  pl.flat = C * H * W;
  pl.dcol = take(dcol_max);
  int k = pl.flat;
  for (int l = 0; l < q->n_dense; ++l) {
    SRLX_REQUIRE(q->dense_k[l] == k && q->dense_off[l] == off_p, "imageq: dense layer %d: dense_k / dense_off do not follow from the shapes (%d at %d expected)", l, k, off_p);
    const int out = q->dense_out[l], last = l == q->n_dense - 1;
    pl.act[l] = take((size_t)B * (out + (last ? 0 : 1)));
    pl.dact[l] = take((size_t)B * out);
    if ((size_t)out > split_max) split_max = out;
    if (l > 0 && (size_t)out * (k + 1) > split_max) split_max = (size_t)out * (k + 1);
    off_p += out * (k + 1);
    k = out;
  }
  SRLX_REQUIRE(q->n_params == off_p, "imageq: n_params = %d, the layers hold %d", q->n_params, off_p);
  pl.ones = take((size_t)B);
  pl.qbuf = take((size_t)3 * B * q->n_actions);
  pl.dq = take((size_t)B * q->n_actions);
  pl.tq = take((size_t)B + 4);
  pl.split_floats = 32 * split_max;
  pl.split = take(pl.split_floats);
  pl.total = off;
  return 0;
}

static unsigned grid_for(long long n) { const long long g = (n + 255) / 256; return (unsigned)(g < 148 * 8 ? (g < 1 ? 1 : g) : 148 * 8); }

// forward of n samples with parameter buffer P; leaves col / cact / act of the pass in the workspace, Q in qout [n][A]
static int imageq_forward(const srlx_imageq* q, const ImageQPlan& pl, const float* P, const void* state, int n, float* qout, cudaStream_t s) {
  float* ws = q->ws;
  const Gate open{nullptr, 0};
  for (int l = 0; l < q->n_conv; ++l) {
    const ConvG& g = pl.g[l];
    const long long rows = (long long)n * pl.rows[l];
    if (l == 0 && q->in_u8) im2col_kernel<unsigned char><<<grid_for(rows * (g.K + 1)), 256, 0, s>>>(g, (const unsigned char*)state, q->in_max_val, ws + pl.col[l], rows);
    else im2col_kernel<float><<<grid_for(rows * (g.K + 1)), 256, 0, s>>>(g, l == 0 ? (const float*)state : ws + pl.cact[l - 1], 1.f, ws + pl.col[l], rows);
    count_launch();
    GemmP p{};
    p.A = ws + pl.col[l]; p.sa_m = g.K + 1; p.sa_k = 1;
    p.B = P + q->conv_off[l]; p.sb_k = 1; p.sb_n = g.K + 1;
    p.C = ws + pl.cact[l]; p.ldc = q->conv_f[l];
    p.M = (int)rows; p.N = q->conv_f[l]; p.K = g.K + 1; p.relu = 1; p.gate = open;
    if (launch_gemm(p, 1, s)) return -1;
  }
  for (int l = 0; l < q->n_dense; ++l) {
    const int k = q->dense_k[l], out = q->dense_out[l], last = l == q->n_dense - 1;
    const float* Wl = P + q->dense_off[l];
    GemmP p{};
    p.B = Wl; p.sb_k = 1; p.sb_n = k + 1;
    p.C = last ? qout : ws + pl.act[l]; p.ldc = last ? out : out + 1;
    p.M = n; p.N = out; p.gate = open;
    if (l == 0) {  // the flattened conv output has no constant column: bias as a second, K = 1 map against the ones vector
      p.A = ws + pl.cact[q->n_conv - 1]; p.sa_m = k; p.sa_k = 1; p.K = k;
      if (launch_gemm(p, 1, s)) return -1;
      GemmP b = p;
      b.A = ws + pl.ones; b.sa_m = 1; b.sa_k = 1; b.B = Wl + k; b.K = 1; b.accumulate = 1; b.relu = !last;
      if (launch_gemm(b, 1, s)) return -1;
    } else {
      p.A = ws + pl.act[l - 1]; p.sa_m = k + 1; p.sa_k = 1; p.K = k + 1; p.relu = !last;
      if (launch_gemm(p, 1, s)) return -1;
    }
  }
  return 0;
}

static int imageq_check(const srlx_imageq* q, ImageQPlan& pl) {
  SRLX_REQUIRE(q != nullptr, "imageq: null handle");
  if (int rc = imageq_plan(q, pl)) return rc;
  SRLX_REQUIRE(q->params && q->target && q->ws && q->ws_floats >= pl.total, "imageq: params / target / ws missing or ws too small (%llu floats needed)", (unsigned long long)pl.total);
  return 0;
}

}  // namespace srlx

using namespace srlx;

extern "C" {

size_t srlx_sizeof_imageq(void) { return sizeof(srlx_imageq); }

int srlx_image_linear_table(int32_t dst, int32_t src, int border_reset, int32_t* idx, int32_t* coef) {
  SRLX_REQUIRE(dst > 0 && src > 0 && idx && coef, "image table: dst, src > 0 and two output arrays");
  const double scale = 1.0 / ((double)dst / (double)src);  // resize.cpp: inv_scale = dsize / ssize, scale = 1 / inv_scale
  for (int d = 0; d < dst; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (border_reset) {
      if (s < 0) { s = 0; f = 0.f; }
      if (s >= src - 1) { s = src - 1; f = 0.f; }
    }
    idx[d] = s;
    coef[2 * d] = (int32_t)lrintf((1.f - f) * 2048.f);  // saturate_cast<short>(float) = cvRound
    coef[2 * d + 1] = (int32_t)lrintf(f * 2048.f);
  }
  return 0;
}

int srlx_image_process(const srlx_image_proc* p, const unsigned char* src, uint32_t n, void* out, uint64_t out_stride, uintptr_t stream) {
  SRLX_REQUIRE(p && src && out, "image_process: null argument");
  SRLX_REQUIRE((p->src_c == 1 || p->src_c == 3) && (p->out_c == 1 || p->out_c == 3), "image_process: 1 or 3 channels (got %d -> %d)", p->src_c, p->out_c);
  SRLX_REQUIRE(p->top >= 0 && p->left >= 0 && p->trim_h >= 1 && p->trim_w >= 1 && p->top + p->trim_h <= p->src_h && p->left + p->trim_w <= p->src_w,
               "image_process: trimming window outside the frame");
  SRLX_REQUIRE(p->resize ? (p->x_idx && p->x_coef && p->y_idx && p->y_coef) : (p->out_h == p->trim_h && p->out_w == p->trim_w),
               "image_process: resize needs the four tables; without resize the output is the trimming window");
  SRLX_REQUIRE(p->normalize >= 0 && p->normalize <= 2 && out_stride >= (uint64_t)p->out_h * p->out_w * p->out_c, "image_process: normalize in 0..2, frame stride >= frame size");
  if (n == 0) return 0;
  const long long total = (long long)n * p->out_h * p->out_w * p->out_c;
  cudaStream_t s = (cudaStream_t)stream;
  if (p->normalize == 0) image_process_kernel<unsigned char><<<grid_for(total), 256, 0, s>>>(*p, src, n, (unsigned char*)out, out_stride);
  else image_process_kernel<float><<<grid_for(total), 256, 0, s>>>(*p, src, n, (float*)out, out_stride);
  count_launch();
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

uint64_t srlx_imageq_ws_floats(const srlx_imageq* q) {
  ImageQPlan pl;
  if (!q || imageq_plan(q, pl)) return 0;
  return pl.total;
}

int srlx_imageq_init(const srlx_imageq* q, uintptr_t stream) {
  ImageQPlan pl;
  if (int rc = imageq_check(q, pl)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const long long B = q->batch_cap;
  fill_kernel<<<grid_for(B), 256, 0, s>>>(q->ws + pl.ones, B, 1, 1.f);
  for (int l = 0; l + 1 < q->n_dense; ++l) fill_kernel<<<grid_for(B), 256, 0, s>>>(q->ws + pl.act[l] + q->dense_out[l], B, q->dense_out[l] + 1, 1.f);
  count_launch(q->n_dense);
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int srlx_imageq_forward(const srlx_imageq* q, int use_target, const void* state, uint32_t n, float* q_out, uintptr_t stream) {
  ImageQPlan pl;
  if (int rc = imageq_check(q, pl)) return rc;
  SRLX_REQUIRE(state && q_out && n >= 1 && n <= (uint32_t)q->batch_cap, "imageq_forward: 1 <= n <= batch_cap (%d), got %u", q->batch_cap, n);
  if (imageq_forward(q, pl, use_target ? q->target : q->params, state, (int)n, q_out, (cudaStream_t)stream)) return -1;
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int srlx_imageq_train(const srlx_imageq* q, const void* state, const void* n_state, const int32_t* action, const float* reward,
                      const float* undone, const float* weights, uint32_t batch, float* pri_out, float* loss_out, float* tq_out, int phases,
                      uintptr_t stream) {
  ImageQPlan pl;
  if (int rc = imageq_check(q, pl)) return rc;
  SRLX_REQUIRE(q->adam_m && q->adam_v && q->grads && q->counters, "imageq_train: adam_m / adam_v / grads / counters missing");
  SRLX_REQUIRE(batch >= 1 && batch <= (uint32_t)q->batch_cap, "imageq_train: 1 <= batch <= batch_cap (%d), got %u", q->batch_cap, batch);
  SRLX_REQUIRE(phases >= 1 && phases <= 3, "imageq_train: phases in 1..3");
  SRLX_REQUIRE(q->target_update_interval >= 1, "imageq_train: target_update_interval >= 1");
  cudaStream_t s = (cudaStream_t)stream;
  float* ws = q->ws;
  const int B = (int)batch, A = q->n_actions;
  const Gate open{nullptr, 0};
  if (phases & 1) {
    SRLX_REQUIRE(state && n_state && action && reward && undone && weights && pri_out && loss_out, "imageq_train: null batch array");
    float* qb = ws + pl.qbuf;
    // calc_target_q: pred_target_q(n_state), pred_q(n_state) for double DQN (dqn.py:154-162), then the online pass whose activations the
    // backward pass reads
    if (imageq_forward(q, pl, q->target, n_state, B, qb + (size_t)2 * B * A, s)) return -1;
    if (q->enable_double_dqn && imageq_forward(q, pl, q->params, n_state, B, qb + (size_t)B * A, s)) return -1;
    if (imageq_forward(q, pl, q->params, state, B, qb, s)) return -1;
    imageq_loss_kernel<<<1, 256, 0, s>>>(*q, qb, qb + (size_t)B * A, qb + (size_t)2 * B * A, action, reward, undone, weights, B, ws + pl.dq, pri_out,
                                         loss_out, tq_out);
    count_launch();
    // dense layers, last to first
    const float* dout = ws + pl.dq;
    int ld_dout = A;
    for (int l = q->n_dense - 1; l >= 0; --l) {
      const int k = q->dense_k[l], out = q->dense_out[l];
      const float* X = l == 0 ? ws + pl.cact[q->n_conv - 1] : ws + pl.act[l - 1];
      const int ldx = l == 0 ? k : k + 1;
      GemmP w{};  // dW[out][k (+1)] = dOut^T x X
      w.A = dout; w.sa_m = 1; w.sa_k = ld_dout;
      w.B = X; w.sb_k = ldx; w.sb_n = 1;
      w.C = q->grads + q->dense_off[l]; w.ldc = k + 1;
      w.M = out; w.N = l == 0 ? k : k + 1; w.K = B; w.gate = open;
      if (launch_gemm(w, 1, s, ws + pl.split, pl.split_floats, true)) return -1;
      if (l == 0) {
        GemmP b = w;
        b.B = ws + pl.ones; b.sb_k = 1; b.sb_n = 1; b.C = q->grads + q->dense_off[l] + k; b.N = 1;
        if (launch_gemm(b, 1, s, ws + pl.split, pl.split_floats)) return -1;
      }
      GemmP x{};  // dX[B][k] = (X > 0) * dOut x W[:, :k]
      x.A = dout; x.sa_m = ld_dout; x.sa_k = 1;
      x.B = q->params + q->dense_off[l]; x.sb_k = k + 1; x.sb_n = 1;
      x.C = l == 0 ? ws + pl.dcact[q->n_conv - 1] : ws + pl.dact[l - 1]; x.ldc = k;
      x.M = B; x.N = k; x.K = out; x.mask = X; x.ldmask = ldx; x.gate = open;
      if (launch_gemm(x, 1, s)) return -1;
      dout = x.C;
      ld_dout = k;
    }
    // conv layers, last to first
    for (int l = q->n_conv - 1; l >= 0; --l) {
      const ConvG& g = pl.g[l];
      const int F = q->conv_f[l];
      const long long rows = (long long)B * pl.rows[l];
      GemmP w{};  // dW[F][K+1] = dOut^T x col
      w.A = ws + pl.dcact[l]; w.sa_m = 1; w.sa_k = F;
      w.B = ws + pl.col[l]; w.sb_k = g.K + 1; w.sb_n = 1;
      w.C = q->grads + q->conv_off[l]; w.ldc = g.K + 1;
      w.M = F; w.N = g.K + 1; w.K = (int)rows; w.gate = open;
      if (launch_gemm(w, 1, s, ws + pl.split, pl.split_floats, true)) return -1;
      if (l == 0) break;
      GemmP x{};  // dcol[rows][K] = dOut x W[:, :K]
      x.A = ws + pl.dcact[l]; x.sa_m = F; x.sa_k = 1;
      x.B = q->params + q->conv_off[l]; x.sb_k = g.K + 1; x.sb_n = 1;
      x.C = ws + pl.dcol; x.ldc = g.K;
      x.M = (int)rows; x.N = g.K; x.K = F; x.gate = open;
      if (launch_gemm(x, 1, s)) return -1;
      const long long n_in = (long long)B * g.H * g.W * g.C;
      col2im_kernel<<<grid_for(n_in), 256, 0, s>>>(g, ws + pl.dcol, ws + pl.cact[l - 1], ws + pl.dcact[l - 1], n_in);
      count_launch();
    }
  }
  if (phases & 2) {
    imageq_adam_kernel<<<grid_for(q->n_params), 256, 0, s>>>(*q);
    imageq_count_kernel<<<1, 1, 0, s>>>(*q);
    count_launch(2);
  }
  SRLX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
