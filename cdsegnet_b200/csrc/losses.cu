// Criteria of the DefaultSegmentorV2 wrapper, forward values (pointcept/models/losses/{builder,misc,lovasz}.py) and the
// elementwise diffusion samplers (pointcept/models/default.py:192-222).  Everything stays on the device: the wrapper's
// `loss` is produced without a host round trip.
//
//   ce_softmax_kernel   one warp per point: softmax probabilities (kept for Lovasz), -log p[target] summed over the points
//                       whose label != ignore_index, valid count, per-class label counts        (misc.py:97-129)
//   lovasz_keys_kernel  per (class, point): error |fg - p_c| as a DESCENDING 31-bit radix key     (lovasz.py:137-140)
//   -> cdseg_argsort_rows (serialize.cu) sorts all classes at once (rows = classes)
//   lovasz_scan_kernel  one CTA per class: running foreground count along the sorted order -> Jaccard gradient -> dot
//                       with the sorted errors                                                     (lovasz.py:22-33, 141-143)
//   mse_kernel          masked mean squared error of the Noise Network's prediction               (misc.py:24-94)
//   criteria_finalize   MSE, CE, Lovasz means and the EW / GLS combinations                        (builder.py:24-49)
#include "common.cuh"
#include "../../include/cdseg_b200.h"

namespace ls {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// acc[0] += sum of -log softmax(logits)[target] over valid points, acc[1] += number of valid points
__global__ void __launch_bounds__(256) ce_softmax_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, int64_t n,
                                                         int C, int64_t ignore, float* __restrict__ prob, double* __restrict__ acc,
                                                         int32_t* __restrict__ cls_count) {
  __shared__ double s_nll[8];
  __shared__ int s_cnt[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double nll = 0.0;
  int cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * 8 + warp; i < n; i += (int64_t)gridDim.x * 8) {
    const float* row = logits + i * C;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(row[c] - m);
    s = warp_sum(s);
    if (prob) {
      const float inv = 1.0f / s;
      for (int c = lane; c < C; c += 32) prob[i * C + c] = expf(row[c] - m) * inv;
    }
    const int64_t t = target[i];
    if (t != ignore && t >= 0 && t < C && lane == 0) {
      nll += (double)(logf(s) + m - row[t]);
      ++cnt;
      if (cls_count) atomicAdd(cls_count + t, 1);
    }
  }
  if (lane == 0) { s_nll[warp] = nll; s_cnt[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    int b = 0;
    for (int w = 0; w < 8; ++w) { a += s_nll[w]; b += s_cnt[w]; }
    if (b) { atomicAdd(acc, a); atomicAdd(acc + 1, (double)b); }
  }
}

// key[c][i] = 0x3F800000 - bits(|fg - p|)  (errors lie in [0, 1], so ascending keys = descending errors);
// ignored points get the largest key: they sort behind every valid point and carry zero error
__global__ void lovasz_keys_kernel(const float* __restrict__ prob, const int64_t* __restrict__ target, int64_t n, int C, int64_t ignore,
                                   int64_t* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (i >= n) return;
  const int64_t t = target[i];
  int64_t k = 0x3F800001ll;
  if (t != ignore && t >= 0 && t < C) {
    const float fg = t == c ? 1.0f : 0.0f;
    const float e = fminf(fabsf(fg - prob[i * C + c]), 1.0f);
    k = 0x3F800000ll - (int64_t)__float_as_uint(e);
  }
  keys[(int64_t)c * n + i] = k;
}

// one CTA per class c: loss[c] = sum_k err_sorted[k] * (J_k - J_{k-1}),  J_k = 1 - (G - F_k) / (G + (k + 1) - F_k),
// F_k = foreground points among the first k + 1 of the descending-error order, G = all foreground points of the class
constexpr int LV_THREADS = 1024;
__global__ void __launch_bounds__(LV_THREADS) lovasz_scan_kernel(const float* __restrict__ prob, const int64_t* __restrict__ target,
                                                                  const int32_t* __restrict__ order, const int32_t* __restrict__ cls_count,
                                                                  int64_t n, int C, int64_t ignore, double* __restrict__ loss,
                                                                  float* __restrict__ gbuf) {
  __shared__ int s_warp[32];
  __shared__ double s_red[32];
  __shared__ int s_carry;
  const int c = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = cls_count[c];
  if (G == 0) { if (threadIdx.x == 0) loss[c] = 0.0; return; }    // absent class: skipped by the reference (labels.unique())
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const float gts = (float)G;
  double part = 0.0;
  const int32_t* ord = order + (int64_t)c * n;
  for (int64_t base = 0; base < n; base += LV_THREADS) {
    const int64_t k = base + threadIdx.x;
    int fg = 0;
    float err = 0.f;
    bool valid = false;
    if (k < n) {
      const int64_t i = ord[k];
      const int64_t t = target[i];
      valid = t != ignore && t >= 0 && t < C;
      if (valid) {
        fg = t == c;
        err = fabsf((fg ? 1.0f : 0.0f) - prob[i * C + c]);
      }
    }
    int inc = fg;                                       // inclusive scan of fg over the CTA
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
      s_warp[lane] = wi - w;                            // exclusive prefix of the warp totals
    }
    __syncthreads();
    const int carry = s_carry;
    const int F = carry + s_warp[warp] + inc;           // foreground among sorted positions 0..k
    if (valid) {
      // the reference evaluates these in fp32 (lovasz.py:27-32): intersection = gts - cumsum(fg), union = gts + cumsum(1 - fg)
      const float Fk = (float)F, Fp = (float)(F - fg);
      const float jk = 1.0f - (gts - Fk) / (gts + ((float)(k + 1) - Fk));
      const float jp = k > 0 ? 1.0f - (gts - Fp) / (gts + ((float)k - Fp)) : 0.0f;
      part += (double)err * (double)(jk - jp);
      // d loss_c / d p[i, c]: the Jaccard gradient is a constant of the sort order (autograd sees errors_sorted only, lovasz.py:141-143);
      // |fg - p| has slope -1 on foreground points and +1 elsewhere
      if (gbuf) gbuf[(int64_t)ord[k] * C + c] = fg ? -(jk - jp) : (jk - jp);
    }
    __syncthreads();
    if (threadIdx.x == LV_THREADS - 1) s_carry = F;
    __syncthreads();
  }
  part = warp_sum_d(part);
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < 32; ++w) a += s_red[w];
    loss[c] = a;
  }
}

// acc[0] += sum over valid points and channels of (pred - target)^2, acc[1] += number of summed elements
__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, const int64_t* __restrict__ label,
                                                  int64_t n, int C, int64_t ignore, int use_ignore, double* __restrict__ acc) {
  __shared__ double s_a[8], s_b[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double a = 0.0, b = 0.0;
  const int64_t total = n * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / C;
    if (use_ignore && label && label[i] == ignore) continue;
    const float d = pred[e] - tgt[e];
    a += (double)(d * d);
    b += 1.0;
  }
  a = warp_sum_d(a); b = warp_sum_d(b);
  if (lane == 0) { s_a[warp] = a; s_b[warp] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0.0, y = 0.0;
    for (int w = 0; w < 8; ++w) { x += s_a[w]; y += s_b[w]; }
    if (y > 0.0) { atomicAdd(acc, x); atomicAdd(acc + 1, y); }
  }
}

// out[0] = MSE, out[1] = CE, out[2] = Lovasz (mean over present classes), out[3] = EW sum, out[4] = GLS (task_num = 2):
// sqrt(MSE * (CE + Lovasz)).  A missing term (weight 0 / has_* = 0) contributes 0.0 exactly like Criteria.__call__.
__global__ void criteria_finalize_kernel(const double* __restrict__ acc_ce, const double* __restrict__ acc_mse, const double* __restrict__ lov,
                                         const int32_t* __restrict__ cls_count, int C, float w_mse, float w_ce, float w_lov, int has_mse,
                                         int has_ce, int has_lov, float* __restrict__ out) {
  if (threadIdx.x || blockIdx.x) return;
  float mse = 0.f, ce = 0.f, lv = 0.f;
  if (has_mse) mse = (float)(acc_mse[0] / acc_mse[1]) * w_mse;          // mean over an empty set = nan, like torch
  if (has_ce) ce = (float)(acc_ce[0] / acc_ce[1]) * w_ce;
  if (has_lov) {
    double s = 0.0;
    int present = 0;
    for (int c = 0; c < C; ++c)
      if (cls_count[c] > 0) { s += lov[c]; ++present; }
    lv = present ? (float)(s / present) * w_lov : 0.f;
  }
  out[0] = mse; out[1] = ce; out[2] = lv;
  out[3] = mse + ce + lv;
  out[4] = sqrtf(mse * (ce + lv));
}

// coef[0] = d loss / d CE-term per valid point = s_seg * w_ce / n_valid, coef[1] = s_seg * w_lov / #present classes,
// coef[2] = s_mse * w_mse * 2 / #summed elements, with (s_seg, s_mse) = (1, 1) for the EW sum and
// (MSE / 2L, (CE + Lovasz) / 2L) for the GLS loss L = sqrt(MSE * (CE + Lovasz))  (builder.py:37-49)
__global__ void criteria_coef_kernel(const double* __restrict__ acc_ce, const double* __restrict__ acc_mse, const int32_t* __restrict__ cls_count,
                                     int C, float w_mse, float w_ce, float w_lov, int gls, const float* __restrict__ out5,
                                     float* __restrict__ coef) {
  if (threadIdx.x || blockIdx.x) return;
  int present = 0;
  for (int c = 0; c < C; ++c) present += cls_count[c] > 0;
  float s_seg = 1.f, s_mse = 1.f;
  if (gls) { const float L = out5[4]; s_seg = out5[0] / (2.f * L); s_mse = (out5[1] + out5[2]) / (2.f * L); }
  coef[0] = acc_ce[1] > 0.0 ? s_seg * w_ce / (float)acc_ce[1] : 0.f;
  coef[1] = present ? s_seg * w_lov / (float)present : 0.f;
  coef[2] = acc_mse[1] > 0.0 ? s_mse * w_mse * 2.f / (float)acc_mse[1] : 0.f;
}

// grad[i, j] = coef_ce (p_ij - [j == t_i]) + coef_lov p_ij (G_ij - sum_c G_ic p_ic) on valid points, 0 elsewhere: CE backward plus the
// Lovasz gradient pulled through the softmax Jacobian; one warp per point
__global__ void __launch_bounds__(256) seg_grad_kernel(const float* __restrict__ prob, const float* __restrict__ gbuf, const int64_t* __restrict__ target,
                                                       int64_t n, int C, int64_t ignore, const float* __restrict__ coef, int has_ce, int has_lov,
                                                       float* __restrict__ grad) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + warp;
  if (i >= n) return;
  const int64_t t = target[i];
  const bool valid = t != ignore && t >= 0 && t < C;
  float dot = 0.f;
  if (valid && has_lov)
    for (int c = lane; c < C; c += 32) dot += gbuf[i * C + c] * prob[i * C + c];
  dot = warp_sum(dot);
  const float a_ce = has_ce ? coef[0] : 0.f, a_lov = has_lov ? coef[1] : 0.f;
  for (int c = lane; c < C; c += 32) {
    float g = 0.f;
    if (valid) {
      const float p = prob[i * C + c];
      g = a_ce * (p - (c == t ? 1.f : 0.f));
      if (has_lov) g += a_lov * p * (gbuf[i * C + c] - dot);
    }
    grad[i * C + c] = g;
  }
}

__global__ void mse_grad_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, const int64_t* __restrict__ label, int64_t n, int C,
                                int64_t ignore, int use_ignore, const float* __restrict__ coef, float* __restrict__ grad) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * C) return;
  const bool skip = use_ignore && label && label[e / C] == ignore;
  grad[e] = skip ? 0.f : coef[2] * (pred[e] - tgt[e]);
}

// AdamW (torch.optim.AdamW, decoupled weight decay) over a table of chunks: entry k = (param, grad, exp_avg, exp_avg_sq, numel) of one
// chunk (<= 64K elements) of one tensor -- a multi-tensor apply: ONE launch updates every parameter of a group; one CTA per chunk
struct AdamEntry { float* p; const float* g; float* m; float* v; long long n; };
__global__ void __launch_bounds__(256) adamw_kernel(const AdamEntry* __restrict__ tab, float lr, float beta1, float beta2, float eps, float wd,
                                                    float bc1, float bc2_sqrt) {
  const AdamEntry e = tab[blockIdx.x];
  const float step = lr / bc1;
  for (long long i = threadIdx.x; i < e.n; i += blockDim.x) {
    const float g = e.g[i];
    float p = e.p[i] * (1.f - lr * wd);
    const float m = beta1 * e.m[i] + (1.f - beta1) * g;
    const float v = beta2 * e.v[i] + (1.f - beta2) * g * g;
    e.m[i] = m; e.v[i] = v;
    p -= step * (m / (sqrtf(v) / bc2_sqrt + eps));
    e.p[i] = p;
  }
}

// out = a[b] * x0 + c[b] * noise per row (b = batch[row]) : q(x_t | x_0), default.py:216-222
__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise, const int32_t* __restrict__ batch,
                                const float* __restrict__ sa, const float* __restrict__ sb, int64_t n, int C, float* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * C) return;
  const int b = batch ? batch[e / C] : 0;
  out[e] = __fadd_rn(__fmul_rn(sa[b], x0[e]), __fmul_rn(sb[b], noise[e]));
}

// one DDIM update with a timestep shared by all rows (default.py:192-214):
//   target "noise": x0 = (x_t - s1m * pred) / sa ;  eps = pred          target "x0": x0 = pred ; eps = (x_t - sa * x0) / s1m
//   last (t == 0): out = x0          else: out = sa_prev * x0 + s1m_prev * eps
__global__ void ddim_step_kernel(const float* __restrict__ xt, const float* __restrict__ pred, int64_t total, float sa, float s1m,
                                 float sa_prev, float s1m_prev, int target_x0, int last, float* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  float x0, eps;
  if (!target_x0) {
    x0 = __fdiv_rn(__fsub_rn(xt[e], __fmul_rn(s1m, pred[e])), sa);
    eps = pred[e];
  } else {
    x0 = pred[e];
    eps = __fdiv_rn(__fsub_rn(xt[e], __fmul_rn(sa, x0)), s1m);
  }
  out[e] = last ? x0 : __fadd_rn(__fmul_rn(sa_prev, x0), __fmul_rn(s1m_prev, eps));
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, float scale, int64_t total) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < total) y[e] = (y[e] + a * x[e]) * scale;
}

}  // namespace ls

CDSEG_API size_t cdseg_criteria_workspace_bytes(int64_t n, int C) {
  const size_t head = (((size_t)(4 + C) * 8 + (size_t)C * 4 + 16) + 255) & ~(size_t)255;
  const size_t prob = ((size_t)n * C * 4 + 255) & ~(size_t)255;
  const size_t keys = (size_t)n * C * 8, ord = (size_t)n * C * 4;
  return head + prob + keys + 2 * ord + cdseg_argsort_workspace_bytes(C, n) + 256;
}

// see include/cdseg_b200.h
static int criteria_impl(const float* n_pred, const int64_t* n_target, int64_t n, int C, int64_t ignore, const float* c_pred,
                         const float* c_target, int Cc, int mse_use_ignore, float w_mse, float w_ce, float w_lov, int has_mse,
                         int has_ce, int has_lov, float* out5, int gls, float* grad_n, float* grad_c, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (n < 0 || C <= 0 || C > 4096 || !n_pred || !n_target || !out5 || !workspace) return CDSEG_EINVAL;
  if (has_mse && (!c_pred || !c_target || Cc <= 0)) return CDSEG_EINVAL;
  if (workspace_bytes < cdseg_criteria_workspace_bytes(n, C)) return CDSEG_ENOSPC;
  cudaStream_t st = (cudaStream_t)stream;
  char* p = (char*)workspace;
  double* acc = (double*)p;                       // [0..1] CE, [2..3] MSE
  double* lov = acc + 4;                          // [C]
  int32_t* cls = (int32_t*)(lov + C);             // [C]
  const size_t head = (((size_t)(4 + C) * 8 + (size_t)C * 4 + 16) + 255) & ~(size_t)255;
  cudaMemsetAsync(p, 0, head, st);
  p += head;
  float* prob = (float*)p; p += ((size_t)n * C * 4 + 255) & ~(size_t)255;
  int64_t* keys = (int64_t*)p; p += (size_t)n * C * 8;
  float* gbuf = (grad_n && has_lov) ? (float*)keys : nullptr;          // the key rows are dead once the sort has run: reuse them for d loss_c / d p
  float* coef = (float*)(cls + C);                                      // 3 floats behind the per-class counts (inside the zeroed header)
  int32_t* order = (int32_t*)p; p += (size_t)n * C * 4;
  int32_t* inverse = (int32_t*)p; p += (size_t)n * C * 4;
  const size_t sort_ws = (size_t)((char*)workspace + workspace_bytes - p);
  if (n > 0) {
    const int blocks = (int)(((n + 7) / 8) < 2368 ? ((n + 7) / 8) : 2368);
    ls::ce_softmax_kernel<<<blocks, 256, 0, st>>>(n_pred, n_target, n, C, ignore, (has_lov || grad_n) ? prob : nullptr, acc, cls);
    CDSEG_COUNT_LAUNCH(1);
    if (has_lov) {
      dim3 g((unsigned)cdseg_div_up(n, 256), (unsigned)C);
      ls::lovasz_keys_kernel<<<g, 256, 0, st>>>(prob, n_target, n, C, ignore, keys);
      CDSEG_COUNT_LAUNCH(1);
      int s = cdseg_argsort_rows(keys, C, n, 31, order, inverse, p, sort_ws, stream);
      if (s != CDSEG_OK) return s;
      if (gbuf) cudaMemsetAsync(gbuf, 0, (size_t)n * C * 4, st);         // absent classes and ignored points keep a zero gradient
      ls::lovasz_scan_kernel<<<C, ls::LV_THREADS, 0, st>>>(prob, n_target, order, cls, n, C, ignore, lov, gbuf);
      CDSEG_COUNT_LAUNCH(1);
    }
    if (has_mse) {
      const int64_t total = n * Cc;
      const int mb = (int)(((total + 255) / 256) < 1184 ? ((total + 255) / 256) : 1184);
      ls::mse_kernel<<<mb, 256, 0, st>>>(c_pred, c_target, n_target, n, Cc, ignore, mse_use_ignore, acc + 2);
      CDSEG_COUNT_LAUNCH(1);
    }
  }
  ls::criteria_finalize_kernel<<<1, 32, 0, st>>>(acc, acc + 2, lov, cls, C, w_mse, w_ce, w_lov, has_mse, has_ce, has_lov, out5);
  CDSEG_COUNT_LAUNCH(1);
  if ((grad_n || grad_c) && n > 0) {
    ls::criteria_coef_kernel<<<1, 32, 0, st>>>(acc, acc + 2, cls, C, w_mse, w_ce, w_lov, gls, out5, coef);
    CDSEG_COUNT_LAUNCH(1);
    if (grad_n) {
      ls::seg_grad_kernel<<<cdseg_div_up(n, 8), 256, 0, st>>>(prob, gbuf, n_target, n, C, ignore, coef, has_ce, has_lov && gbuf, grad_n);
      CDSEG_COUNT_LAUNCH(1);
    }
    if (grad_c && has_mse) {
      ls::mse_grad_kernel<<<cdseg_div_up(n * Cc, 256), 256, 0, st>>>(c_pred, c_target, n_target, n, Cc, ignore, mse_use_ignore, coef, grad_c);
      CDSEG_COUNT_LAUNCH(1);
    }
  }
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

CDSEG_API int cdseg_criteria(const float* n_pred, const int64_t* n_target, int64_t n, int C, int64_t ignore, const float* c_pred,
                             const float* c_target, int Cc, int mse_use_ignore, float w_mse, float w_ce, float w_lov, int has_mse,
                             int has_ce, int has_lov, float* out5, void* workspace, size_t workspace_bytes, void* stream) {
  return criteria_impl(n_pred, n_target, n, C, ignore, c_pred, c_target, Cc, mse_use_ignore, w_mse, w_ce, w_lov, has_mse, has_ce, has_lov,
                       out5, 0, nullptr, nullptr, workspace, workspace_bytes, stream);
}

// forward values + gradients of the selected combination w.r.t. the two network outputs (see include/cdseg_b200.h)
CDSEG_API int cdseg_criteria_grad(const float* n_pred, const int64_t* n_target, int64_t n, int C, int64_t ignore, const float* c_pred,
                                  const float* c_target, int Cc, int mse_use_ignore, float w_mse, float w_ce, float w_lov, int has_mse,
                                  int has_ce, int has_lov, int gls, float* out5, float* grad_n_pred, float* grad_c_pred, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (!grad_n_pred) return CDSEG_EINVAL;
  return criteria_impl(n_pred, n_target, n, C, ignore, c_pred, c_target, Cc, mse_use_ignore, w_mse, w_ce, w_lov, has_mse, has_ce, has_lov,
                       out5, gls, grad_n_pred, grad_c_pred, workspace, workspace_bytes, stream);
}

CDSEG_API int cdseg_adamw_step(const void* table, int n_chunks, float lr, float beta1, float beta2, float eps, float weight_decay,
                               int step, void* stream) {
  if (!table || n_chunks < 0 || step < 1) return CDSEG_EINVAL;
  if (n_chunks == 0) return CDSEG_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  ls::adamw_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>((const ls::AdamEntry*)table, lr, beta1, beta2, eps, weight_decay, (float)bc1,
                                                               (float)sqrt(bc2));
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

CDSEG_API int cdseg_q_sample(const float* x0, const float* noise, const int32_t* batch, const float* sqrt_ab, const float* sqrt_1mab,
                             int64_t n, int C, float* out, void* stream) {
  if (n < 0 || C <= 0 || !x0 || !noise || !sqrt_ab || !sqrt_1mab || !out) return CDSEG_EINVAL;
  if (n == 0) return CDSEG_OK;
  ls::q_sample_kernel<<<cdseg_div_up(n * C, 256), 256, 0, (cudaStream_t)stream>>>(x0, noise, batch, sqrt_ab, sqrt_1mab, n, C, out);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

CDSEG_API int cdseg_ddim_step(const float* x_t, const float* pred, int64_t total, float sqrt_ab, float sqrt_1mab, float sqrt_ab_prev,
                              float sqrt_1mab_prev, int target_is_x0, int last, float* out, void* stream) {
  if (total < 0 || !x_t || !pred || !out) return CDSEG_EINVAL;
  if (total == 0) return CDSEG_OK;
  ls::ddim_step_kernel<<<cdseg_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(x_t, pred, total, sqrt_ab, sqrt_1mab, sqrt_ab_prev,
                                                                                  sqrt_1mab_prev, target_is_x0, last, out);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

CDSEG_API int cdseg_axpy_scale(float* y, const float* x, float a, float scale, int64_t total, void* stream) {
  if (total < 0 || !y || !x) return CDSEG_EINVAL;
  if (total == 0) return CDSEG_OK;
  ls::axpy_kernel<<<cdseg_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(y, x, a, scale, total);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}
