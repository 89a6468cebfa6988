// The consumers either side of the pseudo-label path, as streaming sm_100a kernels.
//
//   seg_loss              fg/bg-balanced cross-entropy on the pseudo-label map        utils/seg_helper.py:800-813
//   seg_refine_by_label   class-masked, temperature-sharpened softmax of the teacher   utils/seg_helper.py:553-568
//   cam_loss              multi-label soft-margin loss, CAM vs. refined segmentation   utils/seg_helper.py:593-602
//
// The per-pixel softmax kernels come in two forms: for the VOC class count (C = 21) the logits of a thread's
// pixels stay in registers (read once, all loads in flight); for any other C they are streamed twice with an
// online max/sum.  A thread owns V = 4 consecutive pixels of a plane when H*W % 4 == 0, else one.
#include <math.h>

#include "common.cuh"

namespace cosa {

template <int V>
__device__ __forceinline__ void load_px(const float *p, float (&v)[V]) {
  if constexpr (V == 4) {
    const float4 t = ldg_stream4(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    v[0] = __ldg(p);
  }
}
template <int V>
__device__ __forceinline__ void store_px(float *p, const float (&v)[V]) {
  if constexpr (V == 4) stg_stream4(p, make_float4(v[0], v[1], v[2], v[3]));
  else p[0] = v[0];
}

struct SegLossStats {       // device-resident between forward and backward
  double sum_bg, sum_fg;    // sum of cross-entropies over the background / foreground pixels
  unsigned long long n_bg, n_fg;
};

// log-sum-exp and softmax statistics of the V pixels of a thread; REG keeps the logits in l[][].
template <int CT, int V>
__device__ __forceinline__ void softmax_stats(const float *lg, size_t HW, int C, float (&l)[CT > 0 ? CT : 1][V],
                                              float (&mx)[V], float (&den)[V]) {
#pragma unroll
  for (int i = 0; i < V; ++i) { mx[i] = -INFINITY; den[i] = 0.0f; }
  if constexpr (CT > 0) {
#pragma unroll
    for (int c = 0; c < CT; ++c) load_px<V>(lg + (size_t)c * HW, l[c]);
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
      for (int i = 0; i < V; ++i) mx[i] = fmaxf(mx[i], l[c][i]);
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
      for (int i = 0; i < V; ++i) den[i] += expf(l[c][i] - mx[i]);
  } else {
    for (int c = 0; c < C; ++c) {
      float t[V];
      load_px<V>(lg + (size_t)c * HW, t);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float d = t[i] - mx[i];
        const float e = expf(-fabsf(d));
        den[i] = d > 0.0f ? fmaf(den[i], e, 1.0f) : den[i] + e;
        mx[i] = fmaxf(mx[i], t[i]);
      }
    }
  }
}

// ---- seg_loss ------------------------------------------------------------------------------------------------
template <int CT, int V>
__global__ void __launch_bounds__(128) seg_ce_forward_kernel(const float *__restrict__ logit,
                                                             const float *__restrict__ label, int ignore,
                                                             SegLossStats *stats, int B, int C, long long HW) {
  const long long per = HW / V, t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float s_bg = 0.0f, s_fg = 0.0f;
  int c_bg = 0, c_fg = 0;
  if (t < (long long)B * per) {
    const int b = (int)(t / per);
    const size_t p = (size_t)(t % per) * V;
    const float *lg = logit + (size_t)b * C * HW + p;
    float l[CT > 0 ? CT : 1][V], mx[V], den[V], labf[V];
    softmax_stats<CT, V>(lg, (size_t)HW, C, l, mx, den);
    load_px<V>(label + (size_t)b * HW + p, labf);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int lab = (int)labf[i];                                   // .long() of an integer-valued float map
      if (lab == ignore || lab < 0 || lab >= C) continue;
      float picked;
      if constexpr (CT > 0) {
        picked = 0.0f;
#pragma unroll
        for (int c = 0; c < CT; ++c) picked = c == lab ? l[c][i] : picked;
      } else {
        picked = __ldg(lg + (size_t)lab * HW + i);
      }
      const float ce = (mx[i] + logf(den[i])) - picked;               // -log_softmax at the label
      if (lab == 0) { s_bg += ce; ++c_bg; } else { s_fg += ce; ++c_fg; }
    }
  }
  __shared__ float sh_f[2][4];
  __shared__ int sh_i[2][4];
  s_bg = warp_sum(s_bg); s_fg = warp_sum(s_fg);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { c_bg += __shfl_xor_sync(0xffffffffu, c_bg, o); c_fg += __shfl_xor_sync(0xffffffffu, c_fg, o); }
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sh_f[0][wid] = s_bg; sh_f[1][wid] = s_fg; sh_i[0][wid] = c_bg; sh_i[1][wid] = c_fg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, f = 0.0;
    unsigned long long nb = 0, nf = 0;
    for (int k = 0; k < 4; ++k) { a += sh_f[0][k]; f += sh_f[1][k]; nb += sh_i[0][k]; nf += sh_i[1][k]; }
    if (nb) { atomicAdd(&stats->sum_bg, a); atomicAdd(&stats->n_bg, nb); }
    if (nf) { atomicAdd(&stats->sum_fg, f); atomicAdd(&stats->n_fg, nf); }
  }
}

__global__ void seg_loss_finalize_kernel(const SegLossStats *stats, float fg_alpha, float *loss_out) {
  // sum / (count + 1e-6) in fp32, then the convex combination (seg_helper.py:806-813)
  const float bg = __fdiv_rn((float)stats->sum_bg, __fadd_rn((float)stats->n_bg, 1e-6f));
  const float fg = __fdiv_rn((float)stats->sum_fg, __fadd_rn((float)stats->n_fg, 1e-6f));
  loss_out[0] = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, fg_alpha), bg), __fmul_rn(fg_alpha, fg));
}

template <int CT, int V>
__global__ void __launch_bounds__(128) seg_ce_backward_kernel(const float *__restrict__ logit,
                                                              const float *__restrict__ label, int ignore,
                                                              const SegLossStats *__restrict__ stats,
                                                              const float *__restrict__ grad_out, float fg_alpha,
                                                              float *__restrict__ grad_logit, int B, int C,
                                                              long long HW) {
  const long long per = HW / V, t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)B * per) return;
  const int b = (int)(t / per);
  const size_t p = (size_t)(t % per) * V;
  const float *lg = logit + (size_t)b * C * HW + p;
  float *out = grad_logit + (size_t)b * C * HW + p;
  const float g = __ldg(grad_out);
  const float k_bg = g * (1.0f - fg_alpha) / ((float)stats->n_bg + 1e-6f);
  const float k_fg = g * fg_alpha / ((float)stats->n_fg + 1e-6f);
  float l[CT > 0 ? CT : 1][V], mx[V], den[V], labf[V], coef[V];
  int lab[V];
  softmax_stats<CT, V>(lg, (size_t)HW, C, l, mx, den);
  load_px<V>(label + (size_t)b * HW + p, labf);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    lab[i] = (int)labf[i];
    const bool valid = lab[i] != ignore && lab[i] >= 0 && lab[i] < C;
    coef[i] = !valid ? 0.0f : (lab[i] == 0 ? k_bg : k_fg);
    den[i] = 1.0f / den[i];
  }
  const int Cn = CT > 0 ? CT : C;
#pragma unroll
  for (int c = 0; c < Cn; ++c) {
    float x[V], o[V];
    if constexpr (CT > 0) {
#pragma unroll
      for (int i = 0; i < V; ++i) x[i] = l[c][i];
    } else {
      if constexpr (V == 4) { const float4 t4 = ldg4c(lg + (size_t)c * HW); x[0] = t4.x; x[1] = t4.y; x[2] = t4.z; x[3] = t4.w; }
      else x[0] = __ldg(lg + (size_t)c * HW);
    }
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] = coef[i] * (expf(x[i] - mx[i]) * den[i] - (c == lab[i] ? 1.0f : 0.0f));
    store_px<V>(out + (size_t)c * HW, o);
  }
}

// ---- seg_refine_by_label -----------------------------------------------------------------------------------
// after_softmax == 0: channels of absent classes are set to -1e5 BEFORE the division by the temperature and the
// softmax (:564-566); after_softmax == 1: plain softmax(seg / T) multiplied by the 0/1 class vector (:559-561).
// The background channel (index 0) is always kept (:556-557).
template <int CT, int V>
__global__ void __launch_bounds__(128) seg_refine_kernel(const float *__restrict__ seg,
                                                         const float *__restrict__ cls_label, float temp,
                                                         int after_softmax, float *__restrict__ out, int B, int C,
                                                         long long HW) {
  const long long per = HW / V, t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)B * per) return;
  const int b = (int)(t / per);
  const size_t p = (size_t)(t % per) * V;
  const float *sg = seg + (size_t)b * C * HW + p;
  float *o = out + (size_t)b * C * HW + p;
  const float *lab = cls_label + (size_t)b * (C - 1);
  auto present = [&](int c) { return c == 0 || (long long)__ldg(lab + c - 1) != 0; };
  // x / temp as q + (x - temp q) * rcp with q = x * rcp, rcp = RN(1 / temp): the quotient of the division routine
  // (correctly rounded but for rare last-bit cases and magnitudes beyond 1e30) at three issue slots instead of a
  // dozen and a branch - the kernel spends C divisions per pixel
  const float rcp = __frcp_rn(temp);
  auto value = [&](int c, float x) {     // the tensor the softmax sees, already divided by the temperature
    const float a = (!after_softmax && !present(c)) ? -1e5f : x;
    const float q = __fmul_rn(a, rcp);
    return __fmaf_rn(__fmaf_rn(-temp, q, a), rcp, q);
  };
  float mx[V], den[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { mx[i] = -INFINITY; den[i] = 0.0f; }
  if constexpr (CT > 0) {
    float l[CT][V];
#pragma unroll
    for (int c = 0; c < CT; ++c) load_px<V>(sg + (size_t)c * HW, l[c]);
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
      for (int i = 0; i < V; ++i) { l[c][i] = value(c, l[c][i]); mx[i] = fmaxf(mx[i], l[c][i]); }
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
      for (int i = 0; i < V; ++i) { l[c][i] = __expf(l[c][i] - mx[i]); den[i] += l[c][i]; }
#pragma unroll
    for (int i = 0; i < V; ++i) den[i] = 1.0f / den[i];
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      const float keep = (after_softmax && !present(c)) ? 0.0f : 1.0f;
      float r[V];
#pragma unroll
      for (int i = 0; i < V; ++i) r[i] = l[c][i] * den[i] * keep;
      store_px<V>(o + (size_t)c * HW, r);
    }
  } else {
    for (int c = 0; c < C; ++c) {
      float x[V];
      load_px<V>(sg + (size_t)c * HW, x);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float v = value(c, x[i]);
        const float d = v - mx[i];
        const float e = __expf(-fabsf(d));
        den[i] = d > 0.0f ? fmaf(den[i], e, 1.0f) : den[i] + e;
        mx[i] = fmaxf(mx[i], v);
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) den[i] = 1.0f / den[i];
    for (int c = 0; c < C; ++c) {
      const float keep = (after_softmax && !present(c)) ? 0.0f : 1.0f;
      float x[V], r[V];
      if constexpr (V == 4) { const float4 t4 = ldg4c(sg + (size_t)c * HW); x[0] = t4.x; x[1] = t4.y; x[2] = t4.z; x[3] = t4.w; }
      else x[0] = __ldg(sg + (size_t)c * HW);
#pragma unroll
      for (int i = 0; i < V; ++i) r[i] = __expf(value(c, x[i]) - mx[i]) * den[i] * keep;
      store_px<V>(o + (size_t)c * HW, r);
    }
  }
}

// ---- cam_loss ------------------------------------------------------------------------------------------------
// target[b,c,y,x] = bilinear(seg_ps[b,c+1] -> (H,W)), align_corners=False; x = relu(cam) (optional);
// loss = mean_{b,y,x} mean_c -( t log sigmoid(x) + (1-t) log sigmoid(-x) )   (F.multilabel_soft_margin_loss)
__global__ void __launch_bounds__(256) cam_loss_forward_kernel(const float *__restrict__ cam,
                                                               const float *__restrict__ seg_ps, int is_relu,
                                                               float *__restrict__ target, double *acc, int B, int C,
                                                               int H, int W, int Hs, int Ws) {
  const long long total = (long long)B * C * H * W;
  float local = 0.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long plane = i / ((long long)W * H);
    const int b = (int)(plane / C), c = (int)(plane % C);
    const float *s = seg_ps + ((size_t)b * (C + 1) + c + 1) * Hs * Ws;
    float tv;
    if (Hs == H && Ws == W) {
      tv = __ldg(s + (size_t)y * W + x);
    } else {
      const Tap ty = tap_half_pixel(y, (float)Hs / (float)H, Hs), tx = tap_half_pixel(x, (float)Ws / (float)W, Ws);
      const float a00 = __ldg(s + (size_t)ty.i0 * Ws + tx.i0), a01 = __ldg(s + (size_t)ty.i0 * Ws + tx.i1);
      const float a10 = __ldg(s + (size_t)ty.i1 * Ws + tx.i0), a11 = __ldg(s + (size_t)ty.i1 * Ws + tx.i1);
      tv = (Hs > H || Ws > W) ? bilerp_down(ty, tx, a00, a01, a10, a11) : bilerp_up(ty, tx, a00, a01, a10, a11);
    }
    target[i] = tv;
    float v = __ldg(cam + i);
    if (is_relu) v = fmaxf(v, 0.0f);
    const float sp = log1pf(expf(-fabsf(v)));              // log sigmoid(v) = min(v,0) - sp, log sigmoid(-v) = min(-v,0) - sp
    local += -(tv * (fminf(v, 0.0f) - sp) + (1.0f - tv) * (fminf(-v, 0.0f) - sp));
  }
  __shared__ float s_part[8];
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += s_part[k];
    atomicAdd(acc, t);
  }
}

__global__ void cam_loss_finalize_kernel(const double *acc, double count, float *loss_out) {
  loss_out[0] = (float)(acc[0] / count);
}

__global__ void __launch_bounds__(256) cam_loss_backward_kernel(const float *__restrict__ cam,
                                                                const float *__restrict__ target,
                                                                const float *__restrict__ grad_out, int is_relu,
                                                                float *__restrict__ grad_cam, long long total) {
  const float k = __ldg(grad_out) / (float)total;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float raw = __ldg(cam + i);
    const float v = is_relu ? fmaxf(raw, 0.0f) : raw;
    const float sig = 1.0f / (1.0f + expf(-v));
    grad_cam[i] = (is_relu && raw <= 0.0f) ? 0.0f : k * (sig - __ldg(target + i));
  }
}

}  // namespace cosa

using namespace cosa;

// C = 21 with whole quads -> register-resident form; other C -> streamed twice; H*W % 4 != 0 -> one pixel per thread
#define COSA_DISPATCH_CV(KERNEL, C, HW, total_px, ...)                                                 \
  do {                                                                                                 \
    if ((HW) % 4 == 0) {                                                                               \
      const unsigned nb_ = (unsigned)ceil_div_ll((total_px) / 4, 128);                                 \
      if ((C) == 21) {                                                                                 \
        auto k_ = KERNEL<21, 4>;                                                                       \
        COSA_LAUNCH_T(#KERNEL, k_, nb_, 128, 0, s, __VA_ARGS__);                                       \
      } else {                                                                                         \
        auto k_ = KERNEL<0, 4>;                                                                        \
        COSA_LAUNCH_T(#KERNEL, k_, nb_, 128, 0, s, __VA_ARGS__);                                       \
      }                                                                                                \
    } else {                                                                                           \
      const unsigned nb_ = (unsigned)ceil_div_ll((total_px), 128);                                     \
      auto k_ = KERNEL<0, 1>;                                                                          \
      COSA_LAUNCH_T(#KERNEL, k_, nb_, 128, 0, s, __VA_ARGS__);                                         \
    }                                                                                                  \
  } while (0)

extern "C" size_t cosa_seg_loss_stats_bytes(void) { return sizeof(SegLossStats); }

extern "C" int cosa_seg_loss_forward(const float *seg_pred, const float *mask_label, float fg_alpha, int ignore_index,
                                     float *loss_out, void *stats, int B, int C, int H, int W, void *stream) {
  if (!seg_pred || !mask_label || !loss_out || !stats || B < 1 || C < 1 || H < 1 || W < 1) return COSA_E_ARG;
  if (!(fg_alpha >= 0.0f && fg_alpha <= 1.0f)) return COSA_E_ARG;   // the reference asserts this (:803)
  cudaStream_t s = (cudaStream_t)stream;
  const long long HW = (long long)H * W, px = (long long)B * HW;
  SegLossStats *st = (SegLossStats *)stats;
  COSA_CUDA(cudaMemsetAsync(st, 0, sizeof(SegLossStats), s));
  COSA_DISPATCH_CV(seg_ce_forward_kernel, C, HW, px, seg_pred, mask_label, ignore_index, st, B, C, HW);
  COSA_LAUNCH(seg_loss_finalize_kernel, 1, 1, 0, s, st, fg_alpha, loss_out);
  return 0;
}

extern "C" int cosa_seg_loss_backward(const float *seg_pred, const float *mask_label, const void *stats,
                                      const float *grad_out, float fg_alpha, int ignore_index, float *grad_pred, int B,
                                      int C, int H, int W, void *stream) {
  if (!seg_pred || !mask_label || !stats || !grad_out || !grad_pred || B < 1 || C < 1 || H < 1 || W < 1)
    return COSA_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const long long HW = (long long)H * W, px = (long long)B * HW;
  const SegLossStats *st = (const SegLossStats *)stats;
  COSA_DISPATCH_CV(seg_ce_backward_kernel, C, HW, px, seg_pred, mask_label, ignore_index, st, grad_out, fg_alpha,
                   grad_pred, B, C, HW);
  return 0;
}

extern "C" int cosa_seg_refine_by_label(const float *seg, const float *cls_label, float softmaxtemp, int after_softmax,
                                        float *out, int B, int C, int H, int W, void *stream) {
  if (!seg || !cls_label || !out || B < 1 || C < 2 || H < 1 || W < 1 || softmaxtemp == 0.0f) return COSA_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const long long HW = (long long)H * W, px = (long long)B * HW;
  COSA_DISPATCH_CV(seg_refine_kernel, C, HW, px, seg, cls_label, softmaxtemp, after_softmax, out, B, C, HW);
  return 0;
}

extern "C" int cosa_cam_loss_forward(const float *cam, const float *seg_ps, int is_relu, float *loss_out,
                                     float *target_ws, double *acc_ws, int B, int C, int H, int W, int Hs, int Ws,
                                     void *stream) {
  if (!cam || !seg_ps || !loss_out || !target_ws || !acc_ws || B < 1 || C < 1 || H < 1 || W < 1 || Hs < 1 || Ws < 1)
    return COSA_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const long long total = (long long)B * C * H * W;
  COSA_CUDA(cudaMemsetAsync(acc_ws, 0, sizeof(double), s));
  const int blocks = (int)max(1LL, min((long long)sm_count() * 8, ceil_div_ll(total, 256)));
  COSA_LAUNCH(cam_loss_forward_kernel, blocks, 256, 0, s, cam, seg_ps, is_relu, target_ws, acc_ws, B, C, H, W, Hs, Ws);
  COSA_LAUNCH(cam_loss_finalize_kernel, 1, 1, 0, s, acc_ws, (double)total, loss_out);
  return 0;
}

extern "C" int cosa_cam_loss_backward(const float *cam, const float *target_ws, const float *grad_out, int is_relu,
                                      float *grad_cam, int B, int C, int H, int W, void *stream) {
  if (!cam || !target_ws || !grad_out || !grad_cam || B < 1 || C < 1 || H < 1 || W < 1) return COSA_E_ARG;
  const long long total = (long long)B * C * H * W;
  const int blocks = (int)max(1LL, min((long long)sm_count() * 8, ceil_div_ll(total, 256)));
  COSA_LAUNCH(cam_loss_backward_kernel, blocks, 256, 0, (cudaStream_t)stream, cam, target_ws, grad_out, is_relu,
              grad_cam, total);
  return 0;
}
