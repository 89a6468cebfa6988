// Dense-CRF energy loss on the GPU lattice.
//
// Reference: utils/seg_helper.py:864-903 DenseEnergyLossFunction (dup utils/rrm_utils.py:352-391),
//            utils/seg_helper.py:191-208 DenseEnergyLoss.forward, :210-230 get_energy_loss.
//
//   energy_gate_kernel      Gate = clamp_min(ROI - max_k S, 0), Gate[unlabel] = 1; S <- S * ROI        :870-880
//   (lattice)               AS = BilateralFilter(S * ROI); AS <- AS * Gate; loss = -<S, AS>/N            :887-893
//   energy_grad_kernel      grad_S = -2 * g * AS / N * ROI                                               :898-903
//   energy_prepare_kernel   get_energy_loss + DenseEnergyLoss.forward fused for the exact 2:1 case:
//                           de-normalise + nearest image, ROI from the boxes, unlabel from the label map,
//                           bilinear(softmax(logit)), gate                                               :199-229
//   energy_logit_grad_kernel  the matching backward: d loss / d logit through the bilinear 2:1 reduction
//                           and the softmax, without materialising the probabilities.
#include <math.h>

#include <stdlib.h>

#include "common.cuh"
#include "lattice.cuh"

namespace cosa {

__global__ void __launch_bounds__(256) energy_gate_kernel(const float *__restrict__ segs,
                                                          const float *__restrict__ rois,
                                                          const unsigned char *__restrict__ unlabel,
                                                          float *__restrict__ s_roi, float *__restrict__ gate, int K,
                                                          int n, long long P) {
  for (long long gp = blockIdx.x * (long long)blockDim.x + threadIdx.x; gp < P;
       gp += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(gp / n), p = (int)(gp % n);
    const float roi = __ldg(rois + gp);
    const float *src = segs + (size_t)b * K * n + p;
    float *dst = s_roi + (size_t)b * K * n + p;
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) {
      const float s = __ldg(src + (size_t)k * n);
      mx = fmaxf(mx, s);
      dst[(size_t)k * n] = __fmul_rn(s, roi);
    }
    float g = __fsub_rn(roi, mx);
    if (unlabel[gp]) g = 1.0f;
    if (g < 0.0f) g = 0.0f;
    gate[gp] = g;
  }
}

__global__ void energy_loss_finalize_kernel(const double *__restrict__ acc, float *__restrict__ loss_out, int N,
                                            float weight, int apply_weight) {
  // np.dot in float32, negated, divided by N (seg_helper.py:890-893); the layer multiplies by its weight (:207)
  float l = __fdiv_rn(-(float)(*acc), (float)N);
  if (apply_weight) l = __fmul_rn(weight, l);
  loss_out[0] = l;
}

__global__ void __launch_bounds__(256) energy_grad_kernel(const float *__restrict__ as_saved,
                                                          const float *__restrict__ rois,
                                                          const float *__restrict__ grad_out,
                                                          float *__restrict__ grad_seg, int K, int n, long long P,
                                                          int N) {
  const float c = __fmul_rn(-2.0f, __ldg(grad_out));
  const float fn = (float)N;
  for (long long gp = blockIdx.x * (long long)blockDim.x + threadIdx.x; gp < P;
       gp += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(gp / n), p = (int)(gp % n);
    const float roi = __ldg(rois + gp);
    const size_t base = (size_t)b * K * n + p;
    for (int k = 0; k < K; ++k) {
      const size_t i = base + (size_t)k * n;
      grad_seg[i] = __fmul_rn(__fdiv_rn(__fmul_rn(c, __ldg(as_saved + i)), fn), roi);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused get_energy_loss front end, exact 2:1 reduction (H, W even; scale_factor 0.5).
// One thread per half-resolution pixel.  The four source pixels of the bilinear reduction are softmax-ed on
// the fly (online max/sum), so the [B,C,H,W] probability tensor is never written.
// ------------------------------------------------------------------------------------------------
struct Affine3 {
  float mean[3], std[3];
};

__device__ __forceinline__ void online_softmax_stats(const float *__restrict__ logit, size_t HW, int C, float &mx,
                                                     float &den) {
  mx = -INFINITY;
  den = 0.0f;
  for (int c = 0; c < C; ++c) {
    const float v = __ldg(logit + (size_t)c * HW);
    if (v > mx) {
      den = den * expf(mx - v) + 1.0f;
      mx = v;
    } else {
      den += expf(v - mx);
    }
  }
}

__global__ void __launch_bounds__(256) energy_prepare_kernel(const float *__restrict__ simg,
                                                             const float *__restrict__ logit,
                                                             const float *__restrict__ label,
                                                             const int *__restrict__ boxes, Affine3 aff,
                                                             float *__restrict__ img_half, float *__restrict__ s_roi,
                                                             float *__restrict__ gate, float *__restrict__ roi_out,
                                                             int C, int H, int W) {
  const int h = H / 2, w = W / 2;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (x >= w || y >= h) return;
  const size_t HW = (size_t)H * W, hw = (size_t)h * w;
  const size_t pix = (size_t)y * w + x;
  const size_t src = (size_t)(2 * y) * W + 2 * x;   // nearest: source index 2i (seg_helper.py:201,203,204)

  if (img_half) {   // null: the prebuilt lattice's owner has written it already
#pragma unroll
    for (int c = 0; c < 3; ++c)   // img * std + mean, two roundings like the two tensor ops (:225-227)
      img_half[((size_t)b * 3 + c) * hw + pix] =
          __fadd_rn(__fmul_rn(__ldg(simg + ((size_t)b * 3 + c) * HW + src), aff.std[c]), aff.mean[c]);
  }

  const int *box = boxes + 4 * b;
  const float roi = (2 * y >= box[0] && 2 * y < box[1] && 2 * x >= box[2] && 2 * x < box[3]) ? 1.0f : 0.0f;
  // label.type(uint8) then == 255 (:229, :205)
  const bool unl = ((int)__ldg(label + (size_t)b * HW + src) & 255) == 255;

  const float *lg = logit + (size_t)b * C * HW + src;
  float mx[4], rden[4];
  const size_t tap[4] = {0, 1, (size_t)W, (size_t)W + 1};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float den;
    online_softmax_stats(lg + tap[t], HW, C, mx[t], den);
    rden[t] = den;
  }
  float smax = -INFINITY;
  float *dst = s_roi + (size_t)b * C * hw + pix;
  for (int c = 0; c < C; ++c) {
    const float *lc = lg + (size_t)c * HW;
    const float p00 = __fdiv_rn(expf(__ldg(lc + tap[0]) - mx[0]), rden[0]);
    const float p01 = __fdiv_rn(expf(__ldg(lc + tap[1]) - mx[1]), rden[1]);
    const float p10 = __fdiv_rn(expf(__ldg(lc + tap[2]) - mx[2]), rden[2]);
    const float p11 = __fdiv_rn(expf(__ldg(lc + tap[3]) - mx[3]), rden[3]);
    // exact 2:1 bilinear, align_corners=False: 0.25 * (((p00 + p01) + p10) + p11)
    const float s = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(p00, p01), p10), p11));
    smax = fmaxf(smax, s);
    dst[(size_t)c * hw] = __fmul_rn(s, roi);
  }
  float g = __fsub_rn(roi, smax);
  if (unl) g = 1.0f;
  if (g < 0.0f) g = 0.0f;
  gate[(size_t)b * hw + pix] = g;
  roi_out[(size_t)b * hw + pix] = roi;
}

// d loss / d logit.  grad wrt the half-resolution S is gS = coef * AS * ROI with coef = -2 * g * weight / N;
// each of the four source pixels receives 0.25 * gS through the bilinear reduction, then the softmax
// backward p * (u - <p, u>) is applied per source pixel.
__global__ void __launch_bounds__(256) energy_logit_grad_kernel(const float *__restrict__ logit,
                                                                const float *__restrict__ as_saved,
                                                                const float *__restrict__ roi_half,
                                                                const float *__restrict__ grad_out, float weight,
                                                                float *__restrict__ grad_logit, int C, int H, int W,
                                                                int N) {
  const int h = H / 2, w = W / 2;
  const int X = blockIdx.x * 32 + (threadIdx.x & 31);
  const int Y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (X >= W || Y >= H) return;
  const size_t HW = (size_t)H * W, hw = (size_t)h * w;
  const size_t pix = (size_t)(Y >> 1) * w + (X >> 1);
  const float roi = __ldg(roi_half + (size_t)b * hw + pix);
  const float coef = __fdiv_rn(__fmul_rn(-2.0f, __fmul_rn(__ldg(grad_out), weight)), (float)N) * roi * 0.25f;
  const float *lg = logit + (size_t)b * C * HW + (size_t)Y * W + X;
  const float *as = as_saved + (size_t)b * C * hw + pix;
  float *out = grad_logit + (size_t)b * C * HW + (size_t)Y * W + X;
  float mx, den;
  online_softmax_stats(lg, HW, C, mx, den);
  const float rden = 1.0f / den;
  float dot = 0.0f;
  for (int c = 0; c < C; ++c) {
    const float p = expf(__ldg(lg + (size_t)c * HW) - mx) * rden;
    dot = fmaf(p, coef * __ldg(as + (size_t)c * hw), dot);
  }
  for (int c = 0; c < C; ++c) {
    const float p = expf(__ldg(lg + (size_t)c * HW) - mx) * rden;
    out[(size_t)c * HW] = p * (coef * __ldg(as + (size_t)c * hw) - dot);
  }
}

// ------------------------------------------------------------------------------------------------
// Vectorised streaming variants (W % 4 == 0): 128-bit loads, two passes over the logits (the second one hits
// L2), online max/sum in the first pass with ONE exponential per element (exp(-|l - max|) serves both the
// "new maximum" rescale and the ordinary accumulation).  Within the 1e-4 budget the fast ex2-based __expf is
// used here (relative error ~1e-6 over the softmax range); nothing on this branch decides a label.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void online_update(float l, float &mx, float &den) {
  const float d = l - mx;
  const float e = __expf(-fabsf(d));
  den = d > 0.0f ? fmaf(den, e, 1.0f) : den + e;
  mx = fmaxf(mx, l);
}

// One thread = 4 x 2 full-resolution pixels = 2 half-resolution pixels.
__global__ void __launch_bounds__(256) energy_prepare_vec_kernel(const float *__restrict__ simg,
                                                                 const float *__restrict__ logit,
                                                                 const float *__restrict__ label,
                                                                 const int *__restrict__ boxes, Affine3 aff,
                                                                 float *__restrict__ img_half,
                                                                 float *__restrict__ s_roi, float *__restrict__ gate,
                                                                 float *__restrict__ roi_out, int B, int C, int H,
                                                                 int W) {
  const int h = H / 2, w = W / 2, wq = W / 4;
  const long long total = (long long)B * h * wq;
  const size_t HW = (size_t)H * W, hw = (size_t)h * w;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(t % wq), yh = (int)((t / wq) % h), b = (int)(t / ((long long)wq * h));
    const int X = xq * 4, Y = yh * 2;
    const size_t src = (size_t)Y * W + X;
    const float *lg = logit + (size_t)b * C * HW + src;
    float mx[8], den[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { mx[i] = -INFINITY; den[i] = 0.0f; }
#pragma unroll 3
    for (int c = 0; c < C; ++c) {
      const float4 a = ldg_stream4(lg + (size_t)c * HW);
      const float4 d = ldg_stream4(lg + (size_t)c * HW + W);
      online_update(a.x, mx[0], den[0]); online_update(a.y, mx[1], den[1]);
      online_update(a.z, mx[2], den[2]); online_update(a.w, mx[3], den[3]);
      online_update(d.x, mx[4], den[4]); online_update(d.y, mx[5], den[5]);
      online_update(d.z, mx[6], den[6]); online_update(d.w, mx[7], den[7]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) den[i] = 1.0f / den[i];

    const int *box = boxes + 4 * b;
    const bool rowin = Y >= box[0] && Y < box[1];
    const float roi0 = (rowin && X >= box[2] && X < box[3]) ? 1.0f : 0.0f;           // nearest: pixel (2y, 2x)
    const float roi1 = (rowin && X + 2 >= box[2] && X + 2 < box[3]) ? 1.0f : 0.0f;
    const size_t pix = (size_t)yh * w + (X >> 1);
    float smax0 = -INFINITY, smax1 = -INFINITY;
    float *dst = s_roi + (size_t)b * C * hw + pix;
#pragma unroll 3
    for (int c = 0; c < C; ++c) {
      const float4 a = ldg4c(lg + (size_t)c * HW);
      const float4 d = ldg4c(lg + (size_t)c * HW + W);
      const float p00 = __expf(a.x - mx[0]) * den[0], p01 = __expf(a.y - mx[1]) * den[1];
      const float q00 = __expf(a.z - mx[2]) * den[2], q01 = __expf(a.w - mx[3]) * den[3];
      const float p10 = __expf(d.x - mx[4]) * den[4], p11 = __expf(d.y - mx[5]) * den[5];
      const float q10 = __expf(d.z - mx[6]) * den[6], q11 = __expf(d.w - mx[7]) * den[7];
      // exact 2:1 bilinear, align_corners=False: 0.25 * (((p00 + p01) + p10) + p11)
      const float s0 = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(p00, p01), p10), p11));
      const float s1 = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(q00, q01), q10), q11));
      smax0 = fmaxf(smax0, s0);
      smax1 = fmaxf(smax1, s1);
      *reinterpret_cast<float2 *>(dst + (size_t)c * hw) = make_float2(__fmul_rn(s0, roi0), __fmul_rn(s1, roi1));
    }
    // image (nearest, de-normalised), unlabel flag, gate, ROI for the two half-resolution pixels
    const float4 lab = ldg_stream4(label + (size_t)b * HW + src);
    if (img_half) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float4 v = ldg_stream4(simg + ((size_t)b * 3 + c) * HW + src);
        *reinterpret_cast<float2 *>(img_half + ((size_t)b * 3 + c) * hw + pix) =
            make_float2(__fadd_rn(__fmul_rn(v.x, aff.std[c]), aff.mean[c]), __fadd_rn(__fmul_rn(v.z, aff.std[c]), aff.mean[c]));
      }
    }
    float g0 = __fsub_rn(roi0, smax0), g1 = __fsub_rn(roi1, smax1);
    if (((int)lab.x & 255) == 255) g0 = 1.0f;
    if (((int)lab.z & 255) == 255) g1 = 1.0f;
    *reinterpret_cast<float2 *>(gate + (size_t)b * hw + pix) = make_float2(fmaxf(g0, 0.0f), fmaxf(g1, 0.0f));
    *reinterpret_cast<float2 *>(roi_out + (size_t)b * hw + pix) = make_float2(roi0, roi1);
  }
}

// One thread = 4 full-resolution pixels of one row (two half-resolution pixels of the saved AS / ROI).
__global__ void __launch_bounds__(256) energy_logit_grad_vec_kernel(const float *__restrict__ logit,
                                                                    const float *__restrict__ as_saved,
                                                                    const float *__restrict__ roi_half,
                                                                    const float *__restrict__ grad_out, float weight,
                                                                    float *__restrict__ grad_logit, int B, int C,
                                                                    int H, int W) {
  const int h = H / 2, w = W / 2, wq = W / 4;
  const long long total = (long long)B * H * wq;
  const size_t HW = (size_t)H * W, hw = (size_t)h * w;
  const float cbase = __fdiv_rn(__fmul_rn(-2.0f, __fmul_rn(__ldg(grad_out), weight)), (float)B) * 0.25f;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(t % wq), Y = (int)((t / wq) % H), b = (int)(t / ((long long)wq * H));
    const int X = xq * 4;
    const size_t pix = (size_t)(Y >> 1) * w + (X >> 1);
    const float *lg = logit + (size_t)b * C * HW + (size_t)Y * W + X;
    const float *as = as_saved + (size_t)b * C * hw + pix;
    float *out = grad_logit + (size_t)b * C * HW + (size_t)Y * W + X;
    const float2 roi = *reinterpret_cast<const float2 *>(roi_half + (size_t)b * hw + pix);
    float mx[4], den[4], num[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { mx[i] = -INFINITY; den[i] = 0.0f; num[i] = 0.0f; }
#pragma unroll 3
    for (int c = 0; c < C; ++c) {
      const float4 l = ldg_stream4(lg + (size_t)c * HW);
      const float2 a = __ldg(reinterpret_cast<const float2 *>(as + (size_t)c * hw));
      const float lv[4] = {l.x, l.y, l.z, l.w};
      const float av[4] = {a.x, a.x, a.y, a.y};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float d = lv[i] - mx[i];
        const float e = __expf(-fabsf(d));
        if (d > 0.0f) { den[i] = fmaf(den[i], e, 1.0f); num[i] = fmaf(num[i], e, av[i]); }
        else          { den[i] += e;                    num[i] = fmaf(e, av[i], num[i]); }
        mx[i] = fmaxf(mx[i], lv[i]);
      }
    }
    float dot[4], cf[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      den[i] = 1.0f / den[i];
      dot[i] = num[i] * den[i];                   // <p, AS> of this pixel
      cf[i] = cbase * (i < 2 ? roi.x : roi.y);
    }
#pragma unroll 3
    for (int c = 0; c < C; ++c) {
      const float4 l = ldg4c(lg + (size_t)c * HW);
      const float2 a = __ldg(reinterpret_cast<const float2 *>(as + (size_t)c * hw));
      float4 o;
      o.x = __expf(l.x - mx[0]) * den[0] * (cf[0] * (a.x - dot[0]));
      o.y = __expf(l.y - mx[1]) * den[1] * (cf[1] * (a.x - dot[1]));
      o.z = __expf(l.z - mx[2]) * den[2] * (cf[2] * (a.y - dot[2]));
      o.w = __expf(l.w - mx[3]) * den[3] * (cf[3] * (a.y - dot[3]));
      stg_stream4(out + (size_t)c * HW, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Register-resident variants for a compile-time class count (C = 21, the VOC shape): every logit is read from HBM
// exactly once and stays in registers between the softmax statistics and the output pass, all C loads of a thread are
// in flight together (no L2 re-read whose hit rate depends on how much of the grid is resident), and each element
// costs one ex2.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ldg_stream2(const float *p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}

// One thread = one half-resolution pixel = 2 x 2 source pixels.
template <int C>
__global__ void __launch_bounds__(128, 4) energy_prepare_reg_kernel(const float *__restrict__ simg,
                                                                    const float *__restrict__ logit,
                                                                    const float *__restrict__ label,
                                                                    const int *__restrict__ boxes, Affine3 aff,
                                                                    float *__restrict__ img_half,
                                                                    float *__restrict__ s_roi, float *__restrict__ gate,
                                                                    float *__restrict__ roi_out, int B, int H, int W) {
  const int h = H / 2, w = W / 2;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)B * h * w) return;
  const int x = (int)(t % w), y = (int)((t / w) % h), b = (int)(t / ((long long)w * h));
  const size_t HW = (size_t)H * W, hw = (size_t)h * w;
  const size_t src = (size_t)(2 * y) * W + 2 * x, pix = (size_t)y * w + x;
  const float *lg = logit + (size_t)b * C * HW + src;
  float2 top[C], bot[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    top[c] = ldg_stream2(lg + (size_t)c * HW);
    bot[c] = ldg_stream2(lg + (size_t)c * HW + W);
  }
  const float lab = __ldg(label + (size_t)b * HW + src);
  float im[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) im[c] = __ldg(simg + ((size_t)b * 3 + c) * HW + src);
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    m0 = fmaxf(m0, top[c].x); m1 = fmaxf(m1, top[c].y);
    m2 = fmaxf(m2, bot[c].x); m3 = fmaxf(m3, bot[c].y);
  }
  float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    top[c].x = __expf(top[c].x - m0); d0 += top[c].x;
    top[c].y = __expf(top[c].y - m1); d1 += top[c].y;
    bot[c].x = __expf(bot[c].x - m2); d2 += bot[c].x;
    bot[c].y = __expf(bot[c].y - m3); d3 += bot[c].y;
  }
  d0 = 1.0f / d0; d1 = 1.0f / d1; d2 = 1.0f / d2; d3 = 1.0f / d3;
  const int *box = boxes + 4 * b;
  const float roi = (2 * y >= box[0] && 2 * y < box[1] && 2 * x >= box[2] && 2 * x < box[3]) ? 1.0f : 0.0f;
  float smax = -INFINITY;
  float *dst = s_roi + (size_t)b * C * hw + pix;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    // exact 2:1 bilinear, align_corners=False: 0.25 * (((p00 + p01) + p10) + p11)
    const float sv = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(top[c].x * d0, top[c].y * d1), bot[c].x * d2),
                                                bot[c].y * d3));
    smax = fmaxf(smax, sv);
    dst[(size_t)c * hw] = __fmul_rn(sv, roi);
  }
  if (img_half) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      img_half[((size_t)b * 3 + c) * hw + pix] = __fadd_rn(__fmul_rn(im[c], aff.std[c]), aff.mean[c]);
  }
  float g = __fsub_rn(roi, smax);
  if (((int)lab & 255) == 255) g = 1.0f;
  gate[(size_t)b * hw + pix] = fmaxf(g, 0.0f);
  roi_out[(size_t)b * hw + pix] = roi;
}

// One thread = 4 full-resolution pixels of one row.
template <int C>
__global__ void __launch_bounds__(128, 2) energy_logit_grad_reg_kernel(const float *__restrict__ logit,
                                                                       const float *__restrict__ as_saved,
                                                                       const float *__restrict__ roi_half,
                                                                       const float *__restrict__ grad_out, float weight,
                                                                       float *__restrict__ grad_logit, int B, int H,
                                                                       int W) {
  const int h = H / 2, w = W / 2, wq = W / 4;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)B * H * wq) return;
  const size_t HW = (size_t)H * W, hw = (size_t)h * w;
  const int xq = (int)(t % wq), Y = (int)((t / wq) % H), b = (int)(t / ((long long)wq * H));
  const int X = xq * 4;
  const size_t pix = (size_t)(Y >> 1) * w + (X >> 1);
  const float *lg = logit + (size_t)b * C * HW + (size_t)Y * W + X;
  const float *as = as_saved + (size_t)b * C * hw + pix;
  float *out = grad_logit + (size_t)b * C * HW + (size_t)Y * W + X;
  float4 l[C];
  float2 a[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    l[c] = ldg_stream4(lg + (size_t)c * HW);
    a[c] = __ldg(reinterpret_cast<const float2 *>(as + (size_t)c * hw));
  }
  const float2 roi = *reinterpret_cast<const float2 *>(roi_half + (size_t)b * hw + pix);
  const float cbase = __fdiv_rn(__fmul_rn(-2.0f, __fmul_rn(__ldg(grad_out), weight)), (float)B) * 0.25f;
  float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    mx.x = fmaxf(mx.x, l[c].x); mx.y = fmaxf(mx.y, l[c].y); mx.z = fmaxf(mx.z, l[c].z); mx.w = fmaxf(mx.w, l[c].w);
  }
  float4 den = make_float4(0.f, 0.f, 0.f, 0.f), num = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    l[c].x = __expf(l[c].x - mx.x); den.x += l[c].x; num.x = fmaf(l[c].x, a[c].x, num.x);
    l[c].y = __expf(l[c].y - mx.y); den.y += l[c].y; num.y = fmaf(l[c].y, a[c].x, num.y);
    l[c].z = __expf(l[c].z - mx.z); den.z += l[c].z; num.z = fmaf(l[c].z, a[c].y, num.z);
    l[c].w = __expf(l[c].w - mx.w); den.w += l[c].w; num.w = fmaf(l[c].w, a[c].y, num.w);
  }
  den.x = 1.0f / den.x; den.y = 1.0f / den.y; den.z = 1.0f / den.z; den.w = 1.0f / den.w;
  const float4 dot = make_float4(num.x * den.x, num.y * den.y, num.z * den.z, num.w * den.w);   // <p, AS> per pixel
  const float c0 = cbase * roi.x, c1 = cbase * roi.y;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float4 o;
    o.x = l[c].x * den.x * (c0 * (a[c].x - dot.x));
    o.y = l[c].y * den.y * (c0 * (a[c].x - dot.y));
    o.z = l[c].z * den.z * (c1 * (a[c].y - dot.z));
    o.w = l[c].w * den.w * (c1 * (a[c].y - dot.w));
    stg_stream4(out + (size_t)c * HW, o);
  }
}

// ------------------------------------------------------------------------------------------------
// Register-resident variants for LARGE compile-time class counts (C = 81, the COCO shape).  A pixel quad per thread
// would need 4 C registers, so a thread owns TWO adjacent pixels of one row (2 C registers, 64-bit loads, a warp's
// load is one 256-byte row segment) and, in the forward kernel, exchanges its probabilities with the thread that owns
// the row below for the 2:1 reduction.  The two-pass kernels they replace are DRAM-bound on reading the logits twice
// (the logits of all resident threads - 786 MB at COCO B = 32 - do not fit in L2): 0.49 of the HBM peak.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ldg_stream2f(const float *p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}

// Warp = 32 full-resolution columns x 2 rows = 16 half-resolution pixels; lane = 16 * row + column pair.
template <int C>
__global__ void __launch_bounds__(128, 2) energy_prepare_pair_kernel(const float *__restrict__ simg,
                                                                     const float *__restrict__ logit,
                                                                     const float *__restrict__ label,
                                                                     const int *__restrict__ boxes, Affine3 aff,
                                                                     float *__restrict__ img_half,
                                                                     float *__restrict__ s_roi, float *__restrict__ gate,
                                                                     float *__restrict__ roi_out, int B, int H, int W) {
  const int h = H / 2, w = W / 2, wc = W / 32;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (warp >= (long long)B * h * wc) return;              // whole warps leave together
  const int lane = threadIdx.x & 31, r = lane >> 4, cp = lane & 15;
  const int xc = (int)(warp % wc), yh = (int)((warp / wc) % h), b = (int)(warp / ((long long)wc * h));
  const int X = 32 * xc + 2 * cp, Y = 2 * yh + r;
  const size_t HW = (size_t)H * W, hw = (size_t)h * w;
  const size_t src = (size_t)Y * W + X;
  const float *lg = logit + (size_t)b * C * HW + src;
  float2 p[C];
#pragma unroll
  for (int c = 0; c < C; ++c) p[c] = ldg_stream2f(lg + (size_t)c * HW);
  const bool writer = r == 0;                              // owns half-resolution pixel (yh, X / 2)
  float lab = 0.0f, im[3] = {0.f, 0.f, 0.f};
  if (writer) {
    lab = __ldg(label + (size_t)b * HW + src);
#pragma unroll
    for (int c = 0; c < 3; ++c) im[c] = __ldg(simg + ((size_t)b * 3 + c) * HW + src);
  }
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int c = 0; c < C; ++c) { m0 = fmaxf(m0, p[c].x); m1 = fmaxf(m1, p[c].y); }
  float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    p[c].x = __expf(p[c].x - m0); d0 += p[c].x;
    p[c].y = __expf(p[c].y - m1); d1 += p[c].y;
  }
  d0 = 1.0f / d0; d1 = 1.0f / d1;
  const int *box = boxes + 4 * b;
  const float roi = (Y >= box[0] && Y < box[1] && X >= box[2] && X < box[3]) ? 1.0f : 0.0f;   // writers: pixel (2y, 2x)
  const size_t pix = (size_t)yh * w + (X >> 1);
  float smax = -INFINITY;
  float *dst = s_roi + (size_t)b * C * hw + pix;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    // exact 2:1 bilinear, align_corners=False: 0.25 * (((p00 + p01) + p10) + p11); the lower row comes from lane + 16
    const float v0 = p[c].x * d0, v1 = p[c].y * d1;
    const float p10 = __shfl_xor_sync(0xffffffffu, v0, 16), p11 = __shfl_xor_sync(0xffffffffu, v1, 16);
    const float sv = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(v0, v1), p10), p11));
    if (writer) {
      smax = fmaxf(smax, sv);
      dst[(size_t)c * hw] = __fmul_rn(sv, roi);
    }
  }
  if (writer) {
    if (img_half) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        img_half[((size_t)b * 3 + c) * hw + pix] = __fadd_rn(__fmul_rn(im[c], aff.std[c]), aff.mean[c]);
    }
    float g = __fsub_rn(roi, smax);
    if (((int)lab & 255) == 255) g = 1.0f;
    gate[(size_t)b * hw + pix] = fmaxf(g, 0.0f);
    roi_out[(size_t)b * hw + pix] = roi;
  }
}

// One thread = two adjacent full-resolution pixels of one row = one half-resolution pixel's AS / ROI.  The C values
// of AS the thread needs twice (for <p, AS> and for the output) are parked in its own shared-memory column by
// cp.async, so that they neither occupy registers next to the 2 C logits nor are requested twice.
template <int C>
__global__ void __launch_bounds__(128, 2) energy_logit_grad_pair_kernel(const float *__restrict__ logit,
                                                                        const float *__restrict__ as_saved,
                                                                        const float *__restrict__ roi_half,
                                                                        const float *__restrict__ grad_out, float weight,
                                                                        float *__restrict__ grad_logit, int B, int H,
                                                                        int W) {
  __shared__ float s_as[C][128];
  const int h = H / 2, w = W / 2;
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool live = t < (long long)B * H * w;
  if (!live) t = (long long)B * H * w - 1;                 // clamp: the thread loads valid data and stores nothing
  const size_t HW = (size_t)H * W, hw = (size_t)h * w;
  const int xh = (int)(t % w), Y = (int)((t / w) % H), b = (int)(t / ((long long)w * H));
  const size_t pix = (size_t)(Y >> 1) * w + xh;
  const float *lg = logit + (size_t)b * C * HW + (size_t)Y * W + 2 * xh;
  const float *as = as_saved + (size_t)b * C * hw + pix;
  float *out = grad_logit + (size_t)b * C * HW + (size_t)Y * W + 2 * xh;
  const unsigned col = (unsigned)__cvta_generic_to_shared(&s_as[0][threadIdx.x]);
#pragma unroll
  for (int c = 0; c < C; ++c)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(col + c * 128 * 4), "l"(as + (size_t)c * hw) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
  float2 l[C];
#pragma unroll
  for (int c = 0; c < C; ++c) l[c] = ldg_stream2f(lg + (size_t)c * HW);
  const float roi = __ldg(roi_half + (size_t)b * hw + pix);
  const float cf = __fdiv_rn(__fmul_rn(-2.0f, __fmul_rn(__ldg(grad_out), weight)), (float)B) * 0.25f * roi;
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int c = 0; c < C; ++c) { m0 = fmaxf(m0, l[c].x); m1 = fmaxf(m1, l[c].y); }
  asm volatile("cp.async.wait_group 0;" ::: "memory");     // each thread reads back only what it copied itself
  float d0 = 0.0f, d1 = 0.0f, n0 = 0.0f, n1 = 0.0f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float a = s_as[c][threadIdx.x];
    l[c].x = __expf(l[c].x - m0); d0 += l[c].x; n0 = fmaf(l[c].x, a, n0);
    l[c].y = __expf(l[c].y - m1); d1 += l[c].y; n1 = fmaf(l[c].y, a, n1);
  }
  d0 = 1.0f / d0; d1 = 1.0f / d1;
  const float dot0 = n0 * d0, dot1 = n1 * d1;              // <p, AS> per pixel
  if (!live) return;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float a = s_as[c][threadIdx.x];
    float2 o;
    o.x = l[c].x * d0 * (cf * (a - dot0));
    o.y = l[c].y * d1 * (cf * (a - dot1));
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(out + (size_t)c * HW), "f"(o.x), "f"(o.y));
  }
}

static int grid1d(long long items) { return (int)max(1LL, min((long long)sm_count() * 8, ceil_div_ll(items, 256))); }

}  // namespace cosa

using namespace cosa;

// Workspace of the autograd Function: S*ROI [N,K,H,W], gate [N,H,W], then the lattice (one chunk).
extern "C" size_t cosa_dense_energy_ws_bytes(int N, int K, int H, int W) {
  if (N < 1 || K < 1 || H < 1 || W < 1) return 0;
  const size_t n = (size_t)H * W;
  return align_up((size_t)N * K * n * sizeof(float), 256) + align_up((size_t)N * n * sizeof(float), 256) + 256 +
         lattice_ws_bytes(lattice_chunk_images(N, K, H, W), K, H, W);
}

static float budget_of(int flags) { return (float)((flags >> 8) & 0xff) / 16.0f; }   // COSA_ENERGY_VERTEX_BUDGET

static int energy_core(const float *images, const float *s_roi, const float *gate, float *as_out, float *loss_out,
                       double *acc, int N, int K, int H, int W, float sigmargb, float sigmaxy, float weight,
                       int apply_weight, void *lattice_ws, cudaStream_t s, bool prebuilt = false, float vpp = 0.0f) {
  const size_t n = (size_t)H * W;
  COSA_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), s));
  const int chunk = lattice_chunk_images(N, K, H, W);
  if (prebuilt && chunk < N) return COSA_E_ARG;   // a prebuilt lattice covers the whole batch or nothing
  for (int n0 = 0; n0 < N; n0 += chunk) {   // chunks reuse the lattice workspace, stream-ordered
    const int nb = min(chunk, N - n0);
    LatticeBufs L;
    lattice_carve(lattice_ws, nb, K, H, W, &L, kLatD, vpp);
    if (!prebuilt)
      COSA_CHECK(lattice_build(L, images + (size_t)n0 * 3 * n, nb, H, W, sigmargb, sigmaxy, n0 == 0, s));
    COSA_CHECK(lattice_splat_blur(L, s_roi + (size_t)n0 * K * n, nb, K, H, W, s));
    COSA_CHECK(lattice_slice(L, s_roi + (size_t)n0 * K * n, gate + (size_t)n0 * n, acc, as_out + (size_t)n0 * K * n,
                             nb, K, H, W, s));
  }
  COSA_LAUNCH(energy_loss_finalize_kernel, 1, 1, 0, s, acc, loss_out, N, weight, apply_weight);
  return 0;
}

extern "C" int cosa_dense_energy_forward(const float *images, const float *segs, const float *rois,
                                         const unsigned char *unlabel, float sigmargb, float sigmaxy, float *as_out,
                                         float *loss_out, int N, int K, int H, int W, void *ws, size_t ws_bytes,
                                         void *stream) {
  if (!images || !segs || !rois || !unlabel || !as_out || !loss_out || !ws || N < 1 || K < 1 || H < 1 || W < 1)
    return COSA_E_ARG;
  if (ws_bytes < cosa_dense_energy_ws_bytes(N, K, H, W)) return COSA_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const int n = H * W;
  const long long P = (long long)N * n;
  Arena a(ws);
  float *s_roi = a.take<float>((size_t)N * K * n);
  float *gate = a.take<float>((size_t)P);
  double *acc = a.take<double>(1);
  void *lws = a.base + a.off;
  COSA_LAUNCH(energy_gate_kernel, grid1d(P), 256, 0, s, segs, rois, unlabel, s_roi, gate, K, n, P);
  return energy_core(images, s_roi, gate, as_out, loss_out, acc, N, K, H, W, sigmargb, sigmaxy, 1.0f, 0, lws, s);
}

extern "C" int cosa_dense_energy_backward(const float *as_saved, const float *rois, const float *grad_out,
                                          float *grad_seg, int N, int K, int H, int W, void *stream) {
  if (!as_saved || !rois || !grad_out || !grad_seg || N < 1 || K < 1 || H < 1 || W < 1) return COSA_E_ARG;
  const int n = H * W;
  const long long P = (long long)N * n;
  COSA_LAUNCH(energy_grad_kernel, grid1d(P), 256, 0, (cudaStream_t)stream, as_saved, rois, grad_out, grad_seg, K, n, P,
              N);
  return 0;
}

// ---- fused get_energy_loss ---------------------------------------------------------------------------
// saved = [ AS gated [B,C,h,w] | ROI [B,h,w] ];  ws = [ img_half [B,3,h,w] | S*ROI [B,C,h,w] | gate [B,h,w] | lattice ]
extern "C" size_t cosa_energy_loss_saved_bytes(int B, int C, int H, int W) {
  const size_t hw = (size_t)(H / 2) * (W / 2);
  return align_up((size_t)B * C * hw * sizeof(float), 256) + align_up((size_t)B * hw * sizeof(float), 256);
}

// bytes in front of the lattice in the workspace of the fused calls (independent of the vertex budget)
extern "C" size_t cosa_energy_loss_lattice_offset(int B, int C, int H, int W) {
  if (B < 1 || C < 1 || H < 2 || W < 2) return 0;
  const size_t hw = (size_t)(H / 2) * (W / 2);
  return align_up((size_t)B * 3 * hw * sizeof(float), 256) + align_up((size_t)B * C * hw * sizeof(float), 256) +
         align_up((size_t)B * hw * sizeof(float), 256) + 256;
}

extern "C" size_t cosa_energy_loss_ws_bytes_ex(int B, int C, int H, int W, int flags) {
  if (B < 1 || C < 1 || H < 2 || W < 2) return 0;
  return cosa_energy_loss_lattice_offset(B, C, H, W) +
         lattice_ws_bytes(lattice_chunk_images(B, C, H / 2, W / 2), C, H / 2, W / 2, kLatD, budget_of(flags));
}

extern "C" size_t cosa_energy_loss_ws_bytes(int B, int C, int H, int W) {
  return cosa_energy_loss_ws_bytes_ex(B, C, H, W, 0);
}

// The image-only half of the forward: de-normalised nearest 2:1 image (the same two roundings as the prepare kernels)
// and the lattice build.  Runs on any stream: the lattice does not depend on the labels or the logits.
__global__ void __launch_bounds__(256) energy_img_half_kernel(const float *__restrict__ simg, Affine3 aff,
                                                              float *__restrict__ img_half, int planes, int H, int W) {
  const int h = H / 2, w2 = W / 4;                 // one thread = two half-resolution pixels from one float4
  const long long total = (long long)planes * h * w2;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int xq = (int)(i % w2);
    const long long t = i / w2;
    const int y = (int)(t % h), p = (int)(t / h), c = p % 3;
    const float4 v = ldg_stream4(simg + ((size_t)p * H + 2 * y) * W + 4 * xq);
    *reinterpret_cast<float2 *>(img_half + ((size_t)p * h + y) * (W / 2) + 2 * xq) =
        make_float2(__fadd_rn(__fmul_rn(v.x, aff.std[c]), aff.mean[c]), __fadd_rn(__fmul_rn(v.z, aff.std[c]), aff.mean[c]));
  }
}

__global__ void __launch_bounds__(256) energy_img_half_scalar_kernel(const float *__restrict__ simg, Affine3 aff,
                                                                     float *__restrict__ img_half, int planes, int H,
                                                                     int W) {
  const int h = H / 2, w = W / 2;
  const long long total = (long long)planes * h * w;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int x = (int)(i % w);
    const long long t = i / w;
    const int y = (int)(t % h), p = (int)(t / h), c = p % 3;
    img_half[i] = __fadd_rn(__fmul_rn(__ldg(simg + ((size_t)p * H + 2 * y) * W + 2 * x), aff.std[c]), aff.mean[c]);
  }
}

extern "C" int cosa_energy_loss_prebuild(const float *simg, const float *mean, const float *std, float sigmargb,
                                         float sigmaxy_scaled, int B, int C, int H, int W, void *ws, size_t ws_bytes,
                                         void *stream) {
  return cosa_energy_loss_prebuild_ex(simg, mean, std, sigmargb, sigmaxy_scaled, B, C, H, W, ws, ws_bytes, 0, stream);
}

extern "C" int cosa_energy_loss_prebuild_ex(const float *simg, const float *mean, const float *std, float sigmargb,
                                            float sigmaxy_scaled, int B, int C, int H, int W, void *ws, size_t ws_bytes,
                                            int flags, void *stream) {
  if (flags & ~COSA_ENERGY_VERTEX_BUDGET_MASK) return COSA_E_ARG;
  if (!simg || !mean || !std || !ws || B < 1 || C < 1) return COSA_E_ARG;
  if (H < 2 || W < 2 || (H & 1) || (W & 1)) return COSA_E_ARG;
  if (ws_bytes < cosa_energy_loss_ws_bytes_ex(B, C, H, W, flags)) return COSA_E_WORKSPACE;
  const int h = H / 2, w = W / 2;
  if (lattice_chunk_images(B, C, h, w) < B) return COSA_E_ARG;   // more than one lattice chunk: nothing to prebuild
  cudaStream_t s = (cudaStream_t)stream;
  const size_t hw = (size_t)h * w;
  Arena a(ws);                                     // the carve of cosa_energy_loss_forward
  float *img_half = a.take<float>((size_t)B * 3 * hw);
  a.take<float>((size_t)B * C * hw);
  a.take<float>((size_t)B * hw);
  a.take<double>(1);
  void *lws = a.base + a.off;
  Affine3 aff;
  for (int c = 0; c < 3; ++c) { aff.mean[c] = mean[c]; aff.std[c] = std[c]; }
  if (W % 4 == 0) {
    COSA_LAUNCH(energy_img_half_kernel, grid1d((long long)B * 3 * h * (W / 4)), 256, 0, s, simg, aff, img_half, B * 3, H, W);
  } else {
    COSA_LAUNCH(energy_img_half_scalar_kernel, grid1d((long long)B * 3 * hw), 256, 0, s, simg, aff, img_half, B * 3, H, W);
  }
  LatticeBufs L;
  lattice_carve(lws, B, C, h, w, &L, kLatD, budget_of(flags));
  return lattice_build(L, img_half, B, h, w, sigmargb, sigmaxy_scaled, true, s);
}

extern "C" int cosa_energy_loss_forward(const float *simg, const float *logit, const float *label, const int *boxes,
                                        const float *mean, const float *std, float weight, float sigmargb,
                                        float sigmaxy_scaled, float *loss_out, void *saved, int B, int C, int H, int W,
                                        void *ws, size_t ws_bytes, void *stream) {
  return cosa_energy_loss_forward_flags(simg, logit, label, boxes, mean, std, weight, sigmargb, sigmaxy_scaled, loss_out,
                                        saved, B, C, H, W, ws, ws_bytes, 0, stream);
}

extern "C" int cosa_energy_loss_forward_flags(const float *simg, const float *logit, const float *label,
                                              const int *boxes, const float *mean, const float *std, float weight,
                                              float sigmargb, float sigmaxy_scaled, float *loss_out, void *saved, int B,
                                              int C, int H, int W, void *ws, size_t ws_bytes, int flags, void *stream) {
  return cosa_energy_loss_forward_ev(simg, logit, label, boxes, mean, std, weight, sigmargb, sigmaxy_scaled, loss_out,
                                     saved, B, C, H, W, ws, ws_bytes, flags, nullptr, stream);
}

extern "C" int cosa_energy_loss_forward_ev(const float *simg, const float *logit, const float *label, const int *boxes,
                                           const float *mean, const float *std, float weight, float sigmargb,
                                           float sigmaxy_scaled, float *loss_out, void *saved, int B, int C, int H,
                                           int W, void *ws, size_t ws_bytes, int flags, void *lattice_ready,
                                           void *stream) {
  if (flags & ~(COSA_ENERGY_LATTICE_PREBUILT | COSA_ENERGY_VERTEX_BUDGET_MASK)) return COSA_E_ARG;
  if (lattice_ready && !(flags & COSA_ENERGY_LATTICE_PREBUILT)) return COSA_E_ARG;
  if (!simg || !logit || !label || !boxes || !mean || !std || !loss_out || !saved || !ws || B < 1 || C < 1)
    return COSA_E_ARG;
  if (H < 2 || W < 2 || (H & 1) || (W & 1)) return COSA_E_ARG;
  if (ws_bytes < cosa_energy_loss_ws_bytes_ex(B, C, H, W, flags)) return COSA_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const int h = H / 2, w = W / 2;
  const size_t hw = (size_t)h * w;
  Arena sv(saved);
  float *as_out = sv.take<float>((size_t)B * C * hw);
  float *roi_half = sv.take<float>((size_t)B * hw);
  Arena a(ws);
  float *img_half_ws = a.take<float>((size_t)B * 3 * hw);
  float *s_roi = a.take<float>((size_t)B * C * hw);
  float *gate = a.take<float>((size_t)B * hw);
  double *acc = a.take<double>(1);
  void *lws = a.base + a.off;
  // a prebuilt lattice comes with its half-resolution image: the prepare kernel leaves it alone, so it may run while
  // the build is still in flight on another stream (lattice_ready)
  float *img_half = (flags & COSA_ENERGY_LATTICE_PREBUILT) ? nullptr : img_half_ws;
  Affine3 aff;
  for (int c = 0; c < 3; ++c) { aff.mean[c] = mean[c]; aff.std[c] = std[c]; }
  // the vector kernels read 8 / 16 bytes at a time: a tensor view that starts at an odd element takes the scalar kernel
  const uintptr_t in_bits = (uintptr_t)simg | (uintptr_t)logit | (uintptr_t)label;
  const bool al8 = in_bits % 8 == 0, al16 = in_bits % 16 == 0;
  if (C == 21 && al8) {   // VOC: register-resident single pass
    const long long threads = (long long)B * h * w;
    COSA_LAUNCH(energy_prepare_reg_kernel<21>, (unsigned)ceil_div_ll(threads, 128), 128, 0, s, simg, logit, label, boxes,
                aff, img_half, s_roi, gate, roi_half, B, H, W);
  } else if (C == 81 && W % 32 == 0 && al8) {   // COCO: register-resident, two pixels per thread
    const long long threads = (long long)B * h * (W / 32) * 32;
    COSA_LAUNCH(energy_prepare_pair_kernel<81>, (unsigned)ceil_div_ll(threads, 128), 128, 0, s, simg, logit, label,
                boxes, aff, img_half, s_roi, gate, roi_half, B, H, W);
  } else if (W % 4 == 0 && al16) {
    const long long threads = (long long)B * h * (W / 4);
    COSA_LAUNCH(energy_prepare_vec_kernel, grid1d(threads), 256, 0, s, simg, logit, label, boxes, aff, img_half, s_roi,
                gate, roi_half, B, C, H, W);
  } else {
    dim3 grid(ceil_div(w, 32), ceil_div(h, 8), B);
    COSA_LAUNCH(energy_prepare_kernel, grid, 256, 0, s, simg, logit, label, boxes, aff, img_half, s_roi, gate,
                roi_half, C, H, W);
  }
  if (lattice_ready) COSA_CUDA(cudaStreamWaitEvent(s, (cudaEvent_t)lattice_ready, 0));
  return energy_core(img_half_ws, s_roi, gate, as_out, loss_out, acc, B, C, h, w, sigmargb, sigmaxy_scaled, weight, 1,
                     lws, s, (flags & COSA_ENERGY_LATTICE_PREBUILT) != 0, budget_of(flags));
}

extern "C" int cosa_energy_loss_backward(const float *logit, const void *saved, const float *grad_out, float weight,
                                         float *grad_logit, int B, int C, int H, int W, void *stream) {
  if (!logit || !saved || !grad_out || !grad_logit || B < 1 || C < 1 || H < 2 || W < 2 || (H & 1) || (W & 1))
    return COSA_E_ARG;
  const size_t hw = (size_t)(H / 2) * (W / 2);
  Arena sv(const_cast<void *>(saved));
  const float *as_saved = sv.take<float>((size_t)B * C * hw);
  const float *roi_half = sv.take<float>((size_t)B * hw);
  dim3 grid(ceil_div(W, 32), ceil_div(H, 8), B);
  cudaStream_t s = (cudaStream_t)stream;
  const uintptr_t io_bits = (uintptr_t)logit | (uintptr_t)grad_logit;
  const bool al8 = io_bits % 8 == 0, al16 = io_bits % 16 == 0;
  if (C == 21 && W % 4 == 0 && al16) {
    const long long threads = (long long)B * H * (W / 4);
    COSA_LAUNCH(energy_logit_grad_reg_kernel<21>, (unsigned)ceil_div_ll(threads, 128), 128, 0, s, logit, as_saved,
                roi_half, grad_out, weight, grad_logit, B, H, W);
  } else if (C == 81 && al8) {
    const long long threads = (long long)B * H * (W / 2);
    COSA_LAUNCH(energy_logit_grad_pair_kernel<81>, (unsigned)ceil_div_ll(threads, 128), 128, 0, s, logit, as_saved,
                roi_half, grad_out, weight, grad_logit, B, H, W);
  } else if (W % 4 == 0 && al16) {
    const long long threads = (long long)B * H * (W / 4);
    COSA_LAUNCH(energy_logit_grad_vec_kernel, grid1d(threads), 256, 0, s, logit, as_saved, roi_half, grad_out, weight,
                grad_logit, B, C, H, W);
  } else {
    COSA_LAUNCH(energy_logit_grad_kernel, grid, 256, 0, s, logit, as_saved, roi_half, grad_out, weight, grad_logit, C,
                H, W, B);
  }
  return 0;
}
