// CAM normalise / validation / labelling kernels (reference: utils/seg_helper.py:264-270, 515-551, 721-797).
//
// All of these are streaming, HBM-bound passes: 128-bit loads where the layout allows, warp shuffles for the
// per-pixel channel reduction, grids sized from the SM count (persistent grid-stride) or one tile per CTA.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "par.cuh"

namespace cosa {

// ------------------------------------------------------------------------------------------------
// cam_validation (seg_helper.py:547-551): out = cls_label[b,c] * cam.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cam_validation_kernel(const float *__restrict__ cam,
                                                             const float *__restrict__ cls_label,
                                                             float *__restrict__ out, long long HW4, long long HW,
                                                             int planes) {
  // grid.y strides over planes, grid.x over float4 chunks of the plane (HW % 4 == 0 path)
  for (int p = blockIdx.y; p < planes; p += gridDim.y) {
    const float l = __ldg(cls_label + p);
    const float *src = cam + (size_t)p * HW;
    float *dst = out + (size_t)p * HW;
    if (l == 0.0f) {
      // absent class: the plane is written as zeros without being read (18 of 20 planes at VOC).  For finite inputs
      // this equals the product up to the sign of zero; a NaN/Inf in an absent plane does not propagate as it would
      // through the reference's multiplication (deliberate differences)
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW4;
           i += (long long)gridDim.x * blockDim.x)
        stg_stream4(dst + 4 * i, z);
      continue;
    }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW4;
         i += (long long)gridDim.x * blockDim.x) {
      float4 v = ldg_stream4(src + 4 * i);
      v.x = __fmul_rn(l, v.x); v.y = __fmul_rn(l, v.y); v.z = __fmul_rn(l, v.z); v.w = __fmul_rn(l, v.w);
      stg_stream4(dst + 4 * i, v);
    }
  }
}
__global__ void cam_validation_scalar_kernel(const float *__restrict__ cam, const float *__restrict__ cls_label,
                                             float *__restrict__ out, long long HW, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = __fmul_rn(__ldg(cls_label + i / HW), cam[i]);
}

// ------------------------------------------------------------------------------------------------
// denormalize_img (utils/torch_helper.py:354-367, caller main.py:117): the [0,1] image cam2mask and PAR read is
// derived on the device from the ImageNet-normalised network input: (uint8)(x * std + mean) / 255, the product and
// the sum rounded separately as torch evaluates them, the cast truncating toward zero.
// ------------------------------------------------------------------------------------------------
struct Affine3f {
  float mean[3], std[3];
};
// u / 255 for an integer u in [0, 255], correctly rounded without the division routine: q = RN(u * RN(1/255)),
// r = u - 255 q (exact in an FMA), RN(q + r * RN(1/255)) - equal to IEEE division for all 256 values (checked).
__device__ __forceinline__ float div255_u8(float u) {
  const float c = 0.0039215688593685627f;   // RN(1/255) = 0x3B808081
  const float q = __fmul_rn(u, c);
  return __fmaf_rn(__fmaf_rn(-255.0f, q, u), c, q);
}
__global__ void __launch_bounds__(256) denormalize_img_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                              Affine3f a, long long HW4, long long HW, int planes) {
  for (int p = blockIdx.y; p < planes; p += gridDim.y) {
    const float sd = a.std[p % 3], mu = a.mean[p % 3];
    const float *src = in + (size_t)p * HW;
    float *dst = out + (size_t)p * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW4;
         i += (long long)gridDim.x * blockDim.x) {
      float4 v = ldg_stream4(src + 4 * i);
      float *e = reinterpret_cast<float *>(&v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float t = __fadd_rn(__fmul_rn(e[k], sd), mu);
        // uint8 cast of an in-range value: truncation (out-of-range inputs are unspecified in the reference itself)
        const int u = min(max((int)t, 0), 255);
        e[k] = div255_u8((float)u);
      }
      stg_stream4(dst + 4 * i, v);
    }
  }
}
__global__ void denormalize_img_scalar_kernel(const float *__restrict__ in, float *__restrict__ out, Affine3f a,
                                              long long HW, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i / HW) % 3);
    const float t = __fadd_rn(__fmul_rn(in[i], a.std[c]), a.mean[c]);
    out[i] = div255_u8((float)min(max((int)t, 0), 255));
  }
}

// ------------------------------------------------------------------------------------------------
// CAM normalise (seg_helper.py:264-270).  Pass 1: per-plane min/max of the scale sum (several CTAs per
// plane, float atomics through the order-preserving int mapping).  Pass 2: (sum + (-min)) / (max' + 1e-5),
// max' = rn(max + (-min)) (rounding is monotonic, so this is the max of the shifted plane).
// ------------------------------------------------------------------------------------------------
constexpr int kMaxScales = 8;
struct ScalePtrs {
  const float *p[kMaxScales];
};

__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void minmax_init_kernel(int *mm, int planes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < planes) {
    mm[2 * i] = float_to_ordered(INFINITY);
    mm[2 * i + 1] = float_to_ordered(-INFINITY);
  }
}

__device__ __forceinline__ float scale_sum(const ScalePtrs &sp, int n_scales, size_t idx) {
  float s = __ldg(sp.p[0] + idx);
  for (int k = 1; k < n_scales; ++k) s = __fadd_rn(s, __ldg(sp.p[k] + idx));
  return s;
}

__global__ void __launch_bounds__(256) cam_minmax_kernel(ScalePtrs sp, int n_scales, int *__restrict__ mm,
                                                         long long HW, int planes) {
  __shared__ float s_min[8], s_max[8];
  for (int p = blockIdx.y; p < planes; p += gridDim.y) {
    float lo = INFINITY, hi = -INFINITY;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW;
         i += (long long)gridDim.x * blockDim.x) {
      const float s = scale_sum(sp, n_scales, (size_t)p * HW + i);
      lo = fminf(lo, s);
      hi = fmaxf(hi, s);
    }
    lo = warp_min(lo);
    hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = lo; s_max[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x < 32) {
      lo = threadIdx.x < 8 ? s_min[threadIdx.x] : INFINITY;
      hi = threadIdx.x < 8 ? s_max[threadIdx.x] : -INFINITY;
      lo = warp_min(lo);
      hi = warp_max(hi);
      if (threadIdx.x == 0) {
        atomicMin(mm + 2 * p, float_to_ordered(lo));
        atomicMax(mm + 2 * p + 1, float_to_ordered(hi));
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) cam_normalize_apply_kernel(ScalePtrs sp, int n_scales,
                                                                  const int *__restrict__ mm, float *__restrict__ out,
                                                                  long long HW, int planes) {
  for (int p = blockIdx.y; p < planes; p += gridDim.y) {
    const float neg_min = -ordered_to_float(mm[2 * p]);                 // max(-cam)
    const float top = __fadd_rn(ordered_to_float(mm[2 * p + 1]), neg_min);   // max(cam - min)
    const float den = __fadd_rn(top, 1e-5f);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW;
         i += (long long)gridDim.x * blockDim.x) {
      const size_t idx = (size_t)p * HW + i;
      out[idx] = __fdiv_rn(__fadd_rn(scale_sum(sp, n_scales, idx), neg_min), den);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// cam_to_label (seg_helper.py:515-545).  A warp covers 32 consecutive pixels as 8 float4 columns x 4 channel
// groups; each lane scans channels cg, cg+4, ... of its 4 pixels with 128-bit loads, then the 4 groups are
// merged with two xor-shuffles (larger value wins, lower channel on ties = torch.max's first index).
// ------------------------------------------------------------------------------------------------
struct LabelArgs {
  float bkg_thre, high_thre, low_thre;
  int ignore_mid;
  long long ignore_index;
};

__device__ __forceinline__ void argmax_merge(float &v, int &idx, float ov, int oidx) {
  // NaN handling follows "first maximal"; CAMs on this path are finite.
  if (ov > v || (ov == v && oidx < idx)) { v = ov; idx = oidx; }
}

__device__ __forceinline__ long long decide_label(float v, int idx, int y, int x, const int *box,
                                                  const LabelArgs &a) {
  long long lab = (long long)idx + 1;
  if (v <= a.bkg_thre) lab = 0;
  if (box) {
    if (a.ignore_mid) {
      if (v <= a.high_thre) lab = a.ignore_index;
      if (v <= a.low_thre) lab = 0;
    }
    if (!(y >= box[0] && y < box[1] && x >= box[2] && x < box[3])) lab = a.ignore_index;
  }
  return lab;
}

template <bool VEC4>
__global__ void __launch_bounds__(256) cam_to_label_kernel(const float *__restrict__ cam,
                                                           const float *__restrict__ cls_label,
                                                           const int *__restrict__ boxes,
                                                           float *__restrict__ valid_out,
                                                           long long *__restrict__ label_out, int B, int C1, int H, int W,
                                                           LabelArgs args) {
  const long long HW = (long long)H * W;
  const int lane = threadIdx.x & 31, cg = lane >> 3, pq = lane & 7;
  const long long warps_per_img = ceil_div_ll(HW, 32);
  const long long total_warps = warps_per_img * B;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long warp_stride = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long wi = warp0; wi < total_warps; wi += warp_stride) {
    const int b = (int)(wi / warps_per_img);
    const long long p0 = (wi % warps_per_img) * 32 + pq * 4;   // first of this lane's 4 pixels
    const float *base = cam + (size_t)b * C1 * HW;
    float *vbase = valid_out ? valid_out + (size_t)b * C1 * HW : nullptr;
    float bv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int bi[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    for (int c = cg; c < C1; c += 4) {
      const float l = cls_label ? __ldg(cls_label + (size_t)b * C1 + c) : 1.0f;
      float v[4];
      if (VEC4) {
        if (p0 < HW) {
          const float4 t = ldg_stream4(base + (size_t)c * HW + p0);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
          v[0] = v[1] = v[2] = v[3] = 0.0f;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (p0 + k < HW) ? __ldg(base + (size_t)c * HW + p0 + k) : 0.0f;
      }
      if (cls_label) {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __fmul_rn(l, v[k]);
      }
      if (vbase) {
        if (VEC4) {
          if (p0 < HW) stg_stream4(vbase + (size_t)c * HW + p0, make_float4(v[0], v[1], v[2], v[3]));
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (p0 + k < HW) vbase[(size_t)c * HW + p0 + k] = v[k];
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (v[k] > bv[k]) { bv[k] = v[k]; bi[k] = c; }   // strictly greater: the first maximum is kept
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv[k], o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi[k], o);
        argmax_merge(bv[k], bi[k], ov, oi);
      }
    }
    if (cg == 0) {
      const int *box = boxes ? boxes + 4 * b : nullptr;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long p = p0 + k;
        if (p < HW) {
          const int y = (int)(p / W), x = (int)(p % W);
          label_out[(size_t)b * HW + p] = decide_label(bv[k], bi[k], y, x, box, args);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// cam2mask (seg_helper.py:721-785), batched over images and over the two threshold stacks.
//
//  keys     : per image the list [0] + [c+1 : cls_label[b,c] != 0]  (torch.nonzero order)         :762-767
//  prepare  : half-resolution image (bilinear, align_corners=False) and, per image, the softmax over the
//             live channels of [threshold, down(cams)] for the high and the low threshold          :737-769
//  (PAR)    : par_kernels.cu, both stacks of every image in one ragged batch                         :790
//  finalize : bilinear up-sampling of the refined stacks, argmax (first max), key lookup, box crop
//             and the high/low merge rule                                                        :793-795, :777-783
// Mask buffers are [B, 2*C, h, w]: channels [0,nc) hold the high stack, [nc,2nc) the low stack.
// ------------------------------------------------------------------------------------------------
__global__ void cam2mask_keys_kernel(const float *__restrict__ cls_labels, int *__restrict__ keys,
                                     int *__restrict__ nc_out, int *__restrict__ nch_out, int B, int C1,
                                     int derive, int cap, int *__restrict__ err) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int *k = keys + (size_t)b * (C1 + 1);
  int n = 0;
  k[n++] = 0;
  for (int c = 0; c < C1; ++c)
    if (cls_labels[(size_t)b * C1 + c] != 0.0f) {
      if (n <= cap) k[n++] = c + 1;
      else atomicExch(err, 1);      // more present classes than the mask buffers were sized for (COSA_CAM2MASK_MAX_CLASSES)
    }
  nc_out[b] = n;
  nch_out[b] = 2 * (n - derive);   // stored channels: both stacks, minus the derived one of each
}

struct ResizeGeom {
  int H, W, h, w;       // full and reduced size
  float sy, sx;         // source-index scale for the reduction (H/h, W/w); 1 when downscale == 0
  int identity;         // no reduction (downscale == 0)
};

// Optional producers folded into the prepare kernels (cosa_cam2mask_ex):
//  * images given as the ImageNet-normalised network input: each source pixel is de-normalised exactly as
//    denormalize_img_kernel does it (utils/torch_helper.py:354-367) before the reduction - no [0,1] image in HBM;
//  * CAMs given before cam_validation: each source value is multiplied by cls_label[b,c] (seg_helper.py:547-551);
//    only the planes of present classes are ever read, so the zero planes cam_validation would write never exist.
struct Denorm {
  float mean[3], std[3];
  int on;
};
__device__ __forceinline__ float denorm_px(float v, const Denorm &d, int c) {
  if (!d.on) return v;
  const float t = __fadd_rn(__fmul_rn(v, d.std[c]), d.mean[c]);
  return div255_u8((float)min(max((int)t, 0), 255));
}

// `scale` multiplies every source value (1 = identity, bit for bit)
__device__ __forceinline__ float sample_down(const float *plane, const ResizeGeom &g, const Tap &ty, const Tap &tx,
                                             float scale = 1.0f) {
  if (g.identity) return __fmul_rn(__ldg(plane + (size_t)ty.i0 * g.W + tx.i0), scale);
  return bilerp_down(ty, tx, __fmul_rn(__ldg(plane + (size_t)ty.i0 * g.W + tx.i0), scale),
                     __fmul_rn(__ldg(plane + (size_t)ty.i0 * g.W + tx.i1), scale),
                     __fmul_rn(__ldg(plane + (size_t)ty.i1 * g.W + tx.i0), scale),
                     __fmul_rn(__ldg(plane + (size_t)ty.i1 * g.W + tx.i1), scale));
}
__device__ __forceinline__ float sample_down_img(const float *plane, const ResizeGeom &g, const Tap &ty, const Tap &tx,
                                                 const Denorm &d, int c) {
  if (g.identity) return denorm_px(__ldg(plane + (size_t)ty.i0 * g.W + tx.i0), d, c);
  return bilerp_down(ty, tx, denorm_px(__ldg(plane + (size_t)ty.i0 * g.W + tx.i0), d, c),
                     denorm_px(__ldg(plane + (size_t)ty.i0 * g.W + tx.i1), d, c),
                     denorm_px(__ldg(plane + (size_t)ty.i1 * g.W + tx.i0), d, c),
                     denorm_px(__ldg(plane + (size_t)ty.i1 * g.W + tx.i1), d, c));
}

constexpr int kCacheC = 8;   // live channels cached in registers by the prepare kernel

__global__ void __launch_bounds__(256) cam2mask_prepare_kernel(const float *__restrict__ images,
                                                               const float *__restrict__ cams,
                                                               const int *__restrict__ keys,
                                                               const int *__restrict__ nc_dev,
                                                               float *__restrict__ img_small,
                                                               float *__restrict__ masks, MaskLayout ml, ResizeGeom g,
                                                               int C1, float thr_high, float thr_low, int derive,
                                                               Denorm dn, const float *__restrict__ cam_scale,
                                                               int cstride) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (x >= g.w || y >= g.h) return;
  const size_t HW = (size_t)g.H * g.W, hw = (size_t)g.h * g.w;
  const size_t pix = (size_t)y * g.w + x;
  Tap ty, tx;
  if (g.identity) {
    ty.i0 = ty.i1 = y; tx.i0 = tx.i1 = x; ty.w0 = tx.w0 = 1.0f; ty.w1 = tx.w1 = 0.0f;
  } else {
    ty = tap_half_pixel(y, g.sy, g.H);
    tx = tap_half_pixel(x, g.sx, g.W);
  }
  if (img_small) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      img_small[((size_t)b * 3 + c) * hw + pix] = sample_down_img(images + ((size_t)b * 3 + c) * HW, g, ty, tx, dn, c);
  }
  const int nc = nc_dev[b];
  const int *key = keys + (size_t)b * (C1 + 1);
  const float *cam_b = cams + (size_t)b * C1 * HW;
  // mask rows may be padded (MaskLayout): interior at column ml.off, ml.padn replicated columns either side
  const size_t mplane = (size_t)g.h * ml.pitch;
  float *m_hi = masks + (size_t)b * cstride * mplane + (size_t)y * ml.pitch + ml.off + x;
  const int ns = nc - derive;   // channels stored per stack (the last live one is derived, see cosa_cam2mask)
  float *m_lo = m_hi + (size_t)ns * mplane;
  auto put = [&](float *dst, float val) {
    *dst = val;
    if (ml.padn) {
      if (x == 0)
        for (int i = 1; i <= ml.padn; ++i) dst[-i] = val;
      if (x == g.w - 1)
        for (int i = 1; i <= ml.padn; ++i) dst[i] = val;
    }
  };

  // live foreground channels: values, their max; the two stacks differ only in channel 0 (the threshold)
  float v[kCacheC];
  float mx = -INFINITY;
  const float *sc_b = cam_scale ? cam_scale + (size_t)b * C1 : nullptr;
  auto scale_of = [&](int j) { return sc_b ? __ldg(sc_b + key[j] - 1) : 1.0f; };
  for (int j = 1; j < nc; ++j) {
    const float t = sample_down(cam_b + (size_t)(key[j] - 1) * HW, g, ty, tx, scale_of(j));
    if (j < kCacheC) v[j] = t;
    mx = fmaxf(mx, t);
  }
  const float mx_hi = fmaxf(mx, thr_high), mx_lo = fmaxf(mx, thr_low);
  float e0_hi = expf(thr_high - mx_hi), e0_lo = expf(thr_low - mx_lo);
  float den_hi = e0_hi, den_lo = e0_lo;
  for (int j = 1; j < nc; ++j) {
    const float t = (j < kCacheC) ? v[j] : sample_down(cam_b + (size_t)(key[j] - 1) * HW, g, ty, tx, scale_of(j));
    den_hi += expf(t - mx_hi);
    den_lo += expf(t - mx_lo);
  }
  if (ns > 0) {
    put(m_hi, e0_hi / den_hi);
    put(m_lo, e0_lo / den_lo);
  }
  for (int j = 1; j < ns; ++j) {
    const float t = (j < kCacheC) ? v[j] : sample_down(cam_b + (size_t)(key[j] - 1) * HW, g, ty, tx, scale_of(j));
    put(m_hi + (size_t)j * mplane, expf(t - mx_hi) / den_hi);
    put(m_lo + (size_t)j * mplane, expf(t - mx_lo) / den_lo);
  }
}

// Exact 2:1 reduction (H = 2h, W = 2w, W % 4 == 0): a thread produces TWO adjacent half-resolution pixels from
// 128-bit loads of the two source rows of every plane (8 scalar loads with their index arithmetic per plane and pixel
// in the general kernel).  Arithmetic identical to cam2mask_prepare_kernel: the taps of the exact ratio are the four
// source pixels with weight 0.25 each, accumulated in bilerp_down's order.
__device__ __forceinline__ void down2_sum(const float4 &a, const float4 &b, float &s0, float &s1);
__device__ __forceinline__ void down2_pair(const float *plane, size_t row0, int W, float &s0, float &s1,
                                           float scale = 1.0f) {
  float4 a = __ldg(reinterpret_cast<const float4 *>(plane + row0));
  float4 b = __ldg(reinterpret_cast<const float4 *>(plane + row0 + W));
  a.x = __fmul_rn(a.x, scale); a.y = __fmul_rn(a.y, scale); a.z = __fmul_rn(a.z, scale); a.w = __fmul_rn(a.w, scale);
  b.x = __fmul_rn(b.x, scale); b.y = __fmul_rn(b.y, scale); b.z = __fmul_rn(b.z, scale); b.w = __fmul_rn(b.w, scale);
  down2_sum(a, b, s0, s1);
}
__device__ __forceinline__ void down2_pair_img(const float *plane, size_t row0, int W, float &s0, float &s1,
                                               const Denorm &d, int c) {
  float4 a = __ldg(reinterpret_cast<const float4 *>(plane + row0));
  float4 b = __ldg(reinterpret_cast<const float4 *>(plane + row0 + W));
  a.x = denorm_px(a.x, d, c); a.y = denorm_px(a.y, d, c); a.z = denorm_px(a.z, d, c); a.w = denorm_px(a.w, d, c);
  b.x = denorm_px(b.x, d, c); b.y = denorm_px(b.y, d, c); b.z = denorm_px(b.z, d, c); b.w = denorm_px(b.w, d, c);
  down2_sum(a, b, s0, s1);
}
__device__ __forceinline__ void down2_sum(const float4 &a, const float4 &b, float &s0, float &s1) {
  float t = __fmul_rn(0.25f, a.x);
  t = __fmaf_rn(0.25f, a.y, t); t = __fmaf_rn(0.25f, b.x, t); t = __fmaf_rn(0.25f, b.y, t);
  s0 = t;
  t = __fmul_rn(0.25f, a.z);
  t = __fmaf_rn(0.25f, a.w, t); t = __fmaf_rn(0.25f, b.z, t); t = __fmaf_rn(0.25f, b.w, t);
  s1 = t;
}

__global__ void __launch_bounds__(256) cam2mask_prepare_x2_kernel(const float *__restrict__ images,
                                                                  const float *__restrict__ cams,
                                                                  const int *__restrict__ keys,
                                                                  const int *__restrict__ nc_dev,
                                                                  float *__restrict__ img_small,
                                                                  float *__restrict__ masks, MaskLayout ml, ResizeGeom g,
                                                                  int C1, float thr_high, float thr_low, int derive,
                                                                  Denorm dn, const float *__restrict__ cam_scale,
                                                                  int cstride) {
  const int xp = blockIdx.x * 32 + (threadIdx.x & 31);     // pair index: half-resolution columns 2 xp, 2 xp + 1
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  const int x = 2 * xp;
  if (x >= g.w || y >= g.h) return;
  const size_t HW = (size_t)g.H * g.W, hw = (size_t)g.h * g.w;
  const size_t pix = (size_t)y * g.w + x;
  const size_t row0 = (size_t)(2 * y) * g.W + 2 * x;       // first of the 4 source columns, upper source row
  if (img_small) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s0, s1;
      down2_pair_img(images + ((size_t)b * 3 + c) * HW, row0, g.W, s0, s1, dn, c);
      *reinterpret_cast<float2 *>(img_small + ((size_t)b * 3 + c) * hw + pix) = make_float2(s0, s1);
    }
  }
  const int nc = nc_dev[b];
  const int *key = keys + (size_t)b * (C1 + 1);
  const float *cam_b = cams + (size_t)b * C1 * HW;
  const size_t mplane = (size_t)g.h * ml.pitch;
  float *m_hi = masks + (size_t)b * cstride * mplane + (size_t)y * ml.pitch + ml.off + x;
  const int ns = nc - derive;
  float *m_lo = m_hi + (size_t)ns * mplane;
  auto put = [&](float *dst, float v0, float v1) {
    *reinterpret_cast<float2 *>(dst) = make_float2(v0, v1);
    if (ml.padn) {
      if (x == 0)
        for (int i = 1; i <= ml.padn; ++i) dst[-i] = v0;
      if (x + 2 == g.w)
        for (int i = 1; i <= ml.padn; ++i) dst[1 + i] = v1;
    }
  };
  float v0[kCacheC], v1[kCacheC];
  float mx0 = -INFINITY, mx1 = -INFINITY;
  const float *sc_b = cam_scale ? cam_scale + (size_t)b * C1 : nullptr;
  auto scale_of = [&](int j) { return sc_b ? __ldg(sc_b + key[j] - 1) : 1.0f; };
  for (int j = 1; j < nc; ++j) {
    float s0, s1;
    down2_pair(cam_b + (size_t)(key[j] - 1) * HW, row0, g.W, s0, s1, scale_of(j));
    if (j < kCacheC) { v0[j] = s0; v1[j] = s1; }
    mx0 = fmaxf(mx0, s0);
    mx1 = fmaxf(mx1, s1);
  }
  const float mh0 = fmaxf(mx0, thr_high), ml0 = fmaxf(mx0, thr_low);
  const float mh1 = fmaxf(mx1, thr_high), ml1 = fmaxf(mx1, thr_low);
  const float eh0 = expf(thr_high - mh0), el0 = expf(thr_low - ml0);
  const float eh1 = expf(thr_high - mh1), el1 = expf(thr_low - ml1);
  float dh0 = eh0, dl0 = el0, dh1 = eh1, dl1 = el1;
  for (int j = 1; j < nc; ++j) {
    float s0, s1;
    if (j < kCacheC) { s0 = v0[j]; s1 = v1[j]; }
    else down2_pair(cam_b + (size_t)(key[j] - 1) * HW, row0, g.W, s0, s1, scale_of(j));
    dh0 += expf(s0 - mh0); dl0 += expf(s0 - ml0);
    dh1 += expf(s1 - mh1); dl1 += expf(s1 - ml1);
  }
  if (ns > 0) {
    put(m_hi, eh0 / dh0, eh1 / dh1);
    put(m_lo, el0 / dl0, el1 / dl1);
  }
  for (int j = 1; j < ns; ++j) {
    float s0, s1;
    if (j < kCacheC) { s0 = v0[j]; s1 = v1[j]; }
    else down2_pair(cam_b + (size_t)(key[j] - 1) * HW, row0, g.W, s0, s1, scale_of(j));
    put(m_hi + (size_t)j * mplane, expf(s0 - mh0) / dh0, expf(s1 - mh1) / dh1);
    put(m_lo + (size_t)j * mplane, expf(s0 - ml0) / dl0, expf(s1 - ml1) / dl1);
  }
}

// argmax over nc channels of the bilinearly up-sampled stack at full-resolution pixel (Y, X)
// `total` > 0: the stack stores only its first nc - 1 channels and the last one is total - (sum of the others) at
// every half-resolution tap (see cosa_cam2mask).
__device__ __forceinline__ int upsampled_argmax(const float *stack, int nc, size_t hw, int w, const Tap &ty,
                                                const Tap &tx, float total = 0.0f) {
  // hw = plane stride, w = row pitch (the caller has already added the interior offset to `stack`)
  const size_t o00 = (size_t)ty.i0 * w + tx.i0, o01 = (size_t)ty.i0 * w + tx.i1;
  const size_t o10 = (size_t)ty.i1 * w + tx.i0, o11 = (size_t)ty.i1 * w + tx.i1;
  float best = -INFINITY;
  int arg = 0;
  float s00 = 0.0f, s01 = 0.0f, s10 = 0.0f, s11 = 0.0f;
  for (int j = 0; j < nc; ++j) {
    float a00, a01, a10, a11;
    if (total > 0.0f && j == nc - 1) {
      a00 = total - s00; a01 = total - s01; a10 = total - s10; a11 = total - s11;
    } else {
      const float *p = stack + (size_t)j * hw;
      a00 = __ldg(p + o00); a01 = __ldg(p + o01); a10 = __ldg(p + o10); a11 = __ldg(p + o11);
      s00 += a00; s01 += a01; s10 += a10; s11 += a11;
    }
    const float val = bilerp_up(ty, tx, a00, a01, a10, a11);
    if (val > best || j == 0) { best = val; arg = j; }
  }
  return arg;
}

__global__ void __launch_bounds__(256) cam2mask_finalize_kernel(const float *__restrict__ refined,
                                                                const int *__restrict__ keys,
                                                                const int *__restrict__ nc_dev,
                                                                const int *__restrict__ boxes,
                                                                float *__restrict__ label_out,
                                                                float *__restrict__ label_hi_out,
                                                                float *__restrict__ label_lo_out, MaskLayout ml,
                                                                ResizeGeom g, int C1, float ignore_index,
                                                                float derive_total, int cstride,
                                                                const int *__restrict__ err) {
  const int X = blockIdx.x * 32 + (threadIdx.x & 31);
  const int Y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (X >= g.W || Y >= g.H) return;
  const size_t HW = (size_t)g.H * g.W, mplane = (size_t)g.h * ml.pitch;
  const size_t out_idx = (size_t)b * HW + (size_t)Y * g.W + X;
  const int *box = boxes + 4 * b;
  float hi = ignore_index, lo = ignore_index;
  if (Y >= box[0] && Y < box[1] && X >= box[2] && X < box[3]) {
    const int nc = nc_dev[b];
    const int *key = keys + (size_t)b * (C1 + 1);
    // up-sampling scale = reduced / full (area_pixel_compute_scale with the output size given)
    const Tap ty = tap_half_pixel(Y, g.identity ? 1.0f : (float)g.h / (float)g.H, g.h);
    const Tap tx = tap_half_pixel(X, g.identity ? 1.0f : (float)g.w / (float)g.W, g.w);
    const float *st = refined + (size_t)b * cstride * mplane + ml.off;
    const int ns = derive_total > 0.0f ? nc - 1 : nc;   // channels stored per stack
    hi = (float)key[upsampled_argmax(st, nc, mplane, ml.pitch, ty, tx, derive_total)];
    lo = (float)key[upsampled_argmax(st + (size_t)ns * mplane, nc, mplane, ml.pitch, ty, tx, derive_total)];
  }
  // merge (seg_helper.py:781-783): high fg stays; high bg becomes ignore unless low also says bg
  float out = hi;
  if (hi == 0.0f) out = ignore_index;
  if (hi + lo == 0.0f) out = 0.0f;
  if (*err) out = hi = lo = __int_as_float(0x7fc00000);   // class budget exceeded: fail loudly
  label_out[out_idx] = out;
  if (label_hi_out) label_hi_out[out_idx] = hi;
  if (label_lo_out) label_lo_out[out_idx] = lo;
}

// The two stacks of one 2 x 2 output block: per stack the argmax over the channels of the bilinearly enlarged values.
// NC > 0: the channel count as a compile-time constant - the loop unrolls, every channel's nine loads are issued up
// front and ptxas schedules across channels (the generic loop waited on its loads channel by channel: 42 % of the
// kernel's stalls); NC = 0: run-time count.  DERIVE: the last channel of a stack is total - (sum of the others).
template <int NC, bool DERIVE>
__device__ __forceinline__ void finalize_stacks(const float *st, size_t mplane, int nc_rt, const int (&off)[3][3],
                                                const Tap (&ty)[2], const Tap (&tx)[2], float derive_total,
                                                int (&arg_hi)[2][2], int (&arg_lo)[2][2]) {
  const int nc = NC ? NC : nc_rt;
  const int nload = DERIVE ? nc - 1 : nc;            // channels of a stack that exist in memory
#pragma unroll
  for (int s = 0; s < 2; ++s) {                       // high stack, low stack
    const float *stack = st + (size_t)s * nload * mplane;
    float best[2][2];
    int arg[2][2] = {{0, 0}, {0, 0}};
    float sum[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll(NC ? NC : 1)
    for (int j = 0; j < nc; ++j) {
      const float *p = stack + (size_t)j * mplane;
      float v[3][3];
      if (DERIVE && j == nc - 1) {   // the channel that was not propagated
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int e = 0; e < 3; ++e) v[a][e] = derive_total - sum[a][e];
      } else {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int e = 0; e < 3; ++e) {
            v[a][e] = __ldg(p + off[a][e]);
            sum[a][e] += v[a][e];
          }
      }
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const float val = bilerp_up(ty[dy], tx[dx], v[dy][dx], v[dy][dx + 1], v[dy + 1][dx], v[dy + 1][dx + 1]);
          if (j == 0 || val > best[dy][dx]) { best[dy][dx] = val; arg[dy][dx] = j; }
        }
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) (s == 0 ? arg_hi : arg_lo)[dy][dx] = arg[dy][dx];
  }
}

// Exact 2x enlargement (H = 2h, W = 2w): one thread per half-resolution cell produces the 2x2 block of labels
// from ONE 3x3 neighbourhood per channel (9 loads instead of 16, index arithmetic shared by the four outputs).
// Output row 2y' interpolates source rows (y'-1, y'), row 2y'+1 rows (y', y'+1); the weights come from the same
// tap function as the general kernel (at the borders they degenerate to (1, 0), so clamped loads are exact).
__global__ void __launch_bounds__(256, 4) cam2mask_finalize_x2_kernel(const float *__restrict__ refined,
                                                                   const int *__restrict__ keys,
                                                                   const int *__restrict__ nc_dev,
                                                                   const int *__restrict__ boxes,
                                                                   float *__restrict__ label_out,
                                                                   float *__restrict__ label_hi_out,
                                                                   float *__restrict__ label_lo_out, MaskLayout ml,
                                                                   ResizeGeom g, int C1, float ignore_index,
                                                                   float derive_total, int cstride,
                                                                   const int *__restrict__ err) {
  const int xs = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ys = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (xs >= g.w || ys >= g.h) return;
  const size_t HW = (size_t)g.H * g.W, mplane = (size_t)g.h * ml.pitch;
  const int *box = boxes + 4 * b;
  const int Y0 = 2 * ys, X0 = 2 * xs;
  bool in[2][2];
  bool any = false;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      in[dy][dx] = (Y0 + dy >= box[0] && Y0 + dy < box[1] && X0 + dx >= box[2] && X0 + dx < box[3]);
      any |= in[dy][dx];
    }
  float hi[2][2], lo[2][2];
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) hi[dy][dx] = lo[dy][dx] = ignore_index;
  if (any) {
    const int nc = nc_dev[b];
    const int *key = keys + (size_t)b * (C1 + 1);
    const Tap ty[2] = {tap_half_pixel(Y0, 0.5f, g.h), tap_half_pixel(Y0 + 1, 0.5f, g.h)};
    const Tap tx[2] = {tap_half_pixel(X0, 0.5f, g.w), tap_half_pixel(X0 + 1, 0.5f, g.w)};
    const size_t r[3] = {(size_t)max(ys - 1, 0) * ml.pitch, (size_t)ys * ml.pitch, (size_t)min(ys + 1, g.h - 1) * ml.pitch};
    const int c[3] = {max(xs - 1, 0), xs, min(xs + 1, g.w - 1)};
    const float *st = refined + (size_t)b * cstride * mplane + ml.off;
    const int off[3][3] = {{(int)r[0] + c[0], (int)r[0] + c[1], (int)r[0] + c[2]},
                           {(int)r[1] + c[0], (int)r[1] + c[1], (int)r[1] + c[2]},
                           {(int)r[2] + c[0], (int)r[2] + c[1], (int)r[2] + c[2]}};
    int arg_hi[2][2], arg_lo[2][2];
    if (derive_total > 0.0f) {
      switch (nc) {
        case 2: finalize_stacks<2, true>(st, mplane, nc, off, ty, tx, derive_total, arg_hi, arg_lo); break;
        case 3: finalize_stacks<3, true>(st, mplane, nc, off, ty, tx, derive_total, arg_hi, arg_lo); break;
        case 4: finalize_stacks<4, true>(st, mplane, nc, off, ty, tx, derive_total, arg_hi, arg_lo); break;
        default: finalize_stacks<0, true>(st, mplane, nc, off, ty, tx, derive_total, arg_hi, arg_lo);
      }
    } else {
      switch (nc) {
        case 2: finalize_stacks<2, false>(st, mplane, nc, off, ty, tx, derive_total, arg_hi, arg_lo); break;
        case 3: finalize_stacks<3, false>(st, mplane, nc, off, ty, tx, derive_total, arg_hi, arg_lo); break;
        case 4: finalize_stacks<4, false>(st, mplane, nc, off, ty, tx, derive_total, arg_hi, arg_lo); break;
        default: finalize_stacks<0, false>(st, mplane, nc, off, ty, tx, derive_total, arg_hi, arg_lo);
      }
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx)
        if (in[dy][dx]) {
          hi[dy][dx] = (float)key[arg_hi[dy][dx]];
          lo[dy][dx] = (float)key[arg_lo[dy][dx]];
        }
  }
  const bool poisoned = *err != 0;      // class budget exceeded: fail loudly
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
    float o[2];
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {   // merge (seg_helper.py:781-783)
      float out = hi[dy][dx];
      if (hi[dy][dx] == 0.0f) out = ignore_index;
      if (hi[dy][dx] + lo[dy][dx] == 0.0f) out = 0.0f;
      if (poisoned) out = hi[dy][dx] = lo[dy][dx] = __int_as_float(0x7fc00000);
      o[dx] = out;
    }
    const size_t idx = (size_t)b * HW + (size_t)(Y0 + dy) * g.W + X0;
    *reinterpret_cast<float2 *>(label_out + idx) = make_float2(o[0], o[1]);
    if (label_hi_out) *reinterpret_cast<float2 *>(label_hi_out + idx) = make_float2(hi[dy][0], hi[dy][1]);
    if (label_lo_out) *reinterpret_cast<float2 *>(label_lo_out + idx) = make_float2(lo[dy][0], lo[dy][1]);
  }
}

// _refine_cams tail for callers that bring their own refined stack (seg_helper.py:793-795)
__global__ void __launch_bounds__(256) upsample_argmax_kernel(const float *__restrict__ refined,
                                                              const long long *__restrict__ valid_key,
                                                              long long *__restrict__ label_out, int nc, int h, int w,
                                                              int H, int W) {
  const int X = blockIdx.x * 32 + (threadIdx.x & 31);
  const int Y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (X >= W || Y >= H) return;
  const size_t hw = (size_t)h * w;
  const Tap ty = tap_half_pixel(Y, (float)h / (float)H, h);
  const Tap tx = tap_half_pixel(X, (float)w / (float)W, w);
  const int arg = upsampled_argmax(refined + (size_t)b * nc * hw, nc, hw, w, ty, tx);
  label_out[(size_t)b * H * W + (size_t)Y * W + X] = valid_key[arg];
}


// ------------------------------------------------------------------------------------------------
// Multi-scale CAM merge (seg_helper.py:245-270, the post-processing of multi_scale_camseg): per scale the raw
// token-grid CAMs of the image batch and of its horizontally flipped copy ([2B, C1, hs, ws]) are enlarged to
// (H, W) (bilinear, align_corners=False), max-ed with the un-flipped flipped copy, ReLU-ed, summed over the
// scales, and the sum is min-max normalised per (b, c) plane.  One thread produces 4 consecutive output pixels;
// the un-normalised sum is written once together with the per-plane min/max, then normalised in place.
// ------------------------------------------------------------------------------------------------
struct RawScales {
  const float *p[kMaxScales];
  int hs[kMaxScales], ws[kMaxScales];
  int n;
};

__device__ __forceinline__ float merged_value(const RawScales &rs, int B, int C1, int b, int c, int y, int x, int H,
                                              int W) {
  float sum = 0.0f;
  for (int s = 0; s < rs.n; ++s) {
    const int hs = rs.hs[s], ws = rs.ws[s];
    const float *a = rs.p[s] + ((size_t)b * C1 + c) * hs * ws;            // original batch
    const float *f = rs.p[s] + ((size_t)(B + b) * C1 + c) * hs * ws;      // flipped batch
    const Tap ty = tap_half_pixel(y, (float)hs / (float)H, hs);
    const Tap tx = tap_half_pixel(x, (float)ws / (float)W, ws);
    const Tap tf = tap_half_pixel(W - 1 - x, (float)ws / (float)W, ws);   // .flip(-1) after the enlargement
    const float v0 = bilerp_up(ty, tx, __ldg(a + ty.i0 * ws + tx.i0), __ldg(a + ty.i0 * ws + tx.i1),
                               __ldg(a + ty.i1 * ws + tx.i0), __ldg(a + ty.i1 * ws + tx.i1));
    const float v1 = bilerp_up(ty, tf, __ldg(f + ty.i0 * ws + tf.i0), __ldg(f + ty.i0 * ws + tf.i1),
                               __ldg(f + ty.i1 * ws + tf.i0), __ldg(f + ty.i1 * ws + tf.i1));
    const float v = fmaxf(fmaxf(v0, v1), 0.0f);                            // torch.max then F.relu
    sum = s == 0 ? v : __fadd_rn(sum, v);
  }
  return sum;
}

// Pass 1 (WRITE == false): per-plane min / max of the merged sum, nothing stored.  Pass 2 (WRITE == true): the merged
// sum is evaluated again (the raw token-grid maps are a few hundred KB and stay in L1/L2) and stored normalised, so
// the [B, C1, H, W] result crosses HBM exactly once.
template <bool WRITE>
__global__ void __launch_bounds__(256) cam_merge_kernel(RawScales rs, float *__restrict__ out, int *__restrict__ mm,
                                                        int B, int C1, int H, int W) {
  __shared__ float s_min[8], s_max[8];
  const int plane = blockIdx.y;                       // b * C1 + c
  const int b = plane / C1, c = plane - b * C1;
  const long long HW = (long long)H * W;
  float lo = INFINITY, hi = -INFINITY, neg_min = 0.0f, den = 1.0f;
  if (WRITE) {
    neg_min = -ordered_to_float(mm[2 * plane]);
    den = __fadd_rn(__fadd_rn(ordered_to_float(mm[2 * plane + 1]), neg_min), 1e-5f);
  }
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < HW;
       i += (long long)gridDim.x * blockDim.x * 4) {
    float v[4];
    const int y = (int)(i / W), x0 = (int)(i % W);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int yy = y, xx = x0 + k;
      if (xx >= W) { yy += xx / W; xx %= W; }
      v[k] = (i + k < HW) ? merged_value(rs, B, C1, b, c, yy, xx, H, W) : 0.0f;
      if (i + k < HW) { lo = fminf(lo, v[k]); hi = fmaxf(hi, v[k]); }
    }
    if (WRITE) {
      float *o = out + (size_t)plane * HW + i;
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = __fdiv_rn(__fadd_rn(v[k], neg_min), den);
      if (i + 3 < HW && (HW & 3) == 0) stg_stream4(o, make_float4(v[0], v[1], v[2], v[3]));
      else
        for (int k = 0; k < 4 && i + k < HW; ++k) o[k] = v[k];
    }
  }
  if (!WRITE) {
    lo = warp_min(lo);
    hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = lo; s_max[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x < 32) {
      lo = threadIdx.x < 8 ? s_min[threadIdx.x] : INFINITY;
      hi = threadIdx.x < 8 ? s_max[threadIdx.x] : -INFINITY;
      lo = warp_min(lo);
      hi = warp_max(hi);
      if (threadIdx.x == 0) {
        atomicMin(mm + 2 * plane, float_to_ordered(lo));
        atomicMax(mm + 2 * plane + 1, float_to_ordered(hi));
      }
    }
  }
}

// Row-walking form of the merge for W % 4 == 0 and up to 5 scales.  A thread owns 4 adjacent output columns of one
// plane and walks down a chunk of rows; per scale it keeps the x-interpolated values of its columns on the two
// source rows that bracket the current output row (plain and flipped copy) and refreshes them only when the source
// row changes (every H/hs output rows), so an output pixel costs one y-interpolation per scale and copy instead of
// two full bilinear evaluations.  Arithmetic order per value is torch's (x-lerp, then y-lerp), as in merged_value.
//   MODE 0: cam, min/max only      MODE 1: cam, normalised store      MODE 2: seg (sum of plain + flipped), store
//   MODE 3: cam, min/max AND the un-normalised sum stored (cam_normalize_inplace_kernel finishes the plane)
template <int NS, int MODE>
__global__ void __launch_bounds__(128) merge_rows_kernel(RawScales rs, float *__restrict__ out, int *__restrict__ mm,
                                                         int B, int C, int H, int W, int rows_per_chunk,
                                                         const float *__restrict__ cls) {
  const int wq = W >> 2, n_chunks = (H + rows_per_chunk - 1) / rows_per_chunk;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)B * C * n_chunks * wq;
  const bool live = t < total;
  const int xq = live ? (int)(t % wq) : 0;
  const int chunk = live ? (int)((t / wq) % n_chunks) : 0;
  const int plane = live ? (int)(t / ((long long)wq * n_chunks)) : 0;
  const int b = plane / C, c = plane - b * C;
  const int x = xq << 2;
  const int y_begin = chunk * rows_per_chunk;
  int y_end = live ? min(H, y_begin + rows_per_chunk) : y_begin;
  // fused cam_validation (seg_helper.py:547-551): the plane is multiplied by its class label; a plane whose label
  // is 0 is neither merged nor scanned for its extrema, only zero-filled
  const float lab = (cls && live) ? __ldg(cls + plane) : 1.0f;
  if (lab == 0.0f) {
    if (MODE == 1) {
      float *z = out + (size_t)plane * H * W + x;
      for (int y = y_begin; y < y_end; ++y) stg_stream4(z + (size_t)y * W, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    y_end = y_begin;
  }
  float neg_min = 0.0f, den = 1.0f;
  if (MODE == 1) {
    neg_min = -ordered_to_float(mm[2 * plane]);
    den = __fadd_rn(__fadd_rn(ordered_to_float(mm[2 * plane + 1]), neg_min), 1e-5f);
  }
  float top[NS][4], bot[NS][4], ftop[NS][4], fbot[NS][4], sy[NS];
  int r0[NS], r1[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) { r0[s] = -1; r1[s] = -1; sy[s] = (float)rs.hs[s] / (float)H; }
  float lo = INFINITY, hi = -INFINITY;
  float *o = out + (size_t)plane * H * W + x;
  for (int y = y_begin; y < y_end; ++y) {
    float acc[4];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int hs = rs.hs[s], ws = rs.ws[s];
      const Tap ty = tap_half_pixel(y, sy[s], hs);
      if (ty.i0 != r0[s] || ty.i1 != r1[s]) {      // new source rows: x-interpolate this thread's columns on them
        const float *a = rs.p[s] + ((size_t)b * C + c) * hs * ws;
        const float *f = rs.p[s] + ((size_t)(B + b) * C + c) * hs * ws;
        const float sx = (float)ws / (float)W;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const Tap tx = tap_half_pixel(x + k, sx, ws), tf = tap_half_pixel(W - 1 - (x + k), sx, ws);
          top[s][k] = lerp_nested(tx.w0, __ldg(a + ty.i0 * ws + tx.i0), tx.w1, __ldg(a + ty.i0 * ws + tx.i1));
          bot[s][k] = lerp_nested(tx.w0, __ldg(a + ty.i1 * ws + tx.i0), tx.w1, __ldg(a + ty.i1 * ws + tx.i1));
          ftop[s][k] = lerp_nested(tf.w0, __ldg(f + ty.i0 * ws + tf.i0), tf.w1, __ldg(f + ty.i0 * ws + tf.i1));
          fbot[s][k] = lerp_nested(tf.w0, __ldg(f + ty.i1 * ws + tf.i0), tf.w1, __ldg(f + ty.i1 * ws + tf.i1));
        }
        r0[s] = ty.i0;
        r1[s] = ty.i1;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float v0 = lerp_nested(ty.w0, top[s][k], ty.w1, bot[s][k]);
        const float v1 = lerp_nested(ty.w0, ftop[s][k], ty.w1, fbot[s][k]);
        const float v = MODE == 2 ? __fadd_rn(v0, v1) : fmaxf(fmaxf(v0, v1), 0.0f);
        acc[k] = s == 0 ? v : __fadd_rn(acc[k], v);
      }
    }
    if (MODE == 0 || MODE == 3) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { lo = fminf(lo, acc[k]); hi = fmaxf(hi, acc[k]); }
      if (MODE == 3) *reinterpret_cast<float4 *>(o + (size_t)y * W) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
      if (MODE == 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = __fmul_rn(lab, __fdiv_rn(__fadd_rn(acc[k], neg_min), den));
      }
      stg_stream4(o + (size_t)y * W, make_float4(acc[0], acc[1], acc[2], acc[3]));
    }
  }
  if (MODE == 0 || MODE == 3) {
    // one pair of atomics per warp when the whole warp works on one plane, else one per thread
    const int p0 = __shfl_sync(0xffffffffu, plane, 0);
    const bool uniform = __all_sync(0xffffffffu, plane == p0 && live);
    if (uniform) {
      lo = warp_min(lo);
      hi = warp_max(hi);
      if ((threadIdx.x & 31) == 0) {
        atomicMin(mm + 2 * plane, float_to_ordered(lo));
        atomicMax(mm + 2 * plane + 1, float_to_ordered(hi));
      }
    } else if (live && y_end > y_begin) {
      atomicMin(mm + 2 * plane, float_to_ordered(lo));
      atomicMax(mm + 2 * plane + 1, float_to_ordered(hi));
    }
  }
}

// Second pass of the CAM merge: the plane holds the un-normalised scale sum (merge_rows_kernel<.., 3>), the extrema
// are known; one streaming read-modify-write finishes it (and applies the class label of the fused cam_validation).
// The bilinear merge is then evaluated once per pixel instead of twice (once for the extrema, once for the store).
// present_only: planes whose label is 0 are left untouched and no label factor is applied (the caller hands the
// result to cam2mask as un-validated CAMs: cosa_multi_scale_cam_merge_present).
__global__ void __launch_bounds__(256) cam_normalize_inplace_kernel(float *__restrict__ out, const int *__restrict__ mm,
                                                                    const float *__restrict__ cls, long long HW4,
                                                                    long long HW, int planes, int present_only) {
  for (int p = blockIdx.y; p < planes; p += gridDim.y) {
    float lab = cls ? __ldg(cls + p) : 1.0f;
    float *dst = out + (size_t)p * HW;
    if (present_only) {
      if (lab == 0.0f) continue;
      lab = 1.0f;
    }
    if (lab == 0.0f) {
      for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW4;
           i += (long long)gridDim.x * blockDim.x)
        stg_stream4(dst + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
      continue;
    }
    const float neg_min = -ordered_to_float(mm[2 * p]);
    const float den = __fadd_rn(__fadd_rn(ordered_to_float(mm[2 * p + 1]), neg_min), 1e-5f);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW4;
         i += (long long)gridDim.x * blockDim.x) {
      float4 v = *reinterpret_cast<const float4 *>(dst + 4 * i);
      v.x = __fmul_rn(lab, __fdiv_rn(__fadd_rn(v.x, neg_min), den));
      v.y = __fmul_rn(lab, __fdiv_rn(__fadd_rn(v.y, neg_min), den));
      v.z = __fmul_rn(lab, __fdiv_rn(__fadd_rn(v.z, neg_min), den));
      v.w = __fmul_rn(lab, __fdiv_rn(__fadd_rn(v.w, neg_min), den));
      stg_stream4(dst + 4 * i, v);
    }
  }
}

template <int MODE>
static int launch_merge_rows(const RawScales &rs, float *out, int *mm, int B, int C, int H, int W, const float *cls,
                             cudaStream_t s) {
  const int rows = 56;   // rows per thread: long enough to amortise the source-row refresh, short enough to fill the GPU
  const long long total = (long long)B * C * ceil_div(H, rows) * (W / 4);
  const unsigned grid = (unsigned)ceil_div_ll(total, 128);
  const char *name = MODE == 0 ? "cam_merge_minmax_kernel"
                     : (MODE == 1 ? "cam_merge_write_kernel" : (MODE == 2 ? "seg_merge_kernel" : "cam_merge_sum_kernel"));
  switch (rs.n) {
    case 1: { auto k = merge_rows_kernel<1, MODE>; COSA_LAUNCH_T(name, k, grid, 128, 0, s, rs, out, mm, B, C, H, W, rows, cls); break; }
    case 2: { auto k = merge_rows_kernel<2, MODE>; COSA_LAUNCH_T(name, k, grid, 128, 0, s, rs, out, mm, B, C, H, W, rows, cls); break; }
    case 3: { auto k = merge_rows_kernel<3, MODE>; COSA_LAUNCH_T(name, k, grid, 128, 0, s, rs, out, mm, B, C, H, W, rows, cls); break; }
    case 4: { auto k = merge_rows_kernel<4, MODE>; COSA_LAUNCH_T(name, k, grid, 128, 0, s, rs, out, mm, B, C, H, W, rows, cls); break; }
    case 5: { auto k = merge_rows_kernel<5, MODE>; COSA_LAUNCH_T(name, k, grid, 128, 0, s, rs, out, mm, B, C, H, W, rows, cls); break; }
    default: return COSA_E_ARG;
  }
  return 0;
}

// seg = sum_s ( up(seg_s[:B]) + up(seg_s[B:]).flip(-1) )     (seg_helper.py:260-262, :273)
__global__ void __launch_bounds__(256) seg_merge_kernel(RawScales rs, float *__restrict__ out, int B, int C, int H,
                                                        int W) {
  const long long total = (long long)B * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long plane = i / ((long long)W * H);
    const int b = (int)(plane / C), c = (int)(plane % C);
    float acc = 0.0f;
    for (int s = 0; s < rs.n; ++s) {
      const int hs = rs.hs[s], ws = rs.ws[s];
      const float *a = rs.p[s] + ((size_t)b * C + c) * hs * ws;
      const float *f = rs.p[s] + ((size_t)(B + b) * C + c) * hs * ws;
      const Tap ty = tap_half_pixel(y, (float)hs / (float)H, hs);
      const Tap tx = tap_half_pixel(x, (float)ws / (float)W, ws);
      const Tap tf = tap_half_pixel(W - 1 - x, (float)ws / (float)W, ws);
      const float v0 = bilerp_up(ty, tx, __ldg(a + ty.i0 * ws + tx.i0), __ldg(a + ty.i0 * ws + tx.i1),
                                 __ldg(a + ty.i1 * ws + tx.i0), __ldg(a + ty.i1 * ws + tx.i1));
      const float v1 = bilerp_up(ty, tf, __ldg(f + ty.i0 * ws + tf.i0), __ldg(f + ty.i0 * ws + tf.i1),
                                 __ldg(f + ty.i1 * ws + tf.i0), __ldg(f + ty.i1 * ws + tf.i1));
      const float v = __fadd_rn(v0, v1);
      acc = s == 0 ? v : __fadd_rn(acc, v);
    }
    out[i] = acc;
  }
}

}  // namespace cosa

using namespace cosa;

extern "C" int cosa_cam_validation(const float *cam, const float *cls_label, float *out, int B, int C1, long long HW,
                                   void *stream) {
  if (!cam || !cls_label || !out || B < 1 || C1 < 1 || HW < 1) return COSA_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const int planes = B * C1;
  const bool vec = (HW % 4 == 0) && (((uintptr_t)cam | (uintptr_t)out) % 16 == 0);
  if (vec) {
    const long long HW4 = HW / 4;
    const int bx = (int)min(ceil_div_ll(HW4, 256), 64LL);
    const int by = min(planes, max(1, sm_count() * 16 / bx));
    COSA_LAUNCH(cam_validation_kernel, dim3(bx, by), 256, 0, s, cam, cls_label, out, HW4, HW, planes);
  } else {
    const long long total = (long long)planes * HW;
    const int blocks = (int)min((long long)sm_count() * 8, ceil_div_ll(total, 256));
    COSA_LAUNCH(cam_validation_scalar_kernel, blocks, 256, 0, s, cam, cls_label, out, HW, total);
  }
  return 0;
}

extern "C" int cosa_denormalize_img(const float *imgs, float *out, int B, long long HW, const float mean[3],
                                    const float std[3], void *stream) {
  if (!imgs || !out || !mean || !std || B < 1 || HW < 1) return COSA_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  Affine3f a;
  for (int c = 0; c < 3; ++c) { a.mean[c] = mean[c]; a.std[c] = std[c]; }
  const int planes = 3 * B;
  if ((HW % 4 == 0) && (((uintptr_t)imgs | (uintptr_t)out) % 16 == 0)) {
    const long long HW4 = HW / 4;
    const int bx = (int)min(ceil_div_ll(HW4, 256), 64LL);
    const int by = min(planes, max(1, sm_count() * 16 / bx));
    COSA_LAUNCH(denormalize_img_kernel, dim3(bx, by), 256, 0, s, imgs, out, a, HW4, HW, planes);
  } else {
    const long long total = (long long)planes * HW;
    const int blocks = (int)min((long long)sm_count() * 8, ceil_div_ll(total, 256));
    COSA_LAUNCH(denormalize_img_scalar_kernel, blocks, 256, 0, s, imgs, out, a, HW, total);
  }
  return 0;
}

extern "C" int cosa_cam_normalize(const float *const *scale_maps, int n_scales, float *out, int planes, long long HW,
                                  float *minmax_ws, void *stream) {
  if (!scale_maps || n_scales < 1 || n_scales > kMaxScales || !out || !minmax_ws || planes < 1 || HW < 1)
    return COSA_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  ScalePtrs sp;
  for (int k = 0; k < kMaxScales; ++k) sp.p[k] = k < n_scales ? scale_maps[k] : nullptr;
  int *mm = (int *)minmax_ws;
  COSA_LAUNCH(minmax_init_kernel, ceil_div(planes, 256), 256, 0, s, mm, planes);
  const int bx = (int)min(ceil_div_ll(HW, 256 * 8), 32LL);
  const int by = min(planes, max(1, sm_count() * 8 / bx));
  COSA_LAUNCH(cam_minmax_kernel, dim3(bx, by), 256, 0, s, sp, n_scales, mm, HW, planes);
  COSA_LAUNCH(cam_normalize_apply_kernel, dim3(bx, by), 256, 0, s, sp, n_scales, mm, out, HW, planes);
  return 0;
}

extern "C" int cosa_cam_to_label(const float *cam, const float *cls_label, const int *boxes, float *valid_cam_out,
                                 long long *label_out, int B, int C1, int H, int W, float bkg_thre, float high_thre,
                                 float low_thre, int ignore_mid, long long ignore_index, void *stream) {
  if (!cam || !label_out || B < 1 || C1 < 1 || H < 1 || W < 1) return COSA_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  LabelArgs a{bkg_thre, high_thre, low_thre, ignore_mid, ignore_index};
  const long long HW = (long long)H * W;
  const long long warps = ceil_div_ll(HW, 32) * B;
  const int blocks = (int)min((long long)sm_count() * 8, ceil_div_ll(warps, 8));
  const bool vec = (HW % 4 == 0) && ((uintptr_t)cam % 16 == 0) && (!valid_cam_out || (uintptr_t)valid_cam_out % 16 == 0);
  if (vec) {
    COSA_LAUNCH(cam_to_label_kernel<true>, blocks, 256, 0, s, cam, cls_label, boxes, valid_cam_out, label_out, B, C1,
                H, W, a);
  } else {
    COSA_LAUNCH(cam_to_label_kernel<false>, blocks, 256, 0, s, cam, cls_label, boxes, valid_cam_out, label_out, B, C1,
                H, W, a);
  }
  return 0;
}

static void cam2mask_geom(int H, int W, int downscale, ResizeGeom *g) {
  g->H = H; g->W = W;
  g->identity = downscale == 0;
  g->h = downscale ? H / downscale : H;
  g->w = downscale ? W / downscale : W;
  g->sy = g->identity ? 1.0f : (float)H / (float)g->h;
  g->sx = g->identity ? 1.0f : (float)W / (float)g->w;
}

// mask rows of the PAR path are padded for the vectorised kernel; pads wider than 24 columns are not used
// (cosa_cam2mask then keeps the plain layout), which bounds the scratch without knowing the dilation values
static MaskLayout cam2mask_layout(int w, int use_par, const int *dilations, int n_dil) {
  if (!use_par) return plain_layout(w);
  MaskLayout l = padded_layout(w, dilations, n_dil);
  return l.padn > 24 ? plain_layout(w) : l;
}

// mask planes per image: both threshold stacks of (1 + present classes) channels; COSA_CAM2MASK_MAX_CLASSES caps the
// present classes the buffers are sized for (0 = all C1)
static int cam2mask_class_cap(int C1, int flags) {
  const int cap = (flags >> 8) & 0xff;
  return cap > 0 ? min(cap, C1) : C1;
}

extern "C" size_t cosa_cam2mask_ws_bytes(int B, int C1, int H, int W, int downscale, int use_par, int n_dil) {
  return cosa_cam2mask_ws_bytes_ex(B, C1, H, W, downscale, use_par, n_dil, 0);
}

extern "C" size_t cosa_cam2mask_ws_bytes_ex(int B, int C1, int H, int W, int downscale, int use_par, int n_dil,
                                            int flags) {
  ResizeGeom g;
  cam2mask_geom(H, W, downscale, &g);
  const size_t hw = (size_t)g.h * g.w;
  const size_t pitch = use_par ? (size_t)max_padded_pitch(g.w) : (size_t)g.w;
  const int cstride = 2 * (cam2mask_class_cap(C1, flags) + 1);
  size_t bytes = align_up((size_t)B * (C1 + 1) * sizeof(int), 256) + 3 * align_up((size_t)B * sizeof(int), 256);
  const int n_mask_bufs = use_par ? 4 : 1;
  bytes += n_mask_bufs * align_up((size_t)B * cstride * g.h * pitch * sizeof(float), 256);
  if (use_par) {
    bytes += align_up((size_t)B * 3 * hw * sizeof(float), 256);
    bytes += align_up(par_affinity_floats(B, n_dil, g.h, g.w) * sizeof(float), 256);
    bytes += align_up(par_tile_flag_ints(B, g.h, g.w) * sizeof(int), 256);
  }
  return bytes;
}

extern "C" int cosa_cam2mask(const float *images, const int *boxes, const float *cams, const float *cls_labels,
                             float threshold_high, float threshold_low, float ignore_index, int downscale, int use_par,
                             const int *dilations, int n_dil, int num_iter, float *label_out, float *label_high_out,
                             float *label_low_out, int B, int C1, int H, int W, void *ws, size_t ws_bytes,
                             void *stream) {
  return cosa_cam2mask_flags(images, boxes, cams, cls_labels, threshold_high, threshold_low, ignore_index, downscale,
                             use_par, dilations, n_dil, num_iter, label_out, label_high_out, label_low_out, B, C1, H, W,
                             ws, ws_bytes, 0, stream);
}

extern "C" int cosa_cam2mask_flags(const float *images, const int *boxes, const float *cams, const float *cls_labels,
                                   float threshold_high, float threshold_low, float ignore_index, int downscale,
                                   int use_par, const int *dilations, int n_dil, int num_iter, float *label_out,
                                   float *label_high_out, float *label_low_out, int B, int C1, int H, int W, void *ws,
                                   size_t ws_bytes, int flags, void *stream) {
  return cosa_cam2mask_ex(images, boxes, cams, cls_labels, threshold_high, threshold_low, ignore_index, downscale,
                          use_par, dilations, n_dil, num_iter, label_out, label_high_out, label_low_out, B, C1, H, W,
                          ws, ws_bytes, flags, nullptr, nullptr, stream);
}

extern "C" int cosa_cam2mask_ex(const float *images, const int *boxes, const float *cams, const float *cls_labels,
                                float threshold_high, float threshold_low, float ignore_index, int downscale,
                                int use_par, const int *dilations, int n_dil, int num_iter, float *label_out,
                                float *label_high_out, float *label_low_out, int B, int C1, int H, int W, void *ws,
                                size_t ws_bytes, int flags, const float *denorm_mean, const float *denorm_std,
                                void *stream) {
  if (flags & ~(COSA_CAM2MASK_REUSE_AFFINITY | COSA_CAM2MASK_ALL_CHANNELS | COSA_CAM2MASK_CAMS_UNVALIDATED |
                COSA_CAM2MASK_MAX_CLASSES_MASK))
    return COSA_E_ARG;
  if ((denorm_mean == nullptr) != (denorm_std == nullptr)) return COSA_E_ARG;
  if (!images || !boxes || !cams || !cls_labels || !label_out || !ws || B < 1 || C1 < 1 || H < 1 || W < 1 ||
      downscale < 0)
    return COSA_E_ARG;
  if (use_par && (!dilations || n_dil < 1 || num_iter < 0)) return COSA_E_ARG;
  ResizeGeom g;
  cam2mask_geom(H, W, downscale, &g);
  if (g.h < 1 || g.w < 1) return COSA_E_ARG;
  if (ws_bytes < cosa_cam2mask_ws_bytes_ex(B, C1, H, W, downscale, use_par, n_dil, flags)) return COSA_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t hw = (size_t)g.h * g.w;
  const int C = C1 + 1;
  const bool refine = use_par && num_iter > 0;
  const MaskLayout lay = cam2mask_layout(g.w, refine, dilations, n_dil);
  const int class_cap = cam2mask_class_cap(C1, flags);
  const int cstride = 2 * (class_cap + 1);
  const size_t mfloats = layout_floats(lay, B, cstride, g.h);
  Arena arena(ws);
  int *keys = arena.take<int>((size_t)B * C);
  int *nc = arena.take<int>(B);
  int *nch = arena.take<int>(B);
  int *err = arena.take<int>(B);   // [0]: an image had more present classes than the buffers were sized for
  float *masks = arena.take<float>(mfloats);

  // The stacks PAR refines are softmax outputs: their channels sum to 1 at every pixel, and one propagation step
  // multiplies that sum by the row sum of the weights (the same 8*n_dil taps at every pixel, replicate padding
  // included).  So the last live channel of each stack is not propagated: after num_iter steps it is
  // row_sum^num_iter minus the sum of the others, evaluated by the labelling kernel at every tap (|error| ~ 1e-6,
  // the same order as the summation-order differences between two fp32 evaluations of the reference).
  // COSA_CAM2MASK_ALL_CHANNELS (a per-call flag) propagates every channel instead.
  ParConst pc;
  pc.row_sum = 1.0;
  if (refine) COSA_CHECK(par_make_constants(dilations, n_dil, &pc));
  const int derive = (refine && !(flags & COSA_CAM2MASK_ALL_CHANNELS)) ? 1 : 0;
  Denorm dn;
  dn.on = denorm_mean != nullptr;
  for (int c = 0; c < 3; ++c) { dn.mean[c] = dn.on ? denorm_mean[c] : 0.0f; dn.std[c] = dn.on ? denorm_std[c] : 1.0f; }
  const float *cam_scale = (flags & COSA_CAM2MASK_CAMS_UNVALIDATED) ? cls_labels : nullptr;
  const float derive_total = derive ? (float)pow(pc.row_sum, (double)num_iter) : 0.0f;
  COSA_CUDA(cudaMemsetAsync(err, 0, sizeof(int), s));
  COSA_LAUNCH(cam2mask_keys_kernel, ceil_div(B, 64), 64, 0, s, cls_labels, keys, nc, nch, B, C1, derive, class_cap, err);
  float *img_small = nullptr, *aff = nullptr, *sa = nullptr, *sb = nullptr, *fin = nullptr;
  int *tile_flags = nullptr;
  if (refine) {
    sa = arena.take<float>(mfloats);
    sb = arena.take<float>(mfloats);
    fin = arena.take<float>(mfloats);
    img_small = arena.take<float>((size_t)B * 3 * hw);
    aff = arena.take<float>(par_affinity_floats(B, n_dil, g.h, g.w));
    tile_flags = arena.take<int>(par_tile_flag_ints(B, g.h, g.w));
  }
  // COSA_CAM2MASK_REUSE_AFFINITY: the previous call on this workspace had the same images, geometry and dilations
  // (main.py:158 and :191 label the CAMs and the auxiliary CAMs of one batch): its affinity planes are still in the
  // workspace, so neither the reduced image nor the affinity is computed again.
  const bool reuse_aff = refine && (flags & COSA_CAM2MASK_REUSE_AFFINITY);
  const bool exact2 = !g.identity && H == 2 * g.h && W == 2 * g.w && W % 4 == 0 && (lay.pitch % 2) == 0 &&
                      (lay.off % 2) == 0 && (((uintptr_t)images | (uintptr_t)cams) % 16) == 0;
  if (exact2) {
    dim3 gs(ceil_div(g.w / 2, 32), ceil_div(g.h, 8), B);
    COSA_LAUNCH_T("cam2mask_prepare_kernel", cam2mask_prepare_x2_kernel, gs, 256, 0, s, images, cams, keys, nc,
                  reuse_aff ? nullptr : img_small, masks, lay, g, C1, threshold_high, threshold_low, derive, dn, cam_scale,
                  cstride);
  } else {
    dim3 gs(ceil_div(g.w, 32), ceil_div(g.h, 8), B);
    COSA_LAUNCH(cam2mask_prepare_kernel, gs, 256, 0, s, images, cams, keys, nc, reuse_aff ? nullptr : img_small, masks,
                lay, g, C1, threshold_high, threshold_low, derive, dn, cam_scale, cstride);
  }
  const float *refined = masks;
  MaskLayout lay_fin = lay;
  lay_fin.padn = 0;   // the labelling kernel never reads the pads
  if (refine) {
    if (reuse_aff) {
      COSA_CHECK(par_launch_iterations(pc, aff, masks, sa, sb, lay, fin, lay_fin, nch, 0, cstride, B, g.h, g.w, num_iter,
                                       tile_flags, s));
    } else {
      COSA_CHECK(par_refine_batch(pc, img_small, aff, masks, sa, sb, lay, fin, lay_fin, nch, 0, cstride, B, g.h, g.w,
                                  num_iter, tile_flags, s));
    }
    refined = fin;
  }
  if (!g.identity && H == 2 * g.h && W == 2 * g.w) {
    dim3 gf(ceil_div(g.w, 32), ceil_div(g.h, 8), B);
    COSA_LAUNCH(cam2mask_finalize_x2_kernel, gf, 256, 0, s, refined, keys, nc, boxes, label_out, label_high_out,
                label_low_out, lay_fin, g, C1, ignore_index, derive_total, cstride, err);
  } else {
    dim3 gf(ceil_div(W, 32), ceil_div(H, 8), B);
    COSA_LAUNCH(cam2mask_finalize_kernel, gf, 256, 0, s, refined, keys, nc, boxes, label_out, label_high_out,
                label_low_out, lay_fin, g, C1, ignore_index, derive_total, cstride, err);
  }
  return 0;
}

extern "C" int cosa_upsample_argmax(const float *refined, const long long *valid_key, long long *label_out, int B,
                                    int nc, int h, int w, int H, int W, void *stream) {
  if (!refined || !valid_key || !label_out || B < 1 || nc < 1 || h < 1 || w < 1 || H < 1 || W < 1) return COSA_E_ARG;
  dim3 gf(ceil_div(W, 32), ceil_div(H, 8), B);
  COSA_LAUNCH(upsample_argmax_kernel, gf, 256, 0, (cudaStream_t)stream, refined, valid_key, label_out, nc, h, w, H, W);
  return 0;
}

static int fill_raw_scales(RawScales *rs, const float *const *raw, const int *hs, const int *ws, int n_scales) {
  if (!raw || !hs || !ws || n_scales < 1 || n_scales > kMaxScales) return COSA_E_ARG;
  rs->n = n_scales;
  for (int k = 0; k < kMaxScales; ++k) {
    rs->p[k] = k < n_scales ? raw[k] : nullptr;
    rs->hs[k] = k < n_scales ? hs[k] : 1;
    rs->ws[k] = k < n_scales ? ws[k] : 1;
    if (k < n_scales && (!raw[k] || hs[k] < 1 || ws[k] < 1)) return COSA_E_ARG;
  }
  return 0;
}

static int cam_merge_impl(const float *const *raw, const int *hs, const int *ws, int n_scales, const float *cls_label,
                          float *out, int B, int C1, int H, int W, float *minmax_ws, void *stream,
                          int present_only = 0) {
  if (!out || !minmax_ws || B < 1 || C1 < 1 || H < 1 || W < 1) return COSA_E_ARG;
  if (present_only && !(W % 4 == 0 && n_scales <= 5 && ((uintptr_t)out % 16) == 0)) return COSA_E_ARG;
  RawScales rs;
  COSA_CHECK(fill_raw_scales(&rs, raw, hs, ws, n_scales));
  cudaStream_t s = (cudaStream_t)stream;
  const int planes = B * C1;
  const long long HW = (long long)H * W;
  int *mm = (int *)minmax_ws;
  COSA_LAUNCH(minmax_init_kernel, ceil_div(planes, 256), 256, 0, s, mm, planes);
  if (W % 4 == 0 && n_scales <= 5) {
    if (((uintptr_t)out % 16) != 0) {   // unaligned output: evaluate the merge twice (extrema, then store)
      COSA_CHECK(launch_merge_rows<0>(rs, out, mm, B, C1, H, W, cls_label, s));
      return launch_merge_rows<1>(rs, out, mm, B, C1, H, W, cls_label, s);
    }
    COSA_CHECK(launch_merge_rows<3>(rs, out, mm, B, C1, H, W, cls_label, s));
    const int bx = (int)min(ceil_div_ll(HW / 4, 256), 64LL);
    const int by = min(planes, max(1, sm_count() * 16 / bx));
    COSA_LAUNCH(cam_normalize_inplace_kernel, dim3(bx, by), 256, 0, s, out, mm, cls_label, HW / 4, HW, planes,
                present_only);
    return 0;
  }
  const int bx = (int)max(1LL, min(ceil_div_ll(HW, 256 * 4 * 4), 32LL));
  COSA_LAUNCH_T("cam_merge_minmax_kernel", cam_merge_kernel<false>, dim3(bx, planes), 256, 0, s, rs, out, mm, B, C1, H, W);
  COSA_LAUNCH_T("cam_merge_write_kernel", cam_merge_kernel<true>, dim3(bx, planes), 256, 0, s, rs, out, mm, B, C1, H, W);
  if (cls_label) return cosa_cam_validation(out, cls_label, out, B, C1, HW, stream);   // generic widths: in place
  return 0;
}

extern "C" int cosa_multi_scale_cam_merge(const float *const *raw, const int *hs, const int *ws, int n_scales,
                                          float *out, int B, int C1, int H, int W, float *minmax_ws, void *stream) {
  return cam_merge_impl(raw, hs, ws, n_scales, nullptr, out, B, C1, H, W, minmax_ws, stream);
}

extern "C" int cosa_multi_scale_cam_merge_valid(const float *const *raw, const int *hs, const int *ws, int n_scales,
                                                const float *cls_label, float *out, int B, int C1, int H, int W,
                                                float *minmax_ws, void *stream) {
  if (!cls_label) return COSA_E_ARG;
  return cam_merge_impl(raw, hs, ws, n_scales, cls_label, out, B, C1, H, W, minmax_ws, stream);
}

extern "C" int cosa_multi_scale_cam_merge_present(const float *const *raw, const int *hs, const int *ws, int n_scales,
                                                  const float *cls_label, float *out, int B, int C1, int H, int W,
                                                  float *minmax_ws, void *stream) {
  if (!cls_label) return COSA_E_ARG;
  return cam_merge_impl(raw, hs, ws, n_scales, cls_label, out, B, C1, H, W, minmax_ws, stream, 1);
}

extern "C" int cosa_multi_scale_seg_merge(const float *const *raw, const int *hs, const int *ws, int n_scales,
                                          float *out, int B, int C, int H, int W, void *stream) {
  if (!out || B < 1 || C < 1 || H < 1 || W < 1) return COSA_E_ARG;
  RawScales rs;
  COSA_CHECK(fill_raw_scales(&rs, raw, hs, ws, n_scales));
  if (W % 4 == 0 && n_scales <= 5) return launch_merge_rows<2>(rs, out, nullptr, B, C, H, W, nullptr, (cudaStream_t)stream);
  const long long total = (long long)B * C * H * W;
  const int blocks = (int)max(1LL, min((long long)sm_count() * 16, ceil_div_ll(total, 256)));
  COSA_LAUNCH(seg_merge_kernel, blocks, 256, 0, (cudaStream_t)stream, rs, out, B, C, H, W);
  return 0;
}
