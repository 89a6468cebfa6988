// Bilinear enlargement of the segmentation logits and its adjoint.
//
// Reference: main.py:167   seg_pred = F.interpolate(seg_pred, size=label.shape[1:], mode='bilinear',
//                                                   align_corners=False)
// The decoder emits the logits on the ViT token grid (28 x 28 for a 448^2 crop); this step brings them to the
// label resolution before seg_loss (seg_helper.py:800-813) and get_energy_loss (:210-230) read them, and autograd
// carries the gradient of both losses back through it.
//
//   upsample_bilinear_kernel        out[p, Y, X] = sum_{i,j} wy(Y,i) wx(X,j) in[p, i, j]        (forward)
//   upsample_adjoint_rows_kernel    tmp[p, k, 0|1, X] = sum over the rows Y with upper tap k of (w0 | w1)(Y) g[p, Y, X]
//   upsample_adjoint_cols_kernel    gin[p, i, j] = sum_X wx(X,j) (tmp[p,i,0,X] + tmp[p,i-1,1,X])          (backward)
//
// Forward arithmetic is torch's (common.cuh: source index scale*(dst+0.5)-0.5 clamped at 0, x-lerp then y-lerp as
// fma(w0, a, rn(w1*b))): bit-exact against the CPU kernel for integer ratios (the path's 16x; dyadic weights), within
// an ulp otherwise.  The adjoint is a gather in two separable passes - pass 1
// streams the full-resolution gradient exactly once with coalesced 128-bit loads, pass 2 works on the H/(2h)-times
// smaller intermediate - instead of the scatter with 4 atomics per element the definition suggests; only its
// summation order differs from torch's.
#include "common.cuh"

namespace cosa {

__global__ void __launch_bounds__(256) upsample_bilinear_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                                long long planes, int h, int w, int H, int W,
                                                                float sy, float sx, int vec) {
  const int Wq = vec ? W >> 2 : W;
  const long long total = planes * H * Wq;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(idx % Wq);
    const long long t = idx / Wq;
    const int Y = (int)(t % H);
    const long long p = t / H;
    const Tap ty = tap_half_pixel(Y, sy, h);
    const float *r0 = in + (size_t)p * h * w + (size_t)ty.i0 * w;
    const float *r1 = in + (size_t)p * h * w + (size_t)ty.i1 * w;
    float *dst = out + ((size_t)p * H + Y) * W;
    if (vec) {
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const Tap tx = tap_half_pixel(4 * xq + k, sx, w);
        v[k] = bilerp_up(ty, tx, __ldg(r0 + tx.i0), __ldg(r0 + tx.i1), __ldg(r1 + tx.i0), __ldg(r1 + tx.i1));
      }
      stg_stream4(dst + 4 * xq, make_float4(v[0], v[1], v[2], v[3]));
    } else {
      const Tap tx = tap_half_pixel(xq, sx, w);
      dst[xq] = bilerp_up(ty, tx, __ldg(r0 + tx.i0), __ldg(r0 + tx.i1), __ldg(r1 + tx.i0), __ldg(r1 + tx.i1));
    }
  }
}

// Row-walking form for W % 4 == 0: a thread owns 4 adjacent output columns of one plane and walks down a chunk of
// rows; it keeps the x-interpolated values of its columns on the two source rows that bracket the current output row
// and refreshes them only when the source row changes (every H/h output rows), so an output pixel costs one
// y-interpolation instead of a full bilinear evaluation with its index arithmetic.  Same arithmetic order per value
// (x-lerp, then y-lerp) as upsample_bilinear_kernel.
__global__ void __launch_bounds__(128) upsample_rows_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                            long long planes, int h, int w, int H, int W, float sy,
                                                            float sx, int rows_per_chunk) {
  const int wq = W >> 2, n_chunks = (H + rows_per_chunk - 1) / rows_per_chunk;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= planes * n_chunks * wq) return;
  const int xq = (int)(t % wq);
  const int chunk = (int)((t / wq) % n_chunks);
  const long long p = t / ((long long)wq * n_chunks);
  const int x = xq << 2;
  const int y_begin = chunk * rows_per_chunk, y_end = min(H, y_begin + rows_per_chunk);
  const float *a = in + (size_t)p * h * w;
  float *o = out + (size_t)p * H * W + x;
  Tap tx[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) tx[k] = tap_half_pixel(x + k, sx, w);
  float top[4], bot[4];
  int r0 = -1, r1 = -1;
  for (int y = y_begin; y < y_end; ++y) {
    const Tap ty = tap_half_pixel(y, sy, h);
    if (ty.i0 != r0 || ty.i1 != r1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        top[k] = lerp_nested(tx[k].w0, __ldg(a + ty.i0 * w + tx[k].i0), tx[k].w1, __ldg(a + ty.i0 * w + tx[k].i1));
        bot[k] = lerp_nested(tx[k].w0, __ldg(a + ty.i1 * w + tx[k].i0), tx[k].w1, __ldg(a + ty.i1 * w + tx[k].i1));
      }
      r0 = ty.i0;
      r1 = ty.i1;
    }
    stg_stream4(o + (size_t)y * W, make_float4(lerp_nested(ty.w0, top[0], ty.w1, bot[0]),
                                               lerp_nested(ty.w0, top[1], ty.w1, bot[1]),
                                               lerp_nested(ty.w0, top[2], ty.w1, bot[2]),
                                               lerp_nested(ty.w0, top[3], ty.w1, bot[3])));
  }
}

// First destination index whose upper-left tap is >= k (tap i0 is non-decreasing in the destination index).
__device__ __forceinline__ int first_dst_with_i0(int k, float scale, int src_size, int dst_size) {
  if (k <= 0) return 0;   // negative source coordinates are clamped to 0: the first block starts at the border
  int d = max(0, (int)floorf(((float)k + 0.5f) / scale - 0.5f) - 2);
  while (d < dst_size && tap_half_pixel(d, scale, src_size).i0 < k) ++d;
  return d;
}

// pass 1: one thread per (plane, source row k, 4 destination columns).  The destination rows whose upper tap is k
// form one contiguous block; the thread streams that block ONCE and keeps two sums, the part that belongs to source
// row k (weights w0) and the part that belongs to the row below (weights w1): tmp[p][k][0 | 1][X].
__global__ void __launch_bounds__(256) upsample_adjoint_rows_kernel(const float *__restrict__ g, float *__restrict__ tmp,
                                                                    long long planes, int h, int H, int W, float sy,
                                                                    int vec) {
  const int Wq = vec ? W >> 2 : W;
  const long long total = planes * h * Wq;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(idx % Wq);
    const long long t = idx / Wq;
    const int k = (int)(t % h);
    const long long p = t / h;
    const int y0 = first_dst_with_i0(k, sy, h, H), y1 = k + 1 < h ? first_dst_with_i0(k + 1, sy, h, H) : H;
    const float *src = g + (size_t)p * H * W;
    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
    for (int Y = y0; Y < y1; ++Y) {
      const Tap ty = tap_half_pixel(Y, sy, h);
      if (vec) {
        const float4 v = ldg_stream4(src + (size_t)Y * W + 4 * xq);
        lo.x = fmaf(ty.w0, v.x, lo.x); lo.y = fmaf(ty.w0, v.y, lo.y);
        lo.z = fmaf(ty.w0, v.z, lo.z); lo.w = fmaf(ty.w0, v.w, lo.w);
        hi.x = fmaf(ty.w1, v.x, hi.x); hi.y = fmaf(ty.w1, v.y, hi.y);
        hi.z = fmaf(ty.w1, v.z, hi.z); hi.w = fmaf(ty.w1, v.w, hi.w);
      } else {
        const float v = __ldg(src + (size_t)Y * W + xq);
        lo.x = fmaf(ty.w0, v, lo.x);
        hi.x = fmaf(ty.w1, v, hi.x);
      }
    }
    float *dst = tmp + (((size_t)p * h + k) * 2) * W;
    if (vec) {
      *reinterpret_cast<float4 *>(dst + 4 * xq) = lo;
      *reinterpret_cast<float4 *>(dst + W + 4 * xq) = hi;
    } else {
      dst[xq] = lo.x;
      dst[W + xq] = hi.x;
    }
  }
}

// pass 2: one thread per (plane, source row i, source column j): row i collects its own w0 part, the w1 part of the
// block above and - in the last row, where the lower tap is clamped onto the row itself - its own w1 part; the
// columns are reduced the same way (destination columns whose left tap is j - 1 or j).
__global__ void __launch_bounds__(256) upsample_adjoint_cols_kernel(const float *__restrict__ tmp, float *__restrict__ gin,
                                                                    long long planes, int h, int w, int W, float sx) {
  const long long total = planes * h * w;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % w);
    const long long t = idx / w;
    const int i = (int)(t % h);
    const long long p = t / h;
    const float *own = tmp + (((size_t)p * h + i) * 2) * W;          // [0]: w0 part of block i, [1]: its w1 part
    const float *above = i > 0 ? own - 2 * W + W : nullptr;           // w1 part of block i - 1
    const bool last = i == h - 1;
    const int x0 = j > 0 ? first_dst_with_i0(j - 1, sx, w, W) : 0;
    const int x1 = j + 1 < w ? first_dst_with_i0(j + 1, sx, w, W) : W;
    float acc = 0.0f;
    for (int X = x0; X < x1; ++X) {
      const Tap tx = tap_half_pixel(X, sx, w);
      const float wgt = (tx.i0 == j ? tx.w0 : 0.0f) + (tx.i1 == j ? tx.w1 : 0.0f);
      float v = __ldg(own + X);
      if (above) v += __ldg(above + X);
      if (last) v += __ldg(own + W + X);
      acc = fmaf(wgt, v, acc);
    }
    gin[idx] = acc;
  }
}

static int grid_for(long long items) { return (int)max(1LL, min((long long)sm_count() * 16, ceil_div_ll(items, 256))); }

}  // namespace cosa

using namespace cosa;

extern "C" int cosa_upsample_bilinear(const float *in, float *out, long long planes, int h, int w, int H, int W,
                                      void *stream) {
  if (!in || !out || planes < 1 || h < 1 || w < 1 || H < 1 || W < 1) return COSA_E_ARG;
  const int vec = (W % 4 == 0 && ((uintptr_t)out % 16) == 0) ? 1 : 0;
  if (vec && H >= 2 * h) {   // enlargement: walk down the rows
    const int rows = 56;
    const long long threads = planes * ceil_div(H, rows) * (W / 4);
    COSA_LAUNCH(upsample_rows_kernel, (unsigned)ceil_div_ll(threads, 128), 128, 0, (cudaStream_t)stream, in, out, planes,
                h, w, H, W, (float)h / (float)H, (float)w / (float)W, rows);
    return 0;
  }
  const long long total = planes * H * (vec ? W / 4 : W);
  COSA_LAUNCH(upsample_bilinear_kernel, grid_for(total), 256, 0, (cudaStream_t)stream, in, out, planes, h, w, H, W,
              (float)h / (float)H, (float)w / (float)W, vec);
  return 0;
}

extern "C" size_t cosa_upsample_bilinear_backward_ws_bytes(long long planes, int h, int W) {
  if (planes < 1 || h < 1 || W < 1) return 0;
  return (size_t)planes * h * W * 2 * sizeof(float);
}

extern "C" int cosa_upsample_bilinear_backward(const float *grad_out, float *grad_in, long long planes, int h, int w,
                                               int H, int W, void *ws, size_t ws_bytes, void *stream) {
  if (!grad_out || !grad_in || !ws || planes < 1 || h < 1 || w < 1 || H < 1 || W < 1) return COSA_E_ARG;
  if (ws_bytes < cosa_upsample_bilinear_backward_ws_bytes(planes, h, W)) return COSA_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  float *tmp = (float *)ws;
  const int vec = (W % 4 == 0 && (((uintptr_t)grad_out | (uintptr_t)tmp) % 16) == 0) ? 1 : 0;
  COSA_LAUNCH(upsample_adjoint_rows_kernel, grid_for(planes * h * (vec ? W / 4 : W)), 256, 0, s, grad_out, tmp, planes,
              h, H, W, (float)h / (float)H, vec);
  COSA_LAUNCH(upsample_adjoint_cols_kernel, grid_for(planes * h * w), 256, 0, s, tmp, grad_in, planes, h, w, W,
              (float)w / (float)W);
  return 0;
}
