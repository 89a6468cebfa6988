// PAR — pixel-adaptive refinement (reference: models/PAR.py:26-91) as sm_100a kernels.
//
//   par_affinity_kernel : A[b,n,y,x] = softmax_n( mean_c -(|I_c(nbr_n)-I_c|/(std48_c+1e-8)/w1)^2 )
//                                      + w2 * softmax_n( -(pos_n/(std(pos)+1e-8)/w1)^2 )          (PAR.py:69-85)
//   par_iterate_kernel  : M'[b,c,y,x] = sum_n A[b,n,y,x] * M[b,c,clamp(y+dy_n),clamp(x+dx_n)]       (PAR.py:87-89)
//   resize_align_corners_kernel : masks -> image size, bilinear align_corners=True                 (PAR.py:66)
//
// Neighbour n = 8*k + m: dilation k in ctor order, direction m in get_kernel() order (PAR.py:10-24):
// (-,-) (-,0) (-,+) (0,-) (0,+) (+,-) (+,0) (+,+), borders replicated (PAR.py:44).
#include <math.h>

#include "common.cuh"
#include "par.cuh"

namespace cosa {

__constant__ int c_dil[kMaxDil];
__constant__ float c_pos_term[kMaxDil * 8];   // w2 * softmax(pos_aff), filled by par_upload_constants

// Host: the position term is a constant vector (PAR.py:51-62,77,82); evaluate it in double.
int par_upload_constants(const int *dilations, int n_dil, cudaStream_t stream) {
  if (n_dil < 1 || n_dil > kMaxDil) return COSA_E_ARG;
  const int nd = 8 * n_dil;
  double pos[kMaxDil * 8], mean = 0.0;
  for (int k = 0; k < n_dil; ++k) {
    if (dilations[k] < 1) return COSA_E_ARG;
    for (int m = 0; m < 8; ++m) {
      const bool diag = (m == 0 || m == 2 || m == 5 || m == 7);
      // PAR.py:54-58: float32 ones with sqrt(2) stored as float32, times the (integer) dilation
      pos[8 * k + m] = (double)((float)(diag ? (float)sqrt(2.0) : 1.0f) * (float)dilations[k]);
      mean += pos[8 * k + m];
    }
  }
  mean /= nd;
  double var = 0.0;
  for (int n = 0; n < nd; ++n) var += (pos[n] - mean) * (pos[n] - mean);
  const double sd = sqrt(var / (nd - 1));            // torch.std is unbiased
  double logit[kMaxDil * 8], mx = -1e300, sum = 0.0;
  for (int n = 0; n < nd; ++n) {
    const double t = pos[n] / (sd + 1e-8) / 0.3;
    logit[n] = -t * t;
    mx = fmax(mx, logit[n]);
  }
  for (int n = 0; n < nd; ++n) sum += exp(logit[n] - mx);
  float term[kMaxDil * 8];
  for (int n = 0; n < nd; ++n) term[n] = 0.01f * (float)(exp(logit[n] - mx) / sum);
  int dil[kMaxDil] = {0};
  for (int k = 0; k < n_dil; ++k) dil[k] = dilations[k];
  COSA_CUDA(cudaMemcpyToSymbolAsync(c_dil, dil, sizeof(dil), 0, cudaMemcpyHostToDevice, stream));
  COSA_CUDA(cudaMemcpyToSymbolAsync(c_pos_term, term, sizeof(float) * nd, 0, cudaMemcpyHostToDevice, stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Affinity.  One thread per pixel; per colour channel the 8*NDIL neighbour differences live in registers
// (two-pass unbiased variance: the one-pass form cancels catastrophically in flat regions).
// ------------------------------------------------------------------------------------------------
template <int NDIL>
__global__ void __launch_bounds__(128) par_affinity_kernel(const float *__restrict__ imgs, float *__restrict__ aff,
                                                           int h, int w) {
  constexpr int ND = 8 * NDIL;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 4 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (x >= w || y >= h) return;
  const size_t plane = (size_t)h * w;
  const float *img = imgs + (size_t)b * 3 * plane;

  float logit[ND];
#pragma unroll
  for (int n = 0; n < ND; ++n) logit[n] = 0.0f;

#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    const float *ch = img + c * plane;
    const float ctr = __ldg(ch + (size_t)y * w + x);
    float v[ND];
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < NDIL; ++k) {
      const int d = c_dil[k];
      const int ym = max(y - d, 0), yp = min(y + d, h - 1);
      const int xm = max(x - d, 0), xp = min(x + d, w - 1);
      const float *r0 = ch + (size_t)ym * w, *r1 = ch + (size_t)y * w, *r2 = ch + (size_t)yp * w;
      v[8 * k + 0] = __ldg(r0 + xm); v[8 * k + 1] = __ldg(r0 + x); v[8 * k + 2] = __ldg(r0 + xp);
      v[8 * k + 3] = __ldg(r1 + xm);                               v[8 * k + 4] = __ldg(r1 + xp);
      v[8 * k + 5] = __ldg(r2 + xm); v[8 * k + 6] = __ldg(r2 + x); v[8 * k + 7] = __ldg(r2 + xp);
    }
#pragma unroll
    for (int n = 0; n < ND; ++n) sum += v[n];
    const float mean = sum / (float)ND;
    float ss = 0.0f;
#pragma unroll
    for (int n = 0; n < ND; ++n) {
      const float t = v[n] - mean;
      ss = fmaf(t, t, ss);
    }
    const float sd = sqrtf(ss / (float)(ND - 1));
    const float inv = 1.0f / ((sd + 1e-8f) * 0.3f);
#pragma unroll
    for (int n = 0; n < ND; ++n) {
      const float t = fabsf(v[n] - ctr) * inv;
      logit[n] = fmaf(t, t, logit[n]);
    }
  }
  // aff = -(sum_c t^2)/3 ; softmax over the ND neighbours; add the position term.
  float mx = -INFINITY;
#pragma unroll
  for (int n = 0; n < ND; ++n) {
    logit[n] = -logit[n] / 3.0f;
    mx = fmaxf(mx, logit[n]);
  }
  float den = 0.0f;
#pragma unroll
  for (int n = 0; n < ND; ++n) {
    logit[n] = expf(logit[n] - mx);
    den += logit[n];
  }
  const float rden = 1.0f / den;
  float *out = aff + (size_t)b * ND * plane + (size_t)y * w + x;
#pragma unroll
  for (int n = 0; n < ND; ++n) out[(size_t)n * plane] = fmaf(logit[n], rden, c_pos_term[n]);
}

// Generic (any n_dil <= kMaxDil) three-pass variant: neighbours are re-read instead of kept in registers.
__global__ void __launch_bounds__(128) par_affinity_generic_kernel(const float *__restrict__ imgs,
                                                                   float *__restrict__ aff, int h, int w, int n_dil) {
  const int nd = 8 * n_dil;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 4 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (x >= w || y >= h) return;
  const size_t plane = (size_t)h * w;
  const float *img = imgs + (size_t)b * 3 * plane;
  float *out = aff + (size_t)b * nd * plane + (size_t)y * w + x;
  auto nbr = [&](const float *ch, int n) {
    const int d = c_dil[n >> 3], m = n & 7;
    const int dy = (m < 3) ? -d : (m < 5 ? 0 : d);
    const int dx = (m == 0 || m == 3 || m == 5) ? -d : ((m == 1 || m == 6) ? 0 : d);
    return __ldg(ch + (size_t)clampi(y + dy, 0, h - 1) * w + clampi(x + dx, 0, w - 1));
  };
  float inv[3], ctr[3];
  for (int c = 0; c < 3; ++c) {
    const float *ch = img + c * plane;
    ctr[c] = __ldg(ch + (size_t)y * w + x);
    float sum = 0.0f;
    for (int n = 0; n < nd; ++n) sum += nbr(ch, n);
    const float mean = sum / (float)nd;
    float ss = 0.0f;
    for (int n = 0; n < nd; ++n) {
      const float t = nbr(ch, n) - mean;
      ss = fmaf(t, t, ss);
    }
    inv[c] = 1.0f / ((sqrtf(ss / (float)(nd - 1)) + 1e-8f) * 0.3f);
  }
  float mx = -INFINITY;
  for (int n = 0; n < nd; ++n) {
    float l = 0.0f;
    for (int c = 0; c < 3; ++c) {
      const float t = fabsf(nbr(img + c * plane, n) - ctr[c]) * inv[c];
      l = fmaf(t, t, l);
    }
    l = -l / 3.0f;
    out[(size_t)n * plane] = l;
    mx = fmaxf(mx, l);
  }
  float den = 0.0f;
  for (int n = 0; n < nd; ++n) den += expf(out[(size_t)n * plane] - mx);
  const float rden = 1.0f / den;
  for (int n = 0; n < nd; ++n)
    out[(size_t)n * plane] = fmaf(expf(out[(size_t)n * plane] - mx), rden, c_pos_term[n]);
}

// ------------------------------------------------------------------------------------------------
// One propagation step.  Thread per pixel, CH mask channels per pass held in registers; the affinity of the
// pixel is streamed once per pass (coalesced per neighbour plane), mask neighbours come through L1.
// nch_dev (optional) gives the number of live channels per image for the ragged cam2mask batch.
// ------------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256) par_iterate_kernel(const float *__restrict__ aff, const float *__restrict__ in,
                                                          float *__restrict__ out, const int *__restrict__ nch_dev,
                                                          int nch_uniform, int c_stride, int h, int w, int n_dil) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (x >= w || y >= h) return;
  const int nch = nch_dev ? nch_dev[b] : nch_uniform;
  const size_t plane = (size_t)h * w;
  const size_t pix = (size_t)y * w + x;
  const float *A = aff + (size_t)b * (8 * n_dil) * plane + pix;
  const float *src = in + (size_t)b * c_stride * plane;
  float *dst = out + (size_t)b * c_stride * plane + pix;

  for (int c0 = 0; c0 < nch; c0 += CH) {
    float acc[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = 0.0f;
    const float *s0 = src + (size_t)c0 * plane;
    const int live = min(CH, nch - c0);
#pragma unroll 1
    for (int kd = 0; kd < n_dil; ++kd) {
      const int d = c_dil[kd];
      const int ym = max(y - d, 0) * w, y0 = y * w, yp = min(y + d, h - 1) * w;
      const int xm = max(x - d, 0), xp = min(x + d, w - 1);
      const int off[8] = {ym + xm, ym + x, ym + xp, y0 + xm, y0 + xp, yp + xm, yp + x, yp + xp};
      float a[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) a[m] = __ldg(A + (size_t)(8 * kd + m) * plane);
      if (live == CH) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          const float *ch = s0 + (size_t)k * plane;
#pragma unroll
          for (int m = 0; m < 8; ++m) acc[k] = fmaf(a[m], __ldg(ch + off[m]), acc[k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          if (k < live) {
            const float *ch = s0 + (size_t)k * plane;
#pragma unroll
            for (int m = 0; m < 8; ++m) acc[k] = fmaf(a[m], __ldg(ch + off[m]), acc[k]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < CH; ++k)
      if (k < live) dst[(size_t)(c0 + k) * plane] = acc[k];
  }
}

__global__ void resize_align_corners_kernel(const float *__restrict__ in, float *__restrict__ out, int planes, int hi,
                                            int wi, int ho, int wo, float sy, float sx) {
  const long long total = (long long)planes * ho * wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % wo), y = (int)((i / wo) % ho);
    const long long p = i / ((long long)wo * ho);
    const Tap ty = tap_align_corners(y, sy, hi), tx = tap_align_corners(x, sx, wi);
    const float *s = in + p * (long long)hi * wi;
    out[i] = bilerp_up(ty, tx, s[(size_t)ty.i0 * wi + tx.i0], s[(size_t)ty.i0 * wi + tx.i1],
                       s[(size_t)ty.i1 * wi + tx.i0], s[(size_t)ty.i1 * wi + tx.i1]);
  }
}

// ------------------------------------------------------------------------------------------------
// Host-side launchers (shared with cam2mask).
// ------------------------------------------------------------------------------------------------
int par_launch_affinity(const float *imgs, float *aff, int B, int h, int w, int n_dil, cudaStream_t stream) {
  dim3 grid(ceil_div(w, 32), ceil_div(h, 4), B), block(128);
  if (n_dil == 6) {
    COSA_LAUNCH(par_affinity_kernel<6>, grid, block, 0, stream, imgs, aff, h, w);
  } else {
    COSA_LAUNCH(par_affinity_generic_kernel, grid, block, 0, stream, imgs, aff, h, w, n_dil);
  }
  return 0;
}

int par_launch_iterations(const float *aff, const float *src0, float *scratch_a, float *scratch_b, float *final_dst,
                          const int *nch_dev, int nch_uniform, int c_stride, int B, int h, int w, int n_dil,
                          int num_iter, cudaStream_t stream) {
  if (num_iter == 0) {
    COSA_CUDA(cudaMemcpyAsync(final_dst, src0, (size_t)B * c_stride * h * w * sizeof(float),
                              cudaMemcpyDeviceToDevice, stream));
    return 0;
  }
  dim3 grid(ceil_div(w, 32), ceil_div(h, 8), B), block(256);
  const bool wide = nch_dev ? (c_stride > 4) : (nch_uniform > 4);
  const float *src = src0;
  for (int it = 0; it < num_iter; ++it) {
    float *dst = (it == num_iter - 1) ? final_dst : ((it & 1) ? scratch_b : scratch_a);
    if (wide) {
      COSA_LAUNCH(par_iterate_kernel<8>, grid, block, 0, stream, aff, src, dst, nch_dev, nch_uniform, c_stride, h, w,
                  n_dil);
    } else {
      COSA_LAUNCH(par_iterate_kernel<4>, grid, block, 0, stream, aff, src, dst, nch_dev, nch_uniform, c_stride, h, w,
                  n_dil);
    }
    src = dst;
  }
  return 0;
}

}  // namespace cosa

using namespace cosa;

extern "C" size_t cosa_par_ws_bytes(int B, int C, int h, int w, int n_dil) {
  const size_t plane = (size_t)h * w;
  return align_up((size_t)B * 8 * n_dil * plane * sizeof(float), 256) +
         2 * align_up((size_t)B * C * plane * sizeof(float), 256);
}

extern "C" int cosa_par_affinity(const float *imgs, float *aff, int B, int h, int w, const int *dilations, int n_dil,
                                 void *stream) {
  if (!imgs || !aff || B < 1 || h < 1 || w < 1) return COSA_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  COSA_CHECK(par_upload_constants(dilations, n_dil, s));
  return par_launch_affinity(imgs, aff, B, h, w, n_dil, s);
}

extern "C" int cosa_par_forward(const float *imgs, const float *masks_in, float *masks_out, int B, int C, int h, int w,
                                int hm, int wm, const int *dilations, int n_dil, int num_iter, void *ws,
                                size_t ws_bytes, void *stream) {
  if (!imgs || !masks_in || !masks_out || !ws || B < 1 || C < 1 || h < 1 || w < 1 || hm < 1 || wm < 1 || num_iter < 0)
    return COSA_E_ARG;
  if (ws_bytes < cosa_par_ws_bytes(B, C, h, w, n_dil)) return COSA_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  COSA_CHECK(par_upload_constants(dilations, n_dil, s));
  const size_t plane = (size_t)h * w;
  Arena arena(ws);
  float *aff = arena.take<float>((size_t)B * 8 * n_dil * plane);
  float *buf_a = arena.take<float>((size_t)B * C * plane);
  float *buf_b = arena.take<float>((size_t)B * C * plane);
  COSA_CHECK(par_launch_affinity(imgs, aff, B, h, w, n_dil, s));
  const float *src0 = masks_in;
  int iters = num_iter;
  if (hm != h || wm != w) {
    // PAR.py:66 - bilinear, align_corners=True.  With no iteration left to run the resize is the output.
    const float sy = h > 1 ? (float)(hm - 1) / (float)(h - 1) : 0.0f;
    const float sx = w > 1 ? (float)(wm - 1) / (float)(w - 1) : 0.0f;
    const long long total = (long long)B * C * plane;
    const int blocks = (int)min((long long)sm_count() * 8, ceil_div_ll(total, 256));
    float *dst = (iters == 0) ? masks_out : buf_b;
    COSA_LAUNCH(resize_align_corners_kernel, blocks, 256, 0, s, masks_in, dst, B * C, hm, wm, h, w, sy, sx);
    if (iters == 0) return 0;
    // first step reads the resized masks from buf_b and writes buf_a (or masks_out when it is the only one)
    COSA_CHECK(par_launch_iterations(aff, buf_b, buf_a, buf_a, (iters == 1) ? masks_out : buf_a, nullptr, C, C, B, h,
                                     w, n_dil, 1, s));
    if (iters == 1) return 0;
    src0 = buf_a;
    iters -= 1;
    // remaining steps ping-pong buf_b / (a third slice is not needed: buf_a is only read by the next step)
    return par_launch_iterations(aff, src0, buf_b, buf_a, masks_out, nullptr, C, C, B, h, w, n_dil, iters, s);
  }
  return par_launch_iterations(aff, src0, buf_a, buf_b, masks_out, nullptr, C, C, B, h, w, n_dil, iters, s);
}
