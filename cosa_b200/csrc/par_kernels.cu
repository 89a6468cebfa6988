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
#include <stdlib.h>

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
// One propagation step, generic form (any width, plain NCHW layout).  Thread per pixel, CH mask channels per
// pass in registers; clamped scalar neighbour loads through L1.
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
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        if (k < live) {
          const float *ch = s0 + (size_t)k * plane;
#pragma unroll
          for (int m = 0; m < 8; ++m) acc[k] = fmaf(a[m], __ldg(ch + off[m]), acc[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < CH; ++k)
      if (k < live) dst[(size_t)(c0 + k) * plane] = acc[k];
  }
}

// ------------------------------------------------------------------------------------------------
// One propagation step, vectorised form (w % 4 == 0, padded rows).  A thread owns 4 horizontally adjacent pixels
// and CH channels: per dilation it streams the 8 affinity quads (128-bit, no L1 allocation) and, per channel and
// neighbour row, three aligned 128-bit loads that cover the column offsets -d, 0, +d of all four pixels
// (d % 4 == 0: the quads at x-d, x, x+d; d < 4: the quads left/centre/right, recombined in registers).
// The replicated column pads make every load unclamped; rows are clamped by index.  A warp covers 4 rows x
// 8 quads, i.e. one 128-byte line per row when the interior is 128-byte aligned.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

__device__ __forceinline__ void fma4(float4 &acc, const float4 &a, const float4 &v) {
  acc.x = fmaf(a.x, v.x, acc.x);
  acc.y = fmaf(a.y, v.y, acc.y);
  acc.z = fmaf(a.z, v.z, acc.z);
  acc.w = fmaf(a.w, v.w, acc.w);
}

// the quads starting at column offsets -d and +d of the centre quad C, from the aligned quads L, C, R (d = 1..3)
__device__ __forceinline__ void shifted_quads(const float4 &L, const float4 &C, const float4 &R, int d, float4 &m,
                                              float4 &p) {
  if (d == 1) {
    m = make_float4(L.w, C.x, C.y, C.z);
    p = make_float4(C.y, C.z, C.w, R.x);
  } else if (d == 2) {
    m = make_float4(L.z, L.w, C.x, C.y);
    p = make_float4(C.z, C.w, R.x, R.y);
  } else {
    m = make_float4(L.y, L.z, L.w, C.x);
    p = make_float4(C.w, R.x, R.y, R.z);
  }
}

static int par_tile_log2() {   // width of the CTA tile in quads (log2); COSA_PAR_TILE_LOG2 overrides
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("COSA_PAR_TILE_LOG2");
    v = e ? atoi(e) : 3;
    if (v < 0 || v > 8) v = 3;
  }
  return v;
}

template <int CH>
__global__ void __launch_bounds__(256, 2)
    par_iterate_vec_kernel(const float *__restrict__ aff, const float *__restrict__ in, MaskLayout li,
                           float *__restrict__ out, MaskLayout lo, const int *__restrict__ nch_dev, int nch_uniform,
                           int c_stride, int h, int w, int n_dil, int tq_log2) {
  // CTA tile = 2^tq_log2 quads x (256 >> tq_log2) rows; a warp then covers 4 rows x 8 quads (one 128-byte line per row)
  const int wq = w >> 2;
  const int xq = (blockIdx.x << tq_log2) + (threadIdx.x & ((1 << tq_log2) - 1));
  const int y = blockIdx.y * (256 >> tq_log2) + (threadIdx.x >> tq_log2);
  const int b = blockIdx.z;
  if (xq >= wq || y >= h) return;
  const int x = xq << 2;
  const int nch = nch_dev ? nch_dev[b] : nch_uniform;
  const size_t plane = (size_t)h * w;
  const size_t iplane = (size_t)h * li.pitch, oplane = (size_t)h * lo.pitch;
  const float *A = aff + (size_t)b * (8 * n_dil) * plane + (size_t)y * w + x;
  const float *src = in + (size_t)b * c_stride * iplane + li.off + x;
  float *dst = out + (size_t)b * c_stride * oplane + (size_t)y * lo.pitch + lo.off + x;

  for (int c0 = 0; c0 < nch; c0 += CH) {
    float4 acc[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int live = min(CH, nch - c0);
#pragma unroll 1
    for (int kd = 0; kd < n_dil; ++kd) {
      const int d = c_dil[kd];
      float4 a[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) a[m] = ldg_stream4(A + (size_t)(8 * kd + m) * plane);
      const size_t rm = (size_t)max(y - d, 0) * li.pitch, r0 = (size_t)y * li.pitch,
                   rp = (size_t)min(y + d, h - 1) * li.pitch;
      if ((d & 3) == 0) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          if (k < live) {
            const float *ch = src + (size_t)(c0 + k) * iplane;
            fma4(acc[k], a[0], ldg4(ch + rm - d));
            fma4(acc[k], a[1], ldg4(ch + rm));
            fma4(acc[k], a[2], ldg4(ch + rm + d));
            fma4(acc[k], a[3], ldg4(ch + r0 - d));
            fma4(acc[k], a[4], ldg4(ch + r0 + d));
            fma4(acc[k], a[5], ldg4(ch + rp - d));
            fma4(acc[k], a[6], ldg4(ch + rp));
            fma4(acc[k], a[7], ldg4(ch + rp + d));
          }
        }
      } else if (d < 4) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          if (k < live) {
            const float *ch = src + (size_t)(c0 + k) * iplane;
            float4 m, p;
            float4 C = ldg4(ch + rm);
            shifted_quads(ldg4(ch + rm - 4), C, ldg4(ch + rm + 4), d, m, p);
            fma4(acc[k], a[0], m); fma4(acc[k], a[1], C); fma4(acc[k], a[2], p);
            C = ldg4(ch + r0);
            shifted_quads(ldg4(ch + r0 - 4), C, ldg4(ch + r0 + 4), d, m, p);
            fma4(acc[k], a[3], m); fma4(acc[k], a[4], p);
            C = ldg4(ch + rp);
            shifted_quads(ldg4(ch + rp - 4), C, ldg4(ch + rp + 4), d, m, p);
            fma4(acc[k], a[5], m); fma4(acc[k], a[6], C); fma4(acc[k], a[7], p);
          }
        }
      } else {   // unaligned large dilation: scalar loads, still unclamped thanks to the pads
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          if (k < live) {
            const float *ch = src + (size_t)(c0 + k) * iplane;
            const size_t rows[3] = {rm, r0, rp};
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
              const float *q = ch + rows[rr];
#pragma unroll
              for (int cc = 0; cc < 3; ++cc) {
                if (rr == 1 && cc == 1) continue;
                const int mi = rr * 3 + cc - (rr * 3 + cc > 4 ? 1 : 0);
                const float *qq = q + (cc - 1) * d;
                fma4(acc[k], a[mi], make_float4(__ldg(qq), __ldg(qq + 1), __ldg(qq + 2), __ldg(qq + 3)));
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      if (k < live) {
        float *o = dst + (size_t)(c0 + k) * oplane;
        *reinterpret_cast<float4 *>(o) = acc[k];
        if (lo.padn) {   // replicate the edge pixels into the column pads for the next step
          if (xq == 0) {
            const float4 e = make_float4(acc[k].x, acc[k].x, acc[k].x, acc[k].x);
            for (int i = 4; i <= lo.padn; i += 4) *reinterpret_cast<float4 *>(o - i) = e;
          }
          if (xq == wq - 1) {
            const float4 e = make_float4(acc[k].w, acc[k].w, acc[k].w, acc[k].w);
            for (int i = 4; i <= lo.padn; i += 4) *reinterpret_cast<float4 *>(o + i) = e;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// One propagation step with the near neighbourhood staged in shared memory by the TMA unit.
//
// CTA tile: 32 rows x 32 pixels (256 threads, 4 pixels x CH channels each, as in the vector kernel).  The mask tile
// of every live channel, with a halo of kHalo = 8 pixels, is brought into shared memory by cp.async.bulk row copies
// (one 192-byte copy per tile row and channel, completion on an mbarrier): 32 of the 48 neighbours (dilations
// 1, 2, 4, 8) are then served by conflict-free 128-bit shared-memory loads; only dilations > 8 go to L1/L2.
// Rows are clamped when the copy is issued (replicate border), columns come from the replicated column pads.
// ------------------------------------------------------------------------------------------------
constexpr int kHalo = 8;
constexpr int kTileW = 32, kTileH = 32;
constexpr int kSmemW = kTileW + 2 * kHalo;   // 48 floats = 192 bytes per staged row
constexpr int kSmemH = kTileH + 2 * kHalo;   // 48 rows

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, unsigned bytes,
                                              unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

template <int CH>
__global__ void __launch_bounds__(256, 2)
    par_iterate_smem_kernel(const float *__restrict__ aff, const float *__restrict__ in, MaskLayout li,
                            float *__restrict__ out, MaskLayout lo, const int *__restrict__ nch_dev, int nch_uniform,
                            int c_stride, int h, int w, int n_dil) {
  extern __shared__ __align__(128) float s_tile[];   // [CH][kSmemH][kSmemW]
  __shared__ __align__(8) unsigned long long s_bar;
  const int wq = w >> 2;
  const int tq = threadIdx.x & 7, tr = threadIdx.x >> 3;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const int xq = (x0 >> 2) + tq, y = y0 + tr, x = x0 + (tq << 2);
  const int b = blockIdx.z;
  const bool active = xq < wq && y < h;            // everybody stays for the staging and the barrier
  const int nch = nch_dev ? nch_dev[b] : nch_uniform;
  const size_t plane = (size_t)h * w;
  const size_t iplane = (size_t)h * li.pitch, oplane = (size_t)h * lo.pitch;
  const float *A = aff + (size_t)b * (8 * n_dil) * plane + (size_t)min(y, h - 1) * w + min(x, w - 4);
  const float *src = in + (size_t)b * c_stride * iplane + li.off;
  float *dst = out + (size_t)b * c_stride * oplane + (size_t)y * lo.pitch + lo.off + x;

  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  unsigned phase = 0;
  for (int c0 = 0; c0 < nch; c0 += CH) {
    const int live = min(CH, nch - c0);
    // ---- stage the halo tile of the live channels: one bulk row copy per (channel, row) ---------------------
    if (threadIdx.x == 0) mbar_expect_tx(&s_bar, (unsigned)(live * kSmemH * kSmemW * sizeof(float)));
    for (int i = threadIdx.x; i < live * kSmemH; i += 256) {
      const int k = i / kSmemH, r = i - k * kSmemH;
      const int gy = min(max(y0 - kHalo + r, 0), h - 1);
      bulk_copy_g2s(s_tile + (size_t)i * kSmemW, src + (size_t)(c0 + k) * iplane + (size_t)gy * li.pitch + (x0 - kHalo),
                    kSmemW * sizeof(float), &s_bar);
    }
    float4 acc[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    bool staged = false;
    // far dilations first (global loads) while the bulk copies are in flight, then the staged near ones
#pragma unroll 1
    for (int it = 0; it < 2 * n_dil; ++it) {
      const int kd = it < n_dil ? it : it - n_dil;
      const int d = c_dil[kd];
      if ((d <= kHalo) != (it >= n_dil)) continue;
      float4 a[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) a[m] = ldg_stream4(A + (size_t)(8 * kd + m) * plane);
      if (d <= kHalo) {
        if (!staged) { mbar_wait(&s_bar, phase); staged = true; }
        // shared-memory coordinates of this thread's quad: row tr + kHalo, column 4 tq + kHalo
        const float *t0 = s_tile + (size_t)(tr + kHalo) * kSmemW + (tq << 2) + kHalo;
        const int rm = -d * kSmemW, rp = d * kSmemW;
        if ((d & 3) == 0) {
#pragma unroll
          for (int k = 0; k < CH; ++k) {
            if (k < live) {
              const float *q = t0 + k * (kSmemH * kSmemW);
              fma4(acc[k], a[0], lds4(q + rm - d)); fma4(acc[k], a[1], lds4(q + rm)); fma4(acc[k], a[2], lds4(q + rm + d));
              fma4(acc[k], a[3], lds4(q - d));                                          fma4(acc[k], a[4], lds4(q + d));
              fma4(acc[k], a[5], lds4(q + rp - d)); fma4(acc[k], a[6], lds4(q + rp)); fma4(acc[k], a[7], lds4(q + rp + d));
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < CH; ++k) {
            if (k < live) {
              const float *q = t0 + k * (kSmemH * kSmemW);
              float4 m, p, C;
              if (d < 4) {
                C = lds4(q + rm); shifted_quads(lds4(q + rm - 4), C, lds4(q + rm + 4), d, m, p);
                fma4(acc[k], a[0], m); fma4(acc[k], a[1], C); fma4(acc[k], a[2], p);
                C = lds4(q); shifted_quads(lds4(q - 4), C, lds4(q + 4), d, m, p);
                fma4(acc[k], a[3], m); fma4(acc[k], a[4], p);
                C = lds4(q + rp); shifted_quads(lds4(q + rp - 4), C, lds4(q + rp + 4), d, m, p);
                fma4(acc[k], a[5], m); fma4(acc[k], a[6], C); fma4(acc[k], a[7], p);
              } else {   // 5, 6, 7: unaligned, scalar shared-memory reads
                const int rows[3] = {rm, 0, rp};
#pragma unroll
                for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                  for (int cc = 0; cc < 3; ++cc) {
                    if (rr == 1 && cc == 1) continue;
                    const int mi = rr * 3 + cc - (rr * 3 + cc > 4 ? 1 : 0);
                    const float *qq = q + rows[rr] + (cc - 1) * d;
                    fma4(acc[k], a[mi], make_float4(qq[0], qq[1], qq[2], qq[3]));
                  }
              }
            }
          }
        }
      } else if (active) {
        const float *g0 = src + (size_t)c0 * iplane + x;
        const size_t rm = (size_t)max(y - d, 0) * li.pitch, r0 = (size_t)y * li.pitch,
                     rp = (size_t)min(y + d, h - 1) * li.pitch;
        if ((d & 3) == 0) {
#pragma unroll
          for (int k = 0; k < CH; ++k) {
            if (k < live) {
              const float *ch = g0 + (size_t)k * iplane;
              fma4(acc[k], a[0], ldg4(ch + rm - d)); fma4(acc[k], a[1], ldg4(ch + rm)); fma4(acc[k], a[2], ldg4(ch + rm + d));
              fma4(acc[k], a[3], ldg4(ch + r0 - d));                                       fma4(acc[k], a[4], ldg4(ch + r0 + d));
              fma4(acc[k], a[5], ldg4(ch + rp - d)); fma4(acc[k], a[6], ldg4(ch + rp)); fma4(acc[k], a[7], ldg4(ch + rp + d));
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < CH; ++k) {
            if (k < live) {
              const float *ch = g0 + (size_t)k * iplane;
              const size_t rows[3] = {rm, r0, rp};
#pragma unroll
              for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                  if (rr == 1 && cc == 1) continue;
                  const int mi = rr * 3 + cc - (rr * 3 + cc > 4 ? 1 : 0);
                  const float *qq = ch + rows[rr] + (cc - 1) * d;
                  fma4(acc[k], a[mi], make_float4(__ldg(qq), __ldg(qq + 1), __ldg(qq + 2), __ldg(qq + 3)));
                }
            }
          }
        }
      }
    }
    if (!staged) mbar_wait(&s_bar, phase);   // no dilation used the tile: still consume the phase
    phase ^= 1;
    if (active) {
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        if (k < live) {
          float *o = dst + (size_t)(c0 + k) * oplane;
          *reinterpret_cast<float4 *>(o) = acc[k];
          if (lo.padn) {   // replicate the edge pixels into the column pads for the next step
            if (xq == 0) {
              const float4 e = make_float4(acc[k].x, acc[k].x, acc[k].x, acc[k].x);
              for (int i = 4; i <= lo.padn; i += 4) *reinterpret_cast<float4 *>(o - i) = e;
            }
            if (xq == wq - 1) {
              const float4 e = make_float4(acc[k].w, acc[k].w, acc[k].w, acc[k].w);
              for (int i = 4; i <= lo.padn; i += 4) *reinterpret_cast<float4 *>(o + i) = e;
            }
          }
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();   // the tile is re-staged (async proxy) by the next channel group
  }
}

// plain [planes, h, w] -> padded layout (interior + replicated column pads)
__global__ void par_pack_kernel(const float *__restrict__ src, float *__restrict__ dst, MaskLayout l, int planes,
                                int h, int w) {
  const int span = w + 2 * l.padn;
  const long long total = (long long)planes * h * span;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % span);
    const long long row = i / span;
    const int x = clampi(c - l.padn, 0, w - 1);
    dst[row * l.pitch + l.off - l.padn + c] = __ldg(src + row * w + x);
  }
}

__global__ void resize_align_corners_kernel(const float *__restrict__ in, float *__restrict__ out, MaskLayout l,
                                            int planes, int hi, int wi, int ho, int wo, float sy, float sx) {
  const int span = wo + 2 * l.padn;
  const long long total = (long long)planes * ho * span;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % span);
    const long long row = i / span;
    const int x = clampi(c - l.padn, 0, wo - 1), y = (int)(row % ho);
    const long long p = row / ho;
    const Tap ty = tap_align_corners(y, sy, hi), tx = tap_align_corners(x, sx, wi);
    const float *s = in + p * (long long)hi * wi;
    out[row * l.pitch + l.off - l.padn + c] =
        bilerp_up(ty, tx, s[(size_t)ty.i0 * wi + tx.i0], s[(size_t)ty.i0 * wi + tx.i1],
                  s[(size_t)ty.i1 * wi + tx.i0], s[(size_t)ty.i1 * wi + tx.i1]);
  }
}

// ------------------------------------------------------------------------------------------------
// Host-side launchers (shared with cam2mask).
// ------------------------------------------------------------------------------------------------
MaskLayout padded_layout(int w, const int *dilations, int n_dil) {
  if (w % 4 != 0 || w < 8) return plain_layout(w);
  int max_dil = 1;
  for (int k = 0; k < n_dil; ++k) max_dil = max(max_dil, dilations[k]);
  MaskLayout l;
  l.padn = (max(max_dil, kHalo) + 3) & ~3;              // the staged tile reads kHalo columns beyond the image
  l.off = (l.padn + 31) & ~31;
  l.pitch = (l.off + ((w + 31) & ~31) + l.padn + 31) & ~31;   // whole 32-pixel tiles plus the right pad
  return l;
}

int par_launch_affinity(const float *imgs, float *aff, int B, int h, int w, int n_dil, cudaStream_t stream) {
  dim3 grid(ceil_div(w, 32), ceil_div(h, 4), B), block(128);
  static int force_generic = -1;
  if (force_generic < 0) force_generic = getenv("COSA_PAR_AFF_GENERIC") ? 1 : 0;
  if (n_dil == 6 && !force_generic) {
    COSA_LAUNCH(par_affinity_kernel<6>, grid, block, 0, stream, imgs, aff, h, w);
  } else {
    COSA_LAUNCH(par_affinity_generic_kernel, grid, block, 0, stream, imgs, aff, h, w, n_dil);
  }
  return 0;
}

int par_launch_pack(const float *src, float *dst, MaskLayout lay, int planes, int h, int w, cudaStream_t stream) {
  const long long total = (long long)planes * h * (w + 2 * lay.padn);
  const int blocks = (int)max(1LL, min((long long)sm_count() * 8, ceil_div_ll(total, 256)));
  COSA_LAUNCH(par_pack_kernel, blocks, 256, 0, stream, src, dst, lay, planes, h, w);
  return 0;
}

int par_launch_iterations(const float *aff, const float *src0, float *scratch_a, float *scratch_b, MaskLayout lay,
                          float *final_dst, MaskLayout lay_final, const int *nch_dev, int nch_uniform, int c_stride,
                          int B, int h, int w, int n_dil, int num_iter, cudaStream_t stream) {
  if (num_iter <= 0) return COSA_E_ARG;   // callers handle the zero-iteration copy themselves
  const bool wide = nch_dev ? (c_stride > 4) : (nch_uniform > 4);
  const bool vec = lay.padn > 0;
  const float *src = src0;
  for (int it = 0; it < num_iter; ++it) {
    const bool last = it == num_iter - 1;
    float *dst = last ? final_dst : ((it & 1) ? scratch_b : scratch_a);
    const MaskLayout lo = last ? lay_final : lay;
    static int step_kind = -1;   // COSA_PAR_STEP=vec selects the L1-only vector kernel (for A/B measurements)
    if (step_kind < 0) {
      const char *e = getenv("COSA_PAR_STEP");
      step_kind = (e && e[0] == 'v') ? 1 : 0;
    }
    if (vec && step_kind == 0) {
      dim3 grid(ceil_div(w, kTileW), ceil_div(h, kTileH), B), block(256);
      static bool attr_set = false;
      if (!attr_set) {
        COSA_CUDA(cudaFuncSetAttribute(par_iterate_smem_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       8 * kSmemH * kSmemW * (int)sizeof(float)));
        COSA_CUDA(cudaFuncSetAttribute(par_iterate_smem_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       4 * kSmemH * kSmemW * (int)sizeof(float)));
        attr_set = true;
      }
      if (wide) {
        COSA_LAUNCH(par_iterate_smem_kernel<8>, grid, block, 8 * kSmemH * kSmemW * sizeof(float), stream, aff, src, lay,
                    dst, lo, nch_dev, nch_uniform, c_stride, h, w, n_dil);
      } else {
        COSA_LAUNCH(par_iterate_smem_kernel<4>, grid, block, 4 * kSmemH * kSmemW * sizeof(float), stream, aff, src, lay,
                    dst, lo, nch_dev, nch_uniform, c_stride, h, w, n_dil);
      }
    } else if (vec) {
      const int tq = par_tile_log2();
      dim3 grid(ceil_div(w / 4, 1 << tq), ceil_div(h, 256 >> tq), B), block(256);
      if (wide) {
        COSA_LAUNCH(par_iterate_vec_kernel<8>, grid, block, 0, stream, aff, src, lay, dst, lo, nch_dev, nch_uniform,
                    c_stride, h, w, n_dil, tq);
      } else {
        COSA_LAUNCH(par_iterate_vec_kernel<4>, grid, block, 0, stream, aff, src, lay, dst, lo, nch_dev, nch_uniform,
                    c_stride, h, w, n_dil, tq);
      }
    } else {
      if (lo.pitch != w || lo.off != 0) return COSA_E_ARG;   // the generic kernel writes plain NCHW only
      dim3 grid(ceil_div(w, 32), ceil_div(h, 8), B), block(256);
      if (wide) {
        COSA_LAUNCH(par_iterate_kernel<8>, grid, block, 0, stream, aff, src, dst, nch_dev, nch_uniform, c_stride, h,
                    w, n_dil);
      } else {
        COSA_LAUNCH(par_iterate_kernel<4>, grid, block, 0, stream, aff, src, dst, nch_dev, nch_uniform, c_stride, h,
                    w, n_dil);
      }
    }
    src = dst;
  }
  return 0;
}

// Affinity + num_iter propagation steps for the whole batch.  (Splitting the batch into L2-sized chunks so that
// the affinity planes are re-read from L2 was measured and is slower: the step is bound by the L1 path of the
// neighbour loads, not by the affinity stream - see profiles/README.md.)
int par_refine_batch(const float *imgs, float *aff, const float *src0, float *scratch_a, float *scratch_b,
                     MaskLayout lay, float *final_dst, MaskLayout lay_final, const int *nch_dev, int nch_uniform,
                     int c_stride, int B, int h, int w, int n_dil, int num_iter, cudaStream_t stream) {
  COSA_CHECK(par_launch_affinity(imgs, aff, B, h, w, n_dil, stream));
  return par_launch_iterations(aff, src0, scratch_a, scratch_b, lay, final_dst, lay_final, nch_dev, nch_uniform,
                               c_stride, B, h, w, n_dil, num_iter, stream);
}

}  // namespace cosa

using namespace cosa;

// scratch: affinity [B,8*n_dil,h,w] + two mask buffers in the widest layout cosa_par_forward uses
// (column pads of at most 24: wider dilations take the plain layout and the generic kernel)
extern "C" size_t cosa_par_ws_bytes(int B, int C, int h, int w, int n_dil) {
  const size_t plane = (size_t)h * w;
  const size_t pitch = (size_t)max_padded_pitch(w);
  return align_up((size_t)B * 8 * n_dil * plane * sizeof(float), 256) +
         2 * align_up((size_t)B * C * h * pitch * sizeof(float), 256);
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
  MaskLayout lay = padded_layout(w, dilations, n_dil);
  if (lay.padn > 24) lay = plain_layout(w);   // keeps the scratch within cosa_par_ws_bytes
  const MaskLayout plain = plain_layout(w);
  Arena arena(ws);
  float *aff = arena.take<float>((size_t)B * 8 * n_dil * plane);
  float *buf_a = arena.take<float>(layout_floats(lay, B, C, h));
  float *buf_b = arena.take<float>(layout_floats(lay, B, C, h));
  const bool resize = (hm != h || wm != w);
  const long long total = (long long)B * C * plane;
  const int blocks = (int)max(1LL, min((long long)sm_count() * 8, ceil_div_ll(total, 256)));
  if (num_iter == 0) {   // PAR.py:66 alone
    if (resize) {
      const float sy = h > 1 ? (float)(hm - 1) / (float)(h - 1) : 0.0f;
      const float sx = w > 1 ? (float)(wm - 1) / (float)(w - 1) : 0.0f;
      COSA_LAUNCH(resize_align_corners_kernel, blocks, 256, 0, s, masks_in, masks_out, plain, B * C, hm, wm, h, w, sy,
                  sx);
    } else {
      COSA_CUDA(cudaMemcpyAsync(masks_out, masks_in, total * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    return 0;
  }
  // stage the input in the iteration layout: buf_b (then steps alternate buf_a / buf_b, the last one writes out)
  const float *src0 = masks_in;
  if (resize) {
    const float sy = h > 1 ? (float)(hm - 1) / (float)(h - 1) : 0.0f;
    const float sx = w > 1 ? (float)(wm - 1) / (float)(w - 1) : 0.0f;
    COSA_LAUNCH(resize_align_corners_kernel, blocks, 256, 0, s, masks_in, buf_b, lay, B * C, hm, wm, h, w, sy, sx);
    src0 = buf_b;
  } else if (lay.padn > 0) {
    COSA_CHECK(par_launch_pack(masks_in, buf_b, lay, B * C, h, w, s));
    src0 = buf_b;
  }
  // step 0 reads src0 (buf_b or the caller's tensor) and writes buf_a, step 1 writes buf_b, ...
  // (src0 is the caller's tensor only in the plain layout, whose strides equal the scratch strides)
  return par_refine_batch(imgs, aff, src0, buf_a, buf_b, lay, masks_out, plain, nullptr, C, C, B, h, w, n_dil, num_iter,
                          s);
}
