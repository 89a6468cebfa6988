// PAR — pixel-adaptive refinement (reference: models/PAR.py:26-91) as sm_100a kernels.
//
//   par_affinity_kernel : A[b,n,y,x] = softmax_n( mean_c -(|I_c(nbr_n)-I_c|/(std48_c+1e-8)/w1)^2 )
//                                      + w2 * softmax_n( -(pos_n/(std(pos)+1e-8)/w1)^2 )          (PAR.py:69-85)
//   par_iterate_kernel  : M'[b,c,y,x] = sum_n A[b,n,y,x] * M[b,c,clamp(y+dy_n),clamp(x+dx_n)]       (PAR.py:87-89)
//   resize_align_corners_kernel : masks -> image size, bilinear align_corners=True                 (PAR.py:66)
//
// Neighbour n = 8*k + m: dilation k in ctor order, direction m in get_kernel() order (PAR.py:10-24):
// (-,-) (-,0) (-,+) (0,-) (0,+) (+,-) (+,0) (+,+), borders replicated (PAR.py:44).
//
// Step kernels: `par_chain_kernel` (the reference dilation set, TMA-staged 32 x 32 tiles, ALL steps in one launch
// chained by tile-level step counters: north_star's single launch, the default), `par_iterate_tile_kernel` (the same
// tiles, one launch per step), `par_iterate_smem_kernel` / `par_iterate_kernel` (any dilation set: padded rows with
// the near neighbourhood staged / plain NCHW).  The variants that were measured and dropped (scalar 512-thread tiles,
// double-buffered persistent CTAs, a cooperative launch with grid barriers, deeper register prefetch, rolling
// affinity refill, L2 tensor prefetch of the affinity tile) are described in DESIGN.md section 4.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "par.cuh"

namespace cosa {

// The dilation list and the constant position term (PAR.py:51-62,77,82), evaluated on the host in double and handed to
// the kernels BY VALUE: no __constant__ upload, hence nothing to order between streams and nothing cached per device.
int par_make_constants(const int *dilations, int n_dil, ParConst *pc) {
  if (!dilations || n_dil < 1 || n_dil > kMaxDil) return COSA_E_ARG;
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
  // row_sum: sum over the neighbours of softmax + w2 * softmax(pos_aff), from the float constants the kernels use
  pc->row_sum = 1.0;
  for (int n = 0; n < kMaxDil * 8; ++n) pc->pos_term[n] = 0.0f;
  for (int n = 0; n < nd; ++n) {
    pc->pos_term[n] = 0.01f * (float)(exp(logit[n] - mx) / sum);
    pc->row_sum += (double)pc->pos_term[n];
  }
  for (int k = 0; k < kMaxDil; ++k) pc->dil[k] = k < n_dil ? dilations[k] : 0;
  pc->n_dil = n_dil;
  static const int kStd[6] = {1, 2, 4, 8, 12, 24};   // the reference's list (PAR.py:94): the compile-time tile kernels
  pc->std_dilations = n_dil == 6;
  for (int k = 0; k < 6 && pc->std_dilations; ++k) pc->std_dilations = pc->dil[k] == kStd[k];
  return 0;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) belongs to the current device's context: once per (kernel, device).
template <typename F>
static int opt_in_smem(F kernel, int bytes, std::atomic<unsigned long long> &done) {
  int dev = 0;
  COSA_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = 1ULL << (dev & 63);
  if (dev < 64 && (done.load(std::memory_order_acquire) & bit)) return 0;
  COSA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (dev < 64) done.fetch_or(bit, std::memory_order_release);
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Affinity.  One thread per pixel; per colour channel the 8*NDIL neighbour differences live in registers
// (two-pass unbiased variance: the one-pass form cancels catastrophically in flat regions).
// ------------------------------------------------------------------------------------------------
// x / 3 correctly rounded without the division routine (MUFU.RCP + Newton + FCHK + slow-path call, ~12 issue slots and
// a branch per quotient; the affinity kernels need one per neighbour): q = RN(x * RN(1/3)) is within one ulp of the
// quotient, r = x - 3q is exact in an FMA, and RN(q + r * RN(1/3)) is the correctly rounded x / 3 - checked against
// IEEE division for every mantissa of two binades.  (Quotients in the denormal range may differ in the last bit.)
__device__ __forceinline__ float div3_rn(float x) {
  const float c = 0.333333343267440796f;   // RN(1/3) = 0x3EAAAAAB
  const float q = __fmul_rn(x, c);
  return __fmaf_rn(__fmaf_rn(-3.0f, q, x), c, q);
}

template <int NDIL>
__global__ void __launch_bounds__(128) par_affinity_kernel(const float *__restrict__ imgs, float *__restrict__ aff,
                                                           int h, int w, const ParConst pc) {
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
      const int d = pc.dil[k];
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
    logit[n] = div3_rn(-logit[n]);
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
  for (int n = 0; n < ND; ++n) out[(size_t)n * plane] = fmaf(logit[n], rden, pc.pos_term[n]);
}

// Generic (any n_dil <= kMaxDil) three-pass variant: neighbours are re-read instead of kept in registers.
__global__ void __launch_bounds__(128) par_affinity_generic_kernel(const float *__restrict__ imgs,
                                                                   float *__restrict__ aff, int h, int w, int n_dil,
                                                                   const ParConst pc) {
  const int nd = 8 * n_dil;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 4 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (x >= w || y >= h) return;
  const size_t plane = (size_t)h * w;
  const float *img = imgs + (size_t)b * 3 * plane;
  float *out = aff + (size_t)b * nd * plane + (size_t)y * w + x;
  auto nbr = [&](const float *ch, int n) {
    const int d = pc.dil[n >> 3], m = n & 7;
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
    l = div3_rn(-l);
    out[(size_t)n * plane] = l;
    mx = fmaxf(mx, l);
  }
  float den = 0.0f;
  for (int n = 0; n < nd; ++n) den += expf(out[(size_t)n * plane] - mx);
  const float rden = 1.0f / den;
  for (int n = 0; n < nd; ++n)
    out[(size_t)n * plane] = fmaf(expf(out[(size_t)n * plane] - mx), rden, pc.pos_term[n]);
}

// ------------------------------------------------------------------------------------------------
// One propagation step, generic form (any width, plain NCHW layout).  Thread per pixel, CH mask channels per
// pass in registers; clamped scalar neighbour loads through L1.
// nch_dev (optional) gives the number of live channels per image for the ragged cam2mask batch.
// ------------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256) par_iterate_kernel(const float *__restrict__ aff, const float *__restrict__ in,
                                                          float *__restrict__ out, const int *__restrict__ nch_dev,
                                                          int nch_uniform, int c_stride, int h, int w, int n_dil,
                                                          const ParConst pc) {
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
      const int d = pc.dil[kd];
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
// Vectorised step kernels (w % 4 == 0, padded rows).  A thread owns 4 horizontally adjacent pixels and CH channels:
// per dilation it streams the 8 affinity quads (128-bit, no L1 allocation) and, per channel and neighbour row, three
// aligned 128-bit loads that cover the column offsets -d, 0, +d of all four pixels (d % 4 == 0: the quads at x-d, x,
// x+d; d < 4: the quads left/centre/right, recombined in registers).  The replicated column pads make every load
// unclamped; rows are clamped by index.
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
// ------------------------------------------------------------------------------------------------
// One propagation step with the near neighbourhood staged in shared memory by the TMA unit.
//
// CTA tile: 32 rows x 32 pixels (256 threads, 4 pixels x CH channels each).  The mask tile
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
                            int c_stride, int h, int w, int n_dil, const ParConst pc) {
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
      const int d = pc.dil[kd];
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

// ------------------------------------------------------------------------------------------------
// One propagation step for the reference dilation set {1,2,4,8,12,24} with the WHOLE 24-pixel neighbourhood staged in
// shared memory (every pixel-step consumes 48 x channels neighbour values; from shared memory every 128-bit load
// costs exactly 4 crossbar cycles, a misaligned global load 5-6 and most far-neighbour loads missed L1).
//
// CTA = 32 x 32 pixels (256 threads, one quad x CH channels each).  The (32+48)^2 halo tile of each staged channel
// arrives as ONE TMA box on its own mbarrier (rows outside the image are zero-filled and repaired in place = replicate
// border; columns come from the replicated pads), so the first dilation of channel 0 starts when 25 KB have landed
// and the other channels land underneath.  The dilations are compile-time constants: every shared-memory offset is an
// immediate and the d = 1, 2 quad recombination is register renaming.  Affinity quads are streamed straight into
// registers one dilation ahead of their use (tile_pass).
// ------------------------------------------------------------------------------------------------
constexpr int kFH = 24;                       // halo of the full-neighbourhood tile
constexpr int kFS = kTileW + 2 * kFH;         // 80 floats per staged row, 80 rows
constexpr unsigned kTileBytes = kFS * kFS * sizeof(float);

// WAIT: channel k's tile is awaited right before its first use (the first dilation of a pass)
// LIVE > 0: the number of live channels is a compile-time constant (the channels' loads and FMAs form one basic block
// that ptxas can interleave: chained kernel 0.684 -> 0.657 ms); LIVE = 0: channel k is guarded by k < live.
template <int D, int CH, bool WAIT = false, int LIVE = 0>
__device__ __forceinline__ void tile_dilation(float4 (&acc)[CH], const float4 (&a)[8], const float *q, int live,
                                              unsigned long long *bars = nullptr, unsigned phase = 0) {
  constexpr int R = D * kFS;
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    if (LIVE ? k < LIVE : k < live) {
      if constexpr (WAIT) mbar_wait(&bars[k], phase);
      const float *p = q + k * (kFS * kFS);
      if constexpr ((D & 3) == 0) {
        fma4(acc[k], a[0], lds4(p - R - D)); fma4(acc[k], a[1], lds4(p - R)); fma4(acc[k], a[2], lds4(p - R + D));
        fma4(acc[k], a[3], lds4(p - D));                                      fma4(acc[k], a[4], lds4(p + D));
        fma4(acc[k], a[5], lds4(p + R - D)); fma4(acc[k], a[6], lds4(p + R)); fma4(acc[k], a[7], lds4(p + R + D));
      } else {
        float4 m, pl, C;
        C = lds4(p - R); shifted_quads(lds4(p - R - 4), C, lds4(p - R + 4), D, m, pl);
        fma4(acc[k], a[0], m); fma4(acc[k], a[1], C); fma4(acc[k], a[2], pl);
        C = lds4(p); shifted_quads(lds4(p - 4), C, lds4(p + 4), D, m, pl);
        fma4(acc[k], a[3], m); fma4(acc[k], a[4], pl);
        C = lds4(p + R); shifted_quads(lds4(p + R - 4), C, lds4(p + R + 4), D, m, pl);
        fma4(acc[k], a[5], m); fma4(acc[k], a[6], C); fma4(acc[k], a[7], pl);
      }
    }
  }
}

// Rows -D and +D of dilation D (plane group a[0..5] = (-,-) (-,0) (-,+) (+,-) (+,0) (+,+)); D = 1, 2, 4.
// WAIT: channel k's tile is awaited right before its first use (the first group of a pass).
template <int D, int CH, bool WAIT, int LIVE>
__device__ __forceinline__ void tile_rows(float4 (&acc)[CH], const float4 (&a)[8], const float *q, int live,
                                          unsigned long long *bars, unsigned phase) {
  constexpr int R = D * kFS;
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    if (LIVE ? k < LIVE : k < live) {
      if constexpr (WAIT) mbar_wait(&bars[k], phase);
      const float *p = q + k * (kFS * kFS);
      if constexpr (D == 4) {
        fma4(acc[k], a[0], lds4(p - R - 4)); fma4(acc[k], a[1], lds4(p - R)); fma4(acc[k], a[2], lds4(p - R + 4));
        fma4(acc[k], a[3], lds4(p + R - 4)); fma4(acc[k], a[4], lds4(p + R)); fma4(acc[k], a[5], lds4(p + R + 4));
      } else {
        float4 m, pl, C;
        C = lds4(p - R); shifted_quads(lds4(p - R - 4), C, lds4(p - R + 4), D, m, pl);
        fma4(acc[k], a[0], m); fma4(acc[k], a[1], C); fma4(acc[k], a[2], pl);
        C = lds4(p + R); shifted_quads(lds4(p + R - 4), C, lds4(p + R + 4), D, m, pl);
        fma4(acc[k], a[3], m); fma4(acc[k], a[4], C); fma4(acc[k], a[5], pl);
      }
    }
  }
}

// The centre-row taps of d = 1, 2, 4 (plane group a[0..5] = d1 (0,-) (0,+), d2 (0,-) (0,+), d4 (0,-) (0,+)) from ONE
// read of the aligned quads p - 4, p, p + 4.
template <int CH, int LIVE>
__device__ __forceinline__ void tile_centre(float4 (&acc)[CH], const float4 (&a)[8], const float *q, int live) {
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    if (LIVE ? k < LIVE : k < live) {
      const float *p = q + k * (kFS * kFS);
      const float4 L = lds4(p - 4), C = lds4(p), Rr = lds4(p + 4);
      float4 m, pl;
      shifted_quads(L, C, Rr, 1, m, pl);
      fma4(acc[k], a[0], m); fma4(acc[k], a[1], pl);
      shifted_quads(L, C, Rr, 2, m, pl);
      fma4(acc[k], a[2], m); fma4(acc[k], a[3], pl);
      fma4(acc[k], a[4], L); fma4(acc[k], a[5], Rr);
    }
  }
}

// Tile-permuted affinity layout of the tile step kernels: [B][48][tiles][8 warps][4 j][32 lanes], the value of pixel
// (4 tq + j, 4 wi + r) of a 32 x 32 tile at wi*128 + j*32 + (r*8 + tq) - the j-th pixel of the quad that lane r*8 + tq
// of warp wi owns.  A warp then fetches one neighbour plane of its 32 quads with FOUR fully coalesced 128-byte
// LDG.32 (one L1 wavefront each) instead of one LDG.128 that spans four lines: the L1 pipeline replays a multi-line
// request at half rate (8.3 against 4 cycles per 512 bytes), and the 48 affinity loads of a pass were as expensive
// on that pipeline as its 100 shared-memory loads.
__host__ __device__ __forceinline__ int aff_tile_offset(int px, int py) {
  return (py >> 2) * 128 + (px & 3) * 32 + (py & 3) * 8 + (px >> 2);
}
// Plane order of the tile-permuted layout: the neighbour (8 * dilation + direction, PAR.py:10-24) stored in plane m.
// The centre-row taps (0, -d) and (0, +d) of d = 1, 2, 4 all read the aligned quads p - 4, p, p + 4 of the mask tile, so
// they form a group of their own and those three loads are made once instead of three times (45 instead of 50
// shared-memory loads per pixel quad and channel):
//   planes  0.. 5  d = 1, rows -1 / +1     6..11  d = 2, rows -2 / +2     12..17  d = 4, rows -4 / +4
//   planes 18..23  the centre-row taps of d = 1, 2, 4                      24..47  d = 8, 12, 24 in neighbour order
__host__ __device__ constexpr int aff_plane_neighbour(int m) {
  return m >= 24 ? m
         : m >= 18 ? 8 * ((m - 18) >> 1) + 3 + ((m - 18) & 1)
                   : 8 * (m / 6) + (m % 6 < 3 ? m % 6 : m % 6 + 2);
}

__device__ __forceinline__ float ldg_stream1(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// the affinity quads of one plane group (N = 6 or 8 planes); A (this lane's slot in the tile-permuted plane) walks
// through the planes with one live 64-bit address instead of 48
template <int N>
__device__ __forceinline__ void load_aff(float4 (&a)[8], const float *&A, size_t plane) {
#pragma unroll
  for (int m = 0; m < N; ++m) {
    a[m] = make_float4(ldg_stream1(A), ldg_stream1(A + 32), ldg_stream1(A + 64), ldg_stream1(A + 96));
    // opaque increment: keeps ptxas from materialising (and spilling) all 48 plane addresses up front
    asm volatile("add.u64 %0, %0, %1;" : "+l"(A) : "l"(plane * sizeof(float)));
  }
}

// One instruction per warp requests the 32 lines (8 planes x 4 x 128 bytes) of this warp's affinity quads of a LATER
// dilation into L2: lane l takes line (l >> 2, l & 3).  ptxas sinks the register loads of a set down to the dilation
// that uses them (register pressure), so without this every dilation starts with a DRAM round trip (chain kernel:
// 0.757 -> 0.724 ms for the ten steps).
// A = this lane's slot in the first plane of that dilation.
__device__ __forceinline__ void prefetch_aff8(const float *A, size_t plane, int n_planes = 8) {
  const int lane = threadIdx.x & 31;
  const float *p = A - lane + (size_t)(lane >> 2) * plane + (lane & 3) * 32;
  if ((lane >> 2) < n_planes) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Orders the refill of one affinity register set behind (a) the arrival of the OTHER set and (b) the end of the
// dilation that used the registers being refilled.  ptxas tracks all global loads of this kernel on ONE scoreboard (the
// other five pipeline the shared-memory loads), and a scoreboard wait drains everything issued on it: with a refill
// already in flight, the first FMA of every dilation waited a full memory latency for loads it did not need (ncu r02l:
// 17 % of all warp stalls).  Making the refill's address depend on a value of the set used next and on the
// accumulators of the dilation just finished (the select is never taken: affinities and masks are finite) puts the
// wait for the next set - the only loads in flight, issued a whole dilation earlier - at the dilation boundary and
// the refill right behind it.
template <int CH>
__device__ __forceinline__ const float *issue_after(const float *A, const float4 &next, const float4 (&acc)[CH]) {
  unsigned bits = __float_as_uint(next.x);
#pragma unroll
  for (int k = 0; k < CH; ++k) bits &= __float_as_uint(acc[k].w);
  return A + (bits == 0xffffffffu ? 4 : 0);
}

__device__ __forceinline__ void tma_load_box(void *dst_smem, const CUtensorMap *tmap, int x, int y, int z,
                                             unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
          "r"(smem_u32(dst_smem)),
      "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}
constexpr int kBoxRows = 16;

// rows [0, r_lo) <- row r_lo, rows [r_hi, 80) <- row r_hi - 1 in every staged channel (whole CTA; ends with a barrier)
__device__ __forceinline__ void replicate_border_rows(float *tile, int live, int r_lo, int r_hi) {
  constexpr int Q = kFS / 4;
  const int bad = r_lo + (kFS - r_hi);
  for (int i = threadIdx.x; i < live * bad * Q; i += blockDim.x) {
    const int k = i / (bad * Q), j = i - k * (bad * Q);
    const int rr = j / Q, c4 = j - rr * Q;
    const int r = rr < r_lo ? rr : r_hi + (rr - r_lo);
    const int from = rr < r_lo ? r_lo : r_hi - 1;
    float4 *base = reinterpret_cast<float4 *>(tile + (size_t)k * kFS * kFS);
    base[r * Q + c4] = base[from * Q + c4];
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Affinity for the reference dilation set from a TMA-staged image tile: the (32+48)^2 halo tile of the three colour
// planes sits in shared memory (replicate border repaired in place), so the 48 x 3 neighbour reads are conflict-free
// shared-memory loads at immediate offsets instead of clamped global loads with per-neighbour index arithmetic.
// 256 threads; a thread produces the 4 pixels (tx, ty + 8 r) one after the other, arithmetic identical (operation
// for operation) to par_affinity_kernel<6>.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void replicate_border(float *tile, int planes, int r_lo, int r_hi, int c_lo, int c_hi) {
  // columns first (valid rows only), then whole rows
  if (c_lo > 0 || c_hi < kFS) {
    const int badc = c_lo + (kFS - c_hi), rows = r_hi - r_lo;
    for (int i = threadIdx.x; i < planes * rows * badc; i += 256) {
      const int k = i / (rows * badc), j = i - k * (rows * badc);
      const int r = r_lo + j / badc, cc = j % badc;
      const int c = cc < c_lo ? cc : c_hi + (cc - c_lo);
      float *row = tile + ((size_t)k * kFS + r) * kFS;
      row[c] = row[cc < c_lo ? c_lo : c_hi - 1];
    }
    __syncthreads();
  }
  if (r_lo > 0 || r_hi < kFS) replicate_border_rows(tile, planes, r_lo, r_hi);
}

template <bool PERM>
__global__ void __launch_bounds__(256, 2)
    par_affinity_tile_kernel(const __grid_constant__ CUtensorMap tm_img, float *__restrict__ aff, int h, int w,
                             const ParConst pc) {
  extern __shared__ __align__(128) float s_tile[];   // [3][80][80]
  __shared__ __align__(8) unsigned long long s_bar;
  constexpr int ND = 48;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH, b = blockIdx.z;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_expect_tx(&s_bar, (unsigned)(3 * kFS * kFS * sizeof(float)));
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < kFS; r += kBoxRows)
        tma_load_box(s_tile + (c * kFS + r) * kFS, &tm_img, x0 - kFH, y0 - kFH + r, b * 3 + c, &s_bar);
  }
  __syncthreads();
  mbar_wait(&s_bar, 0);
  const int r_lo = max(0, kFH - y0), r_hi = min(kFS, h - y0 + kFH);
  const int c_lo = max(0, kFH - x0), c_hi = min(kFS, w - x0 + kFH);
  if (r_lo > 0 || r_hi < kFS || c_lo > 0 || c_hi < kFS) replicate_border(s_tile, 3, r_lo, r_hi, c_lo, c_hi);

  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const size_t plane = (size_t)h * w;
  constexpr int kD[6] = {1, 2, 4, 8, 12, 24};
#pragma unroll 1
  for (int rr = 0; rr < 4; ++rr) {
    const int yl = ty + 8 * rr;
    const int x = x0 + tx, y = y0 + yl;
    if (x >= w || y >= h) continue;
    // packed fp32: neighbour pairs (2i, 2i + 1) share one FADD2 / FMUL2 / FFMA2
    float2 logit2[ND / 2];
#pragma unroll
    for (int i = 0; i < ND / 2; ++i) logit2[i] = make_float2(0.0f, 0.0f);
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      const float *q = s_tile + (c * kFS + kFH + yl) * kFS + kFH + tx;
      const float ctr = q[0];
      float2 v2[ND / 2];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int d = kD[k], R = kD[k] * kFS;
        v2[4 * k + 0] = make_float2(q[-R - d], q[-R]);
        v2[4 * k + 1] = make_float2(q[-R + d], q[-d]);
        v2[4 * k + 2] = make_float2(q[d], q[R - d]);
        v2[4 * k + 3] = make_float2(q[R], q[R + d]);
      }
      float2 s2 = v2[0];
#pragma unroll
      for (int i = 1; i < ND / 2; ++i) s2 = fadd2(s2, v2[i]);
      const float mean = div3_rn(s2.x + s2.y) * 0.0625f;   // sum / 48, correctly rounded (ND = 48)
      const float2 nmean = splat2(-mean);
      float2 ss2 = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int i = 0; i < ND / 2; ++i) {
        const float2 t = fadd2(v2[i], nmean);
        ss2 = ffma2(t, t, ss2);
      }
      const float sd = sqrtf((ss2.x + ss2.y) / (float)(ND - 1));
      const float2 inv = splat2(1.0f / ((sd + 1e-8f) * 0.3f)), nctr = splat2(-ctr);
#pragma unroll
      for (int i = 0; i < ND / 2; ++i) {
        const float2 t = fmul2(fadd2(v2[i], nctr), inv);   // (|v - ctr| * inv)^2 == ((v - ctr) * inv)^2 bit for bit
        logit2[i] = ffma2(t, t, logit2[i]);
      }
    }
    // aff = -(sum_c t^2) / 3 (div3_rn, packed: q = RN(x c), RN(q + RN(x - 3 q) c)); softmax over the 48 neighbours
    float logit[ND];
    float mx = -INFINITY;
    {
      const float2 c3 = splat2(0.333333343267440796f), m3 = splat2(-3.0f);
#pragma unroll
      for (int i = 0; i < ND / 2; ++i) {
        const float2 xneg = make_float2(-logit2[i].x, -logit2[i].y);
        const float2 qq = fmul2(xneg, c3);
        const float2 r = ffma2(ffma2(m3, qq, xneg), c3, qq);
        logit[2 * i] = r.x;
        logit[2 * i + 1] = r.y;
        mx = fmaxf(mx, fmaxf(r.x, r.y));
      }
    }
    float den = 0.0f;
#pragma unroll
    for (int n = 0; n < ND; ++n) {
      logit[n] = expf(logit[n] - mx);
      den += logit[n];
    }
    const float rden = 1.0f / den;
    // PERM: the tile-permuted layout the tile step kernels read (aff_tile_offset); else plain [B, 48, h, w]
    const size_t stride = PERM ? (size_t)gridDim.x * gridDim.y * (kTileW * kTileH) : plane;
    float *out = PERM ? aff + ((size_t)b * ND * gridDim.x * gridDim.y + (size_t)blockIdx.y * gridDim.x + blockIdx.x) *
                                  (kTileW * kTileH) + aff_tile_offset(tx, yl)
                      : aff + (size_t)b * ND * plane + (size_t)y * w + x;
#pragma unroll
    for (int m = 0; m < ND; ++m) {
      const int n = PERM ? aff_plane_neighbour(m) : m;   // PERM: planes in the order the tile step kernels consume them
      *out = fmaf(logit[n], rden, pc.pos_term[n]);
      // walk the planes with one opaque 64-bit add: `out[n * plane]` costs a wide multiply and four more integer
      // instructions per store in a kernel that is bound by instruction issue
      asm volatile("add.u64 %0, %0, %1;" : "+l"(out) : "l"(stride * sizeof(float)));
    }
  }
}

// One pass of a CTA over its tile: <= CH staged channels x the six dilations.  Affinity sets alternate between two
// register sets; a refill is issued behind the wait for the set used next (issue_after), one dilation ahead of its use.
template <int CH, int LIVE>
__device__ __forceinline__ void tile_pass(float4 (&acc)[CH], const float *A, size_t plane, const float *q, int live,
                                          float *s_tile, unsigned long long *bars, unsigned phase, int r_lo, int r_hi) {
#pragma unroll
  for (int k = 0; k < CH; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 a0[8], a1[8];
  load_aff<6>(a0, A, plane);   // rows of d = 1
  load_aff<6>(a1, A, plane);   // rows of d = 2 (in flight together with the first group: the first wait drains both)
  prefetch_aff8(A, plane);                // planes 12..19
  prefetch_aff8(A + 8 * plane, plane);    // planes 20..27
  if (r_lo > 0 || r_hi < kFS) {   // top / bottom tiles: every channel must have landed before the border rows are repaired
    for (int k = 0; k < live; ++k) mbar_wait(&bars[k], phase);
    replicate_border_rows(s_tile, live, r_lo, r_hi);
  }
  tile_rows<1, CH, true, LIVE>(acc, a0, q, live, bars, phase);
  load_aff<6>(a0, A, plane);   // rows of d = 4: nothing else is in flight
  prefetch_aff8(A + 10 * plane, plane);   // planes 28..35
  tile_rows<2, CH, false, LIVE>(acc, a1, q, live, bars, phase);
  A = issue_after<CH>(A, a0[0], acc);
  load_aff<6>(a1, A, plane);   // centre taps of d = 1, 2, 4
  prefetch_aff8(A + 12 * plane, plane);   // planes 36..43
  tile_rows<4, CH, false, LIVE>(acc, a0, q, live, bars, phase);
  A = issue_after<CH>(A, a1[0], acc);
  load_aff<8>(a0, A, plane);   // d = 8
  prefetch_aff8(A + 12 * plane, plane, 4);   // planes 44..47: the last ones of this image
  tile_centre<CH, LIVE>(acc, a1, q, live);
  A = issue_after<CH>(A, a0[0], acc);
  load_aff<8>(a1, A, plane);   // d = 12
  tile_dilation<8, CH, false, LIVE>(acc, a0, q, live);
  A = issue_after<CH>(A, a1[0], acc);
  load_aff<8>(a0, A, plane);   // d = 24
  tile_dilation<12, CH, false, LIVE>(acc, a1, q, live);
  tile_dilation<24, CH, false, LIVE>(acc, a0, q, live);
}

// the pass with the live-channel count as a compile-time constant (2, 3 or CH channels), else the guarded form
template <int CH>
__device__ __forceinline__ void tile_pass_dispatch(float4 (&acc)[CH], const float *A, size_t plane, const float *q,
                                                   int live, float *s_tile, unsigned long long *bars, unsigned phase,
                                                   int r_lo, int r_hi) {
  if (live == CH) tile_pass<CH, CH>(acc, A, plane, q, live, s_tile, bars, phase, r_lo, r_hi);
  else if (CH > 3 && live == 3) tile_pass<CH, (CH > 3 ? 3 : CH)>(acc, A, plane, q, live, s_tile, bars, phase, r_lo, r_hi);
  else if (CH > 2 && live == 2) tile_pass<CH, (CH > 2 ? 2 : CH)>(acc, A, plane, q, live, s_tile, bars, phase, r_lo, r_hi);
  else tile_pass<CH, 0>(acc, A, plane, q, live, s_tile, bars, phase, r_lo, r_hi);
}

// one thread: stage `live` channel tiles (planes gz .. gz + live - 1, origin (gx, gy)), one TMA box and barrier each
__device__ __forceinline__ void stage_tiles(float *s_tile, const CUtensorMap *tm, int gx, int gy, int gz, int live,
                                            unsigned long long *bars) {
  for (int k = 0; k < live; ++k) {
    mbar_expect_tx(&bars[k], kTileBytes);
    tma_load_box(s_tile + (size_t)k * kFS * kFS, tm, gx, gy, gz + k, &bars[k]);
  }
}

template <int CH>
__device__ __forceinline__ void store_quads(const float4 (&acc)[CH], float *dst, size_t oplane, int c0, int live,
                                            const MaskLayout &lo, int xq, int wq) {
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

// Per-step kernel ("tile" mode): one CTA per (image tile, channel split g of gsplit); the CTA walks the channel
// passes of its tile.  The hardware CTA scheduler balances the load; 2 CTAs per SM overlap staging and compute.
// blockIdx.z = image * gsplit + g: gsplit CTAs share an image tile and take the channel groups g, g + gsplit, ...
template <int CH>
__global__ void __launch_bounds__(256, 2)
    par_iterate_tile_kernel(const float *__restrict__ aff, const __grid_constant__ CUtensorMap tmap_in, MaskLayout li,
                            float *__restrict__ out, MaskLayout lo, const int *__restrict__ nch_dev, int nch_uniform,
                            int c_stride, int h, int w, int gsplit) {
  extern __shared__ __align__(128) float s_tile[];   // [CH][80][80]
  __shared__ __align__(8) unsigned long long s_bar[CH];
  const int wq = w >> 2;
  const int tq = threadIdx.x & 7, tr = threadIdx.x >> 3;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const int xq = (x0 >> 2) + tq, y = y0 + tr, x = x0 + (tq << 2);
  const int b = blockIdx.z / gsplit, g = blockIdx.z - b * gsplit;
  const bool active = xq < wq && y < h;
  const int nch = nch_dev ? nch_dev[b] : nch_uniform;
  if (nch <= 0) return;   // an image without live channels (cam2mask: no foreground class)
  // up to CH channels: one CTA (the affinity quads are loaded once); more: split evenly over the gsplit CTAs of
  // this tile, at most CH per pass
  const int n_groups = nch <= CH ? 1 : gsplit * ((nch + gsplit * CH - 1) / (gsplit * CH));
  const int chunk = (nch + n_groups - 1) / n_groups;
  if (g * chunk >= nch) return;
  const size_t plane = (size_t)gridDim.x * gridDim.y * (kTileW * kTileH);   // tile-permuted affinity (aff_tile_offset)
  const size_t oplane = (size_t)h * lo.pitch;
  const float *A = aff + ((size_t)b * 48 * gridDim.x * gridDim.y + (size_t)blockIdx.y * gridDim.x + blockIdx.x) *
                             (kTileW * kTileH) + (threadIdx.x >> 5) * 128 + (threadIdx.x & 31);
  float *dst = out + (size_t)b * c_stride * oplane + (size_t)y * lo.pitch + lo.off + x;
  // staged rows [r_lo, r_hi) exist in the image; the others replicate the border row (PAR.py:44)
  const int r_lo = max(0, kFH - y0), r_hi = min(kFS, h - y0 + kFH);
  const float *q = s_tile + (tr + kFH) * kFS + (tq << 2) + kFH;   // this thread's quad in staged channel 0

  // Programmatic dependent launch: the next step's CTAs may become resident as soon as slots free up in this
  // step's last wave (their barrier set-up and first affinity loads do not depend on this step); they wait for this
  // grid's completion (griddepcontrol.wait, below) before they touch the masks.  Without the launch attribute both
  // instructions are no-ops.
  asm volatile("griddepcontrol.launch_dependents;");
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < CH; ++k) mbar_init(&s_bar[k], 1);
  }
  __syncthreads();
  unsigned phase = 0;
  for (int c0 = g * chunk; c0 < nch; c0 += gsplit * chunk) {
    const int live = min(chunk, nch - c0);
    if (threadIdx.x == 0) {
      // the previous step (the producer of the tiles, and the last reader of the buffer this step overwrites) is
      // complete and its writes are visible; every other thread of the CTA is gated by the mbarriers behind this
      asm volatile("griddepcontrol.wait;" ::: "memory");
      stage_tiles(s_tile, &tmap_in, li.off + x0 - kFH, y0 - kFH, b * c_stride + c0, live, s_bar);
    }
    float4 acc[CH];
    tile_pass_dispatch<CH>(acc, A, plane, q, live, s_tile, s_bar, phase, r_lo, r_hi);
    phase ^= 1;
    if (active) store_quads<CH>(acc, dst, oplane, c0, live, lo, xq, wq);
    if (c0 + gsplit * chunk < nch) {   // the tile is re-staged (async proxy) for the next channel group
      __syncthreads();
      if (threadIdx.x == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
}

struct PropArgs {
  const float *aff;              // [B, 48, h, w]
  float *out_a, *out_b, *out_final;
  MaskLayout li, lo_final;       // scratch layout (inputs and intermediate outputs), layout of the last output
  const int *nch_dev;
  int nch_uniform, c_stride, B, h, w;
  int tiles_x, tiles_y, num_iter;
};

// ------------------------------------------------------------------------------------------------
// Chained step kernel (default): ALL num_iter propagation steps in ONE launch - north_star's single launch - without
// grid barriers.  The grid holds one CTA per (step, image, tile) in step-major order inside groups of `group` images;
// a CTA of step it > 0 waits (warp 0 polls with ld.acquire.gpu) until the <= 9 tiles of its image that overlap its
// halo have published step it - 1 in `done` (one counter per image tile: the number of steps that tile has
// completed), then stages its (32+48)^2 tiles by TMA exactly like par_iterate_tile_kernel.  The same <= 9 tiles are
// the only readers of the region this CTA overwrites in the ping-pong buffer (halo 24 < tile 32), so the one wait
// covers the write-after-read hazard too.  CTAs are dispatched in blockIdx order, so every CTA a waiter depends on is
// already resident or finished: no deadlock, and no wave quantisation at the step boundaries - the SMs stay full over
// all steps.  The distance between a tile's consecutive steps (group * tiles CTAs) must stay above the number of
// resident CTAs, or the later one idles its slot until the earlier one is done.
// One CTA walks ALL live channels of its tile, <= CH per pass (the affinity quads are loaded once per pass).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int CH>
__global__ void __launch_bounds__(256, 2)
    par_chain_kernel(const PropArgs p, const __grid_constant__ CUtensorMap tm_src0,
                     const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                     int *__restrict__ done, int group) {
  extern __shared__ __align__(128) float s_tile[];   // [CH][80][80]
  __shared__ __align__(8) unsigned long long s_bar[CH];
  const int tiles = p.tiles_x * p.tiles_y;
  // blockIdx.x -> (image group, step, image of the group, tile)
  int u = blockIdx.x;
  const int per_group = p.num_iter * group * tiles;
  const int grp = u / per_group;
  u -= grp * per_group;
  const int b0 = grp * group, bg = min(group, p.B - b0);
  const int it = u / (bg * tiles);
  u -= it * bg * tiles;
  const int bi = u / tiles, t = u - bi * tiles, b = b0 + bi;
  const int tyi = t / p.tiles_x, txi = t - tyi * p.tiles_x;
  const int nch = p.nch_dev ? p.nch_dev[b] : p.nch_uniform;
  if (nch <= 0) return;   // an image without live channels: none of its CTAs runs, none waits
  const int n_groups = (nch + CH - 1) / CH;
  const int chunk = (nch + n_groups - 1) / n_groups;

  const CUtensorMap *tm = it == 0 ? &tm_src0 : (((it - 1) & 1) ? &tm_b : &tm_a);
  const bool last = it == p.num_iter - 1;
  float *out = last ? p.out_final : ((it & 1) ? p.out_b : p.out_a);
  const MaskLayout lo = last ? p.lo_final : p.li;
  const int h = p.h, w = p.w, wq = w >> 2;
  const int tq = threadIdx.x & 7, tr = threadIdx.x >> 3;
  const int x0 = txi * kTileW, y0 = tyi * kTileH;
  const int xq = (x0 >> 2) + tq, y = y0 + tr, x = x0 + (tq << 2);
  const bool active = xq < wq && y < h;
  const size_t plane = (size_t)tiles * (kTileW * kTileH);   // tile-permuted affinity (aff_tile_offset)
  const size_t oplane = (size_t)h * lo.pitch;
  const float *A = p.aff + ((size_t)b * 48 * tiles + t) * (kTileW * kTileH) + (threadIdx.x >> 5) * 128 + (threadIdx.x & 31);
  float *dst = out + (size_t)b * p.c_stride * oplane + (size_t)y * lo.pitch + lo.off + x;
  const int r_lo = max(0, kFH - y0), r_hi = min(kFS, h - y0 + kFH);
  const float *q = s_tile + (tr + kFH) * kFS + (tq << 2) + kFH;   // this thread's quad in staged channel 0

  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < CH; ++k) mbar_init(&s_bar[k], 1);
  }
  __syncthreads();
  unsigned phase = 0;
  for (int c0 = 0; c0 < nch; c0 += chunk) {
    const int live = min(chunk, nch - c0);
    if (threadIdx.x < 32) {
      if (c0 == 0 && it > 0) {
        // lanes 0..8: one neighbour tile each (the tile itself included)
        const int lane = threadIdx.x;
        if (lane < 9) {
          const int ny = tyi + lane / 3 - 1, nx = txi + lane % 3 - 1;
          if (ny >= 0 && ny < p.tiles_y && nx >= 0 && nx < p.tiles_x) {
            const int *f = done + (size_t)b * tiles + ny * p.tiles_x + nx;
            const long long t0 = clock64();
            while (ld_acquire_gpu(f) < it) {
              __nanosleep(40);
              if (clock64() - t0 > (1LL << 33)) __trap();   // seconds: the dispatch-order assumption failed - fail loudly
            }
          }
        }
        __syncwarp();
      }
      if (threadIdx.x == 0) {
        // the neighbours' generic-proxy stores (acquired above) precede this CTA's async-proxy (TMA) reads
        asm volatile("fence.proxy.async;" ::: "memory");
        stage_tiles(s_tile, tm, p.li.off + x0 - kFH, y0 - kFH, b * p.c_stride + c0, live, s_bar);
      }
    }
    float4 acc[CH];
    tile_pass_dispatch<CH>(acc, A, plane, q, live, s_tile, s_bar, phase, r_lo, r_hi);
    phase ^= 1;
    if (active) store_quads<CH>(acc, dst, oplane, c0, live, lo, xq, wq);
    __syncthreads();   // every shared-memory read and every output store of this pass has been issued
    if (c0 + chunk < nch && threadIdx.x == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (!last && threadIdx.x == 0) {
    // publish: this tile has completed step `it` (bar.sync above + release: its outputs are visible before the
    // counter is; the consumers read them through the async proxy)
    asm volatile("fence.proxy.async;" ::: "memory");
    st_release_gpu(done + (size_t)b * tiles + t, it + 1);
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
// 3-D tensor map over an fp32 buffer [planes, h, pitch] with box {bx, by, bz}, no swizzle, zero fill outside.
// cuTensorMapEncodeTiled is fetched through the runtime (no link against libcuda).
static int make_tmap3(CUtensorMap *map, const float *base, long long planes, int h, int pitch, int bx, int by, int bz) {
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                               const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult q;
    void *fn = nullptr;
    COSA_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return COSA_E_ARG;
    encode = (EncodeFn)fn;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)h, (cuuint64_t)planes};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch * sizeof(float), (cuuint64_t)h * pitch * sizeof(float)};
  const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : COSA_E_ARG;
}

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

// Step-kernel selection (cosa_par_set_step_mode; process-wide, meant for A/B runs and tests - results are identical):
//   chain (default) TMA-tile kernel, every step in ONE launch (north_star's single launch), tile-level dependencies
//                   through step counters, <= 4 channels per CTA pass ("chain<G>": G images per group, default 8)
//   tile            TMA-tile kernel, one CTA per tile and channel split (<= 3 channels per pass), one launch per
//                   step (steps 2..T as programmatic dependent launches)
//   smem            the generic per-step kernel (what non-reference dilation sets use)
enum { kStepTile = 0, kStepSmem = 2, kStepChain = 3 };
static std::atomic<int> g_step_mode{kStepChain};
static std::atomic<int> g_chain_group{0};   // images per group of the chained kernel, 0 = automatic (A/B knob)
static int parse_step_mode(const char *e) {
  if (!e) return -1;
  switch (e[0]) {
    case 't': return kStepTile;
    case 'c': return kStepChain;
    case 's': return kStepSmem;
    default: return -1;
  }
}

bool par_tile_path(const ParConst &pc, const MaskLayout &lay, const MaskLayout &lay_final) {
  return lay.padn > 0 && pc.std_dilations && lay.padn == kFH && (lay_final.padn == 0 || lay_final.padn == kFH) &&
         g_step_mode.load(std::memory_order_relaxed) != kStepSmem;
}

int par_launch_affinity(const ParConst &pc, const float *imgs, float *aff, int B, int h, int w, bool tile_permuted,
                        cudaStream_t stream) {
  dim3 grid(ceil_div(w, 32), ceil_div(h, 4), B), block(128);
  const int n_dil = pc.n_dil;
  if (pc.std_dilations && w % 4 == 0) {
    static std::atomic<unsigned long long> done{0};
    const int smem = 3 * kFS * kFS * (int)sizeof(float);
    CUtensorMap tm;
    COSA_CHECK(make_tmap3(&tm, imgs, (long long)B * 3, h, w, kFS, kBoxRows, 1));
    const dim3 tgrid(ceil_div(w, kTileW), ceil_div(h, kTileH), B);
    if (tile_permuted) {
      static std::atomic<unsigned long long> done_p{0};
      COSA_CHECK(opt_in_smem(par_affinity_tile_kernel<true>, smem, done_p));
      COSA_LAUNCH_T("par_affinity_tile_kernel", par_affinity_tile_kernel<true>, tgrid, 256, smem, stream, tm, aff, h, w, pc);
    } else {
      COSA_CHECK(opt_in_smem(par_affinity_tile_kernel<false>, smem, done));
      COSA_LAUNCH_T("par_affinity_tile_kernel", par_affinity_tile_kernel<false>, tgrid, 256, smem, stream, tm, aff, h, w, pc);
    }
  } else if (tile_permuted) {
    return COSA_E_ARG;   // the tile step kernels exist for the reference dilation set only
  } else if (n_dil == 6) {
    COSA_LAUNCH(par_affinity_kernel<6>, grid, block, 0, stream, imgs, aff, h, w, pc);
  } else {
    COSA_LAUNCH(par_affinity_generic_kernel, grid, block, 0, stream, imgs, aff, h, w, n_dil, pc);
  }
  return 0;
}

int par_launch_pack(const float *src, float *dst, MaskLayout lay, int planes, int h, int w, cudaStream_t stream) {
  const long long total = (long long)planes * h * (w + 2 * lay.padn);
  const int blocks = (int)max(1LL, min((long long)sm_count() * 8, ceil_div_ll(total, 256)));
  COSA_LAUNCH(par_pack_kernel, blocks, 256, 0, stream, src, dst, lay, planes, h, w);
  return 0;
}

constexpr int kStepCH = 3;    // channels per CTA pass of the per-step tile kernels
constexpr int kChainCH = 4;   // channels per CTA pass of the chained kernel

static int par_launch_propagate(const float *aff, const float *src0, float *scratch_a, float *scratch_b, MaskLayout lay,
                                float *final_dst, MaskLayout lay_final, const int *nch_dev, int nch_uniform,
                                int c_stride, int B, int h, int w, int num_iter, int mode, int *tile_flags,
                                cudaStream_t stream) {
  constexpr int CH = kStepCH;
  PropArgs a;
  a.aff = aff;
  a.out_a = scratch_a; a.out_b = scratch_b; a.out_final = final_dst;
  a.li = lay; a.lo_final = lay_final;
  a.nch_dev = nch_dev; a.nch_uniform = nch_uniform; a.c_stride = c_stride;
  a.B = B; a.h = h; a.w = w;
  a.tiles_x = ceil_div(w, kTileW); a.tiles_y = ceil_div(h, kTileH);
  a.num_iter = num_iter;
  const long long planes = (long long)B * c_stride;
  CUtensorMap t0, ta, tb;   // one box = the whole (32+48)^2 halo tile of one channel
  COSA_CHECK(make_tmap3(&t0, src0, planes, h, lay.pitch, kFS, kFS, 1));
  COSA_CHECK(make_tmap3(&ta, scratch_a ? scratch_a : src0, planes, h, lay.pitch, kFS, kFS, 1));
  COSA_CHECK(make_tmap3(&tb, scratch_b ? scratch_b : src0, planes, h, lay.pitch, kFS, kFS, 1));
  if (mode == kStepChain && tile_flags) {   // every step in one launch, tile-level dependencies
    constexpr int CC = kChainCH;
    const size_t smem_c = (size_t)CC * kTileBytes;
    static std::atomic<unsigned long long> done_chain{0};
    COSA_CHECK(opt_in_smem(par_chain_kernel<CC>, (int)smem_c, done_chain));
    const int tiles = a.tiles_x * a.tiles_y;
    const long long ctas = (long long)a.num_iter * a.B * tiles;
    if (ctas <= 0x7fffffffLL) {
      COSA_CUDA(cudaMemsetAsync(tile_flags, 0, (size_t)a.B * tiles * sizeof(int), stream));
      // a tile's consecutive steps must be further apart in the grid than the CTAs resident at once
      // (a closer pair still completes - the later CTA waits for the earlier one - it only idles a slot meanwhile)
      int group = g_chain_group.load(std::memory_order_relaxed);
      if (group <= 0) {
        // automatic: groups of about 8 images (their affinity planes stay close in L2), all of nearly the same size
        // (a short last group would bring a tile's consecutive steps too close), each large enough for the distance
        const int resident = 2 * sm_count();
        int n_groups = max(1, a.B / 8);
        while (n_groups > 1 && (long long)(a.B / n_groups) * tiles < resident + 2 * a.tiles_x + 2) --n_groups;
        group = ceil_div(a.B, n_groups);
      }
      group = min(group, a.B);
      COSA_LAUNCH(par_chain_kernel<CC>, dim3((unsigned)ctas), 256, smem_c, stream, a, t0, ta, tb, tile_flags, group);
      return 0;
    }
  }
  // two CTAs per tile share the channels of an image with more than CH live ones (the kernel keeps <= CH channels in
  // one CTA, where the affinity quads are loaded once)
  const int gsplit = 2;
  const size_t smem = (size_t)CH * kTileBytes;
  static std::atomic<unsigned long long> done_tile{0};
  COSA_CHECK(opt_in_smem(par_iterate_tile_kernel<CH>, (int)smem, done_tile));
  const dim3 grid(a.tiles_x, a.tiles_y, a.B * gsplit);
  for (int it = 0; it < a.num_iter; ++it) {
    const bool last = it == a.num_iter - 1;
    const CUtensorMap &tm = it == 0 ? t0 : (((it - 1) & 1) ? tb : ta);
    float *dst = last ? a.out_final : ((it & 1) ? a.out_b : a.out_a);
    const MaskLayout lo_it = last ? a.lo_final : a.li;
    // steps 2..T: programmatic dependent launch behind the previous step (see the kernel's prologue); plain launches
    // while the per-kernel event timing is on, so that a step's events bracket that step alone
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (it > 0 && !g_prof_on) ? 1 : 0;
    if (g_prof_on) prof_mark("par_iterate_tile_kernel", stream, true);
    const cudaError_t e = cudaLaunchKernelEx(&cfg, par_iterate_tile_kernel<CH>, a.aff, tm, a.li, dst, lo_it, a.nch_dev,
                                             a.nch_uniform, a.c_stride, a.h, a.w, gsplit);
    ++g_launches;
    if (g_prof_on) prof_mark("par_iterate_tile_kernel", stream, false);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

int par_launch_iterations(const ParConst &pc, const float *aff, const float *src0, float *scratch_a, float *scratch_b,
                          MaskLayout lay, float *final_dst, MaskLayout lay_final, const int *nch_dev, int nch_uniform,
                          int c_stride, int B, int h, int w, int num_iter, int *tile_flags, cudaStream_t stream) {
  if (num_iter <= 0) return COSA_E_ARG;   // callers handle the zero-iteration copy themselves
  const int n_dil = pc.n_dil;
  const bool wide = nch_dev ? (c_stride > 4) : (nch_uniform > 4);
  const bool vec = lay.padn > 0;
  const int mode = g_step_mode.load(std::memory_order_relaxed);
  if (par_tile_path(pc, lay, lay_final))   // `aff` is in the tile-permuted layout (par_launch_affinity)
    return par_launch_propagate(aff, src0, scratch_a, scratch_b, lay, final_dst, lay_final, nch_dev, nch_uniform,
                                c_stride, B, h, w, num_iter, mode, tile_flags, stream);
  const float *src = src0;
  for (int it = 0; it < num_iter; ++it) {
    const bool last = it == num_iter - 1;
    float *dst = last ? final_dst : ((it & 1) ? scratch_b : scratch_a);
    const MaskLayout lo = last ? lay_final : lay;
    if (vec) {
      dim3 grid(ceil_div(w, kTileW), ceil_div(h, kTileH), B), block(256);
      static std::atomic<unsigned long long> done8{0}, done4{0};
      if (wide) {
        COSA_CHECK(opt_in_smem(par_iterate_smem_kernel<8>, 8 * kSmemH * kSmemW * (int)sizeof(float), done8));
        COSA_LAUNCH(par_iterate_smem_kernel<8>, grid, block, 8 * kSmemH * kSmemW * sizeof(float), stream, aff, src, lay,
                    dst, lo, nch_dev, nch_uniform, c_stride, h, w, n_dil, pc);
      } else {
        COSA_CHECK(opt_in_smem(par_iterate_smem_kernel<4>, 4 * kSmemH * kSmemW * (int)sizeof(float), done4));
        COSA_LAUNCH(par_iterate_smem_kernel<4>, grid, block, 4 * kSmemH * kSmemW * sizeof(float), stream, aff, src, lay,
                    dst, lo, nch_dev, nch_uniform, c_stride, h, w, n_dil, pc);
      }
    } else {
      if (lo.pitch != w || lo.off != 0) return COSA_E_ARG;   // the generic kernel writes plain NCHW only
      dim3 grid(ceil_div(w, 32), ceil_div(h, 8), B), block(256);
      if (wide) {
        COSA_LAUNCH(par_iterate_kernel<8>, grid, block, 0, stream, aff, src, dst, nch_dev, nch_uniform, c_stride, h,
                    w, n_dil, pc);
      } else {
        COSA_LAUNCH(par_iterate_kernel<4>, grid, block, 0, stream, aff, src, dst, nch_dev, nch_uniform, c_stride, h,
                    w, n_dil, pc);
      }
    }
    src = dst;
  }
  return 0;
}

// Affinity + num_iter propagation steps for the whole batch.  (Splitting the batch into L2-sized chunks so that
// the affinity planes are re-read from L2 was measured and is slower: the step is bound by the L1 path of the
// neighbour loads, not by the affinity stream - see profiles/README.md.)
int par_refine_batch(const ParConst &pc, const float *imgs, float *aff, const float *src0, float *scratch_a,
                     float *scratch_b, MaskLayout lay, float *final_dst, MaskLayout lay_final, const int *nch_dev,
                     int nch_uniform, int c_stride, int B, int h, int w, int num_iter, int *tile_flags,
                     cudaStream_t stream) {
  COSA_CHECK(par_launch_affinity(pc, imgs, aff, B, h, w, par_tile_path(pc, lay, lay_final), stream));
  return par_launch_iterations(pc, aff, src0, scratch_a, scratch_b, lay, final_dst, lay_final, nch_dev, nch_uniform,
                               c_stride, B, h, w, num_iter, tile_flags, stream);
}

}  // namespace cosa

using namespace cosa;

// scratch: affinity [B,8*n_dil,h,w] + two mask buffers in the widest layout cosa_par_forward uses
// (column pads of at most 24: wider dilations take the plain layout and the generic kernel)
extern "C" size_t cosa_par_ws_bytes(int B, int C, int h, int w, int n_dil) {
  const size_t pitch = (size_t)max_padded_pitch(w);
  return align_up(par_affinity_floats(B, n_dil, h, w) * sizeof(float), 256) +
         2 * align_up((size_t)B * C * h * pitch * sizeof(float), 256) + align_up(par_tile_flag_ints(B, h, w) * sizeof(int), 256);
}

extern "C" int cosa_par_set_step_mode(const char *name) {
  const int m = parse_step_mode(name);
  if (m < 0) return COSA_E_ARG;
  if (m == kStepChain) {   // "chain" or "chain<images per group>"
    const char *d = name;
    while (*d && (*d < '0' || *d > '9')) ++d;
    const int grp = *d ? atoi(d) : 0;
    if (grp < 0) return COSA_E_ARG;
    g_chain_group.store(grp, std::memory_order_relaxed);
  }
  g_step_mode.store(m, std::memory_order_relaxed);
  return 0;
}

extern "C" int cosa_par_affinity(const float *imgs, float *aff, int B, int h, int w, const int *dilations, int n_dil,
                                 void *stream) {
  if (!imgs || !aff || B < 1 || h < 1 || w < 1) return COSA_E_ARG;
  ParConst pc;
  COSA_CHECK(par_make_constants(dilations, n_dil, &pc));
  return par_launch_affinity(pc, imgs, aff, B, h, w, false, (cudaStream_t)stream);
}

extern "C" int cosa_par_forward(const float *imgs, const float *masks_in, float *masks_out, int B, int C, int h, int w,
                                int hm, int wm, const int *dilations, int n_dil, int num_iter, void *ws,
                                size_t ws_bytes, void *stream) {
  if (!imgs || !masks_in || !masks_out || !ws || B < 1 || C < 1 || h < 1 || w < 1 || hm < 1 || wm < 1 || num_iter < 0)
    return COSA_E_ARG;
  if (ws_bytes < cosa_par_ws_bytes(B, C, h, w, n_dil)) return COSA_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  ParConst pc;
  COSA_CHECK(par_make_constants(dilations, n_dil, &pc));
  const size_t plane = (size_t)h * w;
  MaskLayout lay = padded_layout(w, dilations, n_dil);
  if (lay.padn > 24) lay = plain_layout(w);   // keeps the scratch within cosa_par_ws_bytes
  const MaskLayout plain = plain_layout(w);
  Arena arena(ws);
  float *aff = arena.take<float>(par_affinity_floats(B, n_dil, h, w));
  float *buf_a = arena.take<float>(layout_floats(lay, B, C, h));
  float *buf_b = arena.take<float>(layout_floats(lay, B, C, h));
  int *tile_flags = arena.take<int>(par_tile_flag_ints(B, h, w));
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
  return par_refine_batch(pc, imgs, aff, src0, buf_a, buf_b, lay, masks_out, plain, nullptr, C, C, B, h, w, num_iter,
                          tile_flags, s);
}
