// Dense-CRF mean-field inference on the GPU lattices.
//
// Reference: utils/seg_helper.py:961-996 (class DenseCRF / crf_inference_infv2, called at evaluation time from
// evaluation_engine.py:205-211) and :905-922 (crf_inference_inf).  Both drive pydensecrf (Kraehenbuehl & Koltun's
// DenseCRF2D; not vendored in the reference tree, un-pinned upstream - DESIGN.md section 2):
//
//     U  = -log(clip(probs, 1e-5, 1))                                   unary_from_softmax
//     Q  = softmax(-U)                                                  DenseCRF::inference: expAndNormalize(Q, -unary)
//     repeat iter_max times:
//         Q = softmax( -U + pos_w * K_g(Q) + bi_w * K_b(Q) )            PottsCompatibility: -w * filtered, tmp1 -= tmp2
//     K(Q) = norm * filter(norm * Q),  norm = 1 / sqrt(filter(1) + 1e-20)   DenseKernel, NORMALIZE_SYMMETRIC (the default
//                                                                        of addPairwiseGaussian / addPairwiseBilateral)
// K_g filters over the 2-D lattice of (x, y) / pos_xy_std, K_b over the 5-D lattice of (x, y) / bi_xy_std and
// (R, G, B) / bi_rgb_std - the same permutohedral splat / blur / slice as the training loss (lattice_kernels.cu), whose
// D = 2 instantiation exists for this function.
#include <math.h>

#include "common.cuh"
#include "lattice.cuh"

namespace cosa {

// One thread per pixel; channels are planes of n pixels.
__global__ void __launch_bounds__(256) crf_unary_kernel(const float *__restrict__ probs, float *__restrict__ unary,
                                                        float *__restrict__ q, float *__restrict__ ones, int C, int n,
                                                        long long P) {
  for (long long gp = blockIdx.x * (long long)blockDim.x + threadIdx.x; gp < P; gp += (long long)gridDim.x * blockDim.x) {
    const long long b = gp / n, p = gp - b * n;
    const float *src = probs + b * C * n + p;
    float *u = unary + b * C * n + p, *dst = q + b * C * n + p;
    float mn = INFINITY;
    for (int c = 0; c < C; ++c) {
      const float e = -logf(fminf(fmaxf(__ldg(src + (size_t)c * n), 1e-5f), 1.0f));   // unary_from_softmax, clip = 1e-5
      u[(size_t)c * n] = e;
      mn = fminf(mn, e);
    }
    float den = 0.0f;
    for (int c = 0; c < C; ++c) {
      const float e = expf(mn - u[(size_t)c * n]);       // expAndNormalize(-U): subtract the column maximum first
      dst[(size_t)c * n] = e;
      den += e;
    }
    for (int c = 0; c < C; ++c) dst[(size_t)c * n] = dst[(size_t)c * n] / den;
    ones[gp] = 1.0f;
  }
}

__global__ void __launch_bounds__(256) crf_norm_kernel(float *__restrict__ norm, long long P) {
  for (long long gp = blockIdx.x * (long long)blockDim.x + threadIdx.x; gp < P; gp += (long long)gridDim.x * blockDim.x)
    norm[gp] = 1.0f / sqrtf(norm[gp] + 1e-20f);
}

// s_g = norm_g * Q, s_b = norm_b * Q
__global__ void __launch_bounds__(256) crf_scale_kernel(const float *__restrict__ q, const float *__restrict__ norm_g,
                                                        const float *__restrict__ norm_b, float *__restrict__ s_g,
                                                        float *__restrict__ s_b, int C, int n, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / ((long long)C * n), p = i % n;
    const float v = q[i];
    s_g[i] = v * norm_g[b * n + p];
    s_b[i] = v * norm_b[b * n + p];
  }
}

// Q = softmax(-U + pos_w * norm_g * F_g + bi_w * norm_b * F_b) over the channels of every pixel
__global__ void __launch_bounds__(256) crf_update_kernel(const float *__restrict__ unary, const float *__restrict__ f_g,
                                                         const float *__restrict__ f_b, const float *__restrict__ norm_g,
                                                         const float *__restrict__ norm_b, float pos_w, float bi_w,
                                                         float *__restrict__ q, int C, int n, long long P) {
  for (long long gp = blockIdx.x * (long long)blockDim.x + threadIdx.x; gp < P; gp += (long long)gridDim.x * blockDim.x) {
    const long long b = gp / n, p = gp - b * n;
    const size_t at = (size_t)b * C * n + p;
    const float wg = pos_w * norm_g[gp], wb = bi_w * norm_b[gp];
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) {
      const size_t i = at + (size_t)c * n;
      const float t = -unary[i] + wg * f_g[i] + wb * f_b[i];
      q[i] = t;
      mx = fmaxf(mx, t);
    }
    float den = 0.0f;
    for (int c = 0; c < C; ++c) {
      const size_t i = at + (size_t)c * n;
      const float e = expf(q[i] - mx);
      q[i] = e;
      den += e;
    }
    for (int c = 0; c < C; ++c) q[at + (size_t)c * n] = q[at + (size_t)c * n] / den;
  }
}

static int crf_grid(long long items) { return (int)max(1LL, min((long long)sm_count() * 8, ceil_div_ll(items, 256))); }

struct CrfCarve {
  float *unary, *s_g, *s_b, *f_g, *f_b, *norm_g, *norm_b, *ones;
  void *lat_g, *lat_b;
  size_t bytes;
};

static CrfCarve crf_carve(void *ws, int N, int C, int H, int W) {
  const size_t n = (size_t)H * W, plane = (size_t)N * C * n;
  Arena a(ws ? ws : (void *)256);   // a null workspace only measures
  CrfCarve c;
  c.unary = a.take<float>(plane);
  c.s_g = a.take<float>(plane);
  c.s_b = a.take<float>(plane);
  c.f_g = a.take<float>(plane);
  c.f_b = a.take<float>(plane);
  c.norm_g = a.take<float>((size_t)N * n);
  c.norm_b = a.take<float>((size_t)N * n);
  c.ones = a.take<float>((size_t)N * n);
  c.lat_g = a.base + a.off;
  a.off += align_up(lattice_ws_bytes(N, C, H, W, 2), 256);
  c.lat_b = a.base + a.off;
  a.off += align_up(lattice_ws_bytes(N, C, H, W, 5), 256);
  c.bytes = a.off;
  return c;
}

}  // namespace cosa

using namespace cosa;

extern "C" size_t cosa_crf_inference_ws_bytes(int N, int C, int H, int W) {
  if (N < 1 || N > kMaxImagesPerLattice || C < 1 || H < 1 || W < 1) return 0;
  return crf_carve(nullptr, N, C, H, W).bytes;
}

extern "C" int cosa_crf_inference(const float *images, const float *probs, float *q_out, int N, int C, int H, int W,
                                  int iter_max, float pos_w, float pos_xy_std, float bi_w, float bi_xy_std,
                                  float bi_rgb_std, void *ws, size_t ws_bytes, void *stream) {
  if (!images || !probs || !q_out || !ws || N < 1 || N > kMaxImagesPerLattice || C < 1 || H < 1 || W < 1 || iter_max < 0)
    return COSA_E_ARG;
  if (!(pos_xy_std > 0.0f) || !(bi_xy_std > 0.0f) || !(bi_rgb_std > 0.0f)) return COSA_E_ARG;
  if (ws_bytes < cosa_crf_inference_ws_bytes(N, C, H, W)) return COSA_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const int n = H * W;
  const long long P = (long long)N * n, total = P * C;
  const CrfCarve c = crf_carve(ws, N, C, H, W);
  COSA_LAUNCH(crf_unary_kernel, crf_grid(P), 256, 0, s, probs, c.unary, q_out, c.ones, C, n, P);
  if (iter_max == 0) return 0;
  // the two lattices: spatial (d = 2) and bilateral (d = 5); each is carved twice over the same memory - with K = C
  // for the message passing and with K = 1 for the normalisation filter (the buffers the build fills do not depend on K)
  LatticeBufs Lg, Lb, Lg1, Lb1;
  lattice_carve(c.lat_g, N, C, H, W, &Lg, 2);
  lattice_carve(c.lat_g, N, 1, H, W, &Lg1, 2);
  lattice_carve(c.lat_b, N, C, H, W, &Lb, 5);
  lattice_carve(c.lat_b, N, 1, H, W, &Lb1, 5);
  COSA_CHECK(lattice_build(Lg, images, N, H, W, 1.0f, pos_xy_std, true, s));
  COSA_CHECK(lattice_build(Lb, images, N, H, W, bi_rgb_std, bi_xy_std, true, s));
  // norm = 1 / sqrt(filter(1) + 1e-20)
  COSA_CHECK(lattice_splat_blur(Lg1, c.ones, N, 1, H, W, s));
  COSA_CHECK(lattice_slice(Lg1, c.ones, nullptr, nullptr, c.norm_g, N, 1, H, W, s));
  COSA_CHECK(lattice_splat_blur(Lb1, c.ones, N, 1, H, W, s));
  COSA_CHECK(lattice_slice(Lb1, c.ones, nullptr, nullptr, c.norm_b, N, 1, H, W, s));
  COSA_LAUNCH(crf_norm_kernel, crf_grid(P), 256, 0, s, c.norm_g, P);
  COSA_LAUNCH(crf_norm_kernel, crf_grid(P), 256, 0, s, c.norm_b, P);
  for (int it = 0; it < iter_max; ++it) {
    COSA_LAUNCH(crf_scale_kernel, crf_grid(total), 256, 0, s, q_out, c.norm_g, c.norm_b, c.s_g, c.s_b, C, n, total);
    COSA_CHECK(lattice_splat_blur(Lg, c.s_g, N, C, H, W, s));
    COSA_CHECK(lattice_slice(Lg, c.s_g, nullptr, nullptr, c.f_g, N, C, H, W, s));
    COSA_CHECK(lattice_splat_blur(Lb, c.s_b, N, C, H, W, s));
    COSA_CHECK(lattice_slice(Lb, c.s_b, nullptr, nullptr, c.f_b, N, C, H, W, s));
    COSA_LAUNCH(crf_update_kernel, crf_grid(P), 256, 0, s, c.unary, c.f_g, c.f_b, c.norm_g, c.norm_b, pos_w, bi_w, q_out,
                C, n, P);
  }
  return 0;
}
