// Shared helpers for the cosa_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cosa_b200.h"

namespace cosa {

extern unsigned long long g_launches;   // kernels launched through the C-ABI (capi.cu)

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Optional per-kernel timing (cosa_profile_begin / cosa_profile_end): CUDA events recorded on the launching
// stream around every launch while profiling is on.  Off by default; no cost beyond one branch.
extern bool g_prof_on;
void prof_mark(const char *name, cudaStream_t stream, bool is_start);

// Launch bookkeeping: every kernel launch in the library goes through COSA_LAUNCH so that
// cosa_launch_count() is an honest count and launch errors surface as return codes.
#define COSA_LAUNCH(kernel, grid, block, smem, stream, ...)                  \
  do {                                                                       \
    if (::cosa::g_prof_on) ::cosa::prof_mark(#kernel, (stream), true);       \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);              \
    ++::cosa::g_launches;                                                    \
    cudaError_t e_ = cudaGetLastError();                                     \
    if (::cosa::g_prof_on) ::cosa::prof_mark(#kernel, (stream), false);      \
    if (e_ != cudaSuccess) return (int)e_;                                   \
  } while (0)

// Same for template instantiations (commas in the kernel expression): the profile name is given explicitly.
#define COSA_LAUNCH_T(name, kernel, grid, block, smem, stream, ...)          \
  do {                                                                       \
    if (::cosa::g_prof_on) ::cosa::prof_mark(name, (stream), true);          \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);              \
    ++::cosa::g_launches;                                                    \
    cudaError_t e_ = cudaGetLastError();                                     \
    if (::cosa::g_prof_on) ::cosa::prof_mark(name, (stream), false);         \
    if (e_ != cudaSuccess) return (int)e_;                                   \
  } while (0)

#define COSA_CHECK(expr)                        \
  do {                                          \
    int r_ = (expr);                            \
    if (r_ != 0) return r_;                     \
  } while (0)

#define COSA_CUDA(expr)                         \
  do {                                          \
    cudaError_t e_ = (expr);                    \
    if (e_ != cudaSuccess) return (int)e_;      \
  } while (0)

__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over the caller's workspace (256-byte aligned slices).
struct Arena {
  char *base;
  size_t off;
  explicit Arena(void *p) : base((char *)p), off(0) {}
  template <typename T>
  T *take(size_t count) {
    T *p = (T *)(base + off);
    off += align_up(count * sizeof(T), 256);
    return p;
  }
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ---- bilinear resampling, arithmetic as torch's CPU kernels evaluate it (DESIGN.md "resampling") ----
struct Tap {
  int i0, i1;
  float w0, w1;
};

// align_corners=False source index (area_pixel_compute_source_index): scale*(dst+0.5)-0.5 clamped at 0.
__device__ __forceinline__ Tap tap_half_pixel(int dst, float scale, int in_size) {
  float real = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  real = fmaxf(real, 0.0f);
  Tap t;
  t.i0 = min((int)real, in_size - 1);
  t.i1 = min(t.i0 + 1, in_size - 1);
  t.w1 = fminf(fmaxf(__fsub_rn(real, (float)t.i0), 0.0f), 1.0f);
  t.w0 = __fsub_rn(1.0f, t.w1);
  return t;
}

// align_corners=True: scale = (in-1)/(out-1), src = scale*dst.
__device__ __forceinline__ Tap tap_align_corners(int dst, float scale, int in_size) {
  float real = __fmul_rn(scale, (float)dst);
  Tap t;
  t.i0 = min((int)real, in_size - 1);
  t.i1 = min(t.i0 + 1, in_size - 1);
  t.w1 = fminf(fmaxf(__fsub_rn(real, (float)t.i0), 0.0f), 1.0f);
  t.w0 = __fsub_rn(1.0f, t.w1);
  return t;
}

// Up-sampling form: x-lerp then y-lerp, each as fma(w0, a, rn(w1*b)).
__device__ __forceinline__ float lerp_nested(float w0, float a, float w1, float b) {
  return __fmaf_rn(w0, a, __fmul_rn(w1, b));
}
__device__ __forceinline__ float bilerp_up(const Tap &ty, const Tap &tx, float a00, float a01, float a10, float a11) {
  float top = lerp_nested(tx.w0, a00, tx.w1, a01);
  float bot = lerp_nested(tx.w0, a10, tx.w1, a11);
  return lerp_nested(ty.w0, top, ty.w1, bot);
}
// Down-sampling form: four products accumulated left to right (exact 2:1 gives 0.25*(((a+b)+c)+d)).
__device__ __forceinline__ float bilerp_down(const Tap &ty, const Tap &tx, float a00, float a01, float a10, float a11) {
  float t = __fmul_rn(__fmul_rn(ty.w0, tx.w0), a00);
  t = __fmaf_rn(__fmul_rn(ty.w0, tx.w1), a01, t);
  t = __fmaf_rn(__fmul_rn(ty.w1, tx.w0), a10, t);
  t = __fmaf_rn(__fmul_rn(ty.w1, tx.w1), a11, t);
  return t;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Blackwell packed fp32 (sm_100a): two IEEE round-to-nearest operations per issue slot (SASS FFMA2 / FADD2 / FMUL2).
// Each lane rounds exactly like the scalar instruction, so results are those of the scalar code.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

// 128-bit streaming loads/stores (read-once / write-once data: keep it out of L1).
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// 128-bit load that bypasses L1 but may hit L2 (second pass over data the same thread streamed before)
__device__ __forceinline__ float4 ldg4c(const float *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void stg_stream4(float *p, const float4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w));
}

}  // namespace cosa
