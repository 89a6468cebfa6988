// Permutohedral-lattice bilateral filter on the GPU (d = 5).
//
// Reference: utils/bilateralfilter/bilateralfilter.cpp:4-55 (features, per-image driver, batch loop) and
// utils/bilateralfilter/permutohedral.cpp:115-297 (Permutohedral::init, SSE branch), :507-571 (compute).
//
//   build      one thread per pixel: features -> elevate -> round (half-even) -> rank -> barycentric ->
//              6 packed vertex keys, inserted into an open-addressing hash table in HBM with 64-bit CAS;
//              lanes of a warp that hold the same key elect one inserter (warp-aggregated atomics).
//   resolve    table slot -> dense vertex id for every (pixel, vertex)
//   neighbours 12 look-ups per vertex -> blur neighbour table                      (permutohedral.cpp:272-297)
//   splat      values[v] += bary * in, all K channels of a vertex in one row, 128-bit vector reductions
//   blur       6 Jacobi passes  new = old + 0.5 * (old[n1] + old[n2])               (permutohedral.cpp:536-552)
//   slice      out = sum_r bary_r * alpha * values[v_r]  (+ fused gate / energy dot)  (permutohedral.cpp:554-567)
//
// Arithmetic that decides DISCRETE outcomes (rounding, ranks, keys) and the barycentric weights use explicit
// round-to-nearest intrinsics in the reference's operation order, so the vertex set and the weights are
// bit-identical to the CPU code; only the summation order of the splat (atomics) differs.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "lattice.cuh"

namespace cosa {

// ---- packed keys ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack_key(const int q[kLatD], int r, int b, int *bad) {
  unsigned long long k = 0;
#pragma unroll
  for (int i = 0; i < kLatD; ++i) {
    const int v = q[i] + kQBias;
    if (v < 0 || v >= (1 << kQBits)) *bad = 1;
    k |= (unsigned long long)(v & ((1 << kQBits) - 1)) << (kQBits * i);
  }
  k |= (unsigned long long)r << (kQBits * kLatD);
  k |= (unsigned long long)b << (kQBits * kLatD + 3);
  return k;
}

__device__ __forceinline__ unsigned long long hash_key(unsigned long long k) {
  // splitmix64 finaliser: full avalanche, so linear probing sees no structure from the lattice geometry
  k ^= k >> 30; k *= 0xbf58476d1ce4e5b9ULL;
  k ^= k >> 27; k *= 0x94d049bb133111ebULL;
  k ^= k >> 31;
  return k;
}

// Scale factors of the elevation (permutohedral.cpp:156-159): double arithmetic with a float inv_std_dev,
// stored as float.  Evaluated on the host, passed by value.
struct EmbedConst {
  float sf[kLatD];
};

static EmbedConst make_embed_const() {
  EmbedConst c;
  const float inv_std_dev = (float)(sqrt(2.0 / 3.0) * (kLatD + 1));
  for (int i = 0; i < kLatD; ++i) c.sf[i] = (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * inv_std_dev);
  return c;
}

// One point of Permutohedral::init (permutohedral.cpp:176-252).  Outputs q0[i] = rem0[i]/6 (after the wrap),
// rank[i] and the six barycentric weights.
__device__ __forceinline__ void embed_point(const float f[kLatD], const EmbedConst &ec, int q0[kLatD + 1],
                                            int rank[kLatD + 1], float bary[kLatD + 1]) {
  constexpr int D = kLatD;
  const float inv6 = 1.0f / 6.0f;
  float el[D + 1], rem0[D + 1];
  float sm = 0.0f;
#pragma unroll
  for (int j = D; j > 0; --j) {
    const float cf = __fmul_rn(f[j - 1], ec.sf[j - 1]);
    el[j] = __fsub_rn(sm, __fmul_rn((float)j, cf));
    sm = __fadd_rn(sm, cf);
  }
  el[0] = sm;
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    const float v = rintf(__fmul_rn(inv6, el[i]));     // _mm_cvtps_epi32: round half to even
    rem0[i] = __fmul_rn(v, 6.0f);
    sum = __fadd_rn(sum, v);
  }
  float diff[D + 1];
#pragma unroll
  for (int i = 0; i <= D; ++i) { diff[i] = __fsub_rn(el[i], rem0[i]); rank[i] = 0; }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i + 1; j <= D; ++j) {
      const int c = diff[i] < diff[j] ? 1 : 0;
      rank[i] += c;
      rank[j] += 1 - c;
    }
  const int isum = (int)sum;
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    rank[i] += isum;
    if (rank[i] < 0) { rank[i] += D + 1; rem0[i] = __fadd_rn(rem0[i], 6.0f); }
    else if (rank[i] > D) { rank[i] -= D + 1; rem0[i] = __fsub_rn(rem0[i], 6.0f); }
  }
  // barycentric (permutohedral.cpp:222-241): b[5-rank] += v, b[6-rank] -= v, then b[0] += 1 + b[6].
  // rank is a permutation, so b[s] = v(rank = 5-s) - v(rank = 6-s) whatever the visiting order.
  float vr[D + 1];
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    const float v = __fmul_rn(__fsub_rn(el[i], rem0[i]), inv6);
#pragma unroll
    for (int k = 0; k <= D; ++k)
      if (rank[i] == k) vr[k] = v;
    q0[i] = (int)rintf(__fmul_rn(rem0[i], inv6));   // exact: rem0 is a small multiple of 6
  }
#pragma unroll
  for (int s = 1; s <= D; ++s) bary[s] = __fsub_rn(vr[D - s], vr[D + 1 - s]);
  bary[0] = __fadd_rn(vr[D], __fadd_rn(1.0f, -vr[0]));
}

// ---- build ---------------------------------------------------------------------------------------
__global__ void lattice_clear_kernel(unsigned long long *table_keys, unsigned long long cap, int *counters,
                                     int first_chunk) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x * 2;
  for (unsigned long long i = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 2; i < cap; i += stride)
    *reinterpret_cast<ulonglong2 *>(table_keys + i) = make_ulonglong2(kEmptyKey, kEmptyKey);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // 0: M of this chunk   1: key-range error (sticky over the chunks of a call)   2: max probe length (same)
    // 3: table capacity     4: vertices of the earlier chunks of this call
    counters[4] = first_chunk ? 0 : counters[4] + counters[0];
    counters[0] = 0;
    if (first_chunk) { counters[1] = 0; counters[2] = 0; }
    counters[3] = (int)min(cap, (unsigned long long)0x7fffffff);
  }
}

__global__ void __launch_bounds__(256) lattice_build_kernel(LatticeBufs L, const float *__restrict__ images,
                                                            EmbedConst ec, int N, int H, int W, int n_pad,
                                                            float sigmargb, float sigmaxy) {
  __shared__ int s_new, s_base;
  if (threadIdx.x == 0) s_new = 0;
  __syncthreads();
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)N * n_pad;
  const bool in_range = g < total;        // out-of-range threads stay for the block collectives
  const int n = H * W;
  const int b = in_range ? (int)(g / n_pad) : 0, p = in_range ? (int)(g % n_pad) : 0;
  const bool real = in_range && p < n;   // the SSE loop also embeds zero-feature padding pixels (permutohedral.cpp:168-173)

  float f[kLatD] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (real) {
    const float *img = images + (size_t)b * 3 * n;
    f[0] = __fdiv_rn((float)(p % W), sigmaxy);               // bilateralfilter.cpp:9-13
    f[1] = __fdiv_rn((float)(p / W), sigmaxy);
    f[2] = __fdiv_rn(__ldg(img + p), sigmargb);
    f[3] = __fdiv_rn(__ldg(img + n + p), sigmargb);
    f[4] = __fdiv_rn(__ldg(img + 2 * n + p), sigmargb);
  }
  int q0[kLatD + 1], rank[kLatD + 1];
  float bary[kLatD + 1];
  embed_point(f, ec, q0, rank, bary);

  const long long gp = (long long)b * n + p;   // compact pixel index
  int bad = 0, max_probe = 0;
  unsigned new_mask = 0;                        // bit r: this thread created the table entry of vertex r
  unsigned long long key[kLatD + 1], slot[kLatD + 1], cur[kLatD + 1];
  // Nine out of ten look-ups find a vertex that an earlier pixel created (M is ~0.5 n for 6 n look-ups), so the
  // table is probed with plain L2 loads - all six in flight at once - and only an empty slot costs a CAS.  A stale
  // "empty" is harmless (the CAS decides); an occupied slot never changes again.
#pragma unroll
  for (int r = 0; r <= kLatD; ++r) {
    int q[kLatD];
#pragma unroll
    for (int i = 0; i < kLatD; ++i) q[i] = q0[i] - (rank[i] > kLatD - r ? 1 : 0);   // canonical[r][rank] = r or r-6
    key[r] = pack_key(q, r, b, &bad);
    slot[r] = hash_key(key[r]) & L.cap_mask;
    cur[r] = in_range ? __ldcg(L.table_keys + slot[r]) : key[r];
  }
#pragma unroll
  for (int r = 0; r <= kLatD; ++r) {
    if (in_range) {
      unsigned long long s = slot[r], c = cur[r];
      int probes = 0;
      for (;;) {
        if (c == key[r]) break;
        if (c == kEmptyKey) {
          c = atomicCAS(L.table_keys + s, kEmptyKey, key[r]);
          if (c == kEmptyKey) { new_mask |= 1u << r; break; }
          if (c == key[r]) break;
        }
        s = (s + 1) & L.cap_mask;
        c = __ldcg(L.table_keys + s);
        ++probes;
      }
      slot[r] = s;
      max_probe = max(max_probe, probes);
      if (real) {
        L.offsets[(size_t)r * L.P + gp] = (int)s;
        L.bary[(size_t)r * L.P + gp] = bary[r];
      }
    }
  }
  // dense vertex ids: one global atomic per CTA instead of one per new vertex (a single hot address otherwise)
  const int mine = __popc(new_mask);
  int local = 0;
  if (mine) local = atomicAdd(&s_new, mine);
  __syncthreads();
  if (threadIdx.x == 0) s_base = s_new ? atomicAdd(L.counters, s_new) : 0;
  __syncthreads();
  if (mine) {
    int id = s_base + local;
#pragma unroll
    for (int r = 0; r <= kLatD; ++r) {
      if (new_mask & (1u << r)) {
        L.vkeys[id] = key[r];
        L.table_ids[slot[r]] = id + 1;
        ++id;
      }
    }
  }
  if (bad) atomicExch(L.counters + 1, 1);
  if (max_probe > 0) atomicMax(L.counters + 2, max_probe);
}

__global__ void lattice_resolve_kernel(LatticeBufs L) {
  const long long total = 6 * L.P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    L.offsets[i] = L.table_ids[L.offsets[i]];
}

// After the build the vertex count M is known on the device.  The neighbour look-ups (12 per vertex) would be
// DRAM-latency bound in the worst-case-sized build table, so the M keys are re-inserted into a compact table of
// 2^k >= 2M slots that reuses the head of the build table's storage and stays L2-resident.
__device__ __forceinline__ unsigned long long compact_mask(const LatticeBufs &L) {
  const unsigned m2 = 2u * (unsigned)max(L.counters[0], 512);
  const unsigned long long cap = 1ULL << (32 - __clz(m2 - 1));
  return min(cap, L.cap_mask + 1) - 1;
}

__global__ void lattice_compact_clear_kernel(LatticeBufs L) {
  const unsigned long long cap = compact_mask(L) + 1;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < cap;
       i += (unsigned long long)gridDim.x * blockDim.x)
    L.table_keys[i] = kEmptyKey;
}

__global__ void lattice_compact_insert_kernel(LatticeBufs L) {
  const unsigned long long mask = compact_mask(L);
  const long long M = L.counters[0];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long key = L.vkeys[i];
    unsigned long long slot = hash_key(key) & mask;
    while (atomicCAS(L.table_keys + slot, kEmptyKey, key) != kEmptyKey) slot = (slot + 1) & mask;   // keys are distinct
    L.table_ids[slot] = (int)i + 1;
    // every neighbour entry starts as "absent"; lattice_neighbours_kernel fills in the ones that exist
#pragma unroll
    for (int j = 0; j <= kLatD; ++j) L.nbr[(size_t)j * L.m_cap + i] = make_int2(0, 0);
  }
}

__device__ __forceinline__ int table_find(const LatticeBufs &L, unsigned long long mask, unsigned long long key) {
  unsigned long long slot = hash_key(key) & mask;
  for (;;) {
    const unsigned long long cur = L.table_keys[slot];
    if (cur == key) return L.table_ids[slot];
    if (cur == kEmptyKey) return 0;
    slot = (slot + 1) & mask;
  }
}

// Blur neighbours (permutohedral.cpp:282-294): n1 = key - 1 on every stored coordinate and key[j] + 5 on axis j,
// n2 the opposite.  In (q, r) form the residue moves to r-1 / r+1 and q carries when it wraps.
// The relation is symmetric - u = n1_j(v) exactly when v = n2_j(u) - so a thread looks up only n1 (6 instead of 12
// table walks per vertex) and, when it finds u, also records itself as u's n2.  Entries start as "absent" (0).
__global__ void __launch_bounds__(256) lattice_neighbours_kernel(LatticeBufs L) {
  const long long M = L.counters[0];
  const long long total = M * (kLatD + 1);
  const unsigned long long qmask = (1ULL << kQBits) - 1;
  const unsigned long long tmask = compact_mask(L);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / (kLatD + 1);
    const int j = (int)(idx % (kLatD + 1));
    const unsigned long long key = L.vkeys[i];
    const int r = (int)((key >> (kQBits * kLatD)) & 7);
    const unsigned long long bbits = key >> (kQBits * kLatD + 3);
    // n1: coordinates -1, axis +5
    const int r2 = r == 0 ? kLatD : r - 1;
    const int dq = r == 0 ? -1 : 0;
    int qq[kLatD];
    int bad = 0;
#pragma unroll
    for (int k = 0; k < kLatD; ++k)
      qq[k] = (int)((key >> (kQBits * k)) & qmask) - kQBias + dq + (k == j ? 1 : 0);
    const unsigned long long nk = pack_key(qq, r2, 0, &bad) | (bbits << (kQBits * kLatD + 3));
    const int u = bad ? 0 : table_find(L, tmask, nk);
    if (u > 0) {
      int *base = reinterpret_cast<int *>(L.nbr + (size_t)j * L.m_cap);
      base[2 * i] = u;                       // n1 of vertex i
      base[2 * (size_t)(u - 1) + 1] = (int)i + 1;   // vertex i is the n2 of vertex u - 1
    }
  }
}

// ---- compute -------------------------------------------------------------------------------------
__global__ void lattice_zero_values_kernel(LatticeBufs L) {
  const long long rows = (long long)L.counters[0] + 1;
  const long long total4 = rows * (L.Kp / 4);
  float4 *v0 = reinterpret_cast<float4 *>(L.val0);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x)
    v0[i] = z;
  if (blockIdx.x == 0 && threadIdx.x < L.Kp / 4) reinterpret_cast<float4 *>(L.val1)[threadIdx.x] = z;
}

__device__ __forceinline__ void red_add_v4(float *addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// Splat (permutohedral.cpp:526-534).  One thread per pixel; the channels of a vertex are contiguous so each
// (pixel, vertex, 4 channels) is one 128-bit vector reduction in L2.
__global__ void __launch_bounds__(256) lattice_splat_kernel(LatticeBufs L, const float *__restrict__ ins, int K,
                                                            int n) {
  for (long long gp = blockIdx.x * (long long)blockDim.x + threadIdx.x; gp < L.P;
       gp += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(gp / n), p = (int)(gp % n);
    int off[kLatD + 1];
    float w[kLatD + 1];
#pragma unroll
    for (int r = 0; r <= kLatD; ++r) {
      off[r] = L.offsets[(size_t)r * L.P + gp];
      w[r] = L.bary[(size_t)r * L.P + gp];
    }
    const float *src = ins + (size_t)b * K * n + p;
    for (int c0 = 0; c0 < L.Kp; c0 += 4) {
      float4 x;
      x.x = c0 + 0 < K ? __ldg(src + (size_t)(c0 + 0) * n) : 0.f;
      x.y = c0 + 1 < K ? __ldg(src + (size_t)(c0 + 1) * n) : 0.f;
      x.z = c0 + 2 < K ? __ldg(src + (size_t)(c0 + 2) * n) : 0.f;
      x.w = c0 + 3 < K ? __ldg(src + (size_t)(c0 + 3) * n) : 0.f;
#pragma unroll
      for (int r = 0; r <= kLatD; ++r)
        red_add_v4(L.val0 + (size_t)off[r] * L.Kp + c0,
                   make_float4(__fmul_rn(w[r], x.x), __fmul_rn(w[r], x.y), __fmul_rn(w[r], x.z), __fmul_rn(w[r], x.w)));
    }
  }
}

// One blur axis (permutohedral.cpp:538-551): thread per (vertex, 4 channels).
__global__ void __launch_bounds__(256) lattice_blur_kernel(LatticeBufs L, const float *__restrict__ src,
                                                           float *__restrict__ dst, int axis) {
  const int kq = L.Kp / 4;
  const long long total = (long long)L.counters[0] * kq;
  const int2 *nb = L.nbr + (size_t)axis * L.m_cap;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / kq;
    const int c0 = (int)(idx % kq) * 4;
    const int2 nn = nb[i];
    const float4 o = *reinterpret_cast<const float4 *>(src + (size_t)(i + 1) * L.Kp + c0);
    const float4 a = *reinterpret_cast<const float4 *>(src + (size_t)nn.x * L.Kp + c0);
    const float4 c = *reinterpret_cast<const float4 *>(src + (size_t)nn.y * L.Kp + c0);
    float4 rr;
    rr.x = __fadd_rn(o.x, __fmul_rn(0.5f, __fadd_rn(a.x, c.x)));
    rr.y = __fadd_rn(o.y, __fmul_rn(0.5f, __fadd_rn(a.y, c.y)));
    rr.z = __fadd_rn(o.z, __fmul_rn(0.5f, __fadd_rn(a.z, c.z)));
    rr.w = __fadd_rn(o.w, __fmul_rn(0.5f, __fadd_rn(a.w, c.w)));
    *reinterpret_cast<float4 *>(dst + (size_t)(i + 1) * L.Kp + c0) = rr;
  }
}

// Slice (permutohedral.cpp:554-567) with the optional dense-CRF epilogue (seg_helper.py:888-890).
template <bool ENERGY>
__global__ void __launch_bounds__(256) lattice_slice_kernel(LatticeBufs L, const float *__restrict__ values,
                                                            const float *__restrict__ ins,
                                                            const float *__restrict__ gate, double *loss_acc,
                                                            float *__restrict__ outs, int K, int n) {
  const float alpha = 1.0f / (1.0f + 0.03125f);   // 1 / (1 + 2^-d)
  float local = 0.0f;
  for (long long gp = blockIdx.x * (long long)blockDim.x + threadIdx.x; gp < L.P;
       gp += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(gp / n), p = (int)(gp % n);
    int off[kLatD + 1];
    float w[kLatD + 1];
#pragma unroll
    for (int r = 0; r <= kLatD; ++r) {
      off[r] = L.offsets[(size_t)r * L.P + gp];
      w[r] = __fmul_rn(L.bary[(size_t)r * L.P + gp], alpha);
    }
    const float gt = ENERGY ? __ldg(gate + gp) : 1.0f;
    float *dst = outs + (size_t)b * K * n + p;
    const float *src = ins + (size_t)b * K * n + p;
    // two adjacent quads (one 32-byte sector of each vertex row; rows are 32-byte aligned, Kp % 4 == 0) per pass, so
    // that a sector is requested by one pair of back-to-back loads instead of twice, five other rows apart
    for (int c0 = 0; c0 < L.Kp; c0 += 8) {
      const bool two = c0 + 4 < L.Kp;
      float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
#pragma unroll
      for (int r = 0; r <= kLatD; ++r) {
        const float *row = values + (size_t)off[r] * L.Kp + c0;
        const float4 v0 = *reinterpret_cast<const float4 *>(row);
        const float4 v1 = two ? *reinterpret_cast<const float4 *>(row + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc0.x = __fadd_rn(acc0.x, __fmul_rn(w[r], v0.x)); acc0.y = __fadd_rn(acc0.y, __fmul_rn(w[r], v0.y));
        acc0.z = __fadd_rn(acc0.z, __fmul_rn(w[r], v0.z)); acc0.w = __fadd_rn(acc0.w, __fmul_rn(w[r], v0.w));
        acc1.x = __fadd_rn(acc1.x, __fmul_rn(w[r], v1.x)); acc1.y = __fadd_rn(acc1.y, __fmul_rn(w[r], v1.y));
        acc1.z = __fadd_rn(acc1.z, __fmul_rn(w[r], v1.z)); acc1.w = __fadd_rn(acc1.w, __fmul_rn(w[r], v1.w));
      }
      const float a8[8] = {acc0.x, acc0.y, acc0.z, acc0.w, acc1.x, acc1.y, acc1.z, acc1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (c0 + k < K) {
          float o = a8[k];
          if (ENERGY) {
            o = __fmul_rn(o, gt);
            local = fmaf(__ldg(src + (size_t)(c0 + k) * n), o, local);
          }
          dst[(size_t)(c0 + k) * n] = o;
        }
      }
    }
  }
  if (ENERGY) {
    __shared__ float s_part[8];
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0.0f;
      t = warp_sum(t);
      if (threadIdx.x == 0) atomicAdd(loss_acc, (double)t);
    }
  }
}

// ---- tile-local vertex sharing ---------------------------------------------------------------------
// A 32 x 8 pixel tile touches 6 * 256 (pixel, vertex) pairs but only ~22 % as many distinct vertices (the lattice
// cells are sigma_xy = 50 pixels wide; measured on the synthetic VOC images: 0.22 at 32 x 8, 0.15 at 32 x 32).  The
// tile kernels below therefore dedup the vertex ids of a tile in a shared-memory hash table first and touch every
// distinct vertex row in L2 once per tile instead of once per pair:
//   splat  pairs are bucketed per vertex (count -> scan -> fill), each vertex sums its contributions from a
//          shared-memory copy of the tile's input channels and issues ONE sector-wide reduction per 8 channels
//   slice  the distinct vertex rows are staged in shared memory once and the pixels gather from there
constexpr int kTileW = 32;
constexpr int kChunk = 24;                       // channels per pass (K = 21 -> one pass)

__device__ __forceinline__ unsigned tile_hash(int id, int bits) { return ((unsigned)id * 2654435761u) >> (32 - bits); }

// Inserts `id` into the open-addressing table hkey[1 << bits] (-1 = empty); returns the slot.  *won is set for the
// thread whose CAS created the entry.
__device__ __forceinline__ int tile_insert(int *hkey, int bits, int id, bool *won) {
  const unsigned mask = (1u << bits) - 1;
  unsigned h = tile_hash(id, bits);
  for (;;) {
    const int old = atomicCAS(hkey + h, -1, id);
    if (old == -1) { *won = true; return (int)h; }
    if (old == id) { *won = false; return (int)h; }
    h = (h + 1) & mask;
  }
}

// Splat (permutohedral.cpp:526-534) with tile-local pre-reduction.  CTA = 256 threads = TH/8 pixels per thread of a
// 32 x TH tile.  Shared memory: hash keys + (start | count) words, the pair list bucketed by vertex, the list of
// occupied slots and the tile's input channels [kChunk][pixels + 1].
template <int TH>
__global__ void __launch_bounds__(256) lattice_splat_tile_kernel(LatticeBufs L, const float *__restrict__ ins, int K,
                                                                 int H, int W) {
  constexpr int PPT = TH / 8, PIX = kTileW * TH, PAIRS = 6 * PIX;
  constexpr int BITS = PPT == 1 ? 11 : 12, HS = 1 << BITS;   // load factor <= 0.75 when every pair is distinct
  constexpr int PLANE = PIX + 1;                               // odd plane pitch: channel-strided reads hit distinct banks
  extern __shared__ __align__(16) unsigned char s_raw[];
  int *hkey = reinterpret_cast<int *>(s_raw);                  // [HS]
  int *hinfo = hkey + HS;                                      // [HS] count, then (start << 16) | count
  int2 *plist = reinterpret_cast<int2 *>(hinfo + HS);          // [PAIRS] (pixel, weight bits) bucketed by vertex
  unsigned short *ulist = reinterpret_cast<unsigned short *>(plist + PAIRS);   // [PAIRS] occupied slots
  float *in_s = reinterpret_cast<float *>(ulist + PAIRS);      // [kChunk][PLANE]
  __shared__ int s_scan[8];
  __shared__ int s_U;

  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * TH, b = blockIdx.z;
  const int n = H * W;
  const int tid = threadIdx.x, lx = tid & 31, ly0 = tid >> 5;
  for (int i = tid; i < HS; i += 256) { hkey[i] = -1; hinfo[i] = 0; }
  if (tid == 0) s_U = 0;
  __syncthreads();

  int hslot[PPT][6], hpos[PPT][6];
  float wgt[PPT][6];
  // every vertex id and weight of the thread's pixels is requested before the first shared-memory atomic (the
  // compiler does not move loads across atomics)
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int x = x0 + lx, y = y0 + ly0 + 8 * j;
    const bool ok = x < W && y < H;
    const long long gp = (long long)b * n + (long long)y * W + x;
#pragma unroll
    for (int r = 0; r <= kLatD; ++r) {
      hslot[j][r] = ok ? L.offsets[(size_t)r * L.P + gp] : -1;
      wgt[j][r] = ok ? L.bary[(size_t)r * L.P + gp] : 0.0f;
    }
  }
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
#pragma unroll
    for (int r = 0; r <= kLatD; ++r) {
      if (hslot[j][r] >= 0) {
        bool won;
        const int h = tile_insert(hkey, BITS, hslot[j][r], &won);
        hslot[j][r] = h;
        hpos[j][r] = atomicAdd(hinfo + h, 1);
      }
    }
  }
  __syncthreads();
  // exclusive scan of the counts (HS / 256 consecutive slots per thread) + list of the occupied slots
  {
    constexpr int PER = HS / 256;
    int cnt[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) { cnt[i] = hinfo[tid * PER + i]; sum += cnt[i]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) s_scan[tid >> 5] = incl;
    __syncthreads();
    int base = incl - sum;
    for (int wv = 0; wv < (tid >> 5); ++wv) base += s_scan[wv];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      if (cnt[i]) {
        hinfo[tid * PER + i] = (base << 16) | cnt[i];
        ulist[atomicAdd(&s_U, 1)] = (unsigned short)(tid * PER + i);
        base += cnt[i];
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < PPT; ++j)
#pragma unroll
    for (int r = 0; r <= kLatD; ++r)
      if (hslot[j][r] >= 0)
        plist[(hinfo[hslot[j][r]] >> 16) + hpos[j][r]] = make_int2((ly0 + 8 * j) * kTileW + lx, __float_as_int(wgt[j][r]));
  const int U = s_U;

  const int sub = (tid & 31) >> 3, cl = tid & 7, wv = tid >> 5;
  for (int c0 = 0; c0 < L.Kp; c0 += kChunk) {
    const int kc = min(kChunk, L.Kp - c0);
    __syncthreads();   // the pair list is complete / the previous pass has read in_s
    for (int i0 = tid; i0 < kc * PIX; i0 += 4 * 256) {   // four loads in flight per thread
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * 256;
        const int c = i / PIX, pl = i - c * PIX;
        const int x = x0 + (pl & 31), y = y0 + (pl >> 5);
        v[k] = 0.0f;
        if (i < kc * PIX && c0 + c < K && x < W && y < H) v[k] = __ldg(ins + ((size_t)b * K + c0 + c) * n + (size_t)y * W + x);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * 256;
        const int c = i / PIX, pl = i - c * PIX;
        if (i < kc * PIX) in_s[c * PLANE + pl] = v[k];
      }
    }
    __syncthreads();
    // a quarter-warp per vertex: lane cl sums channels cl, cl + 8, cl + 16 over the vertex's pairs
    for (int u = wv * 4 + sub; u < U; u += 32) {
      const int slot = ulist[u];
      const int info = hinfo[slot];
      const int start = info >> 16, cnt = info & 0xffff;
      float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
      for (int e = 0; e < cnt; ++e) {
        const int2 pw = plist[start + e];
        const float wv_ = __int_as_float(pw.y);
        const float *src = in_s + cl * PLANE + pw.x;
        a0 = fmaf(wv_, src[0], a0);
        a1 = fmaf(wv_, src[8 * PLANE], a1);
        a2 = fmaf(wv_, src[16 * PLANE], a2);
      }
      float *dst = L.val0 + (size_t)hkey[slot] * L.Kp + c0 + cl;
      if (cl < kc) atomicAdd(dst, a0);
      if (cl + 8 < kc) atomicAdd(dst + 8, a1);
      if (cl + 16 < kc) atomicAdd(dst + 16, a2);
    }
  }
}

// Slice (permutohedral.cpp:554-567) + optional dense-CRF epilogue with the distinct vertex rows of a 32 x 8 tile staged
// in shared memory (up to kSliceRows of them; pairs beyond that read L2 directly).
constexpr int kSliceRows = 512;
template <bool ENERGY>
__global__ void __launch_bounds__(256, 3) lattice_slice_tile_kernel(LatticeBufs L, const float *__restrict__ values,
                                                                 const float *__restrict__ ins,
                                                                 const float *__restrict__ gate, double *loss_acc,
                                                                 float *__restrict__ outs, int K, int H, int W) {
  constexpr int BITS = 11, HS = 1 << BITS, TH = 8;
  __shared__ int hkey[HS];
  __shared__ int hval[HS];                      // local row index of the slot's vertex
  __shared__ int urow_id[kSliceRows];           // vertex id of local row u
  extern __shared__ __align__(16) unsigned char s_raw[];
  float *rows = reinterpret_cast<float *>(s_raw);   // [kSliceRows][kChunk]
  __shared__ int s_U;
  __shared__ float s_part[8];
  const float alpha = 1.0f / (1.0f + 0.03125f);   // 1 / (1 + 2^-d)
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * TH, b = blockIdx.z;
  const int n = H * W;
  const int tid = threadIdx.x, lx = tid & 31, ly = tid >> 5;
  for (int i = tid; i < HS; i += 256) hkey[i] = -1;
  if (tid == 0) s_U = 0;
  __syncthreads();
  const int x = x0 + lx, y = y0 + ly;
  const bool ok = x < W && y < H;
  const long long gp = (long long)b * n + (long long)y * W + x;
  int id[kLatD + 1], hs[kLatD + 1];
  float w[kLatD + 1];
#pragma unroll
  for (int r = 0; r <= kLatD; ++r) {   // all loads first: the compiler does not move loads across the atomics below
    id[r] = ok ? L.offsets[(size_t)r * L.P + gp] : 0;
    w[r] = ok ? __fmul_rn(L.bary[(size_t)r * L.P + gp], alpha) : 0.0f;
    hs[r] = 0;
  }
  const float gt = (ENERGY && ok) ? __ldg(gate + gp) : 1.0f;
  float s_in[kChunk];                   // energy epilogue: the pixel's own inputs of the first channel pass
  if (ENERGY) {
#pragma unroll
    for (int c = 0; c < kChunk; ++c)
      s_in[c] = (ok && c < K) ? __ldg(ins + ((size_t)b * K + c) * n + (size_t)y * W + x) : 0.0f;
  }
#pragma unroll
  for (int r = 0; r <= kLatD; ++r) {
    if (ok) {
      bool won;
      hs[r] = tile_insert(hkey, BITS, id[r], &won);
      if (won) {
        const int u = atomicAdd(&s_U, 1);
        hval[hs[r]] = u;
        if (u < kSliceRows) urow_id[u] = id[r];
      }
    }
  }
  __syncthreads();
  int u_r[kLatD + 1];
#pragma unroll
  for (int r = 0; r <= kLatD; ++r) u_r[r] = ok ? hval[hs[r]] : 0;
  const int U = min(s_U, kSliceRows);
  float local = 0.0f;
  for (int c0 = 0; c0 < L.Kp; c0 += kChunk) {
    const int kc = min(kChunk, L.Kp - c0), kq = kc >> 2;
    __syncthreads();
    for (int i0 = tid; i0 < U * kq; i0 += 4 * 256) {   // four row quads in flight per thread
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * 256;
        const int u = i / kq, q = i - u * kq;
        if (i < U * kq) v[k] = *reinterpret_cast<const float4 *>(values + (size_t)urow_id[u] * L.Kp + c0 + 4 * q);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * 256;
        const int u = i / kq, q = i - u * kq;
        if (i < U * kq) reinterpret_cast<float4 *>(rows)[u * (kChunk / 4) + q] = v[k];
      }
    }
    __syncthreads();
    if (ok) {
#pragma unroll
      for (int q = 0; q < kChunk / 4; ++q) {
        if (q >= kq) break;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r <= kLatD; ++r) {
          const float4 v = u_r[r] < kSliceRows
                               ? reinterpret_cast<const float4 *>(rows)[u_r[r] * (kChunk / 4) + q]
                               : *reinterpret_cast<const float4 *>(values + (size_t)id[r] * L.Kp + c0 + 4 * q);
          acc.x = __fadd_rn(acc.x, __fmul_rn(w[r], v.x));
          acc.y = __fadd_rn(acc.y, __fmul_rn(w[r], v.y));
          acc.z = __fadd_rn(acc.z, __fmul_rn(w[r], v.z));
          acc.w = __fadd_rn(acc.w, __fmul_rn(w[r], v.w));
        }
        const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = c0 + 4 * q + k;
          if (c < K) {
            float o = a4[k];
            const size_t at = ((size_t)b * K + c) * n + (size_t)y * W + x;
            if (ENERGY) {
              o = __fmul_rn(o, gt);
              local = fmaf(c0 == 0 ? s_in[4 * q + k] : __ldg(ins + at), o, local);
            }
            outs[at] = o;
          }
        }
      }
    }
  }
  if (ENERGY) {
    local = warp_sum(local);
    if ((tid & 31) == 0) s_part[tid >> 5] = local;
    __syncthreads();
    if (tid < 32) {
      float t = tid < 8 ? s_part[tid] : 0.0f;
      t = warp_sum(t);
      if (tid == 0) atomicAdd(loss_acc, (double)t);
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------
static unsigned long long table_capacity(long long m_cap) {
  unsigned long long cap = 1024;
  // m_cap is the worst case (six new vertices per pixel; natural images stay below 2.5 n), so the load factor is
  // at most 0.8 and in practice below 0.1
  while (cap < (unsigned long long)m_cap + (unsigned long long)m_cap / 4) cap <<= 1;
  return cap;
}

// Images per lattice: up to kMaxImagesPerLattice (the key's image field).  Smaller chunks whose two value buffers fit
// in L2 were measured and are slower at every size (VOC B = 32: 2.73 ms per step at 64, 2.80 at 16, 2.95 at 8 - the
// blur passes are not HBM-bound and the smaller launches fill the GPU worse).  COSA_LATTICE_CHUNK overrides (A/B).
int lattice_chunk_images(int N, int K, int H, int W) {
  (void)K; (void)H; (void)W;
  static int forced = -1;
  if (forced < 0) {
    const char *e = getenv("COSA_LATTICE_CHUNK");
    forced = e ? atoi(e) : 0;
  }
  int c = forced > 0 ? min(forced, kMaxImagesPerLattice) : kMaxImagesPerLattice;
  c = min(c, N);
  const int chunks = (N + c - 1) / c;
  return (N + chunks - 1) / chunks;   // even chunks
}

size_t lattice_ws_bytes(int N, int K, int H, int W) {
  const long long n = (long long)H * W, n_pad = (n + 3) & ~3LL;
  const long long P = (long long)N * n, m_cap = 6LL * N * n_pad;
  const unsigned long long cap = table_capacity(m_cap);
  const int Kp = (K + 3) & ~3;
  size_t b = 0;
  b += align_up(cap * sizeof(unsigned long long), 256);
  b += align_up(cap * sizeof(int), 256);
  b += align_up((size_t)m_cap * sizeof(unsigned long long), 256);
  b += align_up(8 * sizeof(int), 256);
  b += 2 * align_up((size_t)6 * P * sizeof(float), 256);
  b += align_up((size_t)6 * m_cap * sizeof(int2), 256);
  b += 2 * align_up((size_t)(m_cap + 1) * Kp * sizeof(float), 256);
  return b;
}

void lattice_carve(void *ws, int N, int K, int H, int W, LatticeBufs *L) {
  const long long n = (long long)H * W, n_pad = (n + 3) & ~3LL;
  L->P = (long long)N * n;
  L->m_cap = 6LL * N * n_pad;
  const unsigned long long cap = table_capacity(L->m_cap);
  L->cap_mask = cap - 1;
  L->Kp = (K + 3) & ~3;
  Arena a(ws);
  L->counters = a.take<int>(8);   // first: the same place for every chunk size (cosa_bilateral_stats)
  L->table_keys = a.take<unsigned long long>(cap);
  L->table_ids = a.take<int>(cap);
  L->vkeys = a.take<unsigned long long>((size_t)L->m_cap);
  L->offsets = a.take<int>((size_t)6 * L->P);
  L->bary = a.take<float>((size_t)6 * L->P);
  L->nbr = a.take<int2>((size_t)6 * L->m_cap);
  L->val0 = a.take<float>((size_t)(L->m_cap + 1) * L->Kp);
  L->val1 = a.take<float>((size_t)(L->m_cap + 1) * L->Kp);
}

static int persistent_blocks(long long work_items, int per_block) {
  return (int)max(1LL, min((long long)sm_count() * 8, ceil_div_ll(work_items, per_block)));
}

int lattice_build(const LatticeBufs &L, const float *images, int N, int H, int W, float sigmargb, float sigmaxy,
                  bool first_chunk, cudaStream_t stream) {
  if (N < 1 || N > kMaxImagesPerLattice) return COSA_E_ARG;
  const long long n = (long long)H * W;
  const int n_pad = (int)((n + 3) & ~3LL);
  const unsigned long long cap = L.cap_mask + 1;
  COSA_LAUNCH(lattice_clear_kernel, persistent_blocks((long long)(cap / 2), 256), 256, 0, stream, L.table_keys, cap,
              L.counters, first_chunk ? 1 : 0);
  const long long total = (long long)N * n_pad;
  COSA_LAUNCH(lattice_build_kernel, (unsigned)ceil_div_ll(total, 256), 256, 0, stream, L, images, make_embed_const(),
              N, H, W, n_pad, sigmargb, sigmaxy);
  COSA_LAUNCH(lattice_resolve_kernel, persistent_blocks(6 * L.P, 256), 256, 0, stream, L);
  // the vertex count lives on the device: size the grid for the SMs and let the kernel read it
  COSA_LAUNCH(lattice_compact_clear_kernel, sm_count() * 8, 256, 0, stream, L);
  COSA_LAUNCH(lattice_compact_insert_kernel, sm_count() * 8, 256, 0, stream, L);
  COSA_LAUNCH(lattice_neighbours_kernel, sm_count() * 8, 256, 0, stream, L);
  return 0;
}

// COSA_LATTICE_PIXEL=1 selects the one-thread-per-pixel splat / slice kernels (A/B runs); COSA_SPLAT_TH=16 the
// 32 x 16 splat tile.
static int lattice_env(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}
static bool lattice_per_pixel() {
  static int v = -1;
  if (v < 0) v = lattice_env("COSA_LATTICE_PIXEL", 0) ? 1 : 0;
  return v == 1;
}

template <int TH>
static int launch_splat_tile(const LatticeBufs &L, const float *ins, int N, int K, int H, int W, cudaStream_t stream) {
  constexpr int PIX = kTileW * TH, PAIRS = 6 * PIX, HS = TH == 8 ? 2048 : 4096;
  const size_t smem = (size_t)HS * 8 + (size_t)PAIRS * 8 + (size_t)PAIRS * 2 + (size_t)kChunk * (PIX + 1) * 4;
  static bool attr = false;
  if (!attr) {
    COSA_CUDA(cudaFuncSetAttribute(lattice_splat_tile_kernel<TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const dim3 grid(ceil_div(W, kTileW), ceil_div(H, TH), N);
  COSA_LAUNCH_T("lattice_splat_tile_kernel", lattice_splat_tile_kernel<TH>, grid, 256, smem, stream, L, ins, K, H, W);
  return 0;
}

int lattice_splat_blur(const LatticeBufs &L, const float *ins, int N, int K, int H, int W, cudaStream_t stream) {
  const int n = H * W;
  COSA_LAUNCH(lattice_zero_values_kernel, sm_count() * 8, 256, 0, stream, L);
  if (lattice_per_pixel()) {
    COSA_LAUNCH(lattice_splat_kernel, persistent_blocks(L.P, 256), 256, 0, stream, L, ins, K, n);
  } else {
    static int th = 0;
    if (!th) th = lattice_env("COSA_SPLAT_TH", 8) == 16 ? 16 : 8;
    if (th == 16) COSA_CHECK(launch_splat_tile<16>(L, ins, N, K, H, W, stream));
    else COSA_CHECK(launch_splat_tile<8>(L, ins, N, K, H, W, stream));
  }
  float *src = L.val0, *dst = L.val1;
  for (int axis = 0; axis <= kLatD; ++axis) {
    COSA_LAUNCH(lattice_blur_kernel, sm_count() * 8, 256, 0, stream, L, src, dst, axis);
    float *t = src; src = dst; dst = t;
  }
  return 0;   // six swaps: the result is back in val0
}

int lattice_slice(const LatticeBufs &L, const float *ins, const float *gate, double *loss_acc, float *outs, int N,
                  int K, int H, int W, cudaStream_t stream) {
  const int n = H * W;
  static int slice_tile = -1;
  if (slice_tile < 0) slice_tile = lattice_env("COSA_SLICE_TILE", 0) ? 1 : 0;   // measured equal: the simple one is the default
  if (lattice_per_pixel() || !slice_tile) {
    const int blocks = persistent_blocks(L.P, 256);
    if (gate) {
      COSA_LAUNCH(lattice_slice_kernel<true>, blocks, 256, 0, stream, L, L.val0, ins, gate, loss_acc, outs, K, n);
    } else {
      COSA_LAUNCH(lattice_slice_kernel<false>, blocks, 256, 0, stream, L, L.val0, ins, gate, loss_acc, outs, K, n);
    }
    return 0;
  }
  const size_t smem = (size_t)kSliceRows * kChunk * sizeof(float);
  static bool attr = false;
  if (!attr) {
    COSA_CUDA(cudaFuncSetAttribute(lattice_slice_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    COSA_CUDA(cudaFuncSetAttribute(lattice_slice_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const dim3 grid(ceil_div(W, kTileW), ceil_div(H, 8), N);
  if (gate) {
    COSA_LAUNCH_T("lattice_slice_tile_kernel", lattice_slice_tile_kernel<true>, grid, 256, smem, stream, L, L.val0, ins,
                  gate, loss_acc, outs, K, H, W);
  } else {
    COSA_LAUNCH_T("lattice_slice_tile_kernel", lattice_slice_tile_kernel<false>, grid, 256, smem, stream, L, L.val0, ins,
                  gate, loss_acc, outs, K, H, W);
  }
  return 0;
}

}  // namespace cosa

using namespace cosa;

extern "C" size_t cosa_bilateral_ws_bytes(int N, int K, int H, int W) {
  if (N < 1 || K < 1 || H < 1 || W < 1) return 0;
  return lattice_ws_bytes(lattice_chunk_images(N, K, H, W), K, H, W);
}

extern "C" int cosa_bilateralfilter_batch(const float *images, const float *ins, float *outs, int N, int K, int H,
                                          int W, float sigmargb, float sigmaxy, void *ws, size_t ws_bytes,
                                          void *stream) {
  if (!images || !ins || !outs || !ws || N < 1 || K < 1 || H < 1 || W < 1) return COSA_E_ARG;
  if (ws_bytes < cosa_bilateral_ws_bytes(N, K, H, W)) return COSA_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)H * W;
  const int chunk = lattice_chunk_images(N, K, H, W);
  for (int n0 = 0; n0 < N; n0 += chunk) {   // chunks share the workspace, stream-ordered
    const int nb = min(chunk, N - n0);
    LatticeBufs L;
    lattice_carve(ws, nb, K, H, W, &L);
    COSA_CHECK(lattice_build(L, images + (size_t)n0 * 3 * n, nb, H, W, sigmargb, sigmaxy, n0 == 0, s));
    COSA_CHECK(lattice_splat_blur(L, ins + (size_t)n0 * K * n, nb, K, H, W, s));
    COSA_CHECK(lattice_slice(L, ins + (size_t)n0 * K * n, nullptr, nullptr, outs + (size_t)n0 * K * n, nb, K, H, W, s));
  }
  return 0;
}

extern "C" int cosa_bilateral_stats(const void *ws, int N, int K, int H, int W, long long stats[4], void *stream) {
  if (!ws || !stats || N < 1 || K < 1 || H < 1 || W < 1) return COSA_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  LatticeBufs L;
  lattice_carve(const_cast<void *>(ws), lattice_chunk_images(N, K, H, W), K, H, W, &L);
  int h[5];
  COSA_CUDA(cudaMemcpyAsync(h, L.counters, sizeof(h), cudaMemcpyDeviceToHost, s));
  COSA_CUDA(cudaStreamSynchronize(s));
  stats[0] = (long long)h[0] + h[4];   // vertices of all chunks of the last call
  stats[1] = h[1]; stats[2] = (long long)(L.cap_mask + 1); stats[3] = h[2];
  return h[1] ? COSA_E_KEYRANGE : 0;
}

extern "C" int cosa_bilateralfilter_batch_host(const float *images, const float *ins, float *outs, int N, int K,
                                               int H, int W, float sigmargb, float sigmaxy) {
  if (!images || !ins || !outs || N < 1 || K < 1 || H < 1 || W < 1) return COSA_E_ARG;
  const size_t n = (size_t)H * W;
  const size_t ws_bytes = cosa_bilateral_ws_bytes(N, K, H, W);
  float *d_img = nullptr, *d_in = nullptr, *d_out = nullptr;
  void *d_ws = nullptr;
  int rc = 0;
  if (cudaMalloc(&d_img, (size_t)N * 3 * n * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&d_in, (size_t)N * K * n * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&d_out, (size_t)N * K * n * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&d_ws, ws_bytes) != cudaSuccess) {
    rc = (int)cudaGetLastError();
    if (rc == 0) rc = (int)cudaErrorMemoryAllocation;
  }
  if (rc == 0) {
    cudaMemcpy(d_img, images, (size_t)N * 3 * n * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(d_in, ins, (size_t)N * K * n * sizeof(float), cudaMemcpyHostToDevice);
    rc = cosa_bilateralfilter_batch(d_img, d_in, d_out, N, K, H, W, sigmargb, sigmaxy, d_ws, ws_bytes, nullptr);
    if (rc == 0) rc = (int)cudaMemcpy(outs, d_out, (size_t)N * K * n * sizeof(float), cudaMemcpyDeviceToHost);
  }
  cudaFree(d_img); cudaFree(d_in); cudaFree(d_out); cudaFree(d_ws);
  return rc;
}
