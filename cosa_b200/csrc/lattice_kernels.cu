// Permutohedral-lattice bilateral filter on the GPU (d = 5).
//
// Reference: utils/bilateralfilter/bilateralfilter.cpp:4-55 (features, per-image driver, batch loop) and
// utils/bilateralfilter/permutohedral.cpp:115-297 (Permutohedral::init, SSE branch), :507-571 (compute).
//
//   tile build one CTA per 32 x 8 pixel tile, one thread per pixel: features -> elevate -> round (half-even) -> rank ->
//              barycentric -> 6 packed vertex keys; the tile's keys are de-duplicated in a shared-memory hash table and
//              the pairs bucketed per distinct vertex (count -> scan -> fill).  No global hash traffic at all.
//   insert     the distinct keys of all tiles (T of them, ~22 % of the 6 n pairs) are found-or-inserted in an
//              open-addressing table in HBM (load first, 64-bit CAS only on an empty slot) sized 2^k >= 2T on the
//              device, so it stays L2-resident; new vertices get dense ids, one global atomic per CTA
//   finish     list entries: table slot -> vertex row; 6 look-ups per vertex -> blur neighbour table
//              (permutohedral.cpp:272-297); value rows zeroed
//   splat      per tile: a quarter-warp per distinct vertex sums its pairs from the tile's staged inputs and issues
//              one reduction per vertex row (permutohedral.cpp:526-534)
//   blur       6 Jacobi passes  new = old + 0.5 * (old[n1] + old[n2])               (permutohedral.cpp:536-552)
//   slice      per tile: the distinct vertex rows are staged in shared memory once, pixels gather from there;
//              out = sum_r bary_r * alpha * values[v_r]  (+ fused gate / energy dot)  (permutohedral.cpp:554-567)
//
// Arithmetic that decides DISCRETE outcomes (rounding, ranks, keys) and the barycentric weights use explicit
// round-to-nearest intrinsics in the reference's operation order, so the vertex set and the weights are
// bit-identical to the CPU code; only the summation order of the splat differs.
#include <math.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "lattice.cuh"

namespace cosa {

// ---- packed keys ---------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ unsigned long long pack_key(const int q[D], int r, int b, int *bad) {
  unsigned long long k = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const int v = q[i] + kQBias;
    if (v < 0 || v >= (1 << kQBits)) *bad = 1;
    k |= (unsigned long long)(v & ((1 << kQBits) - 1)) << (kQBits * i);
  }
  k |= (unsigned long long)r << kKeyRShift;
  k |= (unsigned long long)b << kKeyBShift;
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
  float sf[kMaxLatD];
};

static EmbedConst make_embed_const(int d) {
  EmbedConst c;
  const float inv_std_dev = (float)(sqrt(2.0 / 3.0) * (d + 1));
  for (int i = 0; i < kMaxLatD; ++i) c.sf[i] = i < d ? (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * inv_std_dev) : 0.0f;
  return c;
}

// One point of Permutohedral::init (permutohedral.cpp:176-252), any dimension D.  Outputs q0[i] = rem0[i]/(D+1) (after
// the wrap), rank[i] and the D+1 barycentric weights.  ("6" in the names below is D+1; the path's lattice has D = 5.)
template <int D>
__device__ __forceinline__ void embed_point(const float f[D], const EmbedConst &ec, int q0[D + 1], int rank[D + 1],
                                            float bary[D + 1]) {
  const float inv6 = 1.0f / (float)(D + 1);
  constexpr float six = (float)(D + 1);
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
    rem0[i] = __fmul_rn(v, six);
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
    if (rank[i] < 0) { rank[i] += D + 1; rem0[i] = __fadd_rn(rem0[i], six); }
    else if (rank[i] > D) { rank[i] -= D + 1; rem0[i] = __fsub_rn(rem0[i], six); }
  }
  // barycentric (permutohedral.cpp:222-241): b[5-rank] += v, b[6-rank] -= v, then b[0] += 1 + b[6].
  // rank is a permutation, so b[s] = v(rank = 5-s) - v(rank = 6-s) whatever the visiting order.
  float vr[D + 1];
#pragma unroll
  for (int k = 0; k <= D; ++k) vr[k] = 0.0f;
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    const float v = __fmul_rn(__fsub_rn(el[i], rem0[i]), inv6);
#pragma unroll
    for (int k = 0; k <= D; ++k) vr[k] = (rank[i] == k) ? v : vr[k];   // selects: no dynamically indexed local array
    q0[i] = (int)rintf(__fmul_rn(rem0[i], inv6));   // exact: rem0 is a small multiple of 6
  }
#pragma unroll
  for (int s = 1; s <= D; ++s) bary[s] = __fsub_rn(vr[D - s], vr[D + 1 - s]);
  bary[0] = __fadd_rn(vr[D], __fadd_rn(1.0f, -vr[0]));
}

// ---- build ---------------------------------------------------------------------------------------
// Table capacity of this build: 2^k >= 1.25 T (T = distinct keys summed over the tiles; a vertex is shared by ~2.6
// tiles, so the load factor M / 2^k is ~0.2-0.4, and at most 0.8 when no two tiles share a vertex - there is always an
// empty slot to end a probe sequence), at most what was allocated.
__device__ __forceinline__ unsigned long long live_mask(const LatticeBufs &L) {
  const unsigned t = (unsigned)max(L.counters[5], 1024);
  const unsigned t2 = t + t / 4;
  const unsigned long long cap = 1ULL << (32 - __clz(t2 - 1));
  return min(cap, L.cap_mask + 1) - 1;
}

__global__ void lattice_reset_kernel(int *counters, int first_chunk) {
  counters[4] = first_chunk ? 0 : counters[4] + counters[0];
  counters[0] = 0;
  if (first_chunk) { counters[1] = 0; counters[2] = 0; }
  counters[5] = 0;
  counters[6] = 0;
}

__device__ __forceinline__ unsigned tile_hash64(unsigned long long k) {
  const unsigned lo = (unsigned)k, hi = (unsigned)(k >> 32);
  return ((lo ^ (hi * 0x9E3779B1u)) * 2654435761u) >> 21;      // 11 bits: kTileHS = 2048
}
constexpr int kTileHS = 2048;

// Find-or-insert in the tile's shared-memory table; most keys of a tile are already there (a plain load decides),
// only an empty slot costs a CAS.
__device__ __forceinline__ int tile_insert64(unsigned long long *hkey, unsigned long long key) {
  unsigned h = tile_hash64(key);
  for (;;) {
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(hkey + h);
    if (cur == key) return (int)h;
    if (cur == kEmptyKey) {
      cur = atomicCAS(hkey + h, kEmptyKey, key);
      if (cur == kEmptyKey || cur == key) return (int)h;
    }
    h = (h + 1) & (kTileHS - 1);
  }
}

// The six vertex keys of a point.  canonical[r][rank] is r or r-6 (permutohedral.cpp:148-153), i.e. in (q, r) form
// vertex r has q[i] = q0[i] - [rank[i] >= 6 - r]: going from vertex r-1 to r, exactly the coordinate whose rank is
// 6 - r loses one (the sixth coordinate is not stored), and the residue field gains one.
template <int D>
__device__ __forceinline__ void point_keys(const int q0[D + 1], const int rank[D + 1], int b,
                                           unsigned long long key[D + 1], int *bad) {
  unsigned long long k = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const int v = q0[i] + kQBias;
    if (v < 1 || v >= (1 << kQBits)) *bad = 1;      // v - 1 (the decremented coordinate) must fit as well
    k |= (unsigned long long)(v & ((1 << kQBits) - 1)) << (kQBits * i);
  }
  k |= (unsigned long long)b << kKeyBShift;
  key[0] = k;
#pragma unroll
  for (int r = 1; r <= D; ++r) {
    unsigned long long dec = 0;
#pragma unroll
    for (int i = 0; i < D; ++i)
      if (rank[i] == D + 1 - r) dec = 1ULL << (kQBits * i);
    k = k - dec + (1ULL << kKeyRShift);
    key[r] = k;
  }
}

// Features of pixel (x, y) of image plane set `img` (bilateralfilter.cpp:9-13): D = 5: (x, y) / sigma_xy and
// (R, G, B) / sigma_rgb; D = 2 (the spatial kernel of the dense CRF, densecrf addPairwiseGaussian): (x, y) / sigma_xy.
template <int D>
__device__ __forceinline__ void point_features(float f[D], const float *img, int n, int p, int x, int y, float sigmargb,
                                               float sigmaxy) {
  f[0] = __fdiv_rn((float)x, sigmaxy);
  f[1] = __fdiv_rn((float)y, sigmaxy);
  if constexpr (D == 5) {
    f[2] = __fdiv_rn(__ldg(img + p), sigmargb);
    f[3] = __fdiv_rn(__ldg(img + n + p), sigmargb);
    f[4] = __fdiv_rn(__ldg(img + 2 * n + p), sigmargb);
  }
}

constexpr int kPairFirst = 1 << 11;    // plist code: (pixel in tile << 3) | r, bit 11 = first pair of its vertex

template <int D>
__global__ void __launch_bounds__(kTilePix, 5) lattice_tile_build_kernel(LatticeBufs L, const float *__restrict__ images,
                                                                      EmbedConst ec, int H, int W, int n_pad,
                                                                      float sigmargb, float sigmaxy) {
  constexpr int kLatD = D;                           // (shadows the path's default inside this kernel)
  constexpr int kTilePairs = tile_pairs(D), kPairBlock = pair_block(D), kTileListStride = list_stride(D);
  __shared__ unsigned long long hkey[kTileHS];      // 16 KB
  __shared__ int hinfo[kTileHS];                    //  8 KB  pair count, then (list index << 16) | first pair
  __shared__ __align__(16) unsigned short plist_s[kTileListStride];   // pairs bucketed by vertex + the block table
  __shared__ int s_warp[kTilePix / 32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lx = tid & 31, ly = tid >> 5;
  const int x = blockIdx.x * kTileW + lx, y = blockIdx.y * kTileH + ly, b = blockIdx.z;
  const int n = H * W;
  const bool ok = x < W && y < H;
  for (int i = tid; i < kTileHS; i += kTilePix) { hkey[i] = kEmptyKey; hinfo[i] = 0; }
  for (int i = tid; i < kTileListStride / 2; i += kTilePix) reinterpret_cast<unsigned *>(plist_s)[i] = 0;

  const int p = y * W + x;
  float f[kLatD];
#pragma unroll
  for (int i = 0; i < kLatD; ++i) f[i] = 0.0f;
  if (ok) point_features<D>(f, images + (size_t)b * 3 * n, n, p, x, y, sigmargb, sigmaxy);
  int q0[kLatD + 1], rank[kLatD + 1];
  float bary[kLatD + 1];
  embed_point<D>(f, ec, q0, rank, bary);
  int bad = 0;
  unsigned long long key[kLatD + 1];
  point_keys<D>(q0, rank, b, key, &bad);
  if (!ok) bad = 0;
  __syncthreads();

  int hs[kLatD + 1], pos[kLatD + 1];
#pragma unroll
  for (int r = 0; r <= kLatD; ++r) {
    hs[r] = 0; pos[r] = 0;
    if (ok) {
      hs[r] = tile_insert64(hkey, key[r]);
      pos[r] = atomicAdd(hinfo + hs[r], 1);
    }
  }
  __syncthreads();

  // one scan over the 2048 slots gives every occupied slot its list index and the start of its pair bucket
  constexpr int PER = kTileHS / kTilePix;   // 8 consecutive slots per thread
  int packed[PER], sum = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = hinfo[tid * PER + i];
    packed[i] = c ? ((1 << 16) | c) : 0;    // every occupied slot has at least one pair
    sum += packed[i];
  }
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lx >= o) incl += t;
  }
  if (lx == 31) s_warp[ly] = incl;
  __syncthreads();
  int run = incl - sum, total = 0;
#pragma unroll
  for (int wv = 0; wv < kTilePix / 32; ++wv) {
    const int t = s_warp[wv];
    if (wv < ly) run += t;
    total += t;
  }
  const int U = total >> 16, pairs = total & 0xffff;
  const int tile = (b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  // The SSE loop of the reference also embeds the zero-feature pixels that pad n to a multiple of four
  // (permutohedral.cpp:168-173): their vertices exist (they count in M and take part in the blur) but receive nothing.
  // The first tile of an image appends their six keys behind its own list entries: the insert kernel sees them, the
  // splat and the slice (which walk U entries) do not.
  const bool pad_tile = n_pad > n && blockIdx.x == 0 && blockIdx.y == 0;
  if (tid == 0) {
    const int take = U + (pad_tile ? kLatD + 1 : 0);
    int base = atomicAdd(L.counters + 5, take);
    if ((long long)base + take > L.t_cap) { atomicOr(L.counters + 1, 2); base = -1; }
    s_base = base;
    L.tile_info[tile] = make_int2(max(base, 0), base < 0 ? 0 : (U | (pairs << 16)));
    if (pad_tile && base >= 0) {
      float fz[kLatD];
#pragma unroll
      for (int i = 0; i < kLatD; ++i) fz[i] = 0.0f;
      int qz[kLatD + 1], rz[kLatD + 1], badz = 0;
      float bz[kLatD + 1];
      unsigned long long kz[kLatD + 1];
      embed_point<D>(fz, ec, qz, rz, bz);
      point_keys<D>(qz, rz, b, kz, &badz);
#pragma unroll
      for (int r = 0; r <= kLatD; ++r) L.tkeys[base + U + r] = kz[r];
    }
  }
  __syncthreads();
  const int base = s_base;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    if (packed[i]) {
      const int s = tid * PER + i;
      const int u = run >> 16, start = run & 0xffff, cnt = packed[i] & 0xffff;
      if (base >= 0) L.tkeys[base + u] = hkey[s];
      hinfo[s] = run;
      // block table: the splat walks the pair list in blocks of kPairBlock pairs; block j starts inside vertex u
      for (int j = (start + kPairBlock - 1) / kPairBlock; j * kPairBlock < start + cnt; ++j)
        plist_s[kTilePairs + j] = (unsigned short)u;
      run += packed[i];
    }
  }
  __syncthreads();
  if (ok) {
    const long long gp = (long long)b * n + p;
#pragma unroll
    for (int r = 0; r <= kLatD; ++r) {
      const int info = hinfo[hs[r]];
      plist_s[(info & 0xffff) + pos[r]] = (unsigned short)((tid << 3) | r | (pos[r] == 0 ? kPairFirst : 0));
      L.lidx[(size_t)r * L.P + gp] = (unsigned short)(info >> 16);
      L.bary[(size_t)r * L.P + gp] = bary[r];
    }
  }
  __syncthreads();
  if (tid < kTileListStride / 8)
    reinterpret_cast<uint4 *>(L.plist + (size_t)tile * kTileListStride)[tid] = reinterpret_cast<const uint4 *>(plist_s)[tid];
  if (bad) atomicOr(L.counters + 1, 1);
}

__global__ void lattice_table_clear_kernel(LatticeBufs L) {
  const unsigned long long cap = live_mask(L) + 1;   // a power of two >= 1024
  for (unsigned long long i = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 2; i < cap;
       i += (unsigned long long)gridDim.x * blockDim.x * 2)
    *reinterpret_cast<ulonglong2 *>(L.table_keys + i) = make_ulonglong2(kEmptyKey, kEmptyKey);
}

// The T distinct-per-tile keys into the global table.  A vertex is shared by ~2.6 tiles, so most look-ups find
// the key (plain L2 load); only an empty slot costs a CAS.  A stale "empty" is harmless (the CAS decides); an
// occupied slot never changes again.  New vertices get dense ids: one global atomic per CTA and round.
__global__ void __launch_bounds__(256) lattice_insert_kernel(LatticeBufs L) {
  __shared__ int s_new, s_base;
  const int T = (int)min((long long)L.counters[5], L.t_cap);
  const unsigned long long mask = live_mask(L);
  int max_probe = 0;
  for (int e0 = blockIdx.x * 256; e0 < T; e0 += gridDim.x * 256) {
    if (threadIdx.x == 0) s_new = 0;
    __syncthreads();
    const int e = e0 + threadIdx.x;
    bool is_new = false;
    unsigned long long key = 0, s = 0;
    if (e < T) {
      key = L.tkeys[e];
      s = hash_key(key) & mask;
      unsigned long long c = __ldcg(L.table_keys + s);
      int probes = 0;
      for (;;) {
        if (c == key) break;
        if (c == kEmptyKey) {
          c = atomicCAS(L.table_keys + s, kEmptyKey, key);
          if (c == kEmptyKey) { is_new = true; break; }
          if (c == key) break;
        }
        s = (s + 1) & mask;
        c = __ldcg(L.table_keys + s);
        ++probes;
      }
      max_probe = max(max_probe, probes);
      L.tvid[e] = (int)s;
    }
    const int local = is_new ? atomicAdd(&s_new, 1) : 0;
    __syncthreads();
    if (threadIdx.x == 0 && s_new) s_base = atomicAdd(L.counters, s_new);
    __syncthreads();
    if (is_new) {
      const long long id = (long long)s_base + local;
      if (id < L.m_cap) {
        L.vkeys[id] = key;
        L.table_ids[s] = (int)id + 1;
      } else {
        L.table_ids[s] = 0;
        atomicOr(L.counters + 1, 2);
      }
    }
  }
  if (max_probe > 0) atomicMax(L.counters + 2, max_probe);
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

// (a) list entries: table slot -> vertex row.  (b) blur neighbours (permutohedral.cpp:282-294): n1 = key - 1 on every
// stored coordinate and key[j] + D on axis j, n2 the opposite.  In (q, r) form the residue moves to r-1 / r+1 and q
// carries when it wraps.  Thread (vertex, axis) looks both up in the L2-resident table and writes its own (n1, n2)
// pair: one coalesced 8-byte store per entry.  (Looking up n1 only and recording the symmetric n2 through a scattered
// 4-byte store halved the table walks but made every scattered store a DRAM read-modify-write of its sector: 250 MB
// of reads per build at VOC B = 32.)  Absent neighbours are 0.  (c) the value rows the splat accumulates into are
// zeroed: thread (vertex, j) owns quads j, j + D + 1, ...
template <int D>
__global__ void __launch_bounds__(256) lattice_finish_kernel(LatticeBufs L) {
  constexpr int kLatD = D;
  const long long M = min((long long)L.counters[0], L.m_cap);
  const int T = (int)min((long long)L.counters[5], L.t_cap);
  const long long total = M * (kLatD + 1);
  const unsigned long long qmask = (1ULL << kQBits) - 1;
  const unsigned long long tmask = live_mask(L);
  const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x, gstride = (long long)gridDim.x * blockDim.x;
  for (long long e = gtid; e < T; e += gstride) L.tvid[e] = L.table_ids[L.tvid[e]];
  const int kq = L.Kp / 4;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long idx = gtid; idx < total; idx += gstride) {
    const long long i = idx / (kLatD + 1);
    const int j = (int)(idx % (kLatD + 1));
    for (int q = j; q < kq; q += kLatD + 1) reinterpret_cast<float4 *>(L.val0 + (size_t)(i + 1) * L.Kp)[q] = z;
    const unsigned long long key = L.vkeys[i];
    const int r = (int)((key >> kKeyRShift) & 7);
    const unsigned long long bbits = key >> kKeyBShift;
    int q0[kLatD];
#pragma unroll
    for (int k = 0; k < kLatD; ++k) q0[k] = (int)((key >> (kQBits * k)) & qmask) - kQBias;
    int2 nn;
    {   // n1: coordinates -1, axis +D: residue r - 1 (q carries down when r wraps), axis coordinate one q up
      const int r2 = r == 0 ? kLatD : r - 1, dq = r == 0 ? -1 : 0;
      int qq[kLatD], bad = 0;
#pragma unroll
      for (int k = 0; k < kLatD; ++k) qq[k] = q0[k] + dq + (k == j ? 1 : 0);
      const unsigned long long nk = pack_key<D>(qq, r2, 0, &bad) | (bbits << kKeyBShift);
      nn.x = bad ? 0 : table_find(L, tmask, nk);
    }
    {   // n2: coordinates +1, axis -D: residue r + 1 (q carries up when r wraps), axis coordinate one q down
      const int r2 = r == kLatD ? 0 : r + 1, dq = r == kLatD ? 1 : 0;
      int qq[kLatD], bad = 0;
#pragma unroll
      for (int k = 0; k < kLatD; ++k) qq[k] = q0[k] + dq - (k == j ? 1 : 0);
      const unsigned long long nk = pack_key<D>(qq, r2, 0, &bad) | (bbits << kKeyBShift);
      nn.y = bad ? 0 : table_find(L, tmask, nk);
    }
    L.nbr[(size_t)j * L.m_cap + i] = nn;
  }
  if (blockIdx.x == 0) {
    if (threadIdx.x < kq) {
      reinterpret_cast<float4 *>(L.val0)[threadIdx.x] = z;
      reinterpret_cast<float4 *>(L.val1)[threadIdx.x] = z;
    }
    if (threadIdx.x == 0) {
      L.counters[3] = (int)min(tmask + 1, (unsigned long long)0x7fffffff);
      L.counters[6] = 1;
    }
  }
}

// ---- compute -------------------------------------------------------------------------------------
// Only needed when a lattice is splatted a second time: the build leaves val0 zeroed (counters[6]).
__global__ void lattice_zero_values_kernel(LatticeBufs L) {
  if (L.counters[6]) return;
  const long long rows = min((long long)L.counters[0], L.m_cap) + 1;
  const long long total4 = rows * (L.Kp / 4);
  float4 *v0 = reinterpret_cast<float4 *>(L.val0);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x)
    v0[i] = z;
}

constexpr int kChunk = 24;                       // channels per pass of the K = 21 instantiations (one pass)
constexpr int kChunkWide = 28;                   // ... of the run-time-K instantiations (K = 81 -> 84 = three passes)
constexpr int kPlane = kTilePix + 1;             // odd plane pitch: channel-strided reads hit distinct banks

// Splat (permutohedral.cpp:526-534), one CTA per tile.  The tile's pair list (bucketed by vertex in the build), the
// weights and the input channels are staged in shared memory.  The 32 quarter-warps of the CTA each take one block of
// kPairBlock consecutive pairs of the list - the same amount of work whatever the vertex degrees - and lane cl sums
// channels cl, cl + 8, cl + 16; a quarter-warp flushes its partial sum with one reduction per vertex row whenever the
// list moves on to the next vertex (kPairFirst).  KT > 0: compile-time channel count (one pass); 0: run-time K.
template <int KT, int D, int CHUNK>
__global__ void __launch_bounds__(kTilePix) lattice_splat_tile_kernel(LatticeBufs L, const float *__restrict__ ins,
                                                                      int Krt, int H, int W) {
  static_assert(KT <= CHUNK, "a compile-time channel count must fit one pass");
  constexpr int kLatD = D;
  constexpr int kTilePairs = tile_pairs(D), kPairBlock = pair_block(D), kTileListStride = list_stride(D);
  extern __shared__ __align__(16) unsigned char s_raw[];
  float *in_s = reinterpret_cast<float *>(s_raw);                         // [CHUNK][kPlane]
  float *bary_s = in_s + CHUNK * kPlane;                                 // [D + 1][256]
  int *vid_s = reinterpret_cast<int *>(bary_s + (D + 1) * kTilePix);      // [kTilePairs]
  unsigned short *plist_s = reinterpret_cast<unsigned short *>(vid_s + kTilePairs);   // [kTileListStride], 16-byte aligned

  const int K = KT ? KT : Krt, Kp = (K + 3) & ~3;
  const int tid = threadIdx.x, lx = tid & 31, ly = tid >> 5;
  const int x = blockIdx.x * kTileW + lx, y = blockIdx.y * kTileH + ly, b = blockIdx.z;
  const int n = H * W;
  const bool ok = x < W && y < H;
  const int tile = (b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const int2 info = L.tile_info[tile];
  const int U = info.y & 0xffff, pairs = info.y >> 16;
  const long long gp = (long long)b * n + (long long)y * W + x;
  {
    const float *bp = L.bary + gp;
#pragma unroll
    for (int r = 0; r <= kLatD; ++r) { bary_s[r * kTilePix + tid] = ok ? __ldg(bp) : 0.0f; bp += L.P; }
  }
  if (tid < kTileListStride / 8)
    reinterpret_cast<uint4 *>(plist_s)[tid] = reinterpret_cast<const uint4 *>(L.plist + (size_t)tile * kTileListStride)[tid];
  for (int u = tid; u < U; u += kTilePix) vid_s[u] = L.tvid[info.x + u];
  if (tile == 0 && tid == 0) L.counters[6] = 0;   // val0 is being written

  const int sub = lx >> 3, cl = tid & 7;
  const int j = ly * 4 + sub;                     // this quarter-warp's block of the pair list
  const int e0 = j * kPairBlock, e1 = min(e0 + kPairBlock, pairs);
  const float *src0 = ins + (size_t)b * K * n + (size_t)y * W + x;
  for (int c0 = 0; c0 < Kp; c0 += CHUNK) {
    const int kc = min(CHUNK, Kp - c0);
    if (c0) __syncthreads();   // the previous pass has read in_s
    {
      float v[CHUNK];
      const float *sp = src0 + (size_t)c0 * n;
#pragma unroll
      for (int c = 0; c < CHUNK; ++c) {
        v[c] = (ok && c0 + c < K) ? __ldg(sp) : 0.0f;
        sp += n;
      }
#pragma unroll
      for (int c = 0; c < CHUNK; ++c) in_s[c * kPlane + tid] = v[c];
    }
    __syncthreads();
    if (e0 < e1) {
      int u = plist_s[kTilePairs + j];
      constexpr int NA = (CHUNK + 7) / 8;      // channels cl, cl + 8, ... of the pass
      float acc[NA];
#pragma unroll
      for (int i = 0; i < NA; ++i) acc[i] = 0.0f;
      auto flush = [&]() {
        const int vid = vid_s[u];
        if (vid > 0) {
          float *dst = L.val0 + (size_t)vid * Kp + c0 + cl;
#pragma unroll
          for (int i = 0; i < NA; ++i)
            if (cl + 8 * i < kc) atomicAdd(dst + 8 * i, acc[i]);
        }
      };
      const float *inl = in_s + cl * kPlane;
#pragma unroll 4
      for (int e = e0; e < e1; ++e) {
        const int code = plist_s[e];
        if ((code & kPairFirst) && e != e0) {
          flush();
          ++u;
#pragma unroll
          for (int i = 0; i < NA; ++i) acc[i] = 0.0f;
        }
        const int pix = (code >> 3) & 0xff;
        const float wv = bary_s[(code & 7) * kTilePix + pix];
#pragma unroll
        for (int i = 0; i < NA; ++i)
          if (8 * i + 7 < CHUNK || cl + 8 * i < CHUNK) acc[i] = fmaf(wv, inl[8 * i * kPlane + pix], acc[i]);
      }
      flush();
    }
  }
}

// One blur axis (permutohedral.cpp:538-551): thread per (vertex, 4 channels).
__global__ void __launch_bounds__(256) lattice_blur_kernel(LatticeBufs L, const float *__restrict__ src,
                                                           float *__restrict__ dst, int axis) {
  const int kq = L.Kp / 4;
  const long long total = min((long long)L.counters[0], L.m_cap) * kq;
  const int2 *nb = L.nbr + (size_t)axis * L.m_cap;
  // row 0 of `dst` is the "absent neighbour" row of the next pass: zero it here, so that no pass depends on what a
  // buffer held before (views of one lattice carved for different channel counts place val1 at different offsets)
  if (blockIdx.x == 0 && threadIdx.x < kq) reinterpret_cast<float4 *>(dst)[threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
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

// Slice (permutohedral.cpp:554-567) + optional dense-CRF epilogue (seg_helper.py:888-890), one CTA per tile.  The
// tile's distinct vertex rows (up to kSliceRows of them; list entries beyond that are read from L2 directly) are staged
// in shared memory once per channel pass and the pixels gather from there through their 16-bit list indices.
// KT > 0: compile-time channel count; 0: run-time K.
constexpr int kSliceRows = 512;
constexpr int kRowPitch = 28;                     // >= both chunk sizes; 7 quads: row starts fall on 8 different bank quads
template <bool ENERGY, int KT, int D, int CHUNK>
__global__ void __launch_bounds__(kTilePix, 3) lattice_slice_tile_kernel(LatticeBufs L, const float *__restrict__ values,
                                                                         const float *__restrict__ ins,
                                                                         const float *__restrict__ gate, double *loss_acc,
                                                                         float *__restrict__ outs, int Krt, int H, int W) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  float *rows = reinterpret_cast<float *>(s_raw);   // [kSliceRows][kRowPitch]
  __shared__ int vid_s[kSliceRows];
  __shared__ float s_part[kTilePix / 32];
  static_assert(KT <= CHUNK, "a compile-time channel count must fit one pass");
  constexpr int kLatD = D;
  constexpr int KQ = KT ? (KT + 3) / 4 : 1;       // row quads per vertex (compile-time form of kq)
  const int K = KT ? KT : Krt, Kp = (K + 3) & ~3;
  const float alpha = 1.0f / (1.0f + 1.0f / (float)(1 << D));   // 1 / (1 + 2^-d), exact
  const int tid = threadIdx.x, lx = tid & 31, ly = tid >> 5;
  const int x = blockIdx.x * kTileW + lx, y = blockIdx.y * kTileH + ly, b = blockIdx.z;
  const int n = H * W;
  const bool ok = x < W && y < H;
  const int tile = (b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const int2 info = L.tile_info[tile];
  const int Us = min(info.y & 0xffff, kSliceRows);
  const long long gp = (long long)b * n + (long long)y * W + x;
  int u_r[kLatD + 1];
  float w[kLatD + 1];
  {
    const unsigned short *ip = L.lidx + gp;
    const float *bp = L.bary + gp;
#pragma unroll
    for (int r = 0; r <= kLatD; ++r) {
      u_r[r] = ok ? (int)__ldg(ip) : 0;
      w[r] = ok ? __fmul_rn(__ldg(bp), alpha) : 0.0f;
      ip += L.P; bp += L.P;
    }
  }
  for (int u = tid; u < Us; u += kTilePix) vid_s[u] = L.tvid[info.x + u];
  const float gt = (ENERGY && ok) ? __ldg(gate + gp) : 1.0f;
  const size_t at0 = (size_t)b * K * n + (size_t)y * W + x;
  float local = 0.0f;
  for (int c0 = 0; c0 < Kp; c0 += CHUNK) {
    const int kc = min(CHUNK, Kp - c0), kq = kc >> 2;
    __syncthreads();                     // vid_s is complete / the previous pass has read `rows`
    for (int i0 = tid; i0 < Us * kq; i0 += 4 * kTilePix) {   // four row quads in flight per thread
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * kTilePix;
        const int u = KT ? i / KQ : i / kq, q = i - u * kq;
        if (i < Us * kq) v[k] = *reinterpret_cast<const float4 *>(values + (size_t)vid_s[u] * Kp + c0 + 4 * q);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * kTilePix;
        const int u = KT ? i / KQ : i / kq, q = i - u * kq;
        if (i < Us * kq) *reinterpret_cast<float4 *>(rows + u * kRowPitch + 4 * q) = v[k];
      }
    }
    __syncthreads();
    if (ok) {
      const float *ip = ins + at0 + (size_t)c0 * n;
      float *op = outs + at0 + (size_t)c0 * n;
      // energy epilogue: the pixel's own inputs, requested one channel quad ahead of their use
      float nxt[4] = {0.f, 0.f, 0.f, 0.f};
      if (ENERGY) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c0 + k < K) nxt[k] = __ldg(ip + (size_t)k * n);
      }
#pragma unroll
      for (int q = 0; q < CHUNK / 4; ++q) {
        if (q >= kq) break;
        float cur[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) cur[k] = nxt[k];
        if (ENERGY && q + 1 < kq) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c0 + 4 * (q + 1) + k < K) nxt[k] = __ldg(ip + (size_t)(4 + k) * n);
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r <= kLatD; ++r) {
          const float4 v = u_r[r] < kSliceRows
                               ? *reinterpret_cast<const float4 *>(rows + u_r[r] * kRowPitch + 4 * q)
                               : *reinterpret_cast<const float4 *>(values + (size_t)L.tvid[info.x + u_r[r]] * Kp + c0 + 4 * q);
          acc.x = __fadd_rn(acc.x, __fmul_rn(w[r], v.x));
          acc.y = __fadd_rn(acc.y, __fmul_rn(w[r], v.y));
          acc.z = __fadd_rn(acc.z, __fmul_rn(w[r], v.z));
          acc.w = __fadd_rn(acc.w, __fmul_rn(w[r], v.w));
        }
        const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (c0 + 4 * q + k < K) {
            float o = a4[k];
            if (ENERGY) {
              o = __fmul_rn(o, gt);
              local = fmaf(cur[k], o, local);
            }
            *op = o;
            ip += n; op += n;
          }
        }
      }
    }
  }
  // key range / capacity error: fail loudly, never alias silently - every output and the energy become NaN
  if (L.counters[1] != 0) {
    const float qnan = __int_as_float(0x7fc00000);
    if (ok)
      for (int c = 0; c < K; ++c) outs[at0 + (size_t)c * n] = qnan;
    local = qnan;
  }
  if (ENERGY) {
    local = warp_sum(local);
    if (lx == 0) s_part[ly] = local;
    __syncthreads();
    if (tid < 32) {
      float t = tid < kTilePix / 32 ? s_part[tid] : 0.0f;
      t = warp_sum(t);
      if (tid == 0) atomicAdd(loss_acc, (double)t);
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------
static unsigned long long table_capacity(long long t_cap) {
  unsigned long long cap = 1024;
  // t_cap is the worst case (six new list entries per pixel); the build uses the first 2^k >= 1.25 T slots, so the
  // allocation only bounds the load factor in the worst case (<= 0.8)
  while (cap < (unsigned long long)t_cap + (unsigned long long)t_cap / 4) cap <<= 1;
  return cap;
}

// Images per lattice: up to kMaxImagesPerLattice (the key's image field).  Smaller chunks whose two value buffers fit
// in L2 were measured in round 1 and are slower at every size (the blur passes are not HBM-bound and smaller launches
// fill the GPU worse).
int lattice_chunk_images(int N, int K, int H, int W) {
  (void)K; (void)H; (void)W;
  const int c = min(kMaxImagesPerLattice, N);
  const int chunks = (N + c - 1) / c;
  return (N + chunks - 1) / chunks;   // even chunks
}

struct LatticeDims {
  long long n, n_pad, P, m_cap, t_cap, tiles;
  unsigned long long cap;
  int tiles_x, tiles_y, Kp, d;
};

static LatticeDims lattice_dims(int N, int K, int H, int W, int dim, float vpp) {
  LatticeDims d;
  d.d = dim;
  const long long V = dim + 1;             // vertices of a simplex
  d.n = (long long)H * W;
  d.n_pad = (d.n + 3) & ~3LL;
  d.P = (long long)N * d.n;
  d.t_cap = V * N * d.n_pad + V * N;       // + the keys of the padding pixels, appended once per image
  d.m_cap = d.t_cap;
  if (vpp > 0.0f) {                        // a vertex budget instead of the worst case (lattice.cuh)
    d.m_cap = min(d.m_cap, (long long)ceil((double)vpp * (double)d.P) + V * N + 1024);
    d.t_cap = min(d.t_cap, 2 * d.m_cap);
  }
  d.cap = table_capacity(d.t_cap);
  d.tiles_x = ceil_div(W, kTileW);
  d.tiles_y = ceil_div(H, kTileH);
  d.tiles = (long long)N * d.tiles_x * d.tiles_y;
  d.Kp = (K + 3) & ~3;
  return d;
}

template <typename F>
static void lattice_layout(const LatticeDims &d, F &&take) {
  take(0, (size_t)8 * sizeof(int));                                   // counters first: same place for every chunk size
  take(1, (size_t)d.cap * sizeof(unsigned long long));                // table_keys
  take(2, (size_t)d.cap * sizeof(int));                               // table_ids
  take(3, (size_t)d.m_cap * sizeof(unsigned long long));              // vkeys
  take(4, (size_t)d.tiles * sizeof(int2));                            // tile_info
  take(5, (size_t)d.t_cap * sizeof(unsigned long long));              // tkeys
  take(6, (size_t)d.t_cap * sizeof(int));                             // tvid
  take(7, (size_t)256);                                               // (unused)
  take(8, (size_t)d.tiles * list_stride(d.d) * sizeof(unsigned short));  // plist + block table
  take(9, (size_t)(d.d + 1) * d.P * sizeof(unsigned short));          // lidx
  take(10, (size_t)(d.d + 1) * d.P * sizeof(float));                  // bary
  take(11, (size_t)(d.d + 1) * d.m_cap * sizeof(int2));               // nbr
  take(12, (size_t)(d.m_cap + 1) * d.Kp * sizeof(float));             // val0
  take(13, (size_t)(d.m_cap + 1) * d.Kp * sizeof(float));             // val1
}

size_t lattice_ws_bytes(int N, int K, int H, int W, int dim, float vpp) {
  size_t b = 0;
  lattice_layout(lattice_dims(N, K, H, W, dim, vpp), [&](int, size_t bytes) { b += align_up(bytes, 256); });
  return b;
}

void lattice_carve(void *ws, int N, int K, int H, int W, LatticeBufs *L, int dim, float vpp) {
  const LatticeDims d = lattice_dims(N, K, H, W, dim, vpp);
  L->P = d.P; L->m_cap = d.m_cap; L->t_cap = d.t_cap; L->cap_mask = d.cap - 1;
  L->tiles_x = d.tiles_x; L->tiles_y = d.tiles_y; L->Kp = d.Kp; L->d = d.d;
  char *p = (char *)ws;
  void *slot[14];
  lattice_layout(d, [&](int i, size_t bytes) { slot[i] = p; p += align_up(bytes, 256); });
  L->counters = (int *)slot[0];
  L->table_keys = (unsigned long long *)slot[1];
  L->table_ids = (int *)slot[2];
  L->vkeys = (unsigned long long *)slot[3];
  L->tile_info = (int2 *)slot[4];
  L->tkeys = (unsigned long long *)slot[5];
  L->tvid = (int *)slot[6];
  L->plist = (unsigned short *)slot[8];
  L->lidx = (unsigned short *)slot[9];
  L->bary = (float *)slot[10];
  L->nbr = (int2 *)slot[11];
  L->val0 = (float *)slot[12];
  L->val1 = (float *)slot[13];
}

int lattice_build(const LatticeBufs &L, const float *images, int N, int H, int W, float sigmargb, float sigmaxy,
                  bool first_chunk, cudaStream_t stream) {
  if (N < 1 || N > kMaxImagesPerLattice) return COSA_E_ARG;
  if (L.tiles_y > 65535 || (L.d != 5 && L.d != 2)) return COSA_E_ARG;
  const long long n = (long long)H * W;
  const int n_pad = (int)((n + 3) & ~3LL);
  COSA_LAUNCH(lattice_reset_kernel, 1, 1, 0, stream, L.counters, first_chunk ? 1 : 0);
  const dim3 grid(L.tiles_x, L.tiles_y, N);
  if (L.d == 5) {
    COSA_LAUNCH_T("lattice_tile_build_kernel", lattice_tile_build_kernel<5>, grid, kTilePix, 0, stream, L, images,
                  make_embed_const(5), H, W, n_pad, sigmargb, sigmaxy);
  } else {
    COSA_LAUNCH_T("lattice_tile_build_kernel", lattice_tile_build_kernel<2>, grid, kTilePix, 0, stream, L, images,
                  make_embed_const(2), H, W, n_pad, sigmargb, sigmaxy);
  }
  // T and M live on the device: size the grids for the SMs and let the kernels read the counts
  COSA_LAUNCH(lattice_table_clear_kernel, sm_count() * 8, 256, 0, stream, L);
  COSA_LAUNCH(lattice_insert_kernel, sm_count() * 8, 256, 0, stream, L);
  if (L.d == 5) {
    COSA_LAUNCH_T("lattice_finish_kernel", lattice_finish_kernel<5>, sm_count() * 8, 256, 0, stream, L);
  } else {
    COSA_LAUNCH_T("lattice_finish_kernel", lattice_finish_kernel<2>, sm_count() * 8, 256, 0, stream, L);
  }
  return 0;
}

static size_t splat_smem_bytes(int d, int chunk) {
  return (size_t)chunk * kPlane * 4 + (size_t)(d + 1) * kTilePix * 4 + (size_t)tile_pairs(d) * 4 +
         (size_t)list_stride(d) * 2;
}

int lattice_splat_blur(const LatticeBufs &L, const float *ins, int N, int K, int H, int W, cudaStream_t stream) {
  if (L.Kp != ((K + 3) & ~3)) return COSA_E_ARG;   // the view must have been carved for this channel count
  COSA_LAUNCH(lattice_zero_values_kernel, sm_count() * 8, 256, 0, stream, L);
  const dim3 grid(L.tiles_x, L.tiles_y, N);
  if (L.d == 2) {
    COSA_LAUNCH_T("lattice_splat_tile_kernel", (lattice_splat_tile_kernel<0, 2, kChunkWide>), grid, kTilePix, splat_smem_bytes(2, kChunkWide),
                  stream, L, ins, K, H, W);
  } else if (K == 21) {
    COSA_LAUNCH_T("lattice_splat_tile_kernel", (lattice_splat_tile_kernel<21, 5, kChunk>), grid, kTilePix, splat_smem_bytes(5, kChunk),
                  stream, L, ins, K, H, W);
  } else {
    COSA_LAUNCH_T("lattice_splat_tile_kernel", (lattice_splat_tile_kernel<0, 5, kChunkWide>), grid, kTilePix, splat_smem_bytes(5, kChunkWide),
                  stream, L, ins, K, H, W);
  }
  float *src = L.val0, *dst = L.val1;
  for (int axis = 0; axis <= L.d; ++axis) {
    COSA_LAUNCH(lattice_blur_kernel, sm_count() * 8, 256, 0, stream, L, src, dst, axis);
    float *t = src; src = dst; dst = t;
  }
  // d + 1 swaps: for d = 5 the result is back in val0; for d = 2 (three passes) it is in val1 - the slice reads `src`
  return 0;
}

int lattice_slice(const LatticeBufs &L, const float *ins, const float *gate, double *loss_acc, float *outs, int N,
                  int K, int H, int W, cudaStream_t stream) {
  const size_t smem = (size_t)kSliceRows * kRowPitch * sizeof(float);   // 56 KB: opt-in, per device
  static std::atomic<unsigned long long> attr_done{0};
  int dev = 0;
  COSA_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = 1ULL << (dev & 63);
  if (dev >= 64 || !(attr_done.load(std::memory_order_acquire) & bit)) {
    COSA_CUDA(cudaFuncSetAttribute(lattice_slice_tile_kernel<true, 21, 5, kChunk>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    COSA_CUDA(cudaFuncSetAttribute(lattice_slice_tile_kernel<true, 0, 5, kChunkWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    COSA_CUDA(cudaFuncSetAttribute(lattice_slice_tile_kernel<false, 0, 5, kChunkWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    COSA_CUDA(cudaFuncSetAttribute(lattice_slice_tile_kernel<false, 0, 2, kChunkWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev < 64) attr_done.fetch_or(bit, std::memory_order_release);
  }
  const dim3 grid(L.tiles_x, L.tiles_y, N);
  // d + 1 blur passes ping-pong between val0 and val1: the result is in val0 for odd d, in val1 for even d
  const float *values = ((L.d + 1) & 1) ? L.val1 : L.val0;
  if (L.d == 2) {
    if (gate) return COSA_E_ARG;   // the energy epilogue belongs to the bilateral (d = 5) filter
    COSA_LAUNCH_T("lattice_slice_tile_kernel", (lattice_slice_tile_kernel<false, 0, 2, kChunkWide>), grid, kTilePix, smem, stream, L,
                  values, ins, gate, loss_acc, outs, K, H, W);
  } else if (gate && K == 21) {
    COSA_LAUNCH_T("lattice_slice_tile_kernel", (lattice_slice_tile_kernel<true, 21, 5, kChunk>), grid, kTilePix, smem, stream, L,
                  values, ins, gate, loss_acc, outs, K, H, W);
  } else if (gate) {
    COSA_LAUNCH_T("lattice_slice_tile_kernel", (lattice_slice_tile_kernel<true, 0, 5, kChunkWide>), grid, kTilePix, smem, stream, L,
                  values, ins, gate, loss_acc, outs, K, H, W);
  } else {
    COSA_LAUNCH_T("lattice_slice_tile_kernel", (lattice_slice_tile_kernel<false, 0, 5, kChunkWide>), grid, kTilePix, smem, stream, L,
                  values, ins, gate, loss_acc, outs, K, H, W);
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
  int h[6];
  COSA_CUDA(cudaMemcpyAsync(h, L.counters, sizeof(h), cudaMemcpyDeviceToHost, s));
  COSA_CUDA(cudaStreamSynchronize(s));
  stats[0] = (long long)h[0] + h[4];   // vertices of all chunks of the last call
  stats[1] = h[1]; stats[2] = h[3]; stats[3] = h[2];
  if (h[1] & 2) return COSA_E_WORKSPACE;
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
    if (rc == 0) {   // the host form is synchronous: report a key-range / capacity error instead of returning NaNs
      long long st[4];
      rc = cosa_bilateral_stats(d_ws, N, K, H, W, st, nullptr);
    }
  }
  cudaFree(d_img); cudaFree(d_in); cudaFree(d_out); cudaFree(d_ws);
  return rc;
}
