/*
 * ORACLE — test infrastructure only.  NOT part of the product path.
 *
 * Plain-C, single-precision restatement of the reference's CPU bilateral filter
 * (permutohedral lattice, d = 5) as it is actually compiled on x86-64, i.e. the
 * SSE code path of utils/bilateralfilter/permutohedral.cpp with round-half-even
 * (_mm_cvtps_epi32) and no FMA contraction.  Build with -ffp-contract=off.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (cosa_b200/) never does.
 *
 * Parity pin: bit-exact against the unmodified reference C++ built into
 * oracle/_ref/libbf_ref.so (tests/test_oracle_golden.py::test_c_oracle_is_bit_exact_against_reference_build) and
 * against the committed golden vectors tests/golden/bilateral_*.npz generated from that build; the generic-feature
 * entry (any D) against the reference's own Permutohedral class driven through oracle/ref_lattice_shim.cpp
 * (tests/golden/crf_inference.npz).
 *
 * Reference map (all under /root/reference/utils/bilateralfilter/):
 *   features            bilateralfilter.cpp:4-19
 *   per-image driver    bilateralfilter.cpp:22-40, batch loop :42-55
 *   lattice embedding   permutohedral.cpp:115-254   (Permutohedral::init, SSE branch)
 *   blur neighbours     permutohedral.cpp:256-297
 *   splat/blur/slice    permutohedral.cpp:507-571   (compute(float*), SSE branch, value_size = 1)
 *   hash table          permutohedral.cpp:13-100    (only the key SET matters for the result;
 *                                                    ids here are insertion-ordered like the reference)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Lattice dimension: 5 for the path's bilateral filter (liblattice_oracle.so); the Makefile also builds the file with
 * -DD=2 (liblattice_oracle_d2.so) for the spatial kernel of the dense-CRF inference oracle (oracle/crf_oracle.py). */
#ifndef D
#define D 5
#endif
#define D1 (D + 1)

typedef struct {
  int16_t *keys;   /* [filled][D] */
  int *table;      /* [cap] -> id or -1 */
  size_t cap, filled;
} KeyTable;

static size_t key_hash(const int16_t *k) {
  /* permutohedral.cpp:42-49 */
  size_t r = 0;
  for (int i = 0; i < D; i++) { r += (size_t)(long)k[i]; r *= 1664525u; }
  return r;
}

static void kt_init(KeyTable *t, size_t max_keys) {
  size_t cap = 16;
  while (cap < 4 * max_keys) cap <<= 1;   /* never grows: load factor <= 1/4 */
  t->cap = cap; t->filled = 0;
  t->table = (int *)malloc(cap * sizeof(int));
  memset(t->table, -1, cap * sizeof(int));
  t->keys = (int16_t *)malloc((max_keys + 16) * D * sizeof(int16_t));
}
static void kt_free(KeyTable *t) { free(t->table); free(t->keys); }

static int kt_find(KeyTable *t, const int16_t *k, int create) {
  /* permutohedral.cpp:66-96: linear probing, ids in insertion order */
  size_t h = key_hash(k) & (t->cap - 1);
  for (;;) {
    int e = t->table[h];
    if (e < 0) {
      if (!create) return -1;
      memcpy(t->keys + t->filled * D, k, D * sizeof(int16_t));
      t->table[h] = (int)t->filled;
      return (int)t->filled++;
    }
    if (memcmp(t->keys + (size_t)e * D, k, D * sizeof(int16_t)) == 0) return e;
    h = (h + 1) & (t->cap - 1);
  }
}

typedef struct {
  int n, M;
  int *offset;      /* [n_pad][D1] */
  float *bary;      /* [n_pad][D1] */
  int *nbr;         /* [D1][M][2] */
  int16_t *vkeys;   /* [M][D] */
} Lattice;

static float round_half_even(float v) {
  /* _mm_cvtepi32_ps(_mm_cvtps_epi32(v)) under MXCSR round-to-nearest, permutohedral.cpp:190-194 */
  return (float)lrintf(v);
}

/* One pixel of Permutohedral::init (permutohedral.cpp:176-252). */
static void embed_point(const float f[D], const float sf[D], int16_t keys_out[D1][D], float bary_out[D1]) {
  const float invdplus1 = 1.0f / (D + 1);
  const float dplus1 = (float)(D + 1);
  float el[D1], rem0[D1], rank[D1];
  float sm = 0.0f;
  for (int j = D; j > 0; j--) {
    float cf = f[j - 1] * sf[j - 1];
    el[j] = sm - (float)j * cf;
    sm += cf;
  }
  el[0] = sm;
  float sum = 0.0f;
  for (int i = 0; i <= D; i++) {
    float v = round_half_even(invdplus1 * el[i]);
    rem0[i] = v * dplus1;
    sum += v;
  }
  for (int i = 0; i <= D; i++) rank[i] = 0.0f;
  for (int i = 0; i < D; i++) {
    float di = el[i] - rem0[i];
    for (int j = i + 1; j <= D; j++) {
      float dj = el[j] - rem0[j];
      float c = (di < dj) ? 1.0f : 0.0f;
      rank[i] += c;
      rank[j] += 1.0f - c;
    }
  }
  for (int i = 0; i <= D; i++) {
    rank[i] += sum;
    float add = (rank[i] < 0.0f) ? dplus1 : 0.0f;
    float sub = (rank[i] >= dplus1) ? dplus1 : 0.0f;
    rank[i] += add - sub;
    rem0[i] += add - sub;
  }
  float b[D + 2];
  for (int i = 0; i < D + 2; i++) b[i] = 0.0f;
  for (int i = 0; i <= D; i++) {
    float v = (el[i] - rem0[i]) * invdplus1;
    int p = D - (int)rank[i];
    b[p] += v;
    b[p + 1] -= v;
  }
  b[0] += 1.0f + b[D + 1];
  for (int r = 0; r <= D; r++) {
    for (int i = 0; i < D; i++) {
      int rk = (int)rank[i];
      int canon = (rk <= D - r) ? r : r - (D + 1);   /* permutohedral.cpp:148-153 */
      keys_out[r][i] = (int16_t)(rem0[i] + (float)canon);
    }
    bary_out[r] = b[r];
  }
}

static void scale_factors(float sf[D]) {
  /* permutohedral.cpp:156-159: double arithmetic, float inv_std_dev, result stored as float */
  float inv_std_dev = (float)(sqrt(2.0 / 3.0) * (D + 1));
  for (int i = 0; i < D; i++) sf[i] = (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * inv_std_dev);
}

/* features [n][D] (point-major, as Permutohedral::init takes them) */
static void lattice_build_features(Lattice *L, const float *feat, int n) {
  const int n_pad = (n + 3) & ~3;   /* the SSE loop also embeds the zero-feature padding pixels (:168-173, :241) */
  float sf[D];
  scale_factors(sf);
  L->n = n;
  L->offset = (int *)calloc((size_t)(n_pad + 16) * D1, sizeof(int));
  L->bary = (float *)calloc((size_t)(n_pad + 16) * D1, sizeof(float));
  KeyTable kt;
  kt_init(&kt, (size_t)n_pad * D1);
  for (int p = 0; p < n_pad; p++) {
    float f[D];
    for (int k = 0; k < D; k++) f[k] = p < n ? feat[(size_t)p * D + k] : 0.0f;
    int16_t keys[D1][D];
    float bw[D1];
    embed_point(f, sf, keys, bw);
    for (int r = 0; r <= D; r++) {
      L->offset[(size_t)p * D1 + r] = kt_find(&kt, keys[r], 1);
      L->bary[(size_t)p * D1 + r] = bw[r];
    }
  }
  const int M = (int)kt.filled;
  L->M = M;
  L->nbr = (int *)malloc((size_t)D1 * M * 2 * sizeof(int) + 8);
  L->vkeys = (int16_t *)malloc((size_t)M * D * sizeof(int16_t) + 8);
  memcpy(L->vkeys, kt.keys, (size_t)M * D * sizeof(int16_t));
  for (int j = 0; j <= D; j++) {
    for (int i = 0; i < M; i++) {
      const int16_t *key = kt.keys + (size_t)i * D;
      int16_t n1[D1], n2[D1];
      for (int k = 0; k < D; k++) { n1[k] = (int16_t)(key[k] - 1); n2[k] = (int16_t)(key[k] + 1); }
      /* permutohedral.cpp:289-290 writes index j even for j == D (the implicit coordinate) */
      if (j < D) { n1[j] = (int16_t)(key[j] + D); n2[j] = (int16_t)(key[j] - D); }
      L->nbr[((size_t)j * M + i) * 2 + 0] = kt_find(&kt, n1, 0);
      L->nbr[((size_t)j * M + i) * 2 + 1] = kt_find(&kt, n2, 0);
    }
  }
  kt_free(&kt);
}

#if D == 5
static void lattice_build(Lattice *L, const float *image, int H, int W, float sigmargb, float sigmaxy) {
  const int n = H * W;
  float *feat = (float *)malloc((size_t)n * D * sizeof(float));
  for (int p = 0; p < n; p++) {
    int i = p % W, j = p / W;
    feat[(size_t)p * D + 0] = (float)i / sigmaxy;                  /* bilateralfilter.cpp:9-13 */
    feat[(size_t)p * D + 1] = (float)j / sigmaxy;
    feat[(size_t)p * D + 2] = image[0 * n + p] / sigmargb;
    feat[(size_t)p * D + 3] = image[1 * n + p] / sigmargb;
    feat[(size_t)p * D + 4] = image[2 * n + p] / sigmargb;
  }
  lattice_build_features(L, feat, n);
  free(feat);
}
#endif

static void lattice_free(Lattice *L) { free(L->offset); free(L->bary); free(L->nbr); free(L->vkeys); }

/* compute(float*, value_size = 1): permutohedral.cpp:507-571 */
static void lattice_filter(const Lattice *L, float *out, const float *in, float *values, float *new_values) {
  const int n = L->n, M = L->M;
  for (int i = 0; i < M + 2; i++) values[i] = new_values[i] = 0.0f;
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= D; j++) {
      int o = L->offset[(size_t)i * D1 + j] + 1;
      float w = L->bary[(size_t)i * D1 + j];
      values[o] += w * in[i];
    }
  for (int j = 0; j <= D; j++) {
    for (int i = 0; i < M; i++) {
      int n1 = L->nbr[((size_t)j * M + i) * 2 + 0] + 1;
      int n2 = L->nbr[((size_t)j * M + i) * 2 + 1] + 1;
      new_values[i + 1] = values[i + 1] + 0.5f * (values[n1] + values[n2]);
    }
    float *t = values; values = new_values; new_values = t;
  }
  const float alpha = 1.0f / (1 + powf(2, -D));
  for (int i = 0; i < n; i++) {
    float acc = 0.0f;
    for (int j = 0; j <= D; j++) {
      int o = L->offset[(size_t)i * D1 + j] + 1;
      float w = L->bary[(size_t)i * D1 + j] * alpha;
      acc += w * values[o];
    }
    out[i] = acc;
  }
}

/* Generic entry (any D): lattice of n points with features [n][D], K planes of n values filtered with it
 * (Permutohedral::init + compute per plane, un-normalised).  Returns M. */
int cosa_oracle_filter_features(const float *features, int n, const float *in, float *out, int K) {
  Lattice L;
  lattice_build_features(&L, features, n);
  float *values = (float *)malloc((size_t)(L.M + 2) * sizeof(float));
  float *new_values = (float *)malloc((size_t)(L.M + 2) * sizeof(float));
  for (int k = 0; k < K; k++) lattice_filter(&L, out + (size_t)k * n, in + (size_t)k * n, values, new_values);
  free(values); free(new_values);
  int M = L.M;
  lattice_free(&L);
  return M;
}
int cosa_oracle_lattice_dim(void) { return D; }

#if D == 5
/* bilateralfilter(): bilateralfilter.cpp:22-40.  Returns M. */
int cosa_oracle_bilateralfilter(const float *image, const float *in, float *out, int K, int H, int W,
                                float sigmargb, float sigmaxy) {
  Lattice L;
  lattice_build(&L, image, H, W, sigmargb, sigmaxy);
  const int n = H * W;
  float *values = (float *)malloc((size_t)(L.M + 2) * sizeof(float));
  float *new_values = (float *)malloc((size_t)(L.M + 2) * sizeof(float));
  for (int k = 0; k < K; k++) lattice_filter(&L, out + (size_t)k * n, in + (size_t)k * n, values, new_values);
  free(values); free(new_values);
  int M = L.M;
  lattice_free(&L);
  return M;
}

/* bilateralfilter_batch(): bilateralfilter.cpp:42-55 (OpenMP over images when built with -fopenmp). */
void cosa_oracle_bilateralfilter_batch(const float *images, const float *ins, float *outs, int N, int K, int H,
                                       int W, float sigmargb, float sigmaxy) {
  const size_t n = (size_t)H * W;
#pragma omp parallel for schedule(dynamic)
  for (int b = 0; b < N; b++)
    cosa_oracle_bilateralfilter(images + b * 3 * n, ins + b * K * n, outs + b * K * n, K, H, W, sigmargb, sigmaxy);
}

/* Debug/inspection entry used by the GPU parity tests: the embedding of one image.
 * offsets[n][6] (insertion-ordered ids), bary[n][6], vkeys[M_cap][5]; returns M (or -M if M > M_cap). */
int cosa_oracle_lattice_embed(const float *image, int H, int W, float sigmargb, float sigmaxy, int *offsets,
                              float *bary, int16_t *vkeys, int M_cap) {
  Lattice L;
  lattice_build(&L, image, H, W, sigmargb, sigmaxy);
  const size_t n = (size_t)H * W;
  memcpy(offsets, L.offset, n * D1 * sizeof(int));
  memcpy(bary, L.bary, n * D1 * sizeof(float));
  int M = L.M;
  if (M <= M_cap) memcpy(vkeys, L.vkeys, (size_t)M * D * sizeof(int16_t)); else M = -M;
  lattice_free(&L);
  return M;
}
#endif /* D == 5 */
