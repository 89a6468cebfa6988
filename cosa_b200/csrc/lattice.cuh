// Permutohedral lattice on the GPU (d = 5) — internal interface shared by lattice_kernels.cu and
// energy_kernels.cu.  Reference: utils/bilateralfilter/permutohedral.cpp:115-297 (init), :507-571 (compute).
//
// The whole batch is ONE lattice: the image index is part of the vertex key, so images never mix (the blur
// changes only the five feature coordinates) and every stage is a single launch over all images.
//
// Key packing (64 bit).  A lattice point has coordinates key[i] = 6*q[i] + r with one common residue
// r in [0,5] (rem0 is a multiple of d+1 = 6 and canonical[r][.] is r or r-6, permutohedral.cpp:148-153,247):
//     bits  0..54 : q[0..4] + 1024, 11 bits each   (|key[i]| <= 6143; the reference stores `short`)
//     bits 55..57 : r
//     bits 58..63 : image index within the chunk (chunks of 64 images; EMPTY is all-ones, i.e. r = 7: never a key)
// A coordinate outside the range raises the error flag (COSA_E_KEYRANGE) instead of aliasing.
#pragma once
#include "common.cuh"

namespace cosa {

constexpr int kLatD = 5;
constexpr int kQBits = 11;
constexpr int kQBias = 1 << (kQBits - 1);          // 1024
constexpr int kMaxImagesPerLattice = 64;
constexpr unsigned long long kEmptyKey = ~0ULL;

struct LatticeBufs {
  unsigned long long *table_keys;   // [cap]   packed key or kEmptyKey
  int *table_ids;                   // [cap]   vertex id + 1 of an occupied slot
  unsigned long long *vkeys;        // [m_cap] packed key of vertex id
  int *counters;                    // [8]     0: M   1: key-range error   2: max probe length   3: table capacity   4: M of earlier chunks
  int *offsets;                     // [6][P]  table slot during the build, then vertex id + 1
  float *bary;                      // [6][P]
  int2 *nbr;                        // [6][m_cap]  (n1, n2) as vertex id + 1, 0 = absent
  float *val0, *val1;               // [m_cap + 1][Kp]   row 0 is the all-zero "absent" row
  long long P;                      // N * H * W
  long long m_cap;                  // upper bound on the number of vertices
  unsigned long long cap_mask;      // table capacity - 1 (power of two)
  int Kp;                           // channels rounded up to a multiple of 4
};

size_t lattice_ws_bytes(int N, int K, int H, int W);
// Carves `ws` (>= lattice_ws_bytes) into the buffers above.
void lattice_carve(void *ws, int N, int K, int H, int W, LatticeBufs *L);

// Build the lattice of N (<= kMaxImagesPerLattice) planar RGB images [N,3,H,W].
// first_chunk: this is the first lattice of a call (resets the per-call counters: total vertices, error flag).
int lattice_build(const LatticeBufs &L, const float *images, int N, int H, int W, float sigmargb, float sigmaxy,
                  bool first_chunk, cudaStream_t stream);
// Images per lattice for a batch of N (<= kMaxImagesPerLattice, sized for L2; see lattice_kernels.cu).
int lattice_chunk_images(int N, int K, int H, int W);
// values <- splat(ins [N,K,H,W]); six blur passes.  The blurred values end up in L.val0.
int lattice_splat_blur(const LatticeBufs &L, const float *ins, int N, int K, int H, int W, cudaStream_t stream);
// outs [N,K,H,W] <- slice.  With gate != nullptr the dense-CRF epilogue is fused: outs <- slice * gate and
// *loss_acc += sum(ins * outs)   (utils/seg_helper.py:888-890).
int lattice_slice(const LatticeBufs &L, const float *ins, const float *gate, double *loss_acc, float *outs, int N,
                  int K, int H, int W, cudaStream_t stream);

}  // namespace cosa
