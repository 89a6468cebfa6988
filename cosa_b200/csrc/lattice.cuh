// Permutohedral lattice on the GPU (d = 5) — internal interface shared by lattice_kernels.cu and
// energy_kernels.cu.  Reference: utils/bilateralfilter/permutohedral.cpp:115-297 (init), :507-571 (compute).
//
// The whole batch is ONE lattice: the image index is part of the vertex key, so images never mix (the blur
// changes only the five feature coordinates) and every stage is a single launch over all images.
//
// Key packing (64 bit).  A lattice point has coordinates key[i] = 6*q[i] + r with one common residue
// r in [0,5] (rem0 is a multiple of d+1 = 6 and canonical[r][.] is r or r-6, permutohedral.cpp:148-153,247):
//     bits  0..54 : q[0..4] + 1024, 11 bits each   (|key[i]| <= 6143; the reference stores `short`, +-32767)
//     bits 55..57 : r
//     bits 58..63 : image index within the chunk (chunks of 64 images; EMPTY is all-ones, i.e. r = 7: never a key)
// A coordinate outside the range raises the sticky error flag (counters[1] bit 0): the slice then writes NaN and the
// energy loss is NaN, and cosa_bilateral_stats / the host form return COSA_E_KEYRANGE - never an aliased vertex.
//
// Tile-local vertex lists.  The image is cut into 32 x 8 pixel tiles.  The 1536 (pixel, vertex) pairs of a tile
// touch only ~22 % as many distinct vertices (the lattice cells are sigma_xy pixels wide), so the build
// de-duplicates a tile's keys ONCE in shared memory and writes
//     tile_info[t]          (first list entry, U | pairs << 16: distinct vertices and pairs of the tile)
//     tkeys / tvid          per list entry: packed key, vertex row (id + 1)
//     plist[t][1536 + 32]   the tile's pairs bucketed by list entry: (pixel in tile << 3 | r | first-of-vertex << 11),
//                           then for each block of 48 pairs the list entry its first pair belongs to
//     lidx[r][P], bary[r][P] per pixel: 16-bit index into the tile's list and the barycentric weight
// Only the distinct keys of a tile go to the global hash table (4-5x fewer probes); the splat reduces a tile's pairs
// per vertex from shared memory without hashing, and the slice stages the tile's vertex rows once.
#pragma once
#include "common.cuh"

namespace cosa {

constexpr int kLatD = 5;                            // the path's lattice: (x, y, R, G, B)
constexpr int kMaxLatD = 5;                         // kernels are instantiated for D = 5 and D = 2 (the dense CRF's
                                                    // spatial kernel, utils/seg_helper.py:961-996)
constexpr int kQBits = 11;
constexpr int kQBias = 1 << (kQBits - 1);          // 1024
constexpr int kKeyRShift = kQBits * kMaxLatD;       // 55: residue field (the same place for every D)
constexpr int kKeyBShift = kKeyRShift + 3;          // 58: image index
constexpr int kMaxImagesPerLattice = 64;
constexpr unsigned long long kEmptyKey = ~0ULL;

constexpr int kTileW = 32, kTileH = 8;
constexpr int kTilePix = kTileW * kTileH;           // 256 = threads per tile CTA
constexpr int tile_pairs(int d) { return (d + 1) * kTilePix; }      // 1536 at D = 5
constexpr int pair_block(int d) { return tile_pairs(d) / 32; }      // 48: pairs per quarter-warp of the splat
constexpr int list_stride(int d) { return tile_pairs(d) + 32; }     // u16 per tile: the pair list + the first list
                                                                    // entry of each block

// counters[]: 0: M   1: error flags (1 = key range, 2 = list / vertex capacity)   2: max probe length
//             3: table capacity in use   4: M of the earlier chunks of this call   5: T = list entries
//             6: 1 while val0 is known to be all zero (set by the build, cleared by the splat)
struct LatticeBufs {
  unsigned long long *table_keys;   // [cap]   packed key or kEmptyKey (the build uses the first 2^k >= 2T slots)
  int *table_ids;                   // [cap]   vertex id + 1 of an occupied slot
  unsigned long long *vkeys;        // [m_cap] packed key of vertex id
  int *counters;                    // [8]
  int2 *tile_info;                  // [tiles] (first list entry, U | pairs << 16)
  unsigned long long *tkeys;        // [t_cap] packed key of a list entry
  int *tvid;                        // [t_cap] table slot during the build, then vertex id + 1 (row of val0 / val1)
  unsigned short *plist;            // [tiles][list_stride(d)]
  unsigned short *lidx;             // [d + 1][P]  index into the tile's list
  float *bary;                      // [d + 1][P]
  int2 *nbr;                        // [d + 1][m_cap]  (n1, n2) as vertex id + 1, 0 = absent
  float *val0, *val1;               // [m_cap + 1][Kp]   row 0 is the all-zero "absent" row
  long long P;                      // N * H * W
  long long m_cap;                  // upper bound on the number of vertices
  long long t_cap;                  // upper bound on the number of list entries
  unsigned long long cap_mask;      // allocated table capacity - 1 (power of two)
  int tiles_x, tiles_y;             // tiles per image
  int Kp;                           // channels rounded up to a multiple of 4
  int d;                            // lattice dimension: 5 (bilateral) or 2 (spatial)
};

// vpp: vertex budget in vertices per pixel.  0 = the worst case (d + 1 new vertices per pixel: every size bound is then
// a guarantee).  A positive value sizes the vertex arrays (keys, neighbour table, the two value buffers - the bulk of
// the workspace) for vpp * N * H * W vertices and the per-tile lists for twice as many entries; natural images need
// 0.2 - 0.6 (the synthetic VOC batch: 0.50), uniform noise 2.4.  A lattice that outgrows its budget raises the sticky
// error flag 2: outputs and loss become NaN and cosa_bilateral_stats returns COSA_E_WORKSPACE - never a silent overrun.
size_t lattice_ws_bytes(int N, int K, int H, int W, int d = kLatD, float vpp = 0.0f);
// Carves `ws` (>= lattice_ws_bytes for the same arguments) into the buffers above.
void lattice_carve(void *ws, int N, int K, int H, int W, LatticeBufs *L, int d = kLatD, float vpp = 0.0f);

// Build the lattice (dimension L.d) of N (<= kMaxImagesPerLattice) planar RGB images [N,3,H,W] (d = 2: the images are
// not read, the features are the pixel coordinates alone).
// first_chunk: this is the first lattice of a call (resets the per-call counters: total vertices, error flag).
int lattice_build(const LatticeBufs &L, const float *images, int N, int H, int W, float sigmargb, float sigmaxy,
                  bool first_chunk, cudaStream_t stream);
// Images per lattice for a batch of N (<= kMaxImagesPerLattice).
int lattice_chunk_images(int N, int K, int H, int W);
// values <- splat(ins [N,K,H,W]); six blur passes.  The blurred values end up in L.val0.
int lattice_splat_blur(const LatticeBufs &L, const float *ins, int N, int K, int H, int W, cudaStream_t stream);
// outs [N,K,H,W] <- slice.  With gate != nullptr the dense-CRF epilogue is fused: outs <- slice * gate and
// *loss_acc += sum(ins * outs)   (utils/seg_helper.py:888-890).
int lattice_slice(const LatticeBufs &L, const float *ins, const float *gate, double *loss_acc, float *outs, int N,
                  int K, int H, int W, cudaStream_t stream);

}  // namespace cosa
