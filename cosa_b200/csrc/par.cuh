// Internal interface of the PAR kernels (par_kernels.cu), shared with the cam2mask pipeline.
#pragma once
#include "common.cuh"

namespace cosa {

constexpr int kMaxDil = 8;   // up to 64 neighbours

// The dilation list and the constant position term w2 * softmax(pos_aff) (PAR.py:51-62,82), evaluated on the host and
// passed to the kernels by value.  row_sum: sum over the neighbours of the weights of one pixel, 1 (softmax) + w2 (the
// position term), from the float constants the kernels use - a propagation step multiplies the channel sum of a stack
// by this factor (PAR.py:85-89; replicate padding keeps all 8*n_dil taps).  std_dilations: the list is the reference's
// {1,2,4,8,12,24} (PAR.py:94), which the compile-time tile kernels serve.
struct ParConst {
  int dil[kMaxDil];
  float pos_term[kMaxDil * 8];
  int n_dil, std_dilations;
  double row_sum;
};
int par_make_constants(const int *dilations, int n_dil, ParConst *pc);

// aff from imgs [B,3,h,w]: plain [B, 8*n_dil, h, w], or - tile_permuted, what the tile step kernels read (the
// reference dilation set only) - [B, 48, tiles, 32 x 32] with the pixels of a tile in aff_tile_offset order.
int par_launch_affinity(const ParConst &pc, const float *imgs, float *aff, int B, int h, int w, bool tile_permuted,
                        cudaStream_t stream);
// floats of an affinity buffer in either layout (planes padded to whole 32 x 32 tiles)
inline size_t par_affinity_floats(int B, int n_dil, int h, int w) {
  return (size_t)B * 8 * n_dil * (((h + 31) / 32) * 32) * (((w + 31) / 32) * 32);
}

// Row layout of a mask buffer [B, c_stride, h, pitch]: the w interior columns start at column `off`; `padn`
// replicated columns on either side make every neighbour load of the vectorised kernel an unclamped, 16-byte
// aligned access.  {w, 0, 0} is the plain NCHW layout.
struct MaskLayout {
  int pitch, off, padn;
};
inline MaskLayout plain_layout(int w) { return MaskLayout{w, 0, 0}; }
// Upper bound of the pitch of any padded layout the PAR path uses (pads of at most 24 columns).
inline int max_padded_pitch(int w) { return (32 + ((w + 31) & ~31) + 24 + 31) & ~31; }
// Padded layout for dilations up to max_dil (interior 128-byte aligned), or the plain one when w % 4 != 0.
MaskLayout padded_layout(int w, const int *dilations, int n_dil);
inline size_t layout_floats(const MaskLayout &l, int B, int c_stride, int h) {
  return (size_t)B * c_stride * h * l.pitch;
}

// Step counters of the chained step kernel: one int per 32 x 32 tile of every image (part of the caller's workspace).
inline size_t par_tile_flag_ints(int B, int h, int w) { return (size_t)B * ((h + 31) / 32) * ((w + 31) / 32); }

// num_iter propagation steps src0 -> ... -> final_dst.  The two scratch buffers use layout `lay`; src0 must use
// `lay` too (use par_launch_pack for a plain tensor); final_dst has its own layout (plain for user tensors).
// Live channels per image: nch_dev[b] when given, else nch_uniform.  tile_flags: par_tile_flag_ints(B, h, w) ints of
// scratch (nullptr: one launch per step).
int par_launch_iterations(const ParConst &pc, const float *aff, const float *src0, float *scratch_a, float *scratch_b,
                          MaskLayout lay, float *final_dst, MaskLayout lay_final, const int *nch_dev, int nch_uniform,
                          int c_stride, int B, int h, int w, int num_iter, int *tile_flags, cudaStream_t stream);

// Affinity (into aff [B, 8*n_dil, h, w]) + num_iter steps for the whole batch.  Buffer conventions as in
// par_launch_iterations.
int par_refine_batch(const ParConst &pc, const float *imgs, float *aff, const float *src0, float *scratch_a,
                     float *scratch_b, MaskLayout lay, float *final_dst, MaskLayout lay_final, const int *nch_dev,
                     int nch_uniform, int c_stride, int B, int h, int w, int num_iter, int *tile_flags,
                     cudaStream_t stream);

// plain [planes, h, w] -> layout `lay` (interior + replicated pads)
int par_launch_pack(const float *src, float *dst, MaskLayout lay, int planes, int h, int w, cudaStream_t stream);

}  // namespace cosa
