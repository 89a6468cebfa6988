// Internal interface of the PAR kernels (par_kernels.cu), shared with the cam2mask pipeline.
#pragma once
#include "common.cuh"

namespace cosa {

constexpr int kMaxDil = 8;   // up to 64 neighbours

// Uploads the dilation list and the constant position term (PAR.py:51-62,82) for subsequent launches.
int par_upload_constants(const int *dilations, int n_dil, cudaStream_t stream);

// aff [B, 8*n_dil, h, w] from imgs [B,3,h,w].
int par_launch_affinity(const float *imgs, float *aff, int B, int h, int w, int n_dil, cudaStream_t stream);

// num_iter propagation steps src0 -> ... -> final_dst through the two scratch buffers (all distinct,
// [B, c_stride, h, w]).  Live channels per image: nch_dev[b] when given, else nch_uniform.
int par_launch_iterations(const float *aff, const float *src0, float *scratch_a, float *scratch_b, float *final_dst,
                          const int *nch_dev, int nch_uniform, int c_stride, int B, int h, int w, int n_dil,
                          int num_iter, cudaStream_t stream);

}  // namespace cosa
