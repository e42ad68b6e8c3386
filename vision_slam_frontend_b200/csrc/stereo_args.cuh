// Argument block shared by the stereo kernels (stereo_kernels.cu) and the host
// launch sequences (vsf_api.cu).
#pragma once

#include "vsf_device.cuh"

namespace vsf {

struct StereoArgs {
  const vsf_dmatch* matches;   // L->R matches after the ratio test, query (left) order
  const int* n_matches;        // device count
  const float2* xy_left;       // raw frame pixels
  const float2* xy_right;
  float F[9];                  // row-major, x_left^T F x_right
  const float* thresh_cur;     // stereo_ambig_constraint used for this frame
  float* thresh_next;          // mean(residual) + 2, used for the next frame
  float* resid;                // [n_matches]
  unsigned* chunk_keep;        // [gridDim] survivors per 256-match chunk
  // compaction outputs
  int* kept_left;              // [M] indices into the raw left frame
  int* kept_right;
  int* n_kept;                 // M
  const uint32_t* desc_left;   // raw descriptors [n][words]
  const uint32_t* desc_right;
  uint32_t* desc_left_c;       // compacted descriptors [M][words] (a window ring slot)
  uint32_t* desc_right_c;
  float2* xy_left_c;
  float2* xy_right_c;
  int words;
};

cudaError_t launch_stereo_filter(const StereoArgs& a, int max_matches, cudaStream_t stream);

}  // namespace vsf
