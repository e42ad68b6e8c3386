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
  unsigned* chunk_off;         // [gridDim] their exclusive prefix (written by the last residual CTA)
  unsigned* ticket;            // self-resetting arrival counter of the residual kernel
  int residual_order;          // VSF_OPT_RESIDUAL_ORDER
  int hold_on_empty;           // VSF_OPT_HOLD_THRESHOLD_ON_EMPTY
  // compaction outputs
  int* kept_left;              // [M] indices into the raw left frame
  int* kept_right;
  int* n_kept;                 // M (device: later kernels read it as a row count)
  int* n_kept_host;            // optional mirrors in mapped host memory (pipelined frame path)
  float* thresh_next_host;
  const uint32_t* desc_left;   // raw descriptors [n][words]
  const uint32_t* desc_right;
  uint32_t* desc_left_c;       // compacted descriptors [M][words] (a window ring slot)
  uint32_t* desc_right_c;
  float2* xy_left_c;
  float2* xy_right_c;
  int words;
};

cudaError_t launch_stereo_filter(const StereoArgs& a, int max_matches, bool with_threshold, cudaStream_t stream);

// N1: cv::undistortPoints as Frontend::UndistortFeaturePoints calls it (src/slam_frontend.cc:323-351:
// R empty, P = K_left): normalise with K, five fixed-point iterations of the radial / tangential
// model in double (OpenCV's default criteria), re-project with K, store float.
struct UndistortArgs {
  double fx, fy, cx, cy, k1, k2, p1, p2, k3;
};

inline UndistortArgs make_undistort_args(const float* K9, const float* dist5) {
  UndistortArgs a;
  a.fx = double(K9[0]); a.fy = double(K9[4]); a.cx = double(K9[2]); a.cy = double(K9[5]);
  a.k1 = double(dist5[0]); a.k2 = double(dist5[1]); a.p1 = double(dist5[2]); a.p2 = double(dist5[3]);
  a.k3 = double(dist5[4]);
  return a;
}

__device__ __forceinline__ float2 undistort_one(float2 p, const UndistortArgs& a) {
  const double x0 = (double(p.x) - a.cx) / a.fx, y0 = (double(p.y) - a.cy) / a.fy;
  double x = x0, y = y0;
#pragma unroll 1
  for (int it = 0; it < 5; ++it) {
    const double r2 = x * x + y * y;
    const double icdist = 1.0 / (1.0 + ((a.k3 * r2 + a.k2) * r2 + a.k1) * r2);
    if (icdist < 0) {      // OpenCV restores the starting point for this iteration
      x = x0;
      y = y0;
      continue;
    }
    const double dx = 2 * a.p1 * x * y + a.p2 * (r2 + 2 * x * x);
    const double dy = a.p1 * (r2 + 2 * y * y) + 2 * a.p2 * x * y;
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  return make_float2(float(x * a.fx + a.cx), float(y * a.fy + a.cy));
}

// Extras the fused frame path folds into the triangulation launch (both optional):
//  * the undistorted pixels of the M compacted left keypoints (N1) - thread i handles keypoint i;
//  * the sequential threshold sum of the stereo stage, in one extra CTA, where nothing on the
//    frame's critical path waits for it.
struct TriExtras {
  int do_undistort;
  UndistortArgs und;
  const int* n_kept;           // M (device)
  float2* xy_undist;           // [M]
  int do_threshold;
  StereoArgs stereo;           // resid / n_matches / thresh_cur / thresh_next / hold_on_empty
};

}  // namespace vsf
