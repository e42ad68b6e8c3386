// Kernel 2: stereo epipolar filter + compaction (a5) and linear triangulation (a6).
//
//   stereo_residual_kernel + stereo_compact_kernel   <- Frontend::RemoveAmbigStereo
//                                                       src/slam_frontend.cc:353-398
//   triangulate_*_kernel                             <- cv::triangulatePoints as called
//                                                       from Frontend::Calculate3DPoints
//                                                       src/slam_frontend.cc:136-165
#include "vsf_device.cuh"
#include "stereo_args.cuh"

#include <algorithm>
#include <cstring>

namespace vsf {

constexpr int kStereoThreads = 256;

// `(left_ph^T * F * right_ph).norm()` (src/slam_frontend.cc:380-381) in float32; the *_rn
// intrinsics forbid FMA contraction so the result is bit-identical to the oracle's float32
// restatement (oracle/restate.py: epipolar_residual).  Both products are 3-term dot products
// whose summation order is Eigen's, and Eigen is not part of the reference checkout:
// tree = false: (c0 + c1) + c2, Eigen 3.2's unrolled coefficient product;
// tree = true:  c0 + (c1 + c2), Eigen 3.3's redux_novec_unroller (splits the range in halves).
__device__ __forceinline__ float dot3_ordered(float a0, float a1, float a2, bool tree) {
  return tree ? __fadd_rn(a0, __fadd_rn(a1, a2)) : __fadd_rn(__fadd_rn(a0, a1), a2);
}

__device__ __forceinline__ float epipolar_residual(float2 l, float2 r, const float* F, bool tree) {
  float v[3];
#pragma unroll
  for (int j = 0; j < 3; ++j)
    v[j] = dot3_ordered(__fmul_rn(l.x, F[j]), __fmul_rn(l.y, F[3 + j]), __fmul_rn(1.0f, F[6 + j]), tree);
  const float c = dot3_ordered(__fmul_rn(v[0], r.x), __fmul_rn(v[1], r.y), __fmul_rn(v[2], 1.0f), tree);
  return __fsqrt_rn(__fmul_rn(c, c));
}

// One CTA per 256-match chunk: residuals + survivors of the chunk.  The last CTA to finish (a
// self-resetting ticket) turns the per-chunk counts into exclusive offsets and the total M, so
// the compaction kernel starts from finished offsets.
__global__ void __launch_bounds__(kStereoThreads)
stereo_residual_kernel(const __grid_constant__ StereoArgs a) {
  __shared__ unsigned s_last;
  pdl_wait();
  pdl_launch_dependents();
  const int n = *a.n_matches;
  const int m = blockIdx.x * kStereoThreads + threadIdx.x;
  bool keep = false;
  if (m < n) {
    const vsf_dmatch dm = a.matches[m];
    const float c = epipolar_residual(a.xy_left[dm.queryIdx], a.xy_right[dm.trainIdx], a.F, a.residual_order != 0);
    a.resid[m] = c;
    keep = (c <= *a.thresh_cur);   // false for NaN, as in the reference
  }
  const int cnt = __syncthreads_count(keep);
  if (threadIdx.x == 0) {
    a.chunk_keep[blockIdx.x] = unsigned(cnt);
    __threadfence();
    const unsigned t = atomicAdd(a.ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < 32) {   // exclusive scan of the chunk counts by one warp
    const int lane = threadIdx.x;
    unsigned run = 0;
    for (int c0 = 0; c0 < int(gridDim.x); c0 += 32) {
      const int c = c0 + lane;
      const unsigned v = (c < int(gridDim.x)) ? __ldcg(a.chunk_keep + c) : 0u;
      unsigned incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      if (c < int(gridDim.x)) a.chunk_off[c] = run + incl - v;
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
      *a.n_kept = int(run);
      if (a.n_kept_host) *a.n_kept_host = int(run);
      *a.ticket = 0u;
    }
  }
}

// Mean of the residuals, accumulated sequentially in float32 in match order - the order the
// reference adds them in (src/slam_frontend.cc:382) - so the next threshold is bit-identical.
// The whole CTA stages the residuals through shared memory (coalesced loads); one thread then
// walks them: the dependent FADD chain (4 cycles per add) is all that is left.
constexpr int kThreshStage = 4096;
__device__ void stereo_threshold_cta(const StereoArgs& a, float* s_buf) {
  const int n = *a.n_matches;
  float avg = 0.0f;
  for (int base = 0; base < n; base += kThreshStage) {
    const int len = min(kThreshStage, n - base);
    __syncthreads();
    for (int i = threadIdx.x; i < len; i += blockDim.x) s_buf[i] = a.resid[base + i];
    __syncthreads();
    if (threadIdx.x == 0) {
      int i = 0;
      for (; i + 4 <= len; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(s_buf + i);
        avg = __fadd_rn(avg, v.x);
        avg = __fadd_rn(avg, v.y);
        avg = __fadd_rn(avg, v.z);
        avg = __fadd_rn(avg, v.w);
      }
      for (; i < len; ++i) avg = __fadd_rn(avg, s_buf[i]);
    }
  }
  if (threadIdx.x == 0) {
    // avg_constraint / stereo_matches.size() + padding (src/slam_frontend.cc:392-394);
    // 0/0 = NaN when there were no matches, like the reference - unless the caller asked to
    // keep the threshold through empty frames (VSF_OPT_HOLD_THRESHOLD_ON_EMPTY)
    float next = __fadd_rn(__fdiv_rn(avg, float(n)), 2.0f);
    if (n == 0 && a.hold_on_empty) next = *a.thresh_cur;
    *a.thresh_next = next;
    if (a.thresh_next_host) *a.thresh_next_host = next;
  }
}

__global__ void __launch_bounds__(kStereoThreads) stereo_threshold_kernel(const __grid_constant__ StereoArgs a) {
  __shared__ __align__(16) float s_buf[kThreshStage];
  pdl_wait();
  pdl_launch_dependents();
  stereo_threshold_cta(a, s_buf);
}

// One CTA per 256-match chunk: compacts the chunk behind the survivors of the earlier chunks
// and gathers keypoint pixels and descriptor rows of both frames.
__global__ void __launch_bounds__(kStereoThreads)
stereo_compact_kernel(const __grid_constant__ StereoArgs a) {
  __shared__ unsigned s_warp[kStereoThreads / 32];
  pdl_wait();
  pdl_launch_dependents();
  const int n = *a.n_matches;
  const int tid = threadIdx.x;
  const int m = blockIdx.x * kStereoThreads + tid;
  bool keep = false;
  vsf_dmatch dm = {0, 0, 0, 0.f};
  if (m < n) {
    dm = a.matches[m];
    keep = (a.resid[m] <= *a.thresh_cur);
  }
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  const int lane = tid & 31, warp = tid >> 5;
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  unsigned off = a.chunk_off[blockIdx.x];
  for (int w = 0; w < warp; ++w) off += s_warp[w];
  if (keep) {
    const unsigned dst = off + __popc(bal & ((1u << lane) - 1u));
    a.kept_left[dst] = dm.queryIdx;
    a.kept_right[dst] = dm.trainIdx;
    a.xy_left_c[dst] = a.xy_left[dm.queryIdx];
    a.xy_right_c[dst] = a.xy_right[dm.trainIdx];
    const int v4 = a.words / 4;
    const uint4* sl = reinterpret_cast<const uint4*>(a.desc_left + size_t(dm.queryIdx) * a.words);
    const uint4* sr = reinterpret_cast<const uint4*>(a.desc_right + size_t(dm.trainIdx) * a.words);
    uint4* dl = reinterpret_cast<uint4*>(a.desc_left_c + size_t(dst) * a.words);
    uint4* dr = reinterpret_cast<uint4*>(a.desc_right_c + size_t(dst) * a.words);
    for (int v = 0; v < v4; ++v) {
      dl[v] = __ldg(sl + v);
      dr[v] = __ldg(sr + v);
    }
  }
}

template <typename K>
static cudaError_t launch_stereo_pdl(K kernel, int blocks, cudaStream_t stream, const StereoArgs& a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(kStereoThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, a);
}

// with_threshold = false: the caller runs the threshold sum elsewhere (the fused frame path puts
// it into an extra CTA of the triangulation kernel, where nothing waits for it)
cudaError_t launch_stereo_filter(const StereoArgs& a, int max_matches, bool with_threshold, cudaStream_t stream) {
  const int chunks = (max_matches + kStereoThreads - 1) / kStereoThreads;
  if (chunks <= 0) return cudaSuccess;
  cudaError_t e = launch_stereo_pdl(stereo_residual_kernel, chunks, stream, a);
  if (e == cudaSuccess) e = launch_stereo_pdl(stereo_compact_kernel, chunks, stream, a);
  if (e == cudaSuccess && with_threshold) e = launch_stereo_pdl(stereo_threshold_kernel, 1, stream, a);
  return e;
}

// ---------------------------------------------------------------------------------------------
// Linear (DLT) triangulation of one correspondence: the 4x4 system with two rows
// per view, x*P[2,:]-P[0,:] and y*P[2,:]-P[1,:], built in double from float32
// inputs; its right singular vector of the smallest singular value is found with
// a one-sided (Hestenes) Jacobi SVD in double — the same family of solver
// OpenCV's cv::SVD uses for small matrices — and stored as float32.
struct Proj {
  float P1[12];
  float P2[12];
};

__device__ __forceinline__ float4 triangulate_one(const Proj& pj, float2 p1, float2 p2) {
  double A[4][4];  // A[row][col]
  double V[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    A[0][k] = double(p1.x) * double(pj.P1[8 + k]) - double(pj.P1[k]);
    A[1][k] = double(p1.y) * double(pj.P1[8 + k]) - double(pj.P1[4 + k]);
    A[2][k] = double(p2.x) * double(pj.P2[8 + k]) - double(pj.P2[k]);
    A[3][k] = double(p2.y) * double(pj.P2[8 + k]) - double(pj.P2[4 + k]);
#pragma unroll
    for (int j = 0; j < 4; ++j) V[k][j] = (k == j) ? 1.0 : 0.0;
  }
  const double eps = 2.220446049250313e-16 * 8.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          alpha += A[r][p] * A[r][p];
          beta += A[r][q] * A[r][q];
          gamma += A[r][p] * A[r][q];
        }
        if (fabs(gamma) <= eps * sqrt(alpha * beta) || gamma == 0.0) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t);
        const double s = c * t;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const double ap = A[r][p], aq = A[r][q];
          A[r][p] = c * ap - s * aq;
          A[r][q] = s * ap + c * aq;
          const double vp = V[r][p], vq = V[r][q];
          V[r][p] = c * vp - s * vq;
          V[r][q] = s * vp + c * vq;
        }
      }
    }
    if (!rotated) break;
  }
  // column of smallest norm <-> smallest singular value
  int jmin = 0;
  double best = 1e300;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double nrm = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) nrm += A[r][j] * A[r][j];
    if (nrm < best) {
      best = nrm;
      jmin = j;
    }
  }
  double x[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    x[r] = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) x[r] = (j == jmin) ? V[r][j] : x[r];
  }
  return make_float4(float(x[0]), float(x[1]), float(x[2]), float(x[3]));
}

// Explicit point pairs (the stateless vsf_triangulate entry point).  X4 is 4 x n row-major.
__global__ void triangulate_pairs_kernel(const __grid_constant__ Proj pj, const float2* x1,
                                         const float2* x2, int n, float* X4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 X = triangulate_one(pj, x1[i], x2[i]);
  X4[i] = X.x;
  X4[n + i] = X.y;
  X4[2 * size_t(n) + i] = X.z;
  X4[3 * size_t(n) + i] = X.w;
}

// R'->L' matches (query = compacted right frame, train = compacted left frame):
// left_pt = left.keypoints_[trainIdx].pt, right_pt = right.keypoints_[queryIdx].pt
// (src/slam_frontend.cc:137-139).  Output in match order, one float4 per match.  One warp per
// CTA: a point is a long dependent FP64 chain in one thread, so the launch is spread over as
// many SMs as there are warps.  The last CTA (when ex.do_threshold) is the stereo stage's
// sequential threshold sum; thread i also undistorts compacted left keypoint i (ex.do_undistort).
constexpr int kTriThreads = 32;
__global__ void __launch_bounds__(kTriThreads)
triangulate_matches_kernel(const __grid_constant__ Proj pj, const vsf_dmatch* matches, const int* n_matches,
                           const float2* xy_left_c, const float2* xy_right_c, float4* X4,
                           const __grid_constant__ TriExtras ex) {
  __shared__ __align__(16) float s_buf[kThreshStage];
  pdl_wait();
  pdl_launch_dependents();
  if (ex.do_threshold && blockIdx.x == gridDim.x - 1) {
    stereo_threshold_cta(ex.stereo, s_buf);
    return;
  }
  const int i = blockIdx.x * kTriThreads + threadIdx.x;
  if (ex.do_undistort && i < *ex.n_kept) ex.xy_undist[i] = undistort_one(xy_left_c[i], ex.und);
  if (i >= *n_matches) return;
  const vsf_dmatch dm = matches[i];
  X4[i] = triangulate_one(pj, xy_left_c[dm.trainIdx], xy_right_c[dm.queryIdx]);
}

cudaError_t launch_triangulate_pairs(const float* P1, const float* P2, const float2* x1,
                                     const float2* x2, int n, float* X4, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  Proj pj;
  for (int i = 0; i < 12; ++i) {
    pj.P1[i] = P1[i];
    pj.P2[i] = P2[i];
  }
  triangulate_pairs_kernel<<<(n + 127) / 128, 128, 0, stream>>>(pj, x1, x2, n, X4);
  return cudaGetLastError();
}

cudaError_t launch_triangulate_matches(const float* P1, const float* P2, const vsf_dmatch* matches,
                                       const int* n_matches, int max_matches,
                                       const float2* xy_left_c, const float2* xy_right_c,
                                       float4* X4, const TriExtras* extras, cudaStream_t stream) {
  TriExtras ex;
  std::memset(&ex, 0, sizeof(ex));
  if (extras) ex = *extras;
  int blocks = (std::max(max_matches, 0) + kTriThreads - 1) / kTriThreads + (ex.do_threshold ? 1 : 0);
  if (blocks <= 0) return cudaSuccess;
  Proj pj;
  for (int i = 0; i < 12; ++i) {
    pj.P1[i] = P1[i];
    pj.P2[i] = P2[i];
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(kTriThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, triangulate_matches_kernel, pj, matches, n_matches, xy_left_c, xy_right_c, X4, ex);
}

}  // namespace vsf
