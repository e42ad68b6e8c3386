/* vsf.h — C ABI of libvsf_cuda.so: the B200 (sm_100a) replacement for the
 * descriptor-matching / stereo / triangulation hot path of
 * ut-amrl/vision_slam_frontend.
 *
 * The reference has no plugin or operator registry.  The seam is the set of
 * private Frontend helpers and the two OpenCV calls they make; each entry point
 * below names the reference interface it replaces (paths relative to the
 * reference checkout).  All signatures are plain C: pointers, sizes, scalars.
 * Host-pointer entry points never keep a caller pointer past return.  One
 * vsf_ctx = one CUDA device + one stream; calls on a ctx must be serialised by
 * the caller, different ctxs may be used concurrently (one per GPU / rank).
 * There is no CPU fallback: every entry point fails with VSF_ERR_CUDA when the
 * device or the kernels are unavailable.
 *
 * Return value: 0 (VSF_OK) or a VSF_ERR_* code; vsf_last_error(ctx) gives a
 * human-readable message for the last failure on that ctx.  The library never
 * aborts (the reference's CHECK/exit behaviour is re-created, by choice, in the
 * C++ Frontend mirror above this ABI).
 */
#ifndef VSF_H_
#define VSF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSF_VERSION_STRING "0.1.0"

enum {
  VSF_OK = 0,
  VSF_ERR_BAD_ARG = 1,   /* null pointer, negative size, unsupported width */
  VSF_ERR_CAPACITY = 2,  /* more rows / frames than the ctx was created for */
  VSF_ERR_CUDA = 3,      /* CUDA runtime error (message has the detail)     */
  VSF_ERR_STATE = 4      /* call not valid in the current ctx state         */
};

/* Layout-identical to cv::DMatch (16 bytes): what Frontend::GetMatches returns
 * (src/slam_frontend.cc:521-538).  distance holds an exact small integer. */
typedef struct vsf_dmatch {
  int32_t queryIdx;
  int32_t trainIdx;
  int32_t imgIdx;
  float distance;
} vsf_dmatch;

/* Layout-identical to cv::KeyPoint (28 bytes): Frame::keypoints_
 * (src/slam_frontend.h:107).  Only x,y are read on the device. */
typedef struct vsf_keypoint {
  float x, y;
  float size, angle, response;
  int32_t octave, class_id;
} vsf_keypoint;

/* Layout-identical to slam_types::FeatureMatch (src/slam_types.h:77-89). */
typedef struct vsf_feature_match {
  uint64_t feature_idx_initial;
  uint64_t feature_idx_current;
} vsf_feature_match;

typedef struct vsf_ctx vsf_ctx;

/* ------------------------------------------------------------------ context */

/* device: CUDA ordinal.  max_features: largest number of descriptor rows in
 * one frame.  desc_bytes: descriptor width as stored by the caller (1..64;
 * 32 = ORB, 61 = AKAZE, 64 = BRISK/FREAK); rows are zero-padded to 32 or 64
 * bytes on the device, which leaves Hamming distances unchanged.  window:
 * FrontendConfig::frame_life_ (src/slam_frontend.cc:556), the number of past
 * frames kept resident for vsf_window_* / vsf_observe_features. */
int vsf_create(int device, int max_features, int desc_bytes, int window,
               vsf_ctx** out);
void vsf_destroy(vsf_ctx* ctx);
const char* vsf_last_error(const vsf_ctx* ctx);
const char* vsf_version(void);

/* Run the ctx's work on an existing CUDA stream (cudaStream_t passed as
 * void*).  NULL restores the ctx's own non-blocking stream — note that the
 * legacy default stream has handle 0 and therefore cannot be selected; pass a
 * created stream.  Lets a caller time or overlap the asynchronous *_device
 * entry points with its own events. */
int vsf_set_stream(vsf_ctx* ctx, void* cuda_stream);
int vsf_synchronize(vsf_ctx* ctx);

/* Kernel tuning knobs; (-1, 0, 0, -1) restores the library defaults.
 * popc_mode: 0 = 8 POPC per 256-bit comparison (the naive count the roofline is
 * quoted against), 2 / 3 = carry-save adder trees with 5 / 4 POPC, -1 = default.
 * train_split: POPC engine: number of train-dimension splits per problem; tensor
 * engine: pieces a query block's train tiles are cut into (granularity of the
 * per-CTA work ranges); 0 = automatic.
 * queries_per_thread: 1, 2 or 4, 0 = automatic.  variant: 0..3 kernel structure
 * (producer warp / unroll, see csrc/knn2_kernel.cu), -1 = default.
 * Every setting produces bit-identical results. */
int vsf_set_tuning(vsf_ctx* ctx, int popc_mode, int train_split,
                   int queries_per_thread, int variant);

/* Which hardware pipe computes the distance matrix of the kNN kernel.
 * 0 = automatic (tensor cores for batches large enough to fill the GPU, POPC pipe
 * otherwise), 1 = POPC pipe (XOR + POPC, csrc/knn2_kernel.cu), 2 = tensor cores,
 * int8 operands (tcgen05.mma kind::i8 over +-1 expanded descriptor bits,
 * csrc/knn2_tc_kernel.cu), 3 = tensor cores, e4m3 operands (kind::f8f6f4).
 * Engine 2 exists for every descriptor width (csrc/knn2_tc64_kernel.cu for 33..64
 * bytes), engine 3 for desc_bytes <= 32 only.  Every engine produces bit-identical
 * results.  flags: pass 0; 128 keeps 33..64-byte rows on the single-CTA tensor kernel instead of
 * CTA pairs (A/B timing); 64 keeps the device sort of the pipelined frame stream (sort_mode 0 / 2)
 * on the main stream instead of a side stream with reserved SMs (A/B timing); 512 runs the
 * tensor engine's refine and compaction as two kernels instead of one (A/B timing); 256 makes
 * vsf_window_match_block_device launch pose by pose instead of in groups, 4096 makes the
 * distance kernels of the host-buffer frame stream let their finish kernel launch at their start
 * instead of near their end (A/B timing; 1024 and 2048 are accepted and do nothing any more); 16 / 32 record the per-CTA / per-kernel
 * timelines read by vsf_debug_tc_trace / vsf_debug_kernel_trace (16, and the
 * timing-only flags 2 / 4, exist only in libraries built with VSF_TC_TRACE /
 * VSF_TC_BRINGUP, see vision_slam_frontend_b200/build.py). */
int vsf_set_engine(vsf_ctx* ctx, int engine, int flags);
/* Engine (1..3) used by the most recent kNN launch of this ctx. */
int vsf_last_engine(const vsf_ctx* ctx);

/* Per-kernel timing of the kNN launch sequence, for roofline accounting.  While enabled,
 * CUDA events are recorded on the ctx stream around each kernel of every kNN launch (this
 * removes the programmatic overlap between them, so leave it off for throughput runs).
 * vsf_last_kernel_times waits for the most recent launch and returns 4 durations in ms:
 * [0] train expansion, [1] the distance/selection kernel (knn2_tc_kernel or knn2_kernel),
 * [2] refine, [3] ordered compaction ([0], [2], [3] are 0 for the POPC engine, whose single
 * kernel does everything). */
int vsf_set_profile(vsf_ctx* ctx, int enabled);
int vsf_last_kernel_times(vsf_ctx* ctx, float* ms4);

/* ---------------------------------------------- a1: BFMatcher::knnMatch k=2 */

/* Replaces `matcher_->knnMatch(query.descriptors_, train.descriptors_,
 * matches, 2)` with matcher_ = cv::BFMatcher(NORM_HAMMING)
 * (src/slam_frontend.cc:525-527, :247).
 * q: nq rows of desc_bytes bytes, q_stride bytes apart (cv::Mat::step);
 * idx, dist: nq x 2 int32, neighbour 0 then 1, ordered by (distance, trainIdx)
 * ascending — ties resolve to the lowest train index, as OpenCV does.  Missing
 * neighbours (nt < 2) are reported as idx = -1, dist = -1. */
int vsf_knn2(vsf_ctx* ctx, const uint8_t* q, int nq, size_t q_stride,
             const uint8_t* t, int nt, size_t t_stride,
             int32_t* idx, int32_t* dist);

/* ------------------------------------------------- a2: Frontend::GetMatches */

/* Replaces Frontend::GetMatches(frame_query, frame_train, nn_match_ratio)
 * (src/slam_frontend.cc:521-538): kNN k=2 then `dist1 < ratio * dist2`
 * evaluated in double.  out receives the survivors in ascending queryIdx order
 * (imgIdx = 0); *n_out their number.  nt < 2 yields no match (the reference
 * reads out of bounds there).  VSF_ERR_CAPACITY if cap is too small. */
int vsf_get_matches(vsf_ctx* ctx, const uint8_t* q, int nq, size_t q_stride,
                    const uint8_t* t, int nt, size_t t_stride,
                    double nn_match_ratio, vsf_dmatch* out, int cap, int* n_out);

/* --------------------------- a3/a4: window loop of Frontend::ObserveImage */

/* The sliding window `frame_list_` (src/slam_frontend.h:193) lives on the
 * device.  vsf_window_push appends a frame's descriptors and evicts the oldest
 * one first when the window already holds `window` frames — the order used at
 * src/slam_frontend.cc:467-470. */
int vsf_window_push(vsf_ctx* ctx, uint64_t frame_id, const uint8_t* desc, int n,
                    size_t stride);
/* Push the frame most recently handed to vsf_window_match /
 * vsf_window_feature_matches (it is already on the device) without uploading
 * it again; n = its row count. */
int vsf_window_commit(vsf_ctx* ctx, uint64_t frame_id, int n);
int vsf_window_clear(vsf_ctx* ctx);
int vsf_window_size(const vsf_ctx* ctx);

/* Replaces the loop `for (Frame& past_frame : frame_list_)
 * GetMatches(past_frame, curr_frame, ratio)` (src/slam_frontend.cc:424-434 with
 * :287-288) by ONE kernel launch over all resident past frames (query side)
 * against the current frame (train side).  Output for past frame j (oldest
 * first): frame_ids[j], counts[j] and counts[j] matches at
 * out + j*cap_per_frame, in ascending queryIdx order.  *n_frames = number of
 * past frames.  Does not push the current frame. */
int vsf_window_match(vsf_ctx* ctx, const uint8_t* desc, int n, size_t stride,
                     double nn_match_ratio, uint64_t* frame_ids, int* counts,
                     vsf_dmatch* out, int cap_per_frame, int* n_frames);

/* Same, followed by the rest of Frontend::GetFeatureMatches
 * (src/slam_frontend.cc:289-296): order by distance, keep the first
 * int(count * best_percent), emit FeatureMatch(queryIdx, trainIdx).
 * sort_mode 0: the sort runs on the device and is stable, i.e. ordered by
 * (distance, queryIdx); sort_mode 1: the library calls std::sort on the host
 * exactly as the reference does (its order inside equal-distance groups is
 * libstdc++-specific); sort_mode 2: the same order as sort_mode 1, bit for bit,
 * produced on the device by replaying libstdc++'s introsort (csrc/sort_kernel.cu) -
 * no host cores needed, only the kept FeatureMatches cross PCIe; needs
 * max_features <= 24576 (VSF_ERR_CAPACITY otherwise); sort_mode 3: the reference
 * order wherever it is cheaper - on the host (mode 1) when the ctx may use at least 12
 * host threads (vsf_set_host_threads; 8 for the blocking vsf_window_feature_matches) or
 * max_features > 24576, on the device (mode 2) otherwise, e.g. when the ranks of a
 * many-GPU node share its cores.  Modes 0 and
 * 1 / 2 / 3 differ only inside groups of equal distance. */
#define VSF_SORT_STABLE_DEVICE 0
#define VSF_SORT_EXACT_HOST 1
#define VSF_SORT_EXACT_DEVICE 2
#define VSF_SORT_EXACT_AUTO 3
int vsf_window_feature_matches(vsf_ctx* ctx, const uint8_t* desc, int n,
                               size_t stride, double nn_match_ratio,
                               float best_percent, int sort_mode,
                               uint64_t* frame_ids, int* counts,
                               vsf_feature_match* out, int cap_per_frame,
                               int* n_frames);

/* Pipelined form of vsf_window_feature_matches + vsf_window_commit for a frame
 * stream (the bag loop of src/slam_frontend_main.cc:236-328 feeding
 * ObserveImage): vsf_window_submit enqueues the upload of the frame, the match
 * of every resident past frame against it (src/slam_frontend.cc:424-434), the
 * device sort when sort_mode == 0, then evicts / pushes the frame into the
 * window (:467-470) and returns WITHOUT waiting for the device.  The match
 * lists of frame t do not feed frame t+1 (only its descriptors do), so up to
 * VSF_PIPELINE_DEPTH frames may be in flight: the host sorts and consumes the
 * lists of frame t while the device matches frame t+1 (sort_mode 1: on the
 * library's worker threads, the lists of all frames in flight concurrently).
 * desc is copied to pinned staging before the call returns, unless flags has
 * VSF_SUBMIT_PINNED_DESC: the caller then promises that desc is page-locked
 * (cudaHostAlloc / cudaHostRegister), holds 32- or 64-byte rows back to back,
 * and stays unchanged until the frame has been collected; the copy engine
 * reads it in place.  Uploads, kernels and downloads run on three streams:
 * frame t+1 is uploaded and frame t-1 downloaded while frame t is matched.
 * vsf_window_collect waits for the OLDEST submitted frame and returns exactly
 * what vsf_window_feature_matches would have returned for it (same argument
 * meaning; *frame_id = the id given at submit; sort_mode / best_percent are the
 * ones given at submit).  VSF_ERR_STATE: submit with VSF_PIPELINE_DEPTH frames
 * in flight, or collect with none. */
#define VSF_PIPELINE_DEPTH 16
#define VSF_SUBMIT_PINNED_DESC 1
/* together with VSF_SUBMIT_PINNED_DESC for descriptors narrower than the device row (61 -> 64):
 * the rows are already padded to vsf_device_row_bytes(ctx) with zero bytes */
#define VSF_SUBMIT_PADDED_ROWS 2
int vsf_window_submit(vsf_ctx* ctx, uint64_t frame_id, const uint8_t* desc, int n,
                      size_t stride, double nn_match_ratio, float best_percent,
                      int sort_mode, int flags);
int vsf_window_collect(vsf_ctx* ctx, uint64_t* frame_id, uint64_t* frame_ids,
                       int* counts, vsf_feature_match* out, int cap_per_frame,
                       int* n_frames);
int vsf_window_in_flight(const vsf_ctx* ctx);
/* Host threads the library may use for the reference's std::sort of sort_mode 1
 * (default: min(16, hardware threads)); set it to cores / ranks when several
 * ranks share one host.  Call before the first vsf_window_submit. */
int vsf_set_host_threads(vsf_ctx* ctx, int n);

/* Bytes that crossed PCIe for the most recent vsf_window_match /
 * vsf_window_feature_matches / vsf_window_collect call: the uploaded descriptor
 * rows, and the match records + counts the device stored into host memory
 * (sort_mode 1 brings back every ratio survivor, sort_mode 0 only the kept
 * FeatureMatches).  Bench accounting. */
int vsf_window_last_transfer(const vsf_ctx* ctx, size_t* h2d_bytes, size_t* d2h_bytes);

/* ------------------------- a5: stereo L->R match + RemoveAmbigStereo */

/* Replaces `GetMatches(curr, right, ratio)` + Frontend::RemoveAmbigStereo
 * (src/slam_frontend.cc:414-417, :353-398).
 * fundamental: 9 floats row-major, convention x_left^T F x_right (:380-381).
 * The adaptive threshold (file-scope static at :353 in the reference) lives in
 * the ctx: starts at 10000, becomes mean(residual over all matches) + 2 after
 * every call; vsf_set/get_stereo_threshold expose it (the 1-float halo needed
 * to start a shard in the middle of a sequence).
 * Outputs (all optional except n_kept): kept_left/kept_right = indices into the
 * input frames of the M surviving pairs, in match order, so left'[i] =
 * left[kept_left[i]], right'[i] = right[kept_right[i]]; stereo_matches /
 * residuals / n_stereo = the pre-filter L->R matches and their residuals.
 * (vsf_observe_features is the fused entry point that keeps the compacted
 * frames on the device for the window and triangulation stages.) */
int vsf_stereo_filter(vsf_ctx* ctx,
                      const vsf_keypoint* kp_left, const uint8_t* desc_left,
                      int n_left, size_t stride_left,
                      const vsf_keypoint* kp_right, const uint8_t* desc_right,
                      int n_right, size_t stride_right,
                      const float* fundamental, double nn_match_ratio,
                      int32_t* kept_left, int32_t* kept_right, int* n_kept,
                      vsf_dmatch* stereo_matches, float* residuals, int* n_stereo);
int vsf_set_stereo_threshold(vsf_ctx* ctx, float value);
int vsf_get_stereo_threshold(vsf_ctx* ctx, float* value);

/* Behaviour switches of the stereo stage (both default to 0).
 * VSF_OPT_RESIDUAL_ORDER: float32 summation order of the two 3-term dot products inside
 *   `(left_ph.transpose() * F * right_ph).norm()` (src/slam_frontend.cc:380-381).  The order is
 *   Eigen's, and Eigen is not vendored by the reference: 0 = left to right, (c0 + c1) + c2
 *   (Eigen 3.2's unrolled coefficient products), 1 = c0 + (c1 + c2) (Eigen 3.3's
 *   redux_novec_unroller, which halves the range).  The two differ by at most one ulp of the
 *   residual, i.e. only for pairs sitting exactly on the threshold.
 * VSF_OPT_HOLD_THRESHOLD_ON_EMPTY: 0 = the reference's arithmetic, a frame without stereo
 *   matches sets the threshold to 0/0 + 2 = NaN and every later frame loses all its pairs
 *   (src/slam_frontend.cc:392-394); 1 = such a frame leaves the threshold unchanged. */
enum {
  VSF_OPT_RESIDUAL_ORDER = 1,
  VSF_OPT_HOLD_THRESHOLD_ON_EMPTY = 2,
  VSF_OPT_POSE_GROUP = 3,          /* 1 .. 8, default 4: poses vsf_window_match_block_device / vsf_window_run_sequence launch
                                      as ONE batch (at most 40 frame pairs: one distance kernel, one finish kernel); results
                                      do not depend on it */
  VSF_OPT_DEBUG_SORT_DEPTH = 100   /* tests: depth limit of sort_mode 2's introsort replay, -1 = 2 * floor(lg n) */
};
int vsf_set_option(vsf_ctx* ctx, int option, int value);
int vsf_get_option(const vsf_ctx* ctx, int option, int* value);

/* ------------------------------------------- a6: cv::triangulatePoints */

/* Replaces cv::triangulatePoints(P1, P2, pts1, pts2, out)
 * (src/slam_frontend.cc:152-156).  P1, P2: 3x4 row-major float32.  x1, x2:
 * n (x,y) float32 pairs.  X4: 4 x n row-major float32 like the 4xN CV_32F
 * matrix OpenCV returns (unit-norm homogeneous columns; the sign of a column
 * is arbitrary and cancels in xyz/w).  Internals are fp64. */
int vsf_triangulate(vsf_ctx* ctx, const float* P1, const float* P2,
                    const float* x1, const float* x2, int n, float* X4);

/* N1: cv::undistortPoints(pts, K, dist, R = {}, P = K) as Frontend::UndistortFeaturePoints
 * calls it (src/slam_frontend.cc:323-351).  K: 3x3 row-major, dist: k1 k2 p1 p2 k3,
 * xy / out: n interleaved (x, y) pairs.  Five fixed-point iterations in fp64, float output
 * (matches OpenCV to float rounding). */
int vsf_undistort_points(vsf_ctx* ctx, const float* K, const float* dist, const float* xy, int n, float* out);

/* ----------------- whole matching path of Frontend::ObserveImage, fused */

typedef struct vsf_observe_out {
  /* stereo stage (a5) */
  int32_t* kept_left;          /* [cap] */
  int32_t* kept_right;         /* [cap] */
  int n_kept;                  /* M */
  float stereo_threshold_next; /* threshold after this frame */
  /* window stage (a4): past frames oldest first */
  int n_frames;
  uint64_t* frame_ids;         /* [window] */
  int* window_counts;          /* [window] */
  vsf_dmatch* window_matches;  /* [window][cap], query order */
  /* triangulation stage (a6): R'->L' matches in query order + their points */
  int n_tri;
  vsf_dmatch* tri_matches;     /* [cap] */
  float* tri_X4;               /* [cap][4] homogeneous points, match order */
  int cap;                     /* capacity of the per-frame arrays: must be >= n_left AND >= the
                                * row count of every frame resident in the window (a window list
                                * has one entry per row of the PAST frame at most); the ctx's
                                * max_features always suffices.  VSF_ERR_CAPACITY otherwise, with
                                * the ctx state (window, stereo threshold) left as before the call. */
  /* N1 (src/slam_frontend.cc:323-351), filled by vsf_observe_collect when the frame was submitted
   * with K_left / dist_left: the undistorted pixel of compacted left keypoint i */
  float* xy_undist;            /* [cap][2], optional */
} vsf_observe_out;

/* Replaces everything Frontend::ObserveImage does between ExtractFeatures and
 * the SLAMNode assembly (src/slam_frontend.cc:414-437), device-resident:
 * stereo kNN + ratio + epipolar filter + compaction, window kNN against all
 * resident frames, R'->L' kNN + ratio + triangulation, then pushes the
 * compacted left frame into the window (evicting per :467-469).  The sort /
 * best_percent cut / FeatureMatch book-keeping stay with the caller (the C++
 * Frontend mirror does them with std::sort like the reference). */
int vsf_observe_features(vsf_ctx* ctx, uint64_t frame_id,
                         const vsf_keypoint* kp_left, const uint8_t* desc_left,
                         int n_left, size_t stride_left,
                         const vsf_keypoint* kp_right, const uint8_t* desc_right,
                         int n_right, size_t stride_right,
                         const float* fundamental, const float* P_left,
                         const float* P_right, double nn_match_ratio,
                         vsf_observe_out* out);

/* Pipelined form of vsf_observe_features for a frame stream.  Everything a frame needs from its
 * predecessors lives on the device (the compacted left frames in the window ring with their row
 * counts, the adaptive stereo threshold), so vsf_observe_submit enqueues a whole frame - one
 * upload of both images' descriptors and pixels, stereo kNN + filter + compaction, window kNN,
 * R'->L' kNN, triangulation (+ the undistorted pixels of the kept left keypoints when K_left and
 * dist_left are given, N1) - and returns WITHOUT waiting; the kernels store their results
 * straight into page-locked host memory.  Up to VSF_OBSERVE_DEPTH frames may be in flight;
 * vsf_observe_collect waits for the OLDEST one and fills `out` exactly as vsf_observe_features
 * would have (out->cap as documented above; a collect that fails with VSF_ERR_CAPACITY has
 * consumed the frame, which stays in the window).  vsf_observe_features itself is submit +
 * collect. */
typedef struct vsf_observe_params {
  const float* fundamental;  /* 9, row-major, x_left^T F x_right */
  const float* P_left;       /* 12 */
  const float* P_right;      /* 12 */
  const float* K_left;       /* 9, optional (with dist_left): undistort the kept left pixels */
  const float* dist_left;    /* 5: k1 k2 p1 p2 k3 */
  double nn_match_ratio;
} vsf_observe_params;
#define VSF_OBSERVE_DEPTH 4
int vsf_observe_submit(vsf_ctx* ctx, uint64_t frame_id,
                       const vsf_keypoint* kp_left, const uint8_t* desc_left,
                       int n_left, size_t stride_left,
                       const vsf_keypoint* kp_right, const uint8_t* desc_right,
                       int n_right, size_t stride_right,
                       const vsf_observe_params* params);
int vsf_observe_collect(vsf_ctx* ctx, uint64_t* frame_id, vsf_observe_out* out);
int vsf_observe_in_flight(const vsf_ctx* ctx);

/* ------------------------ device-resident entry points (no host copies) */

/* Asynchronous on the ctx stream; all pointers are device pointers.
 * Descriptor rows must already be padded to vsf_device_row_bytes(ctx). */
int vsf_device_row_bytes(const vsf_ctx* ctx);

/* Window matching for one pose of a device-resident sequence: query frame j
 * has nq[j] rows at d_queries[j], train frame has nt rows at d_train.  Results
 * go to the ctx's device buffers (read them back with vsf_fetch_window). */
int vsf_window_match_device(vsf_ctx* ctx, const void* const* d_queries,
                            const int* nq, int n_frames, const void* d_train,
                            int nt, double nn_match_ratio);
int vsf_fetch_window(vsf_ctx* ctx, int n_frames, int* counts, vsf_dmatch* out,
                     int cap_per_frame);

/* The same for a run of `count` consecutive poses of a device-resident sequence buffer that
 * holds n_poses frames of n rows each (rows vsf_device_row_bytes(ctx) wide, back to back): for
 * k in [0, count) the current frame is c = (first + k) mod (n_poses - window) + window and the
 * query frames are c - window .. c - 1 (the bag loop of src/slam_frontend_main.cc:236-328 with
 * src/slam_frontend.cc:424-434 inside, device-resident).  Asynchronous; the results of the last
 * pose stay in the ctx's device buffers (vsf_fetch_window).  On the tensor engine the poses are
 * launched in groups (VSF_OPT_POSE_GROUP, group x window <= 40 frame pairs), a group as ONE batch
 * of two kernels: the distance kernel, which also expands the next group's frames, and one
 * kernel for refine + ordered compaction (csrc/knn2_tc_kernel.cu: knn2_tc_finish_kernel). */
int vsf_window_match_block_device(vsf_ctx* ctx, const void* d_seq, int n, int n_poses,
                                  long long first, int count, double nn_match_ratio);

/* Host-buffer counterpart: streams `count` frames of a PAGE-LOCKED host buffer (n_poses frames
 * of n rows, device row width, back to back; frame k of the run = frame (first + k) mod n_poses
 * of the buffer, frame id first + k) through vsf_window_submit / vsf_window_collect with `lag`
 * frames submitted ahead of the one being collected (1 .. VSF_PIPELINE_DEPTH - 1), i.e. the
 * loop a C++ caller writes around the two calls.  The FeatureMatch lists of frame k are written
 * to out + (k mod ring) * window * cap_per_frame (list j at + j * cap_per_frame) and their
 * lengths to counts + (k mod ring) * window, so a ring of >= 1 frames of output is enough when
 * the caller only wants the lists to have reached host memory.  The window must hold the frames
 * the caller wants frame `first` matched against (vsf_window_push).  Returns after every frame
 * has been collected; *h2d_bytes / *d2h_bytes (optional) accumulate vsf_window_last_transfer.
 * Knowing the frames ahead, the call uploads min(VSF_OPT_POSE_GROUP, lag / 3) of them before
 * it launches their kernels as one batch (like vsf_window_match_block_device); the lists are
 * the same.  lag = 12 keeps groups of four frames flowing. */
int vsf_window_run_sequence(vsf_ctx* ctx, const uint8_t* h_seq, int n, int n_poses,
                            long long first, int count, double nn_match_ratio,
                            float best_percent, int sort_mode, int lag,
                            vsf_feature_match* out, int* counts, int ring, int cap_per_frame,
                            size_t* h2d_bytes, size_t* d2h_bytes);

/* Where the compaction kernels leave the match lists of the most recent window launch: region j
 * (past frame j, oldest first) = *d_lists + j * *stride records, its length at (*d_counts)[j];
 * *regions = number of regions.  Device pointers, valid for the life of the ctx; contents are
 * ordered on the ctx's stream (vsf_stream).  For consumers that keep the lists on the device
 * (libvsf_nccl's vsf_gather_matches). */
int vsf_device_match_lists(vsf_ctx* ctx, const vsf_dmatch** d_lists, const int** d_counts,
                           int* stride, int* regions);
/* The CUDA stream (cudaStream_t as void*) the ctx currently launches on. */
void* vsf_stream(vsf_ctx* ctx);

/* Synthetic sequence generator (bench / test frame source; counter-based so
 * any pose range can be produced on any rank): pose p observes landmarks
 * [stride*p, stride*p + n) in a per-pose affine permutation, every bit flipped
 * with probability 1/32.  Writes n rows per pose for poses
 * [first_pose, first_pose + n_poses) at d_out, vsf_device_row_bytes(ctx) bytes per row (bytes
 * at and beyond the ctx's desc_bytes are zero). */
int vsf_synth_sequence_device(vsf_ctx* ctx, void* d_out, int n, int first_pose,
                              int n_poses, int stride, uint64_t seed);

/* Integer-pipe probe used for the roofline denominator: runs `iters` dependent
 * instructions of kind (0 = POPC, 1 = LOP3, 2 = POPC+LOP3 interleaved 1:1,
 * 3 = IMAD, 4 = VIMNMX, 5 = POPC+LOP3 1:2, 6 = IADD3) on every SM and returns
 * the measured lane-operations per second. */
int vsf_probe_pipe(vsf_ctx* ctx, int kind, int iters, double* ops_per_second);

int vsf_device_sm_count(const vsf_ctx* ctx);

/* Host side of the sort + cut of sort_mode 1, no device needed (CPU tests): keys are
 * (distance << 22 | position); afterwards keys[0 .. keep) are what std::sort by distance leaves
 * there (csrc/exact_sort.h). */
int vsf_debug_sort_prefix(uint32_t* keys, int n, int keep);
/* The same with introsort's depth limit (normally 2 * floor(lg n)) forced to `depth`, which makes
 * the heapsort fallback reachable in tests. */
int vsf_debug_sort_prefix_depth(uint32_t* keys, int n, int keep, int depth);

/* The device sort + cut of sort_mode 0 / 2 on a caller-supplied list (tests): `n` matches in, the
 * first int(n * best_percent) FeatureMatches in the order of the mode out (exact = 1: the
 * reference's std::sort order, csrc/sort_kernel.cu).  n <= max_features. */
int vsf_debug_sort_device(vsf_ctx* ctx, const vsf_dmatch* matches, int n, float best_percent,
                          int exact, vsf_feature_match* out, int* n_out);

/* Host-side planner of the tensor engine's work partition, no device needed (CPU tests): for a
 * launch of query_blocks 256-query blocks against train_tiles 256-row tiles on sm_count SMs,
 * out5 = {pieces per block, tiles per piece, CTAs, partial slots per query row, segments of
 * block 0}.  rows / partial_cap: the capacity check of the partial-key buffer (uint2 units). */
int vsf_debug_tc_plan(int query_blocks, int train_tiles, int sm_count, int force_split,
                      long long rows, long long partial_cap, int* out5);

/* Kernels launched so far by the kNN paths of this ctx (bench.py's gpu_launches). */
long long vsf_debug_launch_count(const vsf_ctx* ctx);

/* Bring-up aid: per-CTA timeline of the last tensor-engine launch made with engine flag 16
 * (vsf_set_engine(ctx, 2, 16)): 16 values per CTA, layout documented in
 * csrc/knn2_tc_kernel.cu.  Waits for the ctx stream. */
int vsf_debug_tc_trace(vsf_ctx* ctx, long long* out, int max_ctas, int* n_ctas);
/* Bring-up aid: kernel-level timeline of the tensor-engine launches made since
 * vsf_set_engine(ctx, engine, 32): per launch 16 values = earliest start / latest end
 * (globaltimer ns) of the expansion, distance, refine (or refine + compaction: finish) and
 * compaction kernels, then of the refine after its wait; 6 spare values.  At most 256 launches
 * are kept. */
int vsf_debug_kernel_trace(vsf_ctx* ctx, long long* out, int max_records, int* n_records);

#ifdef __cplusplus
}
#endif
#endif  /* VSF_H_ */
