// C ABI of libvsf_cuda.so (declared in include/vsf.h): context, buffers, host<->device
// staging and the launch sequences.  No CPU fallback anywhere: if CUDA is unavailable
// every entry point reports VSF_ERR_CUDA.
#include <algorithm>
#include <atomic>
#include <chrono>
#ifdef __linux__
#include <sys/prctl.h>
#endif
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "exact_sort.h"
#include "stereo_args.cuh"
#include "vsf_device.cuh"

namespace vsf {
cudaError_t launch_knn2(const KnnBatch& batch, int words, int R, int mode, int variant,
                        int max_qblocks, cudaStream_t stream);
cudaError_t launch_expand_train(const void* t, int nt_bound, const int* nt_dev, void* out, int int8,
                                int pdl, cudaStream_t stream, long long* ktrace);
cudaError_t launch_knn2_tc(const KnnBatch& batch, const TcBatch& tc, int int8, int max_nq, int pdl,
                           cudaEvent_t* ev, cudaStream_t stream, FinishArgs* fa, int* launched, int phase);
cudaError_t launch_expand_train_multi(const ExpandMulti& em, int int8, int pdl, cudaStream_t stream);
cudaError_t launch_expand_train64(const void* t, int nt_bound, const int* nt_dev, void* out, int pdl,
                                  cudaStream_t stream);
cudaError_t launch_knn2_tc64(const KnnBatch& batch, const TcBatch& tc, int max_nq, int pdl, cudaEvent_t* ev,
                             cudaStream_t stream, FinishArgs* fa, int* launched, int phase);
cudaError_t launch_expand_train64_multi(const ExpandMulti& em, int pdl, cudaStream_t stream);
cudaError_t launch_synth(uint32_t* out, int n, int first_pose, int n_poses, int stride,
                         uint64_t seed, int words, int desc_bytes, cudaStream_t stream);
int probe_ops_per_step(int kind);
cudaError_t launch_probe(int kind, uint32_t* sink, int iters, int blocks, int threads,
                         cudaStream_t stream);
cudaError_t launch_triangulate_pairs(const float* P1, const float* P2, const float2* x1,
                                     const float2* x2, int n, float* X4, cudaStream_t stream);
cudaError_t launch_triangulate_matches(const float* P1, const float* P2, const vsf_dmatch* matches,
                                       const int* n_matches, int max_matches,
                                       const float2* xy_left_c, const float2* xy_right_c,
                                       float4* X4, const TriExtras* extras, cudaStream_t stream);
cudaError_t launch_undistort_points(const float2* in, int n, const float* K9, const float* dist5, float2* out,
                                    cudaStream_t stream);
cudaError_t launch_sort_cut(const vsf_dmatch* const* matches, const int* const* counts,
                            int n_problems, float best_percent, vsf_feature_match* out,
                            int out_stride, int* out_counts, int bins, int max_rows, int exact,
                            int depth_override, void* gscratch, cudaStream_t stream, long long* trace = nullptr);
int sort_exact_max_rows();
size_t sort_exact_global_scratch(int max_rows);
}  // namespace vsf


using namespace vsf;

constexpr int kKtracePoses = 256;   // launches kept by the kernel-level timeline (engine flag 32)

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct ProblemSpec {
  const void* q;
  int nq;
  const int* nq_dev;
  const void* t;
  int nt;
  const int* nt_dev;
  int region;  // which d_matches region / d_match_count slot receives the survivors
  // optional: the train frame already expanded to +-1 bytes (tensor engine image) by the caller,
  // in the operand kind t_exp_int8 says (1 = int8, 0 = e4m3); run_knn then skips its expansion
  const uint8_t* t_exp = nullptr;
  int t_exp_int8 = 1;
  // optional (batches that span several frames' buffers): where the survivors / their count go
  // instead of the ctx's region `region`, and a mapped host word that also receives the count
  vsf_dmatch* out = nullptr;
  int* out_count = nullptr;
  int* host_count = nullptr;
};

// A stream of poses (vsf_window_match_block_device, vsf_window_run_sequence): how the kernels of
// a launch - one pose, or a group of poses as ONE batch - are launched.
struct PoseLaunch {
  int early = 0;                     // TcBatch::early
  int late = 0;                      // TcBatch::late
  uint2* partial = nullptr;          // partial-key buffer of this launch (nullptr: the ctx's)
  unsigned long long* flags = nullptr;   // look-back words / launch state of the finish kernel
  unsigned long long* state = nullptr;   //   (nullptr: the ctx's)
};

struct vsf_ctx {
  int device = 0, max_features = 0, desc_bytes = 0, row_bytes = 0, words = 0, window = 0;
  int sm_count = 0;
  int host_threads = 1;   // for the host-side std::sort of sort_mode 1
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t pev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // per-kernel timing (vsf_set_profile)
  int profile = 0;
  bool pev_valid = false;
  std::string err;
  int popc_mode = -1, force_split = 0, force_R = 0, variant = -1;   // -1 / 0 = library default
  int engine = 0;         // 0 auto, 1 POPC pipe, 2 tensor cores int8, 3 tensor cores e4m3
  int engine_flags = 0;   // timing experiments only (TcBatch::flags)
  long long* d_tc_trace = nullptr;   // engine flag 16: per-CTA timeline of knn2_tc_kernel
  long long* d_ktrace = nullptr;     // engine flag 32: kernel-level timeline, kKtracePoses records of [8][2]
  long long ktrace_n = 0;
  int last_engine = 0;    // engine the last kNN launch used
  int want_second_index = 0;   // set around vsf_knn2's launch: the tensor engine's refine must produce the exact idx[1]
  // automatic mode: tensor cores from this many comparisons per batch (measured crossover,
  // tools/engine_crossover.py: 1e6 POPC 12 us vs 20; 4e6 18 vs 14; 2.5e7 52 vs 24)
  double tc_auto_min_cmp = 3e6;
  // the same for blocking one-frame calls (vsf_knn2 / vsf_get_matches / vsf_observe_*).  Measured
  // (tools latency probe, C2 2000 x 2000): a higher threshold does NOT pay - the single POPC
  // launch saves ~15 us of host launch time but its kernel is ~15 us longer, 40 us either way, and
  // the whole C3 frame gets slower (99 vs 90 us) - so both thresholds are the same.
  double tc_auto_min_cmp_latency = 3e6;
  uint8_t* d_train_exp[kTcMaxTrains] = {nullptr, nullptr};  // +-1 expanded train images

  int rows_pad = 0;    // max_features rounded up to 128
  int regions = 0;     // window + 2
  // device
  uint8_t* d_ring = nullptr;
  uint8_t *d_raw_left = nullptr, *d_raw_right = nullptr, *d_right_c = nullptr;
  float2 *d_xy_left = nullptr, *d_xy_right = nullptr, *d_xy_left_c = nullptr, *d_xy_right_c = nullptr;
  uint4* d_knn_out = nullptr;
  uint2* d_partial = nullptr;
  uint2* d_partial2 = nullptr;              // streams of poses: every other pose (allocated on first use)
  // groups of poses (vsf_window_match_block_device): per-slot buffers, allocated on first use
  int pose_group = 4;                       // VSF_POSE_GROUP (1 .. kMaxPoseGroup)
  uint8_t* grp_exp[2 * kMaxPoseGroup] = {};   // +-1 images: a group reads one half while the next group's are made
  unsigned long long* grp_flags = nullptr;  // look-back words of a batch of kMaxProblems problems + 2 launch states
  vsf_dmatch* grp_matches = nullptr;        // [kMaxPoseGroup - 1][regions][rows_pad]: survivors of the poses that do not write the ctx's own lists
  int* grp_counts = nullptr;                // [kMaxPoseGroup - 1][kMaxProblems]
  size_t qb_cap = 0;
  long long launches = 0;                   // kernels launched by run_knn (vsf_debug_launch_count)
  double tm_wait = 0.0, tm_copy = 0.0, tm_ready = 0.0;   // VSF_TIMING: collect's wait / copy-out, launch -> ready (us, summed)
  long long tm_n = 0;
  // knn2_tc_finish_kernel: ticket counter, per-block (epoch | survivor count) words, and the
  // host's copies of the running values
  unsigned long long *d_finish_ticket = nullptr, *d_finish_flags = nullptr;   // ticket: two FinishArgs::state (4 words each)
  int finish_parity = 0;
  size_t partial_cap = 0;
  unsigned *d_qblock_arrivals = nullptr, *d_qblock_pass = nullptr, *d_problem_arrivals = nullptr;
  vsf_dmatch* d_matches = nullptr;
  int* d_match_count = nullptr;
  float* d_resid = nullptr;
  unsigned *d_chunk_keep = nullptr, *d_chunk_off = nullptr, *d_ticket = nullptr;
  int opt_residual_order = 0, opt_hold_on_empty = 0;   // vsf_set_option
  int sort_depth_override = -1;                        // VSF_OPT_DEBUG_SORT_DEPTH (tests)
  uint8_t* d_sort_scratch = nullptr;                   // sort_mode 2 on lists too long for shared memory
  int *d_kept_left = nullptr, *d_kept_right = nullptr;
  int* d_slot_rows = nullptr;   // [ring_slots] device-side row count of every ring slot (compacted frames)
  float* d_thresh = nullptr;  // [2], ping-pong
  int thresh_cur = 0;
  float4* d_X4 = nullptr;
  float* d_tri_io = nullptr;  // stateless triangulate: x1 | x2 | X4
  uint32_t* d_sink = nullptr;
  vsf_feature_match* d_fm = nullptr;  // [window][rows_pad] sorted+cut feature matches
  int* d_fm_count = nullptr;
  // pinned host staging
  uint8_t* h_desc[2] = {nullptr, nullptr};
  float2* h_xy[2] = {nullptr, nullptr};
  int* h_counts = nullptr;         // kMaxProblems + 8
  vsf_dmatch* h_matches = nullptr;  // regions * rows_pad, mapped: kernels mirror match lists into it
  int* h_region_counts = nullptr;   // kMaxProblems, mapped: survivor count of every region
  vsf_dmatch* dm_matches = nullptr; // device-side addresses of the two mapped buffers
  int* dm_region_counts = nullptr;
  int* h_kept[2] = {nullptr, nullptr};
  float* h_resid = nullptr;
  float4* h_X4 = nullptr;
  uint4* h_knn = nullptr;
  float* h_tri_io = nullptr;
  float* h_scalar = nullptr;
  vsf_feature_match* h_fm = nullptr;
  uint32_t* h_keys = nullptr;   // plain host scratch for the host sort (allocated on first use)
  // sliding window (frame_list_) state
  std::vector<int> slot_count;      // rows of the frame in a ring slot (an upper bound while slot_dev)
  std::vector<char> slot_dev;       // the exact count is only on the device yet (d_slot_rows)
  std::vector<uint64_t> slot_frame;
  std::deque<int> live;   // slot indices, oldest first
  std::deque<int> free_slots;   // FIFO: a slot evicted at frame t is reused for frame t+2, so the
                                // upload of frame t+1 never touches rows the kernels of frame t read
  int staging_slot = 0;
  int ring_slots = 0;     // window + 2
  int last_n_frames = 0;  // of the last vsf_window_match_device
  // where mirrored launches store their match lists / counts (the ctx's own mapped buffers, or
  // those of a pipelined submission)
  vsf_dmatch* mir_dm = nullptr;
  int* mir_dcounts = nullptr;
  int* mir_hcounts = nullptr;
  // pipelined window matching (vsf_window_submit / vsf_window_collect): buffers allocated on
  // first use, one set per frame in flight
  struct Flight {
    std::chrono::steady_clock::time_point t_flush;   // VSF_TIMING: when the frame's kernels were launched
    cudaEvent_t ev_up = nullptr;        // upload stream: the frame's rows are in the ring
    cudaEvent_t ev_chain = nullptr;     // main stream: the kernels of this frame have finished
    cudaEvent_t done = nullptr;         // download stream: the lists are in host memory
    cudaEvent_t ev_sort = nullptr;      // sort stream: the device sort + cut of this frame has finished
    int* d_counts = nullptr;            // device: [kMaxProblems] survivor counts of this frame
    uint8_t* h_desc = nullptr;          // pinned staging of the submitted frame
    uint8_t* d_train_exp = nullptr;     // device: the frame expanded to +-1 bytes on the upload stream
    vsf_dmatch* d_matches = nullptr;    // device: [window][rows_pad] ratio survivors of this frame
    vsf_feature_match* d_fm = nullptr;  // device: [window][rows_pad] sorted + cut (sort_mode 0)
    vsf_dmatch* h_matches = nullptr;    // pinned copy of d_matches (sort_mode 1)
    vsf_feature_match* h_fm = nullptr;  // pinned copy of d_fm (sort_mode 0) / output of the host sort
    int* h_counts = nullptr;            // mapped: [kMaxProblems], written by the kernels
    int* dm_counts = nullptr;           // its device-side alias
    size_t d2h_bytes = 0;
    int nf = 0, sort_mode = 1;
    float best_percent = 1.f;
    uint64_t frame_id = 0;
    uint64_t fids[kMaxProblems];
    size_t h2d_bytes = 0;
    // host-side finish of sort_mode 1 (worker threads): per past frame the kept count, pending
    // = lists not finished yet (+1 while the device is still running)
    uint32_t* keys = nullptr;           // [window][rows_pad] packed (distance, position) sort keys
    int keep[kMaxProblems];
    std::atomic<int> pending{0};
    int cuda_error = 0;
    // staged by vsf_window_submit, launched by flush_flights (a group of frames at a time)
    bool launched = false;
    std::vector<ProblemSpec> specs;
    double ratio = 0.0;
    int n = 0, slot = 0, max_cnt = 0;
    cudaEvent_t chain = nullptr;        // the ev_chain recorded behind this frame's kernels (its group's last frame's)
  };
  struct SortTask {
    Flight* f;
    int list;   // sort + cut that past frame's list
  };
  // One dispatcher thread waits for each submitted frame's download and turns its lists into
  // tasks for the workers.  Three separate locks so that the caller's thread (submit / collect)
  // never queues behind the workers: disp_* (caller -> dispatcher), pool_* (dispatcher ->
  // workers), done_* (workers -> caller).
  std::vector<std::thread> workers;
  std::thread dispatcher;
  std::mutex disp_mu, pool_mu, done_mu;
  std::condition_variable disp_cv, pool_cv, done_cv;
  std::deque<Flight*> disp_q;
  std::deque<SortTask> pool_q;
  bool pool_stop = false;
  bool dispatch_spin = false;   // dispatcher waits with cudaEventSynchronize (spins on a core) instead of polling
  Flight flights[VSF_PIPELINE_DEPTH];
  cudaStream_t up_stream = nullptr, down_stream = nullptr;   // copy engines of the pipelined path
  // device-resident blocks of poses: the +-1 images of the next group are made on a side stream
  // beside the running distance kernel; ev_img[p] = the images of parity p are complete,
  // ev_grp[p] = the group that read them has finished, ev_blk = what preceded the call has
  bool grp_ready = false;                   // group_state_init has completed
  cudaStream_t exp_stream = nullptr;
  cudaEvent_t ev_img[2] = {nullptr, nullptr}, ev_grp[2] = {nullptr, nullptr}, ev_blk = nullptr;
  cudaEvent_t ev_main = nullptr;
  bool main_dirty = false;   // non-pipelined kernels launched since the last submit may still read the ring
  std::vector<cudaEvent_t> slot_last_chain;   // per ring slot: ev_chain of the last pipelined frame that read it
  size_t last_h2d = 0, last_d2h = 0;   // PCIe bytes of the most recent window call (vsf_window_last_transfer)
  bool flights_ready = false;
  int flight_head = 0, flight_count = 0;   // FIFO: oldest = flights[flight_head]
  int flights_staged = 0;   // the newest flights_staged flights are uploaded but their kernels not launched yet
  int defer_group = 1;      // vsf_window_run_sequence: launch the kernels of this many frames together

  // pipelined full-frame path (vsf_observe_submit / vsf_observe_collect): per frame in flight one
  // upload block and one block of mapped host memory the kernels store their results into
  struct ObsFlight {
    cudaEvent_t ev_up = nullptr, ev_done = nullptr;
    uint8_t *h_in = nullptr, *d_in = nullptr;
    uint8_t* h_out = nullptr;           // mapped; the pointers below carve it up (h_* host view, dm_* device view)
    int *h_kept_left = nullptr, *dm_kept_left = nullptr, *h_kept_right = nullptr, *dm_kept_right = nullptr;
    vsf_dmatch *h_lists = nullptr, *dm_lists = nullptr;   // [window + 1][rows_pad]
    int *h_counts = nullptr, *dm_counts = nullptr;        // [kMaxProblems] list lengths, [kMaxProblems] = M
    float *h_scalar = nullptr, *dm_scalar = nullptr;      // threshold after this frame
    float4 *h_X4 = nullptr, *dm_X4 = nullptr;
    float2 *h_xyu = nullptr, *dm_xyu = nullptr;
    int nf = 0, nl = 0, slot = 0, undistort = 0;
    uint64_t frame_id = 0;
    uint64_t fids[kMaxProblems];
    int bounds[kMaxProblems];
  };
  ObsFlight obs[VSF_OBSERVE_DEPTH];
  bool obs_ready = false;
  int obs_head = 0, obs_count = 0;
  cudaStream_t obs_up_stream = nullptr;

  vsf_dmatch* match_base = nullptr;   // d_matches, or the device buffer of a pipelined submission
  int* count_base = nullptr;          // d_match_count, or the device counters of a pipelined submission
  // the device sort of the pipelined path runs on its own stream beside the next frame's distance
  // kernel, on SMs that kernel leaves free (a persistent CTA per SM otherwise)
  cudaStream_t sort_stream[2] = {nullptr, nullptr};   // alternate frames: the sorts of two frames may overlap
  unsigned sort_rr = 0;
  int reserve_sms = 0;
  int reserve_override = -1;    // VSF_RESERVE_SMS (tuning)
  int sort_streams = 2;         // VSF_SORT_STREAMS (tuning): 1 = the sorts of consecutive frames do not overlap
  uint8_t* slot_ptr(int s) const { return d_ring + size_t(s) * rows_pad * row_bytes; }
  vsf_dmatch* region_ptr(int r) const { return match_base + size_t(r) * rows_pad; }
};

#define VSF_CUDA(ctx, expr)                                                             \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                 \
      return VSF_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

static int fail(vsf_ctx* ctx, int code, const char* msg) {
  if (ctx) ctx->err = msg;
  return code;
}

static void reset_ring(vsf_ctx* c) {
  c->live.clear();
  c->free_slots.clear();
  c->staging_slot = 0;
  for (int s = 1; s < c->ring_slots; ++s) c->free_slots.push_back(s);
}

// frame_list_ eviction + push (src/slam_frontend.cc:467-470)
static void commit_staging(vsf_ctx* c, uint64_t frame_id, int count) {
  if (int(c->live.size()) >= c->window) {
    c->free_slots.push_back(c->live.front());
    c->live.pop_front();
  }
  c->slot_count[c->staging_slot] = count;
  c->slot_dev[c->staging_slot] = 0;
  c->slot_frame[c->staging_slot] = frame_id;
  c->live.push_back(c->staging_slot);
  c->staging_slot = c->free_slots.front();
  c->free_slots.pop_front();
}

// Pack caller rows (stride apart, desc_bytes wide) into pinned staging padded to
// row_bytes, then one async H2D copy.
static void pack_desc(const vsf_ctx* c, uint8_t* h, const uint8_t* src, int n, size_t stride) {
  if (stride == size_t(c->row_bytes) && c->desc_bytes == c->row_bytes) {
    std::memcpy(h, src, size_t(n) * c->row_bytes);
  } else {
    for (int i = 0; i < n; ++i) {
      std::memcpy(h + size_t(i) * c->row_bytes, src + size_t(i) * stride, c->desc_bytes);
      if (c->desc_bytes < c->row_bytes)
        std::memset(h + size_t(i) * c->row_bytes + c->desc_bytes, 0, c->row_bytes - c->desc_bytes);
    }
  }
}

static int upload_desc(vsf_ctx* c, uint8_t* h, const uint8_t* src, int n, size_t stride,
                       uint8_t* d_dst) {
  if (n == 0) return VSF_OK;
  pack_desc(c, h, src, n, stride);
  VSF_CUDA(c, cudaMemcpyAsync(d_dst, h, size_t(n) * c->row_bytes, cudaMemcpyHostToDevice, c->stream));
  return VSF_OK;
}

static int upload_desc(vsf_ctx* c, int which, const uint8_t* src, int n, size_t stride,
                       uint8_t* d_dst) {
  return upload_desc(c, c->h_desc[which], src, n, stride, d_dst);
}

static int upload_xy(vsf_ctx* c, int which, const vsf_keypoint* kp, int n, float2* d_dst) {
  if (n == 0) return VSF_OK;
  float2* h = c->h_xy[which];
  for (int i = 0; i < n; ++i) h[i] = make_float2(kp[i].x, kp[i].y);
  VSF_CUDA(c, cudaMemcpyAsync(d_dst, h, size_t(n) * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
  return VSF_OK;
}

// Granularity of the tensor engine's work slots and number of CTAs (see TcBatch): fills
// pieces / tiles_per_piece / total / grid / slots from tb->tiles and the shape of the launch.
// Pure host arithmetic (also exported as vsf_debug_tc_plan for the CPU tests).
static void plan_tc_partition(TcBatch* tbp, int qblocks, int sm, int force_split, size_t rows, size_t partial_cap) {
  TcBatch& tb = *tbp;
  const long long tile_slots = (long long)qblocks * tb.tiles;
  int S = tb.tiles;                  // pieces per block; default: one tile per piece
  long long G = 0;                   // 0: one CTA per SM (or per slot when there are fewer)
  if (force_split > 0) {
    // tuning knob (tests): about force_split pieces per block, one piece per CTA when they fit
    S = std::min(force_split, tb.tiles);
    if ((long long)qblocks * S <= sm) G = (long long)qblocks * S;
  } else if (tile_slots < 4LL * sm && qblocks <= sm) {
    // small launch: one piece of one block per CTA.  Not as many pieces as there are SMs: these
    // launches are a few tile times long and bound by the latencies of the four kernels, which
    // overlap better (programmatic dependent launch) when the distance kernel leaves about a
    // third of the SMs to its neighbours - measured optimum 64..100 CTAs on 148 SMs for every
    // shape tried, with a cliff above ~120 (3000 x 3000: 72 CTAs 14.9 us, 144 CTAs 22.3)
    const long long cap = std::max<long long>(1, (long long)sm * 11 / 16);
    S = int(std::max<long long>(1, std::min<long long>(tb.tiles, cap / qblocks)));
    G = (long long)qblocks * S;
  } else if (tile_slots > 128LL * sm) {
    // very large launch: a few long pieces per block; maximise (fill of the last round of
    // pieces) x (piece length vs ~1.5 tiles of fixed cost per segment)
    double best = -1.0;
    for (int cand = 1; cand <= std::min(tb.tiles, 16); ++cand) {
      const int tpp = (tb.tiles + cand - 1) / cand;
      const long long P = (long long)qblocks * ((tb.tiles + tpp - 1) / tpp);
      const long long rounds = (P + sm - 1) / sm;
      const double eff = double(P) / double(rounds * sm) * double(tpp) / (double(tpp) + 1.5);
      if (eff > best + 1e-9) { best = eff; S = cand; }
    }
  }
  for (;;) {
    S = std::max(1, std::min(S, tb.tiles));
    tb.tiles_per_piece = (tb.tiles + S - 1) / S;
    tb.pieces = (tb.tiles + tb.tiles_per_piece - 1) / tb.tiles_per_piece;   // drop pieces that would be empty
    tb.total = (long long)qblocks * tb.pieces;
    long long g = G > 0 ? std::min<long long>(G, tb.total) : std::min<long long>(sm, tb.total);
    if (G > 0 && (long long)qblocks * tb.pieces <= sm) g = tb.total;          // one piece per CTA
    // a block's `pieces` consecutive slots cross at most 1 + ceil((pieces - 1) / shortest range) ranges
    const long long lmin = tb.total / g;
    tb.slots = int(std::min<long long>(tb.pieces, 1 + (tb.pieces - 1 + lmin - 1) / lmin));
    tb.grid = int(g);
    if (S == 1 || rows * size_t(tb.slots) * 2 <= partial_cap) break;   // x2: column halves of the epilogue
    S = (S + 1) / 2;                 // coarser pieces, fewer segments per block
  }
}

// Build the batch, choose (R, split) and launch kernel 1.
// mirror: also store the match lists / counts into the mapped host buffers (host-API calls)
// latency: the call is a blocking / one-frame-at-a-time one (its own automatic-engine threshold,
// see tc_auto_min_cmp_latency).
static int run_knn(vsf_ctx* c, const std::vector<ProblemSpec>& specs, double ratio, bool mirror = false,
                   bool latency = false, const PoseLaunch* pose = nullptr) {
  if (specs.empty()) return VSF_OK;
  // VSF_TIMING: host time of a blocking call's stages (prepare, expansion launch, distance + finish launches)
  static const bool timing = std::getenv("VSF_TIMING") != nullptr;
  const auto tk0 = std::chrono::steady_clock::now();
  c->main_dirty = true;
  if (int(specs.size()) > kMaxProblems) return fail(c, VSF_ERR_CAPACITY, "too many problems in one batch");
  KnnBatch b;
  std::memset(&b, 0, sizeof(b));
  b.num_problems = int(specs.size());
  b.ratio = ratio;
  b.exact_second = c->want_second_index;
  b.knn_out = c->d_knn_out;
  b.partial = c->d_partial;
  b.qblock_arrivals = c->d_qblock_arrivals;
  b.qblock_pass = c->d_qblock_pass;
  b.problem_arrivals = c->d_problem_arrivals;
  if (mirror) {
    b.host_matches = c->mir_dm;
    b.host_counts = c->mir_dcounts;
    b.host_region_stride = c->rows_pad;
  }
  int row0 = 0, qb0 = 0, max_nq = 0, max_nt = 0;
  long long total_q = 0;
  for (int i = 0; i < b.num_problems; ++i) {
    const ProblemSpec& s = specs[i];
    if (s.nq > c->max_features || s.nt > c->max_features || s.nq < 0 || s.nt < 0)
      return fail(c, VSF_ERR_CAPACITY, "frame has more rows than max_features");
    KnnProblem& p = b.p[i];
    p.q = static_cast<const uint32_t*>(s.q);
    p.t = static_cast<const uint32_t*>(s.t);
    p.nq_dev = s.nq_dev;
    p.nt_dev = s.nt_dev;
    p.nq = s.nq;
    p.nt = s.nt;
    p.matches = s.out ? s.out : c->region_ptr(s.region);
    p.match_count = s.out_count ? s.out_count : c->count_base + s.region;
    p.host_count = s.host_count;
    p.row0 = row0;
    p.qb0 = qb0;
    p.region = s.region;
    row0 += round_up(std::max(s.nq, 1), 128);
    qb0 += round_up(std::max(s.nq, 1), 128) / 32;
    max_nq = std::max(max_nq, s.nq);
    max_nt = std::max(max_nt, s.nt);
    total_q += s.nq;
  }
  if (max_nq == 0) {
    // nothing to match: every problem reports zero survivors
    for (int i = 0; i < b.num_problems; ++i) {
      VSF_CUDA(c, cudaMemsetAsync(b.p[i].match_count, 0, sizeof(int), c->stream));
      if (mirror) c->mir_hcounts[specs[i].region] = 0;   // no kernel will write it
    }
    return VSF_OK;
  }
  // ---- engine choice.  The tensor-core engine needs 32-byte rows and at most kTcMaxTrains
  // distinct train frames in the batch; automatic mode uses it once the batch is large enough
  // to fill the machine (small batches are launch-bound and the POPC kernel has less fixed cost).
  int engine = c->engine;
  int train_of[kMaxProblems];
  int n_trains = 0;
  const ProblemSpec* train_spec[kTcMaxTrains];
  // (64-byte rows: int8 operands only, engine 3 falls back to the POPC kernel)
  bool tc_ok = (c->words == 8 || (c->words == 16 && c->engine != 3)) && max_nt >= 1;
  if (tc_ok) {
    for (int i = 0; i < b.num_problems && tc_ok; ++i) {
      train_of[i] = -1;
      if (specs[i].t_exp && specs[i].t_exp_int8 == (c->engine == 3 ? 0 : 1)) continue;   // expanded by the caller
      int k = 0;
      for (; k < n_trains; ++k)
        if (train_spec[k]->t == specs[i].t && train_spec[k]->nt == specs[i].nt && train_spec[k]->nt_dev == specs[i].nt_dev) break;
      if (k == n_trains) {
        if (n_trains == kTcMaxTrains) { tc_ok = false; break; }
        train_spec[n_trains++] = &specs[i];
      }
      train_of[i] = k;
    }
  }
  const double auto_min = latency ? c->tc_auto_min_cmp_latency : c->tc_auto_min_cmp;
  if (engine == 0) engine = (tc_ok && double(total_q) * double(max_nt) >= auto_min) ? 2 : 1;
  if (engine >= 2 && !tc_ok) engine = 1;
  c->last_engine = engine;

  if (engine >= 2) {
    const int int8 = engine == 2;
    const bool wide = c->words == 16;                // 64-byte rows: knn2_tc64_kernel.cu
    // 64-byte rows run on CTA pairs (cta_group::2, 256 queries x 256 train rows per MMA) once the
    // launch has work for every pair; engine flag 128 keeps the single-CTA kernel (A/B timing)
    const int sm_avail = std::max(1, c->sm_count - c->reserve_sms);
    bool pair = wide && !(c->engine_flags & 128) && sm_avail >= 2;
    if (pair) {
      long long slots = 0;
      for (const ProblemSpec& sp : specs) slots += (long long)((sp.nq + 255) / 256) * ((max_nt + 255) / 256);
      pair = slots >= sm_avail / 2;
    }
    const int unit_q = wide ? (pair ? 256 : 128) : kTcQ;            // queries per work unit
    const int tile_rows = wide ? (pair ? 256 : 128) : kTcTileRows;  // train rows per tile
    const int pdl = (c->engine_flags & 8) ? 0 : 1;   // flag 8: ordinary launches (A/B timing)
    if (c->profile) VSF_CUDA(c, cudaEventRecord(c->pev[0], c->stream));
    long long* kt = nullptr;
    if ((c->engine_flags & 32) && c->d_ktrace) {   // kernel-level timeline, one record per launch (ring)
      kt = c->d_ktrace + size_t(c->ktrace_n % kKtracePoses) * 16;
      ++c->ktrace_n;
    }
    b.ktrace = kt;
    const uint8_t* exp_image[kTcMaxTrains];
    const auto tk1 = std::chrono::steady_clock::now();
    for (int k = 0; k < n_trains; ++k) {
      exp_image[k] = c->d_train_exp[k];
      if (wide)
        VSF_CUDA(c, launch_expand_train64(train_spec[k]->t, train_spec[k]->nt, train_spec[k]->nt_dev,
                                          c->d_train_exp[k], pdl, c->stream));
      else
        VSF_CUDA(c, launch_expand_train(train_spec[k]->t, train_spec[k]->nt, train_spec[k]->nt_dev,
                                        c->d_train_exp[k], int8, pdl, c->stream, kt));
      ++c->launches;
    }
    const auto tk2 = std::chrono::steady_clock::now();
    TcBatch tb;
    std::memset(&tb, 0, sizeof(tb));
    // Work = (256-query block, piece of train tiles) slots, shared out to the CTAs as equal
    // contiguous ranges (see TcBatch).
    int qblocks = 0;
    for (int i = 0; i < b.num_problems; ++i) {
      tb.t_exp[i] = train_of[i] < 0 ? specs[i].t_exp : exp_image[train_of[i]];   // (< 0: expanded by the caller)
      tb.qb_begin[i] = qblocks;
      qblocks += (specs[i].nq + unit_q - 1) / unit_q;
    }
    tb.qb_begin[b.num_problems] = qblocks;
    tb.tiles = (max_nt + tile_rows - 1) / tile_rows;
    tb.flags = c->engine_flags;
    if (c->engine_flags & 16) {   // per-CTA timeline for tools/tc_timeline.py
      if (!c->d_tc_trace) VSF_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->d_tc_trace), size_t(c->sm_count) * kTcTraceSlots * sizeof(long long)));
      VSF_CUDA(c, cudaMemsetAsync(c->d_tc_trace, 0, size_t(c->sm_count) * kTcTraceSlots * sizeof(long long), c->stream));
      tb.trace = c->d_tc_trace;
    }
    // (64-byte rows store 4 partial key pairs per query and segment, 32-byte rows 2)
    tb.unit_q = unit_q;
    plan_tc_partition(&tb, qblocks, pair ? sm_avail / 2 : sm_avail, c->force_split,
                      size_t(row0) * (wide ? 2 : 1), c->partial_cap);
    b.split = tb.slots;
    {
      // refine + compaction as one kernel (engine flag 512: the two separate kernels, A/B timing)
      FinishArgs fa;
      std::memset(&fa, 0, sizeof(fa));
      // two launch states, used alternately: the CTAs of a pose's finish kernel may draw their
      // tickets while the previous pose's finish kernel is still running (early-start distance
      // kernels let their successor launch at once); their epochs never meet (1.., 2^40 + 1..),
      // so the two can share the look-back words
      fa.state = c->d_finish_ticket + 4 * (c->finish_parity & 1);
      fa.flags = c->d_finish_flags;
      const int phase = 0;
      if (pose) {
        tb.early = pose->early;
        tb.late = pose->late;
        if (pose->partial) b.partial = pose->partial;
        if (pose->flags) fa.flags = pose->flags;
        if (pose->state) fa.state = pose->state;
      }
      int launched = 0;
      FinishArgs* fap = (c->engine_flags & 512) ? nullptr : &fa;
      if (wide)
        VSF_CUDA(c, launch_knn2_tc64(b, tb, max_nq, pdl, c->profile ? c->pev + 1 : nullptr, c->stream, fap, &launched, phase));
      else
        VSF_CUDA(c, launch_knn2_tc(b, tb, int8, max_nq, pdl, c->profile ? c->pev + 1 : nullptr, c->stream, fap, &launched, phase));
      c->launches += launched;
      if (fap && !b.exact_second && !(pose && pose->state)) c->finish_parity ^= 1;
    }
    c->pev_valid = c->profile != 0;
    if (timing && latency) {
      const auto tk3 = std::chrono::steady_clock::now();
      auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::micro>(b - a).count();
      };
      std::fprintf(stderr, "  run_knn: prepare %.1f us, expansion launch(es) %.1f us, plan + distance + finish launches %.1f us\n",
                   us(tk0, tk1), us(tk1, tk2), us(tk2, tk3));
    }
    return VSF_OK;
  }

  // ---- POPC engine
  // Defaults from the round-1 sweep on B200 (profiles/): one query per thread keeps
  // 40 registers/thread and the most resident warps, which is what the carry-save
  // variant (5 POPC + 14 LOP3 per 256-bit comparison) needs to hide its longer
  // dependency chains; it beats R = 2, 4 at every size measured.
  const int R = c->force_R ? ((c->words == 16 && c->force_R > 2) ? 2 : c->force_R) : 1;
  const int mode = c->popc_mode >= 0 ? c->popc_mode : 2;
  const int variant = c->variant >= 0 ? c->variant : 3;
  long long qblocks = 0;
  for (const ProblemSpec& s : specs) qblocks += (s.nq + 32 * R - 1) / (32 * R);
  // Train splits: the work of one CTA is 32*R queries x (nt / S) train rows.  Aim for
  // at least ~8 waves of CTAs (5 resident per SM) so the tail of the launch is short,
  // but keep at least one full 256-row tile per split.
  int S = c->force_split;
  if (S == 0) {
    const long long target = 8LL * 5 * c->sm_count;
    S = int((target + qblocks - 1) / std::max(1LL, qblocks));
    S = std::min(S, std::max(1, max_nt / 256));
  }
  S = std::max(1, std::min(S, 32));
  while (S > 1 && size_t(row0) * S > c->partial_cap) --S;
  b.split = S;
  const int max_qblocks = (max_nq + 32 * R - 1) / (32 * R);
  if (c->profile) {
    VSF_CUDA(c, cudaEventRecord(c->pev[0], c->stream));
    VSF_CUDA(c, cudaEventRecord(c->pev[1], c->stream));
  }
  VSF_CUDA(c, launch_knn2(b, c->words, R, mode, variant, max_qblocks, c->stream));
  ++c->launches;
  if (c->profile) {
    for (int k = 2; k < 5; ++k) VSF_CUDA(c, cudaEventRecord(c->pev[k], c->stream));
    c->pev_valid = true;
  }
  return VSF_OK;
}

// Global scratch of the exact device sort (lists longer than its shared-memory capacity only).
// (two halves: the sorts of two consecutive frames of the pipelined path may run side by side)
static int sort_scratch(vsf_ctx* c, int exact, void** out, unsigned half = 0) {
  *out = nullptr;
  if (!exact) return VSF_OK;
  const size_t per = sort_exact_global_scratch(c->rows_pad) * size_t(std::max(c->window, 1));
  if (per == 0) return VSF_OK;
  if (!c->d_sort_scratch) VSF_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->d_sort_scratch), 2 * per));
  *out = c->d_sort_scratch + size_t(half & 1u) * per;
  return VSF_OK;
}

// ------------------------------------------------------------------------------------ context

extern "C" const char* vsf_version(void) { return VSF_VERSION_STRING; }

extern "C" const char* vsf_last_error(const vsf_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

extern "C" void vsf_destroy(vsf_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->own_stream) cudaStreamSynchronize(c->own_stream);
  if (!c->workers.empty()) {
    {
      std::lock_guard<std::mutex> lk(c->disp_mu);
      std::lock_guard<std::mutex> lk2(c->pool_mu);
      c->pool_stop = true;
    }
    c->disp_cv.notify_all();
    if (c->dispatcher.joinable()) c->dispatcher.join();   // drains its queue into the pool first
    c->pool_cv.notify_all();
    for (std::thread& t : c->workers) t.join();
  }
  std::free(c->h_keys);
  for (uint8_t* p : c->grp_exp)
    if (p && p != c->d_train_exp[0] && p != c->d_train_exp[1]) cudaFree(p);
  if (c->grp_flags) cudaFree(c->grp_flags);
  if (c->grp_matches) cudaFree(c->grp_matches);
  if (c->grp_counts) cudaFree(c->grp_counts);
  void* dev[] = {c->d_ring, c->d_raw_left, c->d_right_c, c->d_xy_left, c->d_xy_right,
                 c->d_xy_left_c, c->d_xy_right_c, c->d_knn_out, c->d_partial, c->d_partial2, c->d_finish_ticket, c->d_finish_flags, c->d_qblock_arrivals,
                 c->d_qblock_pass, c->d_problem_arrivals, c->d_matches, c->d_match_count, c->d_resid,
                 c->d_chunk_keep, c->d_chunk_off, c->d_ticket, c->d_kept_left, c->d_kept_right, c->d_slot_rows, c->d_thresh, c->d_X4,
                 c->d_tri_io, c->d_sink, c->d_fm, c->d_fm_count, c->d_train_exp[0], c->d_train_exp[1], c->d_tc_trace,
                 c->d_ktrace, c->d_sort_scratch};
  for (void* p : dev)
    if (p) cudaFree(p);
  void* host[] = {c->h_desc[0], c->h_xy[0], c->h_xy[1], c->h_counts, c->h_matches, c->h_region_counts,
                  c->h_kept[0], c->h_kept[1], c->h_resid, c->h_X4, c->h_knn, c->h_tri_io,
                  c->h_scalar, c->h_fm};
  for (void* p : host)
    if (p) cudaFreeHost(p);
  if (c->up_stream) cudaStreamSynchronize(c->up_stream);
  if (c->down_stream) cudaStreamSynchronize(c->down_stream);
  for (cudaStream_t st : c->sort_stream)
    if (st) cudaStreamSynchronize(st);
  if (c->exp_stream) {
    cudaStreamSynchronize(c->exp_stream);
    cudaStreamDestroy(c->exp_stream);
  }
  for (cudaEvent_t e : {c->ev_img[0], c->ev_img[1], c->ev_grp[0], c->ev_grp[1], c->ev_blk})
    if (e) cudaEventDestroy(e);
  for (vsf_ctx::Flight& f : c->flights) {
    void* fh[] = {f.h_desc, f.h_matches, f.h_fm, f.h_counts};
    for (void* p : fh)
      if (p) cudaFreeHost(p);
    if (f.d_matches) cudaFree(f.d_matches);
    if (f.d_train_exp) cudaFree(f.d_train_exp);
    if (f.d_fm) cudaFree(f.d_fm);
    if (f.d_counts) cudaFree(f.d_counts);
    for (cudaEvent_t e : {f.ev_up, f.ev_chain, f.done, f.ev_sort})
      if (e) cudaEventDestroy(e);
    std::free(f.keys);
  }
  if (c->obs_up_stream) cudaStreamSynchronize(c->obs_up_stream);
  for (vsf_ctx::ObsFlight& f : c->obs) {
    if (f.h_in) cudaFreeHost(f.h_in);
    if (f.h_out) cudaFreeHost(f.h_out);
    if (f.d_in) cudaFree(f.d_in);
    for (cudaEvent_t e : {f.ev_up, f.ev_done})
      if (e) cudaEventDestroy(e);
  }
  if (c->obs_up_stream) cudaStreamDestroy(c->obs_up_stream);
  if (c->ev_main) cudaEventDestroy(c->ev_main);
  if (c->up_stream) cudaStreamDestroy(c->up_stream);
  if (c->down_stream) cudaStreamDestroy(c->down_stream);
  for (cudaStream_t st : c->sort_stream)
    if (st) cudaStreamDestroy(st);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  for (cudaEvent_t e : c->pev)
    if (e) cudaEventDestroy(e);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

#define VSF_ALLOC(ctx, ptr, bytes)                                                      \
  do {                                                                                  \
    cudaError_t e__ = cudaMalloc(reinterpret_cast<void**>(&(ptr)), (bytes));            \
    if (e__ != cudaSuccess) {                                                           \
      std::fprintf(stderr, "vsf_create: cudaMalloc(%zu) failed: %s\n", size_t(bytes),   \
                   cudaGetErrorString(e__));                                            \
      vsf_destroy(ctx);                                                                 \
      return VSF_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)
#define VSF_ALLOC_HOST(ctx, ptr, bytes)                                                 \
  do {                                                                                  \
    cudaError_t e__ = cudaMallocHost(reinterpret_cast<void**>(&(ptr)), (bytes));        \
    if (e__ != cudaSuccess) {                                                           \
      std::fprintf(stderr, "vsf_create: cudaMallocHost(%zu) failed: %s\n",              \
                   size_t(bytes), cudaGetErrorString(e__));                             \
      vsf_destroy(ctx);                                                                 \
      return VSF_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

extern "C" int vsf_create(int device, int max_features, int desc_bytes, int window, vsf_ctx** out) {
  if (!out) return VSF_ERR_BAD_ARG;
  *out = nullptr;
  if (max_features < 1 || max_features > kMaxRows || desc_bytes < 1 || desc_bytes > 64 ||
      window < 1 || window > kMaxProblems - 2)
    return VSF_ERR_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    std::fprintf(stderr, "vsf_create: CUDA device %d not available (no CPU fallback exists)\n", device);
    return VSF_ERR_CUDA;
  }
  if (cudaSetDevice(device) != cudaSuccess) return VSF_ERR_CUDA;
  vsf_ctx* c = new vsf_ctx();
  c->device = device;
  c->max_features = max_features;
  c->desc_bytes = desc_bytes;
  c->row_bytes = desc_bytes <= 32 ? 32 : 64;
  c->words = c->row_bytes / 4;
  c->window = window;
  c->rows_pad = round_up(max_features, 128);
  c->regions = window + 2;
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  c->host_threads = int(std::max(1u, std::min(16u, std::thread::hardware_concurrency())));
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return VSF_ERR_CUDA;
  }
  c->stream = c->own_stream;
  cudaEventCreate(&c->ev0);
  cudaEventCreate(&c->ev1);
  for (cudaEvent_t& e : c->pev) cudaEventCreate(&e);

  const size_t N = size_t(c->rows_pad);
  const size_t rows_cap = size_t(c->regions) * N;
  // (+ kMaxPoseGroup - 1: vsf_window_run_sequence uploads a group of frames before it launches
  // their kernels; a slot must not be reused while a staged frame still has to read it)
  c->ring_slots = window + 2 + (kMaxPoseGroup - 1);
  VSF_ALLOC(c, c->d_ring, size_t(c->ring_slots) * N * c->row_bytes);
  // (one block: a pair call packs its train rows right behind its query rows and uploads both
  // with one copy)
  VSF_ALLOC(c, c->d_raw_left, 2 * N * c->row_bytes);
  c->d_raw_right = c->d_raw_left + N * c->row_bytes;
  VSF_ALLOC(c, c->d_right_c, N * c->row_bytes);
  VSF_ALLOC(c, c->d_xy_left, N * sizeof(float2));
  VSF_ALLOC(c, c->d_xy_right, N * sizeof(float2));
  VSF_ALLOC(c, c->d_xy_left_c, N * sizeof(float2));
  VSF_ALLOC(c, c->d_xy_right_c, N * sizeof(float2));
  VSF_ALLOC(c, c->d_knn_out, rows_cap * sizeof(uint4));
  c->partial_cap = rows_cap * 16 + 262144;
  VSF_ALLOC(c, c->d_partial, c->partial_cap * sizeof(uint2));
  const size_t qb_cap = rows_cap / 32 + kMaxProblems * 4;
  c->qb_cap = qb_cap;
  if (const char* e = std::getenv("VSF_POSE_GROUP")) c->pose_group = std::max(1, std::min(kMaxPoseGroup, std::atoi(e)));
  VSF_ALLOC(c, c->d_qblock_arrivals, qb_cap * sizeof(unsigned));
  VSF_ALLOC(c, c->d_qblock_pass, qb_cap * sizeof(unsigned));
  VSF_ALLOC(c, c->d_problem_arrivals, kMaxProblems * sizeof(unsigned));
  cudaMemset(c->d_qblock_arrivals, 0, qb_cap * sizeof(unsigned));
  cudaMemset(c->d_qblock_pass, 0, qb_cap * sizeof(unsigned));
  cudaMemset(c->d_problem_arrivals, 0, kMaxProblems * sizeof(unsigned));
  VSF_ALLOC(c, c->d_finish_ticket, 8 * sizeof(unsigned long long));
  VSF_ALLOC(c, c->d_finish_flags, (qb_cap + 8) * sizeof(unsigned long long));
  {
    const unsigned long long init[8] = {0ull, 0ull, 1ull, 0ull, 0ull, 0ull, (1ull << 40) + 1ull, 0ull};   // epoch 0 is what the zeroed look-back words hold
    cudaMemcpy(c->d_finish_ticket, init, sizeof(init), cudaMemcpyHostToDevice);
  }
  cudaMemset(c->d_finish_flags, 0, (qb_cap + 8) * sizeof(unsigned long long));   // epoch 0 is never used
  VSF_ALLOC(c, c->d_matches, rows_cap * sizeof(vsf_dmatch));
  c->match_base = c->d_matches;
  VSF_ALLOC(c, c->d_match_count, kMaxProblems * sizeof(int));
  cudaMemset(c->d_match_count, 0, kMaxProblems * sizeof(int));
  c->count_base = c->d_match_count;
  VSF_ALLOC(c, c->d_resid, N * sizeof(float));
  VSF_ALLOC(c, c->d_chunk_keep, (N / 256 + 2) * sizeof(unsigned));
  VSF_ALLOC(c, c->d_chunk_off, (N / 256 + 2) * sizeof(unsigned));
  VSF_ALLOC(c, c->d_ticket, sizeof(unsigned));
  cudaMemset(c->d_ticket, 0, sizeof(unsigned));
  VSF_ALLOC(c, c->d_kept_left, N * sizeof(int));
  VSF_ALLOC(c, c->d_kept_right, N * sizeof(int));
  VSF_ALLOC(c, c->d_slot_rows, c->ring_slots * sizeof(int));
  cudaMemset(c->d_slot_rows, 0, c->ring_slots * sizeof(int));
  VSF_ALLOC(c, c->d_thresh, 2 * sizeof(float));
  VSF_ALLOC(c, c->d_X4, N * sizeof(float4));
  VSF_ALLOC(c, c->d_tri_io, N * 8 * sizeof(float));
  VSF_ALLOC(c, c->d_sink, 64);
  VSF_ALLOC(c, c->d_fm, size_t(window) * N * sizeof(vsf_feature_match));
  VSF_ALLOC(c, c->d_fm_count, kMaxProblems * sizeof(int));
  for (int k = 0; k < kTcMaxTrains; ++k)   // +-1 images: one byte per descriptor bit
    VSF_ALLOC(c, c->d_train_exp[k], size_t(round_up(max_features, kTcTileRows)) * size_t(c->row_bytes) * 8);
  {
    const float init[2] = {10000.0f, 10000.0f};  // stereo_ambig_constraint (src/slam_frontend.cc:353)
    cudaMemcpy(c->d_thresh, init, sizeof(init), cudaMemcpyHostToDevice);
  }
  VSF_ALLOC_HOST(c, c->h_desc[0], 2 * N * c->row_bytes);
  c->h_desc[1] = c->h_desc[0] + N * c->row_bytes;
  for (int k = 0; k < 2; ++k) {
    VSF_ALLOC_HOST(c, c->h_xy[k], N * sizeof(float2));
    VSF_ALLOC_HOST(c, c->h_kept[k], N * sizeof(int));
  }
  VSF_ALLOC_HOST(c, c->h_counts, (kMaxProblems + 8) * sizeof(int));
  {
    // mapped pinned memory: the compaction tails store survivors straight into it (zero-copy)
    if (cudaHostAlloc(reinterpret_cast<void**>(&c->h_matches), rows_cap * sizeof(vsf_dmatch), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostAlloc(reinterpret_cast<void**>(&c->h_region_counts), kMaxProblems * sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->dm_matches), c->h_matches, 0) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->dm_region_counts), c->h_region_counts, 0) != cudaSuccess) {
      std::fprintf(stderr, "vsf_create: mapped pinned allocation failed\n");
      vsf_destroy(c);
      return VSF_ERR_CUDA;
    }
    std::memset(c->h_region_counts, 0, kMaxProblems * sizeof(int));
    c->mir_dm = c->dm_matches;
    c->mir_dcounts = c->dm_region_counts;
    c->mir_hcounts = c->h_region_counts;
  }
  VSF_ALLOC_HOST(c, c->h_resid, N * sizeof(float));
  VSF_ALLOC_HOST(c, c->h_X4, N * sizeof(float4));
  VSF_ALLOC_HOST(c, c->h_knn, N * sizeof(uint4));
  VSF_ALLOC_HOST(c, c->h_tri_io, N * 8 * sizeof(float));
  VSF_ALLOC_HOST(c, c->h_scalar, 16 * sizeof(float));
  VSF_ALLOC_HOST(c, c->h_fm, size_t(window) * N * sizeof(vsf_feature_match));
  if (const char* e = std::getenv("VSF_RESERVE_SMS")) c->reserve_override = std::atoi(e);
  if (const char* e = std::getenv("VSF_SORT_STREAMS")) c->sort_streams = std::atoi(e) == 1 ? 1 : 2;
  if (const char* e = std::getenv("VSF_ENGINE_FLAGS")) c->engine_flags = std::atoi(e) & (8 | 64 | 128 | 256 | 512 | 4096);   // A/B timing
  if (const char* e = std::getenv("VSF_ENGINE")) {   // test / bench override of the automatic choice
    const int v = std::atoi(e);
    if (v >= 0 && v <= 3 && (v < 3 || c->words == 8)) c->engine = v;
  }
  c->slot_count.assign(c->ring_slots, 0);
  c->slot_dev.assign(c->ring_slots, 0);
  c->slot_frame.assign(c->ring_slots, 0);
  c->slot_last_chain.assign(c->ring_slots, nullptr);
  reset_ring(c);
  // result buffers are read back up to a bound in places: start from zeros, not from whatever
  // the allocator handed out
  cudaMemset(c->d_knn_out, 0, rows_cap * sizeof(uint4));
  cudaMemset(c->d_matches, 0, rows_cap * sizeof(vsf_dmatch));
  cudaMemset(c->d_fm, 0, size_t(window) * N * sizeof(vsf_feature_match));
  cudaMemset(c->d_X4, 0, N * sizeof(float4));
  cudaMemset(c->d_resid, 0, N * sizeof(float));
  cudaMemset(c->d_kept_left, 0, N * sizeof(int));
  cudaMemset(c->d_kept_right, 0, N * sizeof(int));
  cudaMemset(c->d_tri_io, 0, N * 8 * sizeof(float));
  if (cudaDeviceSynchronize() != cudaSuccess) {
    vsf_destroy(c);
    return VSF_ERR_CUDA;
  }
  *out = c;
  return VSF_OK;
}

extern "C" int vsf_set_stream(vsf_ctx* c, void* cuda_stream) {
  if (!c) return VSF_ERR_BAD_ARG;
  cudaSetDevice(c->device);
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->up_stream) VSF_CUDA(c, cudaStreamSynchronize(c->up_stream));
  if (c->down_stream) VSF_CUDA(c, cudaStreamSynchronize(c->down_stream));
  for (cudaStream_t st : c->sort_stream)
    if (st) VSF_CUDA(c, cudaStreamSynchronize(st));
  if (c->obs_up_stream) VSF_CUDA(c, cudaStreamSynchronize(c->obs_up_stream));
  if (c->exp_stream) VSF_CUDA(c, cudaStreamSynchronize(c->exp_stream));
  c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
  return VSF_OK;
}

extern "C" int vsf_synchronize(vsf_ctx* c) {
  if (!c) return VSF_ERR_BAD_ARG;
  cudaSetDevice(c->device);
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->up_stream) VSF_CUDA(c, cudaStreamSynchronize(c->up_stream));
  if (c->down_stream) VSF_CUDA(c, cudaStreamSynchronize(c->down_stream));
  for (cudaStream_t st : c->sort_stream)
    if (st) VSF_CUDA(c, cudaStreamSynchronize(st));
  return VSF_OK;
}

extern "C" int vsf_set_tuning(vsf_ctx* c, int popc_mode, int train_split, int queries_per_thread,
                              int variant) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (!(popc_mode == -1 || popc_mode == 0 || popc_mode == 2 || popc_mode == 3))
    return fail(c, VSF_ERR_BAD_ARG, "popc_mode must be -1, 0, 2 or 3");
  if (train_split < 0 || train_split > 32) return fail(c, VSF_ERR_BAD_ARG, "train_split must be 0..32");
  if (!(queries_per_thread == 0 || queries_per_thread == 1 || queries_per_thread == 2 ||
        queries_per_thread == 4))
    return fail(c, VSF_ERR_BAD_ARG, "queries_per_thread must be 0, 1, 2 or 4");
  if (variant < -1 || variant > 3) return fail(c, VSF_ERR_BAD_ARG, "variant must be -1..3");
  c->popc_mode = popc_mode;
  c->force_split = train_split;
  c->force_R = queries_per_thread;
  c->variant = variant;
  return VSF_OK;
}

extern "C" int vsf_set_engine(vsf_ctx* c, int engine, int flags) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (engine < 0 || engine > 3) return fail(c, VSF_ERR_BAD_ARG, "engine must be 0 (auto), 1 (POPC), 2 (tensor int8) or 3 (tensor e4m3)");
  if (engine == 3 && c->words != 8) return fail(c, VSF_ERR_BAD_ARG, "the e4m3 tensor-core engine needs descriptors of at most 32 bytes (64-byte rows: engine 2)");
  c->engine = engine;
  c->engine_flags = flags;
  if (flags & 32) {
    // kernel-level timeline: starts = +inf, ends = 0
    cudaSetDevice(c->device);
    std::vector<long long> init(size_t(kKtracePoses) * 16);
    for (size_t i = 0; i < init.size(); ++i) init[i] = (i & 1) ? 0 : 0x7fffffffffffffffLL;
    if (!c->d_ktrace) VSF_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->d_ktrace), init.size() * sizeof(long long)));
    VSF_CUDA(c, cudaStreamSynchronize(c->stream));
    VSF_CUDA(c, cudaMemcpy(c->d_ktrace, init.data(), init.size() * sizeof(long long), cudaMemcpyHostToDevice));
    c->ktrace_n = 0;
  }
  return VSF_OK;
}

extern "C" int vsf_last_engine(const vsf_ctx* c) { return c ? c->last_engine : 0; }

extern "C" int vsf_set_profile(vsf_ctx* c, int enabled) {
  if (!c) return VSF_ERR_BAD_ARG;
  c->profile = enabled ? 1 : 0;
  c->pev_valid = false;
  return VSF_OK;
}

extern "C" int vsf_last_kernel_times(vsf_ctx* c, float* ms4) {
  if (!c || !ms4) return VSF_ERR_BAD_ARG;
  if (!c->pev_valid) return fail(c, VSF_ERR_STATE, "no profiled kNN launch (call vsf_set_profile(ctx, 1) first)");
  cudaSetDevice(c->device);
  VSF_CUDA(c, cudaEventSynchronize(c->pev[4]));
  for (int k = 0; k < 4; ++k) VSF_CUDA(c, cudaEventElapsedTime(&ms4[k], c->pev[k], c->pev[k + 1]));
  return VSF_OK;
}

extern "C" int vsf_device_sm_count(const vsf_ctx* c) { return c ? c->sm_count : 0; }
extern "C" int vsf_device_row_bytes(const vsf_ctx* c) { return c ? c->row_bytes : 0; }

// ------------------------------------------------------------------------- a1 / a2 (stateless)

static int check_pair(vsf_ctx* c, const uint8_t* q, int nq, size_t qs, const uint8_t* t, int nt, size_t ts) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (nq < 0 || nt < 0 || (nq > 0 && !q) || (nt > 0 && !t)) return fail(c, VSF_ERR_BAD_ARG, "bad descriptor arguments");
  if ((nq > 0 && qs < size_t(c->desc_bytes)) || (nt > 0 && ts < size_t(c->desc_bytes)))
    return fail(c, VSF_ERR_BAD_ARG, "row stride smaller than desc_bytes");
  if (nq > c->max_features || nt > c->max_features) return fail(c, VSF_ERR_CAPACITY, "more rows than max_features");
  return VSF_OK;
}

static int knn_pair(vsf_ctx* c, const uint8_t* q, int nq, size_t qs, const uint8_t* t, int nt, size_t ts,
                    double ratio, bool mirror = false) {
  cudaSetDevice(c->device);
  static const bool timing = std::getenv("VSF_TIMING") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  // one upload: the train rows are packed right behind the query rows (the left / right staging
  // and device buffers are one block each)
  const size_t qb = size_t(nq) * c->row_bytes, tb = size_t(nt) * c->row_bytes;
  if (nq > 0) pack_desc(c, c->h_desc[0], q, nq, qs);
  if (nt > 0) pack_desc(c, c->h_desc[0] + qb, t, nt, ts);
  if (qb + tb > 0) VSF_CUDA(c, cudaMemcpyAsync(c->d_raw_left, c->h_desc[0], qb + tb, cudaMemcpyHostToDevice, c->stream));
  if (timing) std::fprintf(stderr, "  knn_pair: staging + 1 cudaMemcpyAsync %.1f us\n",
                           std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
  std::vector<ProblemSpec> specs(1);
  specs[0] = ProblemSpec{c->d_raw_left, nq, nullptr, c->d_raw_left + qb, nt, nullptr, c->window + 1};
  return run_knn(c, specs, ratio, mirror, true);
}

extern "C" int vsf_knn2(vsf_ctx* c, const uint8_t* q, int nq, size_t q_stride, const uint8_t* t, int nt,
                        size_t t_stride, int32_t* idx, int32_t* dist) {
  int rc = check_pair(c, q, nq, q_stride, t, nt, t_stride);
  if (rc) return rc;
  if (nq == 0) return VSF_OK;
  if (!idx || !dist) return fail(c, VSF_ERR_BAD_ARG, "null output");
  c->want_second_index = 1;
  rc = knn_pair(c, q, nq, q_stride, t, nt, t_stride, 1.0);
  c->want_second_index = 0;
  if (rc) return rc;
  VSF_CUDA(c, cudaMemcpyAsync(c->h_knn, c->d_knn_out, size_t(nq) * sizeof(uint4), cudaMemcpyDeviceToHost, c->stream));
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < nq; ++i) {
    const uint4 r = c->h_knn[i];
    idx[2 * i] = int32_t(r.x);
    idx[2 * i + 1] = int32_t(r.y);
    dist[2 * i] = int32_t(r.z);
    dist[2 * i + 1] = int32_t(r.w);
  }
  return VSF_OK;
}

extern "C" int vsf_get_matches(vsf_ctx* c, const uint8_t* q, int nq, size_t q_stride, const uint8_t* t,
                               int nt, size_t t_stride, double ratio, vsf_dmatch* out, int cap, int* n_out) {
  int rc = check_pair(c, q, nq, q_stride, t, nt, t_stride);
  if (rc) return rc;
  if (!n_out || cap < 0 || (cap > 0 && !out)) return fail(c, VSF_ERR_BAD_ARG, "null output");
  *n_out = 0;
  if (nq == 0) return VSF_OK;
  static const bool timing = std::getenv("VSF_TIMING") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  if ((rc = knn_pair(c, q, nq, q_stride, t, nt, t_stride, ratio, true))) return rc;
  const auto t1 = std::chrono::steady_clock::now();
  const int region = c->window + 1;
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));   // survivors + count are already in mapped host memory
  const auto t2 = std::chrono::steady_clock::now();
  const int n = c->h_region_counts[region];
  if (n > cap) return fail(c, VSF_ERR_CAPACITY, "output capacity too small");
  if (n > 0) std::memcpy(out, c->h_matches + size_t(region) * c->rows_pad, size_t(n) * sizeof(vsf_dmatch));
  *n_out = n;
  if (timing) {
    const auto t3 = std::chrono::steady_clock::now();
    auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    std::fprintf(stderr, "vsf_get_matches: enqueue %.1f us, wait %.1f us, copy out %.1f us\n", us(t0, t1), us(t1, t2), us(t2, t3));
  }
  return VSF_OK;
}

// ----------------------------------------------------------------------------- a4 (the window)

extern "C" int vsf_window_size(const vsf_ctx* c) { return c ? int(c->live.size()) : 0; }

extern "C" int vsf_window_clear(vsf_ctx* c) {
  if (!c) return VSF_ERR_BAD_ARG;
  reset_ring(c);
  return VSF_OK;
}

extern "C" int vsf_window_push(vsf_ctx* c, uint64_t frame_id, const uint8_t* desc, int n, size_t stride) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (n < 0 || (n > 0 && !desc) || (n > 0 && stride < size_t(c->desc_bytes))) return fail(c, VSF_ERR_BAD_ARG, "bad frame");
  if (n > c->max_features) return fail(c, VSF_ERR_CAPACITY, "more rows than max_features");
  cudaSetDevice(c->device);
  int rc = upload_desc(c, 0, desc, n, stride, c->slot_ptr(c->staging_slot));
  if (rc) return rc;
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));  // pinned staging is reused by the next call
  commit_staging(c, frame_id, n);
  return VSF_OK;
}

extern "C" int vsf_window_commit(vsf_ctx* c, uint64_t frame_id, int n) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (n < 0 || n > c->max_features) return fail(c, VSF_ERR_BAD_ARG, "bad row count");
  commit_staging(c, frame_id, n);
  return VSF_OK;
}

static int window_launch(vsf_ctx* c, const uint8_t* desc, int n, size_t stride, double ratio, bool mirror,
                         uint8_t* h_staging = nullptr) {
  if (n < 0 || (n > 0 && !desc) || (n > 0 && stride < size_t(c->desc_bytes))) return fail(c, VSF_ERR_BAD_ARG, "bad frame");
  if (n > c->max_features) return fail(c, VSF_ERR_CAPACITY, "more rows than max_features");
  cudaSetDevice(c->device);
  int rc = upload_desc(c, h_staging ? h_staging : c->h_desc[0], desc, n, stride, c->slot_ptr(c->staging_slot));
  if (rc) return rc;
  std::vector<ProblemSpec> specs;
  int j = 0;
  for (int s : c->live)
    specs.push_back(ProblemSpec{c->slot_ptr(s), c->slot_count[s], c->slot_dev[s] ? c->d_slot_rows + s : nullptr,
                                c->slot_ptr(c->staging_slot), n, nullptr, j++});
  return run_knn(c, specs, ratio, mirror);
}

// host-API launches mirror their match lists into mapped host memory: one synchronisation and the
// counts (h_counts[j]) / lists (h_matches + region * rows_pad) are there
static int sync_mirrored(vsf_ctx* c, int n_regions, int first_region) {
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int j = 0; j < n_regions; ++j) c->h_counts[j] = c->h_region_counts[first_region + j];
  return VSF_OK;
}

// counts first, then exactly the survivors of every region
static int fetch_regions(vsf_ctx* c, int n_regions, int first_region) {
  if (n_regions == 0) return VSF_OK;
  VSF_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_match_count + first_region, n_regions * sizeof(int),
                              cudaMemcpyDeviceToHost, c->stream));
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int j = 0; j < n_regions; ++j) {
    const int n = c->h_counts[j];
    if (n > 0)
      VSF_CUDA(c, cudaMemcpyAsync(c->h_matches + size_t(first_region + j) * c->rows_pad,
                                  c->region_ptr(first_region + j), size_t(n) * sizeof(vsf_dmatch),
                                  cudaMemcpyDeviceToHost, c->stream));
  }
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));
  return VSF_OK;
}

extern "C" int vsf_window_match(vsf_ctx* c, const uint8_t* desc, int n, size_t stride, double ratio,
                                uint64_t* frame_ids, int* counts, vsf_dmatch* out, int cap_per_frame,
                                int* n_frames) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (!n_frames || !counts || cap_per_frame < 0 || (cap_per_frame > 0 && !out)) return fail(c, VSF_ERR_BAD_ARG, "null output");
  int rc = window_launch(c, desc, n, stride, ratio, true);
  if (rc) return rc;
  const int nf = int(c->live.size());
  if ((rc = sync_mirrored(c, nf, 0))) return rc;
  *n_frames = nf;
  c->last_h2d = size_t(n) * c->row_bytes;
  c->last_d2h = 0;
  for (int j = 0; j < nf; ++j) {
    const int cnt = c->h_counts[j];
    c->last_d2h += size_t(cnt) * sizeof(vsf_dmatch) + sizeof(int);
    if (cnt > cap_per_frame) return fail(c, VSF_ERR_CAPACITY, "cap_per_frame too small");
    counts[j] = cnt;
    if (frame_ids) frame_ids[j] = c->slot_frame[c->live[j]];
    if (cnt) std::memcpy(out + size_t(j) * cap_per_frame, c->h_matches + size_t(j) * c->rows_pad, size_t(cnt) * sizeof(vsf_dmatch));
  }
  return VSF_OK;
}

extern "C" int vsf_window_last_transfer(const vsf_ctx* c, size_t* h2d_bytes, size_t* d2h_bytes) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (h2d_bytes) *h2d_bytes = c->last_h2d;
  if (d2h_bytes) *d2h_bytes = c->last_d2h;
  return VSF_OK;
}

namespace {
// The tail of Frontend::GetFeatureMatches on the host (src/slam_frontend.cc:289-296):
// std::sort(matches) with cv::DMatch::operator< (distance only), keep the first
// int(size * best_percent), emit FeatureMatch(queryIdx, trainIdx).  std::sort is not stable, so
// the order inside equal-distance groups is whatever libstdc++'s introsort does; that sequence
// of moves depends only on the comparison results and the element count, never on the element
// type, so sorting 4-byte (distance << 22 | position) keys with a distance-only comparator
// yields exactly the permutation the reference gets on its 16-byte DMatch records, at a
// fraction of the memory traffic.  Only the part of the sort that can reach the first `keep`
// positions is carried out (exact_sort.h).  Returns the kept count.
int sort_cut_list(const vsf_dmatch* m, int n, float best_percent, uint32_t* keys, vsf_feature_match* out, int cap) {
  const int keep = int(float(size_t(n)) * best_percent);   // float multiply, truncation (:290)
  if (keep > cap) return -1;
  for (int i = 0; i < n; ++i) keys[i] = (uint32_t(int(m[i].distance)) << kIdxBits) | uint32_t(i);
  vsf_exact_sort::sort_prefix<kIdxBits>(keys, n, keep);
  for (int i = 0; i < keep; ++i) {
    const vsf_dmatch& d = m[keys[i] & kIdxMask];
    out[i].feature_idx_initial = uint64_t(d.queryIdx);
    out[i].feature_idx_current = uint64_t(d.trainIdx);
  }
  return keep;
}
}  // namespace

extern "C" int vsf_window_feature_matches(vsf_ctx* c, const uint8_t* desc, int n, size_t stride, double ratio,
                                          float best_percent, int sort_mode, uint64_t* frame_ids, int* counts,
                                          vsf_feature_match* out, int cap_per_frame, int* n_frames) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (!n_frames || !counts || cap_per_frame < 0 || (cap_per_frame > 0 && !out)) return fail(c, VSF_ERR_BAD_ARG, "null output");
  if (sort_mode < 0 || sort_mode > 3) return fail(c, VSF_ERR_BAD_ARG, "sort_mode must be 0 .. 3");
  if (sort_mode == VSF_SORT_EXACT_AUTO)
    sort_mode = (c->host_threads >= 8 || c->rows_pad > sort_exact_max_rows()) ? VSF_SORT_EXACT_HOST : VSF_SORT_EXACT_DEVICE;
  if (sort_mode == 2 && c->rows_pad > sort_exact_max_rows())
    return fail(c, VSF_ERR_CAPACITY, "sort_mode 2 (reference order on the device) needs max_features <= 24576; use sort_mode 1");
  int rc = window_launch(c, desc, n, stride, ratio, sort_mode == 1);
  if (rc) return rc;
  const int nf = int(c->live.size());
  *n_frames = nf;
  if (sort_mode != 1) {
    // device: stable counting sort by distance + best_percent cut, only survivors come back
    std::vector<const vsf_dmatch*> mp(nf);
    std::vector<const int*> cp(nf);
    int max_matches = 0;
    for (int j = 0; j < nf; ++j) {
      mp[j] = c->region_ptr(j);
      cp[j] = c->d_match_count + j;
      max_matches = std::max(max_matches, c->slot_count[c->live[j]]);
    }
    if (nf > 0) {
      void* gs = nullptr;
      if ((rc = sort_scratch(c, sort_mode == 2, &gs))) return rc;
      // (the scratch is sized for rows_pad: pass that bound so the kernel's strides match)
      VSF_CUDA(c, launch_sort_cut(mp.data(), cp.data(), nf, best_percent, c->d_fm, c->rows_pad,
                                  c->d_fm_count, 8 * c->row_bytes + 1, gs ? c->rows_pad : max_matches, sort_mode == 2,
                                  c->sort_depth_override, gs, c->stream));
      VSF_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_fm_count, nf * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      VSF_CUDA(c, cudaStreamSynchronize(c->stream));
      for (int j = 0; j < nf; ++j)
        if (c->h_counts[j] > 0)
          VSF_CUDA(c, cudaMemcpyAsync(c->h_fm + size_t(j) * c->rows_pad, c->d_fm + size_t(j) * c->rows_pad,
                                      size_t(c->h_counts[j]) * sizeof(vsf_feature_match),
                                      cudaMemcpyDeviceToHost, c->stream));
      VSF_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    c->last_h2d = size_t(n) * c->row_bytes;
    c->last_d2h = 0;
    for (int j = 0; j < nf; ++j) {
      const int cnt = c->h_counts[j];
      c->last_d2h += size_t(cnt) * sizeof(vsf_feature_match) + sizeof(int);
      if (cnt > cap_per_frame) return fail(c, VSF_ERR_CAPACITY, "cap_per_frame too small");
      counts[j] = cnt;
      if (frame_ids) frame_ids[j] = c->slot_frame[c->live[j]];
      if (cnt) std::memcpy(out + size_t(j) * cap_per_frame, c->h_fm + size_t(j) * c->rows_pad, size_t(cnt) * sizeof(vsf_feature_match));
    }
    return VSF_OK;
  }
  // sort_mode 1: the reference's own host sequence (src/slam_frontend.cc:289-296), one
  // frame pair per host thread (each list is sorted by the same single-threaded
  // std::sort the reference runs, so the order inside every list is unchanged).
  if ((rc = sync_mirrored(c, nf, 0))) return rc;
  c->last_h2d = size_t(n) * c->row_bytes;
  c->last_d2h = 0;
  for (int j = 0; j < nf; ++j) {
    c->last_d2h += size_t(c->h_counts[j]) * sizeof(vsf_dmatch) + sizeof(int);   // the whole survivor list is mirrored
    const int good = int(float(size_t(c->h_counts[j])) * best_percent);
    if (good > cap_per_frame) return fail(c, VSF_ERR_CAPACITY, "cap_per_frame too small");
    counts[j] = good;
    if (frame_ids) frame_ids[j] = c->slot_frame[c->live[j]];
  }
  if (!c->h_keys) c->h_keys = static_cast<uint32_t*>(std::malloc(size_t(c->window) * c->rows_pad * sizeof(uint32_t)));
  if (!c->h_keys) return fail(c, VSF_ERR_CUDA, "out of host memory");
  const int nthreads = std::max(1, std::min(nf, c->host_threads));
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int j = 0; j < nf; ++j)
    sort_cut_list(c->h_matches + size_t(j) * c->rows_pad, c->h_counts[j], best_percent,
                  c->h_keys + size_t(j) * c->rows_pad, out + size_t(j) * cap_per_frame, cap_per_frame);
  return VSF_OK;
}

// ------------------------------------------------- pipelined window matching (a3/a4, frame stream)

static void pool_finish_one(vsf_ctx* c, vsf_ctx::Flight* f) {
  if (f->pending.fetch_sub(1, std::memory_order_acq_rel) == 1) {
    std::lock_guard<std::mutex> lk(c->done_mu);
    c->done_cv.notify_all();
  }
}

static void pool_worker(vsf_ctx* c) {
  for (;;) {
    vsf_ctx::SortTask t;
    {
      std::unique_lock<std::mutex> lk(c->pool_mu);
      c->pool_cv.wait(lk, [c] { return c->pool_stop || !c->pool_q.empty(); });
      if (c->pool_q.empty()) return;   // stop requested and nothing left
      t = c->pool_q.front();
      c->pool_q.pop_front();
    }
    vsf_ctx::Flight* f = t.f;
    const size_t off = size_t(t.list) * c->rows_pad;
    f->keep[t.list] = sort_cut_list(f->h_matches + off, f->h_counts[t.list], f->best_percent, f->keys + off,
                                    f->h_fm + off, c->rows_pad);
    pool_finish_one(c, f);
  }
}

static void pool_dispatcher(vsf_ctx* c) {
  cudaSetDevice(c->device);
#ifdef __linux__
  prctl(PR_SET_TIMERSLACK, 1UL, 0UL, 0UL, 0UL);   // this thread's short sleeps should be short
#endif
  for (;;) {
    vsf_ctx::Flight* f;
    {
      std::unique_lock<std::mutex> lk(c->disp_mu);
      c->disp_cv.wait(lk, [c] { return c->pool_stop || !c->disp_q.empty(); });
      if (c->disp_q.empty()) return;
      f = c->disp_q.front();
      c->disp_q.pop_front();
    }
    // once the event completes the survivors + counts of the frame are in host memory; every
    // past frame's list then becomes one task, longest first (they bound the finish time)
    // Poll with short sleeps instead of cudaEventSynchronize: that call spins on a core the sort
    // workers can use (a blocking-sync event frees the core too, but its wake-up latency under
    // load cost 20 us per frame on a 16-core host); the few tens of microseconds a sleep may
    // overshoot are hidden by the frames in flight.
    cudaError_t e;
    if (c->dispatch_spin) {
      e = cudaEventSynchronize(f->done);
    } else {
      while ((e = cudaEventQuery(f->done)) == cudaErrorNotReady)
        std::this_thread::sleep_for(std::chrono::microseconds(10));
    }
    if (e != cudaSuccess) f->cuda_error = int(e);
    const int nf = (e == cudaSuccess) ? f->nf : 0;
    if (nf > 0) {
      int order[kMaxProblems];
      for (int j = 0; j < nf; ++j) order[j] = j;
      std::sort(order, order + nf, [f](int a, int b) { return f->h_counts[a] > f->h_counts[b]; });
      f->pending.fetch_add(nf, std::memory_order_acq_rel);
      {
        std::lock_guard<std::mutex> lk(c->pool_mu);
        for (int k = 0; k < nf; ++k) c->pool_q.push_back(vsf_ctx::SortTask{f, order[k]});
      }
      c->pool_cv.notify_all();
    }
    pool_finish_one(c, f);   // the device-wait token taken at submit
  }
}

static int flights_init(vsf_ctx* c) {
  if (c->flights_ready) return VSF_OK;
  cudaSetDevice(c->device);
  const size_t N = size_t(c->rows_pad);
  const size_t list_bytes = size_t(c->window) * N * 16;   // vsf_dmatch and vsf_feature_match: 16 B
  VSF_CUDA(c, cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
  VSF_CUDA(c, cudaStreamCreateWithFlags(&c->down_stream, cudaStreamNonBlocking));
  for (cudaStream_t& st : c->sort_stream) VSF_CUDA(c, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  VSF_CUDA(c, cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
  for (vsf_ctx::Flight& f : c->flights) {
    VSF_CUDA(c, cudaEventCreateWithFlags(&f.ev_up, cudaEventDisableTiming));
    VSF_CUDA(c, cudaEventCreateWithFlags(&f.ev_chain, cudaEventDisableTiming));
    VSF_CUDA(c, cudaEventCreateWithFlags(&f.done, cudaEventDisableTiming));
    VSF_CUDA(c, cudaEventCreateWithFlags(&f.ev_sort, cudaEventDisableTiming));
    VSF_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&f.d_counts), kMaxProblems * sizeof(int)));
    VSF_CUDA(c, cudaMemset(f.d_counts, 0, kMaxProblems * sizeof(int)));
    VSF_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&f.h_desc), N * c->row_bytes));
    VSF_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&f.d_matches), list_bytes));
    VSF_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&f.d_train_exp), size_t(round_up(c->max_features, kTcTileRows)) * size_t(c->row_bytes) * 8));
    VSF_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&f.d_fm), list_bytes));
    // lists are downloaded up to their bound, not their (device-side) length: start from zeros
    VSF_CUDA(c, cudaMemset(f.d_matches, 0, list_bytes));
    VSF_CUDA(c, cudaMemset(f.d_fm, 0, list_bytes));
    VSF_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&f.h_matches), list_bytes));
    VSF_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&f.h_fm), list_bytes));
    VSF_CUDA(c, cudaHostAlloc(reinterpret_cast<void**>(&f.h_counts), kMaxProblems * sizeof(int), cudaHostAllocMapped));
    VSF_CUDA(c, cudaHostGetDevicePointer(reinterpret_cast<void**>(&f.dm_counts), f.h_counts, 0));
    std::memset(f.h_counts, 0, kMaxProblems * sizeof(int));
    f.keys = static_cast<uint32_t*>(std::malloc(size_t(c->window) * N * sizeof(uint32_t)));
    if (!f.keys) return fail(c, VSF_ERR_CUDA, "out of host memory");
  }
  // worker threads for the host-side finish of sort_mode 1: lists of different frames in flight
  // are sorted concurrently, so the host keeps up with the device
  // The dispatcher sleeps between frames; the caller's thread and the CUDA driver's own threads
  // need about two cores (measured on a 16-core host: 13 or 14 workers 58.4 us/pose, 15 workers
  // 64.3), which a small share of a many-GPU host cannot spare (4 threads: 3 workers 158 us/pose,
  // 2 workers 235).
  int nworkers = c->host_threads >= 8 ? c->host_threads - 2 : std::max(1, c->host_threads - 1);
  if (const char* e = std::getenv("VSF_SORT_WORKERS")) nworkers = std::max(1, std::atoi(e));
  if (const char* e = std::getenv("VSF_DISPATCH_SPIN")) c->dispatch_spin = std::atoi(e) != 0;
  for (int i = 0; i < nworkers; ++i) c->workers.emplace_back(pool_worker, c);
  c->dispatcher = std::thread(pool_dispatcher, c);
  c->flights_ready = true;
  return VSF_OK;
}

extern "C" int vsf_window_in_flight(const vsf_ctx* c) { return c ? c->flight_count : 0; }

// State of the finish kernel of a batch that spans a group of poses / frames: look-back words
// for kMaxProblems problems, two launch states used alternately (FinishArgs::state; epochs 1.. and
// 2^40 + 1.. never meet, the zeroed look-back words hold epoch 0), survivor lists of the poses
// that do not write the ctx's own.  Allocated on first use.
static size_t group_flag_words(const vsf_ctx* c) { return size_t(kMaxProblems) * (size_t(c->rows_pad) / 32 + 4) + 8; }
static int group_state_init(vsf_ctx* c) {
  if (c->grp_ready) return VSF_OK;
  // (a call that failed half way is completed by the next one: every piece is made only once)
  const size_t flag_words = group_flag_words(c);
  const size_t rows_cap = size_t(c->regions) * size_t(c->rows_pad);
  if (!c->grp_flags) {
    unsigned long long* p = nullptr;
    VSF_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&p), (flag_words + 8) * sizeof(unsigned long long)));
    const unsigned long long init[8] = {0ull, 0ull, 1ull, 0ull, 0ull, 0ull, (1ull << 40) + 1ull, 0ull};
    cudaError_t e = cudaMemset(p, 0, (flag_words + 8) * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemcpy(p + flag_words, init, sizeof(init), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(p);
      VSF_CUDA(c, e);
    }
    c->grp_flags = p;
  }
  if (!c->grp_matches)
    VSF_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->grp_matches), size_t(kMaxPoseGroup - 1) * rows_cap * sizeof(vsf_dmatch)));
  if (!c->grp_counts)
    VSF_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->grp_counts), size_t(kMaxPoseGroup - 1) * kMaxProblems * sizeof(int)));
  if (!c->d_partial2) VSF_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->d_partial2), c->partial_cap * sizeof(uint2)));
  if (!c->exp_stream) VSF_CUDA(c, cudaStreamCreateWithFlags(&c->exp_stream, cudaStreamNonBlocking));
  for (cudaEvent_t* e : {&c->ev_img[0], &c->ev_img[1], &c->ev_grp[0], &c->ev_grp[1], &c->ev_blk})
    if (!*e) VSF_CUDA(c, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  c->grp_ready = true;
  return VSF_OK;
}

// Launch the kernels of the staged frames (vsf_window_submit stages one frame; ordinarily it is
// flushed at once, vsf_window_run_sequence lets c->defer_group frames accumulate).  On the tensor
// engine a group is launched like a group of poses of vsf_window_match_block_device: ONE batch of
// all its frame pairs - one distance kernel, one finish kernel - whose problems write straight
// into the frames' own buffers; then the device sorts, one event for the whole group.
static int flush_flights(vsf_ctx* c) {
  const int m = c->flights_staged;
  if (m == 0) return VSF_OK;
  cudaSetDevice(c->device);
  vsf_ctx::Flight* fl[VSF_PIPELINE_DEPTH];
  for (int g = 0; g < m; ++g)
    fl[g] = &c->flights[(c->flight_head + c->flight_count - m + g) % VSF_PIPELINE_DEPTH];
  c->flights_staged = 0;
  // the uploads (and expansions) of the whole group precede the last frame's event on the upload stream
  VSF_CUDA(c, cudaStreamWaitEvent(c->stream, fl[m - 1]->ev_up, 0));
  int rc = VSF_OK;
  // one batch for the group: every frame has work for the tensor engine
  bool grouped = m > 1 && (c->words == 8 || (c->words == 16 && c->engine != 3)) && c->engine != 1 &&
                 !(c->engine_flags & (8 | 512)) && !c->profile;
  bool any_side_sort = false;
  for (int g = 0; g < m; ++g) {
    const vsf_ctx::Flight& f = *fl[g];
    if (f.nf == 0 || f.n == 0 || f.max_cnt == 0 || !f.specs[0].t_exp || f.ratio != fl[0]->ratio) grouped = false;
    if (f.sort_mode != 1 && f.nf > 0 && !(c->engine_flags & 64)) any_side_sort = true;
  }
  if (grouped && c->engine == 0) {
    // the automatic choice must come out as the tensor engine (run_knn decides on sum(nq) * max(nt))
    long long tq = 0;
    int mt = 0;
    for (int g = 0; g < m; ++g) {
      for (const ProblemSpec& sp : fl[g]->specs) tq += sp.nq;
      mt = std::max(mt, fl[g]->n);
    }
    if (double(tq) * double(mt) < c->tc_auto_min_cmp) grouped = false;
  }
  if (grouped && (rc = group_state_init(c))) return rc;
  // ---- main stream: every resident past frame (query side) against the frame (train side); the
  // survivors go to the flight's own device buffer, the counts straight to mapped host memory.
  // A device sort (modes 0, 2) runs on its own stream while the main stream goes on with the next
  // frames; their persistent distance kernels then leave a few SMs to it (one sort CTA per
  // list), which costs the distance kernel a few per cent and takes the sort off the frame
  // stream's critical path.  Measured on C4 (10 lists, 148 SMs): the 10 us stable sort is best
  // served by 8 SMs (57.2 us/pose; 58.7 with 10), the 55 us exact sort - whose CTAs of two
  // consecutive frames overlap - by 14-16 (62.7; 64.4 with 10, 68.0 with 8) while a pose took
  // 55 us; now that a pose takes 44 us the sorts of two frames have to run fully side by side:
  // 20 SMs (46.4 us/frame with 15, 44.6 with 20, 45.5 with 24, 48.0 with 30)
  auto frame_state = [&](vsf_ctx::Flight& f) {
    const bool side_sort = f.sort_mode != 1 && f.nf > 0 && !(c->engine_flags & 64);
    c->reserve_sms = !side_sort ? 0
                     : f.sort_mode == 2 ? std::min(2 * f.nf, std::max(1, c->sm_count / 7))
                                        : std::min(f.nf, std::max(1, c->sm_count / 18));
    if (side_sort && c->reserve_override >= 0) c->reserve_sms = std::min(c->reserve_override, c->sm_count - 1);
    c->match_base = f.d_matches;
    c->count_base = f.d_counts;
    c->mir_dm = nullptr;
    c->mir_dcounts = f.dm_counts;
    c->mir_hcounts = f.h_counts;
  };
  auto restore_state = [&]() {
    c->reserve_sms = 0;
    c->match_base = c->d_matches;
    c->count_base = c->d_match_count;
    c->mir_dm = c->dm_matches;
    c->mir_dcounts = c->dm_region_counts;
    c->mir_hcounts = c->h_region_counts;
  };
  if (grouped) {
    // chunks of at most kMaxProblems frame pairs; the first distance kernel follows the event
    // wait, a later chunk's starts early (its inputs were complete before the group started, its
    // partial keys go to the other buffer)
    std::vector<ProblemSpec> specs;
    int chunk = 0;
    for (int g0 = 0; g0 < m && !rc;) {
      specs.clear();
      int g1 = g0;
      while (g1 < m && int(specs.size()) + fl[g1]->nf <= kMaxProblems) {
        vsf_ctx::Flight& f = *fl[g1];
        for (int j = 0; j < f.nf; ++j) {
          ProblemSpec sp = f.specs[j];
          sp.out = f.d_matches + size_t(j) * c->rows_pad;
          sp.out_count = f.d_counts + j;
          sp.host_count = f.sort_mode == 1 ? f.dm_counts + j : nullptr;
          specs.push_back(sp);
        }
        ++g1;
      }
      frame_state(*fl[g0]);   // (the sort reserve of the group's frames; the buffers come from the specs)
      PoseLaunch pl;
      pl.early = chunk > 0 ? 1 : 0;
      pl.late = (c->engine_flags & 4096) ? 0 : 1;   // (the next frames' expansion kernels on the upload stream need the room beside the distance CTAs; flag 4096: A/B timing)
      pl.partial = (chunk & 1) ? c->d_partial2 : c->d_partial;
      pl.flags = c->grp_flags;
      pl.state = c->grp_flags + group_flag_words(c) + 4 * (chunk & 1);
      rc = run_knn(c, specs, fl[g0]->ratio, false, false, &pl);
      ++chunk;
      g0 = g1;
    }
  } else {
    // frame by frame (vsf_window_submit alone; small frames): the finish kernel is let in late here
    // too, the next frame's expansion on the upload stream wants the room beside the distance CTAs
    PoseLaunch pl;
    pl.late = (c->engine_flags & 4096) ? 0 : 1;   // (flag 4096: trigger at the start, A/B timing)
    for (int g = 0; g < m && !rc; ++g) {
      vsf_ctx::Flight& f = *fl[g];
      frame_state(f);
      rc = run_knn(c, f.specs, f.ratio, f.sort_mode == 1, false, &pl);
    }
  }
  // ---- device sort (stable, or the replay of the reference's std::sort) + cut; the kept counts
  // go to the flights' mapped counters
  if (!rc && any_side_sort) {
    cudaError_t e = cudaEventRecord(c->ev_main, c->stream);
    if (e != cudaSuccess) {
      c->err = std::string("cudaEventRecord: ") + cudaGetErrorString(e);
      rc = VSF_ERR_CUDA;
    }
  }
  for (int g = 0; g < m && !rc; ++g) {
    vsf_ctx::Flight& f = *fl[g];
    if (f.sort_mode == 1 || f.nf == 0) continue;
    frame_state(f);
    const bool side_sort = !(c->engine_flags & 64);
    const unsigned half = c->sort_streams > 1 ? ((c->sort_rr++) & 1u) : 0u;
    cudaStream_t side = c->sort_stream[half];
    const vsf_dmatch* mp[kMaxProblems];
    const int* cp[kMaxProblems];
    for (int j = 0; j < f.nf; ++j) {
      mp[j] = c->region_ptr(j);
      cp[j] = c->count_base + j;
    }
    void* gs = nullptr;
    if ((rc = sort_scratch(c, f.sort_mode == 2, &gs, half)) == VSF_OK) {
      cudaError_t e = cudaSuccess;
      if (side_sort) e = cudaStreamWaitEvent(side, c->ev_main, 0);
      if (e == cudaSuccess)
        e = launch_sort_cut(mp, cp, f.nf, f.best_percent, f.d_fm, c->rows_pad, f.dm_counts, 8 * c->row_bytes + 1,
                            gs ? c->rows_pad : f.max_cnt, f.sort_mode == 2, c->sort_depth_override, gs,
                            side_sort ? side : c->stream);
      if (e == cudaSuccess && side_sort) e = cudaEventRecord(f.ev_sort, side);
      if (e != cudaSuccess) {
        c->err = std::string("launch_sort_cut: ") + cudaGetErrorString(e);
        rc = VSF_ERR_CUDA;
      }
    }
  }
  restore_state();
  if (rc) {
    // nothing of the group can be collected: drop its flights (the frames stay in the window; a
    // failed launch leaves the context unusable for anything but vsf_destroy anyway)
    c->flight_count -= m;
    return rc;
  }
  c->main_dirty = false;
  // one event behind the kernels of the whole group
  cudaEvent_t chain = fl[m - 1]->ev_chain;
  VSF_CUDA(c, cudaEventRecord(chain, c->stream));
  for (int g = 0; g < m; ++g) {
    vsf_ctx::Flight& f = *fl[g];
    f.chain = chain;
    f.launched = true;
    f.t_flush = std::chrono::steady_clock::now();
    static const bool trace_events = std::getenv("VSF_TIMING2") != nullptr;   // host time of every launch / collect
    if (g == 0 && trace_events)
      std::fprintf(stderr, "FLUSH %.1f frames %llu..+%d\n", std::chrono::duration<double, std::micro>(f.t_flush.time_since_epoch()).count(),
                   (unsigned long long)f.frame_id, m);
    for (const ProblemSpec& sp : f.specs) {
      const int slot = int((static_cast<const uint8_t*>(sp.q) - c->d_ring) / (size_t(c->rows_pad) * c->row_bytes));
      c->slot_last_chain[slot] = chain;
    }
    c->slot_last_chain[f.slot] = chain;
  }
  // ---- download stream: the lists leave through the copy engine while the main stream goes on
  // with the next frames.  Their lengths are only known on the device, so each list is copied up
  // to its bound (a past frame's row count, cut by best_percent for the sorted lists).
  for (int g = 0; g < m; ++g) {
    vsf_ctx::Flight& f = *fl[g];
    const bool side_sort = f.sort_mode != 1 && f.nf > 0 && !(c->engine_flags & 64);
    VSF_CUDA(c, cudaStreamWaitEvent(c->down_stream, side_sort ? f.ev_sort : chain, 0));
    f.d2h_bytes = size_t(f.nf) * sizeof(int);
    if (f.nf > 0 && f.max_cnt > 0) {
      const size_t pitch = size_t(c->rows_pad) * 16;
      const int rows = f.sort_mode == 1 ? f.max_cnt
                                        : std::min(f.max_cnt, int(float(size_t(f.max_cnt)) * f.best_percent) + 1);   // modes 0, 2: the kept part
      if (rows > 0) {
        VSF_CUDA(c, cudaMemcpy2DAsync(f.sort_mode == 1 ? static_cast<void*>(f.h_matches) : static_cast<void*>(f.h_fm), pitch,
                                      f.sort_mode == 1 ? static_cast<const void*>(f.d_matches) : static_cast<const void*>(f.d_fm), pitch,
                                      size_t(rows) * 16, size_t(f.nf), cudaMemcpyDeviceToHost, c->down_stream));
        f.d2h_bytes += size_t(rows) * 16 * size_t(f.nf);
      }
    }
    VSF_CUDA(c, cudaEventRecord(f.done, c->down_stream));
    f.cuda_error = 0;
    if (f.sort_mode == 1) {
      f.pending.store(1, std::memory_order_release);   // the device-wait token
      {
        std::lock_guard<std::mutex> lk(c->disp_mu);
        c->disp_q.push_back(&f);
      }
      c->disp_cv.notify_one();
    }
  }
  return VSF_OK;
}

extern "C" int vsf_window_submit(vsf_ctx* c, uint64_t frame_id, const uint8_t* desc, int n, size_t stride,
                                 double ratio, float best_percent, int sort_mode, int flags) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (sort_mode < 0 || sort_mode > 3) return fail(c, VSF_ERR_BAD_ARG, "sort_mode must be 0 .. 3");
  // (a stream of frames at 44 us per C4 pose keeps ~6 sort workers busy: the host order pays from
  // 12 threads - 10 workers -; measured 53.2 us per pose on the host against 49.2 on the device with 8)
  if (sort_mode == VSF_SORT_EXACT_AUTO)
    sort_mode = (c->host_threads >= 12 || c->rows_pad > sort_exact_max_rows()) ? VSF_SORT_EXACT_HOST : VSF_SORT_EXACT_DEVICE;
  if (sort_mode == 2 && c->rows_pad > sort_exact_max_rows())
    return fail(c, VSF_ERR_CAPACITY, "sort_mode 2 (reference order on the device) needs max_features <= 24576; use sort_mode 1");
  if (n < 0 || (n > 0 && !desc) || (n > 0 && stride < size_t(c->desc_bytes))) return fail(c, VSF_ERR_BAD_ARG, "bad frame");
  if (n > c->max_features) return fail(c, VSF_ERR_CAPACITY, "more rows than max_features");
  if (c->flight_count >= VSF_PIPELINE_DEPTH)
    return fail(c, VSF_ERR_STATE, "VSF_PIPELINE_DEPTH frames already in flight: call vsf_window_collect first");
  int rc = flights_init(c);
  if (rc) return rc;
  cudaSetDevice(c->device);
  vsf_ctx::Flight& f = c->flights[(c->flight_head + c->flight_count) % VSF_PIPELINE_DEPTH];
  const int nf = int(c->live.size());
  f.nf = nf;
  f.sort_mode = sort_mode;
  f.best_percent = best_percent;
  f.frame_id = frame_id;
  f.h2d_bytes = size_t(n) * c->row_bytes;
  f.launched = false;
  f.ratio = ratio;
  f.n = n;
  int max_cnt = 0;
  for (int j = 0; j < nf; ++j) {
    f.fids[j] = c->slot_frame[c->live[j]];
    max_cnt = std::max(max_cnt, c->slot_count[c->live[j]]);
  }
  f.max_cnt = max_cnt;
  // ---- upload stream: the frame's rows go into the free ring slot while the main stream is
  // still matching the previous frames.  The slot was last read several frames ago.
  const int S = c->staging_slot;
  f.slot = S;
  if (c->main_dirty && c->flights_staged == 0) {   // kernels launched outside the pipeline may still be reading the ring
    VSF_CUDA(c, cudaEventRecord(c->ev_main, c->stream));
    VSF_CUDA(c, cudaStreamWaitEvent(c->up_stream, c->ev_main, 0));
  }
  if (c->slot_last_chain[S]) VSF_CUDA(c, cudaStreamWaitEvent(c->up_stream, c->slot_last_chain[S], 0));
  if (n > 0) {
    const uint8_t* src = f.h_desc;
    if ((flags & VSF_SUBMIT_PINNED_DESC) && stride == size_t(c->row_bytes) &&
        (c->desc_bytes == c->row_bytes || (flags & VSF_SUBMIT_PADDED_ROWS))) {
      src = desc;   // page-locked and already in the device layout: no staging copy
    } else if (stride == size_t(c->row_bytes) && c->desc_bytes == c->row_bytes) {
      std::memcpy(f.h_desc, desc, size_t(n) * c->row_bytes);
    } else {
      for (int i = 0; i < n; ++i) {
        std::memcpy(f.h_desc + size_t(i) * c->row_bytes, desc + size_t(i) * stride, c->desc_bytes);
        if (c->desc_bytes < c->row_bytes)
          std::memset(f.h_desc + size_t(i) * c->row_bytes + c->desc_bytes, 0, c->row_bytes - c->desc_bytes);
      }
    }
    VSF_CUDA(c, cudaMemcpyAsync(c->slot_ptr(S), src, size_t(n) * c->row_bytes, cudaMemcpyHostToDevice, c->up_stream));
  }
  // the tensor engine's +-1 image of the frame is made here too, right behind the copy, so the
  // main stream goes from one frame's compaction straight to the next frame's distance kernel
  const bool pre_expand = n > 0 && nf > 0 && f.d_train_exp && c->engine != 1 && !(c->words == 16 && c->engine == 3);
  const int exp_int8 = c->engine == 3 ? 0 : 1;
  if (pre_expand) {
    if (c->words == 16)
      VSF_CUDA(c, launch_expand_train64(c->slot_ptr(S), n, nullptr, f.d_train_exp, 0, c->up_stream));
    else
      VSF_CUDA(c, launch_expand_train(c->slot_ptr(S), n, nullptr, f.d_train_exp, exp_int8, 0, c->up_stream, nullptr));
  }
  VSF_CUDA(c, cudaEventRecord(f.ev_up, c->up_stream));
  f.specs.clear();
  for (int j = 0; j < nf; ++j) {
    const int s = c->live[j];
    f.specs.push_back(ProblemSpec{c->slot_ptr(s), c->slot_count[s], c->slot_dev[s] ? c->d_slot_rows + s : nullptr,
                                  c->slot_ptr(S), n, nullptr, j});
    if (pre_expand) {
      f.specs.back().t_exp = f.d_train_exp;
      f.specs.back().t_exp_int8 = exp_int8;
    }
  }
  commit_staging(c, frame_id, n);   // eviction + push (src/slam_frontend.cc:467-470)
  ++c->flight_count;
  ++c->flights_staged;
  if (c->flights_staged >= std::max(1, std::min(c->defer_group, kMaxPoseGroup))) return flush_flights(c);
  return VSF_OK;
}

extern "C" int vsf_window_collect(vsf_ctx* c, uint64_t* frame_id, uint64_t* frame_ids, int* counts,
                                  vsf_feature_match* out, int cap_per_frame, int* n_frames) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (!n_frames || !counts || cap_per_frame < 0 || (cap_per_frame > 0 && !out)) return fail(c, VSF_ERR_BAD_ARG, "null output");
  if (c->flight_count == 0) return fail(c, VSF_ERR_STATE, "no submitted frame to collect");
  cudaSetDevice(c->device);
  vsf_ctx::Flight& f = c->flights[c->flight_head];
  if (!f.launched) {   // (only inside vsf_window_run_sequence: a group still filling up)
    const int rc = flush_flights(c);
    if (rc) return rc;
  }
  cudaError_t werr = cudaSuccess;
  const auto tw0 = std::chrono::steady_clock::now();
  if (f.sort_mode == 1) {
    std::unique_lock<std::mutex> lk(c->done_mu);
    c->done_cv.wait(lk, [&f] { return f.pending.load(std::memory_order_acquire) == 0; });
    werr = cudaError_t(f.cuda_error);
  } else {
    werr = cudaEventSynchronize(f.done);
  }
  const auto tw1 = std::chrono::steady_clock::now();
  c->tm_wait += std::chrono::duration<double, std::micro>(tw1 - tw0).count();
  c->tm_ready += std::chrono::duration<double, std::micro>(tw1 - f.t_flush).count();
  ++c->tm_n;
  static const bool trace_events = std::getenv("VSF_TIMING2") != nullptr;
  if (trace_events)
    std::fprintf(stderr, "READY %.1f frame %llu entered %.1f\n", std::chrono::duration<double, std::micro>(tw1.time_since_epoch()).count(),
                 (unsigned long long)f.frame_id, std::chrono::duration<double, std::micro>(tw0.time_since_epoch()).count());
  // the flight is consumed whatever happens next
  c->flight_head = (c->flight_head + 1) % VSF_PIPELINE_DEPTH;
  --c->flight_count;
  VSF_CUDA(c, werr);
  const int nf = f.nf;
  *n_frames = nf;
  if (frame_id) *frame_id = f.frame_id;
  c->last_h2d = f.h2d_bytes;
  c->last_d2h = f.d2h_bytes;
  for (int j = 0; j < nf; ++j) {
    const int keep = f.sort_mode != 1 ? f.h_counts[j] : f.keep[j];
    if (keep > cap_per_frame) return fail(c, VSF_ERR_CAPACITY, "cap_per_frame too small");
    counts[j] = keep;
    if (frame_ids) frame_ids[j] = f.fids[j];
  }
  for (int j = 0; j < nf; ++j)
    if (counts[j]) std::memcpy(out + size_t(j) * cap_per_frame, f.h_fm + size_t(j) * c->rows_pad, size_t(counts[j]) * sizeof(vsf_feature_match));
  c->tm_copy += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tw1).count();
  return VSF_OK;
}

extern "C" int vsf_set_host_threads(vsf_ctx* c, int n) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (n < 1 || n > 256) return fail(c, VSF_ERR_BAD_ARG, "host thread count must be 1..256");
  if (c->flights_ready) return fail(c, VSF_ERR_STATE, "worker threads already started (call before the first vsf_window_submit)");
  c->host_threads = n;
  return VSF_OK;
}

// ------------------------------------------------------------------------------ a5 (stereo)

static void fill_stereo_args(vsf_ctx* c, StereoArgs& a, const float* F) {
  std::memset(&a, 0, sizeof(a));
  const int region = c->window + 1;
  a.matches = c->region_ptr(region);
  a.n_matches = c->d_match_count + region;
  a.xy_left = c->d_xy_left;
  a.xy_right = c->d_xy_right;
  for (int i = 0; i < 9; ++i) a.F[i] = F[i];
  a.thresh_cur = c->d_thresh + c->thresh_cur;
  a.thresh_next = c->d_thresh + (c->thresh_cur ^ 1);
  a.resid = c->d_resid;
  a.chunk_keep = c->d_chunk_keep;
  a.chunk_off = c->d_chunk_off;
  a.ticket = c->d_ticket;
  a.residual_order = c->opt_residual_order;
  a.hold_on_empty = c->opt_hold_on_empty;
  a.kept_left = c->d_kept_left;
  a.kept_right = c->d_kept_right;
  a.n_kept = c->d_slot_rows + c->staging_slot;
  a.desc_left = reinterpret_cast<const uint32_t*>(c->d_raw_left);
  a.desc_right = reinterpret_cast<const uint32_t*>(c->d_raw_right);
  a.desc_left_c = reinterpret_cast<uint32_t*>(c->slot_ptr(c->staging_slot));
  a.desc_right_c = reinterpret_cast<uint32_t*>(c->d_right_c);
  a.xy_left_c = c->d_xy_left_c;
  a.xy_right_c = c->d_xy_right_c;
  a.words = c->words;
}

// L->R kNN + ratio, epipolar filter + compaction on frames that are already on the device (all
// async).  The ctx's threshold ping-pong (thresh_cur) is NOT advanced here: the caller flips it
// once every stage of its call has been launched successfully, so a failed call leaves the
// adaptive threshold as it was.  with_threshold = false: the caller places the threshold sum.
static int stereo_launch(vsf_ctx* c, const uint8_t* d_dl, int nl, const uint8_t* d_dr, int nr, const float2* d_xyl,
                         const float2* d_xyr, const float* F, double ratio, bool with_threshold, StereoArgs* a) {
  std::vector<ProblemSpec> specs(1);
  specs[0] = ProblemSpec{d_dl, nl, nullptr, d_dr, nr, nullptr, c->window + 1};
  int rc;
  if ((rc = run_knn(c, specs, ratio, false, true))) return rc;
  fill_stereo_args(c, *a, F);
  a->desc_left = reinterpret_cast<const uint32_t*>(d_dl);
  a->desc_right = reinterpret_cast<const uint32_t*>(d_dr);
  a->xy_left = d_xyl;
  a->xy_right = d_xyr;
  return VSF_OK;
}

// upload both frames, then stereo_launch + the filter kernels
static int stereo_stage(vsf_ctx* c, const vsf_keypoint* kpl, const uint8_t* dl, int nl, size_t sl,
                        const vsf_keypoint* kpr, const uint8_t* dr, int nr, size_t sr, const float* F,
                        double ratio, bool with_threshold, StereoArgs* args_out = nullptr) {
  if (nl < 0 || nr < 0 || (nl > 0 && (!kpl || !dl)) || (nr > 0 && (!kpr || !dr)) || !F)
    return fail(c, VSF_ERR_BAD_ARG, "bad stereo arguments");
  if ((nl > 0 && sl < size_t(c->desc_bytes)) || (nr > 0 && sr < size_t(c->desc_bytes)))
    return fail(c, VSF_ERR_BAD_ARG, "row stride smaller than desc_bytes");
  if (nl > c->max_features || nr > c->max_features) return fail(c, VSF_ERR_CAPACITY, "more rows than max_features");
  cudaSetDevice(c->device);
  int rc;
  if ((rc = upload_desc(c, 0, dl, nl, sl, c->d_raw_left))) return rc;
  if ((rc = upload_desc(c, 1, dr, nr, sr, c->d_raw_right))) return rc;
  if ((rc = upload_xy(c, 0, kpl, nl, c->d_xy_left))) return rc;
  if ((rc = upload_xy(c, 1, kpr, nr, c->d_xy_right))) return rc;
  StereoArgs a;
  if ((rc = stereo_launch(c, c->d_raw_left, nl, c->d_raw_right, nr, c->d_xy_left, c->d_xy_right, F, ratio,
                          with_threshold, &a)))
    return rc;
  VSF_CUDA(c, launch_stereo_filter(a, std::max(nl, 1), with_threshold, c->stream));
  if (args_out) *args_out = a;
  return VSF_OK;
}

extern "C" int vsf_stereo_filter(vsf_ctx* c, const vsf_keypoint* kpl, const uint8_t* dl, int nl, size_t sl,
                                 const vsf_keypoint* kpr, const uint8_t* dr, int nr, size_t sr, const float* F,
                                 double ratio, int32_t* kept_left, int32_t* kept_right, int* n_kept,
                                 vsf_dmatch* stereo_matches, float* residuals, int* n_stereo) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (!n_kept) return fail(c, VSF_ERR_BAD_ARG, "n_kept is required");
  int rc = stereo_stage(c, kpl, dl, nl, sl, kpr, dr, nr, sr, F, ratio, true);
  if (rc) return rc;
  c->thresh_cur ^= 1;
  const int region = c->window + 1;
  VSF_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_slot_rows + c->staging_slot, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  VSF_CUDA(c, cudaMemcpyAsync(c->h_counts + 1, c->d_match_count + region, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));
  const int M = c->h_counts[0], ns = c->h_counts[1];
  if (M > 0) {
    VSF_CUDA(c, cudaMemcpyAsync(c->h_kept[0], c->d_kept_left, size_t(M) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    VSF_CUDA(c, cudaMemcpyAsync(c->h_kept[1], c->d_kept_right, size_t(M) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  }
  if (ns > 0 && stereo_matches)
    VSF_CUDA(c, cudaMemcpyAsync(c->h_matches, c->region_ptr(region), size_t(ns) * sizeof(vsf_dmatch), cudaMemcpyDeviceToHost, c->stream));
  if (ns > 0 && residuals)
    VSF_CUDA(c, cudaMemcpyAsync(c->h_resid, c->d_resid, size_t(ns) * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));
  *n_kept = M;
  if (n_stereo) *n_stereo = ns;
  if (M > 0 && kept_left) std::memcpy(kept_left, c->h_kept[0], size_t(M) * sizeof(int));
  if (M > 0 && kept_right) std::memcpy(kept_right, c->h_kept[1], size_t(M) * sizeof(int));
  if (ns > 0 && stereo_matches) std::memcpy(stereo_matches, c->h_matches, size_t(ns) * sizeof(vsf_dmatch));
  if (ns > 0 && residuals) std::memcpy(residuals, c->h_resid, size_t(ns) * sizeof(float));
  return VSF_OK;
}

extern "C" int vsf_set_stereo_threshold(vsf_ctx* c, float value) {
  if (!c) return VSF_ERR_BAD_ARG;
  cudaSetDevice(c->device);
  c->h_scalar[0] = value;
  VSF_CUDA(c, cudaMemcpyAsync(c->d_thresh + c->thresh_cur, c->h_scalar, sizeof(float), cudaMemcpyHostToDevice, c->stream));
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));
  return VSF_OK;
}

extern "C" int vsf_set_option(vsf_ctx* c, int option, int value) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (option == VSF_OPT_DEBUG_SORT_DEPTH) {
    if (value < -1 || value > 64) return fail(c, VSF_ERR_BAD_ARG, "sort depth must be -1 .. 64");
    c->sort_depth_override = value;
    return VSF_OK;
  }
  if (option == VSF_OPT_POSE_GROUP) {
    if (value < 1 || value > kMaxPoseGroup) return fail(c, VSF_ERR_BAD_ARG, "pose group must be 1 .. 8");
    c->pose_group = value;
    return VSF_OK;
  }
  if (value != 0 && value != 1) return fail(c, VSF_ERR_BAD_ARG, "option value must be 0 or 1");
  switch (option) {
    case VSF_OPT_RESIDUAL_ORDER: c->opt_residual_order = value; return VSF_OK;
    case VSF_OPT_HOLD_THRESHOLD_ON_EMPTY: c->opt_hold_on_empty = value; return VSF_OK;
    default: return fail(c, VSF_ERR_BAD_ARG, "unknown option");
  }
}

extern "C" int vsf_get_option(const vsf_ctx* c, int option, int* value) {
  if (!c || !value) return VSF_ERR_BAD_ARG;
  switch (option) {
    case VSF_OPT_RESIDUAL_ORDER: *value = c->opt_residual_order; return VSF_OK;
    case VSF_OPT_HOLD_THRESHOLD_ON_EMPTY: *value = c->opt_hold_on_empty; return VSF_OK;
    case VSF_OPT_DEBUG_SORT_DEPTH: *value = c->sort_depth_override; return VSF_OK;
    case VSF_OPT_POSE_GROUP: *value = c->pose_group; return VSF_OK;
    default: return VSF_ERR_BAD_ARG;
  }
}

extern "C" int vsf_get_stereo_threshold(vsf_ctx* c, float* value) {
  if (!c || !value) return VSF_ERR_BAD_ARG;
  cudaSetDevice(c->device);
  VSF_CUDA(c, cudaMemcpyAsync(c->h_scalar, c->d_thresh + c->thresh_cur, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));
  *value = c->h_scalar[0];
  return VSF_OK;
}

// --------------------------------------------------------------------------- a6 (triangulate)

extern "C" int vsf_triangulate(vsf_ctx* c, const float* P1, const float* P2, const float* x1, const float* x2,
                               int n, float* X4) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (n < 0 || !P1 || !P2 || (n > 0 && (!x1 || !x2 || !X4))) return fail(c, VSF_ERR_BAD_ARG, "bad triangulate arguments");
  if (n > c->max_features) return fail(c, VSF_ERR_CAPACITY, "more points than max_features");
  if (n == 0) return VSF_OK;
  cudaSetDevice(c->device);
  std::memcpy(c->h_tri_io, x1, size_t(n) * 2 * sizeof(float));
  std::memcpy(c->h_tri_io + size_t(n) * 2, x2, size_t(n) * 2 * sizeof(float));
  VSF_CUDA(c, cudaMemcpyAsync(c->d_tri_io, c->h_tri_io, size_t(n) * 4 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  float* dX = c->d_tri_io + size_t(c->rows_pad) * 4;
  VSF_CUDA(c, launch_triangulate_pairs(P1, P2, reinterpret_cast<const float2*>(c->d_tri_io),
                                       reinterpret_cast<const float2*>(c->d_tri_io + size_t(n) * 2), n, dX, c->stream));
  VSF_CUDA(c, cudaMemcpyAsync(c->h_tri_io, dX, size_t(n) * 4 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));
  std::memcpy(X4, c->h_tri_io, size_t(n) * 4 * sizeof(float));
  return VSF_OK;
}

// ------------------------------------------------------------------- fused ObserveImage path

extern "C" int vsf_undistort_points(vsf_ctx* c, const float* K, const float* dist, const float* xy, int n, float* out) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (n < 0 || !K || !dist || (n > 0 && (!xy || !out))) return fail(c, VSF_ERR_BAD_ARG, "bad undistort arguments");
  if (n > c->max_features) return fail(c, VSF_ERR_CAPACITY, "more points than max_features");
  if (n == 0) return VSF_OK;
  cudaSetDevice(c->device);
  std::memcpy(c->h_tri_io, xy, size_t(n) * 2 * sizeof(float));
  VSF_CUDA(c, cudaMemcpyAsync(c->d_tri_io, c->h_tri_io, size_t(n) * 2 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  float* d_out = c->d_tri_io + size_t(c->rows_pad) * 4;
  VSF_CUDA(c, launch_undistort_points(reinterpret_cast<const float2*>(c->d_tri_io), n, K, dist,
                                      reinterpret_cast<float2*>(d_out), c->stream));
  float* h_out = c->h_tri_io + size_t(c->rows_pad) * 4;
  VSF_CUDA(c, cudaMemcpyAsync(h_out, d_out, size_t(n) * 2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));
  std::memcpy(out, h_out, size_t(n) * 2 * sizeof(float));
  return VSF_OK;
}

static int obs_init(vsf_ctx* c) {
  if (c->obs_ready) return VSF_OK;
  cudaSetDevice(c->device);
  const size_t N = size_t(c->rows_pad);
  const size_t in_bytes = 2 * N * c->row_bytes + 2 * N * sizeof(float2) + 512;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t o_kl = 0, o_kr = up(o_kl + N * 4), o_li = up(o_kr + N * 4);
  const size_t o_ct = up(o_li + size_t(c->window + 1) * N * sizeof(vsf_dmatch));
  const size_t o_sc = up(o_ct + (kMaxProblems + 8) * sizeof(int)), o_x4 = up(o_sc + 64);
  const size_t o_xu = up(o_x4 + N * sizeof(float4)), out_bytes = up(o_xu + N * sizeof(float2));
  VSF_CUDA(c, cudaStreamCreateWithFlags(&c->obs_up_stream, cudaStreamNonBlocking));
  for (vsf_ctx::ObsFlight& f : c->obs) {
    VSF_CUDA(c, cudaEventCreateWithFlags(&f.ev_up, cudaEventDisableTiming));
    VSF_CUDA(c, cudaEventCreateWithFlags(&f.ev_done, cudaEventDisableTiming));
    VSF_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&f.h_in), in_bytes));
    VSF_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&f.d_in), in_bytes));
    VSF_CUDA(c, cudaHostAlloc(reinterpret_cast<void**>(&f.h_out), out_bytes, cudaHostAllocMapped));
    std::memset(f.h_out, 0, out_bytes);
    uint8_t* dm = nullptr;
    VSF_CUDA(c, cudaHostGetDevicePointer(reinterpret_cast<void**>(&dm), f.h_out, 0));
    f.h_kept_left = reinterpret_cast<int*>(f.h_out + o_kl);      f.dm_kept_left = reinterpret_cast<int*>(dm + o_kl);
    f.h_kept_right = reinterpret_cast<int*>(f.h_out + o_kr);     f.dm_kept_right = reinterpret_cast<int*>(dm + o_kr);
    f.h_lists = reinterpret_cast<vsf_dmatch*>(f.h_out + o_li);   f.dm_lists = reinterpret_cast<vsf_dmatch*>(dm + o_li);
    f.h_counts = reinterpret_cast<int*>(f.h_out + o_ct);         f.dm_counts = reinterpret_cast<int*>(dm + o_ct);
    f.h_scalar = reinterpret_cast<float*>(f.h_out + o_sc);       f.dm_scalar = reinterpret_cast<float*>(dm + o_sc);
    f.h_X4 = reinterpret_cast<float4*>(f.h_out + o_x4);          f.dm_X4 = reinterpret_cast<float4*>(dm + o_x4);
    f.h_xyu = reinterpret_cast<float2*>(f.h_out + o_xu);         f.dm_xyu = reinterpret_cast<float2*>(dm + o_xu);
  }
  c->obs_ready = true;
  return VSF_OK;
}

extern "C" int vsf_observe_in_flight(const vsf_ctx* c) { return c ? c->obs_count : 0; }

extern "C" int vsf_observe_submit(vsf_ctx* c, uint64_t frame_id, const vsf_keypoint* kpl, const uint8_t* dl, int nl,
                                  size_t sl, const vsf_keypoint* kpr, const uint8_t* dr, int nr, size_t sr,
                                  const vsf_observe_params* p) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (!p || !p->fundamental || !p->P_left || !p->P_right) return fail(c, VSF_ERR_BAD_ARG, "null observe parameters");
  if ((p->K_left == nullptr) != (p->dist_left == nullptr)) return fail(c, VSF_ERR_BAD_ARG, "K_left and dist_left go together");
  if (nl < 0 || nr < 0 || (nl > 0 && (!kpl || !dl)) || (nr > 0 && (!kpr || !dr)))
    return fail(c, VSF_ERR_BAD_ARG, "bad stereo arguments");
  if ((nl > 0 && sl < size_t(c->desc_bytes)) || (nr > 0 && sr < size_t(c->desc_bytes)))
    return fail(c, VSF_ERR_BAD_ARG, "row stride smaller than desc_bytes");
  if (nl > c->max_features || nr > c->max_features) return fail(c, VSF_ERR_CAPACITY, "more rows than max_features");
  if (c->obs_count >= VSF_OBSERVE_DEPTH)
    return fail(c, VSF_ERR_STATE, "VSF_OBSERVE_DEPTH frames already in flight: call vsf_observe_collect first");
  int rc = obs_init(c);
  if (rc) return rc;
  cudaSetDevice(c->device);
  vsf_ctx::ObsFlight& f = c->obs[(c->obs_head + c->obs_count) % VSF_OBSERVE_DEPTH];
  const int rb = c->row_bytes;
  // ---- one upload: [left rows | right rows | left pixels | right pixels], packed into the
  // flight's pinned block (it was last read by the frame collected VSF_OBSERVE_DEPTH submits ago)
  auto up = [](size_t v) { return (v + 127) / 128 * 128; };
  const size_t o_dl = 0, o_dr = up(size_t(nl) * rb), o_xl = up(o_dr + size_t(nr) * rb);
  const size_t o_xr = up(o_xl + size_t(nl) * sizeof(float2)), total = o_xr + size_t(nr) * sizeof(float2);
  auto pack = [&](uint8_t* dst, const uint8_t* src, int n, size_t stride) {
    if (n == 0) return;
    if (stride == size_t(rb) && c->desc_bytes == rb) {
      std::memcpy(dst, src, size_t(n) * rb);
    } else {
      for (int i = 0; i < n; ++i) {
        std::memcpy(dst + size_t(i) * rb, src + size_t(i) * stride, c->desc_bytes);
        if (c->desc_bytes < rb) std::memset(dst + size_t(i) * rb + c->desc_bytes, 0, rb - c->desc_bytes);
      }
    }
  };
  pack(f.h_in + o_dl, dl, nl, sl);
  pack(f.h_in + o_dr, dr, nr, sr);
  float2* hxl = reinterpret_cast<float2*>(f.h_in + o_xl);
  float2* hxr = reinterpret_cast<float2*>(f.h_in + o_xr);
  for (int i = 0; i < nl; ++i) hxl[i] = make_float2(kpl[i].x, kpl[i].y);
  for (int i = 0; i < nr; ++i) hxr[i] = make_float2(kpr[i].x, kpr[i].y);
  if (total > 0) VSF_CUDA(c, cudaMemcpyAsync(f.d_in, f.h_in, total, cudaMemcpyHostToDevice, c->obs_up_stream));
  VSF_CUDA(c, cudaEventRecord(f.ev_up, c->obs_up_stream));
  VSF_CUDA(c, cudaStreamWaitEvent(c->stream, f.ev_up, 0));
  // ---- main stream.  a5: stereo L->R + filter; the compacted left frame lands in the ring's
  // staging slot, its row count in d_slot_rows[slot]
  const int nf = int(c->live.size());
  f.nf = nf;
  f.nl = nl;
  f.slot = c->staging_slot;
  f.frame_id = frame_id;
  f.undistort = p->K_left != nullptr;
  StereoArgs sa;
  if ((rc = stereo_launch(c, f.d_in + o_dl, nl, f.d_in + o_dr, nr, reinterpret_cast<const float2*>(f.d_in + o_xl),
                          reinterpret_cast<const float2*>(f.d_in + o_xr), p->fundamental, p->nn_match_ratio, false, &sa)))
    return rc;
  sa.kept_left = f.dm_kept_left;      // straight into mapped host memory
  sa.kept_right = f.dm_kept_right;
  sa.n_kept_host = f.dm_counts + kMaxProblems;
  sa.thresh_next_host = f.dm_scalar;
  VSF_CUDA(c, launch_stereo_filter(sa, std::max(nl, 1), false, c->stream));
  // a4 + a6 matching in one launch sequence: every resident past frame vs the compacted left
  // frame, and compacted right (query) vs compacted left (train); lists and counts are mirrored
  // into the flight's mapped memory by the compaction
  int* d_m = c->d_slot_rows + c->staging_slot;
  std::vector<ProblemSpec> specs;
  uint8_t* cur = c->slot_ptr(c->staging_slot);
  for (int j = 0; j < nf; ++j) {
    const int s = c->live[j];
    f.fids[j] = c->slot_frame[s];
    f.bounds[j] = c->slot_count[s];
    specs.push_back(ProblemSpec{c->slot_ptr(s), c->slot_count[s], c->slot_dev[s] ? c->d_slot_rows + s : nullptr, cur, nl, d_m, j});
  }
  const int tri_region = c->window;
  specs.push_back(ProblemSpec{c->d_right_c, nl, d_m, cur, nl, d_m, tri_region});
  c->mir_dm = f.dm_lists;
  c->mir_dcounts = f.dm_counts;
  c->mir_hcounts = f.h_counts;
  rc = run_knn(c, specs, p->nn_match_ratio, true, true);
  c->mir_dm = c->dm_matches;
  c->mir_dcounts = c->dm_region_counts;
  c->mir_hcounts = c->h_region_counts;
  if (rc) return rc;
  // triangulation; thread i also undistorts compacted left keypoint i (N1) and the extra CTA
  // carries the stereo stage's sequential threshold sum, which nothing in this frame waits for
  TriExtras ex;
  std::memset(&ex, 0, sizeof(ex));
  ex.do_threshold = 1;
  ex.stereo = sa;
  if (f.undistort) {
    ex.do_undistort = 1;
    ex.und = make_undistort_args(p->K_left, p->dist_left);
    ex.n_kept = d_m;
    ex.xy_undist = f.dm_xyu;
  }
  VSF_CUDA(c, launch_triangulate_matches(p->P_left, p->P_right, c->region_ptr(tri_region), c->d_match_count + tri_region, nl,
                                         c->d_xy_left_c, c->d_xy_right_c, f.dm_X4, &ex, c->stream));
  VSF_CUDA(c, cudaEventRecord(f.ev_done, c->stream));
  // every stage is enqueued: advance the adaptive threshold and the window together.  The
  // frame's row count M is only on the device yet; the host keeps the bound nl until collect.
  c->thresh_cur ^= 1;
  const int slot = c->staging_slot;
  commit_staging(c, frame_id, nl);
  c->slot_dev[slot] = 1;
  ++c->obs_count;
  return VSF_OK;
}

extern "C" int vsf_observe_collect(vsf_ctx* c, uint64_t* frame_id, vsf_observe_out* out) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (!out) return fail(c, VSF_ERR_BAD_ARG, "null output");
  if (c->obs_count == 0) return fail(c, VSF_ERR_STATE, "no submitted frame to collect");
  cudaSetDevice(c->device);
  vsf_ctx::ObsFlight& f = c->obs[c->obs_head];
  const cudaError_t werr = cudaEventSynchronize(f.ev_done);
  c->obs_head = (c->obs_head + 1) % VSF_OBSERVE_DEPTH;   // the flight is consumed whatever happens next
  --c->obs_count;
  VSF_CUDA(c, werr);
  const int M = f.h_counts[kMaxProblems];
  const int tri_region = c->window;
  const int n_tri = f.h_counts[tri_region];
  // the exact row count replaces the bound while the frame is still resident
  if (c->slot_frame[f.slot] == f.frame_id && c->slot_dev[f.slot]) {
    for (int s : c->live)
      if (s == f.slot) {
        c->slot_count[s] = M;
        c->slot_dev[s] = 0;
      }
  }
  if (frame_id) *frame_id = f.frame_id;
  int need = std::max(M, n_tri);
  for (int k = 0; k < f.nf; ++k) need = std::max(need, f.h_counts[k]);
  if (out->cap < need) return fail(c, VSF_ERR_CAPACITY, "vsf_observe_out.cap is smaller than one of this frame's lists");
  out->n_kept = M;
  out->stereo_threshold_next = f.h_scalar[0];
  out->n_frames = f.nf;
  out->n_tri = n_tri;
  if (M > 0 && out->kept_left) std::memcpy(out->kept_left, f.h_kept_left, size_t(M) * sizeof(int));
  if (M > 0 && out->kept_right) std::memcpy(out->kept_right, f.h_kept_right, size_t(M) * sizeof(int));
  for (int k = 0; k < f.nf; ++k) {
    if (out->frame_ids) out->frame_ids[k] = f.fids[k];
    if (out->window_counts) out->window_counts[k] = f.h_counts[k];
    if (out->window_matches && f.h_counts[k] > 0)
      std::memcpy(out->window_matches + size_t(k) * out->cap, f.h_lists + size_t(k) * c->rows_pad,
                  size_t(f.h_counts[k]) * sizeof(vsf_dmatch));
  }
  if (n_tri > 0 && out->tri_matches)
    std::memcpy(out->tri_matches, f.h_lists + size_t(tri_region) * c->rows_pad, size_t(n_tri) * sizeof(vsf_dmatch));
  if (n_tri > 0 && out->tri_X4) std::memcpy(out->tri_X4, f.h_X4, size_t(n_tri) * sizeof(float4));
  if (M > 0 && out->xy_undist && f.undistort) std::memcpy(out->xy_undist, f.h_xyu, size_t(M) * sizeof(float2));
  return VSF_OK;
}

extern "C" int vsf_observe_features(vsf_ctx* c, uint64_t frame_id, const vsf_keypoint* kpl, const uint8_t* dl,
                                    int nl, size_t sl, const vsf_keypoint* kpr, const uint8_t* dr, int nr,
                                    size_t sr, const float* F, const float* P_left, const float* P_right,
                                    double ratio, vsf_observe_out* out) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (!out || !P_left || !P_right || !F) return fail(c, VSF_ERR_BAD_ARG, "null argument");
  if (c->obs_count != 0) return fail(c, VSF_ERR_STATE, "frames of vsf_observe_submit are still in flight");
  // a window list holds at most one entry per row of the PAST frame, the stereo / triangulation
  // lists at most one per row of this frame: refuse before anything is launched, so that a
  // failed call leaves the window and the adaptive threshold untouched
  int need = nl;
  for (int s : c->live) need = std::max(need, c->slot_count[s]);
  if (out->cap < need)
    return fail(c, VSF_ERR_CAPACITY, "vsf_observe_out.cap must be >= n_left and >= the row count of every resident frame");
  vsf_observe_params p;
  p.fundamental = F;
  p.P_left = P_left;
  p.P_right = P_right;
  p.K_left = nullptr;
  p.dist_left = nullptr;
  p.nn_match_ratio = ratio;
  int rc = vsf_observe_submit(c, frame_id, kpl, dl, nl, sl, kpr, dr, nr, sr, &p);
  if (rc) return rc;
  return vsf_observe_collect(c, nullptr, out);
}

// ----------------------------------------------------------------- device-resident entry points

extern "C" int vsf_window_match_device(vsf_ctx* c, const void* const* d_queries, const int* nq, int n_frames,
                                       const void* d_train, int nt, double ratio) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (n_frames < 0 || n_frames > c->window || (n_frames > 0 && (!d_queries || !nq)) || nt < 0 || (nt > 0 && !d_train))
    return fail(c, VSF_ERR_BAD_ARG, "bad device window arguments");
  cudaSetDevice(c->device);
  std::vector<ProblemSpec> specs;
  for (int j = 0; j < n_frames; ++j)
    specs.push_back(ProblemSpec{d_queries[j], nq[j], nullptr, d_train, nt, nullptr, j});
  c->last_n_frames = n_frames;
  return run_knn(c, specs, ratio);
}

extern "C" int vsf_window_match_block_device(vsf_ctx* c, const void* d_seq, int n, int n_poses, long long first,
                                             int count, double ratio) {
  if (!c) return VSF_ERR_BAD_ARG;
  const int W = c->window;
  if (!d_seq || n < 0 || n > c->max_features || n_poses <= W || first < 0 || count < 0)
    return fail(c, VSF_ERR_BAD_ARG, "bad device sequence arguments");
  cudaSetDevice(c->device);
  const size_t fb = size_t(n) * c->row_bytes;
  const uint8_t* base = static_cast<const uint8_t*>(d_seq);
  auto pose_frame = [&](int k) { return (first + k) % (n_poses - W) + W; };
  // Tensor engine (engine flag 256: pose-by-pose launches; A/B timing).  The poses are launched
  // in groups of G (ctx pose_group, VSF_POSE_GROUP; G * window <= kMaxProblems), each group as
  // ONE batch of all its frame pairs: one distance kernel - its persistent CTAs walk the work
  // slots of G poses without a kernel boundary in between - and one finish kernel.  The +-1
  // images of the NEXT group's current frames are made by one small kernel on a side stream while
  // this group's distance kernel runs (its 128-thread CTAs fit beside the distance CTAs; the
  // finish kernel is let in late, TcBatch::late, so that it does not take that room), which is
  // how the host-buffer path has always done it on its upload stream.  Buffers: +-1 images 2 G
  // deep, partial keys 2 deep, two finish-kernel launch states; the last pose of the call writes
  // the ctx's own lists (vsf_fetch_window), the others scratch lists.
  const bool wide = c->words == 16;
  const bool tensor = c->engine >= 2 || (c->engine == 0 && double(W) * double(n) * double(n) >= c->tc_auto_min_cmp);
  const bool ahead = tensor && (c->words == 8 || (wide && c->engine != 3)) && n > 0 && count > 0 &&
                     !(c->engine_flags & (256 | 512));   // (vsf_set_profile: events around the kernels of every batch)
  const int int8 = c->engine == 3 ? 0 : 1;
  const int G = ahead ? std::max(1, std::min(c->pose_group, kMaxProblems / W)) : 1;
  std::vector<ProblemSpec> specs;
  auto add_pose = [&](int k, const uint8_t* image, int slot) {
    const long long cur = pose_frame(k);
    const size_t rows_cap = size_t(c->regions) * size_t(c->rows_pad);
    for (int j = 0; j < W; ++j) {
      ProblemSpec sp{base + size_t(cur - W + j) * fb, n, nullptr, base + size_t(cur) * fb, n, nullptr, j};
      if (image) {
        sp.t_exp = image;
        sp.t_exp_int8 = int8;
      }
      if (slot > 0) {   // scratch lists (slot 0 = the ctx's own regions)
        sp.out = c->grp_matches + size_t(slot - 1) * rows_cap + size_t(j) * c->rows_pad;
        sp.out_count = c->grp_counts + size_t(slot - 1) * kMaxProblems + j;
      }
      specs.push_back(sp);
    }
  };
  if (!ahead) {
    for (int k = 0; k < count; ++k) {
      specs.clear();
      add_pose(k, nullptr, 0);
      const int rc = run_knn(c, specs, ratio);
      if (rc) return rc;
    }
    c->last_n_frames = W;
    return VSF_OK;
  }
  {
    const int rcs = group_state_init(c);
    if (rcs) return rcs;
  }
  const size_t image_bytes = size_t(round_up(c->max_features, kTcTileRows)) * size_t(c->row_bytes) * 8;
  for (int i = 0; i < 2 * G; ++i)
    if (!c->grp_exp[i]) {
      if (i < kTcMaxTrains) c->grp_exp[i] = c->d_train_exp[i];
      else VSF_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->grp_exp[i]), image_bytes));
    }
  auto group_images = [&](int k0, int parity) {   // frames of poses k0 .. k0 + G - 1
    ExpandMulti em;
    std::memset(&em, 0, sizeof(em));
    em.nt = n;
    for (int g = 0; g < G && k0 + g < count; ++g) {
      em.src[g] = reinterpret_cast<const uint32_t*>(base + size_t(pose_frame(k0 + g)) * fb);
      em.out[g] = c->grp_exp[parity * G + g];
      em.frames = g + 1;
    }
    return em;
  };
  // the images of a group are made on the side stream (ordinary launches: nothing to overlap with
  // on that stream)
  auto expand_images = [&](const ExpandMulti& em) -> int {
    if (em.frames == 0) return VSF_OK;
    if (wide) VSF_CUDA(c, launch_expand_train64_multi(em, 0, c->exp_stream));
    else VSF_CUDA(c, launch_expand_train_multi(em, int8, 0, c->exp_stream));
    ++c->launches;
    return VSF_OK;
  };
  // whatever the stream holds so far may still read the image buffers
  VSF_CUDA(c, cudaEventRecord(c->ev_blk, c->stream));
  VSF_CUDA(c, cudaStreamWaitEvent(c->exp_stream, c->ev_blk, 0));
  int rc = expand_images(group_images(0, 0));
  if (rc) return rc;
  VSF_CUDA(c, cudaEventRecord(c->ev_img[0], c->exp_stream));
  unsigned long long* const states = c->grp_flags + group_flag_words(c);
  for (int k0 = 0, grp = 0; k0 < count; k0 += G, ++grp) {
    const int parity = grp & 1;
    const int m = std::min(G, count - k0);
    // ---- side stream: the next group's images, beside this group's distance kernel (the
    // buffers they go to were last read by the previous group)
    const ExpandMulti next_em = group_images(k0 + G, parity ^ 1);
    if (next_em.frames > 0) {
      if (grp > 0) VSF_CUDA(c, cudaStreamWaitEvent(c->exp_stream, c->ev_grp[parity ^ 1], 0));
      if ((rc = expand_images(next_em))) return rc;
      VSF_CUDA(c, cudaEventRecord(c->ev_img[parity ^ 1], c->exp_stream));
    }
    // ---- main stream: this group as one batch
    specs.clear();
    for (int g = 0; g < m; ++g) add_pose(k0 + g, c->grp_exp[parity * G + g], (count - 1 - (k0 + g)) % G);
    VSF_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_img[parity], 0));
    PoseLaunch pl;
    pl.late = 1;   // the finish kernel must not sit beside the distance CTAs: the side stream's kernels need that room
    pl.partial = parity ? c->d_partial2 : c->d_partial;
    pl.flags = c->grp_flags;
    pl.state = states + 4 * parity;
    rc = run_knn(c, specs, ratio, false, false, &pl);
    if (rc) return rc;
    if (k0 + G < count) VSF_CUDA(c, cudaEventRecord(c->ev_grp[parity], c->stream));
  }
  c->last_n_frames = W;
  return VSF_OK;
}

extern "C" int vsf_window_run_sequence(vsf_ctx* c, const uint8_t* h_seq, int n, int n_poses, long long first,
                                       int count, double ratio, float best_percent, int sort_mode, int lag,
                                       vsf_feature_match* out, int* counts, int ring, int cap_per_frame,
                                       size_t* h2d_bytes, size_t* d2h_bytes) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (!h_seq || n < 0 || n_poses < 1 || first < 0 || count < 0 || !out || !counts || ring < 1 || cap_per_frame < 0 ||
      lag < 1 || lag > VSF_PIPELINE_DEPTH - 1)
    return fail(c, VSF_ERR_BAD_ARG, "bad host sequence arguments");
  const int W = c->window;
  const size_t fb = size_t(n) * c->row_bytes;
  const int flags = VSF_SUBMIT_PINNED_DESC | VSF_SUBMIT_PADDED_ROWS;   // the buffer is in the device layout
  // the kernels of up to pose_group frames are launched together (flush_flights)
  struct DeferGuard {
    vsf_ctx* c;
    ~DeferGuard() { c->defer_group = 1; }
  } guard{c};
  // (a group in flight, one being sorted / collected, one being staged: a third of the lag)
  c->defer_group = std::max(1, std::min(c->pose_group, lag / 3));
  uint64_t fid = 0, fids[kMaxProblems];
  int nf = 0;
  long long collected = 0;
  auto collect_one = [&]() -> int {
    const size_t slot = size_t(collected % ring);
    const int rc = vsf_window_collect(c, &fid, fids, counts + slot * W, out + slot * W * size_t(cap_per_frame),
                                      cap_per_frame, &nf);
    if (rc) return rc;
    ++collected;
    if (h2d_bytes) *h2d_bytes += c->last_h2d;
    if (d2h_bytes) *d2h_bytes += c->last_d2h;
    return VSF_OK;
  };
  // VSF_TIMING: where the calling thread's time goes (submits that stage a frame, submits that
  // also launch a group, collects)
  static const bool timing = std::getenv("VSF_TIMING") != nullptr;
  double t_stage = 0.0, t_launch = 0.0, t_collect = 0.0;
  int n_launch = 0;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::micro>(b - a).count();
  };
  for (int k = 0; k < count; ++k) {
    const uint8_t* D = h_seq + size_t((first + k) % n_poses) * fb;
    const auto t0 = now();
    const int staged_before = c->flights_staged;
    int rc = vsf_window_submit(c, uint64_t(first + k), D, n, size_t(c->row_bytes), ratio, best_percent, sort_mode, flags);
    if (rc) return rc;
    const auto t1 = now();
    if (c->flights_staged <= staged_before) {
      t_launch += us(t0, t1);
      ++n_launch;
    } else {
      t_stage += us(t0, t1);
    }
    if (c->flight_count > lag && (rc = collect_one())) return rc;
    t_collect += us(t1, now());
  }
  while (c->flight_count > 0) {
    const int rc = collect_one();
    if (rc) return rc;
  }
  if (timing && count > 0) {
    std::fprintf(stderr, "vsf_window_run_sequence: %d frames; submit (stage only) %.1f us each, submit + group launch %.1f us each (%d), "
                 "collect %.1f us per frame (wait %.1f, copy-out %.1f; launch -> ready %.1f us)\n", count,
                 t_stage / std::max(1, count - n_launch), t_launch / std::max(1, n_launch), n_launch, t_collect / count,
                 c->tm_wait / std::max(1LL, c->tm_n), c->tm_copy / std::max(1LL, c->tm_n), c->tm_ready / std::max(1LL, c->tm_n));
    c->tm_wait = c->tm_copy = c->tm_ready = 0.0;
    c->tm_n = 0;
  }
  return VSF_OK;
}

extern "C" int vsf_fetch_window(vsf_ctx* c, int n_frames, int* counts, vsf_dmatch* out, int cap_per_frame) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (n_frames < 0 || n_frames > c->window || (n_frames > 0 && !counts)) return fail(c, VSF_ERR_BAD_ARG, "bad fetch arguments");
  cudaSetDevice(c->device);
  int rc = fetch_regions(c, n_frames, 0);
  if (rc) return rc;
  for (int j = 0; j < n_frames; ++j) {
    counts[j] = c->h_counts[j];
    if (out) {
      if (counts[j] > cap_per_frame) return fail(c, VSF_ERR_CAPACITY, "cap_per_frame too small");
      if (counts[j]) std::memcpy(out + size_t(j) * cap_per_frame, c->h_matches + size_t(j) * c->rows_pad, size_t(counts[j]) * sizeof(vsf_dmatch));
    }
  }
  return VSF_OK;
}

extern "C" int vsf_device_match_lists(vsf_ctx* c, const vsf_dmatch** d_lists, const int** d_counts, int* stride,
                                      int* regions) {
  if (!c || !d_lists || !d_counts || !stride || !regions) return VSF_ERR_BAD_ARG;
  *d_lists = c->d_matches;
  *d_counts = c->d_match_count;
  *stride = c->rows_pad;
  *regions = c->window;
  return VSF_OK;
}

extern "C" long long vsf_debug_launch_count(const vsf_ctx* c) { return c ? c->launches : 0; }

extern "C" void* vsf_stream(vsf_ctx* c) { return c ? static_cast<void*>(c->stream) : nullptr; }

extern "C" int vsf_synth_sequence_device(vsf_ctx* c, void* d_out, int n, int first_pose, int n_poses, int stride,
                                         uint64_t seed) {
  if (!c) return VSF_ERR_BAD_ARG;
  if (!d_out || n < 1 || n_poses < 0 || first_pose < 0 || stride < 0)
    return fail(c, VSF_ERR_BAD_ARG, "bad synth arguments");
  cudaSetDevice(c->device);
  VSF_CUDA(c, launch_synth(static_cast<uint32_t*>(d_out), n, first_pose, n_poses, stride, seed, c->words,
                           c->desc_bytes, c->stream));
  return VSF_OK;
}

extern "C" int vsf_debug_sort_prefix(uint32_t* keys, int n, int keep) {
  if (n < 0 || keep < 0 || keep > n || (n > 0 && !keys)) return VSF_ERR_BAD_ARG;
  vsf_exact_sort::sort_prefix<kIdxBits>(keys, n, keep);
  return VSF_OK;
}

extern "C" int vsf_debug_sort_prefix_depth(uint32_t* keys, int n, int keep, int depth) {
  if (n < 0 || keep < 0 || keep > n || (n > 0 && !keys) || depth < 0) return VSF_ERR_BAD_ARG;
  vsf_exact_sort::sort_prefix<kIdxBits>(keys, n, keep, depth);
  return VSF_OK;
}

extern "C" int vsf_debug_sort_device(vsf_ctx* c, const vsf_dmatch* matches, int n, float best_percent, int exact,
                                     vsf_feature_match* out, int* n_out) {
  if (!c || n < 0 || (n > 0 && (!matches || !out)) || !n_out) return VSF_ERR_BAD_ARG;
  if (n > c->max_features) return fail(c, VSF_ERR_CAPACITY, "more matches than max_features");
  if (exact && c->rows_pad > sort_exact_max_rows()) return fail(c, VSF_ERR_CAPACITY, "exact device sort needs max_features <= 24576");
  cudaSetDevice(c->device);
  *n_out = 0;
  if (n == 0) return VSF_OK;
  std::memcpy(c->h_matches, matches, size_t(n) * sizeof(vsf_dmatch));
  c->h_counts[0] = n;
  VSF_CUDA(c, cudaMemcpyAsync(c->d_matches, c->h_matches, size_t(n) * sizeof(vsf_dmatch), cudaMemcpyHostToDevice, c->stream));
  VSF_CUDA(c, cudaMemcpyAsync(c->d_match_count, c->h_counts, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  const vsf_dmatch* mp[1] = {c->d_matches};
  const int* cp[1] = {c->d_match_count};
  void* gs = nullptr;
  int rcs = sort_scratch(c, exact, &gs);
  if (rcs) return rcs;
  long long* d_trace = nullptr;
  if (std::getenv("VSF_SORT_TRACE")) {
    VSF_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&d_trace), 16 * sizeof(long long)));
    VSF_CUDA(c, cudaMemset(d_trace, 0, 16 * sizeof(long long)));
  }
  VSF_CUDA(c, launch_sort_cut(mp, cp, 1, best_percent, c->d_fm, c->rows_pad, c->d_fm_count, 8 * c->row_bytes + 1,
                              gs ? c->rows_pad : n, exact, c->sort_depth_override, gs, c->stream, d_trace));
  if (d_trace) {
    long long t[16];
    VSF_CUDA(c, cudaStreamSynchronize(c->stream));
    VSF_CUDA(c, cudaMemcpy(t, d_trace, sizeof(t), cudaMemcpyDeviceToHost));
    cudaFree(d_trace);
    std::fprintf(stderr, "sort trace n=%d exact=%d: P1 %lld P2 %lld P4 %lld P5 %lld load %lld total %lld levels %lld n_eff %lld\n",
                 n, exact, t[0], t[1], t[2], t[3], t[7], t[8], t[9], t[11]);
  }
  VSF_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_fm_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));
  const int keep = c->h_counts[0];
  if (keep < 0 || keep > n) return fail(c, VSF_ERR_CUDA, "sort kernel reported a bad count");
  if (keep > 0) {
    VSF_CUDA(c, cudaMemcpyAsync(c->h_fm, c->d_fm, size_t(keep) * sizeof(vsf_feature_match), cudaMemcpyDeviceToHost, c->stream));
    VSF_CUDA(c, cudaStreamSynchronize(c->stream));
    std::memcpy(out, c->h_fm, size_t(keep) * sizeof(vsf_feature_match));
  }
  *n_out = keep;
  return VSF_OK;
}

extern "C" int vsf_debug_tc_plan(int query_blocks, int train_tiles, int sm_count, int force_split,
                                 long long rows, long long partial_cap, int* out5) {
  if (query_blocks < 1 || train_tiles < 1 || sm_count < 1 || force_split < 0 || rows < 0 || partial_cap < 0 || !out5)
    return VSF_ERR_BAD_ARG;
  TcBatch tb;
  std::memset(&tb, 0, sizeof(tb));
  tb.tiles = train_tiles;
  plan_tc_partition(&tb, query_blocks, sm_count, force_split, size_t(rows), size_t(partial_cap));
  out5[0] = tb.pieces;
  out5[1] = tb.tiles_per_piece;
  out5[2] = tb.grid;
  out5[3] = tb.slots;
  out5[4] = tc_block_segments(tb, 0);   // segments of the first block (host copy of the device formula)
  return VSF_OK;
}

extern "C" int vsf_debug_tc_trace(vsf_ctx* c, long long* out, int max_ctas, int* n_ctas) {
  if (!c || !out || !n_ctas || max_ctas < 0) return VSF_ERR_BAD_ARG;
  if (!c->d_tc_trace) return fail(c, VSF_ERR_STATE, "no timeline recorded (vsf_set_engine flag 16, then a tensor-engine launch)");
  cudaSetDevice(c->device);
  const int n = std::min(max_ctas, c->sm_count);
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));
  VSF_CUDA(c, cudaMemcpy(out, c->d_tc_trace, size_t(n) * kTcTraceSlots * sizeof(long long), cudaMemcpyDeviceToHost));
  *n_ctas = n;
  return VSF_OK;
}

extern "C" int vsf_debug_kernel_trace(vsf_ctx* c, long long* out, int max_records, int* n_records) {
  if (!c || !out || !n_records || max_records < 0) return VSF_ERR_BAD_ARG;
  if (!c->d_ktrace) return fail(c, VSF_ERR_STATE, "no kernel timeline recorded (vsf_set_engine flag 32)");
  cudaSetDevice(c->device);
  VSF_CUDA(c, cudaStreamSynchronize(c->stream));
  const int n = int(std::min<long long>(std::min<long long>(max_records, kKtracePoses), c->ktrace_n));
  VSF_CUDA(c, cudaMemcpy(out, c->d_ktrace, size_t(n) * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  *n_records = n;
  return VSF_OK;
}

extern "C" int vsf_probe_pipe(vsf_ctx* c, int kind, int iters, double* ops_per_second) {
  if (!c || !ops_per_second || iters < 1 || kind < 0 || kind > 6) return VSF_ERR_BAD_ARG;
  cudaSetDevice(c->device);
  const int blocks = c->sm_count * 2, threads = 1024;
  VSF_CUDA(c, launch_probe(kind, c->d_sink, iters, blocks, threads, c->stream));  // warm-up
  VSF_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  VSF_CUDA(c, launch_probe(kind, c->d_sink, iters, blocks, threads, c->stream));
  VSF_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  VSF_CUDA(c, cudaEventSynchronize(c->ev1));
  float ms = 0;
  VSF_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  const double ops = double(blocks) * threads * 8.0 * iters * probe_ops_per_step(kind);
  *ops_per_second = ops / (double(ms) * 1e-3);
  return VSF_OK;
}
