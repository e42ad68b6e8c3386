// Shared device-side definitions for libvsf_cuda (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "vsf.h"

namespace vsf {

// ---- packed (distance, trainIdx) keys -------------------------------------
// key = distance << 22 | trainIdx.  Unsigned min over keys is the lexicographic
// minimum of (distance, trainIdx): the lowest train index wins ties for both
// neighbours, which is cv::BFMatcher's rule, and the reduction is independent
// of the order in which train rows are visited (so the train dimension can be
// split over warps and CTAs).  distance <= 512 (64-byte descriptors) needs 10
// bits; 22 bits of index allow 4M train rows.
constexpr int kIdxBits = 22;
constexpr uint32_t kIdxMask = (1u << kIdxBits) - 1u;
constexpr uint32_t kKeySentinel = 0xFFFFFFFFu;
constexpr int kMaxRows = 1 << kIdxBits;

// ---- kernel 1 geometry -----------------------------------------------------
constexpr int kConsumerWarps = 8;                 // train rows of a tile are dealt to these
constexpr int kKnnThreads = (kConsumerWarps + 1) * 32;  // + 1 TMA producer warp
constexpr int kStages = 4;                        // smem ring depth
constexpr int kTileBytes = 8192;                  // per stage (256 rows of 32 B)
constexpr int kMaxProblems = 40;                  // window (<= 38) + stereo pair

// One query-frame x train-frame matching problem.  A launch processes a batch
// of them (the whole sliding window, or stereo L->R / R->L) at once.
struct KnnProblem {
  const uint32_t* q;       // [nq][WORDS]
  const uint32_t* t;       // [nt][WORDS]
  const int* nq_dev;       // when non-null the row counts are read from the
  const int* nt_dev;       //   device (compacted frames), nq/nt are upper bounds
  vsf_dmatch* matches;     // compacted survivors of the ratio test, query order
  int* match_count;
  int nq, nt;
  int row0;                // first row of this problem in knn_out / partial
  int qb0;                 // first slot of this problem in the per-query-block counters
  int region;              // index of `matches` among the ctx's match regions (host mirror slot)
  int* host_count;         // when non-null: the survivor count is also stored here (mapped host memory) instead
                           //   of KnnBatch::host_counts[region] (batches that span several frames' buffers)
};

struct KnnBatch {
  int num_problems;
  int split;               // number of train-dimension splits (gridDim.z)
  double ratio;            // nn_match_ratio, compared in double like the reference
  int exact_second;        // tensor engine's refine: 1 = rescan the second-best bucket too, so that the
                           // second neighbour's INDEX is exact (vsf_knn2); 0 = take its distance from the
                           // bucket's exact maximum dot, which is all the ratio test needs
  uint4* knn_out;          // [rows] {idx0, idx1, d0, d1}
  uint2* partial;          // [rows][split] partial top-2 keys (split > 1)
  unsigned* qblock_arrivals;  // [qblocks]   self-resetting counters
  unsigned* qblock_pass;      // [qblocks]   ratio survivors per query block
  unsigned* problem_arrivals; // [problems]  self-resetting counters
  // Optional zero-copy mirror: when host_matches != nullptr the survivors and their count are
  // also stored straight into mapped pinned host memory (region r at host_matches +
  // r * host_region_stride, count at host_counts[r]), so a host-API call needs one stream
  // synchronisation and no device-to-host copy.
  vsf_dmatch* host_matches;
  int* host_counts;
  int host_region_stride;
  long long* ktrace;   // engine flag 32: [8][2] earliest start / latest end (globaltimer ns) of expand, distance, refine, compaction, refine-after-wait; 5..7 spare
  KnnProblem p[kMaxProblems];
};

// ---- tensor-core engine (knn2_tc_kernel.cu) -------------------------------------------------
constexpr int kTcQ = 256;            // queries per work unit (two M=128 UMMA tiles)
constexpr int kTcTileRows = 256;     // train rows per staged tile (N of one UMMA)
#ifndef VSF_TC_BUCKET
#define VSF_TC_BUCKET 16
#endif
constexpr int kTcBucket = VSF_TC_BUCKET;  // train rows per selection bucket (8, 16 or 32)
constexpr int kTcStages = 2;         // train tiles in flight in shared memory
constexpr int kTcRowBytes = 256;     // one 256-bit descriptor expanded to +-1 bytes
constexpr int kTcABytes = kTcQ * kTcRowBytes;
constexpr int kTcBBytes = kTcTileRows * kTcRowBytes;
constexpr int kTcMaxTrains = 2;      // distinct train frames per batch that can be expanded

// Several frames expanded to +-1 images by one launch (a group of poses).
constexpr int kMaxPoseGroup = 8;
struct ExpandMulti {
  const uint32_t* src[kMaxPoseGroup];
  uint8_t* out[kMaxPoseGroup];
  int frames;
  int nt;                            // rows of every frame
};

// The work of a launch is the flattened list of (256-query block, piece) slots, query block
// major, where a piece is tiles_per_piece consecutive train tiles and a block has `pieces` of
// them: T = blocks * pieces slots.  CTA i of the G CTAs of the persistent kernel owns the
// contiguous range [i*T/G, (i+1)*T/G) ("stream-K"): every SM gets the same number of pieces to
// within one, whatever the frame sizes are.  A range covers runs of pieces ("segments") of one
// or more query blocks; a block is covered by the consecutive CTAs owner(first slot) ..
// owner(last slot), and segment s of a block stores its partial top-2 bucket keys in slot s of
// the block's rows.  The host picks the granularity from the shape of the launch:
//   one tile per piece           the general case (frame-to-window matching at C4 size)
//   G = T, one piece per CTA     small launches that cannot fill the machine with whole blocks
//   a few long pieces per block  very large launches: the CTAs then walk the train tiles nearly
//                                in step, which keeps the tiles they fetch hot in L2
struct TcBatch {
  const uint8_t* t_exp[kMaxProblems];  // expanded train image of every problem
  int qb_begin[kMaxProblems + 1];      // prefix sum of 256-query blocks
  int tiles;                           // train tiles per query block (largest train frame)
  int pieces;                          // slots per query block
  int tiles_per_piece;
  int grid;                            // G: CTAs of knn2_tc_kernel
  int slots;                           // partial segments stored per query row (max over blocks)
  long long total;                     // T = qb_begin[num_problems] * pieces
  int unit_q;                          // 64-byte engine: queries per block (128: one CTA per unit, 256: a CTA pair per unit)
  int flags;                           // bring-up knobs (timing experiments; results invalid): 2 = skip the bucket reduction, 4 = skip the TMEM loads
  long long* trace;                    // flag 16 (builds with -DVSF_TC_TRACE): per-CTA timeline, kTcTraceSlots values per CTA
  // A sequence of batches (vsf_window_match_block_device, vsf_window_run_sequence).
  // early != 0: everything this launch reads was complete before its stream predecessor (the
  // previous launch's finish kernel) let it launch, and its partial keys go to another buffer than
  // the one that kernel reads: it starts without waiting for the predecessor and only waits for
  // it just before it exits (so that "complete" still implies "everything before it complete").
  int early;
  // late != 0: the CTAs let the stream successor (the finish kernel) launch only when their TMA
  // producer has issued its last tile load, not at the start.  A finish kernel launched at the
  // start sits in the room a distance CTA leaves on its SM for as long as the distance kernel runs,
  // and the expansion kernels of the upload stream (host-buffer sequences) find no room.
  int late;
};
constexpr int kTcTraceSlots = 16;


// Arguments of knn2_tc_finish_kernel (refine + ordered compaction in one kernel, see there).
struct FinishArgs {
  // Device-side launch state, three words: [0] ticket = next logical block, [1] CTAs that have
  // finished, [2] epoch of the launch that is running (or will run next).  The last CTA to
  // finish resets [0] and [1] and advances [2]; nothing depends on host-side counters, so a
  // launch captured in a CUDA graph can be replayed.  Launches that share a state (and its
  // look-back words) never overlap: each starts after its predecessor on the state completed.
  unsigned long long* state;
  unsigned long long* flags;         // [32-query units of the batch, KnnProblem::qb0 + unit] (epoch << 16 | survivors of a block)
  int nqb;                           // query blocks (of 64 / 128 / 256 queries) per problem (grid = nqb * num_problems)
};



// CTA that owns slot x = the largest i with floor(i*T/G) <= x
__host__ __device__ __forceinline__ int tc_owner(long long x, long long T, int G) {
  const unsigned long long num = (unsigned long long)(x + 1) * (unsigned long long)G - 1ull;
  if (((num | (unsigned long long)T) >> 32) == 0) return int(unsigned(num) / unsigned(T));   // the usual case: one 32-bit divide
  return int(num / (unsigned long long)T);
}
// number of segments (partial slots in use) of query block gqb
__host__ __device__ __forceinline__ int tc_block_segments(const TcBatch& tc, int gqb) {
  const long long s0 = (long long)gqb * tc.pieces;
  return tc_owner(s0 + tc.pieces - 1, tc.total, tc.grid) - tc_owner(s0, tc.total, tc.grid) + 1;
}

// ---- PTX helpers -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  // make the initialised barriers visible to the async (TMA) proxy
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      " selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier
// (SASS: UBLKCP).  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor in the stream is still running; pdl_wait() blocks
// until the predecessor has completed and its writes are visible, pdl_launch_dependents() lets
// the successor start its own prologue.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void ktrace_start(long long* kt, int k) {
  if (kt) atomicMin(reinterpret_cast<long long*>(kt + 2 * k), globaltimer_ns());
}
__device__ __forceinline__ void ktrace_minmax(long long* kt, int k, long long v) {
  if (kt) {
    atomicMin(reinterpret_cast<long long*>(kt + 2 * k), v);
    atomicMax(reinterpret_cast<long long*>(kt + 2 * k + 1), v);
  }
}
__device__ __forceinline__ void ktrace_end(long long* kt, int k) {
  if (kt) atomicMax(reinterpret_cast<long long*>(kt + 2 * k + 1), globaltimer_ns());
}

__device__ __forceinline__ uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint32_t umax32(uint32_t a, uint32_t b) { return a > b ? a : b; }

// top-2 insert of one key into (m1 <= m2)
__device__ __forceinline__ void top2_insert(uint32_t& m1, uint32_t& m2, uint32_t key) {
  m2 = umin32(m2, umax32(m1, key));
  m1 = umin32(m1, key);
}
// top-2 merge of two sorted pairs
__device__ __forceinline__ void top2_merge(uint32_t& m1, uint32_t& m2, uint32_t b1, uint32_t b2) {
  const uint32_t lo = umin32(m1, b1);
  const uint32_t hi = umax32(m1, b1);
  m2 = umin32(hi, umin32(m2, b2));
  m1 = lo;
}

// 3-input bitwise primitives; the compiler maps each to one LOP3.LUT
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
  return (a & b) | (a & c) | (b & c);
}

// Hamming distance of two 256-bit blocks held as 8 words each.
//   MODE 0: 8 XOR + 8 POPC (the naive count the roofline is quoted against)
//   MODE 2: carry-save adders fold 8 words to 2 "ones" + 3 "twos": 5 POPC, 14 LOP3
//   MODE 3: full Harley-Seal tree: 4 POPC, 20 LOP3
template <int MODE>
__device__ __forceinline__ uint32_t hamming256(const uint32_t* q, const uint32_t* t) {
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = q[i] ^ t[i];
  if (MODE == 0) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) d += __popc(x[i]);
    return d;
  } else {
    const uint32_t sa = xor3(x[0], x[1], x[2]), ca = maj3(x[0], x[1], x[2]);
    const uint32_t sb = xor3(x[3], x[4], x[5]), cb = maj3(x[3], x[4], x[5]);
    const uint32_t sc = xor3(x[6], x[7], sa), cc = maj3(x[6], x[7], sa);
    if (MODE == 2) {
      const uint32_t ones = __popc(sb) + __popc(sc);
      const uint32_t twos = __popc(ca) + __popc(cb) + __popc(cc);
      return ones + 2u * twos;
    } else {
      const uint32_t ones = sb ^ sc, cd = sb & sc;
      const uint32_t ts = xor3(ca, cb, cc), fa = maj3(ca, cb, cc);
      const uint32_t twos = ts ^ cd, fb = ts & cd;
      return __popc(ones) + 2u * __popc(twos) + 4u * (__popc(fa) + __popc(fb));
    }
  }
}

}  // namespace vsf
