// Kernel 1: batched brute-force k=2 Hamming kNN + Lowe ratio test + ordered
// compaction, for sm_100a.
//
// Replaces, for a whole batch of (query frame, train frame) problems in one
// launch, what the reference does per pair with
//   matcher_->knnMatch(q, t, matches, 2)            src/slam_frontend.cc:525-527
//   if (dist1 < nn_match_ratio * dist2) keep first   src/slam_frontend.cc:529-536
// and the loop over the sliding window                src/slam_frontend.cc:424-434.
//
// Mapping
//   grid  = (query blocks, problems, train splits)
//   CTA   = 8 consumer warps + 1 TMA producer warp
//   each consumer thread keeps R query descriptors in registers (32*R queries
//   per CTA); the CTA's train rows stream through a 4-stage shared-memory ring
//   filled by 1-D TMA bulk copies (cp.async.bulk + mbarrier); the rows of a tile
//   are dealt round-robin to the 8 warps, every lane of a warp reads the same
//   train row (LDS.128 broadcast), XOR + POPC on the integer pipe, and the
//   running top-2 lives in registers as packed (distance<<22 | trainIdx) keys
//   so min/max implement the lexicographic (distance, trainIdx) order.
//   Warps are merged through shared memory, train splits through global memory
//   (last-arriver pattern), and the last CTA of a problem compacts the ratio
//   survivors in ascending query order.
#include "knn2_tail.cuh"
#include "vsf_device.cuh"

namespace vsf {

template <int WORDS, int MODE>
__device__ __forceinline__ uint32_t hamming_row(const uint32_t* q, const uint32_t* t) {
  uint32_t d = hamming256<MODE>(q, t);
  if (WORDS == 16) d += hamming256<MODE>(q + 8, t + 8);
  return d;
}

// Kernel variants (selected with vsf_set_tuning):
//   0: dedicated TMA producer warp (9 warps), 4-stage ring, rows unrolled by 2
//   1: no producer warp (8 warps; lane 0 of warp 0 issues the TMA refills one tile
//      behind the consumers), 3-stage ring, rows unrolled by 2 -> more compute warps per SM
//   2: as 1, rows unrolled by 4
//   3: as 0, rows unrolled by 4
template <int VAR>
struct Variant {
  static constexpr bool kProducerWarp = (VAR == 0 || VAR == 3);
  static constexpr int kUnroll = (VAR == 2 || VAR == 3) ? 4 : 2;
  static constexpr int kRing = kProducerWarp ? kStages : 3;
  static constexpr int kThreads = (kConsumerWarps + (kProducerWarp ? 1 : 0)) * 32;
};

template <int WORDS, int R, int MODE, int VAR>
__global__ void __launch_bounds__(Variant<VAR>::kThreads)
knn2_kernel(const __grid_constant__ KnnBatch batch) {
  using V = Variant<VAR>;
  constexpr int QB = 32 * R;
  constexpr int ROW_BYTES = WORDS * 4;
  constexpr int TILE_ROWS = kTileBytes / ROW_BYTES;
  constexpr int RING = V::kRing;

  __shared__ __align__(128) uint8_t s_tile[RING][kTileBytes];
  __shared__ __align__(8) uint64_t s_full[RING];
  __shared__ __align__(8) uint64_t s_empty[RING];
  __shared__ uint32_t s_merge[kConsumerWarps][QB][2];
  __shared__ TailSmem s_tail;

  const KnnProblem& P = batch.p[blockIdx.y];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  int nq = P.nq, nt = P.nt;
  if (P.nq_dev) nq = min(nq, *P.nq_dev);
  if (P.nt_dev) nt = min(nt, *P.nt_dev);

  if (nq <= 0) {
    if (blockIdx.x == 0 && blockIdx.z == 0 && tid == 0) {
      *P.match_count = 0;
      if (batch.host_counts) batch.host_counts[P.region] = 0;
    }
    return;
  }
  const int qb = blockIdx.x;
  const int q0 = qb * QB;
  if (q0 >= nq) return;
  const int nqb = (nq + QB - 1) / QB;

  const int S = batch.split;
  const int z = blockIdx.z;
  const int rows_per_split = (nt + S - 1) / S;
  const int t_begin = min(nt, z * rows_per_split);
  const int t_end = min(nt, t_begin + rows_per_split);
  const int ntiles = (t_end - t_begin + TILE_ROWS - 1) / TILE_ROWS;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < RING; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], kConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  uint32_t m1[R], m2[R];
#pragma unroll
  for (int r = 0; r < R; ++r) m1[r] = m2[r] = kKeySentinel;

  const uint8_t* t_src = reinterpret_cast<const uint8_t*>(P.t) + size_t(t_begin) * ROW_BYTES;
  auto issue_tile = [&](int k) {
    const int s = k % RING;
    const int rows = min(TILE_ROWS, t_end - t_begin - k * TILE_ROWS);
    const uint32_t bytes = uint32_t(rows) * ROW_BYTES;
    mbar_arrive_expect_tx(&s_full[s], bytes);
    tma_load_1d(s_tile[s], t_src + size_t(k) * kTileBytes, bytes, &s_full[s]);
  };

  if (V::kProducerWarp && warp == kConsumerWarps) {
    // ---------------- TMA producer: one elected lane ----------------
    if (lane == 0) {
      for (int k = 0; k < ntiles; ++k) {
        const int u = k / RING;
        if (u > 0) mbar_wait(&s_empty[k % RING], (u - 1) & 1);
        issue_tile(k);
      }
    }
    __syncwarp();
  } else {
    // ---------------- consumers ----------------
    if (!V::kProducerWarp && tid == 0) {
      for (int k = 0; k < min(ntiles, RING - 1); ++k) issue_tile(k);   // prologue: fill the ring
    }
    uint32_t qreg[R][WORDS];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int q = q0 + r * 32 + lane;
      const uint4* src = reinterpret_cast<const uint4*>(P.q + size_t(q) * WORDS);
#pragma unroll
      for (int v = 0; v < WORDS / 4; ++v) {
        uint4 w = make_uint4(0, 0, 0, 0);
        if (q < nq) w = __ldg(src + v);
        qreg[r][4 * v + 0] = w.x;
        qreg[r][4 * v + 1] = w.y;
        qreg[r][4 * v + 2] = w.z;
        qreg[r][4 * v + 3] = w.w;
      }
    }
    for (int k = 0; k < ntiles; ++k) {
      const int s = k % RING;
      const int u = k / RING;
      if (!V::kProducerWarp && tid == 0) {
        // refill one tile behind: tile k+RING-1 goes into the slot tile k-1 used, which
        // every warp released when it finished tile k-1
        const int kn = k + RING - 1;
        if (kn < ntiles) {
          if (k > 0) mbar_wait(&s_empty[(k - 1) % RING], ((k - 1) / RING) & 1);
          issue_tile(kn);
        }
      }
      if (!V::kProducerWarp) __syncwarp();
      mbar_wait(&s_full[s], u & 1);
      const int rows = min(TILE_ROWS, t_end - t_begin - k * TILE_ROWS);
      const uint32_t row_base = uint32_t(t_begin + k * TILE_ROWS);
      const uint4* tile = reinterpret_cast<const uint4*>(s_tile[s]);
#pragma unroll (V::kUnroll)
      for (int row = warp; row < rows; row += kConsumerWarps) {
        uint32_t tw[WORDS];
#pragma unroll
        for (int v = 0; v < WORDS / 4; ++v) {
          const uint4 w = tile[row * (WORDS / 4) + v];  // same address in every lane: broadcast
          tw[4 * v + 0] = w.x;
          tw[4 * v + 1] = w.y;
          tw[4 * v + 2] = w.z;
          tw[4 * v + 3] = w.w;
        }
        const uint32_t idx = row_base + uint32_t(row);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const uint32_t d = hamming_row<WORDS, MODE>(qreg[r], tw);
          const uint32_t key = (d << kIdxBits) + idx;
          top2_insert(m1[r], m2[r], key);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[s]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      s_merge[warp][r * 32 + lane][0] = m1[r];
      s_merge[warp][r * 32 + lane][1] = m2[r];
    }
  }
  __syncthreads();

  // ---------------- merge the 8 warps (thread tid < QB owns query q0+tid) ----------------
  uint32_t k1 = kKeySentinel, k2 = kKeySentinel;
  if (tid < QB) {
#pragma unroll
    for (int w = 0; w < kConsumerWarps; ++w) top2_merge(k1, k2, s_merge[w][tid][0], s_merge[w][tid][1]);
  }

  const int row = P.row0 + q0 + tid;  // row of this thread's query in knn_out / partial
  if (S > 1) {
    // ---------------- merge the train splits: last arriver does it ----------------
    if (tid < QB) {
      batch.partial[size_t(row) * S + z] = make_uint2(k1, k2);
      __threadfence();
    }
    __syncthreads();
    if (tid == 0) {
      const unsigned prev = atomicAdd(&batch.qblock_arrivals[P.qb0 + qb], 1u);
      const int last = (prev == unsigned(S - 1));
      if (last) batch.qblock_arrivals[P.qb0 + qb] = 0u;  // self-reset for the next launch
      s_tail.flag = last;
    }
    __syncthreads();
    if (!s_tail.flag) return;
    __threadfence();
    if (tid < QB) {
      k1 = k2 = kKeySentinel;
      for (int s = 0; s < S; ++s) {
        const uint2 p = __ldcg(&batch.partial[size_t(row) * S + s]);
        top2_merge(k1, k2, p.x, p.y);
      }
    }
  }

  // ---------------- finalize: unpack, ratio test, ordered compaction ----------------
  finalize_and_compact<QB, V::kThreads>(batch, P, blockIdx.y, qb, nqb, nq, k1, k2, s_tail);
}

// ---------------------------------------------------------------------------------------------
template <int WORDS, int R, int MODE, int VAR>
static cudaError_t launch_one(const KnnBatch& batch, int max_qblocks, cudaStream_t stream) {
  dim3 grid(max_qblocks, batch.num_problems, batch.split);
  knn2_kernel<WORDS, R, MODE, VAR><<<grid, Variant<VAR>::kThreads, 0, stream>>>(batch);
  return cudaGetLastError();
}

template <int WORDS, int R, int MODE>
static cudaError_t launch_var(const KnnBatch& b, int var, int mq, cudaStream_t s) {
  switch (var) {
    case 0: return launch_one<WORDS, R, MODE, 0>(b, mq, s);
    case 1: return launch_one<WORDS, R, MODE, 1>(b, mq, s);
    case 2: return launch_one<WORDS, R, MODE, 2>(b, mq, s);
    case 3: return launch_one<WORDS, R, MODE, 3>(b, mq, s);
    default: return cudaErrorInvalidValue;
  }
}

template <int WORDS, int R>
static cudaError_t launch_mode(const KnnBatch& b, int mode, int var, int mq, cudaStream_t s) {
  switch (mode) {
    case 0: return launch_var<WORDS, R, 0>(b, var, mq, s);
    case 2: return launch_var<WORDS, R, 2>(b, var, mq, s);
    case 3: return launch_var<WORDS, R, 3>(b, var, mq, s);
    default: return cudaErrorInvalidValue;
  }
}

// words: 8 or 16; R: queries per thread (1, 2 or 4; 4 only for 8 words).
// max_qblocks: ceil(max nq / (32*R)) over the batch's problems.
cudaError_t launch_knn2(const KnnBatch& batch, int words, int R, int mode, int variant,
                        int max_qblocks, cudaStream_t stream) {
  if (batch.num_problems <= 0 || max_qblocks <= 0) return cudaSuccess;
  if (words == 8) {
    if (R == 1) return launch_mode<8, 1>(batch, mode, variant, max_qblocks, stream);
    if (R == 2) return launch_mode<8, 2>(batch, mode, variant, max_qblocks, stream);
    if (R == 4) return launch_mode<8, 4>(batch, mode, variant, max_qblocks, stream);
  } else if (words == 16) {
    if (R == 1) return launch_mode<16, 1>(batch, mode, variant, max_qblocks, stream);
    if (R == 2) return launch_mode<16, 2>(batch, mode, variant, max_qblocks, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace vsf
