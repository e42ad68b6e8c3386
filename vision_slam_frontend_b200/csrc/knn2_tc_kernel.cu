// Kernel 1, tensor-core engine: the same k=2 Hamming kNN + ratio + ordered compaction as
// knn2_kernel.cu (cv::BFMatcher(NORM_HAMMING)::knnMatch at src/slam_frontend.cc:525-527 and
// the ratio test at :529-536), with the distance matrix computed by the 5th-generation
// tensor cores instead of the POPC pipe.
//
// Idea.  Expand every descriptor bit b to the 8-bit value (+1 if b == 0, -1 if b == 1).  For
// two 256-bit descriptors the dot product of the expansions is
//     dot = (#equal bits) - (#different bits) = 256 - 2 * hamming,
// an exact small integer, so the largest dot is the smallest Hamming distance.  A block of
// 256 query rows x 256 train rows x K = 256 is then two M=128, N=256 UMMA tiles
// (tcgen05.mma, kind::i8 with int32 accumulators or kind::f8f6f4 with exact fp32
// accumulators), with the accumulators in tensor memory.
//
// Selection without packing an index into every element: the epilogue reads its query's row
// of the accumulator (tcgen05.ld, one TMEM lane per query) and keeps only the MAXIMUM dot of
// every bucket of kTcBucket (16) consecutive train rows (3-input VIMNMX3 on packed int16 pairs),
// then the top-2 BUCKETS per query ordered by (max dot desc, bucket asc).  The best neighbour
// (lowest distance, lowest train index on ties) is always inside the best bucket, and the
// second neighbour is inside the best or the second-best bucket, so the refine kernel
// recomputes exact Hamming distances for just those <= 32 train rows per query on the integer
// pipe (XOR/POPC, lexicographic packed keys) and the result is bit-identical to the POPC engine
// and to OpenCV.
//
// Kernels:
//   expand_train_kernel   packed train rows -> +-1 bytes, written tile by tile in the exact
//                         shared-memory image the UMMA descriptors expect (K-major, no swizzle:
//                         [16 K-chunks][256 rows][16 B]), so one 64 KB 1-D TMA bulk copy
//                         (cp.async.bulk, SASS UBLKCP) stages a tile.
//   knn2_tc_kernel        persistent, warp-specialised, one CTA per SM: warps 0-15 = query
//                         expansion into smem + TMEM epilogue, warp 16 = TMA producer, warp 17 =
//                         MMA issuer (one elected thread).  The (256-query block, train tile)
//                         slots of the launch are shared out as equal contiguous ranges (TcBatch).
//   knn2_tc_refine_kernel exact top-2 inside the candidate buckets, ratio test.
//   knn2_compact_kernel   ratio survivors in ascending queryIdx order.
// launch_knn2_tc chains the four on one stream with programmatic dependent launch.
#include <atomic>
#include <cstdlib>
#include <utility>

#include "knn2_tail.cuh"
#include "tc_ptx.cuh"
#include "vsf_device.cuh"

namespace vsf {

// bit j of b16 -> byte j of the result: 0 -> +1, 1 -> -1 in the operand type
template <bool I8>
__device__ __forceinline__ uint32_t expand4(uint32_t nibble) {
  // bit i of the nibble lands on bit 7 of byte i
  const uint32_t sign = (nibble * 0x10204080u) & 0x80808080u;
  if (I8) {
    // 0x80 -> 0xFF (-1), 0x00 -> 0x01 (+1)
    return (sign | (sign - (sign >> 7))) | 0x01010101u;
  } else {
    // e4m3: +1.0 = 0x38, -1.0 = 0xB8
    return sign | 0x38383838u;
  }
}
template <bool I8>
__device__ __forceinline__ uint4 expand16(uint32_t b16) {
  uint4 r;
  r.x = expand4<I8>(b16 & 0xFu);
  r.y = expand4<I8>((b16 >> 4) & 0xFu);
  r.z = expand4<I8>((b16 >> 8) & 0xFu);
  r.w = expand4<I8>((b16 >> 12) & 0xFu);
  return r;
}

// One (row, group of 4 K-chunks) item of a frame's image (the unit of work of every expansion).
template <bool I8>
__device__ __forceinline__ void expand_item(const uint32_t* __restrict__ t, int nt, uint8_t* __restrict__ out, int row, int cg) {
  uint2 w = make_uint2(0u, 0u);
  const bool live = row < nt;
  if (live) w = __ldg(reinterpret_cast<const uint2*>(t + size_t(row) * 8) + cg);
  const int tile = row / kTcTileRows, r = row % kTcTileRows;
  uint8_t* base = out + size_t(tile) * kTcBBytes + size_t(r) * 16;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t word = (c < 2) ? w.x : w.y;
    const uint32_t b16 = (word >> (16 * (c & 1))) & 0xFFFFu;
    uint4 v = live ? expand16<I8>(b16) : make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(base + size_t(4 * cg + c) * (kTcTileRows * 16)) = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Train expansion.  One thread = one (row, group of 4 K-chunks).  Rows >= nt of the last tile
// are written as zeros (their columns are masked in the epilogue).
template <bool I8>
__global__ void __launch_bounds__(256)
expand_train_kernel(const uint32_t* __restrict__ t, int nt_bound, const int* __restrict__ nt_dev,
                    uint8_t* __restrict__ out, long long* kt) {
  if (threadIdx.x == 0) ktrace_start(kt, 0);
  pdl_wait();
  pdl_launch_dependents();
  const int rows_pad = (nt_bound + kTcTileRows - 1) / kTcTileRows * kTcTileRows;
  int nt = nt_bound;
  if (nt_dev) nt = min(nt, *nt_dev);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = idx % rows_pad;
  const int cg = idx / rows_pad;  // 0..3: K-chunks 4*cg .. 4*cg+3  (words 2*cg, 2*cg+1)
  if (cg >= 4) return;   // whole CTAs at most: blockDim divides rows_pad
  uint2 w = make_uint2(0u, 0u);
  const bool live = row < nt;
  if (live) w = __ldg(reinterpret_cast<const uint2*>(t + size_t(row) * 8) + cg);
  const int tile = row / kTcTileRows, r = row % kTcTileRows;
  uint8_t* base = out + size_t(tile) * kTcBBytes + size_t(r) * 16;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t word = (c < 2) ? w.x : w.y;
    const uint32_t b16 = (word >> (16 * (c & 1))) & 0xFFFFu;
    uint4 v = live ? expand16<I8>(b16) : make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(base + size_t(4 * cg + c) * (kTcTileRows * 16)) = v;
  }
  if (threadIdx.x == 0) ktrace_end(kt, 0);
}

// The same for the frames of a group of poses, blockIdx.y = frame.
template <bool I8>
__global__ void __launch_bounds__(256)
expand_train_multi_kernel(const __grid_constant__ ExpandMulti em) {
  pdl_wait();
  pdl_launch_dependents();
  const int rows_pad = (em.nt + kTcTileRows - 1) / kTcTileRows * kTcTileRows;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows_pad * 4) return;
  expand_item<I8>(em.src[blockIdx.y], em.nt, em.out[blockIdx.y], idx % rows_pad, idx / rows_pad);
}

// ---------------------------------------------------------------------------------------------
// TcBatch::flags 2 / 4 (skip the bucket reduction / the TMEM loads: timing experiments whose
// results are invalid) only exist in builds with -DVSF_TC_BRINGUP
#ifdef VSF_TC_BRINGUP
constexpr bool kBringup = true;
#else
constexpr bool kBringup = false;
#endif
constexpr int kTcEpiWarps = 16;
constexpr int kTcColSplit = 2;                      // epilogue warps per (accumulator half, lane quarter)
constexpr int kTcEpiCols = kTcTileRows / kTcColSplit;  // columns of a tile one epilogue warp reduces
constexpr int kTcThreads = (kTcEpiWarps + 2) * 32;  // + TMA producer warp + MMA warp
constexpr int kTcBarriers = 2 * kTcStages + 6;      // full/empty per stage, tfull/tempty/aready per half
constexpr int kTcSmemBytes = kTcABytes + kTcStages * kTcBBytes + kTcBarriers * 8 + 16 + 128;
constexpr int kTcTmemCols = 512;                    // 2 accumulator tiles of 128 lanes x 256 columns
constexpr int kBucketIdBits = 20;
constexpr int kBucketIdMask = (1 << kBucketIdBits) - 1;
constexpr int kTcKeySentinel = int(0x80000000u);    // INT_MIN: "no bucket"

struct TcUnit {
  int problem, qb, slot;
  int nq, nt;
  int q0;
  int t_begin, t_end, ntiles;
  int len;     // tile slots of the CTA's range this segment uses up
  bool skip;   // no queries in this block
};

// Walk over the segments of this CTA's slot range [x, end) (slots = pieces of
// tc.tiles_per_piece train tiles, see TcBatch).  Only the first segment can start in the middle
// of a query block (two divisions, done once, before the kernel waits for its stream
// predecessor); every later one starts at piece 0 of the next block and is the block's first
// segment, so stepping costs a few integer instructions.
struct TcWalk {
  long long x, end;
  int gqb, t0, slot, p;
};

__device__ __forceinline__ TcWalk walk_begin(const TcBatch& tc) {
  TcWalk w;
  w.p = 0;
  w.x = (long long)blockIdx.x * tc.total / tc.grid;
  w.end = (long long)(blockIdx.x + 1) * tc.total / tc.grid;
  const long long gqb = w.x / tc.pieces;
  w.gqb = int(gqb);
  w.t0 = int(w.x - gqb * tc.pieces);
  w.slot = int(blockIdx.x) - tc_owner(gqb * tc.pieces, tc.total, tc.grid);
  return w;
}

__device__ __forceinline__ TcUnit walk_unit(const KnnBatch& batch, const TcBatch& tc, const TcWalk& w) {
  TcUnit U;
  U.len = int(min((long long)(tc.pieces - w.t0), w.end - w.x));
  int p = w.p;
  while (w.gqb >= tc.qb_begin[p + 1]) ++p;
  U.problem = p;
  U.qb = w.gqb - tc.qb_begin[p];
  U.slot = w.slot;
  const KnnProblem& P = batch.p[p];
  int nq = P.nq, nt = P.nt;
  if (P.nq_dev) nq = min(nq, *P.nq_dev);
  if (P.nt_dev) nt = min(nt, *P.nt_dev);
  U.nq = nq;
  U.nt = nt;
  U.q0 = U.qb * kTcQ;
  U.skip = U.q0 >= nq;
  // pieces -> train rows (the last piece of a block may be short)
  const int tile0 = min(tc.tiles, w.t0 * tc.tiles_per_piece);
  const int tile1 = min(tc.tiles, (w.t0 + U.len) * tc.tiles_per_piece);
  U.t_begin = min(nt, tile0 * kTcTileRows);
  U.t_end = min(nt, tile1 * kTcTileRows);
  U.ntiles = (U.t_end - U.t_begin + kTcTileRows - 1) / kTcTileRows;
  return U;
}

__device__ __forceinline__ void walk_next(TcWalk& w, const TcUnit& U) {
  w.x += U.len;
  w.gqb += 1;
  w.t0 = 0;
  w.slot = 0;
  w.p = U.problem;
}
__device__ __forceinline__ bool walk_more(const TcWalk& w) { return w.x < w.end; }

// Timeline slots (flag 16, SM clock cycles unless noted): 0 globaltimer ns at entry, 1 entry,
// 2 barriers/TMEM ready, 3 predecessor complete (griddepcontrol.wait returned), 4 first TMA issued,
// 5 first query half expanded, 6 first train tile landed, 7 first MMA issued, 8 last MMA
// committed, 9 last accumulator complete, 10 last partial keys written, 11 CTA done,
// 12 cycles the MMA warp waited for expanded queries, 13 ... for train tiles, 14 segments,
// 15 globaltimer ns at the end.
__device__ __forceinline__ long long tc_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_TRACE(slot) do { if (tr) tr[slot] = clock64(); } while (0)
#define TC_TRACE_ONCE(slot) do { if (tr && tr[slot] == 0) tr[slot] = clock64(); } while (0)

template <bool I8>
__global__ void __launch_bounds__(kTcThreads, 1)
knn2_tc_kernel(const __grid_constant__ KnnBatch batch, const __grid_constant__ TcBatch tc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kTcABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcABytes + kTcStages * kTcBBytes);
  uint64_t* full = bars;                      // [kTcStages] TMA -> MMA
  uint64_t* empty = bars + kTcStages;         // [kTcStages] MMA -> TMA
  uint64_t* tfull = bars + 2 * kTcStages;     // [2] MMA -> epilogue (per accumulator half)
  uint64_t* tempty = tfull + 2;               // [2] epilogue -> MMA
  uint64_t* aready = tempty + 2;              // [2] query half expanded in smem
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + kTcBarriers);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  // (compiled out by default: even switched off at run time the stamps cost 6 % of the kernel)
#ifdef VSF_TC_TRACE
  long long* const tr = tc.trace ? tc.trace + size_t(blockIdx.x) * kTcTraceSlots : nullptr;
#else
  long long* const tr = nullptr;
#endif
  if (tr && tid == 0) {
    tr[0] = tc_globaltimer();
    tr[1] = clock64();
  }
  if (tid == 0) ktrace_start(batch.ktrace, 1);
  if (tc.early && !tc.late) pdl_launch_dependents();

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mbar_init(&tfull[h], 1);
      mbar_init(&tempty[h], 4 * kTcColSplit);
      mbar_init(&aready[h], 4 * kTcColSplit);
    }
    mbar_fence_init();
  }
  if (warp == kTcEpiWarps + 1) tc::tmem_alloc<kTcTmemCols>(s_tmem);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  // Everything above overlaps the tail of the previous kernel in the stream, and so does what
  // each role does before its own griddepcontrol.wait below: decoding the first work unit and
  // (epilogue warps) loading + expanding its queries.  Only the expanded train image comes from
  // the stream predecessor (expand_train_kernel); the descriptors and device-side row counts
  // were written by kernels before it, which had completed before the predecessor triggered
  // this launch.
  if (tid == 0) TC_TRACE(2);

  // this CTA's contiguous range of (query block, train tile) slots
  TcWalk wk = walk_begin(tc);

  // K-major, no swizzle: [K-chunk c][row][16 B]; chunk stride = rows * 16, 8-row group stride = 128
  const uint32_t lbo = uint32_t(kTcQ * 16);
  const uint32_t sbo = 128u;

  if (warp == kTcEpiWarps) {
    // ------------------------------ TMA producer ------------------------------
    TcUnit U;
    if (lane == 0 && walk_more(wk)) U = walk_unit(batch, tc, wk);
    if (!tc.early) {
      pdl_wait();               // the expanded train image is complete
      if (!tc.late) pdl_launch_dependents();
    }
    if (lane == 0) {
      uint32_t it = 0;
      while (walk_more(wk)) {
        if (!U.skip) {
          const uint8_t* src = tc.t_exp[U.problem] + size_t(U.t_begin / kTcTileRows) * kTcBBytes;
          for (int k = 0; k < U.ntiles; ++k, ++it) {
            const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);
            mbar_arrive_expect_tx(&full[s], kTcBBytes);
            tma_load_1d(sB + size_t(s) * kTcBBytes, src + size_t(k) * kTcBBytes, kTcBBytes, &full[s]);
            if (it == 0) TC_TRACE(4);
          }
        }
        walk_next(wk, U);
        if (walk_more(wk)) U = walk_unit(batch, tc, wk);
      }
    }
    __syncwarp();
    if (tc.late) pdl_launch_dependents();   // the last tile load is on its way
  } else if (warp == kTcEpiWarps + 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (!tc.early) {
      pdl_wait();
      if (!tc.late) pdl_launch_dependents();
    }
    if (lane == 0) {
      uint32_t it = 0, unit_it = 0, acc_use[2] = {0u, 0u};
      long long wait_a = 0, wait_b = 0, nseg = 0;
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
      while (walk_more(wk)) {
        const TcUnit U = walk_unit(batch, tc, wk);
        walk_next(wk, U);
        ++nseg;
        if (U.skip || U.ntiles == 0) continue;
        for (int k = 0; k < U.ntiles; ++k, ++it) {
          const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
          // a ragged last tile only multiplies the train rows that exist (N a multiple of 16;
          // the epilogue masks per row, the accumulator columns beyond N are never looked at)
          const int valid = min(kTcTileRows, U.t_end - (U.t_begin + k * kTcTileRows));
          const uint32_t idesc = tc::instr_desc(I8, 128, (valid + 15) & ~15);
          if (tr) {
            const long long t0 = clock64();
            mbar_wait(&full[s], ph);
            wait_b += clock64() - t0;
            if (it == 0) tr[6] = clock64();
          } else {
            mbar_wait(&full[s], ph);
          }
          tc::fence_after_sync();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (k == 0) {
              const long long t0 = tr ? clock64() : 0;
              mbar_wait(&aready[h], unit_it & 1u);
              if (tr) wait_a += clock64() - t0;
            }
            mbar_wait(&tempty[h], (acc_use[h] & 1u) ^ 1u);
            tc::fence_after_sync();
            if (it == 0 && h == 0) TC_TRACE(7);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              // one MMA consumes K = 32 bytes = 2 chunks
              const uint64_t ad = tc::smem_desc(a_addr + uint32_t(h) * 128u * 16u + uint32_t(kk) * 2u * (kTcQ * 16), lbo, sbo);
              const uint64_t bd = tc::smem_desc(b_addr + s * kTcBBytes + uint32_t(kk) * 2u * (kTcTileRows * 16), lbo, sbo);
              tc::mma_ss<I8>(tmem_base + uint32_t(h) * kTcTileRows, ad, bd, idesc, kk > 0 ? 1u : 0u);
            }
            tc::commit(&tfull[h]);
            ++acc_use[h];
          }
          tc::commit(&empty[s]);
        }
        ++unit_it;
      }
      if (tr) {
        tr[8] = clock64();
        tr[12] = wait_a;
        tr[13] = wait_b;
        tr[14] = nseg;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ query expansion + epilogue ------------------------------
    // 16 warps: warp = (column half ch) * 8 + (accumulator half h) * 4 + (TMEM lane quarter).
    // Two warps per SM sub-partition work on the same accumulator while the tensor core fills
    // the other one.  An accumulator is handed back to the tensor core as soon as it has been
    // copied to registers (tcgen05.ld + wait), before it is reduced.
    const int h = (warp >> 2) & 1;                 // accumulator half = query rows 128h .. 128h+127
    const int ch = warp >> 3;                      // train columns 128ch .. 128ch+127 of every tile
    const int r = (warp & 3) * 32 + lane;          // TMEM lane = row inside the half
    const uint32_t taddr = tmem_base + (uint32_t((warp & 3) * 32) << 16) + uint32_t(h * kTcTileRows + ch * kTcEpiCols);
    constexpr int NB = kTcEpiCols / kTcBucket;     // buckets per warp per tile
    constexpr int TB = kTcTileRows / kTcBucket;    // buckets per tile
    uint32_t acc_use = 0;

    // the query words of the next unit are fetched one unit ahead
    auto load_query = [&](const TcUnit& V) -> uint4 {
      uint4 w = make_uint4(0u, 0u, 0u, 0u);
      const int q = V.q0 + h * 128 + r;
      if (!V.skip && V.ntiles > 0 && q < V.nq)
        w = __ldg(reinterpret_cast<const uint4*>(batch.p[V.problem].q + size_t(q) * 8) + ch);
      return w;
    };
    // expand half of this thread's query row (K-chunks 8ch .. 8ch+7) into the A image
    auto expand_query = [&](const uint4& qw) {
      const uint32_t words[4] = {qw.x, qw.y, qw.z, qw.w};
      uint8_t* dst = sA + size_t(h * 128 + r) * 16 + size_t(ch * 8) * (kTcQ * 16);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t b16 = (words[c >> 1] >> (16 * (c & 1))) & 0xFFFFu;
        *reinterpret_cast<uint4*>(dst + size_t(c) * (kTcQ * 16)) = expand16<I8>(b16);
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&aready[h]);
      if (tid == 0) TC_TRACE_ONCE(5);
    };
    TcUnit U;
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    bool expanded = false;   // the queries of unit U are already in the A image
    if (walk_more(wk)) {
      U = walk_unit(batch, tc, wk);
      w = load_query(U);
      if (!(U.skip || U.ntiles == 0)) {
        expand_query(w);
        expanded = true;
      }
    }
    if (!tc.early) {
      pdl_wait();               // partial keys of the previous launch have been consumed
      if (!tc.late) pdl_launch_dependents();
    }
    if (tid == 0) TC_TRACE(3);
    while (walk_more(wk)) {
      walk_next(wk, U);                       // wk now stands on the unit after U
      const bool more = walk_more(wk);
      TcUnit Un;
      uint4 wn = make_uint4(0u, 0u, 0u, 0u);
      const KnnProblem& P = batch.p[U.problem];
      const int q = U.q0 + h * 128 + r;
      uint2* part = reinterpret_cast<uint2*>(batch.partial) +
                    (size_t(P.row0 + q) * (tc.slots * kTcColSplit) + U.slot * kTcColSplit + ch);
      if (U.skip || U.ntiles == 0) {
        if (!U.skip && q < U.nq) *part = make_uint2(uint32_t(kTcKeySentinel), uint32_t(kTcKeySentinel));
        if (more) {
          Un = walk_unit(batch, tc, wk);
          wn = load_query(Un);
        }
        U = Un; w = wn;
        continue;
      }
      if (!expanded) expand_query(w);
      expanded = false;
      if (more) {
        Un = walk_unit(batch, tc, wk);
        wn = load_query(Un);
      }
      int m1 = kTcKeySentinel, m2 = kTcKeySentinel;
      const int bucket0 = U.t_begin / kTcBucket + ch * NB;
      for (int k = 0; k < U.ntiles; ++k, ++acc_use) {
        mbar_wait(&tfull[h], acc_use & 1u);
        tc::fence_after_sync();
        if (warp == kTcEpiWarps - 1 && lane == 0) TC_TRACE(9);   // overwritten by every tile: the last one stays
        const int row0 = U.t_begin + k * kTcTileRows + ch * kTcEpiCols;   // first train row of my columns
        const int kbase = kBucketIdMask - (bucket0 + k * TB);
        auto push = [&](int bm, int j) {
          const int key = bm * (1 << kBucketIdBits) + (kbase - j);
          m2 = max(m2, min(m1, key));
          m1 = max(m1, key);
        };
        auto release = [&]() {
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[h]);
          // Last tile of the unit: every MMA that reads this half of the A image has completed
          // (that is what tfull[h] said), so the next unit's queries can go in now and the tensor
          // core restarts while this tile is still being reduced.
          if (k == U.ntiles - 1 && more && !(Un.skip || Un.ntiles == 0)) {
            expand_query(wn);
            expanded = true;
          }
        };
        const bool full_tile = row0 + kTcEpiCols <= U.t_end;   // warp-uniform
        if (kBringup && (tc.flags & 4)) {
          release();
        } else if (I8) {
          // int32 accumulators hold |dot| <= 256: read them as packed int16 pairs
          // (tcgen05.ld ... .pack::16b), 2 columns per register, VIMNMX(3).S16x2 reduction
          constexpr int RPB = kTcBucket / 2;        // registers per bucket
          uint32_t v[2][32];
          tc::tmem_ld_32x32_pack16(taddr, v[0]);
          tc::tmem_ld_32x32_pack16(taddr + 64u, v[1]);
          tc::tmem_ld_wait();
          release();
          if (!(kBringup && (tc.flags & 2))) {
#pragma unroll
            for (int j = 0; j < NB; ++j) {
              uint32_t* b = &v[(j * RPB) / 32][(j * RPB) % 32];
              if (!full_tile) {
                const int valid = U.t_end - (row0 + j * kTcBucket);   // warp-uniform
                if (valid <= 0) continue;
                if (valid < kTcBucket) {
#pragma unroll
                  for (int i = 0; i < RPB; ++i) {
                    if (2 * i >= valid) b[i] = 0x80008000u;
                    else if (2 * i + 1 >= valid) b[i] = (b[i] & 0xFFFFu) | 0x80000000u;
                  }
                }
              }
              uint32_t mm = b[0];
#pragma unroll
              for (int i = 1; i < RPB; ++i) mm = __vmaxs2(mm, b[i]);
              push(max(int(short(mm & 0xFFFFu)), int(mm) >> 16), j);
            }
          }
        } else {
          // fp32 accumulators: four loads of 32 columns, the next one in flight while one is reduced
          constexpr int BPL = 32 / kTcBucket;       // buckets per load
          uint32_t v[2][32];
          tc::tmem_ld_32x32(taddr, v[0]);
#pragma unroll
          for (int l = 0; l < kTcEpiCols / 32; ++l) {
            tc::tmem_ld_wait();
            if (l + 1 < kTcEpiCols / 32) tc::tmem_ld_32x32(taddr + uint32_t((l + 1) * 32), v[(l + 1) & 1]);
            else release();
            if (!(kBringup && (tc.flags & 2))) {
#pragma unroll
              for (int jj = 0; jj < BPL; ++jj) {
                const int j = l * BPL + jj;
                uint32_t* b = &v[l & 1][jj * kTcBucket];
                if (!full_tile) {
                  const int valid = U.t_end - (row0 + j * kTcBucket);
                  if (valid <= 0) continue;
                  if (valid < kTcBucket) {
#pragma unroll
                    for (int i = 0; i < kTcBucket; ++i)
                      if (i >= valid) b[i] = 0xFF800000u;   // -inf
                  }
                }
                float mm = __uint_as_float(b[0]);
#pragma unroll
                for (int i = 1; i < kTcBucket; ++i) mm = fmaxf(mm, __uint_as_float(b[i]));
                push(__float2int_rn(mm), j);
              }
            }
          }
        }
      }
      if (q < U.nq) *part = make_uint2(uint32_t(m1), uint32_t(m2));
      if (warp == kTcEpiWarps - 1 && lane == 0) TC_TRACE(10);
      U = Un; w = wn;
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == kTcEpiWarps + 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc<kTcTmemCols>(tmem_base);
  }
  // early start: completion of this kernel must still imply completion of its stream predecessor
  if (tc.early) pdl_wait();
  if (tid == 0) {
    ktrace_end(batch.ktrace, 1);
    if (tr) {
      tr[11] = clock64();
      tr[15] = tc_globaltimer();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Refine.  Two lanes per query: the even lane rescans the query's best bucket, the odd lane the
// second-best one (after merging the per-split candidates), with exact Hamming distances on the
// integer pipe and packed (distance, trainIdx) keys, lowest train index winning ties.  The
// candidate rows of a warp's 32 (query, bucket) pairs are staged through shared memory with
// coalesced 16-byte loads (a bucket is kTcBucket * 32 contiguous bytes), 8 rows per pair per
// pass, so each lane then reads its own rows without bank conflicts.  The pair is merged with
// one shuffle and the even lane applies the ratio test (src/slam_frontend.cc:529-536).
constexpr int kRefineQB = 64;                                // queries per refine CTA
constexpr int kCompactQB = 128;                              // queries per compaction CTA
constexpr int kRefineThreads = 2 * kRefineQB;
constexpr int kRefineRows = 8;                               // rows per pair per pass

// EXACT2 = true: two lanes per query as described (vsf_knn2 reports the second neighbour's index).
// EXACT2 = false (every other entry point: only the second neighbour's DISTANCE matters, for the
// ratio test): one lane per query rescans the best bucket and takes the second-best bucket's
// best distance from its exact maximum dot, dot = 256 - 2 * hamming - half the rows to fetch,
// half the lanes, 128 queries per CTA.
// ROWS = rows per pair per staging pass (8 for the stand-alone kernel; 4 for the chained post
// kernel, whose shared memory has to fit beside a resident CTA of knn2_tc_kernel).
// PDL: the stand-alone kernel waits for its stream predecessor between the query fetch and the
// partial keys.  All kRefineThreads threads of the CTA call this with the same (problem, qb).
template <int ROWS>
struct RefineGeom {
  static constexpr int kPitch = ROWS * 32 + 16;   // bytes per pair in the stage (+16: bank skew)
  static constexpr int kPP = ROWS * 2;            // 16-byte pieces of one pair's run of rows
  static constexpr int kStageBytes = (kRefineThreads / 32) * 32 * kPitch;
  static_assert(kTcBucket % ROWS == 0 && 32 % kPP == 0, "bucket must be a multiple of the refine pass");
};

template <bool EXACT2, int ROWS, bool PDL>
__device__ __forceinline__ void refine_block(const KnnBatch& batch, const TcBatch& tc, int problem, int qb,
                                             uint8_t* s_stage_all) {
  constexpr int kQB = EXACT2 ? kRefineQB : 2 * kRefineQB;      // queries per CTA
  constexpr int kPitch = RefineGeom<ROWS>::kPitch;
  constexpr int kPP = RefineGeom<ROWS>::kPP;
  const KnnProblem& P = batch.p[problem];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  // The descriptors and row counts were written before the launch sequence began (see
  // knn2_tc_kernel): the query words are fetched while the distance kernel is still finishing.
  int nq = P.nq, nt = P.nt;
  if (P.nq_dev) nq = min(nq, *P.nq_dev);
  if (P.nt_dev) nt = min(nt, *P.nt_dev);
  const int q0 = qb * kQB;
  const int q = q0 + (EXACT2 ? (tid >> 1) : tid);
  const int c = EXACT2 ? (tid & 1) : 0;                       // 0: best bucket, 1: second-best bucket
  uint32_t qw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
  if (q < nq) {
    const uint4* src = reinterpret_cast<const uint4*>(P.q + size_t(q) * 8);
    const uint4 w0 = __ldg(src), w1 = __ldg(src + 1);
    qw[0] = w0.x; qw[1] = w0.y; qw[2] = w0.z; qw[3] = w0.w;
    qw[4] = w1.x; qw[5] = w1.y; qw[6] = w1.z; qw[7] = w1.w;
  }
  if (PDL) {
    pdl_wait();                                  // the partial bucket keys are complete
    if (tid == 0) ktrace_start(batch.ktrace, 4);
    pdl_launch_dependents();
  }
  if (nq <= 0) {
    if (qb == 0 && tid == 0) {
      *P.match_count = 0;
      if (batch.host_counts) batch.host_counts[P.region] = 0;
    }
    return;
  }
  if (q0 >= nq) return;

  int key = kTcKeySentinel, key2 = kTcKeySentinel;
  if (q < nq) {
    int b1 = kTcKeySentinel, b2 = kTcKeySentinel;
    // the query block's tile slots were shared out to consecutive CTAs, one partial segment each
    const int nseg = tc_block_segments(tc, tc.qb_begin[problem] + q / kTcQ);
    // a segment's two column-half partials are one aligned 16-byte record; four records are
    // fetched together so that a block cut into many segments costs one L2 round trip per four
    // segments, not one per partial
    static_assert(kTcColSplit == 2, "one uint4 = the two column-half partials of a segment");
    const uint4* part = reinterpret_cast<const uint4*>(reinterpret_cast<const uint2*>(batch.partial) +
                                                        size_t(P.row0 + q) * (tc.slots * kTcColSplit));
    auto merge = [&](int a1, int a2) {   // merge two descending pairs
      const int hi = max(b1, a1), lo = min(b1, a1);
      b2 = max(lo, max(b2, a2));
      b1 = hi;
    };
    for (int z0 = 0; z0 < nseg; z0 += 4) {
      uint4 p[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        p[j] = (z0 + j < nseg) ? __ldcg(part + z0 + j)
                               : make_uint4(uint32_t(kTcKeySentinel), uint32_t(kTcKeySentinel),
                                            uint32_t(kTcKeySentinel), uint32_t(kTcKeySentinel));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        merge(int(p[j].x), int(p[j].y));
        merge(int(p[j].z), int(p[j].w));
      }
    }
    key = c ? b2 : b1;
    key2 = b2;
  }
  // first train row of this lane's candidate bucket, -1 = none
  const int my_row0 = (key == kTcKeySentinel) ? -1 : (kBucketIdMask - (key & kBucketIdMask)) * kTcBucket;
  uint32_t k1 = kKeySentinel, k2 = kKeySentinel;
  if (!EXACT2 && key2 != kTcKeySentinel) {
    // the second-best bucket's best distance is exact (its maximum dot is); its first row stands
    // in as the index, which nobody reads on this path
    const int dot = key2 >> kBucketIdBits;
    k1 = (uint32_t((kTcRowBytes - dot) >> 1) << kIdxBits) + uint32_t((kBucketIdMask - (key2 & kBucketIdMask)) * kTcBucket);
  }
  uint8_t* stage = s_stage_all + size_t(warp) * (32 * kPitch);
#pragma unroll 1
  for (int r0 = 0; r0 < kTcBucket; r0 += ROWS) {
    // stage: 32 pairs x ROWS rows x 32 B in pieces of 16 B, kPP per lane
#pragma unroll
    for (int i = 0; i < kPP; ++i) {
      const int g = i * (32 / kPP) + lane / kPP;          // pair (lane) whose rows this piece belongs to
      const int piece = lane % kPP;                       // 16-byte piece of the pair's run of rows
      const int base = __shfl_sync(0xffffffffu, my_row0, g);
      const int row = base + r0 + (piece >> 1);
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (base >= 0 && row < nt) v = __ldg(reinterpret_cast<const uint4*>(P.t + size_t(row) * 8) + (piece & 1));
      *reinterpret_cast<uint4*>(stage + g * kPitch + piece * 16) = v;
    }
    __syncwarp();
    if (my_row0 >= 0) {
      const uint4* mine = reinterpret_cast<const uint4*>(stage + lane * kPitch);
#pragma unroll
      for (int k = 0; k < ROWS; ++k) {
        const int row = my_row0 + r0 + k;
        const uint4 t0 = mine[2 * k], t1 = mine[2 * k + 1];
        const uint32_t tw[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
        const uint32_t d = hamming256<2>(qw, tw);
        if (row < nt) top2_insert(k1, k2, (d << kIdxBits) + uint32_t(row));
      }
    }
    __syncwarp();
  }
  // merge the pair (both lanes end up with the query's exact top-2)
  if (EXACT2) {
    const uint32_t o1 = __shfl_xor_sync(0xffffffffu, k1, 1), o2 = __shfl_xor_sync(0xffffffffu, k2, 1);
    top2_merge(k1, k2, o1, o2);
  }
  // unpack + ratio test; survivors are compacted by knn2_compact_kernel
  bool pass = false;
  if (q < nq && c == 0) {
    const int i0 = (k1 == kKeySentinel) ? -1 : int(k1 & kIdxMask);
    const int i1 = (k2 == kKeySentinel) ? -1 : int(k2 & kIdxMask);
    const int d0 = (k1 == kKeySentinel) ? -1 : int(k1 >> kIdxBits);
    const int d1 = (k2 == kKeySentinel) ? -1 : int(k2 >> kIdxBits);
    batch.knn_out[P.row0 + q] = make_uint4(uint32_t(i0), uint32_t(i1), uint32_t(d0), uint32_t(d1));
    pass = (i1 >= 0) && (double(d0) < batch.ratio * double(d1));
  }
  // survivor counts per block of kRefineQB (64) queries, which is what the compaction sums
  if (EXACT2) {
    const int npass = __syncthreads_count(pass);
    if (tid == 0) batch.qblock_pass[P.qb0 + qb] = unsigned(npass);
  } else {
    const int n0 = __syncthreads_count(pass && tid < kRefineQB);
    const int n1 = __syncthreads_count(pass && tid >= kRefineQB);
    if (tid == 0) {
      batch.qblock_pass[P.qb0 + 2 * qb] = unsigned(n0);
      batch.qblock_pass[P.qb0 + 2 * qb + 1] = unsigned(n1);
    }
  }
}

template <bool EXACT2>
__global__ void __launch_bounds__(kRefineThreads)
knn2_tc_refine_kernel(const __grid_constant__ KnnBatch batch, const __grid_constant__ TcBatch tc) {
  __shared__ __align__(16) uint8_t s_stage[RefineGeom<kRefineRows>::kStageBytes];
  if (threadIdx.x == 0) ktrace_start(batch.ktrace, 2);
  refine_block<EXACT2, kRefineRows, true>(batch, tc, blockIdx.y, blockIdx.x, s_stage);
  if (threadIdx.x == 0) ktrace_end(batch.ktrace, 2);
}

// Ordered compaction of the ratio survivors, one CTA per block of 128 queries: the CTA's
// output offset is the sum of the survivor counts of the query blocks before it, so the list
// comes out in ascending queryIdx order (what Frontend::GetMatches returns) with no serial pass.
// All QB threads of the CTA call this with the same (problem, qb); the survivor counts (one per
// CNT queries) and knn_out of the whole problem must be complete.
template <int QB, int CNT>
__device__ __forceinline__ void compact_block(const KnnBatch& batch, int problem, int qb, unsigned* s_red,
                                              unsigned* s_woff) {
  const KnnProblem& P = batch.p[problem];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  int nq = P.nq;
  if (P.nq_dev) nq = min(nq, *P.nq_dev);
  if (nq <= 0) return;                       // the refine wrote match_count = 0
  const int q0 = qb * QB;
  if (q0 >= nq) return;
  const int nqb = (nq + QB - 1) / QB;
  unsigned sum = 0;
  for (int i = tid; i < qb * (QB / CNT); i += QB) sum += __ldcg(&batch.qblock_pass[P.qb0 + i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s_red[warp] = sum;
  const int q = q0 + tid;
  bool ok = false;
  uint4 rec = make_uint4(0u, 0u, 0u, 0u);
  if (q < nq) {
    rec = __ldcg(&batch.knn_out[P.row0 + q]);
    ok = (int(rec.y) >= 0) && (double(int(rec.z)) < batch.ratio * double(int(rec.w)));
  }
  const unsigned bal = __ballot_sync(0xffffffffu, ok);
  if (lane == 0) s_woff[warp] = __popc(bal);
  __syncthreads();
  unsigned off = 0, total = 0;
#pragma unroll
  for (int w = 0; w < QB / 32; ++w) {
    off += s_red[w];
    if (w < warp) off += s_woff[w];
    total += s_woff[w];
  }
  if (ok) {
    int4 m;
    m.x = q;                                  // queryIdx
    m.y = int(rec.x);                         // trainIdx
    m.z = 0;                                  // imgIdx
    m.w = __float_as_int(float(int(rec.z)));  // distance
    const unsigned dst = off + __popc(bal & ((1u << lane) - 1u));
    reinterpret_cast<int4*>(P.matches)[dst] = m;
    if (batch.host_matches)
      reinterpret_cast<int4*>(batch.host_matches + size_t(P.region) * batch.host_region_stride)[dst] = m;
  }
  if (qb == nqb - 1 && tid == 0) {
    unsigned base = 0;
#pragma unroll
    for (int w = 0; w < QB / 32; ++w) base += s_red[w];
    *P.match_count = int(base + total);
    if (batch.host_counts) batch.host_counts[P.region] = int(base + total);
  }
}

__global__ void __launch_bounds__(kCompactQB)
knn2_compact_kernel(const __grid_constant__ KnnBatch batch) {
  __shared__ unsigned s_red[kCompactQB / 32];
  __shared__ unsigned s_woff[kCompactQB / 32];
  if (threadIdx.x == 0) ktrace_start(batch.ktrace, 3);
  pdl_wait();
  pdl_launch_dependents();
  compact_block<kCompactQB, kRefineQB>(batch, blockIdx.y, blockIdx.x, s_red, s_woff);
  if (threadIdx.x == 0) ktrace_end(batch.ktrace, 3);
}

// ---------------------------------------------------------------------------------------------
// Refine + ordered compaction in one kernel (every entry point that does not need the second
// neighbour's index).  A CTA owns a block of 256 consecutive queries of one problem:
//   1. one lane per query merges the query's partial bucket keys and gives the early verdict:
//      the best bucket's maximum dot is the exact best distance d0, the second-best bucket's the
//      exact smallest distance d2b outside the best bucket, and the second neighbour's distance
//      is <= d2b, so !(d0 < ratio * d2b) already fails the ratio test;
//   2. the other queries - the candidates - are re-packed onto consecutive groups of four lanes
//      (64 groups per pass): step by step the four lanes read 64 contiguous bytes (two rows) of
//      the candidate's best bucket, neighbouring lanes add up the two halves of a row's exact
//      Hamming distance, and one shuffle step merges the packed (distance, trainIdx) top-2 keys
//      of the even and the odd rows (lowest train index winning ties, as everywhere); whole
//      warps skip the pass when the candidates run out;
//   3. back in query order the survivors are counted and stored in ascending queryIdx order.
// The compaction needs the survivor counts of the query blocks before this one: each CTA
// publishes its count as (epoch << 16 | count) and reads its predecessors' ("decoupled
// look-back": the blocks of a problem are refined side by side, so the wait is short).  A CTA
// only ever waits for CTAs with a smaller logical index, and the logical index is handed out by
// an atomic ticket in the order the CTAs start, so every CTA that is waited for is already
// running: no deadlock whatever order the hardware dispatches the grid in.  The look-back words
// carry the launch's epoch in their upper bits (64-bit words, the epoch only grows) and are
// never reset; the epoch and the ticket counter live in device memory (FinishArgs::state) and
// are advanced / reset by the last CTA of a launch to finish.
constexpr int kFinLanes = 4;                          // lanes per candidate in the rescan
#ifndef VSF_FIN_THREADS
#define VSF_FIN_THREADS 256
#endif
constexpr int kFinThreads = VSF_FIN_THREADS;
constexpr int kFinGroups = kFinThreads / kFinLanes;   // candidates per rescan pass
constexpr int kFinCountBits = 16;
static_assert(kFinLanes == 4 && kTcBucket % 8 == 0, "knn2_tc_finish_kernel: 4 lanes x 4 steps of 2 rows");
// Queries per CTA (QPB, = one survivor counter): 256 keeps the most queries in flight per SM and
// is what a batch that fills the machine wants (a group of poses); a small batch is a chain of
// dependent L2 round trips on a few CTAs, which 64 queries per CTA - one rescan pass, four times
// the CTAs - keeps short.  The first QPB lanes of the CTA do step 1.

// The look-back words carry their payload themselves (epoch | count): relaxed accesses at GPU
// scope are enough, and an acquire load in the polling loop would invalidate the SM's L1 on
// every iteration (CCTL.IVALL).
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// WORDS = 8: 32-byte rows (knn2_tc_kernel's partial keys: one 16-byte record per segment), 16:
// 64-byte rows (knn2_tc64_kernel.cu: two records per segment; a step reads one row, 16 bytes
// per lane, and the four lanes add up their quarters).
template <int WORDS, int QPB>
__global__ void __launch_bounds__(kFinThreads, 1536 / kFinThreads)
knn2_tc_finish_kernel(const __grid_constant__ KnnBatch batch, const __grid_constant__ TcBatch tc,
                      const __grid_constant__ FinishArgs fa) {
  static_assert(WORDS == 8 || WORDS == 16, "32- or 64-byte rows");
  static_assert(QPB % 32 == 0 && QPB <= kFinThreads && QPB <= 256 && QPB < (1 << kFinCountBits),
                "FinishArgs::flags is indexed in 32-query units (KnnProblem::qb0); candidate indices are bytes");
  constexpr int kFinQ = QPB;
  constexpr int kRecs = WORDS / 8;             // 16-byte partial records per segment
  __shared__ unsigned s_woff[kFinThreads / 32];
  __shared__ unsigned s_ccnt[kFinThreads / 32];   // candidates per warp
  __shared__ int2 s_keys[kFinQ];                  // (best, second-best) bucket keys of the block's candidates
  __shared__ int2 s_out[kFinQ];                   // (trainIdx or -1, distance) of the block's candidates
  __shared__ unsigned char s_list[kFinQ];         // the candidates' indices in the block, ascending
  __shared__ unsigned s_base;
  __shared__ int s_where[4];
  __shared__ unsigned long long s_epoch;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int part = tid % kFinLanes;
  if (tid == 0) ktrace_start(batch.ktrace, 2);
  // one block per CTA, taken by ticket (the do-while only gives the early exits a common end)
  do {
    if (tid == 0) {
      // the integer divisions of the block's coordinates are done once per CTA
      s_epoch = ld_relaxed_u64(fa.state + 2);   // stable while any CTA of this launch runs
      const int id = int(atomicAdd(fa.state, 1ull));
      const int pr = id / fa.nqb, b = id - pr * fa.nqb;
      s_where[0] = pr;
      s_where[1] = b;
      // segments per query of the distance kernel's query block(s) this block lies in (64-byte
      // rows, one CTA per unit: two 128-query blocks)
      const int uq = WORDS == 8 ? kTcQ : tc.unit_q;
      const int gqb = tc.qb_begin[pr] + (b * kFinQ) / uq;
      s_where[2] = tc_block_segments(tc, gqb);
      s_where[3] = (uq < kFinQ && gqb + 1 < tc.qb_begin[pr + 1]) ? tc_block_segments(tc, gqb + 1) : s_where[2];
    }
    __syncthreads();
    const int problem = s_where[0], qb = s_where[1];
    const unsigned long long epoch = s_epoch;
    const KnnProblem& P = batch.p[problem];
    // (the descriptors and row counts were written before the launch sequence began, see
    // knn2_tc_kernel)
    int nq = P.nq, nt = P.nt;
    if (P.nq_dev) nq = min(nq, *P.nq_dev);
    if (P.nt_dev) nt = min(nt, *P.nt_dev);
    const int q0 = qb * kFinQ;
    const int q = q0 + tid;
    pdl_wait();                                // the partial bucket keys are complete
    if (tid == 0) ktrace_start(batch.ktrace, 4);
    pdl_launch_dependents();
    if (nq <= 0) {                             // CTA-uniform, like the next test
      if (qb == 0 && tid == 0) {
        *P.match_count = 0;
        if (P.host_count) *P.host_count = 0;
        else if (batch.host_counts) batch.host_counts[P.region] = 0;
      }
      break;
    }
    if (q0 >= nq) break;
    const int nqb = (nq + kFinQ - 1) / kFinQ;

    // ---- 1. merge the query's partial bucket keys, early verdict (one lane per query)
    bool cand = false;
    if (tid < kFinQ && q < nq) {
      int b1 = kTcKeySentinel, b2 = kTcKeySentinel;
      // the query block's tile slots were shared out to consecutive CTAs, one partial segment each
      const int nseg = ((WORDS == 8 || tc.unit_q >= kFinQ || tid < kFinQ / 2) ? s_where[2] : s_where[3]) * kRecs;
      static_assert(kTcColSplit == 2 && kTcQ % kFinQ == 0 && (2 * 128) % kFinQ == 0, "one uint4 = two partial pairs of a segment");
      const uint4* part_keys = reinterpret_cast<const uint4*>(reinterpret_cast<const uint2*>(batch.partial) +
                                                               size_t(P.row0 + q) * (tc.slots * kTcColSplit * kRecs));
      auto merge = [&](int a1, int a2) {   // merge two descending pairs
        const int hi = max(b1, a1), lo = min(b1, a1);
        b2 = max(lo, max(b2, a2));
        b1 = hi;
      };
      for (int z0 = 0; z0 < nseg; z0 += 2) {
        uint4 p[2];
#pragma unroll
        for (int j = 0; j < 2; ++j)
          p[j] = (z0 + j < nseg) ? __ldcg(part_keys + z0 + j)
                                 : make_uint4(uint32_t(kTcKeySentinel), uint32_t(kTcKeySentinel),
                                              uint32_t(kTcKeySentinel), uint32_t(kTcKeySentinel));
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          merge(int(p[j].x), int(p[j].y));
          merge(int(p[j].z), int(p[j].w));
        }
      }
      if (b1 != kTcKeySentinel) {
        cand = true;
        if (b2 != kTcKeySentinel && batch.ratio >= 0.0) {   // (ratio >= 0: the double product is monotone in the distance)
          const int d0b = (32 * WORDS - (b1 >> kBucketIdBits)) >> 1;
          const int d2b = (32 * WORDS - (b2 >> kBucketIdBits)) >> 1;
          cand = double(d0b) < batch.ratio * double(d2b);
        }
        if (cand) s_keys[tid] = make_int2(b1, b2);
      }
    }
    const unsigned cbal = __ballot_sync(0xffffffffu, cand);
    if (lane == 0) s_ccnt[warp] = __popc(cbal);
    __syncthreads();
    unsigned cbase = 0, ncand = 0;
#pragma unroll
    for (int w = 0; w < kFinThreads / 32; ++w) {
      if (w < warp) cbase += s_ccnt[w];
      ncand += s_ccnt[w];
    }
    if (cand) s_list[cbase + __popc(cbal & ((1u << lane) - 1u))] = static_cast<unsigned char>(tid);
    __syncthreads();
    // ---- 2. exact distances to the rows of the candidates' best buckets, kFinGroups candidates
    // per pass.  The bucket is 2 * kTcBucket 16-byte pieces; in step j the four lanes of a group
    // read pieces 4j .. 4j+3 (64 contiguous bytes = rows 2j, 2j+1), a lane and its neighbour add up
    // the two halves of a row.  (64-byte rows: 4 * kTcBucket pieces, a step is one row, the four
    // lanes add up its quarters.)
    for (unsigned c0 = 0; c0 + unsigned(warp * (32 / kFinLanes)) < ncand; c0 += kFinGroups) {   // warp-uniform
      const unsigned ci = c0 + unsigned(tid / kFinLanes);
      const bool have = ci < ncand;
      const int cq = have ? int(s_list[ci]) : 0;
      int ckey = kTcKeySentinel, ckey2 = kTcKeySentinel;
      // lane `part` works on 16-byte half `part & 1` of every second row of the bucket (64-byte
      // rows: on quarter `part` of every row)
      uint4 qh = make_uint4(0u, 0u, 0u, 0u);
      if (have) {
        const int2 kk = s_keys[cq];
        ckey = kk.x;
        ckey2 = kk.y;
        qh = __ldg(reinterpret_cast<const uint4*>(P.q + size_t(q0 + cq) * WORDS) + (WORDS == 8 ? (part & 1) : part));
      }
      uint32_t k1 = kKeySentinel, k2 = kKeySentinel;
      {
        const int row_base = have ? (kBucketIdMask - (ckey & kBucketIdMask)) * kTcBucket : 0;
        const uint4* pieces = reinterpret_cast<const uint4*>(P.t + size_t(row_base) * WORDS) + part;
        const int rsub = WORDS == 8 ? (part >> 1) : 0;
        constexpr int kSteps = WORDS == 8 ? kTcBucket / 2 : kTcBucket;
        constexpr int kRowsPerStep = WORDS == 8 ? 2 : 1;
#pragma unroll
        for (int j0 = 0; j0 < kSteps; j0 += 4) {
          uint4 t[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            t[u] = (have && row_base + kRowsPerStep * (j0 + u) + rsub < nt) ? __ldg(pieces + 4 * (j0 + u)) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t d = __popc(t[u].x ^ qh.x) + __popc(t[u].y ^ qh.y) + __popc(t[u].z ^ qh.z) + __popc(t[u].w ^ qh.w);
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            if (WORDS == 16) d += __shfl_xor_sync(0xffffffffu, d, 2);
            const int row = row_base + kRowsPerStep * (j0 + u) + rsub;
            if (have && row < nt) top2_insert(k1, k2, (d << kIdxBits) + uint32_t(row));
          }
        }
      }
      if (WORDS == 8) {
        // the two lane pairs of a group saw the even / the odd rows
        const uint32_t o1 = __shfl_xor_sync(0xffffffffu, k1, 2), o2 = __shfl_xor_sync(0xffffffffu, k2, 2);
        top2_merge(k1, k2, o1, o2);
      }
      if (ckey2 != kTcKeySentinel) {
        // the second-best bucket's best distance is exact (its maximum dot is); its first row stands
        // in as the index, which nobody reads on this path
        const int dot = ckey2 >> kBucketIdBits;
        top2_insert(k1, k2, (uint32_t((32 * WORDS - dot) >> 1) << kIdxBits) +
                                uint32_t((kBucketIdMask - (ckey2 & kBucketIdMask)) * kTcBucket));
      }
      // the ratio test (src/slam_frontend.cc:529-536), one lane per candidate
      if (have && part == 0) {
        const int ci0 = (k1 == kKeySentinel) ? -1 : int(k1 & kIdxMask);
        const int cd0 = (k1 == kKeySentinel) ? -1 : int(k1 >> kIdxBits);
        const int i1 = (k2 == kKeySentinel) ? -1 : int(k2 & kIdxMask);
        const int d1 = (k2 == kKeySentinel) ? -1 : int(k2 >> kIdxBits);
        const bool ok = (i1 >= 0) && (double(cd0) < batch.ratio * double(d1));
        s_out[cq] = make_int2(ok ? ci0 : -1, cd0);
      }
    }
    __syncthreads();
    // ---- 3. back to one lane per query of the block, in query order
    bool pass = false;
    int i0 = -1, d0 = -1;
    if (cand) {
      const int2 r = s_out[tid];
      pass = r.x >= 0;
      i0 = r.x;
      d0 = r.y;
    }
    // ---- publish this block's survivor count, fetch the ones before it
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if (lane == 0) s_woff[warp] = __popc(bal);
    __syncthreads();
    unsigned total = 0, woff = 0;
#pragma unroll
    for (int w = 0; w < kFinThreads / 32; ++w) {
      if (w < warp) woff += s_woff[w];
      total += s_woff[w];
    }
    unsigned long long* flags = fa.flags + P.qb0;              // qb0 counts 32-query units: every (kFinQ / 32)-th word is used
    if (tid == 0) st_relaxed_u64(flags + qb * (kFinQ / 32), (epoch << kFinCountBits) | total);
    if (warp == 0) {
      // one warp polls (with a back-off: the other CTAs of the SM are still refining)
      unsigned sum = 0;
      for (int i = lane; i < qb; i += 32) {
        unsigned long long v = ld_relaxed_u64(flags + i * (kFinQ / 32));
        while ((v >> kFinCountBits) != epoch) {
          __nanosleep(40);
          v = ld_relaxed_u64(flags + i * (kFinQ / 32));
        }
        sum += unsigned(v & ((1ull << kFinCountBits) - 1ull));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) s_base = sum;
    }
    __syncthreads();
    const unsigned base = s_base;
    // ---- ordered store (ascending queryIdx, what Frontend::GetMatches returns)
    if (pass) {
      int4 m;
      m.x = q;                                  // queryIdx
      m.y = i0;                                 // trainIdx
      m.z = 0;                                  // imgIdx
      m.w = __float_as_int(float(d0));          // distance
      const unsigned dst = base + woff + __popc(bal & ((1u << lane) - 1u));
      reinterpret_cast<int4*>(P.matches)[dst] = m;
      if (batch.host_matches)
        reinterpret_cast<int4*>(batch.host_matches + size_t(P.region) * batch.host_region_stride)[dst] = m;
    }
    if (qb == nqb - 1 && tid == 0) {
      *P.match_count = int(base + total);
      if (P.host_count) *P.host_count = int(base + total);
      else if (batch.host_counts) batch.host_counts[P.region] = int(base + total);
    }
  } while (false);
  // the last CTA to get here leaves the state ready for the next launch on it
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(fa.state + 1, 1ull) == (unsigned long long)(gridDim.x - 1)) {
      const unsigned long long e = ld_relaxed_u64(fa.state + 2);
      st_relaxed_u64(fa.state, 0ull);
      st_relaxed_u64(fa.state + 1, 0ull);
      __threadfence();
      st_relaxed_u64(fa.state + 2, e + 1ull);
    }
    ktrace_end(batch.ktrace, 2);
  }
}

// ---------------------------------------------------------------------------------------------
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

cudaError_t launch_expand_train_multi(const ExpandMulti& em, int int8, int pdl, cudaStream_t stream) {
  if (em.frames <= 0 || em.nt <= 0) return cudaSuccess;
  const int rows_pad = (em.nt + kTcTileRows - 1) / kTcTileRows * kTcTileRows;
  constexpr int kExpandThreads = 128;
  const dim3 grid((rows_pad * 4 + kExpandThreads - 1) / kExpandThreads, em.frames);
  return int8 ? launch_pdl(expand_train_multi_kernel<true>, grid, dim3(kExpandThreads), 0, stream, pdl != 0, em)
              : launch_pdl(expand_train_multi_kernel<false>, grid, dim3(kExpandThreads), 0, stream, pdl != 0, em);
}

cudaError_t launch_expand_train(const void* t, int nt_bound, const int* nt_dev, void* out, int int8,
                                int pdl, cudaStream_t stream, long long* ktrace) {
  if (nt_bound <= 0) return cudaSuccess;
  const int rows_pad = (nt_bound + kTcTileRows - 1) / kTcTileRows * kTcTileRows;
  // 128-thread CTAs: one warp per SM sub-partition at 18 registers per thread fits beside a
  // resident CTA of knn2_tc_kernel (whose 5 warps x 96 registers leave 1024 registers free on
  // two of the four sub-partitions), so the distance kernel of this launch can start its
  // prologue (and expand its first queries) while the expansion is still running
  constexpr int kExpandThreads = 128;
  {
    // ... and the same shared-memory carve-out as that kernel (the largest one): a kernel that
    // prefers another L1 / shared split cannot share an SM with it, the SM is reconfigured only
    // once it has drained
    // (function attributes are per device; one bit per device ordinal, set once, safe when
    // several contexts on several devices launch concurrently)
    static std::atomic<unsigned long long> carveout_done{0ull};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(carveout_done.load(std::memory_order_acquire) & bit)) {
      cudaError_t e = cudaFuncSetAttribute(expand_train_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(expand_train_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) return e;
      carveout_done.fetch_or(bit, std::memory_order_release);
    }
  }
  const int threads = rows_pad * 4;
  const int blocks = (threads + kExpandThreads - 1) / kExpandThreads;
  const uint32_t* tp = static_cast<const uint32_t*>(t);
  uint8_t* op = static_cast<uint8_t*>(out);
  return int8 ? launch_pdl(expand_train_kernel<true>, dim3(blocks), dim3(kExpandThreads), 0, stream, pdl != 0, tp, nt_bound, nt_dev, op, ktrace)
              : launch_pdl(expand_train_kernel<false>, dim3(blocks), dim3(kExpandThreads), 0, stream, pdl != 0, tp, nt_bound, nt_dev, op, ktrace);
}

// the ordered compaction alone (shared with the 64-byte engine, knn2_tc64_kernel.cu)
cudaError_t launch_knn2_compact(const KnnBatch& batch, int max_nq, bool pdl, cudaStream_t stream) {
  dim3 cgrid((max_nq + kCompactQB - 1) / kCompactQB, batch.num_problems);
  return launch_pdl(knn2_compact_kernel, cgrid, dim3(kCompactQB), 0, stream, pdl, batch);
}

// Queries per CTA of the finish kernel (see kFinThreads): by the number of queries of the batch
// (host-side bounds), VSF_FIN_QPB overrides (64, 128, 256; tuning).
static int finish_block_queries(const KnnBatch& batch) {
  static const int forced = [] {
    const char* e = std::getenv("VSF_FIN_QPB");
    const int v = e ? std::atoi(e) : 0;
    return (v == 64 || v == 128 || v == 256) ? v : 0;
  }();
  if (forced) return forced;
  long long total = 0;
  for (int i = 0; i < batch.num_problems; ++i) total += batch.p[i].nq;
  return total >= 150000 ? 256 : total >= 75000 ? 128 : 64;
}

// ev (optional, 4 events): recorded before the main kernel, after it, after the refine and
// after the compaction kernel (per-kernel timing for bench.py's roofline; an event between two
// kernels removes their programmatic overlap, so it is only used in dedicated timing passes).
// fa (optional): refine + compaction as ONE kernel (knn2_tc_finish_kernel; fa->nqb and the
// grid are filled in here); ignored when the second neighbour's index is wanted.
// The finish kernel alone, for the 64-byte engine (knn2_tc64_kernel.cu).
cudaError_t launch_knn2_tc64_finish(const KnnBatch& batch, const TcBatch& tc, int max_nq, bool pdl,
                                    cudaStream_t stream, FinishArgs* fa) {
  const int qpb = finish_block_queries(batch);
  fa->nqb = (max_nq + qpb - 1) / qpb;
  const int grid = fa->nqb * batch.num_problems;
  return qpb == 256 ? launch_pdl(knn2_tc_finish_kernel<16, 256>, dim3(grid), dim3(kFinThreads), 0, stream, pdl, batch, tc, *fa)
       : qpb == 128 ? launch_pdl(knn2_tc_finish_kernel<16, 128>, dim3(grid), dim3(kFinThreads), 0, stream, pdl, batch, tc, *fa)
                    : launch_pdl(knn2_tc_finish_kernel<16, 64>, dim3(grid), dim3(kFinThreads), 0, stream, pdl, batch, tc, *fa);
}

// phase: 0 = the whole sequence, 1 = the distance kernel only, 2 = the finish kernel only (a
// group of poses: first every pose's distance kernel, then every pose's finish kernel).
cudaError_t launch_knn2_tc(const KnnBatch& batch, const TcBatch& tc, int int8, int max_nq, int pdl,
                           cudaEvent_t* ev, cudaStream_t stream, FinishArgs* fa, int* launched, int phase) {
  if (launched) *launched = 0;
  if (batch.num_problems <= 0 || tc.total <= 0) return cudaSuccess;
  cudaError_t e;
  {
    // per-device function attribute, set once per device and operand kind (one bit each)
    static std::atomic<unsigned long long> smem_done[2] = {{0ull}, {0ull}};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    std::atomic<unsigned long long>& done = smem_done[int8 ? 1 : 0];
    if (!(done.load(std::memory_order_acquire) & bit)) {
      e = int8 ? cudaFuncSetAttribute(knn2_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes)
               : cudaFuncSetAttribute(knn2_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
      if (e != cudaSuccess) return e;
      done.fetch_or(bit, std::memory_order_release);
    }
  }
  const bool p = pdl != 0;
  if (ev) cudaEventRecord(ev[0], stream);
  if (phase != 2) {
    e = int8 ? launch_pdl(knn2_tc_kernel<true>, dim3(tc.grid), dim3(kTcThreads), kTcSmemBytes, stream, p, batch, tc)
             : launch_pdl(knn2_tc_kernel<false>, dim3(tc.grid), dim3(kTcThreads), kTcSmemBytes, stream, p, batch, tc);
    if (e != cudaSuccess) return e;
    if (launched) *launched = 1;
  }
  if (ev) cudaEventRecord(ev[1], stream);
  if (phase == 1) return cudaSuccess;
  if (fa && !batch.exact_second) {
    const int qpb = finish_block_queries(batch);
    fa->nqb = (max_nq + qpb - 1) / qpb;
    const int grid = fa->nqb * batch.num_problems;
    e = qpb == 256 ? launch_pdl(knn2_tc_finish_kernel<8, 256>, dim3(grid), dim3(kFinThreads), 0, stream, p, batch, tc, *fa)
      : qpb == 128 ? launch_pdl(knn2_tc_finish_kernel<8, 128>, dim3(grid), dim3(kFinThreads), 0, stream, p, batch, tc, *fa)
                   : launch_pdl(knn2_tc_finish_kernel<8, 64>, dim3(grid), dim3(kFinThreads), 0, stream, p, batch, tc, *fa);
    if (ev) {
      cudaEventRecord(ev[2], stream);
      cudaEventRecord(ev[3], stream);
    }
    if (launched) *launched = phase == 2 ? 1 : 2;
    return e;
  }
  if (phase == 2) return cudaErrorInvalidValue;
  dim3 cgrid((max_nq + kCompactQB - 1) / kCompactQB, batch.num_problems);
  if (batch.exact_second) {
    dim3 rgrid((max_nq + kRefineQB - 1) / kRefineQB, batch.num_problems);
    e = launch_pdl(knn2_tc_refine_kernel<true>, rgrid, dim3(kRefineThreads), 0, stream, p, batch, tc);
  } else {
    dim3 rgrid((max_nq + 2 * kRefineQB - 1) / (2 * kRefineQB), batch.num_problems);
    e = launch_pdl(knn2_tc_refine_kernel<false>, rgrid, dim3(kRefineThreads), 0, stream, p, batch, tc);
  }
  if (e != cudaSuccess) return e;
  if (ev) cudaEventRecord(ev[2], stream);
  e = launch_pdl(knn2_compact_kernel, cgrid, dim3(kCompactQB), 0, stream, p, batch);
  if (ev) cudaEventRecord(ev[3], stream);
  if (launched) *launched = 3;
  return e;
}

}  // namespace vsf
