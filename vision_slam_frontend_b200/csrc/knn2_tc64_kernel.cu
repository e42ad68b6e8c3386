// Tensor-core engine for 64-byte rows (AKAZE's 61 bytes zero-padded, BRISK / FREAK's 64): the
// reference's default extractor is AKAZE (src/slam_frontend.cc:553), so this is the width a
// drop-in sees by default.  Same idea and same selection scheme as knn2_tc_kernel.cu (read that
// file first): descriptor bits become +-1 int8 values, dot = 512 - 2 * hamming, the epilogue
// keeps the best two 16-row buckets per query, the refine kernel recomputes exact distances
// inside them.  What changes is the geometry, because a row expands to 512 bytes:
//   work unit  128 queries (one M = 128 UMMA tile; the A image is 128 x 512 B = 64 KB)
//   train tile 128 rows    (N = 128; 64 KB per stage, two stages: the shared memory of the
//              32-byte kernel exactly), 16 MMAs of K = 32 per tile
//   TMEM       two accumulators of 128 columns used by alternate tiles of a unit; the 16
//              epilogue warps are (tile parity) x (lane quarter) x (column half of 64)
// A unit's queries may only be replaced once every MMA of the unit has completed; the epilogue
// warps of the two parities meet on a named barrier for that.
#include <atomic>
#include <utility>

#include "tc_ptx.cuh"
#include "vsf_device.cuh"

namespace vsf {

cudaError_t launch_knn2_compact(const KnnBatch& batch, int max_nq, bool pdl, cudaStream_t stream);
cudaError_t launch_knn2_tc64_finish(const KnnBatch& batch, const TcBatch& tc, int max_nq, bool pdl,
                                    cudaStream_t stream, FinishArgs* fa);

namespace w64 {

constexpr int kQ = 128;                  // queries per work unit
constexpr int kT = 128;                  // train rows per tile
constexpr int kRowBytes = 512;           // one 512-bit descriptor as +-1 bytes
constexpr int kChunks = kRowBytes / 16;  // 16-byte K-chunks per row
constexpr int kABytes = kQ * kRowBytes;
constexpr int kBBytes = kT * kRowBytes;
constexpr int kStages = 2;
constexpr int kEpiWarps = 16;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kParts = 4;                // partial key pairs per (query, segment): parity x column half
constexpr int kEpiCols = kT / 2;         // 64 columns per epilogue warp
constexpr int kBarriers = 2 * kStages + 5;
constexpr int kSmemBytes = kABytes + kStages * kBBytes + kBarriers * 8 + 16 + 128;
constexpr int kTmemCols = 512;
constexpr int kBucketIdBits = 20;
constexpr int kBucketIdMask = (1 << kBucketIdBits) - 1;
constexpr int kKeyNone = int(0x80000000u);

// bit i of the nibble -> byte i: 0 -> +1, 1 -> -1
__device__ __forceinline__ uint32_t expand4(uint32_t nibble) {
  const uint32_t sign = (nibble * 0x10204080u) & 0x80808080u;
  return (sign | (sign - (sign >> 7))) | 0x01010101u;
}
__device__ __forceinline__ uint4 expand16(uint32_t b16) {
  return make_uint4(expand4(b16 & 0xFu), expand4((b16 >> 4) & 0xFu), expand4((b16 >> 8) & 0xFu),
                    expand4((b16 >> 12) & 0xFu));
}

// Train expansion: one thread = one (row, group of 4 K-chunks); tile image [32 chunks][128 rows][16 B].
__device__ __forceinline__ void expand_rows64(const uint32_t* __restrict__ t, int nt, uint8_t* __restrict__ out, int idx) {
  const int rows_pad = (nt + kT - 1) / kT * kT;
  const int row = idx % rows_pad;
  const int cg = idx / rows_pad;     // 0..7: K-chunks 4*cg .. 4*cg+3 (words 2*cg, 2*cg+1)
  if (cg >= 8) return;
  uint2 w = make_uint2(0u, 0u);
  const bool live = row < nt;
  if (live) w = __ldg(reinterpret_cast<const uint2*>(t + size_t(row) * 16) + cg);
  const int tile = row / kT, r = row % kT;
  uint8_t* base = out + size_t(tile) * kBBytes + size_t(r) * 16;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t word = (c < 2) ? w.x : w.y;
    const uint32_t b16 = (word >> (16 * (c & 1))) & 0xFFFFu;
    const uint4 v = live ? expand16(b16) : make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(base + size_t(4 * cg + c) * (kT * 16)) = v;
  }
}

__global__ void __launch_bounds__(128)
expand_train64_kernel(const uint32_t* __restrict__ t, int nt_bound, const int* __restrict__ nt_dev,
                      uint8_t* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  // (rows_pad comes from the bound: rows between the device-side count and the bound are zeroed)
  const int rows_pad = (nt_bound + kT - 1) / kT * kT;
  int nt = nt_bound;
  if (nt_dev) nt = min(nt, *nt_dev);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = idx % rows_pad;
  const int cg = idx / rows_pad;     // 0..7: K-chunks 4*cg .. 4*cg+3 (words 2*cg, 2*cg+1)
  if (cg >= 8) return;
  uint2 w = make_uint2(0u, 0u);
  const bool live = row < nt;
  if (live) w = __ldg(reinterpret_cast<const uint2*>(t + size_t(row) * 16) + cg);
  const int tile = row / kT, r = row % kT;
  uint8_t* base = out + size_t(tile) * kBBytes + size_t(r) * 16;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t word = (c < 2) ? w.x : w.y;
    const uint32_t b16 = (word >> (16 * (c & 1))) & 0xFFFFu;
    const uint4 v = live ? expand16(b16) : make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(base + size_t(4 * cg + c) * (kT * 16)) = v;
  }
}

struct Unit {
  int problem, qb, slot;
  int nq, nt;
  int q0;
  int t_begin, t_end, ntiles;
  int len;
  bool skip;
};
struct Walk {
  long long x, end;
  int gqb, t0, slot, p;
};
__device__ __forceinline__ Walk walk_begin(const TcBatch& tc) {
  Walk w;
  w.p = 0;
  w.x = (long long)blockIdx.x * tc.total / tc.grid;
  w.end = (long long)(blockIdx.x + 1) * tc.total / tc.grid;
  const long long gqb = w.x / tc.pieces;
  w.gqb = int(gqb);
  w.t0 = int(w.x - gqb * tc.pieces);
  w.slot = int(blockIdx.x) - tc_owner(gqb * tc.pieces, tc.total, tc.grid);
  return w;
}
__device__ __forceinline__ Unit walk_unit(const KnnBatch& batch, const TcBatch& tc, const Walk& w) {
  Unit U;
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
  U.q0 = U.qb * kQ;
  U.skip = U.q0 >= nq;
  const int tile0 = min(tc.tiles, w.t0 * tc.tiles_per_piece);
  const int tile1 = min(tc.tiles, (w.t0 + U.len) * tc.tiles_per_piece);
  U.t_begin = min(nt, tile0 * kT);
  U.t_end = min(nt, tile1 * kT);
  U.ntiles = (U.t_end - U.t_begin + kT - 1) / kT;
  return U;
}
__device__ __forceinline__ void walk_next(Walk& w, const Unit& U) {
  w.x += U.len;
  w.gqb += 1;
  w.t0 = 0;
  w.slot = 0;
  w.p = U.problem;
}
__device__ __forceinline__ bool walk_more(const Walk& w) { return w.x < w.end; }

__device__ __forceinline__ void epi_barrier() {   // the 16 epilogue warps only
  asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
}

// Start of a role's dependent work: wait for the stream predecessor, then let the successor
// launch - unless the launch is an early one (TcBatch::early: a stream of poses, everything the
// kernel reads was complete before its predecessor let it launch; it waits just before it exits).
__device__ __forceinline__ void role_wait(const TcBatch& tc) {
  if (!tc.early) {
    pdl_wait();
    if (!tc.late) pdl_launch_dependents();   // (TcBatch::late: the TMA producer does it after its last load)
  }
}

__global__ void __launch_bounds__(kThreads, 1)
knn2_tc64_kernel(const __grid_constant__ KnnBatch batch, const __grid_constant__ TcBatch tc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kABytes + kStages * kBBytes);
  uint64_t* full = bars;                  // [kStages] TMA -> MMA
  uint64_t* empty = bars + kStages;       // [kStages] MMA -> TMA
  uint64_t* tfull = bars + 2 * kStages;   // [2] MMA -> epilogue (per accumulator)
  uint64_t* tempty = tfull + 2;           // [2] epilogue -> MMA
  uint64_t* aready = tempty + 2;          // [1] the unit's queries are expanded in smem
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + kBarriers);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  if (tc.early && !tc.late) pdl_launch_dependents();

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], kEpiWarps / 2);
    }
    mbar_init(aready, kEpiWarps);
    mbar_fence_init();
  }
  if (warp == kEpiWarps + 1) tc::tmem_alloc<kTmemCols>(s_tmem);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  Walk wk = walk_begin(tc);
  const uint32_t lbo = uint32_t(kQ * 16);   // = kT * 16: K-chunk stride of both operand images
  const uint32_t sbo = 128u;

  if (warp == kEpiWarps) {
    // ------------------------------ TMA producer ------------------------------
    Unit U;
    if (lane == 0 && walk_more(wk)) U = walk_unit(batch, tc, wk);
    role_wait(tc);
    if (lane == 0) {
      uint32_t it = 0;
      while (walk_more(wk)) {
        if (!U.skip) {
          const uint8_t* src = tc.t_exp[U.problem] + size_t(U.t_begin / kT) * kBBytes;
          for (int k = 0; k < U.ntiles; ++k, ++it) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);
            mbar_arrive_expect_tx(&full[s], kBBytes);
            tma_load_1d(sB + size_t(s) * kBBytes, src + size_t(k) * kBBytes, kBBytes, &full[s]);
          }
        }
        walk_next(wk, U);
        if (walk_more(wk)) U = walk_unit(batch, tc, wk);
      }
    }
    __syncwarp();
    if (tc.late) pdl_launch_dependents();   // the last tile load is on its way
  } else if (warp == kEpiWarps + 1) {
    // ------------------------------ MMA issuer ------------------------------
    role_wait(tc);
    if (lane == 0) {
      uint32_t it = 0, unit_it = 0, acc_use[2] = {0u, 0u};
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
      while (walk_more(wk)) {
        const Unit U = walk_unit(batch, tc, wk);
        walk_next(wk, U);
        if (U.skip || U.ntiles == 0) continue;
        for (int k = 0; k < U.ntiles; ++k, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
          const int valid = min(kT, U.t_end - (U.t_begin + k * kT));
          const uint32_t idesc = tc::instr_desc(true, 128, (valid + 15) & ~15);
          const int a = k & 1;                       // accumulator of this tile
          mbar_wait(&full[s], ph);
          if (k == 0) mbar_wait(aready, unit_it & 1u);
          mbar_wait(&tempty[a], (acc_use[a] & 1u) ^ 1u);
          tc::fence_after_sync();
#pragma unroll
          for (int kk = 0; kk < kChunks / 2; ++kk) {   // one MMA consumes K = 32 bytes = 2 chunks
            const uint64_t ad = tc::smem_desc(a_addr + uint32_t(kk) * 2u * (kQ * 16), lbo, sbo);
            const uint64_t bd = tc::smem_desc(b_addr + s * kBBytes + uint32_t(kk) * 2u * (kT * 16), lbo, sbo);
            tc::mma_ss<true>(tmem_base + uint32_t(a) * 256u, ad, bd, idesc, kk > 0 ? 1u : 0u);
          }
          tc::commit(&tfull[a]);
          ++acc_use[a];
          tc::commit(&empty[s]);
        }
        ++unit_it;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ query expansion + epilogue ------------------------------
    // warp = (column half ch) * 8 + (tile parity h) * 4 + (TMEM lane quarter)
    const int h = (warp >> 2) & 1;
    const int ch = warp >> 3;
    const int r = (warp & 3) * 32 + lane;          // TMEM lane = query row of the unit
    const int g = h * 2 + ch;                      // which 16 bytes of the query row this thread expands
    const uint32_t taddr = tmem_base + (uint32_t((warp & 3) * 32) << 16) + uint32_t(h * 256 + ch * kEpiCols);
    constexpr int NB = kEpiCols / kTcBucket;       // buckets per warp per tile
    constexpr int TB = kT / kTcBucket;             // buckets per tile
    constexpr int RPB = kTcBucket / 2;             // packed registers per bucket
    uint32_t acc_use = 0;

    auto load_query = [&](const Unit& V) -> uint4 {
      uint4 w = make_uint4(0u, 0u, 0u, 0u);
      const int q = V.q0 + r;
      if (!V.skip && V.ntiles > 0 && q < V.nq)
        w = __ldg(reinterpret_cast<const uint4*>(batch.p[V.problem].q + size_t(q) * 16) + g);
      return w;
    };
    auto expand_query = [&](const uint4& qw) {
      const uint32_t words[4] = {qw.x, qw.y, qw.z, qw.w};
      uint8_t* dst = sA + size_t(r) * 16 + size_t(g * 8) * (kQ * 16);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t b16 = (words[c >> 1] >> (16 * (c & 1))) & 0xFFFFu;
        *reinterpret_cast<uint4*>(dst + size_t(c) * (kQ * 16)) = expand16(b16);
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(aready);
    };
    Unit U;
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    bool expanded = false;
    if (walk_more(wk)) {
      U = walk_unit(batch, tc, wk);
      w = load_query(U);
      if (!(U.skip || U.ntiles == 0)) {
        expand_query(w);
        expanded = true;
      }
    }
    role_wait(tc);
    while (walk_more(wk)) {
      walk_next(wk, U);
      const bool more = walk_more(wk);
      Unit Un;
      uint4 wn = make_uint4(0u, 0u, 0u, 0u);
      const KnnProblem& P = batch.p[U.problem];
      const int q = U.q0 + r;
      uint2* part = reinterpret_cast<uint2*>(batch.partial) +
                    (size_t(P.row0 + q) * (tc.slots * kParts) + U.slot * kParts + g);
      if (U.skip || U.ntiles == 0) {
        if (!U.skip && q < U.nq) *part = make_uint2(uint32_t(kKeyNone), uint32_t(kKeyNone));
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
      int m1 = kKeyNone, m2 = kKeyNone;
      const int bucket0 = U.t_begin / kTcBucket + ch * NB;
      for (int k = h; k < U.ntiles; k += 2, ++acc_use) {   // this warp's parity of the unit's tiles
        mbar_wait(&tfull[h], acc_use & 1u);
        tc::fence_after_sync();
        const int row0 = U.t_begin + k * kT + ch * kEpiCols;
        const int kbase = kBucketIdMask - (bucket0 + k * TB);
        // |dot| <= 512 fits int16: packed pairs, 64 columns in one load
        uint32_t v[32];
        tc::tmem_ld_32x32_pack16(taddr, v);
        tc::tmem_ld_wait();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[h]);
        const bool full_tile = row0 + kEpiCols <= U.t_end;   // warp-uniform
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          uint32_t* b = &v[j * RPB];
          if (!full_tile) {
            const int valid = U.t_end - (row0 + j * kTcBucket);
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
          const int bm = max(int(short(mm & 0xFFFFu)), int(mm) >> 16);
          const int key = bm * (1 << kBucketIdBits) + (kbase - j);
          m2 = max(m2, min(m1, key));
          m1 = max(m1, key);
        }
      }
      if (q < U.nq) *part = make_uint2(uint32_t(m1), uint32_t(m2));
      // every MMA of the unit has completed once both parities have seen their last tile: only
      // then may the next unit's queries replace this one's
      epi_barrier();
      U = Un; w = wn;
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc<kTmemCols>(tmem_base);
  }
  if (tc.early) pdl_wait();   // "complete" must still imply "the stream predecessor is complete"
}

// ---------------------------------------------------------------------------------------------
// The same engine on CTA PAIRS (cta_group::2): two SMs of one TPC share every MMA.
//   work unit  256 queries: CTA r of the pair holds queries [128 r, 128 r + 128) (its A image)
//   train tile 256 rows:    CTA r loads image tile 2j + r (its half of B, 64 KB per stage)
//   MMA        M = 256, N = 256, K = 32, issued by the pair's rank-0 CTA for both; each CTA gets
//              its 128 queries x 256 train rows in its own TMEM (two accumulators of 256 columns)
// so one 64 KB tile a CTA pulls from L2 now feeds 256 queries instead of 128 and every SM reads
// half of B from its partner's shared memory instead of its own: the two limits of the
// single-CTA kernel (L2 -> SM tile traffic, shared-memory operand bandwidth) both halve.
// Barriers: TMA -> MMA per CTA (the partner's "tile landed" is relayed to rank 0 by the partner's
// otherwise idle MMA warp), MMA -> TMA / MMA -> epilogue by multicast commits, epilogue -> MMA and
// "queries expanded" by remote arrivals on rank 0's barriers.
namespace pair {

constexpr int kQP = 2 * kQ;              // queries per unit (pair)
constexpr int kTP = 2 * kT;              // train rows per pair tile
constexpr int kBarriersP = 3 * kStages + 5;
constexpr int kSmemBytesP = kABytes + kStages * kBBytes + kBarriersP * 8 + 16 + 128;

__device__ __forceinline__ Walk walk_begin_pair(const TcBatch& tc) {
  Walk w;
  const long long pair_id = blockIdx.x >> 1;
  w.p = 0;
  w.x = pair_id * tc.total / tc.grid;
  w.end = (pair_id + 1) * tc.total / tc.grid;
  const long long gqb = w.x / tc.pieces;
  w.gqb = int(gqb);
  w.t0 = int(w.x - gqb * tc.pieces);
  w.slot = int(pair_id) - tc_owner(gqb * tc.pieces, tc.total, tc.grid);
  return w;
}
__device__ __forceinline__ Unit walk_unit_pair(const KnnBatch& batch, const TcBatch& tc, const Walk& w) {
  Unit U;
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
  U.q0 = U.qb * kQP;
  U.skip = U.q0 >= nq;
  const int tile0 = min(tc.tiles, w.t0 * tc.tiles_per_piece);
  const int tile1 = min(tc.tiles, (w.t0 + U.len) * tc.tiles_per_piece);
  U.t_begin = min(nt, tile0 * kTP);
  U.t_end = min(nt, tile1 * kTP);
  U.ntiles = (U.t_end - U.t_begin + kTP - 1) / kTP;
  return U;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
knn2_tc64_pair_kernel(const __grid_constant__ KnnBatch batch, const __grid_constant__ TcBatch tc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kABytes + kStages * kBBytes);
  uint64_t* full = bars;                   // [kStages] this CTA's TMA -> (rank 0: MMA, rank 1: relay)
  uint64_t* empty = bars + kStages;        // [kStages] MMA -> TMA, both CTAs (multicast commit)
  uint64_t* pfull = bars + 2 * kStages;    // [kStages] rank 0 only: the partner's tile has landed
  uint64_t* tfull = bars + 3 * kStages;    // [2] MMA -> epilogue, both CTAs (multicast commit)
  uint64_t* tempty = tfull + 2;            // [2] rank 0 only: both CTAs' epilogues -> MMA
  uint64_t* aready = tempty + 2;           // [1] rank 0 only: both CTAs' queries are expanded
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + kBarriersP);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  if (tc.early && !tc.late) pdl_launch_dependents();
  const uint32_t rank = tc::cluster_ctarank();

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&pfull[s], 1);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], kEpiWarps);          // 8 local + 8 remote warps per accumulator
    }
    mbar_init(aready, 2 * kEpiWarps);
    mbar_fence_init();
  }
  if (warp == kEpiWarps + 1) tc::tmem_alloc_2cta<kTmemCols>(s_tmem);
  tc::fence_before_sync();
  tc::cluster_sync();
  tc::fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  Walk wk = walk_begin_pair(tc);
  const uint32_t lbo = uint32_t(kQ * 16);   // = kT * 16: K-chunk stride of both operand images
  const uint32_t sbo = 128u;

  if (warp == kEpiWarps) {
    // ------------------------------ TMA producer (each CTA: its half of B) ------------------------------
    Unit U;
    if (lane == 0 && walk_more(wk)) U = walk_unit_pair(batch, tc, wk);
    role_wait(tc);
    if (lane == 0) {
      uint32_t it = 0;
      while (walk_more(wk)) {
        if (!U.skip) {
          const uint8_t* src = tc.t_exp[U.problem] + (size_t(U.t_begin / kT) + rank) * kBBytes;
          for (int k = 0; k < U.ntiles; ++k, ++it) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);
            mbar_arrive_expect_tx(&full[s], kBBytes);
            tma_load_1d(sB + size_t(s) * kBBytes, src + size_t(2 * k) * kBBytes, kBBytes, &full[s]);
          }
        }
        walk_next(wk, U);
        if (walk_more(wk)) U = walk_unit_pair(batch, tc, wk);
      }
    }
    __syncwarp();
    if (tc.late) pdl_launch_dependents();   // the last tile load is on its way
  } else if (warp == kEpiWarps + 1) {
    role_wait(tc);
    if (lane == 0 && rank == 0) {
      // ------------------------------ MMA issuer (rank 0, for the pair) ------------------------------
      uint32_t it = 0, unit_it = 0, acc_use[2] = {0u, 0u};
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
      const uint32_t idesc = tc::instr_desc(true, 256, 256);
      while (walk_more(wk)) {
        const Unit U = walk_unit_pair(batch, tc, wk);
        walk_next(wk, U);
        if (U.skip || U.ntiles == 0) continue;
        for (int k = 0; k < U.ntiles; ++k, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
          const int a = k & 1;                       // accumulator of this tile
          mbar_wait(&full[s], ph);
          tc::mbar_wait_cluster(&pfull[s], ph);
          if (k == 0) tc::mbar_wait_cluster(aready, unit_it & 1u);
          tc::mbar_wait_cluster(&tempty[a], (acc_use[a] & 1u) ^ 1u);
          tc::fence_after_sync();
#pragma unroll
          for (int kk = 0; kk < kChunks / 2; ++kk) {   // one MMA consumes K = 32 bytes = 2 chunks
            const uint64_t ad = tc::smem_desc(a_addr + uint32_t(kk) * 2u * (kQ * 16), lbo, sbo);
            const uint64_t bd = tc::smem_desc(b_addr + s * kBBytes + uint32_t(kk) * 2u * (kT * 16), lbo, sbo);
            tc::mma_ss_2cta<true>(tmem_base + uint32_t(a) * 256u, ad, bd, idesc, kk > 0 ? 1u : 0u);
          }
          tc::commit_2cta(&tfull[a], 3);
          ++acc_use[a];
          tc::commit_2cta(&empty[s], 3);
        }
        ++unit_it;
      }
    } else if (lane == 0) {
      // ------------------------------ relay (rank 1): "my half of the tile has landed" -> rank 0 ------------------------------
      uint32_t it = 0;
      while (walk_more(wk)) {
        const Unit U = walk_unit_pair(batch, tc, wk);
        walk_next(wk, U);
        if (U.skip || U.ntiles == 0) continue;
        for (int k = 0; k < U.ntiles; ++k, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
          mbar_wait(&full[s], ph);
          tc::mbar_arrive_cluster(&pfull[s], 0);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ query expansion + epilogue ------------------------------
    // warp = (column half ch) * 8 + (tile parity h) * 4 + (TMEM lane quarter); a warp reduces
    // 128 of the 256 columns of its parity's accumulator
    const int h = (warp >> 2) & 1;
    const int ch = warp >> 3;
    const int r = (warp & 3) * 32 + lane;          // TMEM lane = query row of this CTA's half of the unit
    const int g = h * 2 + ch;                      // which 16 bytes of the query row this thread expands
    const uint32_t taddr = tmem_base + (uint32_t((warp & 3) * 32) << 16) + uint32_t(h * 256 + ch * (kTP / 2));
    constexpr int NB = kEpiCols / kTcBucket;       // buckets per 64-column load
    constexpr int TB = kTP / kTcBucket;            // buckets per pair tile
    constexpr int RPB = kTcBucket / 2;             // packed registers per bucket
    uint32_t acc_use = 0;

    auto load_query = [&](const Unit& V) -> uint4 {
      uint4 w = make_uint4(0u, 0u, 0u, 0u);
      const int q = V.q0 + int(rank) * kQ + r;
      if (!V.skip && V.ntiles > 0 && q < V.nq)
        w = __ldg(reinterpret_cast<const uint4*>(batch.p[V.problem].q + size_t(q) * 16) + g);
      return w;
    };
    auto expand_query = [&](const uint4& qw) {
      const uint32_t words[4] = {qw.x, qw.y, qw.z, qw.w};
      uint8_t* dst = sA + size_t(r) * 16 + size_t(g * 8) * (kQ * 16);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t b16 = (words[c >> 1] >> (16 * (c & 1))) & 0xFFFFu;
        *reinterpret_cast<uint4*>(dst + size_t(c) * (kQ * 16)) = expand16(b16);
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_cluster(aready, 0);
    };
    Unit U;
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    bool expanded = false;
    if (walk_more(wk)) {
      U = walk_unit_pair(batch, tc, wk);
      w = load_query(U);
      if (!(U.skip || U.ntiles == 0)) {
        expand_query(w);
        expanded = true;
      }
    }
    role_wait(tc);
    while (walk_more(wk)) {
      walk_next(wk, U);
      const bool more = walk_more(wk);
      Unit Un;
      uint4 wn = make_uint4(0u, 0u, 0u, 0u);
      const KnnProblem& P = batch.p[U.problem];
      const int q = U.q0 + int(rank) * kQ + r;
      uint2* part = reinterpret_cast<uint2*>(batch.partial) +
                    (size_t(P.row0 + q) * (tc.slots * kParts) + U.slot * kParts + g);
      if (U.skip || U.ntiles == 0) {
        if (!U.skip && q < U.nq) *part = make_uint2(uint32_t(kKeyNone), uint32_t(kKeyNone));
        if (more) {
          Un = walk_unit_pair(batch, tc, wk);
          wn = load_query(Un);
        }
        U = Un; w = wn;
        continue;
      }
      if (!expanded) expand_query(w);
      expanded = false;
      if (more) {
        Un = walk_unit_pair(batch, tc, wk);
        wn = load_query(Un);
      }
      int m1 = kKeyNone, m2 = kKeyNone;
      for (int k = h; k < U.ntiles; k += 2, ++acc_use) {   // this warp's parity of the unit's tiles
        mbar_wait(&tfull[h], acc_use & 1u);
        tc::fence_after_sync();
#pragma unroll
        for (int half = 0; half < 2; ++half) {             // two loads of 64 packed columns
          const int row0 = U.t_begin + k * kTP + ch * (kTP / 2) + half * kEpiCols;
          const int kbase = kBucketIdMask - row0 / kTcBucket;
          uint32_t v[32];
          tc::tmem_ld_32x32_pack16(taddr + uint32_t(half * kEpiCols), v);
          tc::tmem_ld_wait();
          if (half == 1) {                                  // the accumulator may be overwritten now
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster(&tempty[h], 0);
          }
          const bool full_tile = row0 + kEpiCols <= U.t_end;   // warp-uniform
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            uint32_t* b = &v[j * RPB];
            if (!full_tile) {
              const int valid = U.t_end - (row0 + j * kTcBucket);
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
            const int bm = max(int(short(mm & 0xFFFFu)), int(mm) >> 16);
            const int key = bm * (1 << kBucketIdBits) + (kbase - j);
            m2 = max(m2, min(m1, key));
            m1 = max(m1, key);
          }
        }
      }
      if (q < U.nq) *part = make_uint2(uint32_t(m1), uint32_t(m2));
      // every MMA of the unit has completed (in both CTAs: the commits are multicast) once both
      // parities have seen their last tile: only then may the next unit's queries replace this one's
      epi_barrier();
      U = Un; w = wn;
    }
  }

  tc::fence_before_sync();
  tc::cluster_sync();          // nobody leaves while the partner may still arrive on its barriers
  if (warp == kEpiWarps + 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc_2cta<kTmemCols>(tmem_base);
  }
  if (tc.early) pdl_wait();   // "complete" must still imply "the stream predecessor is complete"
}

}  // namespace pair

// ---------------------------------------------------------------------------------------------
// Refine: as knn2_tc_refine_kernel, for 16-word rows.  Two lanes per query (best / second-best
// bucket); a warp stages 4 rows of every pair per pass (256 contiguous bytes per pair).
constexpr int kRefineQB = 64;
constexpr int kRefineThreads = 2 * kRefineQB;
constexpr int kRefineRows = 4;
constexpr int kRefinePitch = kRefineRows * 64 + 16;
constexpr int kRefinePP = kRefineRows * 4;       // 16-byte pieces of one pair's run of rows (16)
static_assert(kTcBucket % kRefineRows == 0 && 32 % kRefinePP == 0, "bucket must be a multiple of the refine pass");

// (EXACT2 as in knn2_tc_refine_kernel: false = one lane per query, the second neighbour's distance
// from the second-best bucket's exact maximum dot, dot = 512 - 2 * hamming)
template <bool EXACT2>
__global__ void __launch_bounds__(kRefineThreads)
knn2_tc64_refine_kernel(const __grid_constant__ KnnBatch batch, const __grid_constant__ TcBatch tc) {
  constexpr int kQB = EXACT2 ? kRefineQB : 2 * kRefineQB;      // queries per CTA
  __shared__ __align__(16) uint8_t s_stage[kRefineThreads / 32][32 * kRefinePitch];
  const KnnProblem& P = batch.p[blockIdx.y];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  int nq = P.nq, nt = P.nt;
  if (P.nq_dev) nq = min(nq, *P.nq_dev);
  if (P.nt_dev) nt = min(nt, *P.nt_dev);
  const int qb = blockIdx.x;
  const int q0 = qb * kQB;
  const int q = q0 + (EXACT2 ? (tid >> 1) : tid);
  const int c = EXACT2 ? (tid & 1) : 0;
  uint32_t qw[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) qw[i] = 0u;
  if (q < nq) {
    const uint4* src = reinterpret_cast<const uint4*>(P.q + size_t(q) * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 v = __ldg(src + i);
      qw[4 * i] = v.x; qw[4 * i + 1] = v.y; qw[4 * i + 2] = v.z; qw[4 * i + 3] = v.w;
    }
  }
  pdl_wait();
  pdl_launch_dependents();
  if (nq <= 0) {
    if (blockIdx.x == 0 && tid == 0) {
      *P.match_count = 0;
      if (batch.host_counts) batch.host_counts[P.region] = 0;
    }
    return;
  }
  if (q0 >= nq) return;

  int key = kKeyNone, key2 = kKeyNone;
  if (q < nq) {
    int b1 = kKeyNone, b2 = kKeyNone;
    const int nseg = tc_block_segments(tc, tc.qb_begin[blockIdx.y] + q / tc.unit_q);
    const uint4* part = reinterpret_cast<const uint4*>(reinterpret_cast<const uint2*>(batch.partial) +
                                                        size_t(P.row0 + q) * (tc.slots * kParts));
    auto merge = [&](int a1, int a2) {
      const int hi = max(b1, a1), lo = min(b1, a1);
      b2 = max(lo, max(b2, a2));
      b1 = hi;
    };
    const int nrec = nseg * (kParts / 2);      // 16-byte records of two partial pairs
    for (int z0 = 0; z0 < nrec; z0 += 4) {
      uint4 p[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        p[j] = (z0 + j < nrec) ? __ldcg(part + z0 + j)
                               : make_uint4(uint32_t(kKeyNone), uint32_t(kKeyNone), uint32_t(kKeyNone), uint32_t(kKeyNone));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        merge(int(p[j].x), int(p[j].y));
        merge(int(p[j].z), int(p[j].w));
      }
    }
    key = c ? b2 : b1;
    key2 = b2;
  }
  const int my_row0 = (key == kKeyNone) ? -1 : (kBucketIdMask - (key & kBucketIdMask)) * kTcBucket;
  uint32_t k1 = kKeySentinel, k2 = kKeySentinel;
  if (!EXACT2 && key2 != kKeyNone) {
    const int dot = key2 >> kBucketIdBits;
    k1 = (uint32_t((kRowBytes - dot) >> 1) << kIdxBits) + uint32_t((kBucketIdMask - (key2 & kBucketIdMask)) * kTcBucket);
  }
  uint8_t* stage = s_stage[warp];
#pragma unroll 1
  for (int r0 = 0; r0 < kTcBucket; r0 += kRefineRows) {
#pragma unroll
    for (int i = 0; i < kRefinePP; ++i) {
      const int gp = i * (32 / kRefinePP) + lane / kRefinePP;   // pair whose rows this piece belongs to
      const int piece = lane % kRefinePP;
      const int base = __shfl_sync(0xffffffffu, my_row0, gp);
      const int row = base + r0 + (piece >> 2);
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (base >= 0 && row < nt) v = __ldg(reinterpret_cast<const uint4*>(P.t + size_t(row) * 16) + (piece & 3));
      *reinterpret_cast<uint4*>(stage + gp * kRefinePitch + piece * 16) = v;
    }
    __syncwarp();
    if (my_row0 >= 0) {
      const uint4* mine = reinterpret_cast<const uint4*>(stage + lane * kRefinePitch);
#pragma unroll
      for (int k = 0; k < kRefineRows; ++k) {
        const int row = my_row0 + r0 + k;
        uint32_t tw[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 t = mine[4 * k + i];
          tw[4 * i] = t.x; tw[4 * i + 1] = t.y; tw[4 * i + 2] = t.z; tw[4 * i + 3] = t.w;
        }
        const uint32_t d = hamming256<2>(qw, tw) + hamming256<2>(qw + 8, tw + 8);
        if (row < nt) top2_insert(k1, k2, (d << kIdxBits) + uint32_t(row));
      }
    }
    __syncwarp();
  }
  if (EXACT2) {
    const uint32_t o1 = __shfl_xor_sync(0xffffffffu, k1, 1), o2 = __shfl_xor_sync(0xffffffffu, k2, 1);
    top2_merge(k1, k2, o1, o2);
  }
  bool pass = false;
  if (q < nq && c == 0) {
    const int i0 = (k1 == kKeySentinel) ? -1 : int(k1 & kIdxMask);
    const int i1 = (k2 == kKeySentinel) ? -1 : int(k2 & kIdxMask);
    const int d0 = (k1 == kKeySentinel) ? -1 : int(k1 >> kIdxBits);
    const int d1 = (k2 == kKeySentinel) ? -1 : int(k2 >> kIdxBits);
    batch.knn_out[P.row0 + q] = make_uint4(uint32_t(i0), uint32_t(i1), uint32_t(d0), uint32_t(d1));
    pass = (i1 >= 0) && (double(d0) < batch.ratio * double(d1));
  }
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

}  // namespace w64

cudaError_t launch_expand_train64(const void* t, int nt_bound, const int* nt_dev, void* out, int pdl,
                                  cudaStream_t stream) {
  if (nt_bound <= 0) return cudaSuccess;
  const int rows_pad = (nt_bound + w64::kT - 1) / w64::kT * w64::kT;
  const int threads = rows_pad * 8;
  const int blocks = (threads + 127) / 128;
  return w64::launch_pdl(w64::expand_train64_kernel, dim3(blocks), dim3(128), 0, stream, pdl != 0,
                         static_cast<const uint32_t*>(t), nt_bound, nt_dev, static_cast<uint8_t*>(out));
}

// The same for the frames of a group of poses, blockIdx.y = frame (ExpandMulti).
namespace w64 {
__global__ void __launch_bounds__(128)
expand_train64_multi_kernel(const __grid_constant__ ExpandMulti em) {
  pdl_wait();
  pdl_launch_dependents();
  expand_rows64(em.src[blockIdx.y], em.nt, em.out[blockIdx.y], blockIdx.x * blockDim.x + threadIdx.x);
}
}  // namespace w64

cudaError_t launch_expand_train64_multi(const ExpandMulti& em, int pdl, cudaStream_t stream) {
  if (em.frames <= 0 || em.nt <= 0) return cudaSuccess;
  const int rows_pad = (em.nt + w64::kT - 1) / w64::kT * w64::kT;
  const dim3 grid((rows_pad * 8 + 127) / 128, em.frames);
  return w64::launch_pdl(w64::expand_train64_multi_kernel, grid, dim3(128), 0, stream, pdl != 0, em);
}

// ev, fa, launched, phase: as launch_knn2_tc
cudaError_t launch_knn2_tc64(const KnnBatch& batch, const TcBatch& tc, int max_nq, int pdl, cudaEvent_t* ev,
                             cudaStream_t stream, FinishArgs* fa, int* launched, int phase) {
  if (launched) *launched = 0;
  if (batch.num_problems <= 0 || tc.total <= 0) return cudaSuccess;
  // per-device function attributes, set once per device (one bit each)
  static std::atomic<unsigned long long> smem_done{0ull};
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  const bool attrs_set = (smem_done.load(std::memory_order_acquire) & bit) != 0;
  cudaError_t e = cudaSuccess;
  if (!attrs_set) {
    e = cudaFuncSetAttribute(w64::knn2_tc64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, w64::kSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(w64::pair::knn2_tc64_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, w64::pair::kSmemBytesP);
    if (e != cudaSuccess) return e;
    smem_done.fetch_or(bit, std::memory_order_release);
  }
  const bool p = pdl != 0;
  if (ev) cudaEventRecord(ev[0], stream);
  if (phase != 2) {
    if (tc.unit_q == 2 * w64::kQ) {
      // CTA pairs: tc.grid counts pairs; cluster dimensions are a compile-time attribute of the kernel
      e = w64::launch_pdl(w64::pair::knn2_tc64_pair_kernel, dim3(2 * tc.grid), dim3(w64::kThreads), w64::pair::kSmemBytesP, stream, p,
                          batch, tc);
    } else {
      e = w64::launch_pdl(w64::knn2_tc64_kernel, dim3(tc.grid), dim3(w64::kThreads), w64::kSmemBytes, stream, p, batch, tc);
    }
    if (e != cudaSuccess) return e;
    if (launched) *launched = 1;
  }
  if (ev) cudaEventRecord(ev[1], stream);
  if (phase == 1) return cudaSuccess;
  if (fa && !batch.exact_second) {
    e = launch_knn2_tc64_finish(batch, tc, max_nq, p, stream, fa);
    if (ev) {
      cudaEventRecord(ev[2], stream);
      cudaEventRecord(ev[3], stream);
    }
    if (launched) *launched = phase == 2 ? 1 : 2;
    return e;
  }
  if (phase == 2) return cudaErrorInvalidValue;
  if (batch.exact_second) {
    dim3 rgrid((max_nq + w64::kRefineQB - 1) / w64::kRefineQB, batch.num_problems);
    e = w64::launch_pdl(w64::knn2_tc64_refine_kernel<true>, rgrid, dim3(w64::kRefineThreads), 0, stream, p, batch, tc);
  } else {
    dim3 rgrid((max_nq + 2 * w64::kRefineQB - 1) / (2 * w64::kRefineQB), batch.num_problems);
    e = w64::launch_pdl(w64::knn2_tc64_refine_kernel<false>, rgrid, dim3(w64::kRefineThreads), 0, stream, p, batch, tc);
  }
  if (e != cudaSuccess) return e;
  if (ev) cudaEventRecord(ev[2], stream);
  e = launch_knn2_compact(batch, max_nq, p, stream);
  if (ev) cudaEventRecord(ev[3], stream);
  if (launched) *launched = 3;
  return e;
}

}  // namespace vsf
