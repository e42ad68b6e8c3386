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
// every bucket of 32 consecutive train rows (3-input VIMNMX3 / FMNMX3: 16 instructions per 32
// comparisons), then the top-2 BUCKETS per query ordered by (max dot desc, bucket asc).  The
// best neighbour (lowest distance, lowest train index on ties) is always inside the best
// bucket, and the second neighbour is inside the best or the second-best bucket, so the
// refine kernel recomputes exact Hamming distances for just those <= 64 train rows per query
// on the integer pipe (XOR/POPC, lexicographic packed keys) and the result is bit-identical
// to the POPC engine and to OpenCV.
//
// Kernels (all on the ctx stream):
//   expand_train_kernel   packed train rows -> +-1 bytes, written tile by tile in the exact
//                         shared-memory image the UMMA descriptors expect (K-major, no swizzle:
//                         [16 K-chunks][256 rows][16 B]), so one 64 KB 1-D TMA bulk copy
//                         (cp.async.bulk, SASS UBLKCP) stages a tile.
//   knn2_tc_kernel        persistent, warp-specialised: warp 8 = TMA producer, warp 9 = MMA
//                         issuer (one elected thread), warps 0-7 = query expansion into smem +
//                         TMEM epilogue.  Work unit = 256 queries x one train split.
//   knn2_tc_refine_kernel exact top-2 inside the candidate buckets, ratio test, compaction.
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

// ---------------------------------------------------------------------------------------------
// Train expansion.  One thread = one (row, group of 4 K-chunks).  Rows >= nt of the last tile
// are written as zeros (their columns are masked in the epilogue).
template <bool I8>
__global__ void __launch_bounds__(256)
expand_train_kernel(const uint32_t* __restrict__ t, int nt_bound, const int* __restrict__ nt_dev,
                    uint8_t* __restrict__ out) {
  const int rows_pad = (nt_bound + kTcTileRows - 1) / kTcTileRows * kTcTileRows;
  int nt = nt_bound;
  if (nt_dev) nt = min(nt, *nt_dev);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = idx % rows_pad;
  const int cg = idx / rows_pad;  // 0..3: K-chunks 4*cg .. 4*cg+3  (words 2*cg, 2*cg+1)
  if (cg >= 4) return;
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
constexpr int kTcEpiWarps = 8;
constexpr int kTcThreads = (kTcEpiWarps + 2) * 32;  // + TMA producer warp + MMA warp
constexpr int kTcBarriers = 2 * kTcStages + 6;      // full/empty per stage, tfull/tempty/aready per half
constexpr int kTcSmemBytes = kTcABytes + kTcStages * kTcBBytes + kTcBarriers * 8 + 16 + 128;
constexpr int kTcTmemCols = 512;                    // 2 accumulator tiles of 128 lanes x 256 columns
constexpr int kBucketIdBits = 20;
constexpr int kBucketIdMask = (1 << kBucketIdBits) - 1;
constexpr int kTcKeySentinel = int(0x80000000u);    // INT_MIN: "no bucket"

struct TcUnit {
  int problem, qb, z;
  int nq, nt;
  int q0;
  int t_begin, t_end, ntiles;
  bool skip;   // no queries in this block
};

__device__ __forceinline__ TcUnit decode_unit(const KnnBatch& batch, const TcBatch& tc, int u) {
  TcUnit U;
  int p = 0;
  while (u >= tc.unit_begin[p + 1]) ++p;
  U.problem = p;
  const int local = u - tc.unit_begin[p];
  U.qb = local / tc.split;
  U.z = local % tc.split;
  const KnnProblem& P = batch.p[p];
  int nq = P.nq, nt = P.nt;
  if (P.nq_dev) nq = min(nq, *P.nq_dev);
  if (P.nt_dev) nt = min(nt, *P.nt_dev);
  U.nq = nq;
  U.nt = nt;
  U.q0 = U.qb * kTcQ;
  U.skip = U.q0 >= nq;
  U.t_begin = min(nt, U.z * tc.rows_per_split);
  U.t_end = min(nt, U.t_begin + tc.rows_per_split);
  U.ntiles = (U.t_end - U.t_begin + kTcTileRows - 1) / kTcTileRows;
  return U;
}

template <bool I8>
__device__ __forceinline__ int bucket_max(const uint32_t (&v)[32]) {
  if (I8) {
    int m[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      m[c] = int(v[8 * c]);
#pragma unroll
      for (int i = 1; i < 8; ++i) m[c] = max(m[c], int(v[8 * c + i]));
    }
    return max(max(m[0], m[1]), max(m[2], m[3]));
  } else {
    float m[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      m[c] = __uint_as_float(v[8 * c]);
#pragma unroll
      for (int i = 1; i < 8; ++i) m[c] = fmaxf(m[c], __uint_as_float(v[8 * c + i]));
    }
    return __float2int_rn(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])));
  }
}

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

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mbar_init(&tfull[h], 1);
      mbar_init(&tempty[h], 4);
      mbar_init(&aready[h], 4);
    }
    mbar_fence_init();
  }
  if (warp == kTcEpiWarps + 1) tc::tmem_alloc<kTcTmemCols>(s_tmem);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  // K-major, no swizzle: [K-chunk c][row][16 B]; chunk stride = rows * 16, 8-row group stride = 128
  const uint32_t lbo = tc.swap_lbo_sbo ? 128u : uint32_t(kTcQ * 16);
  const uint32_t sbo = tc.swap_lbo_sbo ? uint32_t(kTcQ * 16) : 128u;

  if (warp == kTcEpiWarps) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int u = blockIdx.x; u < tc.total_units; u += gridDim.x) {
        const TcUnit U = decode_unit(batch, tc, u);
        if (U.skip) continue;
        const uint8_t* src = tc.t_exp[U.problem] + size_t(U.t_begin / kTcTileRows) * kTcBBytes;
        for (int k = 0; k < U.ntiles; ++k, ++it) {
          const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], kTcBBytes);
          tma_load_1d(sB + size_t(s) * kTcBBytes, src + size_t(k) * kTcBBytes, kTcBBytes, &full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == kTcEpiWarps + 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = tc::instr_desc(I8, 128, kTcTileRows);
      uint32_t it = 0, unit_it = 0, acc_use[2] = {0u, 0u};
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
      for (int u = blockIdx.x; u < tc.total_units; u += gridDim.x) {
        const TcUnit U = decode_unit(batch, tc, u);
        if (U.skip || U.ntiles == 0) continue;
        for (int k = 0; k < U.ntiles; ++k, ++it) {
          const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
          mbar_wait(&full[s], ph);
          tc::fence_after_sync();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (k == 0) mbar_wait(&aready[h], unit_it & 1u);
            mbar_wait(&tempty[h], (acc_use[h] & 1u) ^ 1u);
            tc::fence_after_sync();
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
    }
    __syncwarp();
  } else {
    // ------------------------------ query expansion + epilogue ------------------------------
    const int h = warp >> 2;                       // accumulator half = query rows 128h .. 128h+127
    const int r = (warp & 3) * 32 + lane;          // TMEM lane = row inside the half
    const uint32_t taddr = tmem_base + (uint32_t((warp & 3) * 32) << 16) + uint32_t(h) * kTcTileRows;
    uint32_t acc_use = 0;
    for (int u = blockIdx.x; u < tc.total_units; u += gridDim.x) {
      const TcUnit U = decode_unit(batch, tc, u);
      if (U.skip) continue;
      const KnnProblem& P = batch.p[U.problem];
      const int q = U.q0 + h * 128 + r;
      uint2* part = reinterpret_cast<uint2*>(batch.partial) + (size_t(P.row0 + q) * tc.split + U.z);
      if (U.ntiles == 0) {
        if (q < U.nq) *part = make_uint2(uint32_t(kTcKeySentinel), uint32_t(kTcKeySentinel));
        continue;
      }
      // expand this thread's query row into the A image
      {
        uint4 w0 = make_uint4(0u, 0u, 0u, 0u), w1 = w0;
        if (q < U.nq) {
          const uint4* src = reinterpret_cast<const uint4*>(P.q + size_t(q) * 8);
          w0 = __ldg(src);
          w1 = __ldg(src + 1);
        }
        const uint32_t words[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        uint8_t* dst = sA + size_t(h * 128 + r) * 16;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const uint32_t b16 = (words[c >> 1] >> (16 * (c & 1))) & 0xFFFFu;
          *reinterpret_cast<uint4*>(dst + size_t(c) * (kTcQ * 16)) = expand16<I8>(b16);
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&aready[h]);
      }
      int m1 = kTcKeySentinel, m2 = kTcKeySentinel;
      const int bucket0 = U.t_begin / kTcBucket;
      for (int k = 0; k < U.ntiles; ++k, ++acc_use) {
        mbar_wait(&tfull[h], acc_use & 1u);
        tc::fence_after_sync();
        const int tile_row0 = U.t_begin + k * kTcTileRows;
#pragma unroll 2
        for (int j = 0; j < kTcTileRows / kTcBucket; ++j) {
          uint32_t v[32];
          tc::tmem_ld_32x32(taddr + uint32_t(j * kTcBucket), v);
          tc::tmem_ld_wait();
          const int brow0 = tile_row0 + j * kTcBucket;
          const int valid = U.t_end - brow0;       // train rows of this bucket that exist
          if (valid <= 0) continue;                // warp-uniform
          if (valid < kTcBucket) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i >= valid) v[i] = I8 ? uint32_t(kTcKeySentinel) : 0xFF800000u;  // INT_MIN / -inf
          }
          const int bm = bucket_max<I8>(v);
          const int key = bm * (1 << kBucketIdBits) + (kBucketIdMask - (bucket0 + k * (kTcTileRows / kTcBucket) + j));
          m2 = max(m2, min(m1, key));
          m1 = max(m1, key);
        }
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[h]);
      }
      if (q < U.nq) *part = make_uint2(uint32_t(m1), uint32_t(m2));
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == kTcEpiWarps + 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc<kTcTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Refine: one thread per query.  Merge the per-split top-2 buckets, recompute exact Hamming
// distances inside them (packed keys, lowest train index wins ties), then the shared tail.
constexpr int kRefineQB = 128;

__global__ void __launch_bounds__(kRefineQB)
knn2_tc_refine_kernel(const __grid_constant__ KnnBatch batch, const __grid_constant__ TcBatch tc) {
  __shared__ TailSmem s_tail;
  const KnnProblem& P = batch.p[blockIdx.y];
  const int tid = threadIdx.x;
  int nq = P.nq, nt = P.nt;
  if (P.nq_dev) nq = min(nq, *P.nq_dev);
  if (P.nt_dev) nt = min(nt, *P.nt_dev);
  if (nq <= 0) {
    if (blockIdx.x == 0 && tid == 0) *P.match_count = 0;
    return;
  }
  const int qb = blockIdx.x;
  const int q0 = qb * kRefineQB;
  if (q0 >= nq) return;
  const int nqb = (nq + kRefineQB - 1) / kRefineQB;
  const int q = q0 + tid;

  uint32_t k1 = kKeySentinel, k2 = kKeySentinel;
  if (q < nq) {
    int b1 = kTcKeySentinel, b2 = kTcKeySentinel;
    const uint2* part = reinterpret_cast<const uint2*>(batch.partial) + size_t(P.row0 + q) * tc.split;
    for (int z = 0; z < tc.split; ++z) {
      const uint2 p = __ldcg(part + z);
      const int a1 = int(p.x), a2 = int(p.y);
      // merge two descending pairs
      const int hi = max(b1, a1), lo = min(b1, a1);
      b2 = max(lo, max(b2, a2));
      b1 = hi;
    }
    uint32_t qw[8];
    {
      const uint4* src = reinterpret_cast<const uint4*>(P.q + size_t(q) * 8);
      const uint4 w0 = __ldg(src), w1 = __ldg(src + 1);
      qw[0] = w0.x; qw[1] = w0.y; qw[2] = w0.z; qw[3] = w0.w;
      qw[4] = w1.x; qw[5] = w1.y; qw[6] = w1.z; qw[7] = w1.w;
    }
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int key = c ? b2 : b1;
      if (key == kTcKeySentinel) continue;
      const int bucket = kBucketIdMask - (key & kBucketIdMask);
      const int r0 = bucket * kTcBucket;
      const int r1 = min(nt, r0 + kTcBucket);
      for (int row = r0; row < r1; ++row) {
        const uint4* tp = reinterpret_cast<const uint4*>(P.t + size_t(row) * 8);
        const uint4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
        const uint32_t tw[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
        const uint32_t d = hamming256<2>(qw, tw);
        top2_insert(k1, k2, (d << kIdxBits) + uint32_t(row));
      }
    }
  }
  finalize_and_compact<kRefineQB, kRefineQB>(batch, P, blockIdx.y, qb, nqb, nq, k1, k2, s_tail);
}

// ---------------------------------------------------------------------------------------------
cudaError_t launch_expand_train(const void* t, int nt_bound, const int* nt_dev, void* out, int int8,
                                cudaStream_t stream) {
  if (nt_bound <= 0) return cudaSuccess;
  const int rows_pad = (nt_bound + kTcTileRows - 1) / kTcTileRows * kTcTileRows;
  const int threads = rows_pad * 4;
  const int blocks = (threads + 255) / 256;
  if (int8)
    expand_train_kernel<true><<<blocks, 256, 0, stream>>>(static_cast<const uint32_t*>(t), nt_bound, nt_dev,
                                                          static_cast<uint8_t*>(out));
  else
    expand_train_kernel<false><<<blocks, 256, 0, stream>>>(static_cast<const uint32_t*>(t), nt_bound, nt_dev,
                                                           static_cast<uint8_t*>(out));
  return cudaGetLastError();
}

cudaError_t launch_knn2_tc(const KnnBatch& batch, const TcBatch& tc, int int8, int grid, int max_nq,
                           cudaStream_t stream) {
  if (batch.num_problems <= 0 || tc.total_units <= 0) return cudaSuccess;
  cudaError_t e;
  // per-device attribute; setting it is a cheap host-side call
  e = int8 ? cudaFuncSetAttribute(knn2_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes)
           : cudaFuncSetAttribute(knn2_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
  if (e != cudaSuccess) return e;
  if (int8)
    knn2_tc_kernel<true><<<grid, kTcThreads, kTcSmemBytes, stream>>>(batch, tc);
  else
    knn2_tc_kernel<false><<<grid, kTcThreads, kTcSmemBytes, stream>>>(batch, tc);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  dim3 rgrid((max_nq + kRefineQB - 1) / kRefineQB, batch.num_problems);
  knn2_tc_refine_kernel<<<rgrid, kRefineQB, 0, stream>>>(batch, tc);
  return cudaGetLastError();
}

}  // namespace vsf
