// Shared tail of the kNN kernels: unpack the final top-2 keys of a query block, apply the
// Lowe ratio test (src/slam_frontend.cc:529-536) and let the last CTA of a problem compact
// the survivors in ascending query order (what Frontend::GetMatches returns).
#pragma once

#include "vsf_device.cuh"

namespace vsf {

constexpr int kScanChunk = 256;

struct TailSmem {
  uint32_t off[kScanChunk];
  int flag;     // this CTA is the last one of its problem
  int total;    // running survivor count after the chunk being scanned (its own word: warp 0 may
                // write it while slower warps are still reading `flag`)
};

// Called by every thread of the CTA (blockDim.x = NTHREADS >= QB, a multiple of 32).
// Thread tid < QB owns query q0 + tid with final keys (k1, k2) = packed
// (distance << kIdxBits | trainIdx), kKeySentinel when the neighbour does not exist.
// qb = index of this query block inside the problem, nqb = number of query blocks.
template <int QB, int NTHREADS>
__device__ __forceinline__ void finalize_and_compact(const KnnBatch& batch, const KnnProblem& P,
                                                     int problem, int qb, int nqb, int nq, uint32_t k1,
                                                     uint32_t k2, TailSmem& sm) {
  constexpr int NWARPS = NTHREADS / 32;
  constexpr int GROUPS = QB / 32;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int q0 = qb * QB;
  const int row = P.row0 + q0 + tid;

  bool pass = false;
  if (tid < QB && q0 + tid < nq) {
    const int i0 = (k1 == kKeySentinel) ? -1 : int(k1 & kIdxMask);
    const int i1 = (k2 == kKeySentinel) ? -1 : int(k2 & kIdxMask);
    const int d0 = (k1 == kKeySentinel) ? -1 : int(k1 >> kIdxBits);
    const int d1 = (k2 == kKeySentinel) ? -1 : int(k2 >> kIdxBits);
    batch.knn_out[row] = make_uint4(uint32_t(i0), uint32_t(i1), uint32_t(d0), uint32_t(d1));
    // `dist1 < nn_match_ratio * dist2` in double (src/slam_frontend.cc:533);
    // fewer than 2 train rows: no match passes.
    pass = (i1 >= 0) && (double(d0) < batch.ratio * double(d1));
    __threadfence();
  }
  const int npass = __syncthreads_count(pass);
  if (tid == 0) {
    batch.qblock_pass[P.qb0 + qb] = unsigned(npass);
    __threadfence();
    const unsigned prev = atomicAdd(&batch.problem_arrivals[problem], 1u);
    const int last = (prev == unsigned(nqb - 1));
    if (last) batch.problem_arrivals[problem] = 0u;  // self-reset for the next launch
    sm.flag = last;
  }
  __syncthreads();
  if (!sm.flag) return;
  __threadfence();

  // the problem's last CTA compacts survivors in query order
  int4* hostm = batch.host_matches
                    ? reinterpret_cast<int4*>(batch.host_matches + size_t(P.region) * batch.host_region_stride)
                    : nullptr;
  uint32_t base = 0;
  for (int c0 = 0; c0 < nqb; c0 += kScanChunk) {
    const int cn = min(kScanChunk, nqb - c0);
    if (warp == 0) {
      // exclusive scan of the per-query-block survivor counts of this chunk
      uint32_t run = base;
      for (int i = lane; i < ((cn + 31) & ~31); i += 32) {
        const uint32_t c = (i < cn) ? __ldcg(&batch.qblock_pass[P.qb0 + c0 + i]) : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += n;
        }
        if (i < cn) sm.off[i] = run + incl - c;
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) sm.total = int(run);
    }
    __syncthreads();
    for (int b = warp; b < cn; b += NWARPS) {
      uint32_t off = sm.off[b];
#pragma unroll
      for (int g = 0; g < GROUPS; ++g) {
        const int q = (c0 + b) * QB + g * 32 + lane;
        bool ok = false;
        uint4 rec = make_uint4(0, 0, 0, 0);
        if (q < nq) {
          rec = __ldcg(&batch.knn_out[P.row0 + q]);
          ok = (int(rec.y) >= 0) && (double(int(rec.z)) < batch.ratio * double(int(rec.w)));
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (ok) {
          const uint32_t dst = off + __popc(bal & ((1u << lane) - 1u));
          int4 m;
          m.x = q;            // queryIdx
          m.y = int(rec.x);   // trainIdx
          m.z = 0;            // imgIdx
          m.w = __float_as_int(float(int(rec.z)));  // distance
          reinterpret_cast<int4*>(P.matches)[dst] = m;
          if (hostm) hostm[dst] = m;
        }
        off += __popc(bal);
      }
    }
    __syncthreads();
    base = uint32_t(sm.total);
    __syncthreads();
  }
  if (tid == 0) {
    *P.match_count = int(base);
    if (batch.host_counts) batch.host_counts[P.region] = int(base);
  }
}

}  // namespace vsf
