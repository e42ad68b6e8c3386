// Device-side replacement for the tail of Frontend::GetFeatureMatches
// (src/slam_frontend.cc:289-296): order the ratio survivors by distance, keep the
// first int(n * best_percent), emit FeatureMatch(queryIdx, trainIdx).
//
// Distances are small integers (<= 512), so the order is produced by a STABLE
// counting sort: the result is ordered by (distance, queryIdx).  The reference
// uses std::sort, which is not stable; the two orders differ only inside groups
// of equal distance.  sort_mode 1 of the C ABI keeps the
// reference's exact std::sort sequence on the host instead.
//
// One CTA per problem.  Each of the 16 warps owns a contiguous slice of the
// (query-ordered) input and a private 513-bin histogram; bin starts are the
// exclusive scan over (bin, warp); the scatter walks each slice in order with
// __match_any_sync ranks, which keeps the sort stable without atomics.
#include "vsf_device.cuh"

namespace vsf {

constexpr int kSortThreads = 512;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kBins = 513;

struct SortArgs {
  const vsf_dmatch* matches[kMaxProblems];
  const int* counts[kMaxProblems];
  vsf_feature_match* out;
  int out_stride;
  int* out_counts;
  float best_percent;
  int bins;   // distances are 0 .. 8 * row_bytes: 257 bins for 32-byte descriptors, 513 for 64-byte ones
};

__global__ void __launch_bounds__(kSortThreads) sort_cut_kernel(const __grid_constant__ SortArgs a) {
  __shared__ uint32_t s_hist[kSortWarps][kBins];
  __shared__ uint32_t s_start[kBins + 31];
  const int p = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const vsf_dmatch* m = a.matches[p];
  const int nbins = a.bins;
  // clearing the histograms overlaps the tail of the compaction kernel (programmatic dependent launch)
  for (int b = lane; b < nbins; b += 32) s_hist[warp][b] = 0;   // every warp clears its own histogram
  pdl_wait();
  pdl_launch_dependents();
  const int n = *a.counts[p];
  // `matches.size() * config_.best_percent_` truncated to int (src/slam_frontend.cc:290)
  const int keep = int(__fmul_rn(float(size_t(n)), a.best_percent));
  vsf_feature_match* out = a.out + size_t(p) * a.out_stride;
  __syncthreads();

  const int seg = ((n + kSortWarps - 1) / kSortWarps + 31) & ~31;
  const int beg = min(n, warp * seg), end = min(n, beg + seg);

  for (int i0 = beg; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool ok = i < end;
    const uint32_t d = ok ? uint32_t(int(m[i].distance)) : (0xFFFF0000u + lane);
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if (ok && (peers & ((1u << lane) - 1u)) == 0) s_hist[warp][min(d, uint32_t(nbins - 1))] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // per bin: exclusive offsets across warps, bin totals
  for (int b = tid; b < nbins; b += kSortThreads) {
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t t = s_hist[w][b];
      s_hist[w][b] = tot;
      tot += t;
    }
    s_start[b] = tot;
  }
  __syncthreads();
  if (warp == 0) {
    uint32_t run = 0;
    for (int b0 = 0; b0 < nbins; b0 += 32) {
      const int b = b0 + lane;
      const uint32_t c = (b < nbins) ? s_start[b] : 0u;
      uint32_t incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (b < nbins) s_start[b] = run + incl - c;
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  __syncthreads();
  for (int i0 = beg; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool ok = i < end;
    vsf_dmatch dm = {0, 0, 0, 0.f};
    if (ok) dm = m[i];
    const uint32_t d = ok ? min(uint32_t(int(dm.distance)), uint32_t(nbins - 1)) : (0xFFFF0000u + lane);
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (ok) {
      const uint32_t pos = s_start[d] + s_hist[warp][d] + rank;
      if (pos < uint32_t(keep)) {
        ulonglong2 fm;
        fm.x = uint64_t(uint32_t(dm.queryIdx));   // feature_idx_initial (src/slam_frontend.cc:295)
        fm.y = uint64_t(uint32_t(dm.trainIdx));   // feature_idx_current (:296)
        reinterpret_cast<ulonglong2*>(out)[pos] = fm;
      }
    }
    __syncwarp();
    if (ok && rank == 0) s_hist[warp][d] += __popc(peers);
    __syncwarp();
  }
  if (tid == 0) a.out_counts[p] = keep;
}

cudaError_t launch_sort_cut(const vsf_dmatch* const* matches, const int* const* counts, int n_problems,
                            float best_percent, vsf_feature_match* out, int out_stride, int* out_counts,
                            int bins, cudaStream_t stream) {
  if (n_problems <= 0) return cudaSuccess;
  if (n_problems > kMaxProblems) return cudaErrorInvalidValue;
  SortArgs a;
  for (int i = 0; i < n_problems; ++i) {
    a.matches[i] = matches[i];
    a.counts[i] = counts[i];
  }
  a.out = out;
  a.out_stride = out_stride;
  a.out_counts = out_counts;
  a.best_percent = best_percent;
  a.bins = bins > 1 && bins <= kBins ? bins : kBins;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_problems);
  cfg.blockDim = dim3(kSortThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, sort_cut_kernel, a);
}

}  // namespace vsf
