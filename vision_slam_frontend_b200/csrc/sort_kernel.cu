// Device-side replacement for the tail of Frontend::GetFeatureMatches
// (src/slam_frontend.cc:289-296): order the ratio survivors by distance, keep the
// first int(n * best_percent), emit FeatureMatch(queryIdx, trainIdx).
//
// Two orders, one kernel (one CTA per frame pair, the list's sort keys
// `distance << 22 | position` in shared memory):
//
//  * stable (sort_mode 0): a counting sort by distance, i.e. ordered by (distance, queryIdx).
//  * exact  (sort_mode 2): the order the reference's std::sort produces.  std::sort is libstdc++'s
//    introsort, which is not stable: which of several equal-distance matches end up before the
//    best_percent cut depends on its sequence of swaps.  The kernel replays
//    that sequence - same median-of-3 pivots, same unguarded Hoare partitions, same depth limit
//    2 * floor(lg n) with the heapsort fallback, same 16-element threshold - and then applies the
//    stable counting sort, because libstdc++'s closing __final_insertion_sort IS a stable sort
//    of whatever the partition phase left.  Like the host version (exact_sort.h) it only descends
//    into ranges that can reach the kept prefix.
//
//    What makes the replay parallel: libstdc++'s partition loop
//        while (true) { while (*first < pivot) ++first;  --last;  while (pivot < *last) --last;
//                       if (!(first < last)) return first;  iter_swap(first, last);  ++first; }
//    pairs the k-th element from the left that is not < pivot (position L[k]) with the k-th from
//    the right that is not > pivot (R[k]) and swaps them while L[k] < R[k]; with k* the first k
//    that fails, it returns min(L[k*], R[k*-1]).  Both lists are functions of the range's contents
//    BEFORE the step, so a warp builds them with ballots and prefix counts, finds k* and does all
//    swaps at once.  Ranges are independent once split: the warps of the CTA take them from a
//    queue in shared memory.  tools/exact_sort_model.py is the numpy model of this formulation;
//    both are checked against the real std::sort (oracle/stdsort_oracle.cc).
#include <algorithm>
#include <atomic>

#include "vsf_device.cuh"

namespace vsf {

constexpr int kSortThreads = 512;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kBins = 513;
constexpr int kSortThreshold = 16;          // std::_S_threshold
constexpr int kSortMaxRows = 24576;         // exact mode: keys + the two position lists must fit in shared memory
constexpr unsigned long long kQValid = 1ull << 63;

struct SortArgs {
  const vsf_dmatch* matches[kMaxProblems];
  const int* counts[kMaxProblems];
  vsf_feature_match* out;
  int out_stride;
  int* out_counts;
  float best_percent;
  int bins;    // distances are 0 .. 8 * row_bytes: 257 bins for 32-byte descriptors, 513 for 64-byte ones
  int exact;   // replay libstdc++'s introsort before the stable pass
  int cap;     // keys the dynamic shared memory holds (>= every list's length)
  int qmask;   // queue entries - 1 (power of two >= cap / 16)
  int depth_override;   // tests: force introsort's depth limit (-1 = 2 * floor(lg n))
};

struct SortShared {
  uint32_t hist[kSortWarps][kBins];
  uint32_t start[kBins + 31];
  unsigned q_head, q_tail;
  int pending;
  int sorted_end;
};

__device__ __forceinline__ uint32_t key_dist(uint32_t k) { return k >> kIdxBits; }

// std::__adjust_heap + std::__push_heap on keys compared by distance only
__device__ void heap_adjust(uint32_t* v, int hole, int len, uint32_t value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (key_dist(v[child]) < key_dist(v[child - 1])) --child;
    v[hole] = v[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    v[hole] = v[child - 1];
    hole = child - 1;
  }
  int parent = (hole - 1) / 2;
  while (hole > top && key_dist(v[parent]) < key_dist(value)) {
    v[hole] = v[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  v[hole] = value;
}

// std::__partial_sort(first, last, last): make_heap + sort_heap (one thread; the depth limit is
// practically never reached on tie-heavy distance lists, but the replay must cover it)
__device__ void heap_sort(uint32_t* v, int len) {
  if (len < 2) return;
  for (int parent = (len - 2) / 2;; --parent) {
    heap_adjust(v, parent, len, v[parent]);
    if (parent == 0) break;
  }
  for (int last = len - 1; last >= 1; --last) {
    const uint32_t value = v[last];
    v[last] = v[0];
    heap_adjust(v, 0, last, value);
  }
}

// std::__unguarded_partition_pivot on keys[f, l) by one warp; returns the cut.
__device__ int warp_partition(uint32_t* keys, uint16_t* Ll, uint16_t* Rl, int f, int l, int lane) {
  const unsigned lt = (1u << lane) - 1u;
  // std::__move_median_to_first(first, first + 1, mid, last - 1)
  {
    const int ia = f + 1, ib = f + (l - f) / 2, ic = l - 1;
    const uint32_t da = key_dist(keys[ia]), db = key_dist(keys[ib]), dc = key_dist(keys[ic]);
    int pick;
    if (da < db) pick = (db < dc) ? ib : ((da < dc) ? ic : ia);
    else pick = (da < dc) ? ia : ((db < dc) ? ic : ib);
    __syncwarp();
    if (lane == 0) {
      const uint32_t t = keys[f];
      keys[f] = keys[pick];
      keys[pick] = t;
    }
    __syncwarp();
  }
  const uint32_t piv = key_dist(keys[f]);
  // L: positions of [f + 1, l) whose element is not < pivot, left to right
  int cntL = 0;
  for (int i0 = f + 1; i0 < l; i0 += 32) {
    const int p = i0 + lane;
    const bool fl = p < l && key_dist(keys[p]) >= piv;
    const unsigned b = __ballot_sync(0xffffffffu, fl);
    if (fl) Ll[f + cntL + __popc(b & lt)] = uint16_t(p);
    cntL += __popc(b);
  }
  // R: positions whose element is not > pivot, right to left
  int cntR = 0;
  for (int i0 = l - 1; i0 > f; i0 -= 32) {
    const int p = i0 - lane;
    const bool fl = p > f && key_dist(keys[p]) <= piv;
    const unsigned b = __ballot_sync(0xffffffffu, fl);
    if (fl) Rl[f + cntR + __popc(b & lt)] = uint16_t(p);
    cntR += __popc(b);
  }
  __syncwarp();
  // k* = number of leading pairs with L[k] < R[k] (the predicate is monotone)
  const int mn = min(cntL, cntR);
  int ks = 0;
  for (int k0 = 0; k0 < mn; k0 += 32) {
    const int k = k0 + lane;
    const bool ok = k < mn && Ll[f + k] < Rl[f + k];
    const unsigned b = __ballot_sync(0xffffffffu, ok);
    ks += __popc(b);
    if (b != 0xffffffffu) break;
  }
  for (int k = lane; k < ks; k += 32) {
    const int a = Ll[f + k], c = Rl[f + k];
    const uint32_t t = keys[a];
    keys[a] = keys[c];
    keys[c] = t;
  }
  const int cl = ks < cntL ? int(Ll[f + ks]) : 0x7fffffff;
  const int cr = ks >= 1 ? int(Rl[f + ks - 1]) : 0x7fffffff;
  __syncwarp();
  return min(cl, cr);
}

__device__ __forceinline__ unsigned long long q_pack(int f, int l, int d) {
  return kQValid | (static_cast<unsigned long long>(unsigned(d)) << 40) | (static_cast<unsigned long long>(unsigned(l)) << 20) |
         static_cast<unsigned long long>(unsigned(f));
}

__global__ void __launch_bounds__(kSortThreads) sort_cut_kernel(const __grid_constant__ SortArgs a) {
  __shared__ SortShared sm;
  extern __shared__ __align__(16) uint8_t s_dyn[];
  uint32_t* keys = reinterpret_cast<uint32_t*>(s_dyn);
  const int p = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const vsf_dmatch* m = a.matches[p];
  const int nbins = a.bins;
  // clearing the histograms overlaps the tail of the compaction kernel (programmatic dependent launch)
  for (int b = lane; b < nbins; b += 32) sm.hist[warp][b] = 0;   // every warp clears its own histogram
  pdl_wait();
  pdl_launch_dependents();
  int n = *a.counts[p];
  // `matches.size() * config_.best_percent_` truncated to int (src/slam_frontend.cc:290)
  const int keep = int(__fmul_rn(float(size_t(n)), a.best_percent));
  vsf_feature_match* out = a.out + size_t(p) * a.out_stride;
  if (n > a.cap) {   // cannot happen: the host sizes the shared memory for the longest possible list
    if (tid == 0) a.out_counts[p] = -1;
    return;
  }
  for (int i = tid; i < n; i += kSortThreads) keys[i] = (uint32_t(int(m[i].distance)) << kIdxBits) | uint32_t(i);

  if (a.exact && n > kSortThreshold && keep > 0) {
    // ---- replay of std::__introsort_loop, restricted to ranges that reach below `keep`
    uint16_t* Ll = reinterpret_cast<uint16_t*>(keys + a.cap);
    uint16_t* Rl = Ll + a.cap;
    volatile unsigned long long* q = reinterpret_cast<volatile unsigned long long*>(Rl + a.cap);
    for (int i = tid; i <= a.qmask; i += kSortThreads) q[i] = 0ull;
    if (tid == 0) {
      sm.q_head = 0;
      sm.q_tail = 0;
      sm.pending = 1;
      sm.sorted_end = n;
    }
    __syncthreads();
    if (tid == 0) {
      const int depth = a.depth_override >= 0 ? a.depth_override : 2 * (31 - __clz(n));
      q[0] = q_pack(0, n, depth);
      sm.q_tail = 1;
    }
    volatile unsigned* vhead = &sm.q_head;
    volatile unsigned* vtail = &sm.q_tail;
    volatile int* vpending = &sm.pending;
    for (;;) {
      // ---- pop: lane 0 claims the next entry, or sees that nothing is left anywhere
      unsigned long long e = 0ull;
      if (lane == 0) {
        for (;;) {
          if (*vpending == 0) break;
          const unsigned h = *vhead;
          if (h != *vtail) {
            if (atomicCAS(&sm.q_head, h, h + 1) == h) {
              while (!((e = q[h & a.qmask]) & kQValid)) {
              }
              q[h & a.qmask] = 0ull;
              break;
            }
          } else {
            __nanosleep(40);
          }
        }
        __threadfence_block();
      }
      e = __shfl_sync(0xffffffffu, e, 0);
      if (!(e & kQValid)) break;
      int f = int(e & 0xFFFFFu), l = int((e >> 20) & 0xFFFFFu), d = int((e >> 40) & 0xFFu);
      // ---- the loop of std::__introsort_loop on [f, l): partition, hand the right part to the
      // queue, go on with the left part
      while (l - f > kSortThreshold) {
        if (d == 0) {
          if (lane == 0) heap_sort(keys + f, l - f);
          __syncwarp();
          break;
        }
        --d;
        const int cut = warp_partition(keys, Ll, Rl, f, l, lane);
        if (cut < keep) {
          if (l - cut > kSortThreshold && lane == 0) {
            __threadfence_block();                 // the swaps above, before the entry becomes visible
            atomicAdd(&sm.pending, 1);
            const unsigned idx = atomicAdd(&sm.q_tail, 1u);
            q[idx & a.qmask] = q_pack(cut, l, d);
          }
        } else if (lane == 0) {
          atomicMin(&sm.sorted_end, cut);
        }
        l = cut;
      }
      if (lane == 0) {
        __threadfence_block();
        atomicSub(&sm.pending, 1);
      }
    }
    __syncthreads();
    n = sm.sorted_end;   // everything at or beyond it is >= everything before it and stays unsorted
  } else {
    __syncthreads();
  }

  // ---- stable counting sort of keys[0, n) by distance == std::__final_insertion_sort.  Each of
  // the 16 warps owns a contiguous slice and a private histogram; bin starts are the exclusive
  // scan over (bin, warp); the scatter walks each slice in order with __match_any_sync ranks,
  // which keeps the sort stable without atomics.
  const int seg = ((n + kSortWarps - 1) / kSortWarps + 31) & ~31;
  const int beg = min(n, warp * seg), end = min(n, beg + seg);

  for (int i0 = beg; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool ok = i < end;
    const uint32_t d = ok ? key_dist(keys[i]) : (0xFFFF0000u + lane);
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if (ok && (peers & ((1u << lane) - 1u)) == 0) sm.hist[warp][min(d, uint32_t(nbins - 1))] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // per bin: exclusive offsets across warps, bin totals
  for (int b = tid; b < nbins; b += kSortThreads) {
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t t = sm.hist[w][b];
      sm.hist[w][b] = tot;
      tot += t;
    }
    sm.start[b] = tot;
  }
  __syncthreads();
  if (warp == 0) {
    uint32_t run = 0;
    for (int b0 = 0; b0 < nbins; b0 += 32) {
      const int b = b0 + lane;
      const uint32_t c = (b < nbins) ? sm.start[b] : 0u;
      uint32_t incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (b < nbins) sm.start[b] = run + incl - c;
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  __syncthreads();
  for (int i0 = beg; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool ok = i < end;
    const uint32_t key = ok ? keys[i] : 0u;
    const uint32_t d = ok ? min(key_dist(key), uint32_t(nbins - 1)) : (0xFFFF0000u + lane);
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (ok) {
      const uint32_t pos = sm.start[d] + sm.hist[warp][d] + rank;
      if (pos < uint32_t(keep)) {
        const vsf_dmatch dm = m[key & kIdxMask];
        ulonglong2 fm;
        fm.x = uint64_t(uint32_t(dm.queryIdx));   // feature_idx_initial (src/slam_frontend.cc:295)
        fm.y = uint64_t(uint32_t(dm.trainIdx));   // feature_idx_current (:296)
        reinterpret_cast<ulonglong2*>(out)[pos] = fm;
      }
    }
    __syncwarp();
    if (ok && rank == 0) sm.hist[warp][d] += __popc(peers);
    __syncwarp();
  }
  if (tid == 0) a.out_counts[p] = keep;
}

int sort_exact_max_rows() { return kSortMaxRows; }

// max_rows: an upper bound of every list's length (sizes the shared memory).  exact != 0 needs
// max_rows <= sort_exact_max_rows().
cudaError_t launch_sort_cut(const vsf_dmatch* const* matches, const int* const* counts, int n_problems,
                            float best_percent, vsf_feature_match* out, int out_stride, int* out_counts,
                            int bins, int max_rows, int exact, int depth_override, cudaStream_t stream) {
  if (n_problems <= 0) return cudaSuccess;
  if (n_problems > kMaxProblems || max_rows < 0) return cudaErrorInvalidValue;
  if (exact && max_rows > kSortMaxRows) return cudaErrorInvalidValue;
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
  a.exact = exact ? 1 : 0;
  a.cap = (std::max(max_rows, 32) + 31) & ~31;
  int qsize = 64;
  while (qsize < a.cap / kSortThreshold + 2) qsize *= 2;
  a.qmask = qsize - 1;
  a.depth_override = depth_override;
  size_t dyn = size_t(a.cap) * sizeof(uint32_t);
  if (exact) dyn += size_t(a.cap) * 2 * sizeof(uint16_t) + size_t(qsize) * sizeof(unsigned long long);
  static std::atomic<size_t> dyn_set{0};   // per-device attribute, raised monotonically (every device gets the max)
  if (dyn > 48 * 1024 || dyn_set.load() != 0) {
    cudaError_t e = cudaFuncSetAttribute(sort_cut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(std::max(dyn, size_t(48 * 1024))));
    if (e != cudaSuccess) return e;
    dyn_set.store(dyn);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_problems);
  cfg.blockDim = dim3(kSortThreads);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, sort_cut_kernel, a);
}

}  // namespace vsf
