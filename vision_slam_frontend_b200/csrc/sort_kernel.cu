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
//    BEFORE the step, so every swap of a partition step is independent of the others, and so are
//    the ranges of one recursion level.  The kernel therefore walks the recursion level by level
//    with the whole CTA: one ballot pass flags every element of every live range against its
//    range's pivot, a prefix sum over the ballots gives each flagged element its rank k, the lists
//    are scattered, one thread per range finds k* by bisection (the predicate is monotone) and the
//    cut, one more pass swaps and re-labels the elements with their child range.  Six barriers per
//    level, ~lg(n / 16) + a few levels.  tools/exact_sort_model.py is the numpy model of this
//    formulation; both are checked against the real std::sort (oracle/stdsort_oracle.cc).
#include <algorithm>

#include "vsf_device.cuh"

namespace vsf {

constexpr int kSortThreads = 1024;                 // the partition replay uses all 32 warps
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kCountWarps = 16;                    // the counting sort: 16 slices, 16 private histograms
constexpr int kSortLeaf = 16;                      // (= std::_S_threshold) ranges this short would be finished serially (off: dependent chains in one thread are slower than two more parallel levels)
constexpr int kBins = 513;
constexpr int kSortThreshold = 16;          // std::_S_threshold
constexpr int kSortMaxRows = 24576;         // exact mode: 16-bit positions; the keys must fit in shared memory
constexpr int kSortSmemRows = 16384;        // exact mode: longest list whose partition scratch is kept in shared memory too

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
  int depth_override;   // tests: force introsort's depth limit (-1 = 2 * floor(lg n))
  uint8_t* gscratch;    // partition scratch in global memory (lists too long for shared memory), or null
  size_t gscratch_stride;
  long long* trace;     // bring-up: per CTA 16 values = cycles in P1..P6, leaf phase, load, counting sort, levels, leaves
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
__device__ __noinline__ void heap_sort(uint32_t* v, int len) {
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

constexpr uint16_t kNoSeg = 0xFFFFu;
constexpr uint16_t kDeadPivot = 0xFFFFu;

// The scratch of the partition phase, carved from `base` (shared memory, or global memory for
// lists too long for it).  cap = a multiple of 32 >= n.
struct PartScratch {
  uint16_t *Ll, *Rl;        // [cap] the two position lists, range [f, l) uses entries f ..
  uint16_t* seg_of;         // [cap] live range of every position at the current level
  uint32_t *mL, *mR;        // [cap / 32 + 1] ballots: element not < pivot / not > pivot
  uint32_t *cumL, *cumR;    // [cap / 32 + 1] counts, then exclusive prefix sums of the ballots
  uint32_t* seg_fl[2];      // [max_seg] f | l << 16, current / next level
  uint16_t *seg_piv, *seg_cut, *seg_ks;   // [max_seg]
  uint32_t* seg_child;      // [max_seg] left child | right child << 16 (kNoSeg = none)
};
__host__ __device__ inline int sort_max_seg(int cap) { return cap / (kSortThreshold + 1) + 2; }
__host__ __device__ inline size_t sort_scratch_bytes(int cap) {
  const size_t nch = size_t(cap / 32 + 1), ms = size_t((sort_max_seg(cap) + 1) & ~1);
  return size_t(cap) * 6 + nch * 16 + ms * (8 + 6 + 4) + 64;
}
template <typename BYTE>
__device__ __forceinline__ PartScratch carve_scratch(BYTE* base, int cap) {
  PartScratch S;
  const size_t nch = size_t(cap / 32 + 1), ms = size_t((sort_max_seg(cap) + 1) & ~1);
  uint32_t* w = reinterpret_cast<uint32_t*>(base);
  S.mL = w; w += nch;
  S.mR = w; w += nch;
  S.cumL = w; w += nch;
  S.cumR = w; w += nch;
  S.seg_fl[0] = w; w += ms;
  S.seg_fl[1] = w; w += ms;
  S.seg_child = w; w += ms;
  uint16_t* h = reinterpret_cast<uint16_t*>(w);
  S.Ll = h; h += cap;
  S.Rl = h; h += cap;
  S.seg_of = h; h += cap;
  S.seg_piv = h; h += ms;
  S.seg_cut = h; h += ms;
  S.seg_ks = h;
  return S;
}

// number of flagged positions in [0, x)
__device__ __forceinline__ uint32_t flag_prefix(const uint32_t* cum, const uint32_t* mask, int x) {
  const int c = x >> 5;
  return cum[c] + __popc(mask[c] & ((1u << (x & 31)) - 1u));
}

struct SortShared {
  uint32_t start[kBins + 31];
  int nseg_next[2], hi_next[2];   // per level parity: written by one level, reset during the next
  int sorted_end;
};

// Replay of std::__introsort_loop on keys[0, n), restricted to ranges that reach below `keep`;
// returns (to every thread) the position from which on the array is left unsorted.
#define VSF_SORT_TR(k)                                   \
  do {                                                   \
    if (tr && tid == 0) {                                \
      const long long now__ = clock64();                 \
      acc[k] += now__ - t_last;                          \
      t_last = now__;                                    \
    }                                                    \
  } while (0)

__device__ __forceinline__ int introsort_replay(uint32_t* keys, const PartScratch S, SortShared& sm, int n, int keep,
                                                int depth_limit, long long* tr) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long t_last = clock64();
  long long acc[7] = {0, 0, 0, 0, 0, 0, 0};
  int levels_done = 0;
  const int cap32 = (n + 31) & ~31;
  for (int i = tid; i < cap32; i += kSortThreads) S.seg_of[i] = i < n ? uint16_t(0) : kNoSeg;
  for (int i = tid; i <= cap32 / 32; i += kSortThreads) {
    S.mL[i] = 0u;
    S.mR[i] = 0u;
  }
  if (tid == 0) {
    S.seg_fl[0][0] = uint32_t(n) << 16;
    sm.sorted_end = n;
    sm.nseg_next[0] = sm.nseg_next[1] = 0;
    sm.hi_next[0] = sm.hi_next[1] = 0;
  }
  __syncthreads();
  int nseg = 1, hi = n, cur = 0;
#pragma unroll 1
  for (int level = 0; nseg > 0; ++level) {
    uint32_t* seg_fl = S.seg_fl[cur];
    uint32_t* seg_next = S.seg_fl[cur ^ 1];
    const int nch = (hi + 31) >> 5;
    // ---- P1, one thread per range: depth limit reached -> libstdc++'s heapsort and the range is
    // done; otherwise std::__move_median_to_first(first, first + 1, mid, last - 1)
#pragma unroll 1
    for (int s = tid; s < nseg; s += kSortThreads) {
      const int f = int(seg_fl[s] & 0xFFFFu), l = int(seg_fl[s] >> 16);
      if (level >= depth_limit) {
        heap_sort(keys + f, l - f);
        S.seg_piv[s] = kDeadPivot;
        continue;
      }
      const int ia = f + 1, ib = f + (l - f) / 2, ic = l - 1;
      const uint32_t da = key_dist(keys[ia]), db = key_dist(keys[ib]), dc = key_dist(keys[ic]);
      int pick;
      if (da < db) pick = (db < dc) ? ib : ((da < dc) ? ic : ia);
      else pick = (da < dc) ? ia : ((db < dc) ? ic : ib);
      const uint32_t t = keys[f];
      keys[f] = keys[pick];
      keys[pick] = t;
      S.seg_piv[s] = uint16_t(key_dist(keys[f]));
    }
    if (tid == 0) {            // this level's counters: last read at the end of level - 2, used from P5 on
      sm.nseg_next[level & 1] = 0;
      sm.hi_next[level & 1] = 0;
    }
    __syncthreads();
    VSF_SORT_TR(0);
    // ---- P2, one ballot pass over [0, hi): element of [f + 1, l) not < pivot / not > pivot
#pragma unroll 1
    for (int c = warp; c < nch; c += kSortWarps) {
      const int p = (c << 5) + lane;
      const uint16_t s = S.seg_of[p];
      const uint32_t d = key_dist(keys[p]);
      bool fl = false, fr = false;
      if (s != kNoSeg) {
        const uint16_t piv = S.seg_piv[s];
        if (piv != kDeadPivot && p != int(seg_fl[s] & 0xFFFFu)) {
          fl = d >= piv;
          fr = d <= piv;
        }
      }
      const unsigned bl = __ballot_sync(0xffffffffu, fl), br = __ballot_sync(0xffffffffu, fr);
      if (lane == 0) {
        S.mL[c] = bl;
        S.mR[c] = br;
        S.cumL[c] = __popc(bl);
        S.cumR[c] = __popc(br);
      }
    }
    __syncthreads();
    VSF_SORT_TR(1);
    // ---- P3, warps 0 and 1: exclusive prefix sums of the ballot counts
    if (warp < 2) {
      uint32_t* cnt = warp == 0 ? S.cumL : S.cumR;
      uint32_t run = 0;
#pragma unroll 1
      for (int c0 = 0; c0 < nch; c0 += 32) {
        const int c = c0 + lane;
        const uint32_t v = c < nch ? cnt[c] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += u;
        }
        if (c < nch) cnt[c] = run + incl - v;
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) {
        cnt[nch] = run;
        (warp == 0 ? S.mL : S.mR)[nch] = 0u;
      }
    }
    __syncthreads();
    VSF_SORT_TR(2);
    // ---- P4: scatter the lists.  L[k] = k-th flagged position from the left of its range, R[k]
    // = k-th from the right; both stored from entry f of the range on
#pragma unroll 1
    for (int c = warp; c < nch; c += kSortWarps) {
      const int p = (c << 5) + lane;
      const uint16_t s = S.seg_of[p];
      if (s == kNoSeg) continue;
      const unsigned bl = S.mL[c], br = S.mR[c];
      const bool isl = (bl >> lane) & 1u, isr = (br >> lane) & 1u;
      if (!isl && !isr) continue;
      const int f = int(seg_fl[s] & 0xFFFFu), l = int(seg_fl[s] >> 16);
      if (isl) {
        const uint32_t k = S.cumL[c] + __popc(bl & ((1u << lane) - 1u)) - flag_prefix(S.cumL, S.mL, f + 1);
        S.Ll[f + k] = uint16_t(p);
      }
      if (isr) {
        const uint32_t upto = S.cumR[c] + __popc(br & ((2u << lane) - 1u));      // flagged in [0, p]
        const uint32_t k = flag_prefix(S.cumR, S.mR, l) - upto;
        S.Rl[f + k] = uint16_t(p);
      }
    }
    __syncthreads();
    VSF_SORT_TR(3);
    // ---- P5, one thread per range: k* by bisection, the cut, the child ranges (long ones go to
    // the next level, short ones to the serial finish)
#pragma unroll 1
    for (int s = tid; s < nseg; s += kSortThreads) {
      const int f = int(seg_fl[s] & 0xFFFFu), l = int(seg_fl[s] >> 16);
      if (S.seg_piv[s] == kDeadPivot) {
        S.seg_ks[s] = 0;
        S.seg_cut[s] = uint16_t(l);
        S.seg_child[s] = uint32_t(kNoSeg) | (uint32_t(kNoSeg) << 16);
        continue;
      }
      const int cntL = int(flag_prefix(S.cumL, S.mL, l) - flag_prefix(S.cumL, S.mL, f + 1));
      const int cntR = int(flag_prefix(S.cumR, S.mR, l) - flag_prefix(S.cumR, S.mR, f + 1));
      int lo = 0, up = min(cntL, cntR);     // first k in [0, up] with !(L[k] < R[k]); up = none fails
#pragma unroll 1
      while (lo < up) {
        const int mid = (lo + up) >> 1;
        if (S.Ll[f + mid] < S.Rl[f + mid]) lo = mid + 1;
        else up = mid;
      }
      const int ks = lo;
      const int cl = ks < cntL ? int(S.Ll[f + ks]) : 0x7fffffff;
      const int cr = ks >= 1 ? int(S.Rl[f + ks - 1]) : 0x7fffffff;
      const int cut = min(cl, cr);
      S.seg_ks[s] = uint16_t(ks);
      S.seg_cut[s] = uint16_t(cut);
      int cf[2], cl2[2], nchild = 0;
      if (cut - f > kSortThreshold) {
        cf[nchild] = f;
        cl2[nchild++] = cut;
      }
      if (cut < keep) {
        if (l - cut > kSortThreshold) {
          cf[nchild] = cut;
          cl2[nchild++] = l;
        }
      } else {
        atomicMin(&sm.sorted_end, cut);
      }
      uint32_t child = uint32_t(kNoSeg) | (uint32_t(kNoSeg) << 16);
      for (int k = 0; k < nchild; ++k) {
        const int i = atomicAdd(&sm.nseg_next[level & 1], 1);
        seg_next[i] = uint32_t(cf[k]) | (uint32_t(cl2[k]) << 16);
        atomicMax(&sm.hi_next[level & 1], cl2[k]);
        if (cf[k] == f) child = (child & 0xFFFF0000u) | uint32_t(i);
        else child = (child & 0x0000FFFFu) | (uint32_t(i) << 16);
      }
      S.seg_child[s] = child;
    }
    __syncthreads();
    VSF_SORT_TR(4);
    // ---- P6: all swaps of the level at once, then every element moves to its child range
#pragma unroll 1
    for (int c = warp; c < nch; c += kSortWarps) {
      const int p = (c << 5) + lane;
      const uint16_t s = S.seg_of[p];
      if (s == kNoSeg) continue;
      const unsigned bl = S.mL[c];
      const int f = int(seg_fl[s] & 0xFFFFu);
      if ((bl >> lane) & 1u) {
        const uint32_t k = S.cumL[c] + __popc(bl & ((1u << lane) - 1u)) - flag_prefix(S.cumL, S.mL, f + 1);
        if (k < S.seg_ks[s]) {
          const int partner = S.Rl[f + k];
          const uint32_t t = keys[p];
          keys[p] = keys[partner];
          keys[partner] = t;
        }
      }
      const uint32_t child = S.seg_child[s];
      S.seg_of[p] = p < int(S.seg_cut[s]) ? uint16_t(child & 0xFFFFu) : uint16_t(child >> 16);
    }
    __syncthreads();
    VSF_SORT_TR(5);
    ++levels_done;
    nseg = sm.nseg_next[level & 1];
    hi = sm.hi_next[level & 1];
    cur ^= 1;
  }
  if (tr && tid == 0) {
    for (int k = 0; k < 6; ++k) tr[k] = acc[k];
    tr[9] = levels_done;
  }
  return sm.sorted_end;
}

__global__ void __launch_bounds__(kSortThreads) sort_cut_kernel(const __grid_constant__ SortArgs a) {
  __shared__ SortShared sm;
  extern __shared__ __align__(16) uint8_t s_dyn[];
  // dynamic shared memory: keys [cap], then a region that is the partition scratch first and the
  // 16 per-warp histograms of the counting sort afterwards
  uint32_t* keys = reinterpret_cast<uint32_t*>(s_dyn);
  uint8_t* region = s_dyn + size_t(a.cap) * sizeof(uint32_t);
  uint32_t (*hist)[kBins] = reinterpret_cast<uint32_t (*)[kBins]>(region);
  const int p = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const vsf_dmatch* m = a.matches[p];
  const int nbins = a.bins;
  pdl_wait();
  pdl_launch_dependents();
  long long* tr = a.trace ? a.trace + size_t(p) * 16 : nullptr;
  const long long t_begin = clock64();
  int n = *a.counts[p];
  // `matches.size() * config_.best_percent_` truncated to int (src/slam_frontend.cc:290)
  const int keep = int(__fmul_rn(float(size_t(n)), a.best_percent));
  vsf_feature_match* out = a.out + size_t(p) * a.out_stride;
  if (n > a.cap) {   // cannot happen: the host sizes the shared memory for the longest possible list
    if (tid == 0) a.out_counts[p] = -1;
    return;
  }
  for (int i = tid; i < n; i += kSortThreads) keys[i] = (uint32_t(int(m[i].distance)) << kIdxBits) | uint32_t(i);
  __syncthreads();
  if (tr && tid == 0) {
    for (int k = 0; k < 16; ++k) tr[k] = 0;
    tr[7] = clock64() - t_begin;
  }
  const long long t_sorted = clock64();
  if (a.exact && n > kSortThreshold && keep > 0) {
    const int depth = a.depth_override >= 0 ? a.depth_override : 2 * (31 - __clz(n));
    // two copies of the replay, so that the scratch in shared memory is reached with shared-memory
    // instructions (a pointer that may be either goes through the slower generic path)
    if (a.gscratch) n = introsort_replay(keys, carve_scratch(a.gscratch + size_t(p) * a.gscratch_stride, a.cap), sm, n, keep, depth, tr);
    else n = introsort_replay(keys, carve_scratch(region, a.cap), sm, n, keep, depth, tr);
    __syncthreads();   // (everything from n on is >= everything before it)
  }
  if (warp < kCountWarps)
    for (int b = lane; b < nbins; b += 32) hist[warp][b] = 0;   // every counting warp clears its own histogram
  __syncthreads();

  // ---- stable counting sort of keys[0, n) by distance == std::__final_insertion_sort.  Each of
  // the 16 warps owns a contiguous slice and a private histogram; bin starts are the exclusive
  // scan over (bin, warp); the scatter walks each slice in order with __match_any_sync ranks,
  // which keeps the sort stable without atomics.
  const int seg = ((n + kCountWarps - 1) / kCountWarps + 31) & ~31;
  const int beg = warp < kCountWarps ? min(n, warp * seg) : n, end = min(n, beg + seg);   // warps 16.. have nothing

  for (int i0 = beg; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool ok = i < end;
    const uint32_t d = ok ? key_dist(keys[i]) : (0xFFFF0000u + lane);
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if (ok && (peers & ((1u << lane) - 1u)) == 0) hist[warp][min(d, uint32_t(nbins - 1))] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // per bin: exclusive offsets across warps, bin totals
  for (int b = tid; b < nbins; b += kSortThreads) {
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < kCountWarps; ++w) {
      const uint32_t t = hist[w][b];
      hist[w][b] = tot;
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
      const uint32_t pos = sm.start[d] + hist[warp][d] + rank;
      if (pos < uint32_t(keep)) {
        const vsf_dmatch dm = m[key & kIdxMask];
        ulonglong2 fm;
        fm.x = uint64_t(uint32_t(dm.queryIdx));   // feature_idx_initial (src/slam_frontend.cc:295)
        fm.y = uint64_t(uint32_t(dm.trainIdx));   // feature_idx_current (:296)
        reinterpret_cast<ulonglong2*>(out)[pos] = fm;
      }
    }
    __syncwarp();
    if (ok && rank == 0) hist[warp][d] += __popc(peers);
    __syncwarp();
  }
  if (tid == 0) a.out_counts[p] = keep;
  if (tr && tid == 0) {
    tr[8] = clock64() - t_sorted;      // replay + counting sort
    tr[11] = n;
  }
}

int sort_exact_max_rows() { return kSortMaxRows; }

// Bytes of global scratch one list needs when exact mode cannot keep the partition scratch in
// shared memory (0 when it can); the caller provides n_problems * that many bytes.
size_t sort_exact_global_scratch(int max_rows) {
  const int cap = (std::max(max_rows, 32) + 31) & ~31;
  return cap > kSortSmemRows ? ((sort_scratch_bytes(cap) + 255) & ~size_t(255)) : 0;
}

// max_rows: an upper bound of every list's length (sizes the shared memory).  exact != 0 needs
// max_rows <= sort_exact_max_rows(), and gscratch when sort_exact_global_scratch(max_rows) != 0.
cudaError_t launch_sort_cut(const vsf_dmatch* const* matches, const int* const* counts, int n_problems,
                            float best_percent, vsf_feature_match* out, int out_stride, int* out_counts,
                            int bins, int max_rows, int exact, int depth_override, void* gscratch,
                            cudaStream_t stream, long long* trace) {
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
  a.depth_override = depth_override;
  a.gscratch = nullptr;
  a.gscratch_stride = 0;
  a.trace = trace;
  size_t region = size_t(kCountWarps) * kBins * sizeof(uint32_t);   // the histograms of the counting sort
  if (exact) {
    const size_t g = sort_exact_global_scratch(max_rows);
    if (g) {
      if (!gscratch) return cudaErrorInvalidValue;
      a.gscratch = static_cast<uint8_t*>(gscratch);
      a.gscratch_stride = g;
    } else {
      region = std::max(region, sort_scratch_bytes(a.cap));
    }
  }
  const size_t dyn = size_t(a.cap) * sizeof(uint32_t) + region;
  // more than the 48 KB a kernel gets without opting in for all but the shortest lists, so the
  // opt-in is simply made on every launch (a host-side call, per device, well under a microsecond)
  const cudaError_t ea = cudaFuncSetAttribute(sort_cut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dyn));
  if (ea != cudaSuccess) return ea;
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
