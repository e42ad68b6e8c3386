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
  uint16_t* rk;             // [cap] rank (from the left, inside its range) of every L element
  uint16_t* Rl;             // [cap] the R list: range [f, l) stores R[k] at entry f + k
  uint16_t* seg_of;         // [cap] live range of every position
  uint32_t *mL, *mR, *mS;   // [cap / 32 + 1] ballots: not < pivot / not > pivot / L element that is swapped
  uint32_t *locL, *locR;    // [cap / 32 + 1] L / R elements in the chunks of the same warp before this one
  uint32_t *wtotL, *wtotR;  // [32] L / R elements of every warp's chunks
  // by level parity (separate members, selected with ?: - an array indexed at run time would push
  // the whole struct into local memory and put a local load in front of every access)
  uint32_t *seg_fl0, *seg_fl1;      // [max_seg] f | l << 16
  int *seg_cut0, *seg_cut1;         // [max_seg] the cut (atomicMin over candidate positions)
  uint32_t *seg_child0, *seg_child1;   // [max_seg] left child | right child << 16 (kNoSeg = none)
  uint16_t* seg_piv;        // [max_seg]
};
__host__ __device__ inline int sort_max_seg(int cap) { return cap / (kSortThreshold + 1) + 2; }
__host__ __device__ inline size_t sort_scratch_bytes(int cap) {
  const size_t nch = size_t(cap / 32 + 1), ms = size_t((sort_max_seg(cap) + 1) & ~1);
  return size_t(cap) * 6 + nch * 20 + 64 * 4 + ms * (6 * 4 + 2) + 64;
}
template <typename BYTE>
__device__ __forceinline__ PartScratch carve_scratch(BYTE* base, int cap) {
  PartScratch S;
  const size_t nch = size_t(cap / 32 + 1), ms = size_t((sort_max_seg(cap) + 1) & ~1);
  uint32_t* w = reinterpret_cast<uint32_t*>(base);
  S.mL = w; w += nch;
  S.mR = w; w += nch;
  S.mS = w; w += nch;
  S.locL = w; w += nch;
  S.locR = w; w += nch;
  S.wtotL = w; w += 32;
  S.wtotR = w; w += 32;
  S.seg_fl0 = w; w += ms;
  S.seg_fl1 = w; w += ms;
  S.seg_cut0 = reinterpret_cast<int*>(w); w += ms;
  S.seg_cut1 = reinterpret_cast<int*>(w); w += ms;
  S.seg_child0 = w; w += ms;
  S.seg_child1 = w; w += ms;
  uint16_t* h = reinterpret_cast<uint16_t*>(w);
  S.rk = h; h += cap;
  S.Rl = h; h += cap;
  S.seg_of = h; h += cap;
  S.seg_piv = h;
  return S;
}

#define VSF_SORT_TR(k)                                   \
  do {                                                   \
    if (tr && tid == 0) {                                \
      const long long now__ = clock64();                 \
      acc[k] += now__ - t_last;                          \
      t_last = now__;                                    \
    }                                                    \
  } while (0)

struct SortShared {
  uint32_t start[kBins + 31];
  int nseg_next[2], hi_next[2];   // per level parity: written by one level, reset during the next
  int sorted_end;
};

// Replay of std::__introsort_loop on keys[0, n), restricted to ranges that reach below `keep`;
// returns (to every thread) the position from which on the array is left unsorted.
//
// One recursion level = four barrier-separated phases over all live ranges at once:
//   P1  one thread per range: median-of-3 to the front (or heapsort at the depth limit)
//   P2  every element: move to the child range the previous level assigned, compare with the
//       range's pivot; ballots, and per warp a running count over its contiguous chunks
//   P4  every element: its rank from its own side (k) and the number of candidates on the other
//       side beyond it, from the ballots' prefix sums (the 32 per-warp totals are scanned by
//       every warp with shuffles - no scan phase): an L element is swapped iff at least k + 1 R
//       elements lie to its right, an R element iff at least k + 1 L elements lie to its left;
//       the R list is scattered by rank; the cut = the smallest position among the L elements
//       that are not swapped and the R elements that are (atomicMin per range)
//   P5  swapped L elements exchange with R[k]; one thread per range turns the cut into child ranges
__device__ __forceinline__ int introsort_replay(uint32_t* keys, const PartScratch S, SortShared& sm, int n, int keep,
                                                int depth_limit, long long* tr) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned lt = (1u << lane) - 1u, le = (2u << lane) - 1u;
  long long t_last = clock64();
  long long acc[4] = {0, 0, 0, 0};
  int levels_done = 0;
  const int cap32 = (n + 31) & ~31;
  for (int i = tid; i < cap32; i += kSortThreads) S.seg_of[i] = i < n ? uint16_t(0) : kNoSeg;
  if (tid == 0) {
    S.seg_fl0[0] = uint32_t(n) << 16;
    sm.sorted_end = n;
    sm.nseg_next[0] = sm.nseg_next[1] = 0;
    sm.hi_next[0] = sm.hi_next[1] = 0;
  }
  __syncthreads();
  int nseg = 1, hi = n;
#pragma unroll 1
  for (int level = 0; nseg > 0; ++level) {
    const int cur = level & 1;
    const uint32_t* seg_fl = cur ? S.seg_fl1 : S.seg_fl0;
    uint32_t* seg_next = cur ? S.seg_fl0 : S.seg_fl1;
    int* seg_cut = cur ? S.seg_cut1 : S.seg_cut0;
    const int* cut_prev = cur ? S.seg_cut0 : S.seg_cut1;
    uint32_t* child_cur = cur ? S.seg_child1 : S.seg_child0;
    const uint32_t* child_prev = cur ? S.seg_child0 : S.seg_child1;
    const int nch = (hi + 31) >> 5;
    const int B = (nch + kSortWarps - 1) / kSortWarps;          // chunks per warp, contiguous
    // c / B without a division: exact for c < 2^16 (B = 1 would need 2^32, so it is special-cased)
    const uint32_t Binv = B > 1 ? 0xFFFFFFFFu / uint32_t(B) + 1u : 0u;   // ceil(2^32 / B)
    auto chunk_warp = [&](int c) { return B > 1 ? int(__umulhi(uint32_t(c), Binv)) : c; };
    const int c_begin = min(nch, warp * B), c_end = min(nch, c_begin + B);
    // ---- P1
#pragma unroll 1
    for (int s = tid; s < nseg; s += kSortThreads) {
      const int f = int(seg_fl[s] & 0xFFFFu), l = int(seg_fl[s] >> 16);
      if (level >= depth_limit) {
        heap_sort(keys + f, l - f);
        S.seg_piv[s] = kDeadPivot;
        seg_cut[s] = l;
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
      seg_cut[s] = 0x7fffffff;
    }
    if (tid == 0) {            // this level's counters: last read at the end of level - 2, used in P5
      sm.nseg_next[cur] = 0;
      sm.hi_next[cur] = 0;
    }
    __syncthreads();
    VSF_SORT_TR(0);
    // ---- P2
    {
      uint32_t runL = 0, runR = 0;
#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        const int p = (c << 5) + lane;
        uint16_t s = S.seg_of[p];
        const uint32_t d = key_dist(keys[p]);
        if (level > 0 && s != kNoSeg) {          // the child range the previous level put this element in
          const uint32_t child = child_prev[s];
          s = p < cut_prev[s] ? uint16_t(child & 0xFFFFu) : uint16_t(child >> 16);
          S.seg_of[p] = s;
        }
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
          S.locL[c] = runL;
          S.locR[c] = runR;
        }
        runL += __popc(bl);
        runR += __popc(br);
      }
      if (lane == 0) {
        S.wtotL[warp] = runL;
        S.wtotR[warp] = runR;
      }
    }
    __syncthreads();
    VSF_SORT_TR(1);
    // ---- P4
    {
      // exclusive prefix of the per-warp totals: lane j holds the count of all warps before warp j
      uint32_t baseL = S.wtotL[lane], baseR = S.wtotR[lane];
      {
        const uint32_t vL = baseL, vR = baseR;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t uL = __shfl_up_sync(0xffffffffu, baseL, o), uR = __shfl_up_sync(0xffffffffu, baseR, o);
          if (lane >= o) {
            baseL += uL;
            baseR += uR;
          }
        }
        baseL -= vL;
        baseR -= vR;
      }
      const uint32_t ownL = __shfl_sync(0xffffffffu, baseL, warp), ownR = __shfl_sync(0xffffffffu, baseR, warp);
#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        const int p = (c << 5) + lane;
        const uint16_t s = S.seg_of[p];
        const unsigned bl = S.mL[c], br = S.mR[c];
        const bool live = s != kNoSeg;
        const uint32_t fl_ = live ? seg_fl[s] : (1u << 16);
        const int f = int(fl_ & 0xFFFFu), l = int(fl_ >> 16);
        // L elements in [0, f] (position f itself is never flagged) and R elements in [0, l - 1]
        const int yf = f, yl = l - 1;
        const int cf = yf >> 5, cl = yl >> 5;
        const uint32_t preL_f = __shfl_sync(0xffffffffu, baseL, chunk_warp(cf)) + S.locL[cf] +
                                __popc(S.mL[cf] & ((2u << (yf & 31)) - 1u));
        const uint32_t preR_l = __shfl_sync(0xffffffffu, baseR, chunk_warp(cl)) + S.locR[cl] +
                                __popc(S.mR[cl] & ((2u << (yl & 31)) - 1u));
        const uint32_t myL_before = ownL + S.locL[c] + __popc(bl & lt);   // L elements in [0, p)
        const uint32_t myR_upto = ownR + S.locR[c] + __popc(br & le);     // R elements in [0, p]
        const bool isL = live && ((bl >> lane) & 1u), isR = live && ((br >> lane) & 1u);
        const uint32_t k = myL_before - preL_f;       // L elements of the range left of p
        const uint32_t r_after = preR_l - myR_upto;   // R elements of the range right of p
        const bool swapL = isL && r_after >= k + 1u;
        const bool swapR = isR && k >= r_after + 1u;
        if (isL) S.rk[p] = uint16_t(k);
        if (isR) S.Rl[f + r_after] = uint16_t(p);
        const unsigned bs = __ballot_sync(0xffffffffu, swapL);
        if (lane == 0) S.mS[c] = bs;
        // the cut: lowest candidate position per range; one atomic per (chunk, range).  A range is
        // a run of consecutive lanes: the run starts where the previous lane's range differs
        const bool cand = (isL && !swapL) || swapR;
        const unsigned prev_s = __shfl_up_sync(0xffffffffu, unsigned(s), 1);
        const unsigned starts = __ballot_sync(0xffffffffu, lane == 0 || prev_s != unsigned(s));
        const unsigned cands = __ballot_sync(0xffffffffu, cand);
        const int run0 = 31 - __clz(starts & le);                      // first lane of this lane's run
        if (cand && (cands & lt & ~((1u << run0) - 1u)) == 0u) atomicMin(&seg_cut[s], p);
      }
    }
    __syncthreads();
    VSF_SORT_TR(2);
    // ---- P5: swaps (elements) and child ranges (one thread per range)
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
      if ((S.mS[c] >> lane) & 1u) {
        const int p = (c << 5) + lane;
        const int f = int(seg_fl[S.seg_of[p]] & 0xFFFFu);
        const int partner = S.Rl[f + S.rk[p]];
        const uint32_t t = keys[p];
        keys[p] = keys[partner];
        keys[partner] = t;
      }
    }
#pragma unroll 1
    for (int s = tid; s < nseg; s += kSortThreads) {
      const int f = int(seg_fl[s] & 0xFFFFu), l = int(seg_fl[s] >> 16);
      uint32_t child = uint32_t(kNoSeg) | (uint32_t(kNoSeg) << 16);
      if (S.seg_piv[s] != kDeadPivot) {
        const int cut = seg_cut[s];
        int cf2[2], cl2[2], nchild = 0;
        if (cut - f > kSortThreshold) {
          cf2[nchild] = f;
          cl2[nchild++] = cut;
        }
        if (cut < keep) {
          if (l - cut > kSortThreshold) {
            cf2[nchild] = cut;
            cl2[nchild++] = l;
          }
        } else {
          atomicMin(&sm.sorted_end, cut);
        }
        for (int k = 0; k < nchild; ++k) {
          const int i = atomicAdd(&sm.nseg_next[cur], 1);
          seg_next[i] = uint32_t(cf2[k]) | (uint32_t(cl2[k]) << 16);
          atomicMax(&sm.hi_next[cur], cl2[k]);
          if (cf2[k] == f) child = (child & 0xFFFF0000u) | uint32_t(i);
          else child = (child & 0x0000FFFFu) | (uint32_t(i) << 16);
        }
      }
      child_cur[s] = child;
    }
    __syncthreads();
    VSF_SORT_TR(3);
    ++levels_done;
    nseg = sm.nseg_next[cur];
    hi = sm.hi_next[cur];
  }
  if (tr && tid == 0) {
    for (int k = 0; k < 4; ++k) tr[k] = acc[k];
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
