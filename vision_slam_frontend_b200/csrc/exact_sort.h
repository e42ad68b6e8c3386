// Host side of Frontend::GetFeatureMatches' sort + cut (src/slam_frontend.cc:289-291):
// std::sort(matches) by distance only, then keep the first int(size * best_percent).
//
// std::sort is not stable, so which of several equal-distance matches end up before the cut is
// whatever libstdc++'s introsort does, and the drop-in has to do the same.  It does not have to
// do all of it, though.  Introsort only ever rearranges elements inside the range it is working
// on: once a partition step has split [first, last) at `cut`, nothing done later to [cut, last)
// can change what is in [first, cut).  So a right-hand part that starts at or beyond the cut
// position `keep` is partitioned off and then left alone; every range that reaches below `keep`
// gets exactly the operations std::sort applies to it (same pivots, same swaps, same depth
// limit, same final insertion sort - all of them libstdc++'s own functions, called directly).
// The first `keep` elements are then identical, ties included, to those of a full std::sort, at
// about half the work for best_percent = 0.3 (tests/test_exact_sort.py checks it against the
// real std::sort of oracle/stdsort_oracle.cc; measured 201 -> 97 us for 4500 matches).
//
// "Identical to the reference" therefore means: identical to std::sort of the libstdc++ this
// library is BUILT with (GCC 13 in this image; the introsort of bits/stl_algo.h has had the same
// pivot rule, threshold 16 and depth limit 2*lg(n) since GCC 4.x, so any libstdc++ a ROS-era
// reference build used sorts the same way, but that is libstdc++'s promise, not the standard's).
// The functions called below are libstdc++ internals; the guard keeps other standard libraries
// (and libstdc++ releases that predate __ops::__iter_comp_iter, GCC < 5) on plain std::sort.
#pragma once

#include <algorithm>
#include <cstdint>

#if defined(__GLIBCXX__) && defined(_GLIBCXX_RELEASE) && _GLIBCXX_RELEASE >= 7
#define VSF_EXACT_SORT_LIBSTDCXX 1
#elif defined(__GLIBCXX__) && __GLIBCXX__ >= 20150422
#define VSF_EXACT_SORT_LIBSTDCXX 1
#endif

namespace vsf_exact_sort {

// Keys are (distance << SHIFT | position): compared by distance only, like cv::DMatch::operator<.
template <int SHIFT>
struct KeyLess {
  bool operator()(uint32_t a, uint32_t b) const { return (a >> SHIFT) < (b >> SHIFT); }
};

#ifdef VSF_EXACT_SORT_LIBSTDCXX
// std::__introsort_loop, except that a right-hand part which starts at or beyond keep_end is
// not descended into.  *sorted_end = start of the leftmost such part.
template <typename Comp>
inline void introsort_prefix_loop(uint32_t* first, uint32_t* last, long depth_limit, const uint32_t* keep_end,
                                  uint32_t** sorted_end, Comp comp) {
  while (last - first > 16) {                       // std::_S_threshold
    if (depth_limit == 0) {
      std::__partial_sort(first, last, last, comp);
      return;
    }
    --depth_limit;
    uint32_t* cut = std::__unguarded_partition_pivot(first, last, comp);
    if (cut < keep_end) introsort_prefix_loop(cut, last, depth_limit, keep_end, sorted_end, comp);
    else if (cut < *sorted_end) *sorted_end = cut;
    last = cut;
  }
}
#endif

// After the call keys[0 .. keep) hold exactly what std::sort(keys, keys + n, KeyLess<SHIFT>())
// leaves there; the rest of the array is the remaining elements in unspecified order.
template <int SHIFT>
inline void sort_prefix(uint32_t* keys, long n, long keep, long depth_limit = -1) {
  if (n <= 0 || keep <= 0) return;
#ifdef VSF_EXACT_SORT_LIBSTDCXX
  auto comp = __gnu_cxx::__ops::__iter_comp_iter(KeyLess<SHIFT>());
  uint32_t* sorted_end = keys + n;
  introsort_prefix_loop(keys, keys + n, depth_limit >= 0 ? depth_limit : std::__lg(n) * 2, keys + std::min(keep, n),
                        &sorted_end, comp);
  std::__final_insertion_sort(keys, sorted_end, comp);
#else
  std::sort(keys, keys + n, KeyLess<SHIFT>());      // another standard library: its own order
#endif
}

}  // namespace vsf_exact_sort
