// libstdc++ restatement of the sort in Frontend::GetFeatureMatches.
//
// TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
//
// Follows /root/reference/src/slam_frontend.cc:289:
//     std::sort(matches.begin(), matches.end());
// on std::vector<cv::DMatch>, whose operator< compares `distance` only.
// std::sort is not stable, so the order inside groups of equal distance is
// whatever libstdc++'s introsort produces (quirk Q1 in SURVEY.md).  The
// sequence of comparisons/moves depends only on the comparator results and the
// element count, not on the element size, so sorting {distance, position}
// records with the same comparator yields the reference's permutation.
#include <algorithm>
#include <cstdint>
#include <vector>

namespace {
struct Rec {
  int32_t queryIdx, trainIdx, imgIdx;  // same 16-byte footprint as cv::DMatch
  float distance;
  bool operator<(const Rec& o) const { return distance < o.distance; }
};
}  // namespace

extern "C" void oracle_stdsort_order(const float* distance, int n, int32_t* order) {
  std::vector<Rec> v(n);
  for (int i = 0; i < n; ++i) v[i] = Rec{i, 0, 0, distance[i]};
  std::sort(v.begin(), v.end());
  for (int i = 0; i < n; ++i) order[i] = v[i].queryIdx;
}
