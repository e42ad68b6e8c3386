/* Plain-C restatement of the reference's k=2 Hamming kNN + Lowe ratio test.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ as the
 * fast checker for the full-size configs and by bench.py's CPU legs.  The
 * product library never links or loads this file.
 *
 * Follows: /root/reference/src/slam_frontend.cc:521-538 (Frontend::GetMatches)
 * and the published behaviour of cv::BFMatcher(NORM_HAMMING)::knnMatch(k=2)
 * that it calls at :525-527 (OpenCV is not vendored; pinned EXACT 3.2.0 at
 * CMakeLists.txt:21): train rows are scanned in increasing index order and
 * inserted with strict '<', so equal distances resolve to the lowest train
 * index for both neighbours.  Pinned against OpenCV 4.13 outputs in
 * tests/golden/ (tests/test_oracle_golden.py).
 */
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline int hamming_row(const uint8_t *a, const uint8_t *b, int bytes) {
  int d = 0, i = 0;
  for (; i + 8 <= bytes; i += 8) {
    uint64_t x, y;
    memcpy(&x, a + i, 8);
    memcpy(&y, b + i, 8);
    d += __builtin_popcountll(x ^ y);
  }
  for (; i < bytes; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
  return d;
}

/* idx[nq][2], dist[nq][2]; missing neighbours (nt < 2) are -1. */
void oracle_knn2_hamming(const uint8_t *q, int nq, const uint8_t *t, int nt,
                         int bytes, int32_t *idx, int32_t *dist) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nq; ++i) {
    int32_t b0 = -1, b1 = -1, d0 = INT32_MAX, d1 = INT32_MAX;
    const uint8_t *qi = q + (size_t)i * bytes;
    for (int j = 0; j < nt; ++j) {
      int d = hamming_row(qi, t + (size_t)j * bytes, bytes);
      if (d < d0) { d1 = d0; b1 = b0; d0 = d; b0 = j; }
      else if (d < d1) { d1 = d; b1 = j; }
    }
    idx[2 * i] = b0; idx[2 * i + 1] = b1;
    dist[2 * i] = b0 < 0 ? -1 : d0;
    dist[2 * i + 1] = b1 < 0 ? -1 : d1;
  }
}

/* GetMatches: survivors of `dist1 < ratio * dist2` (double compare, :533) in
 * ascending query order.  out = int32 quadruples {queryIdx, trainIdx, imgIdx=0,
 * distance-as-int}.  Returns the number written.  nt < 2 -> 0 (quirk Q6). */
int oracle_get_matches(const uint8_t *q, int nq, const uint8_t *t, int nt,
                       int bytes, double ratio, int32_t *out,
                       int32_t *scratch_idx, int32_t *scratch_dist) {
  if (nt < 2) return 0;
  oracle_knn2_hamming(q, nq, t, nt, bytes, scratch_idx, scratch_dist);
  int n = 0;
  for (int i = 0; i < nq; ++i) {
    float d1 = (float)scratch_dist[2 * i], d2 = (float)scratch_dist[2 * i + 1];
    if (d1 < ratio * d2) {
      out[4 * n] = i; out[4 * n + 1] = scratch_idx[2 * i];
      out[4 * n + 2] = 0; out[4 * n + 3] = scratch_dist[2 * i];
      ++n;
    }
  }
  return n;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
