// vsf_latency_probe — what a C++ caller of the C ABI sees, one frame at a time (the reference's
// real mode of use: one ObserveImage per camera frame, src/slam_frontend_main.cc:236-328).
// BASELINE configs C2 (2000 x 2000 single pair) and C3 (stereo L/R match + filter + R'->L' match
// + triangulation, 2000 features, KITTI-size 1241x376 intrinsics).  Prints one JSON object.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "synthetic_source.h"

namespace {
double Now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Stats {
  double best, median;
};
Stats Summarise(std::vector<double> v) {
  std::sort(v.begin(), v.end());
  return Stats{v.front() * 1e6, v[v.size() / 2] * 1e6};
}
}  // namespace

int main(int argc, char** argv) {
  const int device = argc > 1 ? std::atoi(argv[1]) : 0;
  const int n = argc > 2 ? std::atoi(argv[2]) : 2000;
  const int width = argc > 3 ? std::atoi(argv[3]) : 32;
  const int W = argc > 4 ? std::atoi(argv[4]) : 1;
  // KITTI-odometry-style rectified rig
  slam::FrontendConfig rig = slam::SyntheticRig(device, n, width, W, true);
  slam::CameraIntrinsics I = {0.f, 0.f, 0.f, 0.f, 0.f, 718.856f, 718.856f, 607.1928f, 185.2157f};
  rig.intrinsics_left = I;
  rig.intrinsics_right = I;
  const float A[12] = {1, 0, 0, -0.5371657f, 0, 1, 0, 0, 0, 0, 1, 0};
  std::memcpy(rig.stereo_extrinsics, A, sizeof(A));
  rig.UpdateDerived();
  {   // residuals in pixels, like SyntheticRig
    float m = 0.f;
    for (float v : rig.fundamental) m = std::max(m, std::fabs(v));
    for (float& v : rig.fundamental) v /= m;
  }
  slam::SyntheticStereoConfig sc;
  sc.features = n;
  sc.image_width = 1241;
  sc.image_height = 376;
  sc.landmark_stride = std::max(1, n / 10);
  slam::SyntheticStereoSource src(rig, sc);
  const int kFrames = 8;
  std::vector<std::vector<cv::KeyPoint>> lk(kFrames), rk(kFrames);
  std::vector<cv::Mat> ld(kFrames), rd(kFrames);
  for (int p = 0; p < kFrames; ++p) src.Frame(uint64_t(p), &lk[p], &ld[p], &rk[p], &rd[p]);

  vsf_ctx* ctx = nullptr;
  if (vsf_create(device, n, width, W, &ctx) != VSF_OK) {
    std::fprintf(stderr, "vsf_create failed (no CUDA device?)\n");
    return 1;
  }
  // ---- C2: GetMatches on one pair
  std::vector<vsf_dmatch> gm(static_cast<size_t>(n), vsf_dmatch{});
  int ng = 0;
  std::vector<double> t2;
  for (int it = 0; it < 220; ++it) {
    const double t0 = Now();
    if (vsf_get_matches(ctx, ld[0].data, n, ld[0].step, ld[1].data, n, ld[1].step, 0.6000000238418579, gm.data(), n, &ng)) return 2;
    if (it >= 20) t2.push_back(Now() - t0);
  }
  const Stats c2 = Summarise(t2);
  const int c2_engine = vsf_last_engine(ctx);

  // ---- C3: the whole frame
  const int cap = n;
  std::vector<int32_t> kl(cap), kr(cap);
  std::vector<uint64_t> fids(W);
  std::vector<int> wc(W);
  std::vector<vsf_dmatch> wm(size_t(W) * cap), tm(cap);
  std::vector<float> X4(size_t(cap) * 4), xyu(size_t(cap) * 2);
  vsf_observe_out out;
  std::memset(&out, 0, sizeof(out));
  out.kept_left = kl.data();
  out.kept_right = kr.data();
  out.frame_ids = fids.data();
  out.window_counts = wc.data();
  out.window_matches = wm.data();
  out.tri_matches = tm.data();
  out.tri_X4 = X4.data();
  out.xy_undist = xyu.data();
  out.cap = cap;
  vsf_observe_params prm;
  prm.fundamental = rig.fundamental;
  prm.P_left = rig.projection_left;
  prm.P_right = rig.projection_right;
  prm.K_left = rig.camera_matrix_left;
  prm.dist_left = rig.distortion_coeffs_left;
  prm.nn_match_ratio = 0.6000000238418579;
  auto submit = [&](int p, uint64_t id) {
    return vsf_observe_submit(ctx, id, reinterpret_cast<const vsf_keypoint*>(lk[p].data()), ld[p].data, n, ld[p].step,
                              reinterpret_cast<const vsf_keypoint*>(rk[p].data()), rd[p].data, n, rd[p].step, &prm);
  };
  uint64_t id = 0;
  std::vector<double> tb, ts, tc;
  for (int it = 0; it < 320; ++it) {          // one frame at a time: submit, then collect
    const int p = it % kFrames;
    const double t0 = Now();
    if (submit(p, id++)) return 3;
    const double t1 = Now();
    if (vsf_observe_collect(ctx, nullptr, &out)) return 4;
    const double t2e = Now();
    if (it >= 20) {
      tb.push_back(t2e - t0);
      ts.push_back(t1 - t0);
      tc.push_back(t2e - t1);
    }
  }
  const Stats c3 = Summarise(tb), c3s = Summarise(ts), c3c = Summarise(tc);
  const int kept = out.n_kept, tri = out.n_tri, nfr = out.n_frames;
  // pipelined: `depth` frames in flight
  double pipelined[VSF_OBSERVE_DEPTH + 1] = {0};
  for (int depth = 2; depth <= VSF_OBSERVE_DEPTH; ++depth) {
    for (int rep = 0; rep < 2; ++rep) {
      const int frames = 400;
      const double t0 = Now();
      for (int it = 0; it < frames; ++it) {
        if (submit(it % kFrames, id++)) return 5;
        if (vsf_observe_in_flight(ctx) >= depth && vsf_observe_collect(ctx, nullptr, &out)) return 6;
      }
      while (vsf_observe_in_flight(ctx) > 0)
        if (vsf_observe_collect(ctx, nullptr, &out)) return 7;
      pipelined[depth] = (Now() - t0) / frames * 1e6;
    }
  }
  std::printf("{\"features\": %d, \"descriptor_bytes\": %d, \"window\": %d, "
              "\"C2_get_matches_us\": {\"best\": %.2f, \"median\": %.2f, \"engine\": %d, \"survivors\": %d}, "
              "\"C3_frame_us\": {\"best\": %.2f, \"median\": %.2f, \"submit_median\": %.2f, \"collect_median\": %.2f, "
              "\"kept\": %d, \"triangulated\": %d, \"window_frames\": %d}, "
              "\"C3_pipelined_us_per_frame\": {\"2\": %.2f, \"3\": %.2f, \"4\": %.2f}}\n",
              n, width, W, c2.best, c2.median, c2_engine, ng, c3.best, c3.median, c3s.median, c3c.median, kept, tri, nfr,
              pipelined[2], pipelined[3], pipelined[4]);
  vsf_destroy(ctx);
  return 0;
}
