#include "synthetic_source.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace slam {

namespace {

inline uint64_t Mix64(uint64_t x) {   // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// counter-based stream: value k of stream (seed, a, b, c)
struct Stream {
  uint64_t key, k = 0;
  Stream(uint64_t seed, uint64_t a, uint64_t b, uint64_t c)
      : key(Mix64(seed ^ Mix64(a * 0x100000001B3ull + Mix64(b * 0x9E3779B1ull + c)))) {}
  uint64_t Next() { return Mix64(key + (k++) * 0xD1B54A32D192ED03ull); }
  double Uniform() { return double(Next() >> 11) * (1.0 / 9007199254740992.0); }   // [0, 1)
  double Normal() {
    const double u1 = 1.0 - Uniform(), u2 = Uniform();
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
  uint32_t Below(uint32_t n) { return uint32_t(Next() % n); }
};

uint32_t Gcd(uint32_t a, uint32_t b) {
  while (b) {
    const uint32_t t = a % b;
    a = b;
    b = t;
  }
  return a;
}

// affine permutation of [0, n) chosen by h
struct Perm {
  uint32_t a, b, n;
  Perm(uint64_t h, uint32_t n_) : n(n_) {
    a = (uint32_t(h % n) | 1u);
    while (Gcd(a, n) != 1u) a += 2u;
    b = uint32_t((h >> 32) % n);
  }
  uint32_t operator()(uint32_t i) const { return uint32_t((uint64_t(a) * i + b) % n); }
};

void Flip(uint8_t* row, int bytes, int max_flips, Stream* s) {
  const int k = max_flips > 0 ? int(s->Below(uint32_t(max_flips) + 1)) : 0;
  for (int f = 0; f < k; ++f) {
    const uint32_t bit = s->Below(uint32_t(bytes) * 8);
    row[bit >> 3] ^= uint8_t(1u << (bit & 7));
  }
}

void Project(const float* P, const float* X, float* u, float* v) {
  const float x = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3];
  const float y = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7];
  const float w = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11];
  *u = x / w;
  *v = y / w;
}

template <typename T>
void Put(std::vector<uint8_t>* b, T v) {
  uint8_t raw[sizeof(T)];
  std::memcpy(raw, &v, sizeof(T));
  b->insert(b->end(), raw, raw + sizeof(T));
}

}  // namespace

SyntheticStereoSource::SyntheticStereoSource(const FrontendConfig& rig, const SyntheticStereoConfig& cfg)
    : rig_(rig), cfg_(cfg) {
  if (cfg_.features < 1 || rig_.descriptor_bytes < 1 || rig_.descriptor_bytes > 64)
    throw std::runtime_error("SyntheticStereoSource: bad configuration");
}

void SyntheticStereoSource::Odometry(uint64_t pose, Eigen::Vector3f* translation, Eigen::Quaternionf* rotation,
                                     double* timestamp) const {
  // a straight drive: every pose clears the min_odom_translation gate (src/slam_frontend.cc:175-186)
  *translation = Eigen::Vector3f(cfg_.step * float(pose + 1), 0.f, 0.f);
  *rotation = Eigen::Quaternionf(1.f, 0.f, 0.f, 0.f);
  *timestamp = 1000.0 + 0.1 * double(pose);
}

void SyntheticStereoSource::Frame(uint64_t pose, std::vector<cv::KeyPoint>* lk, cv::Mat* ld, std::vector<cv::KeyPoint>* rk,
                                  cv::Mat* rd) const {
  const int n = cfg_.features, B = rig_.descriptor_bytes;
  lk->assign(size_t(n), cv::KeyPoint());
  rk->assign(size_t(n), cv::KeyPoint());
  *ld = cv::Mat(n, B, CV_8U);
  *rd = cv::Mat(n, B, CV_8U);
  const uint64_t hp = Mix64(cfg_.seed ^ (0xA5A5A5A5ull + pose * 0x100000001B3ull));
  const Perm landmark_of(hp, uint32_t(n));               // which landmark feature i observes
  const Perm right_slot(Mix64(hp ^ 0x5bd1e995ull), uint32_t(n));   // where its right observation is stored
  const float fx = rig_.camera_matrix_left[0], fy = rig_.camera_matrix_left[4];
  const float cx = rig_.camera_matrix_left[2], cy = rig_.camera_matrix_left[5];
  for (int i = 0; i < n; ++i) {
    const uint64_t L = uint64_t(cfg_.landmark_stride) * pose + landmark_of(uint32_t(i));
    // the landmark's code, then this pose's observation of it
    uint8_t* dl = ld->ptr(i);
    Stream code(cfg_.seed, 1, L, 0);
    for (int b = 0; b < B; b += 8) {
      const uint64_t w = code.Next();
      std::memcpy(dl + b, &w, size_t(std::min(8, B - b)));
    }
    Stream obs(cfg_.seed, 2, pose, uint64_t(i));
    Flip(dl, B, cfg_.temporal_flips, &obs);
    const int j = int(right_slot(uint32_t(i)));
    uint8_t* dr = rd->ptr(j);
    std::memcpy(dr, dl, size_t(B));
    Flip(dr, B, cfg_.stereo_flips, &obs);
    // geometry: a pixel in the left image at a random depth, seen by both cameras
    const float u = float(obs.Uniform()) * float(cfg_.image_width), v = float(obs.Uniform()) * float(cfg_.image_height);
    const float Z = cfg_.depth_min + float(obs.Uniform()) * (cfg_.depth_max - cfg_.depth_min);
    const float X[3] = {(u - cx) * Z / fx, (v - cy) * Z / fy, Z};
    float ul, vl, ur, vr;
    Project(rig_.projection_left, X, &ul, &vl);
    Project(rig_.projection_right, X, &ur, &vr);
    ul += cfg_.pixel_noise * float(obs.Normal());
    vl += cfg_.pixel_noise * float(obs.Normal());
    ur += cfg_.pixel_noise * float(obs.Normal());
    vr += cfg_.pixel_noise * float(obs.Normal());
    const double kind = obs.Uniform();
    if (kind < cfg_.outlier_fraction) {                  // unrelated right feature
      for (int b = 0; b < B; b += 8) {
        const uint64_t w = obs.Next();
        std::memcpy(dr + b, &w, size_t(std::min(8, B - b)));
      }
      ur = float(obs.Uniform()) * float(cfg_.image_width);
      vr = float(obs.Uniform()) * float(cfg_.image_height);
    } else if (kind < cfg_.outlier_fraction + cfg_.bad_geometry_fraction) {
      vr += (obs.Next() & 1) ? 30.f : -30.f;             // passes the ratio test, fails the epipolar one
    }
    (*lk)[size_t(i)] = cv::KeyPoint(ul, vl, 31.f);
    (*rk)[size_t(j)] = cv::KeyPoint(ur, vr, 31.f);
  }
}

FrontendConfig SyntheticRig(int device, int features, int descriptor_bytes, int frame_life, bool exact_std_sort) {
  FrontendConfig cfg;
  cfg.cuda_device = device;
  cfg.max_features = features;
  cfg.descriptor_bytes = descriptor_bytes;
  cfg.frame_life_ = uint32_t(frame_life);
  cfg.exact_std_sort = exact_std_sort;
  // scale F so that x_l^T F x_r is the distance (pixels) of x_r from the epipolar line of the
  // left image centre: the line is F^T x_l = (a, b, c), the distance |a u + b v + c| / |(a, b)|
  const float* F = cfg.fundamental;
  const float cx = cfg.camera_matrix_left[2], cy = cfg.camera_matrix_left[5];
  const float a = F[0] * cx + F[3] * cy + F[6], b = F[1] * cx + F[4] * cy + F[7];
  const float m = std::sqrt(a * a + b * b);
  if (m > 0.f)
    for (float& v : cfg.fundamental) v /= m;
  return cfg;
}

// ------------------------------------------------------------------------- sharded sequences

std::vector<uint8_t> SLAMProblemPiece::Pack() const {
  std::vector<uint8_t> b;
  Put<uint32_t>(&b, n_nodes);
  Put<uint32_t>(&b, n_vision_factors);
  Put<uint32_t>(&b, n_odometry_factors);
  Put<uint64_t>(&b, nodes.size());
  Put<uint64_t>(&b, vision_factors.size());
  Put<uint64_t>(&b, odometry_factors.size());
  b.insert(b.end(), nodes.begin(), nodes.end());
  b.insert(b.end(), vision_factors.begin(), vision_factors.end());
  b.insert(b.end(), odometry_factors.begin(), odometry_factors.end());
  return b;
}

SLAMProblemPiece SLAMProblemPiece::Unpack(const uint8_t* d, size_t n) {
  SLAMProblemPiece p;
  uint64_t sz[3];
  if (n < 36) throw std::runtime_error("SLAMProblemPiece::Unpack: short blob");
  std::memcpy(&p.n_nodes, d, 4);
  std::memcpy(&p.n_vision_factors, d + 4, 4);
  std::memcpy(&p.n_odometry_factors, d + 8, 4);
  std::memcpy(sz, d + 12, 24);
  if (36 + sz[0] + sz[1] + sz[2] != n) throw std::runtime_error("SLAMProblemPiece::Unpack: size mismatch");
  const uint8_t* q = d + 36;
  p.nodes.assign(q, q + sz[0]);
  p.vision_factors.assign(q + sz[0], q + sz[0] + sz[1]);
  p.odometry_factors.assign(q + sz[0] + sz[1], q + sz[0] + sz[1] + sz[2]);
  return p;
}

SLAMProblemPiece RunSequenceShard(const FrontendConfig& rig, const SyntheticStereoSource& source, uint64_t first,
                                  uint64_t last, int in_flight, double* host_us) {
  Frontend fe(rig);
  const uint64_t halo_first = Frontend::ShardHaloStart(first, rig.frame_life_);
  fe.StartShard(halo_first, first);
  Eigen::Vector3f t;
  Eigen::Quaternionf q;
  double ts;
  source.Odometry(0, &t, &q, &ts);            // init_odom_* are relative to the sequence's first message
  fe.ObserveOdometry(t, q, ts);
  std::vector<cv::KeyPoint> lk, rk;
  cv::Mat ld, rd;
  const int depth = std::max(1, std::min(in_flight, Frontend::MaxInFlight()));
  double t_gen = 0, t_submit = 0, t_collect = 0;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (uint64_t p = halo_first; p < last; ++p) {
    source.Odometry(p, &t, &q, &ts);
    fe.ObserveOdometry(t, q, ts);
    const double t0 = now();
    source.Frame(p, &lk, &ld, &rk, &rd);
    const double t1 = now();
    const bool ok = fe.SubmitFeatures(lk, ld, rk, rd, ts);
    const double t2 = now();
    if (ok)
      while (fe.InFlight() >= depth) fe.CollectFeatures();
    t_gen += t1 - t0;
    t_submit += t2 - t1;
    t_collect += now() - t2;
  }
  while (fe.CollectFeatures()) {
  }
  if (host_us) {
    const double nfr = double(std::max<uint64_t>(last - halo_first, 1));
    host_us[0] = 1e6 * t_gen / nfr;
    host_us[1] = 1e6 * t_submit / nfr;
    host_us[2] = 1e6 * t_collect / nfr;
  }
  slam_types::SLAMProblem problem;
  fe.GetSLAMProblem(&problem);
  // serialise the three arrays separately: the wire format is three length-prefixed arrays, so
  // a piece is the array bodies and the merged message is the concatenation behind summed counts
  SLAMProblemPiece piece;
  piece.n_nodes = uint32_t(problem.nodes.size());
  piece.n_vision_factors = uint32_t(problem.vision_factors.size());
  piece.n_odometry_factors = uint32_t(problem.odometry_factors.size());
  slam_types::SLAMProblem part;
  part.nodes = problem.nodes;
  std::vector<uint8_t> w = Frontend::SerializeSLAMProblem(part);          // [n][nodes...][0][0]
  piece.nodes.assign(w.begin() + 4, w.end() - 8);
  part = slam_types::SLAMProblem();
  part.vision_factors = problem.vision_factors;
  w = Frontend::SerializeSLAMProblem(part);                               // [0][n][factors...][0]
  piece.vision_factors.assign(w.begin() + 8, w.end() - 4);
  part = slam_types::SLAMProblem();
  part.odometry_factors = problem.odometry_factors;
  w = Frontend::SerializeSLAMProblem(part);                               // [0][0][n][factors...]
  piece.odometry_factors.assign(w.begin() + 12, w.end());
  return piece;
}

std::vector<uint8_t> MergeSLAMProblemPieces(const std::vector<SLAMProblemPiece>& pieces) {
  std::vector<uint8_t> b;
  uint32_t n = 0;
  for (const SLAMProblemPiece& p : pieces) n += p.n_nodes;
  Put<uint32_t>(&b, n);
  for (const SLAMProblemPiece& p : pieces) b.insert(b.end(), p.nodes.begin(), p.nodes.end());
  n = 0;
  for (const SLAMProblemPiece& p : pieces) n += p.n_vision_factors;
  Put<uint32_t>(&b, n);
  for (const SLAMProblemPiece& p : pieces) b.insert(b.end(), p.vision_factors.begin(), p.vision_factors.end());
  n = 0;
  for (const SLAMProblemPiece& p : pieces) n += p.n_odometry_factors;
  Put<uint32_t>(&b, n);
  for (const SLAMProblemPiece& p : pieces) b.insert(b.end(), p.odometry_factors.begin(), p.odometry_factors.end());
  return b;
}

}  // namespace slam
