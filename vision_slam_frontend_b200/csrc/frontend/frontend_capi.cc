// extern "C" handles onto slam::Frontend so that the Python tests (ctypes) can drive
// the C++ host mirror.  Not part of the product ABI (that is include/vsf.h); it exists
// for tests/test_gpu_frontend.py and as an example of embedding the Frontend.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "slam_frontend.h"
#include "synthetic_source.h"

namespace {
struct Handle {
  slam::Frontend* fe = nullptr;
  std::string err;
  std::vector<uint8_t> wire;
};
}  // namespace

extern "C" {

void* vsff_create(int device, int max_features, int desc_bytes, int frame_life, float best_percent,
                  float nn_match_ratio, const float* P_left, const float* P_right, const float* fundamental,
                  const float* K_left, const float* dist_left, int exact_std_sort) {
  Handle* h = new Handle();
  try {
    slam::FrontendConfig cfg;
    cfg.cuda_device = device;
    cfg.max_features = max_features;
    cfg.descriptor_bytes = desc_bytes;
    cfg.frame_life_ = uint32_t(frame_life);
    cfg.best_percent_ = best_percent;
    cfg.nn_match_ratio_ = nn_match_ratio;
    cfg.exact_std_sort = exact_std_sort != 0;
    if (P_left) std::memcpy(cfg.projection_left, P_left, 12 * sizeof(float));
    if (P_right) std::memcpy(cfg.projection_right, P_right, 12 * sizeof(float));
    if (fundamental) std::memcpy(cfg.fundamental, fundamental, 9 * sizeof(float));
    if (K_left) std::memcpy(cfg.camera_matrix_left, K_left, 9 * sizeof(float));
    if (dist_left) std::memcpy(cfg.distortion_coeffs_left, dist_left, 5 * sizeof(float));
    h->fe = new slam::Frontend(cfg);
  } catch (const std::exception& e) {
    h->err = e.what();
  }
  return h;
}

void vsff_destroy(void* p) {
  Handle* h = static_cast<Handle*>(p);
  if (!h) return;
  delete h->fe;
  delete h;
}

const char* vsff_last_error(void* p) { return static_cast<Handle*>(p)->err.c_str(); }
int vsff_ok(void* p) { return static_cast<Handle*>(p)->fe != nullptr; }

void vsff_default_config(float* P_left, float* P_right, float* fundamental, float* K_left, float* dist_left) {
  slam::FrontendConfig cfg;
  std::memcpy(P_left, cfg.projection_left, 12 * sizeof(float));
  std::memcpy(P_right, cfg.projection_right, 12 * sizeof(float));
  std::memcpy(fundamental, cfg.fundamental, 9 * sizeof(float));
  std::memcpy(K_left, cfg.camera_matrix_left, 9 * sizeof(float));
  std::memcpy(dist_left, cfg.distortion_coeffs_left, 5 * sizeof(float));
}

void vsff_observe_odometry(void* p, const float* t, const float* q_wxyz, double ts) {
  Handle* h = static_cast<Handle*>(p);
  h->fe->ObserveOdometry(Eigen::Vector3f(t[0], t[1], t[2]),
                         Eigen::Quaternionf(q_wxyz[0], q_wxyz[1], q_wxyz[2], q_wxyz[3]), ts);
}

// 1 = node added, 0 = odometry gate said no, -1 = error (see vsff_last_error)
int vsff_observe_features(void* p, const vsf_keypoint* kl, const uint8_t* dl, int nl, const vsf_keypoint* kr,
                          const uint8_t* dr, int nr, int desc_bytes, double time) {
  Handle* h = static_cast<Handle*>(p);
  try {
    std::vector<cv::KeyPoint> lk(nl), rk(nr);
    if (nl) std::memcpy(static_cast<void*>(lk.data()), kl, size_t(nl) * sizeof(vsf_keypoint));
    if (nr) std::memcpy(static_cast<void*>(rk.data()), kr, size_t(nr) * sizeof(vsf_keypoint));
    cv::Mat ld(nl, desc_bytes, CV_8U, const_cast<uint8_t*>(dl));
    cv::Mat rd(nr, desc_bytes, CV_8U, const_cast<uint8_t*>(dr));
    return h->fe->ObserveFeatures(lk, ld, rk, rd, time) ? 1 : 0;
  } catch (const std::exception& e) {
    h->err = e.what();
    return -1;
  }
}

int vsff_get_matches(void* p, const uint8_t* q, int nq, const uint8_t* t, int nt, int desc_bytes, double ratio,
                     vsf_dmatch* out, int cap) {
  Handle* h = static_cast<Handle*>(p);
  try {
    slam::Frame fq(std::vector<cv::KeyPoint>(nq), cv::Mat(nq, desc_bytes, CV_8U, const_cast<uint8_t*>(q)), 0);
    slam::Frame ft(std::vector<cv::KeyPoint>(nt), cv::Mat(nt, desc_bytes, CV_8U, const_cast<uint8_t*>(t)), 1);
    std::vector<cv::DMatch> m = h->fe->GetMatches(fq, ft, ratio);
    if (int(m.size()) > cap) return -2;
    if (!m.empty()) std::memcpy(out, m.data(), m.size() * sizeof(vsf_dmatch));
    return int(m.size());
  } catch (const std::exception& e) {
    h->err = e.what();
    return -1;
  }
}

int vsff_num_poses(void* p) { return static_cast<Handle*>(p)->fe->GetNumPoses(); }
float vsff_stereo_threshold(void* p) { return static_cast<Handle*>(p)->fe->GetStereoAmbigConstraint(); }

int vsff_num_vision_factors(void* p) {
  slam_types::SLAMProblem pr;
  static_cast<Handle*>(p)->fe->GetSLAMProblem(&pr);
  return int(pr.vision_factors.size());
}

int vsff_vision_factor(void* p, int i, uint64_t* pose_initial, uint64_t* pose_current, vsf_feature_match* out,
                       int cap) {
  slam_types::SLAMProblem pr;
  static_cast<Handle*>(p)->fe->GetSLAMProblem(&pr);
  if (i < 0 || i >= int(pr.vision_factors.size())) return -1;
  const slam_types::VisionFactor& v = pr.vision_factors[i];
  *pose_initial = v.pose_idx_initial;
  *pose_current = v.pose_idx_current;
  const int n = int(v.feature_matches.size());
  if (n > cap) return -2;
  for (int k = 0; k < n; ++k) {
    out[k].feature_idx_initial = v.feature_matches[k].feature_idx_initial;
    out[k].feature_idx_current = v.feature_matches[k].feature_idx_current;
  }
  return n;
}

int vsff_node_features(void* p, int node, float* pixel_xy, float* point3d, int cap) {
  slam_types::SLAMProblem pr;
  static_cast<Handle*>(p)->fe->GetSLAMProblem(&pr);
  if (node < 0 || node >= int(pr.nodes.size())) return -1;
  const slam_types::SLAMNode& n = pr.nodes[node];
  const int m = int(n.features.size());
  if (m > cap) return -2;
  for (int k = 0; k < m; ++k) {
    pixel_xy[2 * k] = n.features[k].pixel.x();
    pixel_xy[2 * k + 1] = n.features[k].pixel.y();
    point3d[3 * k] = n.features[k].point3d.x();
    point3d[3 * k + 1] = n.features[k].point3d.y();
    point3d[3 * k + 2] = n.features[k].point3d.z();
  }
  return m;
}

int vsff_node_pose(void* p, int node, float* loc, float* quat_xyzw, double* timestamp) {
  slam_types::SLAMProblem pr;
  static_cast<Handle*>(p)->fe->GetSLAMProblem(&pr);
  if (node < 0 || node >= int(pr.nodes.size())) return -1;
  const slam_types::SLAMNode& n = pr.nodes[node];
  loc[0] = n.pose.loc.x(); loc[1] = n.pose.loc.y(); loc[2] = n.pose.loc.z();
  quat_xyzw[0] = n.pose.angle.x(); quat_xyzw[1] = n.pose.angle.y();
  quat_xyzw[2] = n.pose.angle.z(); quat_xyzw[3] = n.pose.angle.w();
  *timestamp = n.timestamp;
  return 0;
}

int vsff_num_odometry_factors(void* p) {
  slam_types::SLAMProblem pr;
  static_cast<Handle*>(p)->fe->GetSLAMProblem(&pr);
  return int(pr.odometry_factors.size());
}

int vsff_odometry_factor(void* p, int i, uint64_t* pose_i, uint64_t* pose_j, float* translation, float* quat_xyzw) {
  slam_types::SLAMProblem pr;
  static_cast<Handle*>(p)->fe->GetSLAMProblem(&pr);
  if (i < 0 || i >= int(pr.odometry_factors.size())) return -1;
  const slam_types::OdometryFactor& o = pr.odometry_factors[i];
  *pose_i = o.pose_i; *pose_j = o.pose_j;
  translation[0] = o.translation.x(); translation[1] = o.translation.y(); translation[2] = o.translation.z();
  quat_xyzw[0] = o.rotation.x(); quat_xyzw[1] = o.rotation.y(); quat_xyzw[2] = o.rotation.z(); quat_xyzw[3] = o.rotation.w();
  return 0;
}

// Serialise the current SLAMProblem in ROS1 wire format; returns the byte count (call with
// buf = NULL to size the buffer).
size_t vsff_serialize_problem(void* p, uint8_t* buf, size_t cap) {
  Handle* h = static_cast<Handle*>(p);
  slam_types::SLAMProblem pr;
  h->fe->GetSLAMProblem(&pr);
  h->wire = slam::Frontend::SerializeSLAMProblem(pr);
  if (buf && cap >= h->wire.size()) std::memcpy(buf, h->wire.data(), h->wire.size());
  return h->wire.size();
}


// ---- pipelined calls and shards -------------------------------------------------------------
// 1 = frame submitted, 0 = odometry gate said no, -1 = error.  The descriptor rows are copied:
// the Python caller's arrays need not outlive the call.
int vsff_submit_features(void* p, const vsf_keypoint* kl, const uint8_t* dl, int nl, const vsf_keypoint* kr,
                         const uint8_t* dr, int nr, int desc_bytes, double time) {
  Handle* h = static_cast<Handle*>(p);
  try {
    std::vector<cv::KeyPoint> lk(nl), rk(nr);
    if (nl) std::memcpy(static_cast<void*>(lk.data()), kl, size_t(nl) * sizeof(vsf_keypoint));
    if (nr) std::memcpy(static_cast<void*>(rk.data()), kr, size_t(nr) * sizeof(vsf_keypoint));
    cv::Mat ld = cv::Mat(nl, desc_bytes, CV_8U, const_cast<uint8_t*>(dl)).clone();
    cv::Mat rd = cv::Mat(nr, desc_bytes, CV_8U, const_cast<uint8_t*>(dr)).clone();
    return h->fe->SubmitFeatures(lk, ld, rk, rd, time) ? 1 : 0;
  } catch (const std::exception& e) {
    h->err = e.what();
    return -1;
  }
}

int vsff_collect_features(void* p) {
  Handle* h = static_cast<Handle*>(p);
  try {
    return h->fe->CollectFeatures() ? 1 : 0;
  } catch (const std::exception& e) {
    h->err = e.what();
    return -1;
  }
}

int vsff_in_flight(void* p) { return static_cast<Handle*>(p)->fe->InFlight(); }

int vsff_start_shard(void* p, uint64_t halo_first, uint64_t first) {
  Handle* h = static_cast<Handle*>(p);
  try {
    h->fe->StartShard(halo_first, first);
    return 0;
  } catch (const std::exception& e) {
    h->err = e.what();
    return -1;
  }
}

// ---- synthetic frame source (csrc/frontend/synthetic_source.h) --------------------------------
static slam::FrontendConfig SynthRig(int device, int features, int desc_bytes, int frame_life, int exact) {
  return slam::SyntheticRig(device, features, desc_bytes, frame_life, exact != 0);
}

// The rig of the synthetic runs (P_left 12, P_right 12, fundamental 9, K_left 9, dist_left 5).
void vsff_synthetic_rig(float* P_left, float* P_right, float* fundamental, float* K_left, float* dist_left) {
  const slam::FrontendConfig cfg = slam::SyntheticRig(0, 1, 32, 10, true);
  std::memcpy(P_left, cfg.projection_left, 12 * sizeof(float));
  std::memcpy(P_right, cfg.projection_right, 12 * sizeof(float));
  std::memcpy(fundamental, cfg.fundamental, 9 * sizeof(float));
  std::memcpy(K_left, cfg.camera_matrix_left, 9 * sizeof(float));
  std::memcpy(dist_left, cfg.distortion_coeffs_left, 5 * sizeof(float));
}

// One frame + its odometry message (kl / kr: features records, dl / dr: features x desc_bytes).
int vsff_synth_frame(int features, int desc_bytes, uint64_t seed, uint64_t pose, vsf_keypoint* kl, uint8_t* dl,
                     vsf_keypoint* kr, uint8_t* dr, float* odom_t3, float* odom_q_wxyz, double* timestamp) {
  try {
    slam::SyntheticStereoConfig sc;
    sc.features = features;
    sc.landmark_stride = std::max(1, features / 10);
    sc.seed = seed;
    slam::SyntheticStereoSource src(SynthRig(0, features, desc_bytes, 10, 1), sc);
    std::vector<cv::KeyPoint> lk, rk;
    cv::Mat ld, rd;
    src.Frame(pose, &lk, &ld, &rk, &rd);
    std::memcpy(static_cast<void*>(kl), lk.data(), lk.size() * sizeof(vsf_keypoint));
    std::memcpy(static_cast<void*>(kr), rk.data(), rk.size() * sizeof(vsf_keypoint));
    for (int i = 0; i < features; ++i) {
      std::memcpy(dl + size_t(i) * desc_bytes, ld.ptr(i), size_t(desc_bytes));
      std::memcpy(dr + size_t(i) * desc_bytes, rd.ptr(i), size_t(desc_bytes));
    }
    Eigen::Vector3f t;
    Eigen::Quaternionf q;
    src.Odometry(pose, &t, &q, timestamp);
    odom_t3[0] = t.x(); odom_t3[1] = t.y(); odom_t3[2] = t.z();
    odom_q_wxyz[0] = q.w(); odom_q_wxyz[1] = q.x(); odom_q_wxyz[2] = q.y(); odom_q_wxyz[3] = q.z();
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

// Poses [0, n_poses) of the synthetic sequence run as `world` shards - one after the other in
// this process, each on a fresh Frontend with its halo, exactly what the ranks of
// vsf_sequence_driver do - and merged into ONE SLAMProblem message.  Returns the byte count and
// a malloc'ed buffer (vsff_free), 0 on error (message in err, up to err_cap bytes).
size_t vsff_run_synthetic_sequence(int device, int features, int desc_bytes, int frame_life, int n_poses, int world,
                                   int in_flight, uint64_t seed, int exact, uint8_t** out, char* err, int err_cap) {
  try {
    slam::FrontendConfig rig = SynthRig(device, features, desc_bytes, frame_life, exact);
    slam::SyntheticStereoConfig sc;
    sc.features = features;
    sc.landmark_stride = std::max(1, features / 10);
    sc.seed = seed;
    slam::SyntheticStereoSource src(rig, sc);
    std::vector<slam::SLAMProblemPiece> pieces;
    for (int r = 0; r < world; ++r) {
      const uint64_t base = uint64_t(n_poses) / world, rem = uint64_t(n_poses) % world;
      const uint64_t first = r * base + std::min<uint64_t>(r, rem), last = first + base + (uint64_t(r) < rem ? 1 : 0);
      // through Pack / Unpack, the form a piece travels in between ranks
      const std::vector<uint8_t> blob = slam::RunSequenceShard(rig, src, first, last, in_flight).Pack();
      pieces.push_back(slam::SLAMProblemPiece::Unpack(blob.data(), blob.size()));
    }
    const std::vector<uint8_t> wire = slam::MergeSLAMProblemPieces(pieces);
    *out = static_cast<uint8_t*>(std::malloc(wire.size() ? wire.size() : 1));
    std::memcpy(*out, wire.data(), wire.size());
    return wire.size();
  } catch (const std::exception& e) {
    if (err && err_cap > 0) std::snprintf(err, size_t(err_cap), "%s", e.what());
    return 0;
  }
}

void vsff_free(void* p) { std::free(p); }

}  // extern "C"
