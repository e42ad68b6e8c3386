// Host orchestration of the B200 path; see slam_frontend.h.  Every block cites the
// reference lines (relative to the reference checkout) whose behaviour it restates.
#include "slam_frontend.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>
#include <stdexcept>

namespace slam {

using slam_types::FeatureMatch;
using slam_types::OdometryFactor;
using slam_types::RobotPose;
using slam_types::SLAMNode;
using slam_types::SLAMProblem;
using slam_types::VisionFactor;
using slam_types::VisionFeature;

namespace {

void MatMul3x3_3x4(const float* K, const float* A, float* P) {   // float, left-to-right sums
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      float acc = K[r * 3 + 0] * A[0 * 4 + c];
      acc += K[r * 3 + 1] * A[1 * 4 + c];
      acc += K[r * 3 + 2] * A[2 * 4 + c];
      P[r * 4 + c] = acc;
    }
}

void Inv3(const double* m, double* o) {
  const double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
                   m[2] * (m[3] * m[7] - m[4] * m[6]);
  o[0] = (m[4] * m[8] - m[5] * m[7]) / d;
  o[1] = (m[2] * m[7] - m[1] * m[8]) / d;
  o[2] = (m[1] * m[5] - m[2] * m[4]) / d;
  o[3] = (m[5] * m[6] - m[3] * m[8]) / d;
  o[4] = (m[0] * m[8] - m[2] * m[6]) / d;
  o[5] = (m[2] * m[3] - m[0] * m[5]) / d;
  o[6] = (m[3] * m[7] - m[4] * m[6]) / d;
  o[7] = (m[1] * m[6] - m[0] * m[7]) / d;
  o[8] = (m[0] * m[4] - m[1] * m[3]) / d;
}

void Mul3(const double* a, const double* b, double* o) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) o[r * 3 + c] = a[r * 3] * b[c] + a[r * 3 + 1] * b[3 + c] + a[r * 3 + 2] * b[6 + c];
}

void Transpose3(const double* a, double* o) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) o[c * 3 + r] = a[r * 3 + c];
}

void CameraMatrix(const CameraIntrinsics& I, float* M) {   // src/slam_frontend.cc:542-548
  const float m[9] = {I.fx, 0.f, I.cx, 0.f, I.fy, I.cy, 0.f, 0.f, 1.f};
  std::memcpy(M, m, sizeof(m));
}

struct ByDistance {   // cv::DMatch::operator< — distance only
  bool operator()(const cv::DMatch& a, const cv::DMatch& b) const { return a.distance < b.distance; }
};

}  // namespace

// ------------------------------------------------------------------------------------ config

FrontendConfig::FrontendConfig() {
  // defaults of the reference (src/slam_frontend.cc:550-559)
  debug_images_ = false;   // the reference defaults to true; debug images are out of scope here
  descriptor_extract_type_ = DescriptorExtractorType::AKAZE;
  best_percent_ = 0.3f;
  nn_match_ratio_ = 0.6f;
  frame_life_ = 10;
  min_odom_rotation = static_cast<float>(10.0 / 180.0 * M_PI);
  min_odom_translation = 0.2f;
  min_vision_matches = 10;
  bf_matcher_param_ = cv::NORM_HAMMING;
  // PointGrey stereo calibration used by the reference (src/slam_frontend.cc:565-583)
  intrinsics_left.fx = 527.873518f;
  intrinsics_left.cx = 482.823413f;
  intrinsics_left.fy = 527.276819f;
  intrinsics_left.cy = 298.033945f;
  intrinsics_left.k1 = -0.153137f;
  intrinsics_left.k2 = 0.075666f;
  intrinsics_left.p1 = -0.000227f;
  intrinsics_left.p2 = -0.000320f;
  intrinsics_left.k3 = 0.f;
  intrinsics_right.fx = 530.158021f;
  intrinsics_right.cx = 475.540633f;
  intrinsics_right.fy = 529.682234f;
  intrinsics_right.cy = 299.995465f;
  intrinsics_right.k1 = -0.156833f;
  intrinsics_right.k2 = 0.081841f;
  intrinsics_right.p1 = -0.000779f;
  intrinsics_right.p2 = -0.000356f;
  intrinsics_right.k3 = -0.000779f;
  // right-camera extrinsics [R | t] (src/slam_frontend.cc:603-610)
  const float A_right[12] = {
      0.999593617649873f, 0.021411909431148f, -0.018818333830411f, -0.131707087331978f,
      -0.021140534893290f, 0.999671312094879f, 0.014503294761121f, 0.003232397463343f,
      0.019122691705565f, -0.014099571235136f, 0.999717722536176f, -0.001146108483477f};
  std::memcpy(stereo_extrinsics, A_right, sizeof(A_right));
  // camera -> robot transform (src/slam_frontend.cc:613-618)
  const float XT[3] = {-0.01f, 0.06f, 0.5299999713897705f};
  const float RT[9] = {0.009916590468f, -0.2835522866f, 0.9589055021f,
                       -0.9998698619f, -0.01501486552f, 0.005900269087f,
                       0.01272480238f, -0.9588392225f, -0.2836642819f};
  std::memcpy(left_cam_to_robot_translation, XT, sizeof(XT));
  std::memcpy(left_cam_to_robot_rotation, RT, sizeof(RT));
  cuda_device = 0;
  max_features = 20000;
  descriptor_bytes = 61;   // AKAZE MLDB, the reference's default extractor
  exact_std_sort = true;
  strict_reference_threshold = true;
  residual_order = 0;
  UpdateDerived();
}

void FrontendConfig::UpdateDerived() {
  CameraMatrix(intrinsics_left, camera_matrix_left);     // src/slam_frontend.cc:585-593
  CameraMatrix(intrinsics_right, camera_matrix_right);
  const float A_left[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  MatMul3x3_3x4(camera_matrix_left, A_left, projection_left);            // :595-600, :621
  MatMul3x3_3x4(camera_matrix_right, stereo_extrinsics, projection_right);  // :602-611, :622
  const CameraIntrinsics* I[2] = {&intrinsics_left, &intrinsics_right};
  float* D[2] = {distortion_coeffs_left, distortion_coeffs_right};
  for (int k = 0; k < 2; ++k) {                          // :623-634
    D[k][0] = I[k]->k1;
    D[k][1] = I[k]->k2;
    D[k][2] = I[k]->p1;
    D[k][3] = I[k]->p2;
    D[k][4] = I[k]->k3;
  }
  // Fundamental matrix.  The reference's own construction (:635-644) indexes a
  // 3-vector with [3] and is undefined; this is the textbook matrix in the reference's
  // convention x_left^T F x_right = 0:  F = (K_r^-T [t]x R K_l^-1)^T, in double.
  double Kl[9], Kr[9], R[9], Kli[9], Kri[9], KriT[9], tx[9], a[9], b[9], Fstd[9], F[9];
  for (int i = 0; i < 9; ++i) {
    Kl[i] = camera_matrix_left[i];
    Kr[i] = camera_matrix_right[i];
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = stereo_extrinsics[r * 4 + c];
  const double t[3] = {stereo_extrinsics[3], stereo_extrinsics[7], stereo_extrinsics[11]};
  const double txm[9] = {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0};
  std::memcpy(tx, txm, sizeof(tx));
  Inv3(Kl, Kli);
  Inv3(Kr, Kri);
  Transpose3(Kri, KriT);
  Mul3(KriT, tx, a);
  Mul3(a, R, b);
  Mul3(b, Kli, Fstd);
  Transpose3(Fstd, F);
  for (int i = 0; i < 9; ++i) fundamental[i] = static_cast<float>(F[i]);
}

void FrontendConfig::Load(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("FrontendConfig::Load: cannot open " + path);
  std::string line;
  while (std::getline(in, line)) {
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line.resize(hash);
    std::istringstream ss(line);
    std::string key;
    double v;
    if (!(ss >> key)) continue;
    if (!(ss >> v)) throw std::runtime_error("FrontendConfig::Load: no value for " + key);
    if (key == "best_percent") best_percent_ = float(v);
    else if (key == "nn_match_ratio") nn_match_ratio_ = float(v);
    else if (key == "frame_life") frame_life_ = uint32_t(v);
    else if (key == "min_odom_translation") min_odom_translation = float(v);
    else if (key == "min_odom_rotation") min_odom_rotation = float(v);
    else if (key == "min_vision_matches") min_vision_matches = uint32_t(v);
    else if (key == "max_features") max_features = int(v);
    else if (key == "descriptor_bytes") descriptor_bytes = int(v);
    else if (key == "cuda_device") cuda_device = int(v);
    else if (key == "exact_std_sort") exact_std_sort = (v != 0);
    else if (key == "strict_reference_threshold") strict_reference_threshold = (v != 0);
    else if (key == "residual_order") residual_order = int(v);
    else throw std::runtime_error("FrontendConfig::Load: unknown key " + key);
  }
}

// ------------------------------------------------------------------------------------- Frame

Frame::Frame(const std::vector<cv::KeyPoint>& keypoints, const cv::Mat& descriptors, uint64_t frame_ID) {
  // src/slam_frontend.cc:511-519
  keypoints_ = keypoints;
  descriptors_ = descriptors;
  frame_ID_ = frame_ID;
  is_initial_ = std::vector<bool>(keypoints_.size(), true);
  initial_ids_ = std::vector<int64_t>(keypoints_.size(), -1);
}

// ---------------------------------------------------------------------------------- Frontend

// Both constructors start the odometry members from zero / identity explicitly (the reference
// leaves them indeterminate, src/slam_frontend.cc:188-190; with the real Eigen types a default
// constructed Vector3f / Quaternionf holds garbage).
Frontend::Frontend(const std::string& config_path)
    : odom_initialized_(false),
      init_odom_translation_(0.f, 0.f, 0.f), init_odom_rotation_(1.f, 0.f, 0.f, 0.f),
      prev_odom_translation_(0.f, 0.f, 0.f), prev_odom_rotation_(1.f, 0.f, 0.f, 0.f),
      odom_translation_(0.f, 0.f, 0.f), odom_rotation_(1.f, 0.f, 0.f, 0.f),
      odom_timestamp_(0), ctx_(nullptr), curr_frame_ID_(0), first_output_ID_(0) {
  // The reference ignores its argument (src/slam_frontend.cc:188-190); a non-empty
  // path is honoured here through FrontendConfig::Load.
  if (!config_path.empty()) config_.Load(config_path);
  Check(vsf_create(config_.cuda_device, config_.max_features, config_.descriptor_bytes,
                   int(config_.frame_life_), &ctx_), "vsf_create");
  ApplyOptions();
}

void Frontend::ApplyOptions() {
  Check(vsf_set_option(ctx_, VSF_OPT_HOLD_THRESHOLD_ON_EMPTY, config_.strict_reference_threshold ? 0 : 1), "vsf_set_option");
  Check(vsf_set_option(ctx_, VSF_OPT_RESIDUAL_ORDER, config_.residual_order ? 1 : 0), "vsf_set_option");
}

Frontend::Frontend(const FrontendConfig& config)
    : odom_initialized_(false),
      init_odom_translation_(0.f, 0.f, 0.f), init_odom_rotation_(1.f, 0.f, 0.f, 0.f),
      prev_odom_translation_(0.f, 0.f, 0.f), prev_odom_rotation_(1.f, 0.f, 0.f, 0.f),
      odom_translation_(0.f, 0.f, 0.f), odom_rotation_(1.f, 0.f, 0.f, 0.f),
      odom_timestamp_(0), config_(config), ctx_(nullptr), curr_frame_ID_(0), first_output_ID_(0) {
  Check(vsf_create(config_.cuda_device, config_.max_features, config_.descriptor_bytes,
                   int(config_.frame_life_), &ctx_), "vsf_create");
  ApplyOptions();
}

Frontend::~Frontend() {
  if (ctx_) vsf_destroy(ctx_);
}

void Frontend::Check(int rc, const char* what) {
  if (rc == VSF_OK) return;
  std::string msg = std::string(what) + " failed (code " + std::to_string(rc) + ")";
  if (ctx_) msg += std::string(": ") + vsf_last_error(ctx_);
  else msg += ": no CUDA device / library (there is no CPU fallback)";
#ifdef VSF_FRONTEND_ABORT
  std::fprintf(stderr, "%s\n", msg.c_str());
  std::abort();   // the reference's glog CHECK behaviour
#else
  throw std::runtime_error(msg);
#endif
}

bool Frontend::OdomCheck() {   // src/slam_frontend.cc:175-186
  if (!odom_initialized_) return false;
  if ((prev_odom_translation_ - odom_translation_).norm() > config_.min_odom_translation) return true;
  if (prev_odom_rotation_.angularDistance(odom_rotation_) > config_.min_odom_rotation) return true;
  return false;
}

void Frontend::ObserveOdometry(const Eigen::Vector3f& translation, const Eigen::Quaternionf& rotation,
                               double timestamp) {   // src/slam_frontend.cc:250-263
  if (!odom_initialized_) {
    init_odom_rotation_ = rotation;
    init_odom_translation_ = translation;
    // (the reference copies the still-uninitialised odom_* members here; the first
    // OdomCheck therefore compares against indeterminate values.  The constructors
    // initialise them to zero / identity, so that is what is copied.)
    prev_odom_rotation_ = odom_rotation_;
    prev_odom_translation_ = odom_translation_;
    odom_initialized_ = true;
  }
  odom_translation_ = translation;
  odom_rotation_ = rotation;
  odom_timestamp_ = timestamp;
}

void Frontend::AddOdometryFactor() {   // src/slam_frontend.cc:311-321
  const Eigen::Vector3f translation = prev_odom_rotation_.inverse() * (odom_translation_ - prev_odom_translation_);
  const Eigen::Quaternionf rotation(odom_rotation_ * prev_odom_rotation_.inverse());
  odometry_factors_.push_back(OdometryFactor(curr_frame_ID_ - 1, curr_frame_ID_, translation, rotation));
}

std::vector<cv::DMatch> Frontend::GetMatches(const Frame& frame_query, const Frame& frame_train,
                                             double nn_match_ratio) {
  // src/slam_frontend.cc:521-538: knnMatch(k=2) + ratio test, on the device
  const int nq = frame_query.descriptors_.rows, nt = frame_train.descriptors_.rows;
  if ((nq > 0 && frame_query.descriptors_.cols != config_.descriptor_bytes) ||
      (nt > 0 && frame_train.descriptors_.cols != config_.descriptor_bytes))
    throw std::runtime_error("Frontend::GetMatches: descriptor width differs from FrontendConfig::descriptor_bytes");
  std::vector<vsf_dmatch> out(size_t(std::max(nq, 1)));
  int n = 0;
  Check(vsf_get_matches(ctx_, frame_query.descriptors_.data, nq, frame_query.descriptors_.step,
                        frame_train.descriptors_.data, nt, frame_train.descriptors_.step, nn_match_ratio,
                        out.data(), int(out.size()), &n), "vsf_get_matches");
  std::vector<cv::DMatch> best_matches(static_cast<size_t>(n));
  static_assert(sizeof(cv::DMatch) == sizeof(vsf_dmatch), "DMatch layout");
  if (n) std::memcpy(static_cast<void*>(best_matches.data()), out.data(), size_t(n) * sizeof(vsf_dmatch));
  return best_matches;
}

VisionFactor Frontend::FinishFeatureMatches(std::vector<cv::DMatch>* matches_ptr, float best_percent,
                                            Frame* past_frame_ptr, Frame* curr_frame_ptr,
                                            std::vector<cv::DMatch>* sorted_out) {
  std::vector<cv::DMatch>& matches = *matches_ptr;
  Frame& past_frame = *past_frame_ptr;
  Frame& curr_frame = *curr_frame_ptr;
  // src/slam_frontend.cc:289 — std::sort with DMatch::operator< (distance only).  The
  // input is in ascending queryIdx order exactly like the reference's, so the
  // libstdc++ permutation is the reference's permutation.  exact_std_sort = false
  // uses the (distance, queryIdx) order the device-side sort produces.
  if (config_.exact_std_sort) std::sort(matches.begin(), matches.end(), ByDistance());
  else std::stable_sort(matches.begin(), matches.end(), ByDistance());
  const int num_good_matches = static_cast<int>(matches.size() * best_percent);   // :290
  matches.erase(matches.begin() + num_good_matches, matches.end());               // :291
  std::vector<FeatureMatch> pairs;
  pairs.reserve(matches.size());
  for (const cv::DMatch& match : matches) {                                        // :293-306
    pairs.push_back(FeatureMatch(match.queryIdx, match.trainIdx));
    if (curr_frame.is_initial_[match.trainIdx]) {
      curr_frame.is_initial_[match.trainIdx] = false;
      curr_frame.initial_ids_[match.trainIdx] = past_frame.is_initial_[match.queryIdx]
                                                    ? int64_t(past_frame.frame_ID_)
                                                    : past_frame.initial_ids_[match.queryIdx];
    }
  }
  if (sorted_out) *sorted_out = matches;
  return VisionFactor(past_frame.frame_ID_, curr_frame.frame_ID_, pairs);          // :308
}

VisionFactor* Frontend::GetFeatureMatches(Frame* past_frame_ptr, Frame* curr_frame_ptr) {
  // src/slam_frontend.cc:282-309.  The caller owns the result (delete it).
  std::vector<cv::DMatch> matches = GetMatches(*past_frame_ptr, *curr_frame_ptr, config_.nn_match_ratio_);
  return new VisionFactor(
      FinishFeatureMatches(&matches, config_.best_percent_, past_frame_ptr, curr_frame_ptr, nullptr));
}

void Frontend::UndistortFeaturePoints(std::vector<VisionFeature>* features_ptr) {
  // src/slam_frontend.cc:323-351: cv::undistortPoints(pts, K_left, dist_left, R = {},
  // P = K_left).  Published algorithm: normalise, five fixed-point iterations of the
  // radial/tangential model in double, re-project with K, store float.
  std::vector<VisionFeature>& features = *features_ptr;
  const double fx = config_.camera_matrix_left[0], fy = config_.camera_matrix_left[4];
  const double cx = config_.camera_matrix_left[2], cy = config_.camera_matrix_left[5];
  const double k1 = config_.distortion_coeffs_left[0], k2 = config_.distortion_coeffs_left[1];
  const double p1 = config_.distortion_coeffs_left[2], p2 = config_.distortion_coeffs_left[3];
  const double k3 = config_.distortion_coeffs_left[4];
  for (VisionFeature& f : features) {
    const double x0 = (double(f.pixel.x()) - cx) / fx, y0 = (double(f.pixel.y()) - cy) / fy;
    double x = x0, y = y0;
    for (int it = 0; it < 5; ++it) {
      const double r2 = x * x + y * y;
      const double icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2);
      if (icdist < 0) {
        x = x0;
        y = y0;
        continue;
      }
      const double dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
      const double dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
      x = (x0 - dx) * icdist;
      y = (y0 - dy) * icdist;
    }
    f.pixel = Eigen::Vector2f(float(x * fx + cx), float(y * fy + cy));
  }
}

float Frontend::GetStereoAmbigConstraint() {
  float v = 0.f;
  Check(vsf_get_stereo_threshold(ctx_, &v), "vsf_get_stereo_threshold");
  return v;
}

void Frontend::SetStereoAmbigConstraint(float v) { Check(vsf_set_stereo_threshold(ctx_, v), "vsf_set_stereo_threshold"); }

bool Frontend::ObserveImage(const cv::Mat& left_image, const cv::Mat& right_image, double time) {
  // src/slam_frontend.cc:400-412: odometry gate, then ExtractFeatures on both images
  if (!OdomCheck()) return false;
  if (!extractor_)
    throw std::runtime_error("Frontend::ObserveImage: no FeatureExtractor installed (keypoint extraction "
                             "is the caller's input producer; use SetFeatureExtractor or ObserveFeatures)");
  std::vector<cv::KeyPoint> kl, kr;
  cv::Mat dl, dr;
  extractor_(left_image, &kl, &dl);
  extractor_(right_image, &kr, &dr);
  return ObserveFeatures(kl, dl, kr, dr, time);
}

void Frontend::StartShard(uint64_t halo_first, uint64_t first) {
  if (!nodes_.empty() || !frame_list_.empty() || !pending_.empty() || halo_first > first)
    throw std::runtime_error("Frontend::StartShard: call on a fresh Frontend, halo_first <= first");
  curr_frame_ID_ = halo_first;
  first_output_ID_ = first;
}

uint64_t Frontend::ShardHaloStart(uint64_t first, uint32_t frame_life) {
  const uint64_t halo = uint64_t(frame_life) + 1;
  return first > halo ? first - halo : 0;
}

bool Frontend::ObserveFeatures(const std::vector<cv::KeyPoint>& left_keypoints, const cv::Mat& left_descriptors,
                               const std::vector<cv::KeyPoint>& right_keypoints,
                               const cv::Mat& right_descriptors, double time) {
  if (!pending_.empty())
    throw std::runtime_error("Frontend::ObserveFeatures: frames of SubmitFeatures are still pending (CollectFeatures)");
  if (!SubmitFeatures(left_keypoints, left_descriptors, right_keypoints, right_descriptors, time)) return false;
  CollectFeatures();
  return true;
}

bool Frontend::SubmitFeatures(const std::vector<cv::KeyPoint>& left_keypoints, const cv::Mat& left_descriptors,
                              const std::vector<cv::KeyPoint>& right_keypoints, const cv::Mat& right_descriptors,
                              double /*time*/) {
  if (!OdomCheck()) return false;                                     // src/slam_frontend.cc:404
  const int nl = int(left_keypoints.size()), nr = int(right_keypoints.size());
  if (left_descriptors.rows != nl || right_descriptors.rows != nr)
    throw std::runtime_error("Frontend::ObserveFeatures: keypoint / descriptor row mismatch");
  if ((nl > 0 && left_descriptors.cols != config_.descriptor_bytes) ||
      (nr > 0 && right_descriptors.cols != config_.descriptor_bytes))
    throw std::runtime_error("Frontend::ObserveFeatures: descriptor width differs from FrontendConfig::descriptor_bytes");
  // host-side consistency checks come BEFORE the device call, which commits the frame
  if (int(pending_.size()) >= MaxInFlight())
    throw std::runtime_error("Frontend::SubmitFeatures: MaxInFlight() frames pending, call CollectFeatures first");
  const int expect = int(std::min<size_t>(frame_list_.size() + pending_.size(), config_.frame_life_));
  if (vsf_window_size(ctx_) != expect)
    throw std::runtime_error("Frontend::ObserveFeatures: device window out of sync with frame_list_");
  static_assert(sizeof(cv::KeyPoint) == sizeof(vsf_keypoint), "KeyPoint layout");
  vsf_observe_params prm;
  prm.fundamental = config_.fundamental;
  prm.P_left = config_.projection_left;
  prm.P_right = config_.projection_right;
  prm.K_left = config_.camera_matrix_left;              // N1 on the device, in the triangulation launch
  prm.dist_left = config_.distortion_coeffs_left;
  prm.nn_match_ratio = double(config_.nn_match_ratio_);
  // One fused device pass: stereo L->R kNN + ratio (:414-416), RemoveAmbigStereo (:417),
  // window kNN (:424-434), R'->L' kNN + triangulation (:437), undistortion (:443), window push
  // (:467-470).
  Check(vsf_observe_submit(ctx_, curr_frame_ID_, reinterpret_cast<const vsf_keypoint*>(left_keypoints.data()),
                           left_descriptors.data, nl, left_descriptors.step,
                           reinterpret_cast<const vsf_keypoint*>(right_keypoints.data()), right_descriptors.data, nr,
                           right_descriptors.step, &prm),
        "vsf_observe_submit");
  Pending pd;
  pd.left_keypoints = left_keypoints;
  pd.right_keypoints = right_keypoints;
  pd.left_descriptors = left_descriptors;      // shared, like cv::Mat: the caller keeps the rows alive
  pd.right_descriptors = right_descriptors;
  pd.frame_ID = curr_frame_ID_;
  pd.odom_translation = odom_translation_;
  pd.odom_rotation = odom_rotation_;
  pd.prev_odom_translation = prev_odom_translation_;
  pd.prev_odom_rotation = prev_odom_rotation_;
  pd.odom_timestamp = odom_timestamp_;
  pending_.push_back(std::move(pd));
  // :455-457 - what the next frame's odometry gate compares against does not depend on results
  prev_odom_rotation_ = odom_rotation_;
  prev_odom_translation_ = odom_translation_;
  curr_frame_ID_++;
  return true;
}

bool Frontend::CollectFeatures() {
  if (pending_.empty()) return false;
  Pending pd = std::move(pending_.front());
  pending_.pop_front();
  const std::vector<cv::KeyPoint>& left_keypoints = pd.left_keypoints;
  const std::vector<cv::KeyPoint>& right_keypoints = pd.right_keypoints;
  const int nl = int(left_keypoints.size());
  const int W = int(config_.frame_life_);
  // a window list has one entry per row of the PAST frame at most: size for the largest
  // resident frame as well as for this one
  int cap = std::max(nl, 1);
  for (const Frame& f : frame_list_) cap = std::max(cap, int(f.keypoints_.size()));
  kept_left_.resize(cap);
  kept_right_.resize(cap);
  frame_ids_.resize(W);
  window_counts_.resize(W);
  window_matches_.resize(size_t(W) * cap);
  tri_matches_.resize(cap);
  tri_X4_.resize(size_t(cap) * 4);
  xy_undist_.resize(size_t(cap) * 2);
  vsf_observe_out out;
  std::memset(&out, 0, sizeof(out));
  out.kept_left = kept_left_.data();
  out.kept_right = kept_right_.data();
  out.frame_ids = frame_ids_.data();
  out.window_counts = window_counts_.data();
  out.window_matches = window_matches_.data();
  out.tri_matches = tri_matches_.data();
  out.tri_X4 = tri_X4_.data();
  out.xy_undist = xy_undist_.data();
  out.cap = cap;
  uint64_t fid = 0;
  Check(vsf_observe_collect(ctx_, &fid, &out), "vsf_observe_collect");
  if (fid != pd.frame_ID) throw std::runtime_error("Frontend::CollectFeatures: frames collected out of order");

  // Host copies of the compacted frames (src/slam_frontend.cc:386-389, :396-397).
  const int M = out.n_kept;
  const int wbytes = config_.descriptor_bytes;
  std::vector<cv::KeyPoint> lk(M), rk(M);
  cv::Mat ld(M, wbytes, CV_8U), rd(M, wbytes, CV_8U);
  for (int i = 0; i < M; ++i) {
    lk[i] = left_keypoints[kept_left_[i]];
    rk[i] = right_keypoints[kept_right_[i]];
    std::memcpy(ld.ptr(i), pd.left_descriptors.ptr(kept_left_[i]), size_t(wbytes));
    std::memcpy(rd.ptr(i), pd.right_descriptors.ptr(kept_right_[i]), size_t(wbytes));
  }
  Frame curr_frame(lk, ld, pd.frame_ID);
  Frame right_temp_frame(rk, rd, pd.frame_ID);
  // halo frames of a shard (StartShard) only rebuild the window and the threshold
  const bool output = pd.frame_ID >= first_output_ID_;

  // Window loop (:424-434): one VisionFactor per resident past frame, oldest first,
  // pushed unconditionally.
  if (out.n_frames != int(frame_list_.size())) throw std::runtime_error("window out of sync with the device");
  for (int j = 0; j < out.n_frames; ++j) {
    Frame& past_frame = frame_list_[j];
    if (frame_ids_[j] != past_frame.frame_ID_) throw std::runtime_error("window frame id mismatch");
    const vsf_dmatch* src = window_matches_.data() + size_t(j) * cap;
    std::vector<cv::DMatch> matches(static_cast<size_t>(window_counts_[j]));
    if (!matches.empty()) std::memcpy(static_cast<void*>(matches.data()), src, matches.size() * sizeof(vsf_dmatch));
    VisionFactor factor = FinishFeatureMatches(&matches, config_.best_percent_, &past_frame, &curr_frame, nullptr);
    if (output) vision_factors_.push_back(std::move(factor));
  }

  // Calculate3DPoints (:117-173): GetFeatureMatches(right, left) with best_percent 1.0,
  // points in sorted-match order, Vector3f(x, y, z) / w.
  std::vector<Eigen::Vector3f> points;
  {
    std::vector<cv::DMatch> matches(static_cast<size_t>(out.n_tri));
    for (int i = 0; i < out.n_tri; ++i) {
      std::memcpy(static_cast<void*>(&matches[i]), &tri_matches_[i], sizeof(vsf_dmatch));
      matches[i].imgIdx = i;   // remember the position of the point triangulated for this match
    }
    std::vector<cv::DMatch> sorted;
    FinishFeatureMatches(&matches, 1.0f, &right_temp_frame, &curr_frame, &sorted);
    points.reserve(sorted.size());
    for (const cv::DMatch& m : sorted) {
      const float* X = tri_X4_.data() + size_t(m.imgIdx) * 4;
      points.push_back(Eigen::Vector3f(X[0], X[1], X[2]) / X[3]);
    }
  }

  // :438-443.  The reference indexes points[i] by keypoint index although `points` is in
  // sorted-match order and may be shorter (out-of-bounds read); NaN where it is missing.  The
  // pixel is the undistorted one (UndistortFeaturePoints, :443), computed on the device for the
  // compacted left keypoints.
  std::vector<VisionFeature> features;
  features.reserve(M);
  const float nan = std::numeric_limits<float>::quiet_NaN();
  for (int i = 0; i < M; ++i) {
    const Eigen::Vector3f p = (size_t(i) < points.size()) ? points[i] : Eigen::Vector3f(nan, nan, nan);
    features.push_back(VisionFeature(i, Eigen::Vector2f(xy_undist_[2 * i], xy_undist_[2 * i + 1]), p));
  }
  const Eigen::Vector3f loc = init_odom_rotation_.inverse() * (pd.odom_translation - init_odom_translation_);
  const Eigen::Quaternionf angle = pd.odom_rotation * init_odom_rotation_.inverse();   // :444-446
  if (output) nodes_.push_back(SLAMNode(pd.frame_ID, pd.odom_timestamp, RobotPose(loc, angle), features));
  if (output && pd.frame_ID > 0) {                                                      // :452-454
    const Eigen::Vector3f translation = pd.prev_odom_rotation.inverse() * (pd.odom_translation - pd.prev_odom_translation);
    const Eigen::Quaternionf rotation(pd.odom_rotation * pd.prev_odom_rotation.inverse());
    odometry_factors_.push_back(OdometryFactor(pd.frame_ID - 1, pd.frame_ID, translation, rotation));
  }
  if (frame_list_.size() >= config_.frame_life_) frame_list_.erase(frame_list_.begin());   // :467-469
  frame_list_.push_back(curr_frame);                                                    // :470
  return true;
}

void Frontend::GetSLAMProblem(SLAMProblem* problem) const {   // src/slam_frontend.cc:498-503
  *problem = SLAMProblem(nodes_, vision_factors_, odometry_factors_);
}

int Frontend::GetNumPoses() { return int(nodes_.size()); }   // :505-507

// ------------------------------------------------------------------------- ROS1 wire format

namespace {
template <typename T>
void Put(std::vector<uint8_t>* b, T v) {
  uint8_t raw[sizeof(T)];
  std::memcpy(raw, &v, sizeof(T));   // little-endian host
  b->insert(b->end(), raw, raw + sizeof(T));
}
}  // namespace

std::vector<uint8_t> Frontend::SerializeSLAMProblem(const SLAMProblem& problem) {
  // msg/SLAMProblem.msg = SLAMNode[] nodes, VisionFactor[] vision_factors,
  // OdometryFactor[] odometry_factors; ROS1: little-endian scalars, uint32 length
  // prefix per array, geometry_msgs/Point|Vector3 = 3 x float64, Quaternion = x,y,z,w.
  std::vector<uint8_t> b;
  Put<uint32_t>(&b, uint32_t(problem.nodes.size()));
  for (const SLAMNode& n : problem.nodes) {                 // src/slam_to_ros.h:73-83
    Put<uint64_t>(&b, n.node_idx);
    Put<double>(&b, n.timestamp);
    Put<double>(&b, n.pose.loc.x());
    Put<double>(&b, n.pose.loc.y());
    Put<double>(&b, n.pose.loc.z());
    Put<double>(&b, n.pose.angle.x());
    Put<double>(&b, n.pose.angle.y());
    Put<double>(&b, n.pose.angle.z());
    Put<double>(&b, n.pose.angle.w());
    Put<uint32_t>(&b, uint32_t(n.features.size()));
    for (const VisionFeature& f : n.features) {             // :44-58 (pixel.z = 0)
      Put<uint64_t>(&b, f.feature_idx);
      Put<double>(&b, f.pixel.x());
      Put<double>(&b, f.pixel.y());
      Put<double>(&b, 0.0);
      Put<double>(&b, f.point3d.x());
      Put<double>(&b, f.point3d.y());
      Put<double>(&b, f.point3d.z());
    }
  }
  Put<uint32_t>(&b, uint32_t(problem.vision_factors.size()));
  for (const VisionFactor& v : problem.vision_factors) {    // :85-94
    Put<uint64_t>(&b, v.pose_idx_initial);
    Put<uint64_t>(&b, v.pose_idx_current);
    Put<uint32_t>(&b, uint32_t(v.feature_matches.size()));
    for (const FeatureMatch& m : v.feature_matches) {       // :36-42 (id_initial, id_current)
      Put<uint64_t>(&b, m.feature_idx_initial);
      Put<uint64_t>(&b, m.feature_idx_current);
    }
  }
  Put<uint32_t>(&b, uint32_t(problem.odometry_factors.size()));
  for (const OdometryFactor& o : problem.odometry_factors) {   // :96-109
    Put<uint64_t>(&b, o.pose_i);
    Put<uint64_t>(&b, o.pose_j);
    Put<double>(&b, o.translation.x());
    Put<double>(&b, o.translation.y());
    Put<double>(&b, o.translation.z());
    Put<double>(&b, o.rotation.x());
    Put<double>(&b, o.rotation.y());
    Put<double>(&b, o.rotation.z());
    Put<double>(&b, o.rotation.w());
  }
  return b;
}

}  // namespace slam
