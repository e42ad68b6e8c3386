// C++ host side of the B200 path: slam::Frontend / slam::Frame / slam::FrontendConfig
// with the reference's public interface (src/slam_frontend.h:58-204) on top of the
// C ABI in include/vsf.h.  Same member names, argument meaning and call sequence as
// the reference; the matching / stereo / triangulation work runs in libvsf_cuda.so
// (one fused device pass per frame), the std::sort + best_percent cut + FeatureMatch
// book-keeping of GetFeatureMatches stay on the host exactly as the reference has them.
//
// Deliberate differences (each one a reference quirk documented in SURVEY.md 8(a)):
//  * feature extraction is the caller's input producer: either call ObserveFeatures
//    with keypoints + descriptors, or install a FeatureExtractor for ObserveImage;
//  * the fundamental matrix is a config input defaulting to the textbook matrix (the
//    reference's own construction reads out of bounds, src/slam_frontend.cc:640-642);
//  * the adaptive stereo threshold lives in the vsf context, not in a process global;
//  * features[i].point3d for i >= points.size() is NaN (the reference reads out of
//    bounds there, src/slam_frontend.cc:439-441);
//  * errors throw std::runtime_error (the reference aborts through glog CHECK).
#ifndef VSF_SLAM_FRONTEND_H_
#define VSF_SLAM_FRONTEND_H_

#include <cstdint>
#include <deque>
#include <functional>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "cv_shim.h"
#include "slam_types.h"
#include "vsf.h"

namespace slam {

// Pinhole intrinsics + radial/tangential distortion (src/slam_frontend.h:42-55).
struct CameraIntrinsics {
  float k1, k2, k3;
  float p1, p2;
  float fx, fy, cx, cy;
};

struct FrontendConfig {
 public:
  enum class DescriptorExtractorType { AKAZE, ORB, BRISK, SURF, SIFT, FREAK };
  FrontendConfig();
  // Declared but never defined in the reference (src/slam_frontend.h:69).  Here: a
  // plain "key value" text file overriding the scalar defaults; unknown keys throw.
  void Load(const std::string& path);
  // Re-derive camera matrices, projections and the fundamental matrix from the
  // intrinsics / stereo extrinsics (call after editing them).
  void UpdateDerived();

  bool debug_images_;
  DescriptorExtractorType descriptor_extract_type_;
  float best_percent_;
  float nn_match_ratio_;
  float min_odom_translation;
  float min_odom_rotation;
  uint32_t min_vision_matches;
  uint32_t frame_life_;
  int bf_matcher_param_;                      // cv::NORM_HAMMING
  CameraIntrinsics intrinsics_left, intrinsics_right;
  // right camera pose: X_right = R * X_left + t, row-major 3x4 [R | t]
  float stereo_extrinsics[12];
  // derived (row-major float32, the layouts cv::Mat(3,3/3,4,CV_32F) would have)
  float camera_matrix_left[9], camera_matrix_right[9];
  float distortion_coeffs_left[5], distortion_coeffs_right[5];
  float projection_left[12], projection_right[12];
  float fundamental[9];                       // convention x_left^T F x_right
  float left_cam_to_robot_translation[3];
  float left_cam_to_robot_rotation[9];
  // B200 path parameters (no reference counterpart)
  int cuda_device;
  int max_features;        // capacity of one frame
  int descriptor_bytes;    // 32 ORB, 61 AKAZE, 64 BRISK/FREAK
  bool exact_std_sort;     // true: std::sort on the host like the reference
  // The reference's arithmetic sets the adaptive stereo threshold to 0/0 + 2 = NaN after a frame
  // without stereo matches, which empties the next frame (src/slam_frontend.cc:392-394); false
  // keeps the previous threshold instead (VSF_OPT_HOLD_THRESHOLD_ON_EMPTY).
  bool strict_reference_threshold;
  // float32 summation order of the epipolar residual's dot products: 0 = Eigen 3.2's
  // (c0 + c1) + c2, 1 = Eigen 3.3's c0 + (c1 + c2) (VSF_OPT_RESIDUAL_ORDER)
  int residual_order;
};

class Frame {
 public:
  Frame(const std::vector<cv::KeyPoint>& keypoints, const cv::Mat& descriptors, uint64_t frame_ID);
  Frame() : frame_ID_(0) {}
  uint64_t frame_ID_;
  std::vector<cv::KeyPoint> keypoints_;
  std::vector<bool> is_initial_;
  std::vector<int64_t> initial_ids_;
  cv::Mat descriptors_;
  std::unordered_map<uint64_t, std::pair<uint64_t, uint64_t>> initial_appearances;
  cv::Mat debug_image_;
};

class Frontend {
 public:
  using FeatureExtractor =
      std::function<void(const cv::Mat& image, std::vector<cv::KeyPoint>* keypoints, cv::Mat* descriptors)>;

  explicit Frontend(const std::string& config_path);
  explicit Frontend(const FrontendConfig& config);
  ~Frontend();
  Frontend(const Frontend&) = delete;
  Frontend& operator=(const Frontend&) = delete;

  // The reference's detectAndCompute (src/slam_frontend.cc:266-280) stays an input
  // producer: plug it in here to use ObserveImage.
  void SetFeatureExtractor(FeatureExtractor extractor) { extractor_ = std::move(extractor); }

  // Returns true iff a new SLAM node was added (src/slam_frontend.cc:400-472).
  bool ObserveImage(const cv::Mat& left_image, const cv::Mat& right_image, double time);
  // Same, entered after feature extraction.
  bool ObserveFeatures(const std::vector<cv::KeyPoint>& left_keypoints, const cv::Mat& left_descriptors,
                       const std::vector<cv::KeyPoint>& right_keypoints, const cv::Mat& right_descriptors,
                       double time);
  // Pipelined form for frame streams (the bag loop of src/slam_frontend_main.cc:236-328):
  // SubmitFeatures gates on the odometry exactly like ObserveFeatures, snapshots the odometry
  // state, enqueues the frame on the device (vsf_observe_submit) and returns; CollectFeatures
  // finishes the OLDEST submitted frame on the host (sort + cut, FeatureMatch book-keeping, node
  // assembly) in submission order.  Up to MaxInFlight() frames may be pending; results are
  // identical to calling ObserveFeatures frame by frame.
  bool SubmitFeatures(const std::vector<cv::KeyPoint>& left_keypoints, const cv::Mat& left_descriptors,
                      const std::vector<cv::KeyPoint>& right_keypoints, const cv::Mat& right_descriptors,
                      double time);
  bool CollectFeatures();                 // false when nothing is pending
  int InFlight() const { return int(pending_.size()); }
  static int MaxInFlight() { return VSF_OBSERVE_DEPTH; }
  // Sharded sequences: this instance owns the poses from `first` on of a longer
  // sequence.  Frames are numbered from halo_first; frames below `first` are the halo - they
  // rebuild the sliding window and the adaptive stereo threshold and produce no nodes or factors.
  // halo_first = ShardHaloStart(first, frame_life_) reproduces the unsharded state exactly:
  // frame_life_ frames for the window (src/slam_frontend.cc:424-434) plus one whose only purpose
  // is its `mean + 2` threshold (:392-394).  The caller feeds ObserveOdometry with the sequence's
  // FIRST odometry message before the halo's (init_odom_* are relative to it, :252-256).
  void StartShard(uint64_t halo_first, uint64_t first);
  static uint64_t ShardHaloStart(uint64_t first, uint32_t frame_life);
  void ObserveOdometry(const Eigen::Vector3f& translation, const Eigen::Quaternionf& rotation,
                       double timestamp);
  void GetSLAMProblem(slam_types::SLAMProblem* problem) const;
  int GetNumPoses();
  FrontendConfig GetConfig() { return config_; }

  // The reference's private helpers; public so that tests and integrators can reach
  // the operator boundary directly.
  std::vector<cv::DMatch> GetMatches(const Frame& frame_query, const Frame& frame_train,
                                     double nn_match_ratio);
  slam_types::VisionFactor* GetFeatureMatches(Frame* past_frame_ptr, Frame* curr_frame_ptr);
  void UndistortFeaturePoints(std::vector<slam_types::VisionFeature>* features);
  float GetStereoAmbigConstraint();
  void SetStereoAmbigConstraint(float v);

  // Little-endian ROS1 serialisation of the SLAMProblem message (msg/SLAMProblem.msg
  // through src/slam_to_ros.h:36-124), so the output stays byte-compatible without ROS.
  static std::vector<uint8_t> SerializeSLAMProblem(const slam_types::SLAMProblem& problem);

 private:
  bool OdomCheck();
  void AddOdometryFactor();
  // sort + best_percent cut + FeatureMatch/is_initial_ book-keeping on a query-ordered
  // match list (the part of GetFeatureMatches after GetMatches).
  slam_types::VisionFactor FinishFeatureMatches(std::vector<cv::DMatch>* matches, float best_percent,
                                                Frame* past_frame, Frame* curr_frame,
                                                std::vector<cv::DMatch>* sorted_out);
  void Check(int rc, const char* what);
  void ApplyOptions();

  bool odom_initialized_;
  Eigen::Vector3f init_odom_translation_;
  Eigen::Quaternionf init_odom_rotation_;
  Eigen::Vector3f prev_odom_translation_;
  Eigen::Quaternionf prev_odom_rotation_;
  Eigen::Vector3f odom_translation_;
  Eigen::Quaternionf odom_rotation_;
  double odom_timestamp_;
  FrontendConfig config_;
  vsf_ctx* ctx_;             // replaces cv::Ptr<cv::BFMatcher> matcher_
  uint64_t curr_frame_ID_;
  uint64_t first_output_ID_;   // StartShard: frames below this id produce no nodes / factors
  std::vector<Frame> frame_list_;
  FeatureExtractor extractor_;
  std::vector<slam_types::VisionFactor> vision_factors_;
  std::vector<slam_types::SLAMNode> nodes_;
  std::vector<slam_types::OdometryFactor> odometry_factors_;
  // a submitted frame waiting for its device results
  struct Pending {
    std::vector<cv::KeyPoint> left_keypoints, right_keypoints;
    cv::Mat left_descriptors, right_descriptors;
    uint64_t frame_ID;
    Eigen::Vector3f odom_translation, prev_odom_translation;
    Eigen::Quaternionf odom_rotation, prev_odom_rotation;
    double odom_timestamp;
  };
  std::deque<Pending> pending_;
  // scratch for vsf_observe_collect
  std::vector<float> xy_undist_;
  std::vector<int32_t> kept_left_, kept_right_;
  std::vector<uint64_t> frame_ids_;
  std::vector<int> window_counts_;
  std::vector<vsf_dmatch> window_matches_, tri_matches_;
  std::vector<float> tri_X4_;
};

}  // namespace slam

#endif  // VSF_SLAM_FRONTEND_H_
