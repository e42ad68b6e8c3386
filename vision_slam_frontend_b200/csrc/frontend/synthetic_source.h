// Frame source for the drop-in when there is no rosbag: a deterministic synthetic stereo
// sequence at the point where the reference's input producers hand over - keypoints and
// descriptors of the left and right image of every pose plus the odometry message that gates
// it (the bag loop of src/slam_frontend_main.cc:236-328 feeding ObserveOdometry /
// ObserveImage).  Counter-based: any pose can be produced on any rank without producing the
// ones before it, which is what lets a sequence be sharded into pose ranges.
#ifndef VSF_SYNTHETIC_SOURCE_H_
#define VSF_SYNTHETIC_SOURCE_H_

#include <cstdint>
#include <vector>

#include "cv_shim.h"
#include "slam_frontend.h"

namespace slam {

struct SyntheticStereoConfig {
  int features = 2000;             // per image
  int image_width = 960, image_height = 600;   // pixel range of the left keypoints
  int landmark_stride = 200;       // new landmarks per pose: pose p sees landmarks [stride*p, stride*p + features)
  int temporal_flips = 8;          // descriptor bits flipped (up to) between a landmark's code and its observation
  int stereo_flips = 16;           // and between the left and the right observation
  float pixel_noise = 0.5f;        // Gaussian sigma, pixels
  float outlier_fraction = 0.10f;  // right features unrelated to any left feature
  float bad_geometry_fraction = 0.05f;   // right features 30 px off their epipolar line
  float depth_min = 4.f, depth_max = 80.f;
  float step = 0.5f;               // metres the robot moves per pose (odometry)
  uint64_t seed = 1;
};

class SyntheticStereoSource {
 public:
  SyntheticStereoSource(const FrontendConfig& rig, const SyntheticStereoConfig& cfg);
  // Keypoints + descriptors (rig.descriptor_bytes wide, CV_8U) of both images of `pose`.
  void Frame(uint64_t pose, std::vector<cv::KeyPoint>* left_keypoints, cv::Mat* left_descriptors,
             std::vector<cv::KeyPoint>* right_keypoints, cv::Mat* right_descriptors) const;
  // The odometry message preceding that frame.
  void Odometry(uint64_t pose, Eigen::Vector3f* translation, Eigen::Quaternionf* rotation, double* timestamp) const;
  const SyntheticStereoConfig& config() const { return cfg_; }

 private:
  FrontendConfig rig_;
  SyntheticStereoConfig cfg_;
};

// The rig of the synthetic runs: the reference's PointGrey calibration with the fundamental
// matrix rescaled so that the epipolar residual is (about) a distance in pixels and the adaptive `mean + 2` threshold (src/slam_frontend.cc:392-394) actually
// separates the 30-pixel outliers the source plants.
FrontendConfig SyntheticRig(int device, int features, int descriptor_bytes, int frame_life, bool exact_std_sort);

// What one rank of a sharded run contributes to the SLAMProblem message: the bodies of its three
// arrays in ROS1 wire format (without the uint32 length prefixes) and their element counts.
struct SLAMProblemPiece {
  uint32_t n_nodes = 0, n_vision_factors = 0, n_odometry_factors = 0;
  std::vector<uint8_t> nodes, vision_factors, odometry_factors;
  std::vector<uint8_t> Pack() const;                       // one blob for the gather
  static SLAMProblemPiece Unpack(const uint8_t* data, size_t n);
};

// Run poses [first, last) of the source's sequence through a fresh Frontend built from `rig`
// (halo included, Frontend::StartShard), `in_flight` frames pipelined (1 = blocking calls).
// host_us (optional, 3 values): mean microseconds per frame spent in the frame source, in
// SubmitFeatures and in CollectFeatures.
SLAMProblemPiece RunSequenceShard(const FrontendConfig& rig, const SyntheticStereoSource& source, uint64_t first,
                                  uint64_t last, int in_flight, double* host_us = nullptr);
// The whole message from the ranks' pieces, in rank order: byte-identical to
// Frontend::SerializeSLAMProblem of an unsharded run over the same poses.
std::vector<uint8_t> MergeSLAMProblemPieces(const std::vector<SLAMProblemPiece>& pieces);

}  // namespace slam

#endif  // VSF_SYNTHETIC_SOURCE_H_
