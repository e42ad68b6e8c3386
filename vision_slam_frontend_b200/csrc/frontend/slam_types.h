// Output data model of the frontend — same names, fields and meaning as the
// reference's src/slam_types.h:60-187 (VisionFeature, FeatureMatch, VisionFactor,
// RobotPose, OdometryFactor, SLAMNode, SLAMProblem).  Restated, not copied: the
// Eigen types come from cv_shim.h when Eigen is not installed.
#ifndef VSF_SLAM_TYPES_H_
#define VSF_SLAM_TYPES_H_

#include <cstdint>
#include <vector>

#include "cv_shim.h"

namespace slam_types {

struct VisionFeature {
  uint64_t feature_idx = 0;   // index inside its node's feature vector
  Eigen::Vector2f pixel;      // pixel location (undistorted once the node is built)
  Eigen::Vector3f point3d;    // triangulated position in the left-camera frame
  VisionFeature() {}
  VisionFeature(uint64_t idx, const Eigen::Vector2f& p, const Eigen::Vector3f& p3)
      : feature_idx(idx), pixel(p), point3d(p3) {}
};

struct FeatureMatch {
  uint64_t feature_idx_initial = 0;   // feature id in the earlier pose
  uint64_t feature_idx_current = 0;   // feature id in the later pose
  FeatureMatch() {}
  FeatureMatch(uint64_t fid_initial, uint64_t fid_current)
      : feature_idx_initial(fid_initial), feature_idx_current(fid_current) {}
};

struct VisionFactor {
  uint64_t pose_idx_initial = 0;
  uint64_t pose_idx_current = 0;
  std::vector<FeatureMatch> feature_matches;
  VisionFactor() {}
  VisionFactor(uint64_t pose_initial, uint64_t pose_current, const std::vector<FeatureMatch>& m)
      : pose_idx_initial(pose_initial), pose_idx_current(pose_current), feature_matches(m) {}
};

struct RobotPose {
  Eigen::Vector3f loc;
  Eigen::Quaternionf angle;
  RobotPose() {}
  RobotPose(const Eigen::Vector3f& l, const Eigen::Quaternionf& a) : loc(l), angle(a) {}
};

struct OdometryFactor {
  uint64_t pose_i = 0, pose_j = 0;
  Eigen::Vector3f translation;
  Eigen::Quaternionf rotation;
  OdometryFactor() {}
  OdometryFactor(uint64_t i, uint64_t j, const Eigen::Vector3f& t, const Eigen::Quaternionf& r)
      : pose_i(i), pose_j(j), translation(t), rotation(r) {}
};

struct SLAMNode {
  uint64_t node_idx = 0;
  double timestamp = 0.0;
  RobotPose pose;
  std::vector<VisionFeature> features;
  SLAMNode() {}
  SLAMNode(uint64_t idx, double t, const RobotPose& p, const std::vector<VisionFeature>& f)
      : node_idx(idx), timestamp(t), pose(p), features(f) {}
};

struct SLAMProblem {
  std::vector<SLAMNode> nodes;
  std::vector<VisionFactor> vision_factors;
  std::vector<OdometryFactor> odometry_factors;
  SLAMProblem() {}
  SLAMProblem(const std::vector<SLAMNode>& n, const std::vector<VisionFactor>& v,
              const std::vector<OdometryFactor>& o)
      : nodes(n), vision_factors(v), odometry_factors(o) {}
};

}  // namespace slam_types

#endif  // VSF_SLAM_TYPES_H_
