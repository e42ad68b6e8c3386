// Minimal stand-ins for the OpenCV / Eigen types that appear in the public
// interface of slam::Frontend / slam::Frame / slam_types (reference:
// src/slam_frontend.h:100-142, src/slam_types.h:60-187).  The image this project
// is built in has neither OpenCV C++ headers nor Eigen; where they exist
// (__has_include) the real types are used instead and these shims vanish.
// Layouts match the originals: cv::KeyPoint 28 bytes, cv::DMatch 16 bytes,
// cv::Mat rows of `step` bytes.
#ifndef VSF_CV_SHIM_H_
#define VSF_CV_SHIM_H_

#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#if defined(__has_include)
#if __has_include(<opencv2/core.hpp>) && !defined(VSF_FORCE_SHIMS)
#define VSF_HAVE_OPENCV 1
#endif
#if __has_include(<eigen3/Eigen/Dense>) && !defined(VSF_FORCE_SHIMS)
#define VSF_HAVE_EIGEN 1
#endif
#endif

#ifdef VSF_HAVE_OPENCV
#include <opencv2/core.hpp>
#else
namespace cv {

struct Point2f {
  float x = 0.f, y = 0.f;
  Point2f() {}
  Point2f(float x_, float y_) : x(x_), y(y_) {}
};

struct KeyPoint {
  Point2f pt;
  float size = 0.f, angle = -1.f, response = 0.f;
  int octave = 0, class_id = -1;
  KeyPoint() {}
  KeyPoint(float x, float y, float size_) : pt(x, y), size(size_) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct DMatch {
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = 3.4028235e38f;
  DMatch() {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
  DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
  bool operator<(const DMatch& m) const { return distance < m.distance; }   // distance only
};
static_assert(sizeof(DMatch) == 16, "cv::DMatch layout");

}  // namespace cv
// type codes are preprocessor macros in OpenCV (interface.h), not members of namespace cv
#ifndef CV_8U
#define CV_8U 0
#define CV_32F 5
#endif
namespace cv {
enum NormTypes { NORM_L2 = 4, NORM_HAMMING = 6 };

// Row-major byte matrix with shared ownership — the subset of cv::Mat the path uses
// (descriptor matrices N x W of CV_8U, grey images).
class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  uint8_t* data = nullptr;
  Mat() {}
  Mat(int r, int c, int /*type*/) { create(r, c); }   // like cv::Mat: the type is not optional
  Mat(int r, int c, int /*type*/, void* external, size_t step_ = 0)
      : rows(r), cols(c), step(step_ ? step_ : size_t(c)), data(static_cast<uint8_t*>(external)) {}
  void create(int r, int c) {
    rows = r;
    cols = c;
    step = size_t(c);
    store_ = std::shared_ptr<uint8_t>(new uint8_t[size_t(r) * c + 1], std::default_delete<uint8_t[]>());
    data = store_.get();
  }
  bool empty() const { return rows == 0 || cols == 0 || data == nullptr; }
  uint8_t* ptr(int r) { return data + size_t(r) * step; }
  const uint8_t* ptr(int r) const { return data + size_t(r) * step; }
  Mat clone() const {
    Mat m;
    if (!empty()) {
      m.create(rows, cols);
      for (int r = 0; r < rows; ++r) std::memcpy(m.ptr(r), ptr(r), size_t(cols));
    }
    return m;
  }

 private:
  std::shared_ptr<uint8_t> store_;
};

}  // namespace cv
#endif  // VSF_HAVE_OPENCV

#ifdef VSF_HAVE_EIGEN
#include <eigen3/Eigen/Dense>
#include <eigen3/Eigen/Geometry>
#else
namespace Eigen {

struct Vector2f {
  float v[2] = {0.f, 0.f};
  Vector2f() {}
  Vector2f(float x, float y) { v[0] = x; v[1] = y; }
  float x() const { return v[0]; }
  float y() const { return v[1]; }
  float& x() { return v[0]; }
  float& y() { return v[1]; }
};

struct Vector3f {
  float v[3] = {0.f, 0.f, 0.f};
  Vector3f() {}
  Vector3f(float x, float y, float z) { v[0] = x; v[1] = y; v[2] = z; }
  float x() const { return v[0]; }
  float y() const { return v[1]; }
  float z() const { return v[2]; }
  float norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
  Vector3f operator-(const Vector3f& o) const { return Vector3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
  Vector3f operator+(const Vector3f& o) const { return Vector3f(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
  Vector3f operator/(float s) const { return Vector3f(v[0] / s, v[1] / s, v[2] / s); }
};

// Unit quaternion, Eigen's (w, x, y, z) constructor order and x,y,z,w storage order.
struct Quaternionf {
  float qx = 0.f, qy = 0.f, qz = 0.f, qw = 1.f;
  Quaternionf() {}
  Quaternionf(float w, float x, float y, float z) : qx(x), qy(y), qz(z), qw(w) {}
  static Quaternionf Identity() { return Quaternionf(1.f, 0.f, 0.f, 0.f); }
  float x() const { return qx; }
  float y() const { return qy; }
  float z() const { return qz; }
  float w() const { return qw; }
  Quaternionf inverse() const {
    const float n2 = qx * qx + qy * qy + qz * qz + qw * qw;
    return Quaternionf(qw / n2, -qx / n2, -qy / n2, -qz / n2);
  }
  Quaternionf operator*(const Quaternionf& b) const {
    return Quaternionf(qw * b.qw - qx * b.qx - qy * b.qy - qz * b.qz,
                       qw * b.qx + qx * b.qw + qy * b.qz - qz * b.qy,
                       qw * b.qy - qx * b.qz + qy * b.qw + qz * b.qx,
                       qw * b.qz + qx * b.qy - qy * b.qx + qz * b.qw);
  }
  Vector3f operator*(const Vector3f& p) const {   // rotate: q * p * q^-1 for unit q
    const float ux = qx, uy = qy, uz = qz;
    const float cx = uy * p.v[2] - uz * p.v[1], cy = uz * p.v[0] - ux * p.v[2], cz = ux * p.v[1] - uy * p.v[0];
    const float tx = 2.f * cx, ty = 2.f * cy, tz = 2.f * cz;
    return Vector3f(p.v[0] + qw * tx + (uy * tz - uz * ty), p.v[1] + qw * ty + (uz * tx - ux * tz),
                    p.v[2] + qw * tz + (ux * ty - uy * tx));
  }
  float angularDistance(const Quaternionf& o) const {
    const Quaternionf d = (*this) * o.inverse();
    const float vn = std::sqrt(d.qx * d.qx + d.qy * d.qy + d.qz * d.qz);
    return 2.f * std::atan2(vn, std::fabs(d.qw));
  }
};

}  // namespace Eigen
#endif  // VSF_HAVE_EIGEN

#endif  // VSF_CV_SHIM_H_
