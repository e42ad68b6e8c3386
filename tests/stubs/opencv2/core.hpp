// TEST STUB, not OpenCV: the declarations of <opencv2/core.hpp> that csrc/frontend touches, with
// the real library's shapes (type codes are macros, cv::Mat takes (rows, cols, type[, data, step]),
// Mat::step is a MatStep, KeyPoint / DMatch layouts), so that a syntax-only compile catches code
// that only works against cv_shim.h.  tests/test_frontend_real_headers.py.
#pragma once
#include <cstddef>
#define CV_8U 0
#define CV_32F 5
namespace cv {
typedef unsigned char uchar;
template <typename T>
class Point_ {
 public:
  Point_();
  Point_(T x_, T y_);
  T x, y;
};
typedef Point_<float> Point2f;
class KeyPoint {
 public:
  KeyPoint();
  KeyPoint(Point2f pt_, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1);
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1);
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
};
class DMatch {
 public:
  DMatch();
  DMatch(int q, int t, float d);
  DMatch(int q, int t, int i, float d);
  int queryIdx, trainIdx, imgIdx;
  float distance;
  bool operator<(const DMatch& m) const;
};
struct MatStep {
  MatStep();
  explicit MatStep(size_t s);
  operator size_t() const;
  MatStep& operator=(size_t s);
  size_t* p;
  size_t buf[2];
};
class Mat {
 public:
  enum { AUTO_STEP = 0 };
  Mat();
  Mat(int rows, int cols, int type);
  Mat(int rows, int cols, int type, void* data, size_t step = AUTO_STEP);
  Mat(const Mat& m);
  ~Mat();
  Mat& operator=(const Mat& m);
  Mat clone() const;
  void create(int rows, int cols, int type);
  bool empty() const;
  uchar* ptr(int i0 = 0);
  const uchar* ptr(int i0 = 0) const;
  int flags, dims, rows, cols;
  uchar* data;
  MatStep step;
};
enum NormTypes { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6, NORM_HAMMING2 = 7 };
}  // namespace cv
