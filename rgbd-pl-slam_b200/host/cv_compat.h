// cv_compat.h — the few OpenCV / OpenCV-contrib / Eigen types the reference's front-end signatures
// mention, so that the drop-in classes compile in an environment without those libraries (this
// container has no OpenCV or Eigen C++ headers).  Define PLSLAM_WITH_OPENCV to use the real
// headers instead: the class signatures below are then literally the reference's.
#pragma once
#ifdef PLSLAM_WITH_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#include <opencv2/line_descriptor/descriptor.hpp>
#include <eigen3/Eigen/Core>
#else
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {

struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float a, float b) : x(a), y(b) {} };
struct Point { int x = 0, y = 0; Point() {} Point(int a, int b) : x(a), y(b) {} };
typedef Point Point2i;

// cv::KeyPoint: 28 bytes, same field order (pt, size, angle, response, octave, class_id)
struct KeyPoint {
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct DMatch {
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = 0;
  DMatch() {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};

// single-channel matrix with shared ownership (enough of cv::Mat for images, descriptor blocks and the small CV_32F pose
// matrices the matchers read: CV_8U and CV_32F)
class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;  // bytes per row
  uint8_t* data = nullptr;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* ext, size_t st = 0) : rows(r), cols(c), step(st ? st : (size_t)c * esz(type)), data((uint8_t*)ext), type_(type) {}
  void create(int r, int c, int type) {
    if (r == rows && c == cols && type == type_ && own_) return;
    rows = r; cols = c; type_ = type; step = (size_t)c * esz(type);
    const size_t bytes = (size_t)r * step;
    own_ = std::shared_ptr<uint8_t>(new uint8_t[bytes > 0 ? bytes : 1], std::default_delete<uint8_t[]>());
    data = own_.get();
  }
  void release() { rows = cols = 0; step = 0; data = nullptr; own_.reset(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return type_; }
  size_t elemSize() const { return esz(type_); }
  uint8_t* ptr(int r = 0) { return data + (size_t)r * step; }
  const uint8_t* ptr(int r = 0) const { return data + (size_t)r * step; }
  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
  template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
  template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
  template <typename T> T& at(int i) { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }
  template <typename T> const T& at(int i) const { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }
  Mat row(int r) const { Mat m; m.rows = 1; m.cols = cols; m.step = step; m.type_ = type_; m.data = data + (size_t)r * step; m.own_ = own_; return m; }
  Mat clone() const { Mat m(rows, cols, type_); for (int r = 0; r < rows; ++r) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * esz(type_)); return m; }
 private:
  static size_t esz(int type) { return type == 5 ? 4 : 1; }
  int type_ = 0;
  std::shared_ptr<uint8_t> own_;
};

// InputArray / OutputArray as used by ORBextractor::operator()
class _InputArray {
 public:
  _InputArray() {}
  _InputArray(const Mat& m) : m_(&m) {}
  Mat getMat() const { return m_ ? *m_ : Mat(); }
  bool empty() const { return !m_ || m_->empty(); }
 private:
  const Mat* m_ = nullptr;
};
class _OutputArray {
 public:
  _OutputArray(Mat& m) : m_(&m) {}
  void create(int r, int c, int t) const { m_->create(r, c, t); }
  void release() const { m_->release(); }
  Mat getMat() const { return *m_; }
  Mat& getMatRef() const { return *m_; }
 private:
  Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

namespace line_descriptor {
// cv::line_descriptor::KeyLine (OpenCV-contrib), same field order as plslam_keyline_t
struct KeyLine {
  float angle = 0;
  int class_id = -1;
  int octave = 0;
  Point2f pt;
  float response = 0;
  float size = 0;
  float startPointX = 0, startPointY = 0, endPointX = 0, endPointY = 0;
  float sPointInOctaveX = 0, sPointInOctaveY = 0, ePointInOctaveX = 0, ePointInOctaveY = 0;
  float lineLength = 0;
  int numOfPixels = 0;
};
static_assert(sizeof(KeyLine) == 68, "KeyLine layout");
}  // namespace line_descriptor
}  // namespace cv

namespace Eigen {
struct Vector3d {
  double v[3] = {0, 0, 0};
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
};
}  // namespace Eigen
#endif
