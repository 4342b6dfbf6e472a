// Minimal stand-in for <opencv2/core/core.hpp>, just enough to compile the reference's vendored DBoW2
// sources (Thirdparty/DBoW2/DBoW2/*.cpp, TemplatedVocabulary.h) UNMODIFIED into oracle/_ref/ in an
// environment without OpenCV.  Only 8-bit row descriptors are ever stored in these cv::Mat objects.
// cv::FileStorage / cv::FileNode exist syntactically for the YAML save/load members, which are never called.
#pragma once
// the real header pulls these in transitively; the DBoW2 sources rely on that
#include <math.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {

class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  uint8_t* data = nullptr;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    const size_t es = type == CV_32F ? 4 : 1;
    rows = r; cols = c; step = (size_t)c * es;
    // zero-initialised (value-init): the reference reads a never-written descriptor for the phantom node a trailing
    // empty line of the vocabulary file creates (TemplatedVocabulary.h:1401-1443); zeros make that defined here
    own_ = std::shared_ptr<uint8_t>(new uint8_t[(size_t)r * c * es > 0 ? (size_t)r * c * es : 1](), std::default_delete<uint8_t[]>());
    data = own_.get();
  }
  static Mat zeros(int r, int c, int t) { Mat m(r, c, t); std::memset(m.data, 0, (size_t)r * c); return m; }
  bool empty() const { return !data || rows == 0 || cols == 0; }
  Mat clone() const { Mat m(rows, cols, CV_8U); for (int r = 0; r < rows; ++r) std::memcpy(m.data + r * m.step, data + r * step, cols); return m; }
  Mat row(int r) const { Mat m; m.rows = 1; m.cols = cols; m.step = step; m.data = data + (size_t)r * step; m.own_ = own_; return m; }
  void release() { rows = cols = 0; data = nullptr; own_.reset(); }
  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
  template <typename T> T& at(int r, int c) { return reinterpret_cast<T*>(data + (size_t)r * step)[c]; }
  template <typename T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(data + (size_t)r * step)[c]; }
 private:
  std::shared_ptr<uint8_t> own_;
};

class FileNode {
 public:
  enum { NONE = 0, SEQ = 5, MAP = 6 };
  FileNode operator[](const char*) const { std::abort(); }
  FileNode operator[](const std::string&) const { std::abort(); }
  FileNode operator[](int) const { std::abort(); }
  template <typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  operator T() const { std::abort(); }
  operator std::string() const { std::abort(); }
  size_t size() const { return 0; }
  int type() const { return NONE; }
};
class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const char*, int) {}
  FileStorage(const std::string&, int) {}
  bool isOpened() const { return false; }
  void release() {}
  FileNode operator[](const char*) const { std::abort(); }
  FileNode operator[](const std::string&) const { std::abort(); }
  template <typename T> FileStorage& operator<<(const T&) { return *this; }
};

}  // namespace cv
