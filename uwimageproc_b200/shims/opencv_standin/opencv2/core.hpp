// Minimal stand-in for the parts of <opencv2/core.hpp> the drop-in shims touch.  OpenCV C++ is not in
// this image, so the shims are type-checked (and their host logic unit-tested) against
// this header; where OpenCV exists the real headers are used instead (-I order in check.sh).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32FC1 5
#define UWIP_OPENCV_STANDIN 1

namespace cv {
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };

class Mat {
 public:
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  struct Step { size_t v = 0; operator size_t() const { return v; } } step;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    step.v = (size_t)c * elemSize();
    buf_ = std::shared_ptr<unsigned char>((unsigned char*)std::calloc((size_t)r * step.v + 1, 1), std::free);
    data = buf_.get();
  }
  int type() const { return type_; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t elemSize() const { return (size_t)channels() * ((type_ & 7) == CV_32F ? 4 : 1); }
  bool empty() const { return data == nullptr; }
  bool isContinuous() const { return step.v == (size_t)cols * elemSize(); }
  Size size() const { return Size(cols, rows); }
  template <class T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step.v + (size_t)c * sizeof(T)); }
  template <class T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step.v + (size_t)c * sizeof(T)); }

 private:
  int type_ = 0;
  std::shared_ptr<unsigned char> buf_;  // header copies share the pixel buffer, like cv::Mat
};

namespace cuda {
class GpuMat {
 public:
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;  // device pointer
  size_t step = 0;
  int type() const { return type_; }
  bool isContinuous() const { return step == (size_t)cols; }
  int type_ = 0;
};
}  // namespace cuda
}  // namespace cv
