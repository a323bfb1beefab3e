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

// ---- the few pieces of highgui / imgcodecs / utility the CLI shims touch (stand-in only) ------------------------
// imread / imwrite of the stand-in speak binary PPM (P6) so that the CLI shims can be run end to end without OpenCV;
// with the real headers they are cv::imread / cv::imwrite and read what OpenCV reads.
#include <chrono>
#include <cstdio>
#include <iostream>
#include <map>
#include <sstream>
#define CV_VERSION "stand-in"
namespace cv {
typedef std::string String;
enum { IMREAD_COLOR = 1 };
inline Mat imread(const String& path, int = IMREAD_COLOR) {
  Mat m;
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return m;
  int w = 0, h = 0, mx = 0;
  char magic[3] = {0, 0, 0};
  if (std::fscanf(f, "%2s %d %d %d", magic, &w, &h, &mx) == 4 && magic[0] == 'P' && magic[1] == '6' && mx == 255 && w > 0 && h > 0) {
    std::fgetc(f);
    m.create(h, w, CV_8UC3);
    std::vector<unsigned char> row((size_t)w * 3);
    for (int y = 0; y < h; y++) {
      if (std::fread(row.data(), 1, row.size(), f) != row.size()) { m = Mat(); break; }
      for (int x = 0; x < w; x++) {  // PPM is RGB, cv::Mat is BGR
        m.data[(size_t)y * m.step + 3 * x] = row[3 * x + 2];
        m.data[(size_t)y * m.step + 3 * x + 1] = row[3 * x + 1];
        m.data[(size_t)y * m.step + 3 * x + 2] = row[3 * x];
      }
    }
  }
  std::fclose(f);
  return m;
}
inline bool imwrite(const String& path, const Mat& m) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f || m.empty() || m.channels() != 3) { if (f) std::fclose(f); return false; }
  std::fprintf(f, "P6\n%d %d\n255\n", m.cols, m.rows);
  std::vector<unsigned char> row((size_t)m.cols * 3);
  for (int y = 0; y < m.rows; y++) {
    for (int x = 0; x < m.cols; x++) {
      row[3 * x] = m.data[(size_t)y * m.step + 3 * x + 2];
      row[3 * x + 1] = m.data[(size_t)y * m.step + 3 * x + 1];
      row[3 * x + 2] = m.data[(size_t)y * m.step + 3 * x];
    }
    std::fwrite(row.data(), 1, row.size(), f);
  }
  std::fclose(f);
  return true;
}
inline long long getTickCount() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline double getTickFrequency() { return 1e9; }
// `-key=value` / positional arguments, the subset of cv::CommandLineParser the reference mains use
class CommandLineParser {
 public:
  CommandLineParser(int argc, const char* const argv[], const String& keys) : keys_(keys) {
    for (int i = 1; i < argc; i++) {
      String a = argv[i];
      if (a.size() > 1 && a[0] == '-' && !(a[1] >= '0' && a[1] <= '9')) {
        size_t b = a.find_first_not_of('-'), e = a.find('=');
        String name = a.substr(b, e == String::npos ? String::npos : e - b);
        opt_[name] = e == String::npos ? "true" : a.substr(e + 1);
      } else pos_.push_back(a);
    }
  }
  void about(const String& s) { about_ = s; }
  bool has(const String& name) const { return opt_.count(name) || (name == "help" && (opt_.count("h") || opt_.count("usage") || opt_.count("?"))); }
  template <class T> T get(const String& name) const {
    T v = T();
    std::map<String, String>::const_iterator it = opt_.find(name);
    if (it != opt_.end()) { std::istringstream ss(it->second); ss >> v; }
    return v;
  }
  template <class T> T get(int index) const {
    T v = T();
    if (index < (int)pos_.size()) { std::istringstream ss(pos_[index]); ss >> v; }
    return v;
  }
  bool check() const { return true; }
  void printErrors() const {}
  void printMessage() const { std::cout << about_ << std::endl << keys_ << std::endl; }
 private:
  String keys_, about_;
  std::map<String, String> opt_;
  std::vector<String> pos_;
};
}  // namespace cv
