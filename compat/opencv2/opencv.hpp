// Minimal stand-in for <opencv2/opencv.hpp> (standalone builds only; OpenCV C++ headers are absent in the authoring image).
// cv::Mat here is a non-owning or owning 8-bit single-channel image view: rows / cols / data / step / empty() / ptr().
#ifndef AIRFE_COMPAT_OPENCV
#define AIRFE_COMPAT_OPENCV
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8UC1 0
#define CV_8U 0
typedef unsigned char uchar;

namespace cv {
class Mat {
 public:
  Mat() : rows(0), cols(0), data(nullptr), step(0) {}
  Mat(int r, int c, int /*type*/) : rows(r), cols(c), step((size_t)c), own_(new uchar[(size_t)r * c]()) { data = own_.get(); }
  Mat(int r, int c, int /*type*/, void* d, size_t s = 0) : rows(r), cols(c), data((uchar*)d), step(s ? s : (size_t)c) {}
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  uchar* ptr(int r = 0) { return data + (size_t)r * step; }
  const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
  int channels() const { return 1; }
  int rows, cols;
  uchar* data;
  size_t step;
 private:
  std::shared_ptr<uchar[]> own_;
};
struct Point { int x, y; Point() : x(0), y(0) {} Point(int a, int b) : x(a), y(b) {} Point(float a, float b) : x((int)a), y((int)b) {} };
struct DMatch {
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(0.f) {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
  int queryIdx, trainIdx, imgIdx;
  float distance;
};
}  // namespace cv
#endif
