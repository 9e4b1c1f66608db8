// Minimal stand-in for <opencv2/opencv.hpp> (standalone builds only; OpenCV C++ headers are absent in the authoring image).
// cv::Mat here is a non-owning or owning 8-bit single-channel image view: rows / cols / data / step / empty() / ptr().
#ifndef AIRFE_COMPAT_OPENCV
#define AIRFE_COMPAT_OPENCV
#include <cmath>
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

// ---- cv::findFundamentalMat(points0, points1, FM_RANSAC, threshold, confidence, mask) stand-in --------------------------------------
// PointMatcher::MatchingPoints' outlier-rejection hook (src/point_matcher.cc:95-105) calls OpenCV's RANSAC.  OpenCV's C++ headers are absent in
// the authoring image, so standalone builds get this compact, deterministic stand-in (normalised 8-point model from 8 sampled correspondences,
// symmetric epipolar distance against `threshold`, fixed LCG seed): same contract -- a 0/1 inlier mask of the input size -- not the same random
// stream as OpenCV.  In a real AirSLAM tree the genuine <opencv2/opencv.hpp> is found first and the hook runs OpenCV's implementation.
enum { FM_RANSAC = 8 };
namespace compat_detail {
inline void jacobi9_smallest(double A[9][9], double v[9]) {          // eigenvector of the smallest eigenvalue of a symmetric 9 x 9 matrix
  double V[9][9];
  for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) V[i][j] = i == j;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int p = 0; p < 9; ++p) for (int q = p + 1; q < 9; ++q) off += A[p][q] * A[p][q];
    if (off < 1e-22) break;
    for (int p = 0; p < 9; ++p)
      for (int q = p + 1; q < 9; ++q) {
        if (A[p][q] == 0) continue;
        const double th = (A[q][q] - A[p][p]) / (2 * A[p][q]);
        const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1)), c = 1 / std::sqrt(t * t + 1), s2 = t * c;
        for (int k = 0; k < 9; ++k) { const double a = A[k][p], b = A[k][q]; A[k][p] = c * a - s2 * b; A[k][q] = s2 * a + c * b; }
        for (int k = 0; k < 9; ++k) { const double a = A[p][k], b = A[q][k]; A[p][k] = c * a - s2 * b; A[q][k] = s2 * a + c * b; }
        for (int k = 0; k < 9; ++k) { const double a = V[k][p], b = V[k][q]; V[k][p] = c * a - s2 * b; V[k][q] = s2 * a + c * b; }
      }
  }
  int m = 0;
  for (int i = 1; i < 9; ++i) if (A[i][i] < A[m][m]) m = i;
  for (int i = 0; i < 9; ++i) v[i] = V[i][m];
}
}  // namespace compat_detail
struct NoArray {};
inline int findFundamentalMat(const std::vector<Point>& p0, const std::vector<Point>& p1, int /*method*/, double threshold, double /*confidence*/,
                              std::vector<uchar>& mask) {
  const size_t n = p0.size();
  mask.assign(n, 0);
  if (n < 8 || p1.size() != n) return 0;
  // Hartley normalisation (one transform per image)
  double mx0 = 0, my0 = 0, mx1 = 0, my1 = 0, s0 = 0, s1 = 0;
  for (size_t i = 0; i < n; ++i) { mx0 += p0[i].x; my0 += p0[i].y; mx1 += p1[i].x; my1 += p1[i].y; }
  mx0 /= n; my0 /= n; mx1 /= n; my1 /= n;
  for (size_t i = 0; i < n; ++i) { s0 += std::hypot(p0[i].x - mx0, p0[i].y - my0); s1 += std::hypot(p1[i].x - mx1, p1[i].y - my1); }
  s0 = s0 > 0 ? std::sqrt(2.0) * n / s0 : 1; s1 = s1 > 0 ? std::sqrt(2.0) * n / s1 : 1;
  unsigned long long rng = 0x9E3779B97F4A7C15ull;
  size_t best = 0;
  std::vector<uchar> cur(n);
  for (int it = 0; it < 500; ++it) {
    size_t idx[8];
    for (int k = 0; k < 8; ++k) {
      bool dup;
      do {
        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
        idx[k] = (size_t)((rng >> 33) % n);
        dup = false;
        for (int j = 0; j < k; ++j) dup = dup || idx[j] == idx[k];
      } while (dup);
    }
    double A[9][9] = {};
    for (int k = 0; k < 8; ++k) {
      const double x0 = (p0[idx[k]].x - mx0) * s0, y0 = (p0[idx[k]].y - my0) * s0, x1 = (p1[idx[k]].x - mx1) * s1, y1 = (p1[idx[k]].y - my1) * s1;
      const double r[9] = {x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, 1};
      for (int a = 0; a < 9; ++a) for (int b = 0; b < 9; ++b) A[a][b] += r[a] * r[b];
    }
    double f[9];
    compat_detail::jacobi9_smallest(A, f);
    // F = T1^T Fn T0 in pixel coordinates
    const double T0[3][3] = {{s0, 0, -s0 * mx0}, {0, s0, -s0 * my0}, {0, 0, 1}}, T1[3][3] = {{s1, 0, -s1 * mx1}, {0, s1, -s1 * my1}, {0, 0, 1}};
    double Fn[3][3] = {{f[0], f[1], f[2]}, {f[3], f[4], f[5]}, {f[6], f[7], f[8]}}, tmp[3][3] = {}, F[3][3] = {};
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int c = 0; c < 3; ++c) tmp[a][b] += Fn[a][c] * T0[c][b];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int c = 0; c < 3; ++c) F[a][b] += T1[c][a] * tmp[c][b];
    size_t cnt = 0;
    for (size_t i = 0; i < n; ++i) {
      const double x0 = p0[i].x, y0 = p0[i].y, x1 = p1[i].x, y1 = p1[i].y;
      const double l1[3] = {F[0][0] * x0 + F[0][1] * y0 + F[0][2], F[1][0] * x0 + F[1][1] * y0 + F[1][2], F[2][0] * x0 + F[2][1] * y0 + F[2][2]};   // line in image 1
      const double l0[3] = {F[0][0] * x1 + F[1][0] * y1 + F[2][0], F[0][1] * x1 + F[1][1] * y1 + F[2][1], F[0][2] * x1 + F[1][2] * y1 + F[2][2]};   // line in image 0
      const double e = x1 * l1[0] + y1 * l1[1] + l1[2];
      const double d1 = e * e / (l1[0] * l1[0] + l1[1] * l1[1] + 1e-300), d0 = e * e / (l0[0] * l0[0] + l0[1] * l0[1] + 1e-300);
      cur[i] = (d0 > d1 ? d0 : d1) <= threshold * threshold;
      cnt += cur[i];
    }
    if (cnt > best) { best = cnt; mask = cur; }
  }
  return (int)best;
}
}  // namespace cv
#endif
