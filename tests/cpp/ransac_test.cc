// CPU unit test of the outlier-rejection hook of PointMatcher::MatchingPoints (src/point_matcher.cc:95-105) as restated in
// src/frontend/frontend.cc: planted epipolar geometry (rectified stereo + a general two-view case) with planted outliers.
// Prints "kept_inliers total_inliers kept_outliers total_outliers" per case; tests/test_cpp_surface.py checks the rates.  No GPU needed.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "feature_detector.h"
#include "point_matcher.h"

namespace airfe_cpp {
void RejectOutliersByFundamental(const Eigen::Matrix<float, 259, Eigen::Dynamic>&, const Eigen::Matrix<float, 259, Eigen::Dynamic>&, std::vector<cv::DMatch>&);
}

static double urand(unsigned long long& s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (double)((s >> 11) & ((1ull << 53) - 1)) / (double)(1ull << 53); }

int main() {
  for (int mode = 0; mode < 2; ++mode) {
    unsigned long long s = 12345 + mode;
    const int n_in = 200, n_out = 50, n = n_in + n_out;
    Eigen::Matrix<float, 259, Eigen::Dynamic> f0, f1;
    f0.resize(259, n); f1.resize(259, n);
    std::vector<cv::DMatch> matches;
    for (int i = 0; i < n; ++i) {
      // a 3-D point in front of camera 0, projected into both views (f = 450, c = (376, 240))
      const double X = (urand(s) - 0.5) * 8, Y = (urand(s) - 0.5) * 5, Z = 4 + urand(s) * 12;
      const double x0 = 450 * X / Z + 376, y0 = 450 * Y / Z + 240;
      double x1, y1;
      if (mode == 0) { x1 = 450 * (X - 0.4) / Z + 376; y1 = y0; }                                    // rectified stereo: baseline along x
      else { const double Xc = 0.98 * X + 0.199 * Z - 0.5, Zc = -0.199 * X + 0.98 * Z + 0.3, Yc = Y + 0.2; x1 = 450 * Xc / Zc + 376; y1 = 450 * Yc / Zc + 240; }
      if (i >= n_in) { x1 = 20 + urand(s) * 700; y1 = 20 + urand(s) * 440; }                       // outlier: unrelated position
      f0(1, i) = (float)x0; f0(2, i) = (float)y0; f1(1, i) = (float)x1; f1(2, i) = (float)y1;
      matches.emplace_back(i, i, 0.f);
    }
    airfe_cpp::RejectOutliersByFundamental(f0, f1, matches);
    int ki = 0, ko = 0;
    for (auto& m : matches) (m.queryIdx < n_in ? ki : ko)++;
    printf("%d %d %d %d\n", ki, n_in, ko, n_out);
  }
  return 0;
}
