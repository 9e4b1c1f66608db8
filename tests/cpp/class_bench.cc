// Latency / throughput of the reference-facing C++ class surface, one stereo pair per call, exactly as the reference's VO feature thread
// drives it (src/map_builder.cc:28-29 construction, :85-86 keyframe branch):
//     _feature_detector->Detect(image_left_rect, image_right_rect, left_features, right_features, left_lines, right_lines, junctions);
//     _point_matcher->MatchingPoints(left_features, right_features, stereo_matches, false);
// Frames live in pageable memory (cv::Mat over a std::vector), results come back in Eigen matrices / std::vectors on the host.
// Usage: class_bench <weights_dir> <left.raw> <right.raw> <w> <h> <n_frames> <iters> <warmup> <matcher 0|1> <use_superpoint 0|1> <max_kp> <line_thr>
// Prints one JSON line.  Driven by `bench.py --mode class`.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "feature_detector.h"
#include "point_matcher.h"

typedef std::chrono::steady_clock Clock;
static double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }
static double pct(std::vector<double> v, double p) {
  if (v.empty()) return 0;
  std::sort(v.begin(), v.end());
  size_t i = (size_t)(p * (v.size() - 1) + 0.5);
  return v[i < v.size() ? i : v.size() - 1];
}

int main(int argc, char** argv) {
  if (argc < 13) { std::cerr << "usage: class_bench weights l.raw r.raw w h n_frames iters warmup matcher use_superpoint max_kp line_thr" << std::endl; return 2; }
  const int w = atoi(argv[4]), h = atoi(argv[5]), nfr = atoi(argv[6]), iters = atoi(argv[7]), warm = atoi(argv[8]), matcher = atoi(argv[9]);
  PLNetConfig plnet_config;
  plnet_config.use_superpoint = atoi(argv[10]);
  plnet_config.max_keypoints = atoi(argv[11]);
  plnet_config.keypoint_threshold = 0.004f; plnet_config.remove_borders = 4;
  plnet_config.line_threshold = (float)atof(argv[12]); plnet_config.line_length_threshold = 50.f;
  plnet_config.SetModelPath(argv[1]);
  PointMatcherConfig pm;
  pm.matcher = matcher; pm.image_width = w; pm.image_height = h;
  pm.onnx_file = std::string(argv[1]) + (matcher ? "/superglue_indoor_sim_int32.onnx" : "/superpoint_lightglue.onnx");
  FeatureDetectorPtr _feature_detector = std::shared_ptr<FeatureDetector>(new FeatureDetector(plnet_config));
  PointMatcherPtr _point_matcher = std::shared_ptr<PointMatcher>(new PointMatcher(pm));

  const size_t one = (size_t)w * h;
  std::vector<unsigned char> lb(one * nfr), rb(one * nfr);
  {
    std::ifstream fl(argv[2], std::ios::binary), fr(argv[3], std::ios::binary);
    fl.read((char*)lb.data(), lb.size());
    fr.read((char*)rb.data(), rb.size());
    if ((size_t)fl.gcount() != lb.size() || (size_t)fr.gcount() != rb.size()) { std::cerr << "short read" << std::endl; return 2; }
  }
  std::vector<double> lat, lat_det, lat_match;
  double matches_sum = 0, d2h = 0;
  Clock::time_point t_all;
  bool all_ok = true;
  for (int i = -warm; i < iters; ++i) {
    if (i == 0) t_all = Clock::now();
    const int f = ((i % nfr) + nfr) % nfr;
    cv::Mat image_left_rect(h, w, CV_8UC1, lb.data() + one * f), image_right_rect(h, w, CV_8UC1, rb.data() + one * f);
    Eigen::Matrix<float, 259, Eigen::Dynamic> left_features, right_features, junctions;
    std::vector<Eigen::Vector4d> left_lines, right_lines;
    std::vector<cv::DMatch> stereo_matches;
    Clock::time_point t0 = Clock::now();
    bool ok = plnet_config.use_superpoint ? _feature_detector->Detect(image_left_rect, image_right_rect, left_features, right_features)
                                          : _feature_detector->Detect(image_left_rect, image_right_rect, left_features, right_features, left_lines, right_lines, junctions);
    const double t_det = ms_since(t0);
    const int n = _point_matcher->MatchingPoints(left_features, right_features, stereo_matches, false);
    const double t = ms_since(t0);
    all_ok = all_ok && ok;
    if (i >= 0) {
      lat.push_back(t); lat_det.push_back(t_det); lat_match.push_back(t - t_det);
      matches_sum += n;
      d2h += (double)(left_features.cols() + right_features.cols() + junctions.cols()) * 259 * 4 + (double)(left_lines.size() + right_lines.size()) * 32 + (double)n * 12;
    }
  }
  const double total_s = ms_since(t_all) * 1e-3;
  double mean = 0;
  for (double v : lat) mean += v;
  mean /= lat.empty() ? 1 : lat.size();
  printf("{\"iters\": %d, \"seconds\": %.6f, \"pairs_per_s\": %.3f, \"mean_ms\": %.4f, \"p50_ms\": %.4f, \"p99_ms\": %.4f, \"detect_p50_ms\": %.4f, \"match_p50_ms\": %.4f, "
         "\"mean_matches\": %.2f, \"d2h_bytes_per_call\": %.0f, \"ok\": %d}\n",
         iters, total_s, iters / total_s, mean, pct(lat, 0.5), pct(lat, 0.99), pct(lat_det, 0.5), pct(lat_match, 0.5), matches_sum / (iters ? iters : 1), d2h / (iters ? iters : 1),
         (int)all_ok);
  return all_ok ? 0 : 1;
}
