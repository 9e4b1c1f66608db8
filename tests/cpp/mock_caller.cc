// Mock of the reference's callers (src/map_builder.cc:28-29,85-107 and src/map_user.cc:31-32,111,369): holds the front-end only
// through FeatureDetectorPtr / PointMatcherPtr and calls only the public members, exactly as MapBuilder / MapUser do.
// Usage: mock_caller <weights_dir> <left.pgm-raw> <right.pgm-raw> <w> <h> <matcher 0|1> [use_superpoint]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "feature_detector.h"
#include "point_matcher.h"

static std::vector<unsigned char> slurp(const char* p, size_t n) {
  std::vector<unsigned char> b(n);
  std::ifstream f(p, std::ios::binary);
  f.read((char*)b.data(), n);
  if ((size_t)f.gcount() != n) { std::cerr << "short read " << p << std::endl; exit(2); }
  return b;
}

int main(int argc, char** argv) {
  if (argc < 7) { std::cerr << "usage" << std::endl; return 2; }
  const int w = atoi(argv[4]), h = atoi(argv[5]), matcher = atoi(argv[6]);
  PLNetConfig plnet_config;
  plnet_config.use_superpoint = argc > 7 ? atoi(argv[7]) : 0;
  plnet_config.max_keypoints = 400; plnet_config.keypoint_threshold = 0.004f; plnet_config.remove_borders = 4;
  plnet_config.line_threshold = 0.75f; plnet_config.line_length_threshold = 50.f;
  plnet_config.SetModelPath(argv[1]);
  PointMatcherConfig pm;
  pm.matcher = matcher; pm.image_width = w; pm.image_height = h;
  pm.onnx_file = std::string(argv[1]) + (matcher ? "/superglue_indoor_sim_int32.onnx" : "/superpoint_lightglue.onnx");
  FeatureDetectorPtr _feature_detector = std::shared_ptr<FeatureDetector>(new FeatureDetector(plnet_config));   // map_builder.cc:28
  PointMatcherPtr _point_matcher = std::shared_ptr<PointMatcher>(new PointMatcher(pm));                        // map_builder.cc:29

  std::vector<unsigned char> lb = slurp(argv[2], (size_t)w * h), rb = slurp(argv[3], (size_t)w * h);
  cv::Mat image_left_rect(h, w, CV_8UC1, lb.data()), image_right_rect(h, w, CV_8UC1, rb.data());
  Eigen::Matrix<float, 259, Eigen::Dynamic> left_features, right_features, junctions;
  std::vector<Eigen::Vector4d> left_lines, right_lines;
  std::vector<cv::DMatch> stereo_matches;
  // keyframe branch, map_builder.cc:85-86
  bool ok = _feature_detector->Detect(image_left_rect, image_right_rect, left_features, right_features, left_lines, right_lines, junctions);
  int n = _point_matcher->MatchingPoints(left_features, right_features, stereo_matches, false);
  printf("ok %d left %ld right %ld llines %zu rlines %zu junctions %ld matches %d\n", (int)ok, left_features.cols(), right_features.cols(),
         left_lines.size(), right_lines.size(), junctions.cols(), n);
  // non-keyframe branch, map_builder.cc:94 (+ :101 temporal match against the last keyframe)
  Eigen::Matrix<float, 259, Eigen::Dynamic> features;
  ok = _feature_detector->Detect(image_left_rect, features);
  std::vector<cv::DMatch> matches;
  n = _point_matcher->MatchingPoints(left_features, features, matches, false);
  printf("mono ok %d features %ld matches %d\n", (int)ok, features.cols(), n);
  // relocalization query, map_user.cc:111
  Eigen::Matrix<float, 259, Eigen::Dynamic> q, qj;
  std::vector<Eigen::Vector4d> ql;
  ok = _feature_detector->Detect(image_left_rect, q, ql, qj);
  printf("reloc ok %d features %ld lines %zu junctions %ld\n", (int)ok, q.cols(), ql.size(), qj.cols());
  // dump for the python-side comparison
  FILE* f = fopen("mock_caller_out.bin", "wb");
  if (f) {
    long hdr[6] = {left_features.cols(), right_features.cols(), (long)left_lines.size(), (long)right_lines.size(), junctions.cols(), (long)stereo_matches.size()};
    fwrite(hdr, sizeof(long), 6, f);
    fwrite(left_features.data(), 4, (size_t)left_features.cols() * 259, f);
    fwrite(right_features.data(), 4, (size_t)right_features.cols() * 259, f);
    for (auto& l : left_lines) fwrite(&l[0], 8, 4, f);
    for (auto& l : right_lines) fwrite(&l[0], 8, 4, f);
    fwrite(junctions.data(), 4, (size_t)junctions.cols() * 259, f);
    for (auto& m : stereo_matches) { int a[2] = {m.queryIdx, m.trainIdx}; fwrite(a, 4, 2, f); fwrite(&m.distance, 4, 1, f); }
    fclose(f);
  }
  return ok ? 0 : 1;
}
