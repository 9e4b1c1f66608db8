// Class surfaces of the AirSLAM learned front-end over the libairfe C ABI (include/airfe_c.h).
// Each method restates the CONTROL FLOW of the reference method it replaces (cited), while all arithmetic runs in the
// hand-written sm_100a kernels behind the C ABI.  Error convention = the reference's: bool / int returns + std::cout.
#include <cstdlib>
#include <iostream>

#include "feature_detector.h"
#include "point_matcher.h"

namespace airfe_cpp {
std::string weights_dir_from(const std::string& onnx_path) {
  if (const char* e = std::getenv("AIRFE_WEIGHTS_DIR")) return e;
  const size_t p = onnx_path.find_last_of('/');
  return p == std::string::npos ? std::string(".") : onnx_path.substr(0, p);
}
static airfe_ctx* make_ctx(const airfe_config& cfg) {
  airfe_ctx* c = nullptr;
  int dev = 0;
  if (const char* e = std::getenv("AIRFE_DEVICE")) dev = std::atoi(e);
  if (airfe_create(&cfg, dev, &c) != AIRFE_OK) {
    std::cout << "airfe: " << airfe_last_error(nullptr) << std::endl;
    return nullptr;
  }
  return c;
}
// The outlier-rejection hook of PointMatcher::MatchingPoints exactly as the reference has it (src/point_matcher.cc:95-105): a host step on
// integer-truncated keypoints through cv::findFundamentalMat(FM_RANSAC, 20, 0.99).  It depends on OpenCV's random stream, so it is outside the
// bit-exact parity contract; against real OpenCV this calls OpenCV, in standalone builds the stand-in of compat/opencv2/opencv.hpp.
void RejectOutliersByFundamental(const Eigen::Matrix<float, 259, Eigen::Dynamic>& features0, const Eigen::Matrix<float, 259, Eigen::Dynamic>& features1,
                                 std::vector<cv::DMatch>& matches) {
  std::vector<cv::Point> points0, points1;
  for (auto& m : matches) {
    points0.emplace_back(features0(1, m.queryIdx), features0(2, m.queryIdx));
    points1.emplace_back(features1(1, m.trainIdx), features1(2, m.trainIdx));
  }
  std::vector<uchar> inliers;
  cv::findFundamentalMat(points0, points1, cv::FM_RANSAC, 20, 0.99, inliers);
  int j = 0;
  for (size_t i = 0; i < matches.size(); i++) {
    if (inliers[i]) matches[j++] = matches[i];
  }
  matches.resize(j);
}
}  // namespace airfe_cpp
using airfe_cpp::make_ctx;
typedef Eigen::Matrix<float, 259, Eigen::Dynamic> Features;

// ---------------------------------------------------------------- SuperPoint (src/super_point.cpp:12-144) ----------
SuperPoint::SuperPoint(const SuperPointConfig& c) : super_point_config_(c) {}
bool SuperPoint::build() {
  airfe_config cfg;
  airfe_default_config(&cfg);
  const std::string wd = airfe_cpp::weights_dir_from(super_point_config_.onnx_file);
  cfg.weights_dir = wd.c_str();
  cfg.max_batch = 1;
  cfg.max_keypoints = super_point_config_.max_keypoints;
  cfg.keypoint_threshold = super_point_config_.keypoint_threshold;
  cfg.remove_borders = super_point_config_.remove_borders;
  cfg.enable_superpoint = 1; cfg.enable_plnet = 0; cfg.enable_lightglue = 0; cfg.enable_superglue = 0;
  ctx_.reset(make_ctx(cfg));
  return ctx_ != nullptr;
}
bool SuperPoint::infer(const cv::Mat& image, Features& features) {
  if (!ctx_ || image.empty()) return false;
  const int cap = super_point_config_.max_keypoints;
  features.resize(259, cap);
  int n = 0;
  if (airfe_detect(ctx_.get(), AIRFE_NET_SUPERPOINT, image.data, image.cols, image.rows, (int)image.step, features.data(), cap, &n, nullptr, 0,
                   nullptr, nullptr, 0, nullptr) != AIRFE_OK) {
    std::cout << "airfe: " << airfe_last_error(ctx_.get()) << std::endl;
    return false;
  }
  Features out;
  out.resize(259, n);
  for (long i = 0; i < (long)n * 259; ++i) out.data()[i] = features.data()[i];
  features = out;
  return true;
}
void SuperPoint::save_engine() {}
bool SuperPoint::deserialize_engine() { return false; }

// ---------------------------------------------------------------- PLNet (src/plnet.cpp:17-585) -----------------------
PLNet::PLNet(PLNetConfig& c) : plnet_config_(c) {}
bool PLNet::build() {
  airfe_config cfg;
  airfe_default_config(&cfg);
  const std::string wd = airfe_cpp::weights_dir_from(plnet_config_.plnet_s0_onnx);
  cfg.weights_dir = wd.c_str();
  cfg.max_batch = 1;                  // one stereo pair = 2 images per device batch
  cfg.max_keypoints = plnet_config_.max_keypoints;
  cfg.keypoint_threshold = plnet_config_.keypoint_threshold;
  cfg.remove_borders = plnet_config_.remove_borders;
  cfg.line_threshold = plnet_config_.line_threshold;
  cfg.line_length_threshold = plnet_config_.line_length_threshold;
  cfg.enable_superpoint = 0; cfg.enable_plnet = 1; cfg.enable_lightglue = 0; cfg.enable_superglue = 0;
  ctx_.reset(make_ctx(cfg));
  return ctx_ != nullptr;
}
static void shrink(Features& f, int n) {
  Features out;
  out.resize(259, n);
  for (long i = 0; i < (long)n * 259; ++i) out.data()[i] = f.data()[i];
  f = out;
}
bool PLNet::infer(const cv::Mat& image, Features& features, std::vector<Eigen::Vector4d>& lines, Features& junctions, bool junction_detection) {
  if (!ctx_ || image.empty()) return false;                      // plnet.cpp:247
  const int cap = plnet_config_.max_keypoints, lcap = 4096, jcap = 1024;
  features.resize(259, cap);
  std::vector<double> l((size_t)lcap * 4);
  int n = 0, nl = 0, nj = 0;
  if (junction_detection) junctions.resize(259, jcap);
  if (airfe_detect(ctx_.get(), AIRFE_NET_PLNET, image.data, image.cols, image.rows, (int)image.step, features.data(), cap, &n, l.data(), lcap, &nl,
                   junction_detection ? junctions.data() : nullptr, jcap, &nj) != AIRFE_OK) {
    std::cout << "airfe: " << airfe_last_error(ctx_.get()) << std::endl;
    return false;
  }
  shrink(features, n);
  if (junction_detection) shrink(junctions, nj);
  for (int i = 0; i < nl; ++i) lines.emplace_back(l[i * 4], l[i * 4 + 1], l[i * 4 + 2], l[i * 4 + 3]);   // appends, like plnet.cpp:551
  return true;
}
bool PLNet::infer_pair(const cv::Mat& left, const cv::Mat& right, Features& lf, Features& rf, std::vector<Eigen::Vector4d>& ll,
                       std::vector<Eigen::Vector4d>& rl, Features* lj) {
  if (!ctx_ || left.empty() || right.empty()) return false;
  if (left.cols != right.cols || left.rows != right.rows || left.step != right.step) {   // different sizes: two mono calls
    Features dummy;
    bool a = infer(left, lf, ll, lj ? *lj : dummy, lj != nullptr);
    bool b = infer(right, rf, rl, dummy, false);
    return a & b;
  }
  const int cap = plnet_config_.max_keypoints, lcap = 4096, jcap = 1024;
  std::vector<float> f((size_t)2 * cap * 259), j(lj ? (size_t)2 * jcap * 259 : 0);
  std::vector<double> l((size_t)2 * lcap * 4);
  std::vector<unsigned char> both((size_t)2 * left.rows * left.step);
  memcpy(both.data(), left.data, (size_t)left.rows * left.step);
  memcpy(both.data() + (size_t)left.rows * left.step, right.data, (size_t)left.rows * left.step);
  int n[2] = {0, 0}, nl[2] = {0, 0}, nj[2] = {0, 0};
  if (airfe_detect_batch(ctx_.get(), AIRFE_NET_PLNET, 2, both.data(), left.cols, left.rows, (int)left.step, (long long)left.rows * left.step,
                         f.data(), cap, n, l.data(), lcap, nl, lj ? j.data() : nullptr, jcap, nj) != AIRFE_OK) {
    std::cout << "airfe: " << airfe_last_error(ctx_.get()) << std::endl;
    return false;
  }
  lf.resize(259, n[0]); rf.resize(259, n[1]);
  for (long i = 0; i < (long)n[0] * 259; ++i) lf.data()[i] = f[i];
  for (long i = 0; i < (long)n[1] * 259; ++i) rf.data()[i] = f[(size_t)cap * 259 + i];
  for (int i = 0; i < nl[0]; ++i) ll.emplace_back(l[i * 4], l[i * 4 + 1], l[i * 4 + 2], l[i * 4 + 3]);
  for (int i = 0; i < nl[1]; ++i) rl.emplace_back(l[(size_t)lcap * 4 + i * 4], l[(size_t)lcap * 4 + i * 4 + 1], l[(size_t)lcap * 4 + i * 4 + 2], l[(size_t)lcap * 4 + i * 4 + 3]);
  if (lj) { lj->resize(259, nj[0]); for (long i = 0; i < (long)nj[0] * 259; ++i) lj->data()[i] = j[i]; }   // left image only (feature_detector.cc:100-101)
  return true;
}
void PLNet::save_engine() {}
bool PLNet::deserialize_engine() { return false; }

// ---------------------------------------------------------------- FeatureDetector (src/feature_detector.cc:7-108) ----
FeatureDetector::FeatureDetector(const PLNetConfig& plnet_config) : _plnet_config(plnet_config) {
  if (_plnet_config.use_superpoint) {
    SuperPointConfig sc;
    sc.max_keypoints = plnet_config.max_keypoints;
    sc.keypoint_threshold = plnet_config.keypoint_threshold;
    sc.remove_borders = plnet_config.remove_borders;
    sc.dla_core = -1;
    sc.input_tensor_names.push_back("input");
    sc.output_tensor_names.push_back("scores");
    sc.output_tensor_names.push_back("descriptors");
    sc.onnx_file = plnet_config.superpoint_onnx;
    sc.engine_file = plnet_config.superpoint_engine;
    _superpoint = std::shared_ptr<SuperPoint>(new SuperPoint(sc));
    if (!_superpoint->build()) {
      std::cout << "Error in SuperPoint building" << std::endl;
      exit(0);                                                     // feature_detector.cc:23-26
    }
  }
  _plnet = std::shared_ptr<PLNet>(new PLNet(_plnet_config));
  if (!_plnet->build()) std::cout << "Error in FeatureDetector building" << std::endl;   // non-fatal, :30-33
}
#define AIRFE_REPORT(ok) do { if (!(ok)) std::cout << "Failed when extracting point features !" << std::endl; } while (0)
bool FeatureDetector::Detect(cv::Mat& image, Features& features) {
  bool good;
  if (_plnet_config.use_superpoint) good = _superpoint->infer(image, features);
  else { std::vector<Eigen::Vector4d> lines; return Detect(image, features, lines); }
  AIRFE_REPORT(good);
  return good;
}
bool FeatureDetector::Detect(cv::Mat& image, Features& features, std::vector<Eigen::Vector4d>& lines) {
  Features junctions;
  bool good = _plnet->infer(image, features, lines, junctions);
  AIRFE_REPORT(good);
  return good;
}
bool FeatureDetector::Detect(cv::Mat& image, Features& features, std::vector<Eigen::Vector4d>& lines, Features& junctions) {
  bool good = _plnet->infer(image, features, lines, junctions, true);
  AIRFE_REPORT(good);
  return good;
}
bool FeatureDetector::Detect(cv::Mat& l, cv::Mat& r, Features& lf, Features& rf) {
  bool a = Detect(l, lf), b = Detect(r, rf);
  AIRFE_REPORT(a & b);
  return a & b;
}
bool FeatureDetector::Detect(cv::Mat& l, cv::Mat& r, Features& lf, Features& rf, std::vector<Eigen::Vector4d>& ll, std::vector<Eigen::Vector4d>& rl) {
  bool good = _plnet->infer_pair(l, r, lf, rf, ll, rl, nullptr);       // the reference runs two mono calls (:88-89); one device batch here
  AIRFE_REPORT(good);
  return good;
}
bool FeatureDetector::Detect(cv::Mat& l, cv::Mat& r, Features& lf, Features& rf, std::vector<Eigen::Vector4d>& ll, std::vector<Eigen::Vector4d>& rl,
                             Features& junctions) {
  bool good = _plnet->infer_pair(l, r, lf, rf, ll, rl, &junctions);
  AIRFE_REPORT(good);
  return good;
}

// ---------------------------------------------------------------- matchers ------------------------------------------------
static airfe_ctx* matcher_ctx(const PointMatcherConfig& pc, bool superglue) {
  airfe_config cfg;
  airfe_default_config(&cfg);
  const std::string wd = airfe_cpp::weights_dir_from(pc.onnx_file);
  cfg.weights_dir = wd.c_str();
  cfg.max_batch = 1;
  cfg.max_keypoints = 1024;           // TRT profile maximum of the reference (src/light_glue.cpp:52, src/super_glue.cpp:54)
  cfg.image_width = pc.image_width;
  cfg.image_height = pc.image_height;
  cfg.enable_superpoint = 0; cfg.enable_plnet = 0;
  cfg.enable_lightglue = superglue ? 0 : 1;
  cfg.enable_superglue = superglue ? (pc.onnx_file.find("outdoor") != std::string::npos ? 2 : 1) : 0;
  return make_ctx(cfg);
}
SuperPointLightGlue::SuperPointLightGlue(const PointMatcherConfig& c) : lightglue_config_(c) {}
bool SuperPointLightGlue::build() { ctx_.reset(matcher_ctx(lightglue_config_, false)); return ctx_ != nullptr; }
void SuperPointLightGlue::save_engine() {}
bool SuperPointLightGlue::deserialize_engine() { return false; }
bool SuperPointLightGlue::infer(const Eigen::Matrix<float, 258, Eigen::Dynamic>& f0, const Eigen::Matrix<float, 258, Eigen::Dynamic>& f1,
                                Eigen::Matrix<int, Eigen::Dynamic, 2>& matches_index, Eigen::Matrix<float, Eigen::Dynamic, 1>& matches_score) {
  if (!ctx_) return false;
  const int n0 = (int)f0.cols(), n1 = (int)f1.cols();
  const int cap = n0 > n1 ? n0 : n1;
  std::vector<float> a((size_t)cap * 259, 0.f), b((size_t)cap * 259, 0.f);        // 258 -> 259 rows: row 0 (score) unused by LightGlue
  for (int j = 0; j < n0; ++j) for (int i = 0; i < 258; ++i) a[(size_t)j * 259 + 1 + i] = f0(i, j);
  for (int j = 0; j < n1; ++j) for (int i = 0; i < 258; ++i) b[(size_t)j * 259 + 1 + i] = f1(i, j);
  std::vector<int> i0(1024), i1(1024);
  std::vector<float> sc(1024);
  int nm = 0;
  if (airfe_match_batch_prenormalized(ctx_.get(), AIRFE_MATCHER_LIGHTGLUE, 1, a.data(), &n0, b.data(), &n1, cap, i0.data(), i1.data(), sc.data(), 1024, &nm) != AIRFE_OK) {
    std::cout << "airfe: " << airfe_last_error(ctx_.get()) << std::endl;
    return false;
  }
  matches_index.resize(nm, 2);
  matches_score.resize(nm, 1);
  for (int k = 0; k < nm; ++k) { matches_index(k, 0) = i0[k]; matches_index(k, 1) = i1[k]; matches_score(k) = sc[k]; }
  return true;
}
SuperGlue::SuperGlue(const PointMatcherConfig& c) : superglue_config_(c) {}
bool SuperGlue::build() { ctx_.reset(matcher_ctx(superglue_config_, true)); return ctx_ != nullptr; }
void SuperGlue::save_engine() {}
bool SuperGlue::deserialize_engine() { return false; }
bool SuperGlue::infer(const Features& f0, const Features& f1, Eigen::VectorXi& indices0, Eigen::VectorXi& indices1, Eigen::VectorXd& mscores0,
                      Eigen::VectorXd& mscores1) {
  if (!ctx_) return false;
  const int n0 = (int)f0.cols(), n1 = (int)f1.cols();
  const int cap = n0 > n1 ? n0 : n1;
  std::vector<float> a((size_t)cap * 259, 0.f), b((size_t)cap * 259, 0.f);
  for (long i = 0; i < (long)n0 * 259; ++i) a[i] = f0.data()[i];
  for (long i = 0; i < (long)n1 * 259; ++i) b[i] = f1.data()[i];
  std::vector<int> i0(cap), i1(cap);
  std::vector<float> m0(cap), m1(cap);
  if (airfe_superglue_batch(ctx_.get(), 1, a.data(), &n0, b.data(), &n1, cap, 1, i0.data(), i1.data(), m0.data(), m1.data(), cap) != AIRFE_OK) {
    std::cout << "airfe: " << airfe_last_error(ctx_.get()) << std::endl;
    return false;
  }
  indices0.resize(n0); mscores0.resize(n0); indices1.resize(n1); mscores1.resize(n1);
  for (int k = 0; k < n0; ++k) { indices0(k) = i0[k]; mscores0(k) = (double)m0[k]; }     // float -> double, super_glue.cpp:462-469
  for (int k = 0; k < n1; ++k) { indices1(k) = i1[k]; mscores1(k) = (double)m1[k]; }
  return true;
}

// ---------------------------------------------------------------- PointMatcher (src/point_matcher.cc:6-108) ----------
PointMatcher::PointMatcher(const PointMatcherConfig& config) : _config(config) {
  if (_config.matcher == 0) {
    _config.dla_core = -1;
    _lightglue = std::shared_ptr<SuperPointLightGlue>(new SuperPointLightGlue(_config));
    if (!_lightglue->build()) std::cout << "Erron lightglue building" << std::endl;
  } else if (_config.matcher == 1) {
    _config.dla_core = -1;
    _superglue = std::shared_ptr<SuperGlue>(new SuperGlue(_config));
    if (!_superglue->build()) std::cout << "Erron superglue building" << std::endl;
  } else {
    std::cout << "Plese select the point matcher! (0 for lightglue and 1 for superglue)" << std::endl;
    exit(0);
  }
}
void PointMatcher::NormalizeKeypoints(const Features& features, Features& normalized, int width, int height, float scale) {
  normalized = features;
  float L_inv = 1.0 / std::max(width, height) * scale;
  for (long col = 0; col < features.cols(); ++col) {
    normalized(1, col) = (features(1, col) - width / 2) * L_inv;
    normalized(2, col) = (features(2, col) - height / 2) * L_inv;
  }
}
int PointMatcher::MatchingPoints(const Features& features0, const Features& features1, std::vector<cv::DMatch>& matches, bool outlier_rejection) {
  if (features0.cols() < 1 || features1.cols() < 1) return 0;
  Features nf0, nf1;
  float scale = _config.matcher ? 0.7 : 0.5;
  NormalizeKeypoints(features0, nf0, _config.image_width, _config.image_height, scale);
  NormalizeKeypoints(features1, nf1, _config.image_width, _config.image_height, scale);
  matches.clear();
  if (_config.matcher == 0) {
    if (!_lightglue) return 0;
    Eigen::Matrix<float, 258, Eigen::Dynamic> a, b;
    a.resize(258, nf0.cols()); b.resize(258, nf1.cols());
    for (long j = 0; j < nf0.cols(); ++j) for (int i = 0; i < 258; ++i) a(i, j) = nf0(i + 1, j);     // bottomRows(258)
    for (long j = 0; j < nf1.cols(); ++j) for (int i = 0; i < 258; ++i) b(i, j) = nf1(i + 1, j);
    Eigen::Matrix<int, Eigen::Dynamic, 2> mi;
    Eigen::Matrix<float, Eigen::Dynamic, 1> ms;
    _lightglue->infer(a, b, mi, ms);
    for (long i = 0; i < mi.rows(); i++) matches.emplace_back(mi(i, 0), mi(i, 1), 1.0 - ms(i));
  } else {
    if (!_superglue) return 0;
    Eigen::VectorXi indices0, indices1;
    Eigen::VectorXd mscores0, mscores1;
    _superglue->infer(nf0, nf1, indices0, indices1, mscores0, mscores1);
    for (long i = 0; i < indices0.size(); i++) {
      if (indices0(i) < indices1.size() && indices0(i) >= 0 && indices1(indices0(i)) == i) {
        double d = 1.0 - (mscores0[i] + mscores1[indices0[i]]) / 2.0;
        matches.emplace_back((int)i, indices0[i], d);
      }
    }
  }
  if (outlier_rejection && matches.size() > 8) airfe_cpp::RejectOutliersByFundamental(features0, features1, matches);
  return matches.size();
}
