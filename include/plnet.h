// Drop-in for the reference's include/plnet.h:17-91 (same class name, public members and typedef).
#ifndef PLNET_PLNET_H
#define PLNET_PLNET_H

#include <Eigen/Core>
#include <memory>
#include <opencv2/opencv.hpp>
#include <string>
#include <vector>

#include "read_configs.h"
#include "airfe_handle.h"

class PLNet {
 public:
  PLNet(PLNetConfig& plnet_config);

  bool build();

  bool infer(const cv::Mat &image, Eigen::Matrix<float, 259, Eigen::Dynamic> &features,
      std::vector<Eigen::Vector4d>& lines, Eigen::Matrix<float, 259, Eigen::Dynamic>& junctions, bool junction_detection = false);

  void save_engine();
  bool deserialize_engine();

  // extension used by FeatureDetector's stereo overloads: both images in one device batch (same results as two calls)
  bool infer_pair(const cv::Mat& left, const cv::Mat& right, Eigen::Matrix<float, 259, Eigen::Dynamic>& lf,
                  Eigen::Matrix<float, 259, Eigen::Dynamic>& rf, std::vector<Eigen::Vector4d>& ll, std::vector<Eigen::Vector4d>& rl,
                  Eigen::Matrix<float, 259, Eigen::Dynamic>* left_junctions);

 private:
  PLNetConfig plnet_config_;
  airfe_cpp::CtxPtr ctx_;
};

typedef std::shared_ptr<PLNet> PLNetPtr;

#endif  // PLNET_PLNET_H
