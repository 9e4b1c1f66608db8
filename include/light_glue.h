// Drop-in for the reference's include/light_glue.h:21-65.
#ifndef LIGHT_GLUE_H_
#define LIGHT_GLUE_H_

#include <string>
#include <memory>
#include <Eigen/Core>
#include <opencv2/opencv.hpp>

#include "read_configs.h"
#include "airfe_handle.h"

class SuperPointLightGlue {
public:
    SuperPointLightGlue() {};
    explicit SuperPointLightGlue(const PointMatcherConfig &lightglue_config);

    bool build();

    // features: rows 0-1 = keypoints ALREADY normalised by PointMatcher::NormalizeKeypoints, rows 2..257 = descriptor
    bool infer(const Eigen::Matrix<float, 258, Eigen::Dynamic> &features0,
               const Eigen::Matrix<float, 258, Eigen::Dynamic> &features1,
               Eigen::Matrix<int, Eigen::Dynamic, 2> &matches_index,
               Eigen::Matrix<float, Eigen::Dynamic, 1> &matches_score);

    void save_engine();
    bool deserialize_engine();

    airfe_ctx* ctx() { return ctx_.get(); }

private:
    PointMatcherConfig lightglue_config_;
    airfe_cpp::CtxPtr ctx_;
};

typedef std::shared_ptr<SuperPointLightGlue> SuperPointLightGluePtr;

#endif //LIGHT_GLUE_H_
