// Drop-in for the reference's include/super_glue.h:20-75.
#ifndef SUPER_GLUE_H_
#define SUPER_GLUE_H_

#include <string>
#include <memory>
#include <Eigen/Core>
#include <opencv2/opencv.hpp>

#include "read_configs.h"
#include "airfe_handle.h"

class SuperGlue {
public:
    SuperGlue() {};

    explicit SuperGlue(const PointMatcherConfig &superglue_config);

    bool build();

    // features: row 0 score, rows 1-2 keypoints ALREADY normalised by PointMatcher::NormalizeKeypoints, rows 3..258 descriptor
    bool infer(const Eigen::Matrix<float, 259, Eigen::Dynamic> &features0,
               const Eigen::Matrix<float, 259, Eigen::Dynamic> &features1,
               Eigen::VectorXi &indices0,
               Eigen::VectorXi &indices1,
               Eigen::VectorXd &mscores0,
               Eigen::VectorXd &mscores1);

    void save_engine();
    bool deserialize_engine();

    airfe_ctx* ctx() { return ctx_.get(); }

private:
    PointMatcherConfig superglue_config_;
    airfe_cpp::CtxPtr ctx_;
};

typedef std::shared_ptr<SuperGlue> SuperGluePtr;

#endif //SUPER_GLUE_H_
