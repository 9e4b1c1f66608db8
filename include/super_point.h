// Drop-in for the reference's include/super_point.h:20-67 (same class name, public members and typedef).  The TensorRT
// members (NvInfer.h, tensorrt_buffer) are gone: the engine is libairfe's hand-written sm_100a pipeline behind airfe_c.h.
#ifndef SUPER_POINT_H_
#define SUPER_POINT_H_

#include <string>
#include <memory>
#include <Eigen/Core>
#include <opencv2/opencv.hpp>

#include "read_configs.h"
#include "airfe_handle.h"

class SuperPoint {
public:
    explicit SuperPoint(const SuperPointConfig &super_point_config);

    bool build();                       // loads weights, allocates the device arena (replaces TRT engine build / deserialise)

    bool infer(const cv::Mat &image, Eigen::Matrix<float, 259, Eigen::Dynamic> &features);

    void save_engine();                 // no-op: there is no JIT-built engine to cache
    bool deserialize_engine();          // returns false like "no cached engine"; build() does not need it

private:
    SuperPointConfig super_point_config_;
    airfe_cpp::CtxPtr ctx_;
};

typedef std::shared_ptr<SuperPoint> SuperPointPtr;

#endif //SUPER_POINT_H_
