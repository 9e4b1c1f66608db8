// Subset of the reference's include/read_configs.h (the three front-end config structs, :9-103) WITHOUT the yaml-cpp
// loaders and utils.h (which pulls in g2o / Boost).  Field names and meaning are the reference's.  In a real AirSLAM tree
// the original read_configs.h is used instead of this file.
#ifndef READ_CONFIGS_H_
#define READ_CONFIGS_H_
#include <string>
#include <vector>

struct PLNetConfig {
  std::string superpoint_onnx, superpoint_engine;
  std::string plnet_s0_onnx, plnet_s0_engine, plnet_s1_onnx, plnet_s1_engine;
  int use_superpoint = 0;
  int max_keypoints = 400;
  float keypoint_threshold = 0.004f;
  int remove_borders = 4;
  float line_threshold = 0.75f;
  float line_length_threshold = 50.f;
  void SetModelPath(std::string model_dir) {
    auto cat = [&](const char* f) { return model_dir + (model_dir.empty() || model_dir.back() == '/' ? "" : "/") + f; };
    if (use_superpoint) { superpoint_onnx = cat("superpoint_v1_sim_int32.onnx"); superpoint_engine = cat("superpoint_v1_sim_int32.engine"); }
    plnet_s0_onnx = cat("plnet_s0.onnx"); plnet_s0_engine = cat("plnet_s0.engine");
    plnet_s1_onnx = cat("plnet_s1.onnx"); plnet_s1_engine = cat("plnet_s1.engine");
  }
};

struct SuperPointConfig {
  int max_keypoints = 400;
  float keypoint_threshold = 0.004f;
  int remove_borders = 4;
  int dla_core = -1;
  std::vector<std::string> input_tensor_names, output_tensor_names;
  std::string onnx_file, engine_file;
};

struct PointMatcherConfig {
  int matcher = 0;
  int image_width = 752, image_height = 480;
  int dla_core = -1;
  std::vector<std::string> input_tensor_names, output_tensor_names;
  std::string onnx_file, engine_file;
};
#endif
