// LightGlue (G4) device pipeline.  Replaces SuperPointLightGlue::infer (src/light_glue.cpp:120-170) including its host
// process_input (:172-212) and filter_matches (:214-266), and PointMatcher::NormalizeKeypoints (src/point_matcher.cc:39-48).
#pragma once
#include "engine.h"

namespace airfe {

struct MatcherConfig {
  int max_pairs = 8;
  int cap = 512;            // rows per image slot (multiple of 128, <= 1024 = the reference's TRT profile maximum)
  int image_width = 752, image_height = 480;
};

struct MatchOutputs {       // device pointers, per pair p
  int* idx = nullptr;       // [P][cap][2]  (index0, index1), ascending in index0
  float* score = nullptr;   // [P][cap]     exp(log score)
  int* count = nullptr;     // [P]
  float* dense = nullptr;   // [P][cap][cap] log-assignment scores (parity tap; filled only when requested)
};

class LightGlue {
 public:
  bool init(const MatcherConfig& cfg, const std::string& weights_dir);
  // feat: device [2*pairs][feat_cap][259] (slot = 2*pair + side), n: device [2*pairs].  Asynchronous on st.
  bool run(const float* d_feat, const int* d_n, int feat_cap, int pairs, bool want_dense, cudaStream_t st);
  const MatchOutputs& out() const { return out_; }
  int cap() const { return cfg_.cap; }
  double tc_flops(int pairs);
  int launches(int pairs);
  float* x_state() { return x_; }

 private:
  bool build_ops(int pairs);
  MatcherConfig cfg_;
  Arena arena_;
  std::map<int, OpList> ops_;
  struct Layer { DenseW qkv, out, ffn0, ffn3, c_qk, c_v, c_out, c_ffn0, c_ffn3; float *ln_g, *ln_b, *c_ln_g, *c_ln_b; } L_[9];
  DenseW final_;
  __half* wr_ = nullptr; __half* wm_ = nullptr; float bm_ = 0.f;
  float *x_ = nullptr, *qkv_ = nullptr, *S_ = nullptr, *h_ = nullptr, *rot_ = nullptr, *sim_ = nullptr, *logsig_ = nullptr, *lse_ = nullptr, *row_val_ = nullptr;
  __half *cat16_ = nullptr, *q16_ = nullptr, *k16_ = nullptr, *v16_ = nullptr, *P_ = nullptr, *ctx16_ = nullptr, *h16_ = nullptr, *md16_ = nullptr;
  int *row_arg_ = nullptr, *col_arg_ = nullptr;
  int* n_ = nullptr;   // [2*max_pairs] keypoint counts per slot (device copy owned by the matcher: plans bake this pointer)
  MatchOutputs out_;
};

}  // namespace airfe
