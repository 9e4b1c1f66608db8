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
  // prenormalised: keypoints in rows 1-2 were already passed through PointMatcher::NormalizeKeypoints by the caller
  // d_feat_ptrs (optional, device array of 2*pairs pointers): slot s reads its [n][259] rows from d_feat_ptrs[s] instead of d_feat
  bool run(const float* d_feat, const int* d_n, int feat_cap, int pairs, bool want_dense, cudaStream_t st, bool prenormalised = false,
           const float* const* d_feat_ptrs = nullptr);
  const MatchOutputs& out() const { return out_; }
  int cap() const { return cfg_.cap; }
  double tc_flops(int pairs);
  int launches(int pairs);
  float* x_state() { return x_; }
  const int* counts() const { return n_; }   // [2*pairs] keypoints per slot of the last run (device)

 private:
  bool build_ops(int pairs);
  MatcherConfig cfg_;
  Arena arena_;
  std::map<int, OpList> ops_;
  struct Layer { DenseW qkv, out, ffn0, ffn3, c_qk, c_v, c_qv /* [to_qk ; to_v] rows concatenated: one launch */, c_out, c_ffn0, c_ffn3; float *ln_g, *ln_b, *c_ln_g, *c_ln_b; } L_[9];
  DenseW final_;
  __half* wr_ = nullptr; __half* wm_ = nullptr; float bm_ = 0.f;
  float *x_ = nullptr, *qkv_ = nullptr, *S_ = nullptr, *h_ = nullptr, *rot_ = nullptr, *sim_ = nullptr, *logsig_ = nullptr, *lse_ = nullptr, *row_val_ = nullptr;
  __half *cat16_ = nullptr, *q16_ = nullptr, *k16_ = nullptr, *v16_ = nullptr, *P_ = nullptr, *ctx16_ = nullptr, *h16_ = nullptr, *md16_ = nullptr;
  int *row_arg_ = nullptr, *col_arg_ = nullptr;
  int* n_ = nullptr;   // [2*max_pairs] keypoint counts per slot (device copy owned by the matcher: plans bake this pointer)
  int* row_off_ = nullptr;     // [2*max_pairs + 1] packed row offsets (prefix sums of n_), written by lg_offsets_kernel every run
  __half* md16_pad_ = nullptr; // slot-padded copy of the final projection (operand of the similarity GEMM) in packed mode
  bool packed_ = false;
  MatchOutputs out_;
};

}  // namespace airfe

namespace airfe {

// SuperGlue (G5) device pipeline.  Replaces SuperGlue::infer (src/super_glue.cpp:137-197): process_input (:199-246), the
// engine (keypoint encoder, 18 GNN layers, final projection, 100 in-graph Sinkhorn iterations) and decode (:339-367).
struct SuperGlueOutputs {
  int* idx0 = nullptr;      // [P][cap]  indices0 (-1 = unmatched)
  int* idx1 = nullptr;      // [P][cap]
  float* ms0 = nullptr;     // [P][cap]  mscores0
  float* ms1 = nullptr;
  int* m_idx = nullptr;     // [P][cap][2] mutual matches as PointMatcher::MatchingPoints forms them
  float* m_score = nullptr; // (mscores0[i] + mscores1[j]) / 2
  int* m_count = nullptr;   // [P]
  float* dense = nullptr;   // [P][cap+1][cap+1] final score matrix (parity tap)
};

class SuperGlue {
 public:
  bool init(const MatcherConfig& cfg, const std::string& weights_dir, bool outdoor);
  // max_n: host-side upper bound of the keypoint counts behind d_n (-1 = unknown), see launch_sg_sinkhorn_decode
  bool run(const float* d_feat, const int* d_n, int feat_cap, int pairs, bool want_dense, cudaStream_t st, bool prenormalised = false,
           const float* const* d_feat_ptrs = nullptr, int max_n = -1);
  const SuperGlueOutputs& out() const { return out_; }
  int cap() const { return cfg_.cap; }
  double tc_flops(int pairs) { return build_ops(pairs) ? ops_[pairs].tc_flops : 0.0; }
  int launches(int pairs) { return build_ops(pairs) ? ops_[pairs].launches + 210 : 0; }
  const int* counts() const { return n_; }

 private:
  bool build_ops(int pairs);
  MatcherConfig cfg_;
  Arena arena_;
  std::map<int, OpList> ops_;
  DenseW kenc_[5], final_;
  struct Layer { DenseW qkv, merge, mlp0, mlp3; } L_[18];
  float bin_score_ = 0.f;
  float *x_ = nullptr, *S_ = nullptr, *sim_ = nullptr, *Z_ = nullptr, *u_ = nullptr, *v_ = nullptr, *val0_ = nullptr;
  __half *kin16_ = nullptr, *k1_ = nullptr, *k2_ = nullptr, *cat16_ = nullptr, *qkv16_ = nullptr, *P_ = nullptr, *ctx16_ = nullptr, *h16_ = nullptr, *md16_ = nullptr;
  int *n_ = nullptr, *arg0_ = nullptr, *arg1_ = nullptr;
  SuperGlueOutputs out_;
};

}  // namespace airfe
