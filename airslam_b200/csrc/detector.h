// Detector: SuperPoint (G1) and PLNet (G2 + G3) device pipelines behind one object.
// Replaces SuperPoint::infer (src/super_point.cpp:103-144) and PLNet::infer (src/plnet.cpp:221-244).
#pragma once
#include "engine.h"

namespace airfe {

struct DetectorConfig {
  int max_batch = 8;
  int max_keypoints = 400;       // configs/visual_odometry/vo_euroc.yaml:3
  float keypoint_threshold = 0.004f;
  int remove_borders = 4;
  float line_threshold = 0.75f;
  float line_length_threshold = 50.f;
  bool enable_lines = true;      // build the PLNet line branch (G2 tail + G3)
};

constexpr int kKpCap = 1024;       // >= max_keypoints; TRT profile max of the matchers (src/light_glue.cpp:52)
constexpr int kCandCap = 16384;    // post-NMS candidates (radius-4 NMS bounds this by 512*512/25)
constexpr int kJunc = 300;         // TopK junctions (G2)
constexpr int kProp = 3 * 128 * 128;
constexpr int kLineCap = 16384;    // unique junction pairs verified per image (G3 rows); worst case in theory 44850

struct DetectOutputs {
  // device pointers, valid after run(); per image b
  float* feat = nullptr;       // [B][kKpCap][259]
  int* n_feat = nullptr;       // [B]
  float* lines = nullptr;      // [B][kLineCap][4] (x1,y1,x2,y2) image px, fp32 (widened to double at the class surface)
  int* n_lines = nullptr;      // [B]
  float* junc = nullptr;       // [B][kKpCap][259]
  int* n_junc = nullptr;       // [B]
};

class Detector {
 public:
  bool init(const DetectorConfig& cfg, const std::string& weights_dir, bool use_plnet_weights);
  // images already on the device: u8 [B][h][stride]; asynchronous on `st`
  // remap (optional): d_images are RAW camera frames; Camera::UndistortImage (cv::remap) is applied inside the resize kernel
  bool run(const uint8_t* d_images, int batch, int w, int h, int stride, long long img_stride, bool lines, bool junctions,
           cudaStream_t st, const RemapMaps* remap = nullptr);
  const DetectOutputs& out() const { return out_; }
  double tc_flops(int batch, bool lines);
  int launches(int batch, bool lines);
  // debug / parity taps (device pointers)
  float* heat() { return heat_; }
  float* scores() { return scores_; }
  float* desc_raw() { return desc_raw_; }
  __half* x16() { return x16_; }
  const int* n_unique() const { return n_unique_; }   // [B] unique junction pairs = rows of the stage-1 MLP
  struct Taps;  // stage outputs for stage-wise parity tests
  std::map<std::string, std::pair<void*, size_t>> taps;   // name -> (device ptr, bytes per image)

 private:
  bool build_ops(int batch);
  bool ensure_tables(int w, int h);

  DetectorConfig cfg_;
  bool plnet_ = false;
  Arena arena_;
  std::map<int, OpList> trunk_ops_, line_ops_, mlp_ops_;
  std::map<std::pair<int, int>, ResizeTables> tables_;
  // weights
  DenseW w1a_raw_;  // unused by tc path
  __half* w_conv1a_ = nullptr; float* b_conv1a_ = nullptr;
  DenseW w1b_, w2a_, w2b_, w3a_, w3b_, w4a_, w4b_, wPD_, wPb_, wDb_;
  // line branch
  DenseW l1a_, l1b_, l2a_, l2b_, fc2_, heads0_, heads2_, fc1_, fc34_;
  bool fuse_up_ = true;
  struct HG { DenseW c[5][2], dec[4], aup[4], bup[4]; } hg_[2];
  DenseW s1_fc0_, s1_fc2_, s1_fc4_, s1_res_;
  float* s1_head_w_ = nullptr;  // [2][128] fp32
  float* s1_head_b_ = nullptr;  // [2]
  float* s1_tspan_ = nullptr;   // [30]
  // buffers (all sized for max_batch)
  uint8_t* img_u8_ = nullptr;
  __half* x16_ = nullptr;
  Act a1_, r1_, p1_, a2_, cat2_, p2_, a3_, cat3_, p3_, a4_, r7_, pd_, logits_, descraw_;
  float *heat_ = nullptr, *scores_ = nullptr, *desc_raw_ = nullptr, *kp_ = nullptr;
  uint8_t *mask_a_ = nullptr, *mask_b_ = nullptr;
  int *cand_ = nullptr, *cand_count_ = nullptr;
  // line buffers
  Act l1a_o_, l1b_o_, l2a_o_, l2b_o_, fc2_o_, hmid_o_, heads9_o_, loi_o_, thinaux_o_;
  struct HGBuf { Act a[5], r[5], pool[4], up[4], cat[4], u[4]; } hgb_[2];
  float *lines_pred_ = nullptr, *juncs_ = nullptr, *jloc_ = nullptr;
  int *imin_ = nullptr, *imax_ = nullptr, *pair_table_ = nullptr, *uid_pairs_ = nullptr, *uid_first_ = nullptr, *n_unique_ = nullptr;
  uint8_t *iskeep_ = nullptr, *junc_map_ = nullptr;
  Act feat496_, mlp_a_, mlp_b_, mlp_c_, mlp_r_;
  float* line_score_ = nullptr;
  float* adj_ = nullptr;
  int* junc_idx_ = nullptr;
  __half* jfeat_ = nullptr;
  float* jkp_ = nullptr;
  int* jkp_count_ = nullptr;
  DetectOutputs out_;
  friend struct DetectorAccess;
};

}  // namespace airfe
