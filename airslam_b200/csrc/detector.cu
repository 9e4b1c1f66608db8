#include "detector.h"
#include "line_kernels.h"
#include <stdlib.h>

namespace airfe {

static_assert(kJunc == kJunctions && kProp == kProposals, "constants out of sync");

static Act make_act(Arena* ar, int bmax, int H, int W, int C, bool f32 = false, int ps = 0) {
  Act a;
  a.C = C; a.H = H; a.W = W; a.ps = ps ? ps : C; a.f32 = f32;
  a.p = ar->alloc((size_t)bmax * H * W * a.ps * (f32 ? 4 : 2));
  return a;
}

static bool upload_f32(Arena* ar, const std::vector<float>& v, float** out) {
  *out = ar->alloc_n<float>(v.size());
  if (!*out) return false;
  AIRFE_CUDA_OK(cudaMemcpy(*out, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  return true;
}

bool Detector::init(const DetectorConfig& cfg, const std::string& wdir, bool use_plnet_weights) {
  cfg_ = cfg;
  plnet_ = use_plnet_weights;
  if (cfg.max_keypoints > kKpCap) { set_error("max_keypoints %d exceeds capacity %d", cfg.max_keypoints, kKpCap); return false; }
  WeightFile wf;
  if (!wf.load(wdir + (plnet_ ? "/plnet.afw" : "/superpoint.afw"))) return false;
  const std::string P = plnet_ ? "plnet.pd." : "sp.";
  const int B = cfg.max_batch;
  const bool lines = plnet_ && cfg.enable_lines;
  size_t bytes = (size_t)B * (lines ? 420u : 150u) * 1024 * 1024 + (64u << 20);
  if (!arena_.init(bytes)) return false;
  Arena* ar = &arena_;

  // ---- weights -------------------------------------------------------------------------------------------------------
  {
    const WTensor* t = wf.find(P + "conv1a.weight");
    std::vector<float> b;
    if (!t || !wf.get_f32(P + "conv1a.bias", &b)) { set_error("missing conv1a"); return false; }
    w_conv1a_ = ar->alloc_n<__half>(576);
    AIRFE_CUDA_OK(cudaMemcpy(w_conv1a_, t->data, 576 * 2, cudaMemcpyHostToDevice));
    if (!upload_f32(ar, b, &b_conv1a_)) return false;
  }
  auto pk = [&](const std::string& name, int cin, DenseW* out) { return pack_dense(wf, {{P + name + ".weight", P + name + ".bias", 0}}, cin, ar, out); };
  if (!pk("conv1b", 64, &w1b_) || !pk("conv2a", 64, &w2a_) || !pk("conv2b", 64, &w2b_) || !pk("conv3a", 64, &w3a_) ||
      !pk("conv3b", 128, &w3b_) || !pk("conv4a", 128, &w4a_) || !pk("conv4b", 128, &w4b_) || !pk("convPb", 256, &wPb_) ||
      !pk("convDb", 256, &wDb_))
    return false;
  if (!pack_dense(wf, {{P + "convPa.weight", P + "convPa.bias", 0}, {P + "convDa.weight", P + "convDa.bias", 0}}, 128, ar, &wPD_)) return false;

  // ---- trunk buffers -----------------------------------------------------------------------------------------------------
  img_u8_ = ar->alloc_n<uint8_t>((size_t)B * 1280 * 1024);   // staging for up to 1280x1024 frames
  x16_ = ar->alloc_n<__half>((size_t)B * 262144);
  a1_ = make_act(ar, B, 512, 512, 64);
  r1_ = make_act(ar, B, 512, 512, 64);
  p1_ = make_act(ar, B, 256, 256, 64);
  a2_ = make_act(ar, B, 256, 256, 64);
  cat2_ = make_act(ar, B, 256, 256, 96);     // [line pool(32) | SP relu_3 (64)]
  p2_ = make_act(ar, B, 128, 128, 64);
  a3_ = make_act(ar, B, 128, 128, 128);
  cat3_ = make_act(ar, B, 128, 128, 256);    // [line pool_1 (128) | SP relu_5 (128)]
  p3_ = make_act(ar, B, 64, 64, 128);
  a4_ = make_act(ar, B, 64, 64, 128);
  r7_ = make_act(ar, B, 64, 64, 128);
  pd_ = make_act(ar, B, 64, 64, 512);        // [convPa | convDa]
  logits_ = make_act(ar, B, 64, 64, 80, true);
  descraw_ = make_act(ar, B, 64, 64, 256, true);
  desc_raw_ = (float*)descraw_.p;
  heat_ = ar->alloc_n<float>((size_t)B * 262144);
  scores_ = ar->alloc_n<float>((size_t)B * 262144);
  mask_a_ = ar->alloc_n<uint8_t>((size_t)B * 262144);
  mask_b_ = ar->alloc_n<uint8_t>((size_t)B * 262144);
  cand_ = ar->alloc_n<int>((size_t)B * kCandCap);
  cand_count_ = ar->alloc_n<int>(B);
  kp_ = ar->alloc_n<float>((size_t)B * kKpCap * 3);
  out_.feat = ar->alloc_n<float>((size_t)B * kKpCap * 259);
  out_.n_feat = ar->alloc_n<int>(B);

  if (lines) {
    const std::string L = "plnet.";
    auto pl = [&](const std::string& name, int cin, DenseW* out) { return pack_dense(wf, {{L + name + ".weight", L + name + ".bias", 0}}, cin, ar, out); };
    if (!pl("conv1a", 64, &l1a_) || !pack_dense(wf, {{L + "conv1b.weight", L + "conv1b.bias", 0}}, 32, ar, &l1b_, nullptr, nullptr, (conv3x3_halo_enabled() && !getenv("AIRFE_CONV_K64")) ? 32 : 64) || !pack_dense(wf, {{L + "conv2a.weight", L + "conv2a.bias", 0}}, 96, ar, &l2a_, nullptr, nullptr, (conv3x3_halo_enabled() && !getenv("AIRFE_CONV_K64")) ? 32 : 64 /* 96 = 3 x 32: no K padding on the SWIZZLE_64B path */) || !pl("conv2b", 128, &l2b_) ||
        !pl("fc2", 128, &fc2_) || !pl("fc1", 256, &fc1_))
      return false;
    for (int s = 0; s < 2; ++s) {
      const std::string S = "stack" + std::to_string(s + 1) + ".";
      for (int l = 0; l < 5; ++l) {
        if (!pl(S + "conv" + std::to_string(l + 1) + "a", (s == 0 && l == 0) ? 256 : 128, &hg_[s].c[l][0])) return false;
        if (!pl(S + "conv" + std::to_string(l + 1) + "b", 128, &hg_[s].c[l][1])) return false;
      }
      for (int u = 0; u < 4; ++u) {   // u = 0 is the coarsest up step (deconv1, conv4a_up, conv4b_up)
        if (!pl(S + "deconv" + std::to_string(u + 1), 128, &hg_[s].dec[u])) return false;
        if (!pl(S + "conv" + std::to_string(4 - u) + "a_up", 128, &hg_[s].aup[u])) return false;
        if (!pl(S + "conv" + std::to_string(4 - u) + "b_up", 128, &hg_[s].bup[u])) return false;
      }
    }
    // five head stems 256 -> 64 (3x3) fused into one 256 -> 320 conv; five 1x1 heads fused block-diagonally into 320 -> 9
    {
      std::vector<PackSrc> a, c;
      for (int i = 0; i < 5; ++i) {
        a.push_back({L + "score1.heads." + std::to_string(i) + ".0.weight", L + "score1.heads." + std::to_string(i) + ".0.bias", 0});
        c.push_back({L + "score1.heads." + std::to_string(i) + ".2.weight", L + "score1.heads." + std::to_string(i) + ".2.bias", 64 * i});
      }
      if (!pack_dense(wf, a, 256, ar, &heads0_) || !pack_dense(wf, c, 320, ar, &heads2_)) return false;
    }
    if (!pack_dense(wf, {{L + "fc3.weight", L + "fc3.bias", 0}, {L + "fc4.weight", L + "fc4.bias", 0}}, 256, ar, &fc34_)) return false;
    // stage 1 (G3) MLP
    auto p1 = [&](const std::string& name, int cin, DenseW* out) { return pack_dense(wf, {{L + "s1." + name + ".weight", L + "s1." + name + ".bias", 0}}, cin, ar, out); };
    if (!p1("fc2.0", 496, &s1_fc0_) || !p1("fc2.2", 128, &s1_fc2_) || !p1("fc2.4", 128, &s1_fc4_) || !p1("fc2_res.0", 240, &s1_res_)) return false;
    {
      std::vector<float> hw, hb, t, tc;
      if (!wf.get_f32(L + "s1.fc2_head.weight", &hw) || !wf.get_f32(L + "s1.fc2_head.bias", &hb) || !wf.get_f32(L + "s1.tspan", &t) ||
          !wf.get_f32(L + "s1.tspan_c", &tc))
        return false;
      t.insert(t.end(), tc.begin(), tc.end());
      if (!upload_f32(ar, hw, &s1_head_w_) || !upload_f32(ar, hb, &s1_head_b_) || !upload_f32(ar, t, &s1_tspan_)) return false;
    }
    // line buffers
    l1a_o_ = make_act(ar, B, 512, 512, 32);
    l1b_o_ = make_act(ar, B, 512, 512, 32);
    l2a_o_ = make_act(ar, B, 256, 256, 128);
    l2b_o_ = make_act(ar, B, 256, 256, 128);
    // The Resize (nearest, 2x) node in front of every `deconv` conv is fused into the epilogue of the conv that produces its input: that conv
    // stores each pixel four times, straight into the up-sampled map (the low-resolution map itself has no other reader).  AIRFE_UP2_FUSE=0
    // keeps the separate upsample2 launches (8 per frame batch, 0.47 ms per 94 frames) for A/B timing.
    fuse_up_ = conv3x3_halo_enabled() && !(getenv("AIRFE_UP2_FUSE") && atoi(getenv("AIRFE_UP2_FUSE")) == 0);
    for (int s = 0; s < 2; ++s) {
      int hw_ = 128;
      for (int l = 0; l < 5; ++l) {
        hgb_[s].a[l] = make_act(ar, B, hw_, hw_, 128);
        hgb_[s].r[l] = make_act(ar, B, hw_, hw_, 128);
        if (l < 4) hgb_[s].pool[l] = make_act(ar, B, hw_ / 2, hw_ / 2, 128);
        hw_ /= 2;
      }
      hw_ = 16;
      for (int u = 0; u < 4; ++u) {
        hgb_[s].up[u] = make_act(ar, B, hw_, hw_, 128);
        hgb_[s].cat[u] = make_act(ar, B, hw_, hw_, 128);   // [deconv relu (64) | skip conv relu (64)]
        if (!(fuse_up_ && u < 3)) hgb_[s].u[u] = make_act(ar, B, hw_, hw_, 128);     // fused up-sampling: decoder levels 0-2 are stored up-sampled only
        hw_ *= 2;
      }
    }
    fc2_o_ = make_act(ar, B, 128, 128, 256);
    hmid_o_ = make_act(ar, B, 128, 128, 320);
    heads9_o_ = make_act(ar, B, 128, 128, 16, true);
    // fc1 (LOI features) and fc3 | fc4 (thin / aux) stay two GEMMs over the fc2 map: one merged 256 -> 136 GEMM was measured in round 2
    // (profiles/r02e_fc134_merge_ab.txt) -- it saves a read of the fc2 map (0.41 -> 0.32 ms) but leaves thin / aux strided by the LOI row, and the
    // LOI sampler's 240 scattered 16-byte loads per line then touch 640-byte-strided pixels instead of a compact 32-byte-per-pixel map (+0.07 .. 0.10 ms).
    loi_o_ = make_act(ar, B, 128, 128, 128, true);
    thinaux_o_ = make_act(ar, B, 128, 128, 8, true);
    lines_pred_ = ar->alloc_n<float>((size_t)B * kProp * 4);
    jloc_ = ar->alloc_n<float>((size_t)B * 16384);
    juncs_ = ar->alloc_n<float>((size_t)B * kJunc * 2);
    imin_ = ar->alloc_n<int>((size_t)B * kProp);
    imax_ = ar->alloc_n<int>((size_t)B * kProp);
    iskeep_ = ar->alloc_n<uint8_t>((size_t)B * kProp);
    pair_table_ = ar->alloc_n<int>((size_t)B * kJunc * kJunc);
    uid_pairs_ = ar->alloc_n<int>((size_t)B * kLineCap * 2);
    uid_first_ = ar->alloc_n<int>((size_t)B * kLineCap);
    n_unique_ = ar->alloc_n<int>(B);
    junc_map_ = ar->alloc_n<uint8_t>((size_t)B * 262144);
    // stage-1 matrices: rows = B * kLineCap (image b owns rows [b*kLineCap, ...))
    feat496_ = make_act(ar, B, 1, kLineCap, 512);
    mlp_a_ = make_act(ar, B, 1, kLineCap, 128);
    mlp_b_ = make_act(ar, B, 1, kLineCap, 128);
    mlp_c_ = make_act(ar, B, 1, kLineCap, 128, true);
    mlp_r_ = make_act(ar, B, 1, kLineCap, 128, true);
    line_score_ = ar->alloc_n<float>((size_t)B * kLineCap);
    adj_ = ar->alloc_n<float>((size_t)B * kLineCap * 4);
    junc_idx_ = ar->alloc_n<int>((size_t)B * kJunc);
    jfeat_ = ar->alloc_n<__half>((size_t)B * kJunc * 128);      // LOI endpoint features per junction (junc_feat_kernel)
    jkp_ = ar->alloc_n<float>((size_t)B * kKpCap * 3);
    jkp_count_ = ar->alloc_n<int>(B);
    out_.lines = ar->alloc_n<float>((size_t)B * kLineCap * 4);
    out_.n_lines = ar->alloc_n<int>(B);
    out_.junc = ar->alloc_n<float>((size_t)B * kKpCap * 259);
    out_.n_junc = ar->alloc_n<int>(B);
  }
  if (!ar->ok()) return false;
  taps["x16"] = {x16_, 262144 * 2};
  taps["relu_1"] = {r1_.p, (size_t)262144 * 64 * 2};
  taps["relu_7"] = {r7_.p, (size_t)4096 * 128 * 2};
  taps["logits"] = {logits_.p, (size_t)4096 * 80 * 4};
  taps["desc_raw"] = {desc_raw_, (size_t)4096 * 256 * 4};
  taps["heat"] = {heat_, 262144 * 4};
  taps["scores"] = {scores_, 262144 * 4};
  taps["kp"] = {kp_, (size_t)kKpCap * 3 * 4};
  if (lines) {
    taps["cat3"] = {cat3_.p, (size_t)16384 * 256 * 2};
    taps["stack1_out"] = {hgb_[0].u[3].p, (size_t)16384 * 128 * 2};
    taps["fc2"] = {fc2_o_.p, (size_t)16384 * 256 * 2};
    taps["heads9"] = {heads9_o_.p, (size_t)16384 * 16 * 4};
    taps["loi"] = {loi_o_.p, (size_t)16384 * 128 * 4};
    taps["thinaux"] = {thinaux_o_.p, (size_t)16384 * 8 * 4};
    taps["lines_pred"] = {lines_pred_, (size_t)kProp * 4 * 4};
    taps["jloc"] = {jloc_, 16384 * 4};
    taps["juncs_pred"] = {juncs_, kJunc * 2 * 4};
    taps["imin"] = {imin_, (size_t)kProp * 4};
    taps["imax"] = {imax_, (size_t)kProp * 4};
    taps["iskeep"] = {iskeep_, (size_t)kProp};
    taps["uid_pairs"] = {uid_pairs_, (size_t)kLineCap * 2 * 4};
    taps["uid_first"] = {uid_first_, (size_t)kLineCap * 4};
    taps["n_unique"] = {n_unique_, 4};
    taps["feat496"] = {feat496_.p, (size_t)kLineCap * 512 * 2};
    taps["line_score"] = {line_score_, (size_t)kLineCap * 4};
    taps["lines_adjusted"] = {adj_, (size_t)kLineCap * 4 * 4};
    taps["junc_idx"] = {junc_idx_, (size_t)kJunc * 4};
  }
  return true;
}

bool Detector::ensure_tables(int w, int h) {
  auto key = std::make_pair(w, h);
  if (tables_.count(key)) return true;
  std::vector<int> host(6 * 512);
  build_resize_tables_host(w, h, &host[0], &host[512], &host[1024], &host[1536], &host[2048], &host[2560]);
  int* d = nullptr;
  AIRFE_CUDA_OK(cudaMalloc(&d, host.size() * 4));
  AIRFE_CUDA_OK(cudaMemcpy(d, host.data(), host.size() * 4, cudaMemcpyHostToDevice));
  ResizeTables t{d, d + 512, d + 1024, d + 1536, d + 2048, d + 2560};
  tables_[key] = t;
  return true;
}

bool Detector::build_ops(int B) {
  if (trunk_ops_.count(B)) return true;
  OpList t;
  const bool lines = plnet_ && cfg_.enable_lines;
  auto up = [&](OpList* ol, const Act& in, const Act& out) {
    const Act i = in, o = out;
    ol->push("upsample2", 0, [=](cudaStream_t st) { launch_upsample2((const __half*)i.p, i.C, i.H, i.W, B, i.ps, (__half*)o.p, o.ps, st); return true; });
    ol->launches++;
  };
  // SuperPoint trunk + heads (G1; G2 /backbone/point_detector/*)
  const Act r3 = cat2_.slice(32, 64), r5 = cat3_.slice(128, 128);
  {
    // conv1a (C_in = 1) on CUDA cores (packed FFMA2), then conv1b on tensor cores.  Computing conv1a inside the conv1b kernel was measured in
    // round 1 with FFMA2 producer warps (slower: profiles/r01_fused_conv1a_trace.txt); the drafted tensor-core producer was never validated
    // and has been removed from the tree.
    const __half* x = x16_; const __half* w = w_conv1a_; const float* bb = b_conv1a_; __half* o = (__half*)a1_.p;
    t.push("conv1a 1->64", 0, [=](cudaStream_t st) { launch_conv1a(x, w, bb, o, B, 512, 512, st); return true; });
    t.launches++;
    if (!add_conv3x3(&t, a1_, w1b_, plnet_ && cfg_.enable_lines ? &r1_ : nullptr, &p1_, B, true)) return false;      // conv1b (+ fused pool); only PLNet's line branch reads the full-resolution map
  }
  if (!add_conv3x3(&t, p1_, w2a_, &a2_, nullptr, B, true) || !add_conv3x3(&t, a2_, w2b_, &r3, &p2_, B, true)) return false;
  if (!add_conv3x3(&t, p2_, w3a_, &a3_, nullptr, B, true) || !add_conv3x3(&t, a3_, w3b_, &r5, &p3_, B, true)) return false;
  if (!add_conv3x3(&t, p3_, w4a_, &a4_, nullptr, B, true) || !add_conv3x3(&t, a4_, w4b_, &r7_, nullptr, B, true)) return false;
  if (!add_conv3x3(&t, r7_, wPD_, &pd_, nullptr, B, true)) return false;
  if (!add_dense(&t, pd_.slice(0, 256), wPb_, logits_, B, false, 65, 80)) return false;
  if (!add_dense(&t, pd_.slice(256, 256), wDb_, descraw_, B, false)) return false;
  {
    const float* lg = (const float*)logits_.p; float* heat = heat_; float* sc = scores_; uint8_t* ma = mask_a_; uint8_t* mb = mask_b_;
    t.push("softmax_d2s+nms", 0, [=](cudaStream_t st) { launch_softmax_d2s(lg, 80, heat, B, st); launch_simple_nms(heat, sc, ma, mb, B, st); return true; });
    t.launches += getenv("AIRFE_NMS_V1") ? 4 : 2;      // softmax + fused NMS (v1: three NMS passes)
  }
  trunk_ops_[B] = std::move(t);

  if (lines) {
    OpList l;
    const bool halo = conv3x3_halo_enabled();
    const Act c2lo = cat2_.slice(0, 32), c3lo = cat3_.slice(0, 128);
    if (!add_conv3x3(&l, r1_, l1a_, &l1a_o_, nullptr, B, true)) return false;
    if (!add_conv3x3(&l, l1a_o_, l1b_, halo ? nullptr : &l1b_o_, &c2lo, B, true)) return false;       // only the pooled map is consumed
    if (!add_conv3x3(&l, cat2_, l2a_, &l2a_o_, nullptr, B, true)) return false;
    if (!add_conv3x3(&l, l2a_o_, l2b_, halo ? nullptr : &l2b_o_, &c3lo, B, true)) return false;
    Act x = cat3_;
    for (int s = 0; s < 2; ++s) {
      HGBuf& g = hgb_[s];
      Act in = x;
      const bool fuse_up = fuse_up_;     // Resize nodes fused into the producing conv's store (see the buffer allocation)
      for (int lv = 0; lv < 5; ++lv) {
        if (!add_conv3x3(&l, in, hg_[s].c[lv][0], &g.a[lv], nullptr, B, true)) return false;
        if (lv == 4 && fuse_up) { if (!add_conv3x3(&l, g.a[lv], hg_[s].c[lv][1], &g.up[0], nullptr, B, true, true)) return false; }
        else if (!add_conv3x3(&l, g.a[lv], hg_[s].c[lv][1], &g.r[lv], lv < 4 ? &g.pool[lv] : nullptr, B, true)) return false;
        if (lv < 4) in = g.pool[lv];
      }
      Act u = g.r[4];
      for (int k = 0; k < 4; ++k) {      // k = 0: 8 -> 16 (deconv1, skip = level-3 relu), ... k = 3: 64 -> 128 (skip = level-0 relu)
        if (!fuse_up) up(&l, u, g.up[k]);
        const Act clo = g.cat[k].slice(0, 64), chi = g.cat[k].slice(64, 64);
        if (!add_conv3x3(&l, g.up[k], hg_[s].dec[k], &clo, nullptr, B, true)) return false;
        if (!add_conv3x3(&l, g.r[3 - k], hg_[s].aup[k], &chi, nullptr, B, true)) return false;
        if (fuse_up && k < 3) { if (!add_conv3x3(&l, g.cat[k], hg_[s].bup[k], &g.up[k + 1], nullptr, B, true, true)) return false; }
        else if (!add_conv3x3(&l, g.cat[k], hg_[s].bup[k], &g.u[k], nullptr, B, true)) return false;
        u = g.u[k];
      }
      x = u;
    }
    if (!add_dense(&l, x, fc2_, fc2_o_, B, false)) return false;                 // no ReLU after fc2 (graph)
    if (!add_conv3x3(&l, fc2_o_, heads0_, &hmid_o_, nullptr, B, true)) return false;   // 5 x (3x3 256->64) + ReLU
    if (!add_dense(&l, hmid_o_, heads2_, heads9_o_, B, false, 9, 16)) return false;
    if (!add_dense(&l, fc2_o_, fc1_, loi_o_, B, false)) return false;
    if (!add_dense(&l, fc2_o_, fc34_, thinaux_o_, B, false, 8, 16)) return false;
    line_ops_[B] = std::move(l);
    OpList m;
    m.dyn_kind = kDynLines;
    if (!add_dense(&m, feat496_.slice(0, 496), s1_fc0_, mlp_a_, B, true, -1, 0, n_unique_)) return false;
    if (!add_dense(&m, mlp_a_, s1_fc2_, mlp_b_, B, true, -1, 0, n_unique_)) return false;
    if (!add_dense(&m, mlp_b_, s1_fc4_, mlp_c_, B, false, -1, 0, n_unique_)) return false;
    if (!add_dense(&m, feat496_.slice(256, 240), s1_res_, mlp_r_, B, true, -1, 0, n_unique_)) return false;
    mlp_ops_[B] = std::move(m);
  }
  return true;
}

bool Detector::run(const uint8_t* d_images, int B, int w, int h, int stride, long long img_stride, bool lines, bool junctions,
                   cudaStream_t st, const RemapMaps* remap) {
  if (B < 1 || B > cfg_.max_batch) { set_error("batch %d outside [1,%d]", B, cfg_.max_batch); return false; }
  lines = lines && plnet_ && cfg_.enable_lines;
  if (!ensure_tables(w, h) || !build_ops(B)) return false;
  timed("resize", st, [&] { launch_resize_u8_to_f16(d_images, w, h, stride, img_stride, B, tables_[{w, h}], x16_, nullptr, st, remap); });
  if (!trunk_ops_[B].run(st)) return false;
  const float w_scale = (float)w / 512.f, h_scale = (float)h / 512.f;
  timed("select_keypoints", st, [&] {
    launch_select_keypoints(scores_, B, cfg_.keypoint_threshold, cfg_.remove_borders, cfg_.max_keypoints, cand_, kCandCap, cand_count_, kp_,
                            kKpCap, out_.n_feat, st);
  });
  timed("sample_descriptors", st, [&] { launch_sample_descriptors(desc_raw_, kp_, out_.n_feat, kKpCap, B, w_scale, h_scale, out_.feat, st); });
  if (lines) {
    if (!line_ops_[B].run(st)) return false;
    const float* heads = (const float*)heads9_o_.p;
    timed("hafm_decode", st, [&] { launch_hafm_decode(heads, 16, lines_pred_, jloc_, B, st); });
    // scratch: peaks list reuses cand_ (ints, >= 16384 per image), is_peak reuses mask_b_
    timed("junction_topk", st, [&] { launch_junctions(jloc_, heads, 16, cand_, cand_count_, mask_b_, juncs_, junc_idx_, B, st); });
    timed("association+unique", st, [&] { launch_association(lines_pred_, juncs_, imin_, imax_, iskeep_, pair_table_, uid_pairs_, uid_first_, n_unique_, kLineCap, B, st); });
    timed("loi_gather", st, [&] {
      launch_loi_gather((const float*)loi_o_.p, (int)loi_o_.ps, (const float*)thinaux_o_.p, (int)thinaux_o_.ps, juncs_, lines_pred_, uid_pairs_, uid_first_, n_unique_, kLineCap,
                        s1_tspan_, (__half*)feat496_.p, adj_, B, st, jfeat_);
    });
    if (!mlp_ops_[B].run(st)) return false;      // stage-1 MLP on tensor cores; row counts are read on the device
    timed("line_head+accept", st, [&] {
    launch_line_head((const float*)mlp_c_.p, (const float*)mlp_r_.p, s1_head_w_, s1_head_b_, n_unique_, kLineCap, line_score_, B, st);
    launch_line_accept(adj_, line_score_, n_unique_, kLineCap, cfg_.line_threshold, cfg_.line_length_threshold,
                       cfg_.remove_borders > 0 ? cfg_.remove_borders : 0, junc_map_, out_.lines, out_.n_lines, B, st);
    });
    if (junctions) {
      launch_junction_scan(junc_map_, scores_, cfg_.remove_borders > 0 ? cfg_.remove_borders : 0, jkp_, kKpCap, jkp_count_, B, st);
      launch_sample_descriptors(desc_raw_, jkp_, jkp_count_, kKpCap, B, w_scale, h_scale, out_.junc, st);
      AIRFE_CUDA_OK(cudaMemcpyAsync(out_.n_junc, jkp_count_, sizeof(int) * B, cudaMemcpyDeviceToDevice, st));
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("detector launch error: %s", cudaGetErrorString(e)); return false; }
  return true;
}

double Detector::tc_flops(int batch, bool lines) {
  if (!build_ops(batch)) return 0;
  return trunk_ops_[batch].tc_flops + ((lines && line_ops_.count(batch)) ? line_ops_[batch].tc_flops : 0.0);
}
int Detector::launches(int batch, bool lines) {
  if (!build_ops(batch)) return 0;
  return trunk_ops_[batch].launches + ((lines && line_ops_.count(batch)) ? line_ops_[batch].launches : 0);
}

}  // namespace airfe
