#include "matcher.h"
#include "match_kernels.h"

namespace airfe {

bool SuperGlue::init(const MatcherConfig& cfg, const std::string& wdir, bool outdoor) {
  cfg_ = cfg;
  if (cfg.cap % 128 || cfg.cap > 1024 || cfg.cap < 128) { set_error("matcher cap %d must be a multiple of 128 <= 1024", cfg.cap); return false; }
  WeightFile wf;
  if (!wf.load(wdir + (outdoor ? "/superglue_outdoor.afw" : "/superglue_indoor.afw"))) return false;
  const int S = 2 * cfg.max_pairs, cap = cfg.cap;
  const size_t R = (size_t)S * cap;
  const size_t ld = cap + 1;
  size_t bytes = (size_t)64 << 20;
  bytes += R * (256 * 4 + 64 * 2 + 256 * 2 * 2 + 512 * 2 + 768 * 2 + 256 * 2 + 512 * 2 + 256 * 2 + 64);
  bytes += (size_t)S * 4 * cap * cap * 6 + (size_t)cfg.max_pairs * (cap * cap * 4 + ld * ld * 8 + ld * 16 + cap * 48) + (1 << 20);
  if (!arena_.init(bytes)) return false;
  Arena* ar = &arena_;
  // head layout of the graph: channel c -> (d = c / 4, h = c % 4)  ==> head-major position h*64 + d
  std::vector<int> hp(256), hp3(768);
  for (int c = 0; c < 256; ++c) hp[c] = (c % 4) * 64 + c / 4;
  for (int s = 0; s < 3; ++s) for (int c = 0; c < 256; ++c) hp3[s * 256 + c] = s * 256 + hp[c];
  const int kin[5] = {3, 32, 64, 128, 256};
  const int kidx[5] = {0, 3, 6, 9, 12};
  for (int i = 0; i < 5; ++i) {
    const std::string nm = "sg.kenc.encoder." + std::to_string(kidx[i]);
    if (!pack_dense(wf, {{nm + ".weight", nm + ".bias", 0}}, kin[i], ar, &kenc_[i])) return false;
  }
  for (int l = 0; l < 18; ++l) {
    const std::string T = "sg.gnn.layers." + std::to_string(l) + ".";
    if (!pack_dense(wf, {{T + "attn.proj.0.weight", T + "attn.proj.0.bias", 0}, {T + "attn.proj.1.weight", T + "attn.proj.1.bias", 0},
                         {T + "attn.proj.2.weight", T + "attn.proj.2.bias", 0}}, 256, ar, &L_[l].qkv, nullptr, &hp3))
      return false;
    if (!pack_dense(wf, {{T + "attn.merge.weight", T + "attn.merge.bias", 0}}, 256, ar, &L_[l].merge, &hp)) return false;
    if (!pack_dense(wf, {{T + "mlp.0.weight", T + "mlp.0.bias", 0}}, 512, ar, &L_[l].mlp0)) return false;
    if (!pack_dense(wf, {{T + "mlp.3.weight", T + "mlp.3.bias", 0}}, 512, ar, &L_[l].mlp3)) return false;
  }
  if (!pack_dense(wf, {{"sg.final_proj.weight", "sg.final_proj.bias", 0}}, 256, ar, &final_)) return false;
  {
    std::vector<float> b;
    if (!wf.get_f32("sg.bin_score", &b)) return false;
    bin_score_ = b[0];
  }
  x_ = ar->alloc_n<float>(R * 256);
  kin16_ = ar->alloc_n<__half>(R * 64);
  k1_ = ar->alloc_n<__half>(R * 256);
  k2_ = ar->alloc_n<__half>(R * 256);
  cat16_ = ar->alloc_n<__half>(R * 512);
  qkv16_ = ar->alloc_n<__half>(R * 768);
  ctx16_ = ar->alloc_n<__half>(R * 256);
  h16_ = ar->alloc_n<__half>(R * 512);
  md16_ = ar->alloc_n<__half>(R * 256);
  S_ = ar->alloc_n<float>((size_t)S * 4 * cap * cap);
  P_ = ar->alloc_n<__half>((size_t)S * 4 * cap * cap);
  sim_ = ar->alloc_n<float>((size_t)cfg.max_pairs * cap * cap);
  Z_ = ar->alloc_n<float>((size_t)cfg.max_pairs * ld * ld);
  out_.dense = ar->alloc_n<float>((size_t)cfg.max_pairs * ld * ld);
  u_ = ar->alloc_n<float>((size_t)cfg.max_pairs * ld);
  v_ = ar->alloc_n<float>((size_t)cfg.max_pairs * ld);
  const size_t PC = (size_t)cfg.max_pairs * cap;
  arg0_ = ar->alloc_n<int>(PC); arg1_ = ar->alloc_n<int>(PC); val0_ = ar->alloc_n<float>(PC);
  out_.idx0 = ar->alloc_n<int>(PC); out_.idx1 = ar->alloc_n<int>(PC); out_.ms0 = ar->alloc_n<float>(PC); out_.ms1 = ar->alloc_n<float>(PC);
  out_.m_idx = ar->alloc_n<int>(PC * 2); out_.m_score = ar->alloc_n<float>(PC); out_.m_count = ar->alloc_n<int>(cfg.max_pairs);
  n_ = ar->alloc_n<int>(S);
  return ar->ok();
}

bool SuperGlue::build_ops(int P) {
  if (ops_.count(P)) return true;
  OpList ol;
  const int S = 2 * P, cap = cfg_.cap;
  auto rows = [&](void* p, int C, int ps, bool f32) { Act a; a.p = p; a.C = C; a.H = 1; a.W = cap; a.ps = ps; a.f32 = f32; return a; };
  const Act kin = rows(kin16_, 64, 64, false), k1 = rows(k1_, 256, 256, false), k2 = rows(k2_, 256, 256, false);
  const Act x16 = rows(cat16_, 256, 512, false), msg16 = rows(cat16_ + 256, 256, 512, false), cat = rows(cat16_, 512, 512, false);
  const Act xf = rows(x_, 256, 256, true), qkv = rows(qkv16_, 768, 768, false), ctx = rows(ctx16_, 256, 256, false);
  const Act h16 = rows(h16_, 512, 512, false), md = rows(md16_, 256, 256, false);
  const int* n = n_;
  // keypoint encoder 3 -> 32 -> 64 -> 128 -> 256 -> 256, then desc += kenc (residual epilogue), fp16 operand copy
  Act a = kin, b;
  Act tmp[2] = {k1, k2};
  for (int i = 0; i < 4; ++i) {
    b = tmp[i & 1];
    b.C = kenc_[i].n_rows;
    if (!add_dense(&ol, a, kenc_[i], b, S, true, -1, 0, n)) return false;
    a = b;
  }
  if (!add_dense(&ol, a, kenc_[4], xf, S, false, -1, 0, n, 1.f, x_, &x16)) return false;

  for (int l = 0; l < 18; ++l) {
    const int xr = l & 1;   // names = ['self','cross'] * 9
    Layer& Y = L_[l];
    // q (scaled by 1/8 = 1/sqrt(64), exact in fp16), k, v in head-major layout
    if (!add_dense(&ol, x16, Y.qkv, qkv, S, false, -1, 0, n)) return false;
    if (attn_fused_enabled() && cap <= 512) {
      if (!add_fused_attention(&ol, qkv16_, qkv16_ + 256, qkv16_ + 512, 768, ctx16_, n, S, cap, xr, 0.125f)) return false;
      if (ffn_fused_enabled()) {          // merge + mlp.0 + ReLU + mlp.3 + residual in one kernel (tc_ffn.cuh, ReLU variant)
        if (!add_fused_ffn(&ol, ctx16_, cat16_, x_, Y.merge, Y.mlp0, Y.mlp3, nullptr, nullptr, n, S, cap, true)) return false;
        continue;
      }
      if (!add_dense(&ol, ctx, Y.merge, msg16, S, false, -1, 0, n)) return false;
      if (!add_dense(&ol, cat, Y.mlp0, h16, S, true, -1, 0, n)) return false;
      if (!add_dense(&ol, h16, Y.mlp3, xf, S, false, -1, 0, n, 1.f, x_, &x16)) return false;
      continue;
    }
    TcGemmDesc d;
    d.a = qkv16_; d.a_C = 64; d.W = cap; d.H = 4; d.B = S; d.a_sx = 768; d.a_sy = 64; d.a_sb = (long long)cap * 768;
    d.bw = qkv16_ + 256; d.k_total = 64; d.n_rows = cap; d.bw_sn = 768; d.b_heads = 4; d.bw_shead = 64; d.b_batches = S; d.bw_sbatch = (long long)cap * 768;
    d.b_batch_xor = xr; d.taps = 1; d.c_in_pad = 64; d.block_n = 128; d.out_f32 = 1; d.out = S_;
    d.scale = 0.125f; d.scale_cols = cap;   // scores / 8 (power of two: commutes with the operand rounding)
    d.out_sb = 4ll * cap * cap; d.out_sy = (long long)cap * cap; d.out_sx = cap; d.n_valid = cap; d.tw = 128; d.th = 1; d.tb = 1; d.dyn_w = n;
    if (!add_gemm(&ol, d, 2.0 * S * 4 * (double)cap * cap * 64, xr ? kDynAttCross : kDynAttSelf)) return false;
    {
      const float* Sp = S_; __half* Pp = P_;
      ol.push("softmax_rows", 0, [=](cudaStream_t st) { launch_softmax_rows(Sp, Pp, n, S, cap, xr, st); return true; });
      ol.launches++;
    }
    TcGemmDesc e;
    e.a = P_; e.a_C = cap; e.W = cap; e.H = 4; e.B = S; e.a_sx = cap; e.a_sy = (long long)cap * cap; e.a_sb = 4ll * cap * cap;
    e.bw = qkv16_ + 512; e.k_total = cap; e.n_rows = 64; e.bw_sn = 768; e.b_heads = 4; e.bw_shead = 64; e.b_batches = S; e.bw_sbatch = (long long)cap * 768;
    e.b_batch_xor = xr; e.b_mn_major = 1; e.taps = 1; e.c_in_pad = cap; e.block_n = 64; e.out_f32 = 0; e.out = ctx16_;
    e.out_sb = (long long)cap * 256; e.out_sy = 64; e.out_sx = 256; e.n_valid = 64; e.tw = 128; e.th = 1; e.tb = 1; e.dyn_w = n;
    if (!add_gemm(&ol, e, 2.0 * S * 4 * (double)cap * cap * 64, xr ? kDynAttCross : kDynAttSelf)) return false;
    if (!add_dense(&ol, ctx, Y.merge, msg16, S, false, -1, 0, n)) return false;
    if (!add_dense(&ol, cat, Y.mlp0, h16, S, true, -1, 0, n)) return false;
    if (!add_dense(&ol, h16, Y.mlp3, xf, S, false, -1, 0, n, 1.f, x_, &x16)) return false;
  }
  if (!add_dense(&ol, x16, final_, md, S, false, -1, 0, n, 0.25f)) return false;   // scores / 16 = (0.25 d0) . (0.25 d1)
  {
    TcGemmDesc d;
    d.a = md16_; d.a_C = 256; d.W = cap; d.H = 1; d.B = P; d.a_sx = 256; d.a_sy = 0; d.a_sb = 2ll * cap * 256;
    d.bw = md16_ + (size_t)cap * 256; d.k_total = 256; d.n_rows = cap; d.bw_sn = 256; d.b_batches = P > 1 ? P : 0; d.bw_sbatch = 2ll * cap * 256;
    d.taps = 1; d.c_in_pad = 256; d.block_n = 128; d.out_f32 = 1; d.out = sim_; d.out_sb = (long long)cap * cap; d.out_sy = 0; d.out_sx = cap;
    d.n_valid = cap; d.tw = 128; d.th = 1; d.tb = 1; d.dyn_w = n; d.dyn_w_stride = 2;
    if (!add_gemm(&ol, d, 2.0 * P * (double)cap * cap * 256, kDynSim)) return false;
  }
  ops_[P] = std::move(ol);
  return true;
}

bool SuperGlue::run(const float* d_feat, const int* d_n, int feat_cap, int P, bool want_dense, cudaStream_t st, bool prenorm, const float* const* d_feat_ptrs, int max_n) {
  if (P < 1 || P > cfg_.max_pairs) { set_error("pairs %d outside [1,%d]", P, cfg_.max_pairs); return false; }
  if (!build_ops(P)) return false;
  const int S = 2 * P, cap = cfg_.cap;
  AIRFE_CUDA_OK(cudaMemcpyAsync(n_, d_n, sizeof(int) * S, cudaMemcpyDeviceToDevice, st));
  // scale = 0.7 for SuperGlue (src/point_matcher.cc:58)
  const float l_inv = (float)(1.0 / (double)(cfg_.image_width > cfg_.image_height ? cfg_.image_width : cfg_.image_height) * (double)0.7f);
  timed("sg_prepare", st, [&] { launch_sg_prepare(d_feat, d_feat_ptrs, n_, S, cap, feat_cap, prenorm ? 0 : cfg_.image_width, prenorm ? 0 : cfg_.image_height, prenorm ? 1.f : l_inv, x_, kin16_, st); });
  if (!ops_[P].run(st)) return false;
  timed("sg_sinkhorn+decode", st, [&] {
    launch_sg_sinkhorn_decode(sim_, n_, P, cap, max_n, bin_score_, 100, Z_, u_, v_, 0.2f, arg0_, val0_, arg1_, out_.idx0, out_.idx1, out_.ms0, out_.ms1,
                              out_.m_idx, out_.m_score, out_.m_count, want_dense ? out_.dense : nullptr, st);
  });
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("superglue launch error: %s", cudaGetErrorString(e)); return false; }
  return true;
}

}  // namespace airfe
