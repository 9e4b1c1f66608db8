#include "matcher.h"
#include <stdlib.h>
#include "match_kernels.h"

#include <math.h>

namespace airfe {

static bool up_f32(Arena* ar, const WeightFile& wf, const std::string& name, float** out) {
  std::vector<float> v;
  if (!wf.get_f32(name, &v)) return false;
  *out = ar->alloc_n<float>(v.size());
  if (!*out) return false;
  AIRFE_CUDA_OK(cudaMemcpy(*out, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  return true;
}

bool LightGlue::init(const MatcherConfig& cfg, const std::string& wdir) {
  cfg_ = cfg;
  if (cfg.cap % 128 || cfg.cap > 1024 || cfg.cap < 128) { set_error("matcher cap %d must be a multiple of 128 <= 1024", cfg.cap); return false; }
  WeightFile wf;
  if (!wf.load(wdir + "/lightglue.afw")) return false;
  const int S = 2 * cfg.max_pairs, cap = cfg.cap;
  const size_t R = (size_t)S * cap;
  size_t bytes = (size_t)64 << 20;                                  // weights (23 MB fp16) + slack
  bytes += R * (256 * 4 + 512 * 2 + 768 * 4 + 4 * 256 * 2 + 512 * 4 + 512 * 2 + 64 * 4 + 256 * 2 + 256 * 2 + 64);
  bytes += (size_t)S * 4 * cap * cap * 6 + (size_t)cfg.max_pairs * cap * cap * 8 + (size_t)cfg.max_pairs * cap * 32 + (1 << 20);
  if (!arena_.init(bytes)) return false;
  Arena* ar = &arena_;
  // Wqkv rows: graph order is feature = h*192 + d*3 + s (s in q,k,v); repack to [q | k | v] with head-major 64-wide blocks
  std::vector<int> perm(768);
  for (int h = 0; h < 4; ++h) for (int d = 0; d < 64; ++d) for (int s = 0; s < 3; ++s) perm[h * 192 + d * 3 + s] = s * 256 + h * 64 + d;
  for (int l = 0; l < 9; ++l) {
    const std::string T = "lg.transformers." + std::to_string(l) + ".";
    auto pk = [&](const std::string& n, int cin, DenseW* o, const std::vector<int>* op = nullptr) {
      return pack_dense(wf, {{T + n + ".weight", T + n + ".bias", 0}}, cin, ar, o, nullptr, op);
    };
    Layer& Y = L_[l];
    if (!pk("self_attn.Wqkv", 256, &Y.qkv, &perm) || !pk("self_attn.out_proj", 256, &Y.out) || !pk("self_attn.ffn.0", 512, &Y.ffn0) ||
        !pk("self_attn.ffn.3", 512, &Y.ffn3) || !pk("cross_attn.to_qk", 256, &Y.c_qk) || !pk("cross_attn.to_v", 256, &Y.c_v) ||
        !pk("cross_attn.to_out", 256, &Y.c_out) || !pk("cross_attn.ffn.0", 512, &Y.c_ffn0) || !pk("cross_attn.ffn.3", 512, &Y.c_ffn3))
      return false;
    {   // cross attention: to_qk and to_v read the same input -> one [512 x 256] GEMM whose two column sections go to q16_ / v16_
      Y.c_qv = Y.c_qk;
      Y.c_qv.n_rows = 512;
      Y.c_qv.w = ar->alloc_n<__half>(512 * 256);
      Y.c_qv.bias = ar->alloc_n<float>(512);
      if (!ar->ok()) return false;
      AIRFE_CUDA_OK(cudaMemcpy(Y.c_qv.w, Y.c_qk.w, 256 * 256 * 2, cudaMemcpyDeviceToDevice));
      AIRFE_CUDA_OK(cudaMemcpy(Y.c_qv.w + 256 * 256, Y.c_v.w, 256 * 256 * 2, cudaMemcpyDeviceToDevice));
      AIRFE_CUDA_OK(cudaMemcpy(Y.c_qv.bias, Y.c_qk.bias, 256 * 4, cudaMemcpyDeviceToDevice));
      AIRFE_CUDA_OK(cudaMemcpy(Y.c_qv.bias + 256, Y.c_v.bias, 256 * 4, cudaMemcpyDeviceToDevice));
    }
    if (!up_f32(ar, wf, T + "self_attn.ffn.1.weight", &Y.ln_g) || !up_f32(ar, wf, T + "self_attn.ffn.1.bias", &Y.ln_b) ||
        !up_f32(ar, wf, T + "cross_attn.ffn.1.weight", &Y.c_ln_g) || !up_f32(ar, wf, T + "cross_attn.ffn.1.bias", &Y.c_ln_b))
      return false;
  }
  if (!pack_dense(wf, {{"lg.log_assignment.8.final_proj.weight", "lg.log_assignment.8.final_proj.bias", 0}}, 256, ar, &final_)) return false;
  {
    const WTensor* t = wf.find("lg.posenc.Wr.weight");
    const WTensor* m = wf.find("lg.log_assignment.8.matchability.weight");
    std::vector<float> mb;
    if (!t || !m || !wf.get_f32("lg.log_assignment.8.matchability.bias", &mb)) { set_error("missing LightGlue tensors"); return false; }
    wr_ = ar->alloc_n<__half>(64);
    wm_ = ar->alloc_n<__half>(256);
    AIRFE_CUDA_OK(cudaMemcpy(wr_, t->data, 64 * 2, cudaMemcpyHostToDevice));
    AIRFE_CUDA_OK(cudaMemcpy(wm_, m->data, 256 * 2, cudaMemcpyHostToDevice));
    bm_ = mb[0];
  }
  x_ = ar->alloc_n<float>(R * 256);
  cat16_ = ar->alloc_n<__half>(R * 512);
  qkv_ = ar->alloc_n<float>(R * 768);
  q16_ = ar->alloc_n<__half>(R * 768);        // q | k | v as three [R][256] matrices, R*256 elements apart (split-column GEMM stores)
  k16_ = q16_ + R * 256;
  v16_ = q16_ + 2 * R * 256;
  ctx16_ = ar->alloc_n<__half>(R * 256);
  md16_ = ar->alloc_n<__half>(R * 256);
  h_ = ar->alloc_n<float>(R * 512);
  h16_ = ar->alloc_n<__half>(R * 512);
  rot_ = ar->alloc_n<float>(R * 64);
  S_ = ar->alloc_n<float>((size_t)S * 4 * cap * cap);
  P_ = ar->alloc_n<__half>((size_t)S * 4 * cap * cap);
  sim_ = ar->alloc_n<float>((size_t)cfg.max_pairs * cap * cap);
  out_.dense = ar->alloc_n<float>((size_t)cfg.max_pairs * cap * cap);
  logsig_ = ar->alloc_n<float>(R);
  lse_ = ar->alloc_n<float>(R);
  row_val_ = ar->alloc_n<float>((size_t)cfg.max_pairs * cap);
  row_arg_ = ar->alloc_n<int>((size_t)cfg.max_pairs * cap);
  col_arg_ = ar->alloc_n<int>((size_t)cfg.max_pairs * cap);
  n_ = ar->alloc_n<int>(S);
  row_off_ = ar->alloc_n<int>(S + 1);
  md16_pad_ = ar->alloc_n<__half>(R * 256);
  // Packed rows (slot s at rows [row_off[s], +n[s]) instead of [s * cap, ...)): the row GEMMs, the fused block tails and the attention row
  // tiles then cover sum(n) rows instead of slots x cap -- at 400 keypoints 200 instead of 256 row tiles per 32 pairs.  Needs the fused
  // attention / FFN kernels (cap <= 512); AIRFE_LG_PADDED=1 keeps the slot-padded layout for A/B runs.
  packed_ = attn_fused_enabled() && ffn_fused_enabled() && cap <= 512 && getenv("AIRFE_LG_PADDED") == nullptr && getenv("AIRFE_LG_UNFUSED_PROJ") == nullptr;
  out_.idx = ar->alloc_n<int>((size_t)cfg.max_pairs * cap * 2);
  out_.score = ar->alloc_n<float>((size_t)cfg.max_pairs * cap);
  out_.count = ar->alloc_n<int>(cfg.max_pairs);
  return ar->ok();
}

bool LightGlue::build_ops(int P) {
  if (ops_.count(P)) return true;
  OpList ol;
  const int S = 2 * P, cap = cfg_.cap;
  // packed: ONE matrix of S * cap rows whose valid extent row_off[S] is read on the device; padded: S slots of cap rows with n[s] valid each
  const int RW = packed_ ? S * cap : cap;      // rows per "batch" of the row GEMMs
  const int RB = packed_ ? 1 : S;              // batches
  auto rows = [&](void* p, int C, int ps, bool f32) { Act a; a.p = p; a.C = C; a.H = 1; a.W = RW; a.ps = ps; a.f32 = f32; return a; };
  const Act x16 = rows(cat16_, 256, 512, false), msg16 = rows(cat16_ + 256, 256, 512, false), cat = rows(cat16_, 512, 512, false);
  const Act xf = rows(x_, 256, 256, true), qkv = rows(qkv_, 768, 768, true), q16 = rows(q16_, 256, 256, false), v16 = rows(v16_, 256, 256, false);
  const Act ctx = rows(ctx16_, 256, 256, false), hf = rows(h_, 512, 512, true), h16 = rows(h16_, 512, 512, false), md = rows(md16_, 256, 256, false);
  const float sc = 0.35355339059327379f;   // 64^-1/4
  const int* n = n_;
  const int* nrow = packed_ ? row_off_ + S : n_;                    // valid rows per batch of the row GEMMs (device)
  const int* roff = packed_ ? row_off_ : nullptr;
  const size_t Rall = (size_t)2 * cfg_.max_pairs * cap;             // rows of the q / k / v matrices as allocated
  static const bool fused_proj = getenv("AIRFE_LG_UNFUSED_PROJ") == nullptr;

  auto attention = [&](const __half* qa, const __half* kb, const __half* vb, int xr) -> bool {
    if (attn_fused_enabled() && cap <= 512) return add_fused_attention(&ol, qa, kb, vb, 256, ctx16_, n, S, cap, xr, 1.f, roff);
    // unfused path (cap > 512): S[s][h] = Q_s,h . K_(s^xr),h^T (fp32 in HBM), softmax rows, ctx_s = P . V_(s^xr)
    TcGemmDesc d;
    d.a = qa; d.a_C = 64; d.W = cap; d.H = 4; d.B = S; d.a_sx = 256; d.a_sy = 64; d.a_sb = (long long)cap * 256;
    d.bw = kb; d.k_total = 64; d.n_rows = cap; d.bw_sn = 256; d.b_heads = 4; d.bw_shead = 64; d.b_batches = S; d.bw_sbatch = (long long)cap * 256;
    d.b_batch_xor = xr; d.taps = 1; d.c_in_pad = 64; d.block_n = 128; d.out_f32 = 1; d.out = S_;
    d.out_sb = 4ll * cap * cap; d.out_sy = (long long)cap * cap; d.out_sx = cap; d.n_valid = cap; d.tw = 128; d.th = 1; d.tb = 1; d.dyn_w = n;
    if (!add_gemm(&ol, d, 2.0 * S * 4 * (double)cap * cap * 64, xr ? kDynAttCross : kDynAttSelf)) return false;
    {
      const float* Sp = S_; __half* Pp = P_;
      ol.push("softmax_rows", 0, [=](cudaStream_t st) { launch_softmax_rows(Sp, Pp, n, S, cap, xr, st); return true; });
      ol.launches++;
    }
    TcGemmDesc e;
    e.a = P_; e.a_C = cap; e.W = cap; e.H = 4; e.B = S; e.a_sx = cap; e.a_sy = (long long)cap * cap; e.a_sb = 4ll * cap * cap;
    e.bw = vb; e.k_total = cap; e.n_rows = 64; e.bw_sn = 256; e.b_heads = 4; e.bw_shead = 64; e.b_batches = S; e.bw_sbatch = (long long)cap * 256;
    e.b_batch_xor = xr; e.b_mn_major = 1; e.taps = 1; e.c_in_pad = cap; e.block_n = 64; e.out_f32 = 0; e.out = ctx16_;
    e.out_sb = (long long)cap * 256; e.out_sy = 64; e.out_sx = 256; e.n_valid = 64; e.tw = 128; e.th = 1; e.tb = 1; e.dyn_w = n;
    return add_gemm(&ol, e, 2.0 * S * 4 * (double)cap * cap * 64, xr ? kDynAttCross : kDynAttSelf);
  };
  auto ffn = [&](const DenseW& w_out, const DenseW& w0, const DenseW& w3, const float* g, const float* b) -> bool {
    if (ffn_fused_enabled()) return add_fused_ffn(&ol, ctx16_, cat16_, x_, w_out, w0, w3, g, b, nrow, RB, RW);
    if (!add_dense(&ol, ctx, w_out, msg16, S, false, -1, 0, n)) return false;                     // msg -> [x | msg] operand buffer
    if (!add_dense(&ol, cat, w0, hf, S, false, -1, 0, n)) return false;
    {
      const float* hp = h_; __half* ho = h16_;
      ol.push("ln_gelu", 0, [=](cudaStream_t st) { launch_ln_gelu(hp, g, b, n, S, cap, ho, st); return true; });
      ol.launches++;
    }
    return add_dense(&ol, h16, w3, xf, S, false, -1, 0, n, 1.f, x_, &x16);                       // x += ffn ; fp16 operand copy
  };
  for (int l = 0; l < 9; ++l) {
    Layer& Y = L_[l];
    if (fused_proj) {
      // Wqkv with the rotary embedding, the 64^-1/4 scaling of q,k and the fp16 rounding in the GEMM epilogue; q | k | v land in their
      // own matrices.  (The unfused path wrote 50 MB of fp32 qkv and re-read it in a separate rotary kernel.)
      DenseExtra ex;
      ex.scale_cols = 512; ex.rot = rot_; ex.rot_cols = 512; ex.out_split = 256; ex.out_split_stride = (long long)Rall * 256;
      if (!add_dense(&ol, x16, Y.qkv, rows(q16_, 768, 256, false), RB, false, -1, 0, nrow, sc, nullptr, nullptr, &ex)) return false;
    } else {
      if (!add_dense(&ol, x16, Y.qkv, qkv, S, false, -1, 0, n)) return false;
      const float* qp = qkv_; const float* rp = rot_; __half* q = q16_; __half* k = k16_; __half* v = v16_;
      ol.push("rotary", 0, [=](cudaStream_t st) { launch_lg_rotary(qp, rp, n, S, cap, q, k, v, st); return true; });
      ol.launches++;
    }
    if (!attention(q16_, k16_, v16_, 0) || !ffn(Y.out, Y.ffn0, Y.ffn3, Y.ln_g, Y.ln_b)) return false;
    if (fused_proj) {
      DenseExtra ex;
      ex.scale_cols = 256; ex.out_split = 256; ex.out_split_stride = (long long)Rall * 512;   // section 0 -> q16_, section 1 -> v16_
      if (!add_dense(&ol, x16, Y.c_qv, rows(q16_, 512, 256, false), RB, false, -1, 0, nrow, sc, nullptr, nullptr, &ex)) return false;
    } else {
      if (!add_dense(&ol, x16, Y.c_qk, q16, S, false, -1, 0, n, sc)) return false;
      if (!add_dense(&ol, x16, Y.c_v, v16, S, false, -1, 0, n)) return false;
    }
    if (!attention(q16_, q16_, v16_, 1) || !ffn(Y.c_out, Y.c_ffn0, Y.c_ffn3, Y.c_ln_g, Y.c_ln_b)) return false;
  }
  if (!add_dense(&ol, x16, final_, md, RB, false, -1, 0, nrow, 0.25f)) return false;             // / 256^(1/4)
  {
    // matchability logits of the final state + (packed layout) the slot-padded copy of md the similarity GEMM reads
    const float* xs = x_; const __half* wm = wm_; const float bm = bm_; float* ls = logsig_; const __half* mdp = md16_; __half* mdo = packed_ ? md16_pad_ : nullptr;
    const int* ro = roff;
    ol.push("lg_matchability", 0, [=](cudaStream_t st) { launch_lg_matchability(xs, wm, bm, n, ro, P, cap, ls, mdp, mdo, st); return true; });
    ol.launches++;
  }
  {
    const __half* mdsrc = packed_ ? md16_pad_ : md16_;
    TcGemmDesc d;
    d.a = mdsrc; d.a_C = 256; d.W = cap; d.H = 1; d.B = P; d.a_sx = 256; d.a_sy = 0; d.a_sb = 2ll * cap * 256;
    d.bw = mdsrc + (size_t)cap * 256; d.k_total = 256; d.n_rows = cap; d.bw_sn = 256; d.b_batches = P > 1 ? P : 0; d.bw_sbatch = 2ll * cap * 256;
    d.taps = 1; d.c_in_pad = 256; d.block_n = 128; d.out_f32 = 1; d.out = sim_; d.out_sb = (long long)cap * cap; d.out_sy = 0; d.out_sx = cap;
    d.n_valid = cap; d.tw = 128; d.th = 1; d.tb = 1; d.dyn_w = n; d.dyn_w_stride = 2;
    if (!add_gemm(&ol, d, 2.0 * P * (double)cap * cap * 256, kDynSim)) return false;
  }
  ops_[P] = std::move(ol);
  return true;
}

bool LightGlue::run(const float* d_feat, const int* d_n, int feat_cap, int P, bool want_dense, cudaStream_t st, bool prenorm, const float* const* d_feat_ptrs) {
  if (P < 1 || P > cfg_.max_pairs) { set_error("pairs %d outside [1,%d]", P, cfg_.max_pairs); return false; }
  if (!build_ops(P)) return false;
  const int S = 2 * P, cap = cfg_.cap;
  AIRFE_CUDA_OK(cudaMemcpyAsync(n_, d_n, sizeof(int) * S, cudaMemcpyDeviceToDevice, st));
  // float L_inv = 1.0 / std::max(width, height) * scale;  scale = 0.5 for LightGlue (src/point_matcher.cc:43,58)
  const float l_inv = (float)(1.0 / (double)(cfg_.image_width > cfg_.image_height ? cfg_.image_width : cfg_.image_height) * (double)0.5f);
  // prenormalised input: (x - 0) * 1 reproduces the value bit for bit
  timed("lg_prepare", st, [&] { launch_lg_prepare(d_feat, d_feat_ptrs, n_, packed_ ? row_off_ : nullptr, S, cap, feat_cap, prenorm ? 0 : cfg_.image_width, prenorm ? 0 : cfg_.image_height, prenorm ? 1.f : l_inv, wr_, x_, cat16_, rot_, st); });
  if (!ops_[P].run(st)) return false;
  timed("lg_assignment+filter", st, [&] {
    launch_lg_assignment(sim_, n_, P, cap, logsig_, lse_, row_arg_, row_val_, col_arg_, 0.1f, out_.idx, out_.score, out_.count,
                         want_dense ? out_.dense : nullptr, st);
  });
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("matcher launch error: %s", cudaGetErrorString(e)); return false; }
  return true;
}

double LightGlue::tc_flops(int P) { return build_ops(P) ? ops_[P].tc_flops : 0.0; }
int LightGlue::launches(int P) { return build_ops(P) ? ops_[P].launches + 8 : 0; }

}  // namespace airfe
