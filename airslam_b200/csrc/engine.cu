#include "engine.h"

#include <algorithm>

namespace airfe {

// ---- weight container ------------------------------------------------------------------------------------------------
bool WeightFile::load(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { set_error("cannot open weight file %s", path.c_str()); return false; }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  blob_.resize((size_t)sz);
  size_t rd = fread(blob_.data(), 1, (size_t)sz, f);
  fclose(f);
  if (rd != (size_t)sz || sz < 16 || memcmp(blob_.data(), "AIRFEW01", 8) != 0) { set_error("bad weight file %s", path.c_str()); return false; }
  uint32_t count;
  memcpy(&count, blob_.data() + 8, 4);
  const size_t size = blob_.size();
  if ((size - 16) / 160 < (size_t)count) { set_error("weight file %s: entry table (%u entries) exceeds the file", path.c_str(), count); return false; }
  static const size_t esz[3] = {4, 2, 4};   // f32, f16, i32
  for (uint32_t i = 0; i < count; ++i) {
    const uint8_t* e = blob_.data() + 16 + (size_t)160 * i;
    char name[121];
    memcpy(name, e, 120);
    name[120] = 0;
    WTensor t;
    uint32_t dt, nd, dims[4];
    uint64_t off, nb;
    memcpy(&dt, e + 120, 4); memcpy(&nd, e + 124, 4); memcpy(dims, e + 128, 16); memcpy(&off, e + 144, 8); memcpy(&nb, e + 152, 8);
    if (dt > 2 || nd > 4) { set_error("weight file %s: tensor %s has dtype %u / rank %u", path.c_str(), name, dt, nd); return false; }
    t.dtype = (int)dt; t.ndim = (int)nd;
    uint64_t numel = 1;
    for (int k = 0; k < 4; ++k) {
      t.dims[k] = k < (int)nd ? (int)dims[k] : 1;
      if (k < (int)nd && (dims[k] == 0 || dims[k] > (1u << 28))) { set_error("weight file %s: tensor %s has a bad extent", path.c_str(), name); return false; }
      numel *= (uint64_t)t.dims[k];
      if (numel > ((uint64_t)1 << 40)) { set_error("weight file %s: tensor %s is too large", path.c_str(), name); return false; }
    }
    if (off > size || nb > size - off) { set_error("weight file %s truncated (tensor %s)", path.c_str(), name); return false; }   // no off + nb overflow
    if (numel * esz[dt] != nb) { set_error("weight file %s: tensor %s: %llu elements do not fill %llu bytes", path.c_str(), name, (unsigned long long)numel, (unsigned long long)nb); return false; }
    t.data = blob_.data() + off;
    t.nbytes = nb;
    index_[name] = t;
  }
  return true;
}

const WTensor* WeightFile::find(const std::string& name) const {
  auto it = index_.find(name);
  return it == index_.end() ? nullptr : &it->second;
}

static inline float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }

bool WeightFile::get_f32(const std::string& name, std::vector<float>* out) const {
  const WTensor* t = find(name);
  if (!t) { set_error("missing tensor %s", name.c_str()); return false; }
  out->resize(t->numel());
  if (t->dtype == 0) memcpy(out->data(), t->data, t->numel() * 4);
  else if (t->dtype == 1) { const uint16_t* h = (const uint16_t*)t->data; for (size_t i = 0; i < t->numel(); ++i) (*out)[i] = h2f(h[i]); }
  else { const int32_t* v = (const int32_t*)t->data; for (size_t i = 0; i < t->numel(); ++i) (*out)[i] = (float)v[i]; }
  return true;
}

// ---- arena --------------------------------------------------------------------------------------------------------------
Arena::~Arena() { if (base_) cudaFree(base_); }
bool Arena::init(size_t bytes) {
  if (cudaMalloc(&base_, bytes) != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError())); return false; }
  cap_ = bytes;
  cudaMemset(base_, 0, bytes);
  return true;
}
void* Arena::alloc(size_t bytes, size_t align) {
  size_t o = (off_ + align - 1) / align * align;
  if (o + bytes > cap_) { failed_ = true; set_error("device arena exhausted (%zu + %zu > %zu)", o, bytes, cap_); return nullptr; }
  off_ = o + bytes;
  return base_ + o;
}

// ---- packing --------------------------------------------------------------------------------------------------------------
bool pack_dense(const WeightFile& wf, const std::vector<PackSrc>& srcs, int c_in_total, Arena* arena, DenseW* out,
                const std::vector<int>* in_perm, const std::vector<int>* out_perm, int k_align) {
  int n_rows = 0, taps = 0;
  for (auto& s : srcs) {
    const WTensor* t = wf.find(s.weight);
    if (!t || t->dtype != 1) { set_error("pack_dense: missing fp16 tensor %s", s.weight.c_str()); return false; }
    const int tp = t->ndim == 4 ? t->dims[2] * t->dims[3] : 1;
    if (taps && tp != taps) { set_error("pack_dense: mixed kernel sizes"); return false; }
    taps = tp;
    n_rows += t->dims[0];
  }
  const int c_pad = (c_in_total + k_align - 1) / k_align * k_align;
  const size_t k_total = (size_t)taps * c_pad;
  std::vector<uint16_t> hw((size_t)n_rows * k_total, 0);
  std::vector<float> hb((size_t)n_rows, 0.f);
  int row0 = 0;
  for (auto& s : srcs) {
    const WTensor* t = wf.find(s.weight);
    const int o = t->dims[0], ci = t->dims[1];
    const uint16_t* src = (const uint16_t*)t->data;
    if (s.in_offset + ci > c_in_total) { set_error("pack_dense: %s exceeds c_in_total", s.weight.c_str()); return false; }
    for (int r = 0; r < o; ++r)
      for (int c = 0; c < ci; ++c)
        for (int tp = 0; tp < taps; ++tp) {
          int cdst = s.in_offset + c;
          if (in_perm) cdst = (*in_perm)[cdst];
          const int rdst = out_perm ? (*out_perm)[row0 + r] : row0 + r;
          hw[(size_t)rdst * k_total + (size_t)tp * c_pad + cdst] = src[((size_t)r * ci + c) * taps + tp];
        }
    if (!s.bias.empty()) {
      std::vector<float> b;
      if (!wf.get_f32(s.bias, &b)) return false;
      for (int r = 0; r < o; ++r) hb[out_perm ? (*out_perm)[row0 + r] : row0 + r] = b[r];
    }
    row0 += o;
  }
  out->n_rows = n_rows; out->c_in = c_in_total; out->c_in_pad = c_pad; out->taps = taps;
  out->w = arena->alloc_n<__half>(hw.size());
  out->bias = arena->alloc_n<float>(hb.size());
  if (!out->w || !out->bias) return false;
  AIRFE_CUDA_OK(cudaMemcpy(out->w, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  AIRFE_CUDA_OK(cudaMemcpy(out->bias, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
  return true;
}

// ---- profiler ---------------------------------------------------------------------------------------------------------------
Profiler& profiler() { static Profiler p; return p; }
void Profiler::record(const std::string& name, double fl, cudaStream_t st, const std::function<bool(cudaStream_t)>& fn, bool* ok, int kind) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a, st);
  *ok = fn(st);
  cudaEventRecord(b, st);
  OpProfile r;
  r.name = name; r.flops = fl; r.kind = kind;
  recs.push_back(r);
  evs.push_back({a, b});
}
void Profiler::finish() {
  for (size_t i = 0; i < evs.size(); ++i) {
    cudaEventSynchronize(evs[i].second);
    cudaEventElapsedTime(&recs[i].ms, evs[i].first, evs[i].second);
    cudaEventDestroy(evs[i].first);
    cudaEventDestroy(evs[i].second);
  }
  evs.clear();
  on = false;
}

// ---- op construction -----------------------------------------------------------------------------------------------------
bool add_gemm(OpList* ol, const TcGemmDesc& d, double flops, int kind) {
  TcGemmPlan plan;
  if (!tc_gemm_plan(d, &plan)) return false;
  ol->tc_flops += flops;
  ol->launches += 1;
  char nm[160];
  snprintf(nm, sizeof(nm), "tc_gemm attn M=%dx%dx%d K=%d N=%d%s", d.W, d.H, d.B, d.c_in_pad, d.n_valid, d.b_mn_major ? " (P.V)" : "");
  ol->push(nm, flops, [plan](cudaStream_t st) { return tc_gemm_launch(plan, st); }, kind);
  return true;
}

bool add_dense(OpList* ol, const Act& in, const DenseW& w, const Act& out, int batch, bool relu, int n_valid, int block_n,
               const int* dyn_rows, float scale, const float* resid, const Act* out2, const DenseExtra* ex) {
  TcGemmDesc d;
  d.a = in.p; d.a_C = in.C; d.W = in.W; d.H = in.H; d.B = batch;
  d.a_sx = in.ps; d.a_sy = in.ps * in.W; d.a_sb = in.ps * in.W * in.H;
  d.bw = w.w; d.k_total = w.taps * w.c_in_pad; d.n_rows = w.n_rows; d.bw_sn = d.k_total;
  d.taps = w.taps; d.c_in_pad = w.c_in_pad;
  if (n_valid < 0) n_valid = w.n_rows;
  d.n_valid = n_valid;
  if (!block_n) {
    const int n16 = (n_valid + 15) / 16 * 16;
    if (in.H == 1 && n16 >= 256 && n16 % 128 == 0) block_n = 128;   // row GEMMs of the matchers: more CTAs beat wider tiles
    else if (n16 <= 256) block_n = n16;
    else if (n16 % 256 == 0) block_n = 256;
    else if (n16 % 160 == 0) block_n = 160;
    else if (n16 % 128 == 0) block_n = 128;
    else block_n = 64;
  }
  d.block_n = block_n;
  d.bias = w.bias; d.relu = relu; d.out_f32 = out.f32;
  d.out = out.p; d.out_sx = out.ps; d.out_sy = out.ps * out.W; d.out_sb = out.ps * out.W * out.H;
  if (in.H == 1) { d.tw = 128; d.th = 1; d.tb = 1; }
  else if (in.W >= 16) { d.tw = 16; d.th = 8; d.tb = 1; }
  else { d.tw = 8; d.th = 8; d.tb = 2; }
  d.dyn_w = dyn_rows;
  if (scale != 1.f) { d.scale = scale; d.scale_cols = (ex && ex->scale_cols) ? ex->scale_cols : n_valid; }
  if (ex) { d.rot = ex->rot; d.rot_cols = ex->rot_cols; d.out_split = ex->out_split; d.out_split_stride = ex->out_split_stride; }
  d.resid = resid;
  if (out2) { d.out2 = out2->p; d.out2_sx = out2->ps; d.out2_sy = out2->ps * out2->W; d.out2_sb = out2->ps * out2->W * out2->H; }
  if (in.C > w.c_in_pad || in.C < w.c_in) { set_error("add_dense: activation has %d channels, weights expect %d", in.C, w.c_in); return false; }
  TcGemmPlan plan;
  if (!tc_gemm_plan(d, &plan)) return false;
  const double fl = 2.0 * (double)in.W * in.H * batch * (double)n_valid * w.taps * w.c_in;
  ol->tc_flops += fl;
  ol->launches += 1;
  char nm[160];
  snprintf(nm, sizeof(nm), "tc_gemm %s %d->%d @%dx%dx%d", w.taps == 9 ? "conv3x3" : (in.H == 1 ? "linear" : "conv1x1"), w.c_in, n_valid, in.W, in.H, batch);
  ol->push(nm, fl, [plan](cudaStream_t st) { return tc_gemm_launch(plan, st); }, dyn_rows ? ol->dyn_kind : kDynNone);
  return true;
}

}  // namespace airfe
