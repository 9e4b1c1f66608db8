// Host runtime of libairfe: weight containers, device arena, op lists (the replacement for the reference's
// TensorRT engines + tensorrt_buffer::BufferManager, 3rdparty/tensorrtbuffer/include/buffers.h:237-417, which
// cudaMalloc/cudaFree every binding on every infer() call; here everything is allocated once per context).
#pragma once
#include "common.h"
#include "kernels.h"

#include <functional>
#include <map>
#include <string>
#include <vector>

namespace airfe {

// ---- AIRFEW01 weight container (tools/make_weights.py) -----------------------------------------------------------
struct WTensor {
  int dtype = 0;  // 0 f32, 1 f16, 2 i32
  int ndim = 0;
  int dims[4] = {1, 1, 1, 1};
  const uint8_t* data = nullptr;
  size_t nbytes = 0;
  size_t numel() const { return (size_t)dims[0] * dims[1] * dims[2] * dims[3]; }
};
class WeightFile {
 public:
  bool load(const std::string& path);
  const WTensor* find(const std::string& name) const;
  bool has(const std::string& name) const { return find(name) != nullptr; }
  // fetch as float (converting fp16), empty vector on error
  bool get_f32(const std::string& name, std::vector<float>* out) const;
 private:
  std::vector<uint8_t> blob_;
  std::map<std::string, WTensor> index_;
};

// ---- bump allocator over one cudaMalloc --------------------------------------------------------------------------
class Arena {
 public:
  ~Arena();
  bool init(size_t bytes);
  void* alloc(size_t bytes, size_t align = 1024);
  template <typename T> T* alloc_n(size_t n) { return reinterpret_cast<T*>(alloc(n * sizeof(T))); }
  size_t used() const { return off_; }
  bool ok() const { return !failed_; }
 private:
  uint8_t* base_ = nullptr;
  size_t cap_ = 0, off_ = 0;
  bool failed_ = false;
};

// ---- packed weights of one dense layer (device) -------------------------------------------------------------------
struct DenseW {
  __half* w = nullptr;    // [n_rows][taps * c_in_pad]
  float* bias = nullptr;  // [n_rows]
  int n_rows = 0, c_in = 0, c_in_pad = 0, taps = 1;
};

// NHWC fp16/fp32 activation view on a [Bmax, H, W, ps] buffer (channel offset folded into p)
struct Act {
  void* p = nullptr;
  int C = 0, H = 0, W = 0;
  long long ps = 0;  // pixel stride in elements
  bool f32 = false;
  Act slice(int c0, int c) const {
    Act a = *this;
    a.p = (uint8_t*)p + (size_t)c0 * (f32 ? 4 : 2);
    a.C = c;
    return a;
  }
};

using Op = std::function<bool(cudaStream_t)>;
// How an op's algorithmic work depends on the device-side counts (keypoints per slot, verified lines per image): the op lists are built for
// the CAPACITY of a slot (cap rows), but rows beyond the counts are skipped on the device, so roofline accounting must scale the FLOPs to the
// rows actually processed (airfe_profile_stereo reads the counts back once and applies the factor).
enum DynKind { kDynNone = 0, kDynRows = 1 /* ~ sum n_s */, kDynAttSelf = 2 /* ~ sum n_s^2 */, kDynAttCross = 3 /* ~ sum n_s n_(s^1) */,
               kDynSim = 4 /* ~ sum n_2p n_2p+1 */, kDynLines = 5 /* ~ sum unique lines */ };
struct OpProfile {            // filled when profiling is on (airfe_profile_begin): one record per executed op
  std::string name;
  double flops = 0;           // algorithmic tensor-core FLOPs at slot capacity (0 for non-GEMM ops)
  int kind = kDynNone;
  float ms = 0;
};
struct Profiler {
  bool on = false;
  std::vector<OpProfile> recs;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evs;
  void begin() { on = true; recs.clear(); }
  void record(const std::string& name, double flops, cudaStream_t st, const std::function<bool(cudaStream_t)>& fn, bool* ok, int kind = kDynNone);
  void finish();              // synchronises and converts events to milliseconds
};
Profiler& profiler();

struct OpList {
  std::vector<Op> ops;
  std::vector<std::string> names;   // parallel to ops (tensor-core ops carry their shape in the name)
  std::vector<double> flops;        // parallel to ops
  std::vector<int> kinds;           // parallel to ops (DynKind)
  double tc_flops = 0;   // algorithmic FLOPs of the tensor-core ops in this list (at slot capacity)
  int launches = 0;
  int dyn_kind = kDynRows;          // what a dyn_rows GEMM of this list scales with (the detector's stage-1 MLP list sets kDynLines)
  void push(const std::string& name, double fl, Op op, int kind = kDynNone) { ops.push_back(std::move(op)); names.push_back(name); flops.push_back(fl); kinds.push_back(kind); }
  bool run(cudaStream_t st) const {
    Profiler& pr = profiler();
    for (size_t i = 0; i < ops.size(); ++i) {
      if (pr.on) {
        bool ok = true;
        pr.record(i < names.size() ? names[i] : "op", i < flops.size() ? flops[i] : 0.0, st, ops[i], &ok, i < kinds.size() ? kinds[i] : 0);
        if (!ok) return false;
      } else if (!ops[i](st)) {
        return false;
      }
    }
    return true;
  }
};

// Run `f()` (a group of plain kernel launches on `st`), recording it as one profiled op when profiling is on.
template <class F>
inline void timed(const char* name, cudaStream_t st, F&& f) {
  Profiler& pr = profiler();
  if (pr.on) {
    bool ok = true;
    pr.record(name, 0.0, st, [&](cudaStream_t) { f(); return true; }, &ok);
  } else {
    f();
  }
}

// Pack helpers (host): conv OIHW / linear [out,in] fp16 -> tap-major K-major rows, several sources concatenated along N.
struct PackSrc {
  std::string weight, bias;  // tensor names in the container (bias may be empty)
  int in_offset = 0;         // place this source's input channels at [in_offset, in_offset + c_in) of the packed K (block-diagonal / concat remap)
};
bool pack_dense(const WeightFile& wf, const std::vector<PackSrc>& srcs, int c_in_total, Arena* arena, DenseW* out,
                const std::vector<int>* in_perm = nullptr, const std::vector<int>* out_perm = nullptr /* old row -> new row */,
                int k_align = 64 /* K padding per tap: 64, or 32 for the SWIZZLE_64B conv path */);

// Optional epilogue extras of add_dense (the LightGlue projections): scale only the first scale_cols columns, rotary embedding on the
// first rot_cols columns, column sections of out_split stored out_split_stride elements apart.
struct DenseExtra {
  int scale_cols = 0;
  const float* rot = nullptr; int rot_cols = 0;
  int out_split = 0; long long out_split_stride = 0;
};
// Append a tcgen05 conv / GEMM op.
bool add_dense(OpList* ol, const Act& in, const DenseW& w, const Act& out, int batch, bool relu, int n_valid = -1, int block_n = 0,
               const int* dyn_rows = nullptr, float scale = 1.f, const float* resid = nullptr, const Act* out2 = nullptr, const DenseExtra* ex = nullptr);
// Append a 3x3 convolution on the halo-reuse tcgen05 kernel (tc_conv3x3.cuh); `out` and/or `pool_out` (fused 2x2 max-pool).
// Falls back to the generic streaming-tap kernel (+ pool kernel) when AIRFE_CONV_V1 is set or the map is narrower than 8.
// up2 = true: `out` is the nearest-neighbour 2x up-sampled map (out->W == 2 * in.W): the epilogue stores every pixel four times, which replaces
// the separate upsample2 kernel between a hourglass decoder conv and the `deconv` conv that follows the Resize node (halo kernel only).
bool add_conv3x3(OpList* ol, const Act& in, const DenseW& w, const Act* out, const Act* pool_out, int batch, bool relu, bool up2 = false);
bool conv3x3_halo_enabled();
void match_set_trace(long long* dev_buf);      // authoring aid, see airfe_debug_match_trace
long long* match_trace_buf();
void conv3x3_set_trace(long long* dev_buf);   // authoring aid, see airfe_debug_conv_trace
// Append fused multi-head attention (tc_attn.cuh): ctx = softmax(q k^T * scale) v per (slot, head); keys/values of slot ^ slot_xor.
// row_off (optional, device int[slots + 1]): packed row layout, slot s at rows [row_off[s], row_off[s] + n[s]).
bool add_fused_attention(OpList* ol, const __half* q, const __half* k, const __half* v, long long row_stride, __half* ctx, const int* n, int slots,
                         int cap, int slot_xor, float scale, const int* row_off = nullptr);
bool attn_fused_enabled();
// Append the fused LightGlue block tail (tc_ffn.cuh): x += ffn3(gelu(LN(ffn0([x | out_proj(ctx)])))); refreshes the fp16 copy of x.
// relu = true: SuperGlue's block tail (merge + MLP with ReLU instead of LayerNorm + GELU; ln_g / ln_b unused).
bool add_fused_ffn(OpList* ol, const __half* ctx16, __half* cat16, float* x, const DenseW& w_out, const DenseW& w0, const DenseW& w3, const float* ln_g,
                   const float* ln_b, const int* n, int slots, int cap, bool relu = false);
bool ffn_fused_enabled();
// Append a raw tcgen05 GEMM described by `d` (attention products).
bool add_gemm(OpList* ol, const TcGemmDesc& d, double flops, int kind = kDynNone);

}  // namespace airfe
