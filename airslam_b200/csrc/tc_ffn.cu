// Host side of the fused LightGlue transformer-block kernel.
#include "common.h"
#include "engine.h"
#include "tc_ffn.cuh"

#include <stdlib.h>

namespace airfe {

bool ffn_fused_enabled() {
  static int v = -1;
  if (v < 0) v = getenv("AIRFE_FFN_V1") ? 0 : 1;
  return v == 1;
}

static int ffn_sm_count() { return device_sm_count(); }

static long long* g_match_trace = nullptr;
long long* match_trace_buf() { return g_match_trace; }
void match_set_trace(long long* dev_buf) { g_match_trace = dev_buf; }

bool add_fused_ffn(OpList* ol, const __half* ctx16, __half* cat16, float* x, const DenseW& w_out, const DenseW& w0, const DenseW& w3, const float* ln_g,
                   const float* ln_b, const int* n, int slots, int cap, bool relu) {
  if (w_out.n_rows != 256 || w_out.c_in_pad != 256 || w0.n_rows != 512 || w0.c_in_pad != 512 || w3.n_rows != 256 || w3.c_in_pad != 512) {
    set_error("add_fused_ffn: unexpected layer shapes");
    return false;
  }
  FfnParams p;
  memset(&p, 0, sizeof(p));
  {
    uint64_t dims[4] = {256, (uint64_t)cap, 1, (uint64_t)slots};
    uint32_t box[4] = {64, 128, 1, 1};
    uint64_t s_ctx[3] = {256 * 2, (uint64_t)cap * 256 * 2, (uint64_t)cap * 256 * 2};
    uint64_t s_x[3] = {512 * 2, (uint64_t)cap * 512 * 2, (uint64_t)cap * 512 * 2};
    if (!make_tmap_f16(&p.tmCtx, ctx16, 4, dims, s_ctx, box) || !make_tmap_f16(&p.tmX16, cat16, 4, dims, s_x, box)) return false;
  }
  auto wmap = [&](CUtensorMap* tm, const DenseW& w) {
    uint64_t dims[4] = {(uint64_t)w.c_in_pad, (uint64_t)w.n_rows, 1, 1};
    uint64_t str[3] = {(uint64_t)w.c_in_pad * 2, (uint64_t)w.c_in_pad * 2 * w.n_rows, (uint64_t)w.c_in_pad * 2 * w.n_rows};
    uint32_t box[4] = {64, 128, 1, 1};
    return make_tmap_f16(tm, w.w, 4, dims, str, box);
  };
  if (!wmap(&p.tmWo, w_out) || !wmap(&p.tmW0, w0) || !wmap(&p.tmW3, w3)) return false;
  p.b_out = w_out.bias; p.b0 = w0.bias; p.b3 = w3.bias;
  p.ln_g = relu ? w0.bias : ln_g; p.ln_b = relu ? w0.bias : ln_b;     // the ReLU variant stages but never reads the LayerNorm slots
  p.x = x; p.x16 = cat16; p.n = n; p.slots = slots; p.cap = cap;
  static const int prewait = getenv("AIRFE_PREWAIT") ? 1 : 0;   // measured in round 2: no gain (profiles/r02_prewait_ab.txt), off by default
  p.prewait = prewait;
  const int tiles = slots * (cap / 128);
  const int grid = tiles < ffn_sm_count() ? tiles : ffn_sm_count();
  const double fl = 2.0 * (double)slots * cap * (256.0 * 256 + 512.0 * 512 + 512.0 * 256);
  ol->tc_flops += fl;
  ol->launches += 1;
  // sixteen epilogue warps (EW = 4, tc_ffn.cuh): measured 2.14 -> 2.03 ms per 18 LightGlue launches and 0.745 -> 0.710 ms per 18 SuperGlue launches
  // (profiles/r02e_wide_epilogue_ab.txt) -> default; AIRFE_FFN_WIDE=0 selects the eight-warp kernel (read at plan time; the choice depends on
  // nothing else, so a pair inside a batch stays bit-identical to the same pair alone)
  const int wide = getenv("AIRFE_FFN_WIDE") ? (atoi(getenv("AIRFE_FFN_WIDE")) ? 1 : 0) : 1;
  ol->push(relu ? "tc_ffn fused block (merge+mlp0+ReLU+mlp3+residual)" : "tc_ffn fused block (out_proj+ffn0+LN+GELU+ffn3+residual)", fl, [p, grid, relu, wide](cudaStream_t st) {
    static bool attr_set[kMaxDevices][4] = {};
    const int dev = current_device();
    using FfnKernel = void (*)(const FfnParams);
    static const FfnKernel kerns[4] = {tc_ffn_kernel<false, 2>, tc_ffn_kernel<true, 2>, tc_ffn_kernel<false, 4>, tc_ffn_kernel<true, 4>};
    const int key = (wide ? 2 : 0) | (relu ? 1 : 0);
    FfnKernel kern = kerns[key];
    if (!attr_set[dev][key]) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnSmemBytes) != cudaSuccess) {
        set_error("cudaFuncSetAttribute(tc_ffn_kernel) failed: %s", cudaGetErrorString(cudaGetLastError()));
        return false;
      }
      attr_set[dev][key] = true;
    }
    FfnParams pp = p;
    pp.trace = g_match_trace;                      // authoring aid, normally nullptr
    cudaError_t e = launch_pdl(kern, grid, wide ? kFfnWideThreads : kFfnThreads, kFfnSmemBytes, st, pp);
    if (e != cudaSuccess) { set_error("tc_ffn launch failed: %s", cudaGetErrorString(e)); return false; }
    return true;
  }, kDynRows);
  return true;
}

}  // namespace airfe
