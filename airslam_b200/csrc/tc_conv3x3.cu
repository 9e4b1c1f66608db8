// Host side of the halo-reuse 3x3 convolution kernel: tiling plan, shared-memory budget, launch.
#include "common.h"
#include "engine.h"
#include "tc_conv3x3.cuh"
#include "tc_conv3x3_fold.cuh"

#include <stdlib.h>

namespace airfe {

struct ConvPlan {
  ConvParams p;
  int grid = 0, smem_bytes = 0;
  int wide = 0;       // nine-tap halo kernel with 16 epilogue warps (EW = 4)
};

static int sm_count() { return device_sm_count(); }

static long long* g_conv_trace = nullptr;
void conv3x3_set_trace(long long* dev_buf) { g_conv_trace = dev_buf; }

using ConvKernel = void (*)(const ConvParams);

static int conv_key(const ConvParams& p, int wide) {
  if (p.fold == 2 && p.kblocks == 2) return 16;
  if (p.fold) return 8 + (p.fold == 2 ? 4 : 0) + (p.kw == 32 ? 2 : 0) + (p.block_n == 64 ? 1 : 0);
  return (wide ? 17 : 0) + ((p.kw == 32 ? 4 : 0) | (p.strips == 2 ? 2 : 0) | (p.b_resident ? 1 : 0));
}

static ConvKernel conv_kernel_for(const ConvParams& p, int wide) {
  if (p.fold == 2 && p.kblocks == 2) return tc_conv3x3_fold_kernel<64, 32, true, 1, 2>;
  if (p.fold == 2) {
    if (p.kw == 64) return p.block_n == 64 ? tc_conv3x3_fold_kernel<64, 64, true> : tc_conv3x3_fold_kernel<64, 32, true>;
    return p.block_n == 64 ? tc_conv3x3_fold_kernel<32, 64, true> : tc_conv3x3_fold_kernel<32, 32, true>;
  }
  if (p.fold) {
    if (p.kw == 64) return p.block_n == 64 ? tc_conv3x3_fold_kernel<64, 64, false> : tc_conv3x3_fold_kernel<64, 32, false>;
    return p.block_n == 64 ? tc_conv3x3_fold_kernel<32, 64, false> : tc_conv3x3_fold_kernel<32, 32, false>;
  }
  const int key = (p.kw == 32 ? 4 : 0) | (p.strips == 2 ? 2 : 0) | (p.b_resident ? 1 : 0);
  if (wide) {
    switch (key) {
      case 0: return tc_conv3x3_kernel<64, 1, false, 4>;
      case 1: return tc_conv3x3_kernel<64, 1, true, 4>;
      case 2: return tc_conv3x3_kernel<64, 2, false, 4>;
      case 3: return tc_conv3x3_kernel<64, 2, true, 4>;
      case 4: return tc_conv3x3_kernel<32, 1, false, 4>;
      case 5: return tc_conv3x3_kernel<32, 1, true, 4>;
      case 6: return tc_conv3x3_kernel<32, 2, false, 4>;
      default: return tc_conv3x3_kernel<32, 2, true, 4>;
    }
  }
  switch (key) {
    case 0: return tc_conv3x3_kernel<64, 1, false>;
    case 1: return tc_conv3x3_kernel<64, 1, true>;
    case 2: return tc_conv3x3_kernel<64, 2, false>;
    case 3: return tc_conv3x3_kernel<64, 2, true>;
    case 4: return tc_conv3x3_kernel<32, 1, false>;
    case 5: return tc_conv3x3_kernel<32, 1, true>;
    case 6: return tc_conv3x3_kernel<32, 2, false>;
    default: return tc_conv3x3_kernel<32, 2, true>;
  }
}

static bool conv_launch(const ConvPlan& plan, cudaStream_t st) {
  static bool attr_set[kMaxDevices][25] = {};
  const int dev = current_device();
  const int key = conv_key(plan.p, plan.wide);
  const int threads = plan.p.fold == 2 ? kFoldWideThreads : (plan.wide ? kConvWideThreads : kConvThreads);
  ConvKernel kern = conv_kernel_for(plan.p, plan.wide);
  if (!attr_set[dev][key]) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(tc_conv3x3_kernel) failed: %s", cudaGetErrorString(cudaGetLastError()));
      return false;
    }
    attr_set[dev][key] = true;
  }
  cudaError_t e;
  if (g_conv_trace) {
    ConvPlan traced = plan;
    traced.p.trace = g_conv_trace;
    e = launch_pdl(kern, traced.grid, threads, traced.smem_bytes, st, traced.p);
  } else {
    e = launch_pdl(kern, plan.grid, threads, plan.smem_bytes, st, plan.p);
  }
  if (e != cudaSuccess) { set_error("tc_conv3x3 launch failed: %s", cudaGetErrorString(e)); return false; }
  return true;
}

// kx-folded variant (tc_conv3x3_fold.cuh) for resident-weights small-N layers.  Measured on B200 (profiles/r02c_fold_ab.txt, 16 frames of
// 512 x 512, nine-tap -> folded with 16 epilogue warps): 64->32 0.233 -> 0.182 ms, 32->32 pool-only 0.150 -> 0.137, 64->64 pool-only 0.284 -> 0.284,
// 64->64 0.295 -> 0.326, 64->64 + pool 0.348 -> 0.351: the fold cuts the MMA time by 1.3-1.9x and triples the accumulator columns the epilogue has
// to read, shift and add (the folded kernel is bound by the epilogue's instruction issue), so it pays for C_out = 32 and not for C_out = 64.
// Default: C_out = 32 layers.  AIRFE_CONV_FOLD=0: never, =2: every eligible layer (tests, A/B).
static int conv3x3_fold_mode() {
  const char* e = getenv("AIRFE_CONV_FOLD");   // read at plan time (not cached: the operator tests switch it between calls)
  return e ? atoi(e) : 1;
}

bool conv3x3_halo_enabled() {
  static int v = -1;
  if (v < 0) v = getenv("AIRFE_CONV_V1") ? 0 : 1;
  return v == 1;
}

bool add_conv3x3(OpList* ol, const Act& in, const DenseW& w, const Act* out, const Act* pool_out, int batch, bool relu, bool up2) {
  if (w.taps != 9 || in.f32 || (out && out->f32) || (pool_out && pool_out->f32)) { set_error("add_conv3x3: unsupported layer"); return false; }
  if (up2) {
    // fused nearest 2x up-sampling store: halo kernel only, whole 16-channel chunks, 32-byte aligned pixel rows of the up-sampled buffer
    const bool ok = conv3x3_halo_enabled() && in.W >= 8 && in.W % 8 == 0 && out && !pool_out && out->W == 2 * in.W && out->H == 2 * in.H && w.n_rows % 16 == 0 &&
                    w.n_rows > 64 && ((uintptr_t)out->p & 31) == 0 && out->ps % 16 == 0;
    if (!ok) { set_error("add_conv3x3: layer cannot store an up-sampled map"); return false; }
  }
  if (!conv3x3_halo_enabled() || in.W < 8 || (in.W % 8)) {
    // generic streaming-tap kernel (+ separate pool kernel)
    const Act* full = out;
    if (!full) { set_error("add_conv3x3: v1 path needs a full-resolution output buffer"); return false; }
    if (!add_dense(ol, in, w, *full, batch, relu)) return false;
    if (pool_out) {
      const Act i = *full, o = *pool_out;
      ol->push("maxpool2", 0, [=](cudaStream_t st) { launch_maxpool2((const __half*)i.p, i.C, i.H, i.W, batch, i.ps, (__half*)o.p, o.ps, st); return true; });
      ol->launches++;
    }
    return true;
  }
  ConvPlan plan;
  ConvParams& p = plan.p;
  memset(&p, 0, sizeof(p));
  const int n_valid = w.n_rows;
  const int n16 = (n_valid + 15) / 16 * 16;
  int block_n = n16 <= 256 ? n16 : (n16 % 256 == 0 ? 256 : (n16 % 160 == 0 ? 160 : 128));
  {
    // small maps (hourglass bottom: 8x8 .. 32x32 x batch) give fewer M tiles than SMs: split N further so more CTAs share the
    // layer (each then streams only its slice of the weights).  AIRFE_CONV_SPLIT_N=0 switches this off for A/B timing.
    static const int split_n = getenv("AIRFE_CONV_SPLIT_N") ? atoi(getenv("AIRFE_CONV_SPLIT_N")) : 1;
    const int strips0 = (in.W >= 16 && 4 * conv_acc_stride(block_n) <= 512) ? 2 : 1;
    const int m_tiles = ((in.W + 8 * strips0 - 1) / (8 * strips0)) * ((in.H + kConvTH - 1) / kConvTH) * batch;
    while (split_n && block_n >= 64 && block_n % 32 == 0 && m_tiles * ((n_valid + block_n - 1) / block_n) * 2 <= sm_count()) block_n /= 2;
  }
  p.block_n = block_n;
  p.n_valid = n_valid;
  p.n_tiles = (n_valid + block_n - 1) / block_n;
  p.strips = (in.W >= 16 && 4 * conv_acc_stride(block_n) <= 512) ? 2 : 1;
  p.nacc = conv_nacc(block_n, p.strips);
  static const int prewait = getenv("AIRFE_PREWAIT") ? 1 : 0;   // measured in round 2: no gain (profiles/r02_prewait_ab.txt), off by default
  p.prewait = prewait;
  p.kw = (w.c_in_pad % 64) ? 32 : 64;
  p.kblocks = w.c_in_pad / p.kw;
  p.c_in_pad = w.c_in_pad;
  p.W = in.W; p.H = in.H; p.B = batch;
  p.tiles_x = (in.W + 8 * p.strips - 1) / (8 * p.strips);
  p.tiles_y = (in.H + kConvTH - 1) / kConvTH;
  p.bias = w.bias; p.relu = relu;
  if (out) { p.out = (__half*)out->p; p.out_sx = out->ps; p.out_sy = out->ps * out->W; p.out_sb = out->ps * out->W * out->H; p.up2 = up2 ? 1 : 0; }
  if (pool_out) { p.pool_out = (__half*)pool_out->p; p.pool_sx = pool_out->ps; p.pool_sy = pool_out->ps * pool_out->W; p.pool_sb = pool_out->ps * pool_out->W * pool_out->H; }
  const int a_bytes = conv_a_bytes(p.strips, p.kw), b_bytes = conv_b_bytes(block_n, p.kw);
  const int budget = 212 * 1024;
  const int res_bytes = 9 * p.kblocks * b_bytes;
  if (p.n_tiles == 1 && res_bytes + 2 * a_bytes <= budget && res_bytes <= 120 * 1024) {
    p.b_resident = 1;
    p.stages_a = (budget - res_bytes) / a_bytes;
    if (p.stages_a > 4) p.stages_a = 4;
    static const int cap_a = getenv("AIRFE_CONV_STAGES_A") ? atoi(getenv("AIRFE_CONV_STAGES_A")) : 0;   // experiments: cap the halo-tile ring depth
    if (cap_a >= 2 && p.stages_a > cap_a) p.stages_a = cap_a;
    p.stages_b = 0;
  } else {
    p.b_resident = 0;
    p.stages_a = 2;
    p.stages_b = (budget - 2 * a_bytes) / b_bytes;
    if (p.stages_b > 12) p.stages_b = 12;
    if (p.stages_b < 3) { set_error("add_conv3x3: shared memory budget too small"); return false; }
  }
  const int fold_mode = conv3x3_fold_mode();
  // The folded kernel has no partial-chunk store path: the layer's channel count must be a whole number of N tiles and every store 32-byte aligned.
  const bool fold_aligned = (!out || (((uintptr_t)out->p & 31) == 0 && out->ps % 16 == 0)) && (!pool_out || (((uintptr_t)pool_out->p & 31) == 0 && pool_out->ps % 16 == 0));
  int fold_n = 0, fold_s = kFoldStrips, fold_kb = 1;
  if (fold_mode && fold_aligned && n_valid == n16) {
    if (p.kblocks == 1 && (n16 == 32 || (n16 == 64 && fold_mode == 2))) fold_n = n16;                       // C_in <= 64: one resident N tile
    else if (p.kblocks == 2 && p.kw == 64 && n16 == 64 && (fold_mode == 2 || (in.W >= 32 && in.H >= 32))) {   // the choice must not depend on the batch: a pair inside a batch is bit-identical to the same pair alone     // C_in = 128 -> 64: two resident N tiles of 32, one-strip halo tiles
      static const int kb2 = getenv("AIRFE_CONV_FOLD_KB2") ? atoi(getenv("AIRFE_CONV_FOLD_KB2")) : 1;      // 0: keep the nine-tap kernel for these layers (A/B timing)
      if (kb2) { fold_n = 32; fold_s = 1; fold_kb = 2; }
    }
  }
  if (fold_n) {
    // kx taps folded into N (tc_conv3x3_fold.cuh): tiles of 14 output columns x 8S rows, one TMEM buffer of 3N columns per 8-row strip
    static const int wide = getenv("AIRFE_FOLD_WIDE") ? atoi(getenv("AIRFE_FOLD_WIDE")) : 1;   // 16 epilogue warps (0: the 8-warp software-pipelined epilogue, for A/B timing)
    p.fold = (wide || fold_kb == 2) ? 2 : 1;
    p.block_n = block_n = fold_n;
    p.n_tiles = n_valid / fold_n;
    p.b_resident = 1;
    p.strips = fold_s;
    p.tiles_x = (in.W + kFoldTX - 1) / kFoldTX;
    p.tiles_y = (in.H + 8 * fold_s - 1) / (8 * fold_s);
    p.nacc = fold_nbuf(block_n);
    const int fa = fold_kb * fold_a_bytes(p.kw, fold_s), fb = 9 * fold_kb * block_n * p.kw * 2;
    p.stages_a = (budget - fb) / fa;
    if (p.stages_a > 4) p.stages_a = 4;
    uint64_t dims[4] = {(uint64_t)in.C, (uint64_t)in.W, (uint64_t)in.H, (uint64_t)batch};
    uint64_t str[3] = {(uint64_t)in.ps * 2, (uint64_t)in.ps * in.W * 2, (uint64_t)in.ps * in.W * in.H * 2};
    uint32_t box[4] = {(uint32_t)p.kw, 16, (uint32_t)(8 * fold_s + 2), 1};
    if (!make_tmap_f16(&p.tmA, in.p, 4, dims, str, box, p.kw * 2)) return false;
    const uint64_t k_total = (uint64_t)9 * w.c_in_pad;
    uint64_t bd[4] = {k_total, (uint64_t)w.n_rows, 1, 1};
    uint64_t bs[3] = {k_total * 2, k_total * 2 * w.n_rows, k_total * 2 * w.n_rows};
    uint32_t bb[4] = {(uint32_t)p.kw, (uint32_t)block_n, 1, 1};
    if (!make_tmap_f16(&p.tmB, w.w, 4, bd, bs, bb, p.kw * 2)) return false;
    plan.smem_bytes = p.stages_a * fa + fb + 1024 + (2 * p.stages_a + 1 + 8) * 8 + 16;
    const int total = p.tiles_x * p.tiles_y * batch * p.n_tiles;
    plan.grid = total < sm_count() ? total : sm_count();
    plan.grid -= plan.grid % p.n_tiles;                  // a CTA keeps one N tile resident: t % n_tiles must not change as t += grid
  } else {
    uint64_t dims[4] = {(uint64_t)in.C, (uint64_t)in.W, (uint64_t)in.H, (uint64_t)batch};
    uint64_t str[3] = {(uint64_t)in.ps * 2, (uint64_t)in.ps * in.W * 2, (uint64_t)in.ps * in.W * in.H * 2};
    uint32_t box[4] = {(uint32_t)p.kw, (uint32_t)(8 * p.strips + 2), (uint32_t)(kConvTH + 2), 1};
    if (!make_tmap_f16(&p.tmA, in.p, 4, dims, str, box, p.kw * 2)) return false;
    const uint64_t k_total = (uint64_t)9 * w.c_in_pad;
    uint64_t bd[4] = {k_total, (uint64_t)w.n_rows, 1, 1};
    uint64_t bs[3] = {k_total * 2, k_total * 2 * w.n_rows, k_total * 2 * w.n_rows};
    uint32_t bb[4] = {(uint32_t)p.kw, (uint32_t)block_n, 1, 1};
    if (!make_tmap_f16(&p.tmB, w.w, 4, bd, bs, bb, p.kw * 2)) return false;
  }
  if (!p.fold) {
    // 16 epilogue warps (EW = 4).  Measured per layer on the B200 (profiles/r02e_wide_epilogue_ab.txt): -4 % on conv1b (64->64 + pool on the
    // 512 x 512 map: the epilogue moves two full-resolution maps), +0 .. 3 % (slower) on every other nine-tap layer -> default: full-resolution maps
    // only.  AIRFE_CONV_WIDE_MAXN=n overrides: every layer with block_n <= n (0: never).  The choice depends on the layer only, never on the batch.
    if (getenv("AIRFE_CONV_WIDE_MAXN")) plan.wide = block_n <= atoi(getenv("AIRFE_CONV_WIDE_MAXN")) ? 1 : 0;
    else plan.wide = (in.W >= 512 && in.H >= 512) ? 1 : 0;
    const int n_b_slots = p.b_resident ? 9 * p.kblocks : p.stages_b;
    plan.smem_bytes = p.stages_a * a_bytes + n_b_slots * b_bytes + 1024 + (2 * p.stages_a + 2 * (p.b_resident ? 1 : p.stages_b) + 8) * 8 + 16;
    const int total = p.tiles_x * p.tiles_y * batch * p.n_tiles;
    plan.grid = total < sm_count() ? total : sm_count();
  }
  if (in.C > w.c_in_pad || in.C < w.c_in) { set_error("add_conv3x3: activation has %d channels, weights expect %d", in.C, w.c_in); return false; }
  const double fl = 2.0 * (double)in.W * in.H * batch * (double)n_valid * 9 * w.c_in;
  ol->tc_flops += fl;
  ol->launches += 1;
  char nm[160];
  snprintf(nm, sizeof(nm), "tc_conv3x3 %d->%d @%dx%dx%d%s%s S%d%s", w.c_in, n_valid, in.W, in.H, batch, p.fold ? (p.kblocks == 2 ? " Bres kx-fold 2xN32" : " Bres kx-fold") : (p.b_resident ? " Bres" : ""), pool_out ? (out ? " +pool" : " pool-only") : (up2 ? " +up2" : ""), p.strips, p.kw == 32 ? (plan.wide ? " K32 W16" : " K32") : (plan.wide ? " W16" : ""));
  ol->push(nm, fl, [plan](cudaStream_t st) { return conv_launch(plan, st); });
  return true;
}

}  // namespace airfe
