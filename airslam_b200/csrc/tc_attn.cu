// Host side of the fused attention kernel: tensor maps over the q / k / v operand matrices and launch.
#include "common.h"
#include "engine.h"
#include "tc_attn.cuh"

#include <stdlib.h>

namespace airfe {

bool attn_fused_enabled() {
  static int v = -1;
  if (v < 0) v = getenv("AIRFE_ATTN_V1") ? 0 : 1;
  return v == 1;
}

// q, k, v: fp16 matrices with `row_stride` elements per keypoint row, head h at columns h*64; slot s at rows [s*cap, (s+1)*cap).
bool add_fused_attention(OpList* ol, const __half* q, const __half* k, const __half* v, long long row_stride, __half* ctx, const int* n, int slots,
                         int cap, int slot_xor, float scale, const int* row_off) {
  if (cap > 512 || cap % 128) { set_error("fused attention supports cap <= 512"); return false; }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  // packed layout: one "slot" of slots * cap rows; the kernel adds the slot's row base to the row coordinate
  uint64_t dims[4] = {64, row_off ? (uint64_t)cap * slots : (uint64_t)cap, 4, row_off ? 1u : (uint64_t)slots};
  uint64_t str[3] = {(uint64_t)row_stride * 2, 64 * 2, (uint64_t)cap * row_stride * 2 * (row_off ? slots : 1)};
  uint32_t box_q[4] = {64, 128, 1, 1}, box_kv[4] = {64, 256, 1, 1};
  if (!make_tmap_f16(&p.tmQ, q, 4, dims, str, box_q) || !make_tmap_f16(&p.tmK, k, 4, dims, str, box_kv) || !make_tmap_f16(&p.tmV, v, 4, dims, str, box_kv))
    return false;
  p.n = n; p.slots = slots; p.cap = cap; p.slot_xor = slot_xor; p.scale = scale; p.ctx = ctx; p.row_off = row_off;
  // Small batches (the per-call class surface: 2 slots; config 3: 16) leave most SMs idle with one CTA per (slot, head): split the query tiles
  // of a (slot, head) over up to cap / 128 CTAs while that still fits one wave (each part re-reads K / V from L2, so large batches do not split).
  int q_split = 1;
  while (q_split * 2 <= cap / 128 && slots * 4 * q_split * 2 <= device_sm_count()) q_split *= 2;
  p.q_split = q_split;
  const double fl = 2.0 * 2.0 * slots * 4 * (double)cap * cap * 64;   // same accounting as the unfused pair of GEMMs
  ol->tc_flops += fl;
  ol->launches += 1;
  char nm[96];
  snprintf(nm, sizeof(nm), "tc_attn fused %s slots=%d cap=%d", slot_xor ? "cross" : "self", slots, cap);
  ol->push(nm, fl, [p, slots, q_split](cudaStream_t st) {
    static bool attr_set[kMaxDevices][2] = {};
    static const int wide = getenv("AIRFE_ATTN_NP") ? (atoi(getenv("AIRFE_ATTN_NP")) == 4) : 0;   // AIRFE_ATTN_NP=4: sixteen softmax warps.  Measured equal to eight (80 vs 82 us per launch, profiles/r02c_lightglue_ab.txt): the exp pass is bound by the XU pipe (MUFU.EX2 + fp16 conversions: 51 k exponentials per tile in ~6 k cycles), not by latency
    auto kern = wide ? tc_attn_kernel<4> : tc_attn_kernel<2>;
    const int dev = current_device();
    if (!attr_set[dev][wide]) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes) != cudaSuccess) {
        set_error("cudaFuncSetAttribute(tc_attn_kernel) failed: %s", cudaGetErrorString(cudaGetLastError()));
        return false;
      }
      attr_set[dev][wide] = true;
    }
    AttnParams pp = p;
    pp.trace = match_trace_buf() ? match_trace_buf() + 1024 : nullptr;    // authoring aid, normally nullptr
    cudaError_t e = launch_pdl(kern, slots * 4 * q_split, wide ? kAttnThreadsWide : kAttnThreads, kAttnSmemBytes, st, pp);
    if (e != cudaSuccess) { set_error("tc_attn launch failed: %s", cudaGetErrorString(e)); return false; }
    return true;
  }, slot_xor ? kDynAttCross : kDynAttSelf);
  return true;
}

}  // namespace airfe
