// Launchers of the matcher's non-GEMM kernels (see match_kernels.cu).
#pragma once
#include "common.h"

namespace airfe {

// row_off (optional, device int[slots + 1], written here): packed row layout, see lg_offsets_kernel
void launch_lg_prepare(const float* feat, const float* const* feat_ptrs /* optional per-slot bases */, const int* n, int* row_off, int slots, int cap, int feat_cap, int width, int height, float l_inv, const __half* wr,
                       float* x, __half* cat16, float* rot, cudaStream_t st);
void launch_lg_rotary(const float* qkv, const float* rot, const int* n, int slots, int cap, __half* q16, __half* k16, __half* v16, cudaStream_t st);
void launch_softmax_rows(const float* S, __half* P, const int* n, int slots, int cap, int col_xor, cudaStream_t st);
void launch_ln_gelu(const float* h, const float* gamma, const float* beta, const int* n, int slots, int cap, __half* out, cudaStream_t st);
// matchability logits of the final state x (packed when row_off != nullptr) -> logsig [slot][cap]; optionally unpacks md16 (packed) -> md16_pad
void launch_lg_matchability(const float* x, const __half* wm, float bm, const int* n, const int* row_off, int pairs, int cap, float* logsig, const __half* md16,
                            __half* md16_pad, cudaStream_t st);
void launch_lg_assignment(const float* sim, const int* n, int pairs, int cap, float* logsig,
                          float* lse, int* row_arg, float* row_val, int* col_arg, float thr, int* m_idx, float* m_score, int* m_count,
                          float* scores_out, cudaStream_t st);

}  // namespace airfe

namespace airfe {
// ---- SuperGlue (G5) ------------------------------------------------------------------------------------------------------------
void launch_sg_prepare(const float* feat, const float* const* feat_ptrs, const int* n, int slots, int cap, int feat_cap, int width, int height, float l_inv, float* x,
                       __half* kin16, cudaStream_t st);
// couplings Z [pair][cap+1][cap+1] from sim (already divided by 16) + bin score; 100 log-Sinkhorn iterations; decode.
// max_n: host-side upper bound of the keypoints per slot (-1 = unknown: the slot capacity is assumed).  It only selects the size class of the
// fused cluster kernel (<= 415: 13 columns per lane; <= 512: 17); capacities above 512 run one launch per pass.
void launch_sg_sinkhorn_decode(const float* sim, const int* n, int pairs, int cap, int max_n, float bin_score, int iters, float* Z, float* u, float* v,
                               float thr, int* arg0, float* val0, int* arg1, int* idx0, int* idx1, float* ms0, float* ms1, int* m_idx,
                               float* m_score, int* m_count, float* dense_out, cudaStream_t st);
}  // namespace airfe
