// Launchers of the matcher's non-GEMM kernels (see match_kernels.cu).
#pragma once
#include "common.h"

namespace airfe {

void launch_lg_prepare(const float* feat, const int* n, int slots, int cap, int feat_cap, int width, int height, float l_inv, const __half* wr,
                       float* x, __half* cat16, float* rot, cudaStream_t st);
void launch_lg_rotary(const float* qkv, const float* rot, const int* n, int slots, int cap, __half* q16, __half* k16, __half* v16, cudaStream_t st);
void launch_softmax_rows(const float* S, __half* P, const int* n, int slots, int cap, int col_xor, cudaStream_t st);
void launch_ln_gelu(const float* h, const float* gamma, const float* beta, const int* n, int slots, int cap, __half* out, cudaStream_t st);
void launch_lg_assignment(const float* sim, const float* x, const __half* wm, float bm, const int* n, int pairs, int cap, float* logsig,
                          float* lse, int* row_arg, float* row_val, int* col_arg, float thr, int* m_idx, float* m_score, int* m_count,
                          float* scores_out, cudaStream_t st);

}  // namespace airfe
