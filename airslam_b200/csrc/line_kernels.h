// Launchers of the PLNet line-path kernels (see line_kernels.cu).
#pragma once
#include "common.h"

namespace airfe {

constexpr int kJunctions = 300;             // TopK junctions in G2 (plnet_s0.onnx node /TopK)
constexpr int kProposals = 3 * 128 * 128;   // HAFM line proposals per image

void launch_hafm_decode(const float* heads, int ld, float* lines, float* jloc, int batch, cudaStream_t st);
void launch_junctions(const float* jloc, const float* heads, int ld, int* peaks, int* n_peaks, uint8_t* is_peak, float* juncs,
                      int* junc_idx, int batch, cudaStream_t st);
void launch_association(const float* lines, const float* juncs, int* imin, int* imax, uint8_t* keep, int* pair_table, int* uid_pairs,
                        int* uid_first, int* n_unique, int line_cap, int batch, cudaStream_t st);
void launch_loi_gather(const float* loi, int loi_ld, const float* thinaux, int ta_ld, const float* juncs, const float* lines,
                       const int* uid_pairs, const int* uid_first, const int* n_unique, int line_cap, const float* tspan, __half* feat,
                       float* adj, int batch, cudaStream_t st, __half* junc_feat = nullptr);
void launch_line_head(const float* h1, const float* h2, const float* w, const float* bias, const int* n_unique, int line_cap, float* score,
                      int batch, cudaStream_t st);
void launch_line_accept(const float* adj, const float* score, const int* n_unique, int line_cap, float line_thr, float len_thr, int border,
                        uint8_t* junc_map, float* lines_out, int* n_lines, int batch, cudaStream_t st);
void launch_junction_scan(const uint8_t* junc_map, const float* scores, int border, float* kp, int kp_cap, int* kp_count, int batch,
                          cudaStream_t st);

}  // namespace airfe
