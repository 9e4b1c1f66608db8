// Launcher of the point <-> line association / stereo line matching kernels (assoc_kernels.cu).
#pragma once
#include "common.h"

namespace airfe {

// lines512: device [2*pairs][line_stride][4] fp32 (x1,y1,x2,y2 in the 512 x 512 network frame), n_lines [2*pairs]; feat: device [2*pairs][feat_stride][259];
// m_idx [pairs][m_cap][2], m_count [pairs] (the matcher's outputs).  Image slot 2p = left, 2p+1 = right.  At most max_lines lines per image take
// part (more are ignored and flagged), rel_cap points per line, 8 lines per point; any overflow sets *overflow = 1.
void launch_line_assoc(const float* lines512, const int* n_lines, int line_stride, const float* feat, const int* n_feat, int feat_stride, double ws, double hs,
                       const int* m_idx, const int* m_count, int m_cap, int pairs, double min_x_diff, double max_x_diff, double max_y_diff, int max_lines,
                       int rel_cap, int* rel_n, int* rel_idx, float* rel_dist, int* cnt, int* row_loc, int* line_matches, int* overflow, cudaStream_t st);

}  // namespace airfe
