// the step right after the path, on data that is already on the device.
//   AssignPointsToLines (src/line_processor.cc:68-120): O(lines x points) point-to-segment tests in double precision;
//   MatchLines (:122-187): stereo line matches voted by the stereo POINT matches that sit on them, after the disparity filter of
//   Frame::AddRightFeatures (src/frame.cc:141-155).
// Inputs are the detector's / matcher's device-resident outputs of the last stereo call: lines in 512-space fp32 (scaled to image pixels in
// double exactly like the C ABI does on the host), features [n][259] fp32, match index pairs.  Arithmetic follows the reference expression by
// expression with explicit round-to-nearest double operations (no FMA contraction), so integer outputs and float distances are bit-exact
// against the numpy restatement in oracle/host.py.
#include "assoc_kernels.h"

namespace airfe {

// ---- AssignPointsToLines: one CTA per (line, image); points in ascending order (std::map iteration order) --------------------------------
__global__ void __launch_bounds__(128) assign_points_kernel(const float* __restrict__ lines512, const int* __restrict__ n_lines, int line_stride,
                                                            const float* __restrict__ feat, const int* __restrict__ n_feat, int feat_stride,
                                                            double ws, double hs, int max_lines, int rel_cap, int* __restrict__ rel_n,
                                                            int* __restrict__ rel_idx, float* __restrict__ rel_dist, int* __restrict__ overflow) {
  const int b = blockIdx.y, i = blockIdx.x;
  const int nl = min(n_lines[b], max_lines);
  if (i >= nl) return;
  const float* l = lines512 + ((long long)b * line_stride + i) * 4;
  // Vector4d(x1, y1, x2, y2) in 512-space floats, then *= w_scale / h_scale in double (src/plnet.cpp:551, 577-582)
  const double lx1 = __dmul_rn((double)l[0], ws), ly1 = __dmul_rn((double)l[1], hs), lx2 = __dmul_rn((double)l[2], ws), ly2 = __dmul_rn((double)l[3], hs);
  const double A = __dsub_rn(ly2, ly1), B = __dsub_rn(lx1, lx2);
  const double C = __dsub_rn(__dmul_rn(lx2, ly1), __dmul_rn(lx1, ly2));
  const double D = __dsqrt_rn(__dadd_rn(__dmul_rn(A, A), __dmul_rn(B, B)));
  const double min_lx = lx1 > lx2 ? lx2 : lx1, max_lx = lx1 > lx2 ? lx1 : lx2;
  const double min_ly = ly1 > ly2 ? ly2 : ly1, max_ly = ly1 > ly2 ? ly1 : ly2;
  const double line_side = __dmul_rn(D, D);
  const int np = n_feat[b];
  const float* f = feat + (long long)b * feat_stride * 259;
  __shared__ int s_warp[4];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const long long o = ((long long)b * max_lines + i) * rel_cap;
  for (int j0 = 0; j0 < np; j0 += 128) {
    const int j = j0 + threadIdx.x;
    bool ok = false;
    float dist = 0.f;
    if (j < np) {
      const double px = (double)f[(long long)j * 259 + 1], py = (double)f[(long long)j * 259 + 2];
      if (!(px < __dsub_rn(min_lx, 3.0) || px > __dadd_rn(max_lx, 3.0) || py < __dsub_rn(min_ly, 3.0) || py > __dadd_rn(max_ly, 3.0))) {
        dist = (float)__ddiv_rn(fabs(__dadd_rn(__dadd_rn(__dmul_rn(A, px), __dmul_rn(B, py)), C)), D);
        if (!(dist > 3.f)) {
          const double ax = __dsub_rn(lx1, px), ay = __dsub_rn(ly1, py), bx = __dsub_rn(lx2, px), by = __dsub_rn(ly2, py);
          const double side1 = __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), side2 = __dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by));
          ok = side1 <= 9.0 || side2 <= 9.0 || ((side1 < __dadd_rn(line_side, side2)) && (side2 < __dadd_rn(line_side, side1)));
        }
      }
    }
    // ordered compaction of this chunk
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const int pos = before + __popc(bal & ((1u << lane) - 1));
    if (ok) {
      if (pos < rel_cap) { rel_idx[o + pos] = j; rel_dist[o + pos] = dist; }
      else atomicExch(overflow, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base += s_warp[0] + s_warp[1] + s_warp[2] + s_warp[3];
    __syncthreads();
  }
  if (threadIdx.x == 0) rel_n[(long long)b * max_lines + i] = min(s_base, rel_cap);
}

// ---- MatchLines: one CTA per stereo pair -----------------------------------------------------------------------------------------------------
constexpr int kPtLines = 8;     // lines one point may belong to (assigned_lines0[idx]); more raises the overflow flag
__global__ void __launch_bounds__(256) match_lines_kernel(const int* __restrict__ n_lines, const int* __restrict__ n_feat, int max_lines, int rel_cap,
                                                          const int* __restrict__ rel_n, const int* __restrict__ rel_idx, const float* __restrict__ feat,
                                                          int feat_stride, const int* __restrict__ m_idx, const int* __restrict__ m_count, int m_cap,
                                                          double min_x_diff, double max_x_diff, double max_y_diff, int* __restrict__ cnt /*[pairs][max_lines^2]*/,
                                                          int* __restrict__ row_loc /*[pairs][max_lines]*/, int* __restrict__ line_matches /*[pairs][max_lines]*/,
                                                          int* __restrict__ overflow) {
  extern __shared__ int sm[];
  const int p = blockIdx.x;
  const int L0 = min(n_lines[2 * p], max_lines), L1 = min(n_lines[2 * p + 1], max_lines);
  const int N0 = n_feat[2 * p], N1 = n_feat[2 * p + 1];
  int* lm = line_matches + (long long)p * max_lines;
  for (int i = threadIdx.x; i < max_lines; i += blockDim.x) lm[i] = -1;
  if (N0 == 0 || N1 == 0 || L0 == 0 || L1 == 0) return;
  // per-point line lists of both images in shared memory: [kKp][1 + kPtLines]
  const int kp_cap = feat_stride;
  int* pl0 = sm;
  int* pl1 = sm + kp_cap * (1 + kPtLines);
  for (int i = threadIdx.x; i < 2 * kp_cap; i += blockDim.x) sm[i * (1 + kPtLines)] = 0;
  int* c = cnt + (long long)p * max_lines * max_lines;
  for (int i = threadIdx.x; i < L0 * L1; i += blockDim.x) c[i] = 0;
  __syncthreads();
  for (int side = 0; side < 2; ++side) {
    const int L = side ? L1 : L0;
    int* pl = side ? pl1 : pl0;
    const long long b = 2 * p + side;
    for (int e = threadIdx.x; e < L * rel_cap; e += blockDim.x) {
      const int i = e / rel_cap, k = e - i * rel_cap;
      if (k < rel_n[b * max_lines + i]) {
        const int j = rel_idx[(b * max_lines + i) * rel_cap + k];
        const int slot = atomicAdd(&pl[j * (1 + kPtLines)], 1);
        if (slot < kPtLines) pl[j * (1 + kPtLines) + 1 + slot] = i;
        else atomicExch(overflow, 1);
      }
    }
  }
  __syncthreads();
  // stereo filter (frame.cc:141-155) + votes
  const int nm = min(m_count[p], m_cap);
  const float* f0 = feat + (long long)(2 * p) * feat_stride * 259;
  const float* f1 = feat + (long long)(2 * p + 1) * feat_stride * 259;
  for (int m = threadIdx.x; m < nm; m += blockDim.x) {
    const int q = m_idx[((long long)p * m_cap + m) * 2], t = m_idx[((long long)p * m_cap + m) * 2 + 1];
    const double dx = (double)fabsf(__fsub_rn(f0[(long long)q * 259 + 1], f1[(long long)t * 259 + 1]));
    const double dy = (double)fabsf(__fsub_rn(f0[(long long)q * 259 + 2], f1[(long long)t * 259 + 2]));
    if (!(dx > min_x_diff && dx < max_x_diff && dy <= max_y_diff)) continue;
    const int n0 = min(pl0[q * (1 + kPtLines)], kPtLines), n1 = min(pl1[t * (1 + kPtLines)], kPtLines);
    for (int a = 0; a < n0; ++a)
      for (int bq = 0; bq < n1; ++bq) atomicAdd(&c[pl0[q * (1 + kPtLines) + 1 + a] * L1 + pl1[t * (1 + kPtLines) + 1 + bq]], 1);
  }
  __threadfence_block();
  __syncthreads();
  int* rl = row_loc + (long long)p * max_lines;
  for (int i = threadIdx.x; i < L0; i += blockDim.x) {        // Eigen maxCoeff(&index): first maximum
    int best = c[i * L1], loc = 0;
    for (int j = 1; j < L1; ++j) { const int v = c[i * L1 + j]; if (v > best) { best = v; loc = j; } }
    rl[i] = loc;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < L1; j += blockDim.x) {
    int best = c[j], loc = 0;
    for (int i = 1; i < L0; ++i) { const int v = c[i * L1 + j]; if (v > best) { best = v; loc = i; } }
    if (best < 2 || rl[loc] != j) continue;
    const int s0 = rel_n[(long long)(2 * p) * max_lines + loc], s1 = rel_n[(long long)(2 * p + 1) * max_lines + j];
    const float score = __fdiv_rn((float)(best * best), (float)(s0 < s1 ? s0 : s1));
    if ((double)score < 0.8) continue;
    lm[loc] = j;                                             // a row has one maximum location: no two columns write the same entry
  }
}

void launch_line_assoc(const float* lines512, const int* n_lines, int line_stride, const float* feat, const int* n_feat, int feat_stride, double ws, double hs,
                       const int* m_idx, const int* m_count, int m_cap, int pairs, double min_x_diff, double max_x_diff, double max_y_diff, int max_lines,
                       int rel_cap, int* rel_n, int* rel_idx, float* rel_dist, int* cnt, int* row_loc, int* line_matches, int* overflow, cudaStream_t st) {
  cudaMemsetAsync(overflow, 0, 4, st);
  cudaMemsetAsync(rel_n, 0, (size_t)2 * pairs * max_lines * 4, st);
  assign_points_kernel<<<dim3(max_lines, 2 * pairs), 128, 0, st>>>(lines512, n_lines, line_stride, feat, n_feat, feat_stride, ws, hs, max_lines, rel_cap, rel_n, rel_idx,
                                                                 rel_dist, overflow);
  const int smem = 2 * feat_stride * (1 + kPtLines) * 4;
  static bool attr_set[kMaxDevices] = {};
  const int dev = current_device();
  if (!attr_set[dev]) { cudaFuncSetAttribute(match_lines_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 1024 * (1 + kPtLines) * 4); attr_set[dev] = true; }
  match_lines_kernel<<<pairs, 256, smem, st>>>(n_lines, n_feat, max_lines, rel_cap, rel_n, rel_idx, feat, feat_stride, m_idx, m_count, m_cap, min_x_diff, max_x_diff,
                                                max_y_diff, cnt, row_loc, line_matches, overflow);
}

}  // namespace airfe
