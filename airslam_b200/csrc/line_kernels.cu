// PLNet line path on the device: HAFM decode, junction NMS + top-300, line<->junction association,
// unique junction pairs (wireframe_matcher, src/plnet.cpp:272-307), LOI feature gather for stage 1 (G3), line scoring head,
// line acceptance + junction map (src/plnet.cpp:519-558) and junction keypoints (junction_detector, :425-448).
// The reference does all of this either inside the TensorRT graph or on the CPU with three PCIe hops; here it is a chain of
// small HBM/shared-memory kernels that never leaves the device.  Arithmetic follows the ONNX graph op by op (no FMA
// contraction where a discrete decision depends on the result: __fmul_rn / __fadd_rn).
#include "line_kernels.h"
#include <stdlib.h>
#include <math.h>

namespace airfe {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float clampf_(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// ---- K9a: HAFM decode -> lines_pred [B][3*128*128][4], jloc [B][128*128]   (G2 nodes 279-450, App. B3) -----------------
__global__ void hafm_decode_kernel(const float* __restrict__ heads, int ld, float* __restrict__ lines, float* __restrict__ jloc) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // pixel 0..16383
  const float* h = heads + ((long long)b * 16384 + i) * ld;
  const float md0 = sigmoidf_(h[0]), md1 = sigmoidf_(h[1]), md2 = sigmoidf_(h[2]);
  const float dis = sigmoidf_(h[3]), res = sigmoidf_(h[4]);
  const float l0 = h[5], l1 = h[6];
  const float m = fmaxf(l0, l1);
  const float e0 = expf(l0 - m), e1 = expf(l1 - m);
  jloc[(long long)b * 16384 + i] = e1 / (e0 + e1);
  const float pi = 3.1415927410125732f;
  const float th = __fmul_rn(__fmul_rn(__fsub_rn(md0, 0.5f), pi), 2.f);
  const float ta = tanf(__fdiv_rn(__fmul_rn(md1, pi), 2.f));
  const float tb = tanf(__fdiv_rn(__fmul_rn(-md2, pi), 2.f));
  const float c = cosf(th), s = sinf(th);
  const float ux_a = __fsub_rn(c, __fmul_rn(s, ta)), uy_a = __fadd_rn(s, __fmul_rn(c, ta));
  const float ux_b = __fsub_rn(c, __fmul_rn(s, tb)), uy_b = __fadd_rn(s, __fmul_rn(c, tb));
  const float x = (float)(i & 127), y = (float)(i >> 7);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float d = clampf_(__fadd_rn(dis, __fmul_rn(res, (float)(k - 1))), 0.f, 1.f);
    float4 o;
    o.x = clampf_(__fadd_rn(__fmul_rn(__fmul_rn(ux_a, d), 2.f), x), 0.f, 127.f);
    o.y = clampf_(__fadd_rn(__fmul_rn(__fmul_rn(uy_a, d), 2.f), y), 0.f, 127.f);
    o.z = clampf_(__fadd_rn(__fmul_rn(__fmul_rn(ux_b, d), 2.f), x), 0.f, 127.f);
    o.w = clampf_(__fadd_rn(__fmul_rn(__fmul_rn(uy_b, d), 2.f), y), 0.f, 127.f);
    *reinterpret_cast<float4*>(lines + ((long long)b * kProposals + k * 16384 + i) * 4) = o;
  }
}

// ---- K9b: junction 3x3 NMS + TopK(300) (value desc, index asc) + offsets -> juncs [B][300][2] ---------------------------
__global__ void junc_peaks_kernel(const float* __restrict__ jloc, int* __restrict__ peaks, int* __restrict__ n_peaks, uint8_t* __restrict__ is_peak) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float* j = jloc + (long long)b * 16384;
  const int x = i & 127, y = i >> 7;
  const float v = j[i];
  float m = v;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = x + dx, yy = y + dy;
      if (xx >= 0 && xx < 128 && yy >= 0 && yy < 128) m = fmaxf(m, j[yy * 128 + xx]);
    }
  const bool pk = (v == m) && (v > 0.f);   // J = jloc * (jloc == maxpool): non-peaks become exactly 0
  is_peak[(long long)b * 16384 + i] = pk;
  const unsigned bal = __ballot_sync(0xffffffffu, pk);
  if (!bal) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(n_peaks + b, __popc(bal));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (pk) peaks[(long long)b * 16384 + base + __popc(bal & ((1u << lane) - 1))] = i;
}

__global__ void junc_topk_kernel(const float* __restrict__ jloc, const float* __restrict__ heads, int ld, const int* __restrict__ peaks,
                                 const int* __restrict__ n_peaks, const uint8_t* __restrict__ is_peak, float* __restrict__ juncs,
                                 int* __restrict__ junc_idx) {
  const int b = blockIdx.y;
  const int n = n_peaks[b];
  const float* j = jloc + (long long)b * 16384;
  const int* pk = peaks + (long long)b * 16384;
  __shared__ int s_idx[256];
  __shared__ float s_val[256];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int my_idx = 0;
  float my_val = 0.f;
  if (i < n) { my_idx = pk[i]; my_val = j[my_idx]; }
  int rank = 0;
  for (int base = 0; base < n; base += 256) {
    const int q = base + threadIdx.x;
    __syncthreads();
    if (q < n) { const int id = pk[q]; s_idx[threadIdx.x] = id; s_val[threadIdx.x] = j[id]; }
    __syncthreads();
    const int lim = min(256, n - base);
    if (i < n)
      for (int k = 0; k < lim; ++k) rank += (s_val[k] > my_val) || (s_val[k] == my_val && s_idx[k] < my_idx);
  }
  auto emit = [&](int pos, int idx) {
    const float* h = heads + ((long long)b * 16384 + idx) * ld;
    const float jox = sigmoidf_(h[7]) - 0.5f, joy = sigmoidf_(h[8]) - 0.5f;
    float* o = juncs + ((long long)b * kJunctions + pos) * 2;
    o[0] = __fadd_rn(__fadd_rn((float)(idx & 127), jox), 0.5f);
    o[1] = __fadd_rn(__fadd_rn((float)(idx >> 7), joy), 0.5f);
    junc_idx[(long long)b * kJunctions + pos] = idx;
  };
  if (i < n && rank < kJunctions) emit(rank, my_idx);
  // fewer than 300 peaks: TopK continues with the zero-valued cells in index order
  if (n < kJunctions && blockIdx.x == 0 && threadIdx.x == 0) {
    int pos = n;
    const uint8_t* ip = is_peak + (long long)b * 16384;
    for (int idx = 0; idx < 16384 && pos < kJunctions; ++idx)
      if (!ip[idx]) emit(pos++, idx);
  }
}

// ---- K10a: association: nearest junction (first index on ties) of both endpoints of every proposal ----------------------
// Only proposals whose two endpoints both lie within sqrt(10) px of a junction survive (keep), and nothing downstream reads
// imin / imax of the others, so the 300-junction scan is pruned with a 32x32 grid of 4x4-px cells built per CTA in shared memory:
// a junction closer than sqrt(10) must sit in the 3x3 cells around the endpoint.  Candidates are compared on (distance, index) so
// the result is the sequential first-minimum of the reference.  A cell with more than kCellCap junctions (TopK padding with
// non-peak cells makes runs of adjacent junctions possible) sends the whole CTA down the exact full scan.
constexpr int kCellCap = 8;
constexpr int kAssocPer = 4;     // proposals per thread
__global__ void __launch_bounds__(256) assoc_kernel(const float* __restrict__ lines, const float* __restrict__ juncs, int* __restrict__ imin_o,
                                                    int* __restrict__ imax_o, uint8_t* __restrict__ keep_o, int* __restrict__ pair_table) {
  __shared__ float2 sj[kJunctions];
  __shared__ int s_cnt[1024];
  __shared__ unsigned short s_list[1024 * kCellCap];
  __shared__ int s_overflow;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < kJunctions; i += blockDim.x) sj[i] = reinterpret_cast<const float2*>(juncs)[(long long)b * kJunctions + i];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_cnt[i] = 0;
  if (threadIdx.x == 0) s_overflow = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < kJunctions; i += blockDim.x) {
    const float2 p = sj[i];
    const int cx = min(31, max(0, (int)(p.x * 0.25f))), cy = min(31, max(0, (int)(p.y * 0.25f)));
    const int slot = atomicAdd(&s_cnt[cy * 32 + cx], 1);
    if (slot < kCellCap) s_list[(cy * 32 + cx) * kCellCap + slot] = (unsigned short)i;
    else s_overflow = 1;
  }
  __syncthreads();
  const bool full_scan = s_overflow != 0;
#pragma unroll 1
  for (int r = 0; r < kAssocPer; ++r) {
    const int k = (blockIdx.x * kAssocPer + r) * blockDim.x + threadIdx.x;
    if (k >= kProposals) break;
    const float4 l = *reinterpret_cast<const float4*>(lines + ((long long)b * kProposals + k) * 4);
    float m1 = INFINITY, m2 = INFINITY;
    int i1 = 0, i2 = 0;
    if (full_scan) {
      for (int q = 0; q < kJunctions; ++q) {
        const float2 p = sj[q];
        const float dx1 = __fsub_rn(l.x, p.x), dy1 = __fsub_rn(l.y, p.y);
        const float dx2 = __fsub_rn(l.z, p.x), dy2 = __fsub_rn(l.w, p.y);
        const float d1 = __fadd_rn(__fmul_rn(dx1, dx1), __fmul_rn(dy1, dy1));
        const float d2 = __fadd_rn(__fmul_rn(dx2, dx2), __fmul_rn(dy2, dy2));
        if (d1 < m1) { m1 = d1; i1 = q; }
        if (d2 < m2) { m2 = d2; i2 = q; }
      }
    } else {
      auto nearest = [&](float ex, float ey, float& m, int& im) {
        const int cx0 = max(0, (int)floorf((ex - 3.17f) * 0.25f)), cx1 = min(31, (int)floorf((ex + 3.17f) * 0.25f));
        const int cy0 = max(0, (int)floorf((ey - 3.17f) * 0.25f)), cy1 = min(31, (int)floorf((ey + 3.17f) * 0.25f));
        for (int cy = cy0; cy <= cy1; ++cy)
          for (int cx = cx0; cx <= cx1; ++cx) {
            const int c = cy * 32 + cx, n = s_cnt[c];
            for (int e = 0; e < n; ++e) {
              const int q = s_list[c * kCellCap + e];
              const float2 p = sj[q];
              const float dx = __fsub_rn(ex, p.x), dy = __fsub_rn(ey, p.y);
              const float d = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
              if (d < m || (d == m && q < im)) { m = d; im = q; }
            }
          }
      };
      nearest(l.x, l.y, m1, i1);
      nearest(l.z, l.w, m2, i2);
      if (!(m1 < 10.f) || !(m2 < 10.f)) { i1 = 0; i2 = 0; }   // not kept: indices are never read
    }
    const int lo = min(i1, i2), hi = max(i1, i2);
    const bool keep = (lo < hi) && (m1 < 10.f) && (m2 < 10.f);
    const long long o = (long long)b * kProposals + k;
    imin_o[o] = lo; imax_o[o] = hi; keep_o[o] = keep;
    if (keep) atomicMin(pair_table + (long long)b * kJunctions * kJunctions + lo * kJunctions + hi, k);
  }
}

// ---- block-wide exclusive scan helper (1024 threads) --------------------------------------------------------------------
__device__ int block_excl_scan_1024(int v, int* total) {
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  const int base = warp ? warp_sums[warp - 1] : 0;
  *total = warp_sums[31];
  return base + inc - v;
}

// ---- K10b: unique (min,max) pairs in first-seen order (wireframe_matcher).  One CTA per image. -----------------------------
// A kept proposal k is the FIRST of its pair iff pair_table[pair] == k; its unique id is the number of first-proposals before it.
__global__ void __launch_bounds__(1024) unique_pairs_kernel(const int* __restrict__ imin, const int* __restrict__ imax,
                                                            const uint8_t* __restrict__ keep, const int* __restrict__ pair_table,
                                                            int* __restrict__ uid_pairs, int* __restrict__ uid_first,
                                                            int* __restrict__ n_unique, int line_cap) {
  const int b = blockIdx.x;
  constexpr int PER = kProposals / 1024;  // 48
  const long long o = (long long)b * kProposals;
  const int* tab = pair_table + (long long)b * kJunctions * kJunctions;
  const int k0 = threadIdx.x * PER;
  unsigned long long bits = 0;
  int cnt = 0;
  for (int q = 0; q < PER; ++q) {
    const int k = k0 + q;
    const bool first = keep[o + k] && tab[imin[o + k] * kJunctions + imax[o + k]] == k;
    if (first) { bits |= 1ull << q; ++cnt; }
  }
  int total;
  int pos = block_excl_scan_1024(cnt, &total);
  for (int q = 0; q < PER; ++q)
    if (bits >> q & 1) {
      const int k = k0 + q;
      if (pos < line_cap) {
        uid_pairs[((long long)b * line_cap + pos) * 2 + 0] = imax[o + k];   // (max, min): the swap at plnet.cpp:301
        uid_pairs[((long long)b * line_cap + pos) * 2 + 1] = imin[o + k];
        uid_first[(long long)b * line_cap + pos] = k;
      }
      ++pos;
    }
  if (threadIdx.x == 0) n_unique[b] = min(total, line_cap);
}

// ---- K11a: LOI feature gather (G3, App. B5).  One warp per unique pair -> 512 fp16 (496 used) -------------------------------
__device__ __forceinline__ void bil_setup(float px, float py, int& x0, int& y0, int& x1, int& y1, float& w00, float& w10, float& w01,
                                          float& w11) {
  px -= 0.5f; py -= 0.5f;
  const float fx0 = clampf_(floorf(px), 0.f, 127.f), fy0 = clampf_(floorf(py), 0.f, 127.f);
  const float fx1 = fminf(fx0 + 1.f, 127.f), fy1 = fminf(fy0 + 1.f, 127.f);
  x0 = (int)fx0; y0 = (int)fy0; x1 = (int)fx1; y1 = (int)fy1;
  w00 = (fy1 - py) * (fx1 - px);   // f[y0][x0]
  w10 = (py - fy0) * (fx1 - px);   // f[y1][x0]
  w01 = (fy1 - py) * (px - fx0);   // f[y0][x1]
  w11 = (py - fy0) * (px - fx0);   // f[y1][x1]
}

// Endpoint features once per junction (300 per image) instead of once per line end (~2 x 4300 per image): a junction is an end point of ~29
// candidate lines, and the bilinear sample of the 128-channel LOI map at a junction does not depend on the line.  One warp per junction,
// the arithmetic of the per-line version below verbatim (results are bit-identical); the line kernel then copies two 256-byte rows.
__global__ void junc_feat_kernel(const float* __restrict__ loi, int loi_ld, const float* __restrict__ juncs, __half* __restrict__ jf) {
  const int b = blockIdx.y;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= kJunctions) return;
  const float2 p = reinterpret_cast<const float2*>(juncs)[(long long)b * kJunctions + j];
  const float* L = loi + (long long)b * 16384 * loi_ld;
  int x0, y0, x1, y1; float w00, w10, w01, w11;
  bil_setup(p.x, p.y, x0, y0, x1, y1, w00, w10, w01, w11);
  const float4 a = *reinterpret_cast<const float4*>(L + (long long)(y0 * 128 + x0) * loi_ld + lane * 4);
  const float4 c = *reinterpret_cast<const float4*>(L + (long long)(y1 * 128 + x0) * loi_ld + lane * 4);
  const float4 d = *reinterpret_cast<const float4*>(L + (long long)(y0 * 128 + x1) * loi_ld + lane * 4);
  const float4 g = *reinterpret_cast<const float4*>(L + (long long)(y1 * 128 + x1) * loi_ld + lane * 4);
  const float r0 = a.x * w00 + c.x * w10 + d.x * w01 + g.x * w11;
  const float r1 = a.y * w00 + c.y * w10 + d.y * w01 + g.y * w11;
  const float r2 = a.z * w00 + c.z * w10 + d.z * w01 + g.z * w11;
  const float r3 = a.w * w00 + c.w * w10 + d.w * w01 + g.w * w11;
  __half2 h0 = __floats2half2_rn(r0, r1), h1 = __floats2half2_rn(r2, r3);
  *reinterpret_cast<uint2*>(jf + ((long long)b * kJunctions + j) * 128 + lane * 4) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
}

__global__ void loi_gather_kernel(const float* __restrict__ loi, int loi_ld, const float* __restrict__ thinaux, int ta_ld,
                                  const float* __restrict__ juncs, const float* __restrict__ lines, const int* __restrict__ uid_pairs,
                                  const int* __restrict__ uid_first, const int* __restrict__ n_unique, int line_cap,
                                  const float* __restrict__ tspan, __half* __restrict__ feat, float* __restrict__ adj_out,
                                  const __half* __restrict__ jf /* junc_feat_kernel's rows, or nullptr: sample per line end */) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int n_lines_b = n_unique[b];
  // grid-stride over the image's lines: the grid holds kLineGridX blocks per image instead of one warp per CAPACITY row (16 384 rows, of which
  // a typical image uses ~4 300: three quarters of the former 2 048 blocks per image only read the count and left)
  for (int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n_lines_b; u += (gridDim.x * blockDim.x) >> 5) {
  const long long ub = (long long)b * line_cap + u;
  const int ja = uid_pairs[ub * 2], jb = uid_pairs[ub * 2 + 1];
  const float2 pa = reinterpret_cast<const float2*>(juncs)[(long long)b * kJunctions + ja];
  const float2 pb = reinterpret_cast<const float2*>(juncs)[(long long)b * kJunctions + jb];
  if (lane == 0) *reinterpret_cast<float4*>(adj_out + ub * 4) = make_float4(pa.x, pa.y, pb.x, pb.y);
  __half* f = feat + ub * 512;
  const float* L = loi + (long long)b * 16384 * loi_ld;
  // endpoint features: 128 channels, 4 per lane
  if (jf) {
    *reinterpret_cast<uint2*>(f + lane * 4) = *reinterpret_cast<const uint2*>(jf + ((long long)b * kJunctions + ja) * 128 + lane * 4);
    *reinterpret_cast<uint2*>(f + 128 + lane * 4) = *reinterpret_cast<const uint2*>(jf + ((long long)b * kJunctions + jb) * 128 + lane * 4);
  } else
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const float2 p = e ? pb : pa;
    int x0, y0, x1, y1; float w00, w10, w01, w11;
    bil_setup(p.x, p.y, x0, y0, x1, y1, w00, w10, w01, w11);
    const float4 a = *reinterpret_cast<const float4*>(L + (long long)(y0 * 128 + x0) * loi_ld + lane * 4);
    const float4 c = *reinterpret_cast<const float4*>(L + (long long)(y1 * 128 + x0) * loi_ld + lane * 4);
    const float4 d = *reinterpret_cast<const float4*>(L + (long long)(y0 * 128 + x1) * loi_ld + lane * 4);
    const float4 g = *reinterpret_cast<const float4*>(L + (long long)(y1 * 128 + x1) * loi_ld + lane * 4);
    const float r0 = a.x * w00 + c.x * w10 + d.x * w01 + g.x * w11;
    const float r1 = a.y * w00 + c.y * w10 + d.y * w01 + g.y * w11;
    const float r2 = a.z * w00 + c.z * w10 + d.z * w01 + g.z * w11;
    const float r3 = a.w * w00 + c.w * w10 + d.w * w01 + g.w * w11;
    __half2 h0 = __floats2half2_rn(r0, r1), h1 = __floats2half2_rn(r2, r3);
    uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
    *reinterpret_cast<uint2*>(f + e * 128 + lane * 4) = pk;
  }
  // 30 samples along the adjusted line on `thin` (channels 0-3) and along the original proposal on `aux` (channels 4-7)
  const float4 orig = *reinterpret_cast<const float4*>(lines + ((long long)b * kProposals + uid_first[ub]) * 4);
  const float* TA = thinaux + (long long)b * 16384 * ta_ld;
  if (lane < 30) {
    const float t = tspan[lane], tc = tspan[30 + lane];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const float ax = s ? orig.x : pa.x, ay = s ? orig.y : pa.y, bx = s ? orig.z : pb.x, by = s ? orig.w : pb.y;
      const float px = __fadd_rn(__fmul_rn(ax, t), __fmul_rn(bx, tc));
      const float py = __fadd_rn(__fmul_rn(ay, t), __fmul_rn(by, tc));
      int x0, y0, x1, y1; float w00, w10, w01, w11;
      bil_setup(px, py, x0, y0, x1, y1, w00, w10, w01, w11);
      const float4 a = *reinterpret_cast<const float4*>(TA + (long long)(y0 * 128 + x0) * ta_ld + s * 4);
      const float4 c = *reinterpret_cast<const float4*>(TA + (long long)(y1 * 128 + x0) * ta_ld + s * 4);
      const float4 d = *reinterpret_cast<const float4*>(TA + (long long)(y0 * 128 + x1) * ta_ld + s * 4);
      const float4 g = *reinterpret_cast<const float4*>(TA + (long long)(y1 * 128 + x1) * ta_ld + s * 4);
      __half* o = f + 256 + s * 120;   // channel-major: [ch][30]
      o[0 * 30 + lane] = __float2half_rn(a.x * w00 + c.x * w10 + d.x * w01 + g.x * w11);
      o[1 * 30 + lane] = __float2half_rn(a.y * w00 + c.y * w10 + d.y * w01 + g.y * w11);
      o[2 * 30 + lane] = __float2half_rn(a.z * w00 + c.z * w10 + d.z * w01 + g.z * w11);
      o[3 * 30 + lane] = __float2half_rn(a.w * w00 + c.w * w10 + d.w * w01 + g.w * w11);
    }
  }
  if (lane < 16) f[496 + lane] = __float2half_rn(0.f);
  }
}

// ---- K11b: residual add + fc2_head (128 -> 2) + softmax[1]; one warp per line ------------------------------------------------
__global__ void line_head_kernel(const float* __restrict__ h1, const float* __restrict__ h2, const float* __restrict__ w /*[2][128]*/,
                                 const float* __restrict__ bias, const int* __restrict__ n_unique, int line_cap,
                                 float* __restrict__ score) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int n_lines_b = n_unique[b];
  for (int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n_lines_b; u += (gridDim.x * blockDim.x) >> 5) {
  const long long ub = (long long)b * line_cap + u;
  const float4 a = *reinterpret_cast<const float4*>(h1 + ub * 128 + lane * 4);
  const float4 c = *reinterpret_cast<const float4*>(h2 + ub * 128 + lane * 4);
  // operand rounding of the head GEMM (fp16 operands, fp32 accumulate), like every other contraction on the path
  const float x0 = __half2float(__float2half_rn(a.x + c.x)), x1 = __half2float(__float2half_rn(a.y + c.y));
  const float x2 = __half2float(__float2half_rn(a.z + c.z)), x3 = __half2float(__float2half_rn(a.w + c.w));
  const float4 w0 = *reinterpret_cast<const float4*>(w + lane * 4);
  const float4 w1 = *reinterpret_cast<const float4*>(w + 128 + lane * 4);
  float l0 = x0 * w0.x + x1 * w0.y + x2 * w0.z + x3 * w0.w;
  float l1 = x0 * w1.x + x1 * w1.y + x2 * w1.z + x3 * w1.w;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    l0 += __shfl_xor_sync(0xffffffffu, l0, o);
    l1 += __shfl_xor_sync(0xffffffffu, l1, o);
  }
  if (lane == 0) {
    l0 += bias[0]; l1 += bias[1];
    const float m = fmaxf(l0, l1);
    const float e0 = expf(l0 - m), e1 = expf(l1 - m);
    score[ub] = e1 / (e0 + e1);
  }
  }
}

// ---- K12a: line acceptance + junction map (src/plnet.cpp:519-558).  One CTA per image, ordered compaction. ------------------
__global__ void __launch_bounds__(1024) line_accept_kernel(const float* __restrict__ adj, const float* __restrict__ score,
                                                           const int* __restrict__ n_unique, int line_cap, float line_thr, float len2_thr,
                                                           int border, uint8_t* __restrict__ junc_map, float* __restrict__ lines_out,
                                                           int* __restrict__ n_lines) {
  const int b = blockIdx.x;
  const int n = n_unique[b];
  uint8_t* jm = junc_map + (long long)b * 262144;
  int written = 0;
  for (int base = 0; base < n; base += 1024) {
    const int u = base + threadIdx.x;
    bool acc = false;
    float x1 = 0, y1 = 0, x2 = 0, y2 = 0;
    if (u < n) {
      const long long ub = (long long)b * line_cap + u;
      const float sc = score[ub];
      if (!(sc < 0.5f)) {
        const float4 a = *reinterpret_cast<const float4*>(adj + ub * 4);
        x1 = __fmul_rn(a.x, 4.f); y1 = __fmul_rn(a.y, 4.f); x2 = __fmul_rn(a.z, 4.f); y2 = __fmul_rn(a.w, 4.f);
        const int xi1 = (int)__fadd_rn(x1, 0.1f), yi1 = (int)__fadd_rn(y1, 0.1f), xi2 = (int)__fadd_rn(x2, 0.1f), yi2 = (int)__fadd_rn(y2, 0.1f);
        const bool p1 = xi1 > border && xi1 < 512 - border && yi1 > border && yi1 < 512 - border;
        const bool p2 = xi2 > border && xi2 < 512 - border && yi2 > border && yi2 < 512 - border;
        // validity depends on the pixel only, so "last writer wins" of the reference == "any writer"
        if (p1) jm[yi1 * 512 + xi1] = 1;
        if (p2) jm[yi2 * 512 + xi2] = 1;
        if (!(sc < line_thr)) {
          const float dx = __fsub_rn(x2, x1), dy = __fsub_rn(y2, y1);
          const float l2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
          acc = !(l2 < len2_thr);
        }
      }
    }
    int total;
    const int pos = written + block_excl_scan_1024(acc ? 1 : 0, &total);
    if (acc && pos < line_cap) *reinterpret_cast<float4*>(lines_out + ((long long)b * line_cap + pos) * 4) = make_float4(x1, y1, x2, y2);
    written += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) n_lines[b] = min(written, line_cap);
}

// ---- K12b: junction keypoints: raster scan of the junction map inside the border (junction_detector) ----------------------
__global__ void __launch_bounds__(512) junction_scan_kernel(const uint8_t* __restrict__ junc_map, const float* __restrict__ scores,
                                                            int border, float* __restrict__ kp, int kp_cap, int* __restrict__ kp_count) {
  const int b = blockIdx.x;
  const int y = threadIdx.x;
  const uint8_t* jm = junc_map + (long long)b * 262144 + y * 512;
  const bool row_ok = (y >= border) && (y < 512 - border);
  int cnt = 0;
  if (row_ok) {   // 16 bytes at a time; the border columns are masked out afterwards
    const uint4* jv = reinterpret_cast<const uint4*>(jm);
    for (int q = 0; q < 32; ++q) {
      const uint4 v = jv[q];
      if ((v.x | v.y | v.z | v.w) == 0) continue;
      for (int e = 0; e < 16; ++e) {
        const int x = q * 16 + e;
        if (x >= border && x < 512 - border && jm[x]) ++cnt;
      }
    }
  }
  // exclusive scan over 512 rows
  __shared__ int ws[16];
  const int lane = y & 31, warp = y >> 5;
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < 16 ? ws[lane] : 0;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    if (lane < 16) ws[lane] = w;
  }
  __syncthreads();
  int pos = (warp ? ws[warp - 1] : 0) + inc - cnt;
  if (y == 0) kp_count[b] = min(ws[15], kp_cap);
  if (row_ok && cnt)
    for (int x = border; x < 512 - border; ++x)
      if (jm[x]) {   // rows with junctions are rare (<= 300 per image): this scalar pass runs on a handful of rows only
        if (pos < kp_cap) {
          float* o = kp + ((long long)b * kp_cap + pos) * 3;
          o[0] = scores[(long long)b * 262144 + y * 512 + x];
          o[1] = (float)x;
          o[2] = (float)y;
        }
        ++pos;
      }
}

// ---- launchers ------------------------------------------------------------------------------------------------------------
void launch_hafm_decode(const float* heads, int ld, float* lines, float* jloc, int batch, cudaStream_t st) {
  hafm_decode_kernel<<<dim3(16384 / 256, batch), 256, 0, st>>>(heads, ld, lines, jloc);
}
void launch_junctions(const float* jloc, const float* heads, int ld, int* peaks, int* n_peaks, uint8_t* is_peak, float* juncs,
                      int* junc_idx, int batch, cudaStream_t st) {
  cudaMemsetAsync(n_peaks, 0, sizeof(int) * batch, st);
  junc_peaks_kernel<<<dim3(16384 / 256, batch), 256, 0, st>>>(jloc, peaks, n_peaks, is_peak);
  junc_topk_kernel<<<dim3(16384 / 256, batch), 256, 0, st>>>(jloc, heads, ld, peaks, n_peaks, is_peak, juncs, junc_idx);
}
void launch_association(const float* lines, const float* juncs, int* imin, int* imax, uint8_t* keep, int* pair_table, int* uid_pairs,
                        int* uid_first, int* n_unique, int line_cap, int batch, cudaStream_t st) {
  cudaMemsetAsync(pair_table, 0x7f, sizeof(int) * (size_t)batch * kJunctions * kJunctions, st);
  assoc_kernel<<<dim3((kProposals + 256 * kAssocPer - 1) / (256 * kAssocPer), batch), 256, 0, st>>>(lines, juncs, imin, imax, keep, pair_table);
  unique_pairs_kernel<<<batch, 1024, 0, st>>>(imin, imax, keep, pair_table, uid_pairs, uid_first, n_unique, line_cap);
}
// blocks of 8 warps per image for the one-warp-per-line kernels (grid-stride over the device-side line count): 256 blocks = 2 048 lines per sweep
static int line_grid_x(int line_cap) { const int full = (line_cap * 32 + 255) / 256; return full < 256 ? full : 256; }

void launch_loi_gather(const float* loi, int loi_ld, const float* thinaux, int ta_ld, const float* juncs, const float* lines,
                       const int* uid_pairs, const int* uid_first, const int* n_unique, int line_cap, const float* tspan, __half* feat,
                       float* adj, int batch, cudaStream_t st, __half* junc_feat) {
  static const bool use_jf = !(getenv("AIRFE_LOI_JF") && atoi(getenv("AIRFE_LOI_JF")) == 0);     // 0: per-line endpoint sampling (A/B timing; identical results)
  __half* jf = use_jf ? junc_feat : nullptr;
  if (jf) junc_feat_kernel<<<dim3((kJunctions * 32 + 255) / 256, batch), 256, 0, st>>>(loi, loi_ld, juncs, jf);
  loi_gather_kernel<<<dim3(line_grid_x(line_cap), batch), 256, 0, st>>>(loi, loi_ld, thinaux, ta_ld, juncs, lines, uid_pairs, uid_first,
                                                                             n_unique, line_cap, tspan, feat, adj, jf);
}
void launch_line_head(const float* h1, const float* h2, const float* w, const float* bias, const int* n_unique, int line_cap, float* score,
                      int batch, cudaStream_t st) {
  line_head_kernel<<<dim3(line_grid_x(line_cap), batch), 256, 0, st>>>(h1, h2, w, bias, n_unique, line_cap, score);
}
void launch_line_accept(const float* adj, const float* score, const int* n_unique, int line_cap, float line_thr, float len_thr, int border,
                        uint8_t* junc_map, float* lines_out, int* n_lines, int batch, cudaStream_t st) {
  cudaMemsetAsync(junc_map, 0, (size_t)batch * 262144, st);
  line_accept_kernel<<<batch, 1024, 0, st>>>(adj, score, n_unique, line_cap, line_thr, len_thr * len_thr, border, junc_map, lines_out, n_lines);
}
void launch_junction_scan(const uint8_t* junc_map, const float* scores, int border, float* kp, int kp_cap, int* kp_count, int batch,
                          cudaStream_t st) {
  junction_scan_kernel<<<batch, 512, 0, st>>>(junc_map, scores, border, kp, kp_cap, kp_count);
}

}  // namespace airfe
