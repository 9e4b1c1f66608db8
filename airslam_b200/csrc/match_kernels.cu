// Hand-written sm_100a kernels for the matchers' non-GEMM stages.
// Data layout: "slot" s = 2 * pair + side owns rows [s*CAP, s*CAP + n[s]) of every [rows, C] matrix; all kernels read the
// per-slot keypoint counts n[] from device memory, so nothing here depends on host-side knowledge of N.
#include "match_kernels.h"
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

namespace airfe {

__device__ __forceinline__ float warp_sum_(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- K13: unpack 259xN host features -> normalised keypoints, fp32 state x, fp16 operand copy, rotary table ---------------
// NormalizeKeypoints (src/point_matcher.cc:39-48): (x - width/2) * L_inv with integer width/2; L_inv = float(1.0/max(w,h)*scale).
// `feat_ptrs` (optional): per-slot base pointers instead of the dense [slot][feat_cap][259] array -- relocalization jobs pair a query
// with a keyframe of the device-resident cache without copying either (airfe_reloc_match).
// `row_off` (optional): PACKED row layout -- slot s owns rows [row_off[s], row_off[s] + n[s]) of every [rows, C] matrix instead of
// [s * cap, ...): the row GEMMs and the fused block tail then process sum(n) rows instead of slots x cap (400 of 512 rows used at 400 keypoints).
__global__ void lg_offsets_kernel(const int* __restrict__ n, int slots, int cap, int* __restrict__ row_off /*[slots + 1]*/) {
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int s = 0; s < slots; ++s) { row_off[s] = acc; acc += min(max(n[s], 0), cap); }
    row_off[slots] = acc;
  }
}

__global__ void lg_prepare_kernel(const float* __restrict__ feat, const float* const* __restrict__ feat_ptrs, const int* __restrict__ n, const int* __restrict__ row_off, int cap, int feat_cap, int width, int height,
                                  float l_inv, const __half* __restrict__ wr /*[32][2]*/, float* __restrict__ x, __half* __restrict__ cat16,
                                  float* __restrict__ rot /*[slots][cap][64] cos | sin interleaved as (cos,sin) per freq*/) {
  const int s = blockIdx.y;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n[s]) return;
  const float* f = feat_ptrs ? feat_ptrs[s] + (long long)r * 259 : feat + ((long long)s * feat_cap + r) * 259;
  const float kx = __fmul_rn(__fsub_rn(f[1], (float)(width / 2)), l_inv);
  const float ky = __fmul_rn(__fsub_rn(f[2], (float)(height / 2)), l_inv);
  const long long row = row_off ? (long long)row_off[s] + r : (long long)s * cap + r;
  // descriptors: fp32 residual stream + fp16 operand copy (cols 0..255 of the [x | msg] concat buffer)
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float v = f[3 + lane * 8 + e];
    x[row * 256 + lane * 8 + e] = v;
    cat16[row * 512 + lane * 8 + e] = __float2half_rn(v);
  }
  // rotary: freq f_j = kpt . Wr[j] (operands rounded to fp16, fp32 accumulate), j = lane
  const float kxr = __half2float(__float2half_rn(kx)), kyr = __half2float(__float2half_rn(ky));
  const float fr = kxr * __half2float(wr[lane * 2]) + kyr * __half2float(wr[lane * 2 + 1]);
  rot[row * 64 + lane * 2] = cosf(fr);
  rot[row * 64 + lane * 2 + 1] = sinf(fr);
}

// ---- K14a: rotary + 64^-1/4 scaling of q,k ; v passthrough.  qkv fp32 [rows][768] = [q | k | v] (weights pre-permuted) ------
__global__ void lg_rotary_kernel(const float* __restrict__ qkv, const float* __restrict__ rot, const int* __restrict__ n, int cap,
                                 __half* __restrict__ q16, __half* __restrict__ k16, __half* __restrict__ v16, float sc) {
  const int s = blockIdx.y;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n[s]) return;
  const long long row = (long long)s * cap + r;
  const float c = rot[row * 64 + lane * 2], sn = rot[row * 64 + lane * 2 + 1];   // pair index j = lane within a head
  const float* in = qkv + row * 768;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const int o = h * 64 + lane * 2;
    const float q0 = in[o], q1 = in[o + 1];
    const float k0 = in[256 + o], k1 = in[256 + o + 1];
    // q*cos + rot_half(q)*sin with rot_half(x0,x1) = (-x1, x0)
    const float rq0 = (q0 * c + (-q1) * sn) * sc, rq1 = (q1 * c + q0 * sn) * sc;
    const float rk0 = (k0 * c + (-k1) * sn) * sc, rk1 = (k1 * c + k0 * sn) * sc;
    *reinterpret_cast<__half2*>(q16 + row * 256 + o) = __floats2half2_rn(rq0, rq1);
    *reinterpret_cast<__half2*>(k16 + row * 256 + o) = __floats2half2_rn(rk0, rk1);
    *reinterpret_cast<__half2*>(v16 + row * 256 + o) = __floats2half2_rn(in[512 + o], in[512 + o + 1]);
  }
}

// ---- K14b: row softmax of attention scores S fp32 [slot][head][cap][cap] -> P fp16, zero beyond the valid columns ------------
// One warp per row.  Valid columns = n[slot ^ col_xor] (cross attention looks at the partner's keypoints).
__global__ void softmax_rows_kernel(const float* __restrict__ S, __half* __restrict__ P, const int* __restrict__ n, int cap, int col_xor) {
  const int s = blockIdx.z, h = blockIdx.y;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n[s]) return;
  const int nc = n[s ^ col_xor];
  const long long base = (((long long)s * 4 + h) * cap + r) * cap;
  const float* in = S + base;
  float m = -INFINITY;
  for (int j = lane; j < nc; j += 32) m = fmaxf(m, in[j]);
  m = warp_max_(m);
  float sum = 0.f;
  for (int j = lane; j < nc; j += 32) sum += expf(in[j] - m);
  sum = warp_sum_(sum);
  const float inv = 1.f / sum;
  __half* out = P + base;
  // zero the whole padded row: columns beyond nc may hold stale probabilities of an earlier, larger pair
  for (int j = lane; j < cap; j += 32) out[j] = __float2half_rn(j < nc ? expf(in[j] - m) * inv : 0.f);
}

// ---- FFN middle: LayerNorm(512, eps 1e-5) + exact GELU, fp32 in -> fp16 operand out.  One warp per row. -----------------------
__global__ void ln_gelu_kernel(const float* __restrict__ h, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const int* __restrict__ n, int cap, __half* __restrict__ out) {
  const int s = blockIdx.y;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n[s]) return;
  const long long row = (long long)s * cap + r;
  const float* in = h + row * 512;
  float v[16];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 a = *reinterpret_cast<const float4*>(in + i * 128 + lane * 4);
    v[i * 4] = a.x; v[i * 4 + 1] = a.y; v[i * 4 + 2] = a.z; v[i * 4 + 3] = a.w;
    sum += a.x + a.y + a.z + a.w;
  }
  const float mean = warp_sum_(sum) * (1.f / 512.f);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; var += d * d; }
  var = warp_sum_(var) * (1.f / 512.f);
  const float rstd = 1.f / sqrtf(var + 1e-5f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = i * 128 + lane * 4 + e;
      const float y = (v[i * 4 + e] - mean) * rstd * gamma[c] + beta[c];
      o[e] = 0.5f * y * (1.f + erff(y * 0.70710678118654752440f));
    }
    __half2 a = __floats2half2_rn(o[0], o[1]), b = __floats2half2_rn(o[2], o[3]);
    *reinterpret_cast<uint2*>(out + row * 512 + i * 128 + lane * 4) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  }
}

// ---- K15: log assignment.  sim fp32 [pair][cap][cap] (rows = image 0, cols = image 1); z = matchability logit ------------------
// x may be packed (row_off != nullptr); logsig is always slot-padded ([slot][cap]), like sim.  With md16_pad != nullptr the kernel also copies the
// packed final-projection rows md16 into the slot-padded operand of the similarity GEMM.
__global__ void matchability_kernel(const float* __restrict__ x, const __half* __restrict__ w /*[256]*/, float bias, const int* __restrict__ n,
                                    const int* __restrict__ row_off, int cap, float* __restrict__ logsig, const __half* __restrict__ md16, __half* __restrict__ md16_pad) {
  const int s = blockIdx.y;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n[s]) return;
  const long long row = (long long)s * cap + r;
  const long long xrow = row_off ? (long long)row_off[s] + r : row;
  if (md16_pad) *reinterpret_cast<uint4*>(md16_pad + row * 256 + lane * 8) = *reinterpret_cast<const uint4*>(md16 + xrow * 256 + lane * 8);
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) acc += __half2float(__float2half_rn(x[xrow * 256 + lane * 8 + e])) * __half2float(w[lane * 8 + e]);
  acc = warp_sum_(acc) + bias;
  // logsigmoid(z) = min(z,0) - log1p(exp(-|z|))
  if (lane == 0) logsig[row] = fminf(acc, 0.f) - log1pf(expf(-fabsf(acc)));
}

__global__ void row_lse_kernel(const float* __restrict__ sim, const int* __restrict__ n, int cap, float* __restrict__ lse) {
  const int p = blockIdx.y;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n[2 * p]) return;
  const int nc = n[2 * p + 1];
  const float* in = sim + ((long long)p * cap + r) * cap;
  float m = -INFINITY;
  for (int j = lane; j < nc; j += 32) m = fmaxf(m, in[j]);
  m = warp_max_(m);
  float sum = 0.f;
  for (int j = lane; j < nc; j += 32) sum += expf(in[j] - m);
  sum = warp_sum_(sum);
  if (lane == 0) lse[(long long)(2 * p) * cap + r] = m + logf(sum);
}

// column LSE: block of 32x8 threads sweeps rows; thread (tx) owns column, ty strides rows; smem combine
__global__ void col_lse_kernel(const float* __restrict__ sim, const int* __restrict__ n, int cap, float* __restrict__ lse) {
  const int p = blockIdx.y;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int nr = n[2 * p], nc = n[2 * p + 1];
  __shared__ float sm[8][33], ss[8][33];
  const float* in = sim + (long long)p * cap * cap;
  float m = -INFINITY;
  if (c < nc) for (int r = threadIdx.y; r < nr; r += 8) m = fmaxf(m, in[(long long)r * cap + c]);
  sm[threadIdx.y][threadIdx.x] = m;
  __syncthreads();
  float mm = -INFINITY;
#pragma unroll
  for (int k = 0; k < 8; ++k) mm = fmaxf(mm, sm[k][threadIdx.x]);
  float sum = 0.f;
  if (c < nc) for (int r = threadIdx.y; r < nr; r += 8) sum += expf(in[(long long)r * cap + c] - mm);
  ss[threadIdx.y][threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.y == 0 && c < nc) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += ss[k][threadIdx.x];
    lse[(long long)(2 * p + 1) * cap + c] = mm + logf(t);
  }
}

// scores(i,j) = (sim - rowlse_i) + (sim - collse_j) + ls0_i + ls1_j   (graph: log_softmax(dim cols) + log_softmax(dim rows) + ...)
__device__ __forceinline__ float lg_score(float sim, float rl, float cl, float l0, float l1) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fsub_rn(sim, rl), __fsub_rn(sim, cl)), l0), l1);
}

// ---- K16: mutual nearest neighbour + threshold (filter_matches, src/light_glue.cpp:214-266) -----------------------------------
__global__ void lg_rowmax_kernel(const float* __restrict__ sim, const float* __restrict__ lse, const float* __restrict__ logsig,
                                 const int* __restrict__ n, int cap, int* __restrict__ row_arg, float* __restrict__ row_val,
                                 float* __restrict__ scores_out /*optional dense [pair][cap][cap]*/) {
  const int p = blockIdx.y;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n[2 * p]) return;
  const int nc = n[2 * p + 1];
  const float* in = sim + ((long long)p * cap + r) * cap;
  const float rl = lse[(long long)(2 * p) * cap + r], l0 = logsig[(long long)(2 * p) * cap + r];
  const float* cl = lse + (long long)(2 * p + 1) * cap;
  const float* l1 = logsig + (long long)(2 * p + 1) * cap;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  for (int j = lane; j < nc; j += 32) {
    const float v = lg_score(in[j], rl, cl[j], l0, l1[j]);
    if (scores_out) scores_out[((long long)p * cap + r) * cap + j] = v;
    if (v > best) { best = v; arg = j; }   // ascending j per lane: first max kept
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  if (lane == 0) { row_arg[(long long)p * cap + r] = (arg == 0x7fffffff) ? -1 : arg; row_val[(long long)p * cap + r] = best; }
}

__global__ void lg_colmax_kernel(const float* __restrict__ sim, const float* __restrict__ lse, const float* __restrict__ logsig,
                                 const int* __restrict__ n, int cap, int* __restrict__ col_arg) {
  const int p = blockIdx.y;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int nr = n[2 * p], nc = n[2 * p + 1];
  __shared__ float sv[8][33];
  __shared__ int sa[8][33];
  const float* in = sim + (long long)p * cap * cap;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  if (c < nc) {
    const float cl = lse[(long long)(2 * p + 1) * cap + c], l1 = logsig[(long long)(2 * p + 1) * cap + c];
    for (int r = threadIdx.y; r < nr; r += 8) {
      const float v = lg_score(in[(long long)r * cap + c], lse[(long long)(2 * p) * cap + r], cl, logsig[(long long)(2 * p) * cap + r], l1);
      if (v > best) { best = v; arg = r; }
    }
  }
  sv[threadIdx.y][threadIdx.x] = best;
  sa[threadIdx.y][threadIdx.x] = arg;
  __syncthreads();
  if (threadIdx.y == 0 && c < nc) {
    for (int k = 1; k < 8; ++k) {
      const float ob = sv[k][threadIdx.x];
      const int oa = sa[k][threadIdx.x];
      if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    col_arg[(long long)p * cap + c] = arg;
  }
}

// ordered compaction of mutual matches with exp(score) > threshold; one CTA (1024 threads) per pair; cap <= 1024
__global__ void __launch_bounds__(1024) lg_filter_kernel(const int* __restrict__ row_arg, const float* __restrict__ row_val,
                                                         const int* __restrict__ col_arg, const int* __restrict__ n, int cap, float thr,
                                                         int* __restrict__ m_idx /*[pair][cap][2]*/, float* __restrict__ m_score,
                                                         int* __restrict__ m_count) {
  const int p = blockIdx.x;
  const int r = threadIdx.x;
  const int nr = n[2 * p];
  bool ok = false;
  int j = 0;
  float e = 0.f;
  if (r < nr && r < cap) {
    j = row_arg[(long long)p * cap + r];
    if (j >= 0 && j < n[2 * p + 1] && col_arg[(long long)p * cap + j] == r) {
      e = expf(row_val[(long long)p * cap + r]);
      ok = e > thr;
    }
  }
  // block exclusive scan of `ok`
  __shared__ int ws[32];
  const int lane = r & 31, warp = r >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, ok);
  if (lane == 0) ws[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    int w = ws[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    ws[lane] = w;
  }
  __syncthreads();
  const int pos = (warp ? ws[warp - 1] : 0) + __popc(bal & ((1u << lane) - 1));
  if (ok) {
    m_idx[((long long)p * cap + pos) * 2] = r;
    m_idx[((long long)p * cap + pos) * 2 + 1] = j;
    m_score[(long long)p * cap + pos] = e;
  }
  if (r == 0) m_count[p] = ws[31];
}

// ---- launchers ------------------------------------------------------------------------------------------------------------------
void launch_lg_prepare(const float* feat, const float* const* feat_ptrs, const int* n, int* row_off, int slots, int cap, int feat_cap, int width, int height, float l_inv, const __half* wr,
                       float* x, __half* cat16, float* rot, cudaStream_t st) {
  if (row_off) lg_offsets_kernel<<<1, 32, 0, st>>>(n, slots, cap, row_off);
  lg_prepare_kernel<<<dim3((cap + 7) / 8, slots), 256, 0, st>>>(feat, feat_ptrs, n, row_off, cap, feat_cap, width, height, l_inv, wr, x, cat16, rot);
}
void launch_lg_rotary(const float* qkv, const float* rot, const int* n, int slots, int cap, __half* q16, __half* k16, __half* v16, cudaStream_t st) {
  lg_rotary_kernel<<<dim3((cap + 7) / 8, slots), 256, 0, st>>>(qkv, rot, n, cap, q16, k16, v16, 0.35355339059327379f /* 64^-1/4 */);
}
void launch_softmax_rows(const float* S, __half* P, const int* n, int slots, int cap, int col_xor, cudaStream_t st) {
  softmax_rows_kernel<<<dim3((cap + 7) / 8, 4, slots), 256, 0, st>>>(S, P, n, cap, col_xor);
}
void launch_ln_gelu(const float* h, const float* gamma, const float* beta, const int* n, int slots, int cap, __half* out, cudaStream_t st) {
  ln_gelu_kernel<<<dim3((cap + 7) / 8, slots), 256, 0, st>>>(h, gamma, beta, n, cap, out);
}
void launch_lg_matchability(const float* x, const __half* wm, float bm, const int* n, const int* row_off, int pairs, int cap, float* logsig, const __half* md16,
                            __half* md16_pad, cudaStream_t st) {
  matchability_kernel<<<dim3((cap + 7) / 8, 2 * pairs), 256, 0, st>>>(x, wm, bm, n, row_off, cap, logsig, md16, md16_pad);
}
void launch_lg_assignment(const float* sim, const int* n, int pairs, int cap, float* logsig,
                          float* lse, int* row_arg, float* row_val, int* col_arg, float thr, int* m_idx, float* m_score, int* m_count,
                          float* scores_out, cudaStream_t st) {
  row_lse_kernel<<<dim3((cap + 7) / 8, pairs), 256, 0, st>>>(sim, n, cap, lse);
  col_lse_kernel<<<dim3((cap + 31) / 32, pairs), dim3(32, 8), 0, st>>>(sim, n, cap, lse);
  lg_rowmax_kernel<<<dim3((cap + 7) / 8, pairs), 256, 0, st>>>(sim, lse, logsig, n, cap, row_arg, row_val, scores_out);
  lg_colmax_kernel<<<dim3((cap + 31) / 32, pairs), dim3(32, 8), 0, st>>>(sim, lse, logsig, n, cap, col_arg);
  lg_filter_kernel<<<pairs, 1024, 0, st>>>(row_arg, row_val, col_arg, n, cap, thr, m_idx, m_score, m_count);
}

}  // namespace airfe

// =====================================================================================================================
// SuperGlue (G5): keypoint-encoder input, log-domain Sinkhorn (App. B7), decode (src/super_glue.cpp:339-367) and the
// mutual-match extraction of PointMatcher::MatchingPoints (src/point_matcher.cc:73-92).
// =====================================================================================================================
namespace airfe {

// kenc input rows [x', y', score, 0...] (fp16, K padded to 64) and residual stream x = descriptors (fp32)
__global__ void sg_prepare_kernel(const float* __restrict__ feat, const float* const* __restrict__ feat_ptrs, const int* __restrict__ n, int cap, int feat_cap, int width, int height,
                                  float l_inv, float* __restrict__ x, __half* __restrict__ kin16) {
  const int s = blockIdx.y;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n[s]) return;
  const float* f = feat_ptrs ? feat_ptrs[s] + (long long)r * 259 : feat + ((long long)s * feat_cap + r) * 259;
  const long long row = (long long)s * cap + r;
#pragma unroll
  for (int e = 0; e < 8; ++e) x[row * 256 + lane * 8 + e] = f[3 + lane * 8 + e];
  __half2 o = __floats2half2_rn(0.f, 0.f);
  if (lane == 0) o = __floats2half2_rn(__fmul_rn(__fsub_rn(f[1], (float)(width / 2)), l_inv), __fmul_rn(__fsub_rn(f[2], (float)(height / 2)), l_inv));
  if (lane == 1) o = __floats2half2_rn(f[0], 0.f);
  *reinterpret_cast<__half2*>(kin16 + row * 64 + lane * 2) = o;
}

void launch_sg_prepare(const float* feat, const float* const* feat_ptrs, const int* n, int slots, int cap, int feat_cap, int width, int height, float l_inv, float* x,
                       __half* kin16, cudaStream_t st) {
  sg_prepare_kernel<<<dim3((cap + 7) / 8, slots), 256, 0, st>>>(feat, feat_ptrs, n, cap, feat_cap, width, height, l_inv, x, kin16);
}

// Z = [[S, bin],[bin, bin]] with leading dimension ld = cap + 1; u = v = 0
__global__ void sg_couplings_kernel(const float* __restrict__ sim, const int* __restrict__ n, int cap, float bin, float* __restrict__ Z,
                                    float* __restrict__ u, float* __restrict__ v) {
  const int p = blockIdx.z;
  const int m = n[2 * p], nn = n[2 * p + 1];
  const int i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int ld = cap + 1;
  if (i > m || j > nn) return;
  float val = bin;
  if (i < m && j < nn) val = sim[((long long)p * cap + i) * cap + j];
  Z[((long long)p * ld + i) * ld + j] = val;
  if (j == 0) u[(long long)p * ld + i] = 0.f;
  if (i == 0) v[(long long)p * ld + j] = 0.f;
}

// u_i = log_mu_i - LSE_j(Z_ij + v_j): one warp per row
__global__ void sg_row_pass_kernel(const float* __restrict__ Z, const int* __restrict__ n, int cap, float* __restrict__ u, const float* __restrict__ v) {
  const int p = blockIdx.y;
  const int m = n[2 * p], nn = n[2 * p + 1];
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i > m) return;
  const int ld = cap + 1;
  const float* z = Z + ((long long)p * ld + i) * ld;
  const float* vv = v + (long long)p * ld;
  float mx = -INFINITY;
  for (int j = lane; j <= nn; j += 32) mx = fmaxf(mx, z[j] + vv[j]);
  mx = warp_max_(mx);
  float s = 0.f;
  for (int j = lane; j <= nn; j += 32) s += expf(z[j] + vv[j] - mx);
  s = warp_sum_(s);
  if (lane == 0) {
    const float norm = -logf((float)(m + nn));
    const float log_mu = (i < m) ? norm : logf((float)nn) + norm;
    u[(long long)p * ld + i] = log_mu - (mx + logf(s));
  }
}

// v_j = log_nu_j - LSE_i(Z_ij + u_i): 32 columns per block, 8 row-lanes
__global__ void sg_col_pass_kernel(const float* __restrict__ Z, const int* __restrict__ n, int cap, const float* __restrict__ u, float* __restrict__ v) {
  const int p = blockIdx.y;
  const int m = n[2 * p], nn = n[2 * p + 1];
  const int j = blockIdx.x * 32 + threadIdx.x;
  const int ld = cap + 1;
  __shared__ float sm[8][33], ss[8][33];
  const float* z = Z + (long long)p * ld * ld;
  const float* uu = u + (long long)p * ld;
  float mx = -INFINITY;
  if (j <= nn) for (int i = threadIdx.y; i <= m; i += 8) mx = fmaxf(mx, z[(long long)i * ld + j] + uu[i]);
  sm[threadIdx.y][threadIdx.x] = mx;
  __syncthreads();
  float mm = -INFINITY;
#pragma unroll
  for (int k = 0; k < 8; ++k) mm = fmaxf(mm, sm[k][threadIdx.x]);
  float s = 0.f;
  if (j <= nn) for (int i = threadIdx.y; i <= m; i += 8) s += expf(z[(long long)i * ld + j] + uu[i] - mm);
  ss[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && j <= nn) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += ss[k][threadIdx.x];
    const float norm = -logf((float)(m + nn));
    const float log_nu = (j < nn) ? norm : logf((float)m) + norm;
    v[(long long)p * ld + j] = log_nu - (mm + logf(t));
  }
}


// ---- K18 fused: all 2 x iters log-Sinkhorn passes of one pair in ONE kernel ------------------------------------------------------
// The coupling matrix Z never changes during the iterations -- only u and v do -- so a thread-block CLUSTER of 8 CTAs keeps the whole
// (M+1) x (N+1) fp32 matrix of a pair in REGISTERS (401^2 floats / (8 x 512 threads) = 52 per thread at 400 keypoints) and the 200
// passes exchange nothing but per-column partial log-sum-exps:
//   row i = l * 8 + rank is owned by CTA `rank`, local row l = warp * RPW + r by one warp; lane holds columns lane + 32 k (k < CPL);
//   row pass : warp-shuffle LSE over the row -> u_i (stays in registers: only this warp's rows need it);
//   col pass : per-thread (max, sum) over its rows -> across warps through shared memory -> across the 8 CTAs through DSMEM
//              (double-buffered partials, one cluster barrier per iteration) -> v (every CTA computes the full v redundantly).
// Replaces 200 kernel launches that re-read Z from L2 each time (sg_row_pass_kernel / sg_col_pass_kernel: 2.1 ms per 8 pairs).
// The kernel is instruction-issue bound (~10^3 instructions per thread and iteration), so the exponentials / logarithms of the inner loops
// are the 2-instruction MUFU forms (__expf / __logf: 2^-22 relative error, far inside the 1e-6 the log-sum-exps need).
namespace cg = cooperative_groups;
constexpr int kSkCluster = 8;

__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {     // (m, s) <- logsumexp-combine, -inf safe
  const float mm = fmaxf(m, m2);
  const float ms = (mm == -INFINITY) ? 0.f : mm;
  s = s * __expf(m - ms) + s2 * __expf(m2 - ms);
  m = mm;
}

template <int CPL, int RPW, int NW>
__global__ void __cluster_dims__(kSkCluster, 1, 1) __launch_bounds__(NW * 32, 1)
sg_sinkhorn_cluster_kernel(const float* __restrict__ Z, const int* __restrict__ n, int cap, int iters, float* __restrict__ u_out, float* __restrict__ v_out) {
  constexpr int COLS = 32 * CPL;
  extern __shared__ __align__(16) uint8_t sk_smem[];
  float2* part = reinterpret_cast<float2*>(sk_smem);            // [NW][COLS]   per-warp column partials (max, sum)
  float2* cta_part = part + NW * COLS;                          // [2][COLS]    this CTA's column partials, double buffered, read by the whole cluster
  float* v_s = reinterpret_cast<float*>(cta_part + 2 * COLS);   // [COLS]
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int p = blockIdx.x / kSkCluster;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = n[2 * p], N = n[2 * p + 1];
  const int ld = cap + 1;
  if (M < 1 || N < 1 || M + 1 > kSkCluster * NW * RPW || N + 1 > COLS) return;   // uniform over the cluster: nobody reaches a barrier
  const float norm = -logf((float)(M + N));
  const float* z0 = Z + (long long)p * ld * ld;
  float z[RPW][CPL];
  float u[RPW];
  int row[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    row[r] = (warp * RPW + r) * kSkCluster + rank;
    u[r] = 0.f;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int j = lane + 32 * k;
      z[r][k] = (row[r] <= M && j <= N) ? z0[(long long)row[r] * ld + j] : -INFINITY;
    }
  }
  float v[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) v[k] = 0.f;

  for (int it = 0; it < iters; ++it) {
    // ---- row pass: u_i = log_mu_i - LSE_j(Z_ij + v_j)
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      if (row[r] <= M) {                       // warp-uniform
        float t[CPL];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < CPL; ++k) { t[k] = z[r][k] + v[k]; mx = fmaxf(mx, t[k]); }
        mx = warp_max_(mx);
        float sm = 0.f;
#pragma unroll
        for (int k = 0; k < CPL; ++k) sm += __expf(t[k] - mx);
        sm = warp_sum_(sm);
        const float log_mu = (row[r] < M) ? norm : logf((float)N) + norm;
        u[r] = log_mu - (mx + __logf(sm));
      }
    }
    // ---- col pass: v_j = log_nu_j - LSE_i(Z_ij + u_i), partials: thread -> warp rows (already thread-local) -> CTA -> cluster
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      float mx = -INFINITY;
      float t[RPW];
#pragma unroll
      for (int r = 0; r < RPW; ++r) { t[r] = z[r][k] + u[r]; mx = fmaxf(mx, t[r]); }      // rows beyond M hold -inf
      const float ms = (mx == -INFINITY) ? 0.f : mx;
      float sm = 0.f;
#pragma unroll
      for (int r = 0; r < RPW; ++r) sm += __expf(t[r] - ms);
      part[warp * COLS + lane + 32 * k] = make_float2(mx, sm);
    }
    __syncthreads();
    float2* mine = cta_part + (it & 1) * COLS;
    for (int j = threadIdx.x; j < COLS; j += NW * 32) {
      float mx = -INFINITY;
#pragma unroll
      for (int w = 0; w < NW; ++w) mx = fmaxf(mx, part[w * COLS + j].x);
      const float ms = (mx == -INFINITY) ? 0.f : mx;
      float sm = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) { const float2 q = part[w * COLS + j]; sm += q.y * __expf(q.x - ms); }
      mine[j] = make_float2(mx, sm);
    }
    cluster.sync();                            // every CTA's partials of this iteration are visible (and `part` may be rewritten)
    for (int j = threadIdx.x; j < COLS; j += NW * 32) {
      float2 q[kSkCluster];
#pragma unroll
      for (int c = 0; c < kSkCluster; ++c) q[c] = *cluster.map_shared_rank(mine + j, c);      // DSMEM, fixed rank order: deterministic
      float mx = q[0].x;
#pragma unroll
      for (int c = 1; c < kSkCluster; ++c) mx = fmaxf(mx, q[c].x);
      const float ms = (mx == -INFINITY) ? 0.f : mx;
      float sm = 0.f;
#pragma unroll
      for (int c = 0; c < kSkCluster; ++c) sm += q[c].y * __expf(q[c].x - ms);
      const float log_nu = (j < N) ? norm : logf((float)M) + norm;
      v_s[j] = (j <= N) ? log_nu - (mx + __logf(sm)) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CPL; ++k) v[k] = v_s[lane + 32 * k];
  }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < RPW; ++r)
      if (row[r] <= M) u_out[(long long)p * ld + row[r]] = u[r];
  }
  if (rank == 0)
    for (int j = threadIdx.x; j <= N; j += NW * 32) v_out[(long long)p * ld + j] = v_s[j];
  cluster.sync();                              // no CTA may exit while others still read its shared memory
}

template <int CPL, int RPW, int NW>
static bool launch_sinkhorn_cluster(const float* Z, const int* n, int pairs, int cap, int iters, float* u, float* v, cudaStream_t st) {
  constexpr int smem = (NW * 32 * CPL + 2 * 32 * CPL) * (int)sizeof(float2) + 32 * CPL * (int)sizeof(float);
  static bool attr_set[kMaxDevices] = {};
  const int dev = current_device();
  auto kern = sg_sinkhorn_cluster_kernel<CPL, RPW, NW>;
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) { cudaGetLastError(); return false; }
    attr_set[dev] = true;
  }
  kern<<<pairs * kSkCluster, NW * 32, smem, st>>>(Z, n, cap, iters, u, v);
  return cudaGetLastError() == cudaSuccess;
}

__device__ __forceinline__ float sg_score(float z, float u, float v, float norm) { return __fsub_rn(__fadd_rn(__fadd_rn(z, u), v), norm); }

// decode: row / column argmax over the inner M x N block of Z + u + v - norm (strict '<' scan: first max)
__global__ void sg_rowmax_kernel(const float* __restrict__ Z, const float* __restrict__ u, const float* __restrict__ v, const int* __restrict__ n,
                                 int cap, int* __restrict__ arg0, float* __restrict__ val0, float* __restrict__ dense_out) {
  const int p = blockIdx.y;
  const int m = n[2 * p], nn = n[2 * p + 1];
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int ld = cap + 1;
  if (i > m) return;
  const float norm = -logf((float)(m + nn));
  const float* z = Z + ((long long)p * ld + i) * ld;
  const float ui = u[(long long)p * ld + i];
  const float* vv = v + (long long)p * ld;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  for (int j = lane; j <= nn; j += 32) {
    const float s = sg_score(z[j], ui, vv[j], norm);
    if (dense_out) dense_out[((long long)p * ld + i) * ld + j] = s;
    if (j < nn && i < m && s > best) { best = s; arg = j; }
  }
  if (i >= m) return;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  if (lane == 0) { arg0[(long long)p * cap + i] = arg == 0x7fffffff ? 0 : arg; val0[(long long)p * cap + i] = best; }
}

__global__ void sg_colmax_kernel(const float* __restrict__ Z, const float* __restrict__ u, const float* __restrict__ v, const int* __restrict__ n,
                                 int cap, int* __restrict__ arg1) {
  const int p = blockIdx.y;
  const int m = n[2 * p], nn = n[2 * p + 1];
  const int j = blockIdx.x * 32 + threadIdx.x;
  const int ld = cap + 1;
  __shared__ float sv[8][33];
  __shared__ int sa[8][33];
  const float norm = -logf((float)(m + nn));
  const float* z = Z + (long long)p * ld * ld;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  if (j < nn) {
    const float vj = v[(long long)p * ld + j];
    for (int i = threadIdx.y; i < m; i += 8) {
      const float s = sg_score(z[(long long)i * ld + j], u[(long long)p * ld + i], vj, norm);
      if (s > best) { best = s; arg = i; }
    }
  }
  sv[threadIdx.y][threadIdx.x] = best;
  sa[threadIdx.y][threadIdx.x] = arg;
  __syncthreads();
  if (threadIdx.y == 0 && j < nn) {
    for (int k = 1; k < 8; ++k) {
      const float ob = sv[k][threadIdx.x];
      const int oa = sa[k][threadIdx.x];
      if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    arg1[(long long)p * cap + j] = arg == 0x7fffffff ? 0 : arg;
  }
}

// decode() + the mutual filter of MatchingPoints; one CTA of 1024 threads per pair (cap <= 1024)
__global__ void __launch_bounds__(1024) sg_decode_kernel(const int* __restrict__ arg0, const float* __restrict__ val0, const int* __restrict__ arg1,
                                                         const int* __restrict__ n, int cap, float thr, int* __restrict__ idx0, int* __restrict__ idx1,
                                                         float* __restrict__ ms0, float* __restrict__ ms1, int* __restrict__ m_idx,
                                                         float* __restrict__ m_score, int* __restrict__ m_count) {
  const int p = blockIdx.x;
  const int t = threadIdx.x;
  const int m = n[2 * p], nn = n[2 * p + 1];
  __shared__ float s_ms0[1024];
  __shared__ unsigned char s_valid0[1024];
  __shared__ int s_i1[1024];
  const long long o = (long long)p * cap;
  // side 0
  float my_ms0 = 0.f;
  bool valid0 = false;
  int a0 = 0;
  if (t < m) {
    a0 = arg0[o + t];
    const bool mutual0 = (arg1[o + a0] == t);
    my_ms0 = mutual0 ? expf(val0[o + t]) : 0.f;
    valid0 = mutual0 && (my_ms0 > thr);
  }
  s_ms0[t] = my_ms0;
  s_valid0[t] = valid0;
  __syncthreads();
  if (t < m) { idx0[o + t] = valid0 ? a0 : -1; ms0[o + t] = my_ms0; }
  // side 1
  int my_i1 = -1;
  float my_ms1 = 0.f;
  if (t < nn) {
    const int a1 = arg1[o + t];
    const bool mutual1 = (arg0[o + a1] == t);
    my_ms1 = mutual1 ? s_ms0[a1] : 0.f;
    const bool valid1 = mutual1 && s_valid0[a1];
    my_i1 = valid1 ? a1 : -1;
    idx1[o + t] = my_i1;
    ms1[o + t] = my_ms1;
  }
  s_i1[t] = my_i1;
  __syncthreads();
  // PointMatcher: i with indices0[i] in range and indices1[indices0[i]] == i -> match (i, indices0[i]), score (ms0[i] + ms1[j]) / 2
  bool ok = false;
  int j = 0;
  if (t < m && valid0) {
    j = a0;
    ok = (j >= 0 && j < nn && s_i1[j] == t);
  }
  __shared__ int ws[32];
  const int lane = t & 31, warp = t >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, ok);
  if (lane == 0) ws[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    int w = ws[lane];
#pragma unroll
    for (int q = 1; q < 32; q <<= 1) { const int tt = __shfl_up_sync(0xffffffffu, w, q); if (lane >= q) w += tt; }
    ws[lane] = w;
  }
  __syncthreads();
  const int pos = (warp ? ws[warp - 1] : 0) + __popc(bal & ((1u << lane) - 1));
  if (ok) {
    m_idx[(o + pos) * 2] = t;
    m_idx[(o + pos) * 2 + 1] = j;
    // mscores1[j] for a mutual valid pair equals mscores0[i]; keep the reference's expression
    m_score[o + pos] = (float)(((double)my_ms0 + (double)s_ms0[t]) / 2.0);
  }
  if (t == 0) m_count[p] = ws[31];
}

void launch_sg_sinkhorn_decode(const float* sim, const int* n, int pairs, int cap, int max_n, float bin_score, int iters, float* Z, float* u, float* v,
                               float thr, int* arg0, float* val0, int* arg1, int* idx0, int* idx1, float* ms0, float* ms1, int* m_idx,
                               float* m_score, int* m_count, float* dense_out, cudaStream_t st) {
  sg_couplings_kernel<<<dim3((cap + 1 + 127) / 128, cap + 1, pairs), 128, 0, st>>>(sim, n, cap, bin_score, Z, u, v);
  // fused: one cluster kernel runs all 2 x iters passes with Z in registers.  Size class by slot capacity: <= 415 keypoints per side
  // (13 columns per lane) is what the configs of AirSLAM use (max_keypoints 350 .. 450 -> callers size cap accordingly); up to 512 takes the
  // 17-column instantiation; larger capacities (the 1024-keypoint profile) keep the pass-per-launch path.  AIRFE_SINKHORN_V1=1 forces it.
  static const bool v1 = getenv("AIRFE_SINKHORN_V1") != nullptr;
  bool fused = false;
  if (!v1 && cap <= 512) {
    fused = (max_n >= 0 && max_n <= 415) ? launch_sinkhorn_cluster<13, 4, 16>(Z, n, pairs, cap, iters, u, v, st)
                                         : launch_sinkhorn_cluster<17, 6, 12>(Z, n, pairs, cap, iters, u, v, st);
  }
  if (!fused)
    for (int it = 0; it < iters; ++it) {
      sg_row_pass_kernel<<<dim3((cap + 1 + 7) / 8, pairs), 256, 0, st>>>(Z, n, cap, u, v);
      sg_col_pass_kernel<<<dim3((cap + 1 + 31) / 32, pairs), dim3(32, 8), 0, st>>>(Z, n, cap, u, v);
    }
  sg_rowmax_kernel<<<dim3((cap + 1 + 7) / 8, pairs), 256, 0, st>>>(Z, u, v, n, cap, arg0, val0, dense_out);
  sg_colmax_kernel<<<dim3((cap + 31) / 32, pairs), dim3(32, 8), 0, st>>>(Z, u, v, n, cap, arg1);
  sg_decode_kernel<<<pairs, 1024, 0, st>>>(arg0, val0, arg1, n, cap, thr, idx0, idx1, ms0, ms1, m_idx, m_score, m_count);
}

}  // namespace airfe
