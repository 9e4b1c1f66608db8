// Hand-written sm_100a kernels for the detector's non-GEMM stages: resize, first conv, pool / upsample, detector softmax,
// NMS, keypoint selection, descriptor sampling.  All are HBM/L2-bandwidth kernels: coalesced, vectorised (16-byte
// accesses where the layout allows), warp-shuffle reductions, no tensor cores.
#include "kernels.h"
#include <math.h>
#include <stdlib.h>

namespace airfe {

// =====================================================================================================================
// K1 resize.  Bit-exact restatement of OpenCV's 8-bit INTER_LINEAR (fixed-point, 11-bit coefficients), SURVEY.md App. B9.
// The per-column / per-row coefficient tables are built on the host with the same float/double casts as OpenCV.
// =====================================================================================================================
void build_resize_tables_host(int src_w, int src_h, int* sx, int* a0, int* a1, int* sy, int* b0, int* b1) {
  const double scale_x = (double)src_w / 512.0, scale_y = (double)src_h / 512.0;
  for (int x = 0; x < 512; ++x) {
    float fx = (float)((x + 0.5) * scale_x - 0.5);
    int s = (int)floorf(fx);
    fx -= (float)s;
    if (s < 0) { s = 0; fx = 0.f; }
    if (s >= src_w - 1) { s = src_w - 1; fx = 0.f; }
    sx[x] = s;
    a0[x] = (int)rintf((1.f - fx) * 2048.f);
    a1[x] = (int)rintf(fx * 2048.f);
  }
  for (int y = 0; y < 512; ++y) {
    float fy = (float)((y + 0.5) * scale_y - 0.5);
    int s = (int)floorf(fy);
    fy -= (float)s;
    sy[y] = s;
    b0[y] = (int)rintf((1.f - fy) * 2048.f);
    b1[y] = (int)rintf(fy * 2048.f);
  }
}

// cv::remap, CV_8UC1, INTER_LINEAR, BORDER_CONSTANT(0): one rectified pixel (src/camera.cc:163).  Restated from OpenCV's remapBilinear with the
// FixedPtCast<int, uchar, 15> cast: bilinear weights (32 - fx)(32 - fy) * 32 etc. are exact in the 15-bit table, so no table is needed.
void build_remap_host(const float* map_x, const float* map_y, int w, int h, short* xy, unsigned short* a) {
  auto sat16 = [](int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); };
  for (long long i = 0; i < (long long)w * h; ++i) {
    const int sx = (int)lrintf(map_x[i] * 32.f), sy = (int)lrintf(map_y[i] * 32.f);      // cvRound: round half to even
    xy[2 * i] = sat16(sx >> 5);
    xy[2 * i + 1] = sat16(sy >> 5);
    a[i] = (unsigned short)(((sy & 31) << 5) | (sx & 31));
  }
}
__device__ __forceinline__ int remap_px(const uint8_t* __restrict__ img, int w, int h, int stride, const RemapSide m, int r, int c) {
  const long long i = (long long)r * w + c;
  const short2 xy = *reinterpret_cast<const short2*>(m.xy + 2 * i);
  const int a = m.a[i];
  const int sx = xy.x, sy = xy.y, fx = a & 31, fy = a >> 5;
  if (sx >= w || sx + 1 < 0 || sy >= h || sy + 1 < 0) return 0;
  const bool x0 = sx >= 0, x1 = sx + 1 < w, y0 = sy >= 0, y1 = sy + 1 < h;
  const uint8_t* p = img + (long long)sy * stride + sx;
  const int v00 = (x0 && y0) ? p[0] : 0, v01 = (x1 && y0) ? p[1] : 0, v10 = (x0 && y1) ? p[stride] : 0, v11 = (x1 && y1) ? p[stride + 1] : 0;
  const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
  return (v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11 + (1 << 14)) >> 15;
}

// u8 -> network input: the reference evaluates float(u8) / 255.0 in DOUBLE and stores a float (src/plnet.cpp:264); the fp16 operand rounding follows.
// A double division per pixel on a GPU whose FP64 rate is 1/64 of FP32 made this the whole cost of the kernel, so the 256 possible results are
// computed once per device with exactly that expression and looked up (bit-identical by construction).
__device__ __half g_u8_to_half[256];
__global__ void u8_lut_init_kernel() { g_u8_to_half[threadIdx.x] = __float2half_rn((float)((double)(int)threadIdx.x / 255.0)); }
static void ensure_u8_lut() {
  static bool done[kMaxDevices] = {};
  const int dev = current_device();
  if (done[dev]) return;
  u8_lut_init_kernel<<<1, 256>>>();          // default stream + device synchronisation: once per device, at the first (eager) resize
  cudaDeviceSynchronize();
  done[dev] = true;
}

template <bool REMAP>
__global__ void resize_kernel(const uint8_t* __restrict__ src, int src_w, int src_h, int src_stride, long long src_img_stride,
                              ResizeTables t, __half* __restrict__ dst, uint8_t* __restrict__ dst_u8, RemapMaps maps) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (x >= 512) return;
  const uint8_t* img = src + (long long)b * src_img_stride;
  const int s0 = t.sx[x], s1 = min(s0 + 1, src_w - 1);
  const int a0 = t.a0[x], a1 = t.a1[x];
  const int r0 = min(max(t.sy[y], 0), src_h - 1), r1 = min(max(t.sy[y] + 1, 0), src_h - 1);
  int p00, p01, p10, p11;
  if (REMAP) {       // the four pixels of the RECTIFIED image this output needs, computed from the raw frame on the fly
    const RemapSide m = maps.side[maps.mode == 2 ? (b & 1) : 0];
    p00 = remap_px(img, src_w, src_h, src_stride, m, r0, s0); p01 = remap_px(img, src_w, src_h, src_stride, m, r0, s1);
    p10 = remap_px(img, src_w, src_h, src_stride, m, r1, s0); p11 = remap_px(img, src_w, src_h, src_stride, m, r1, s1);
  } else {
    p00 = img[(long long)r0 * src_stride + s0]; p01 = img[(long long)r0 * src_stride + s1];
    p10 = img[(long long)r1 * src_stride + s0]; p11 = img[(long long)r1 * src_stride + s1];
  }
  const int h0 = p00 * a0 + p01 * a1;
  const int h1 = p10 * a0 + p11 * a1;
  int v = (((t.b0[y] * (h0 >> 4)) >> 16) + ((t.b1[y] * (h1 >> 4)) >> 16) + 2) >> 2;
  v = min(max(v, 0), 255);
  const long long o = ((long long)b * 512 + y) * 512 + x;
  dst[o] = g_u8_to_half[v];     // = half(float(double(v) / 255.0)), see u8_lut_init_kernel
  if (dst_u8) dst_u8[o] = (uint8_t)v;
}

void launch_resize_u8_to_f16(const uint8_t* src, int src_w, int src_h, int src_stride, long long src_img_stride, int batch,
                             ResizeTables t, __half* dst, uint8_t* dst_u8, cudaStream_t st, const RemapMaps* remap) {
  dim3 grid(512 / 128, 512, batch);
  ensure_u8_lut();
  if (remap && remap->mode) resize_kernel<true><<<grid, 128, 0, st>>>(src, src_w, src_h, src_stride, src_img_stride, t, dst, dst_u8, *remap);
  else resize_kernel<false><<<grid, 128, 0, st>>>(src, src_w, src_h, src_stride, src_img_stride, t, dst, dst_u8, RemapMaps{});
}

__global__ void remap_kernel(const uint8_t* __restrict__ src, int w, int h, int stride, long long img_stride, RemapMaps maps, uint8_t* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= w) return;
  const RemapSide m = maps.side[maps.mode == 2 ? (b & 1) : 0];
  dst[((long long)b * h + y) * w + x] = (uint8_t)remap_px(src + (long long)b * img_stride, w, h, stride, m, y, x);
}
void launch_remap_u8(const uint8_t* src, int w, int h, int stride, long long img_stride, int batch, RemapMaps maps, uint8_t* dst, cudaStream_t st) {
  remap_kernel<<<dim3((w + 127) / 128, h, batch), 128, 0, st>>>(src, w, h, stride, img_stride, maps, dst);
}

// =====================================================================================================================
// First layer: 3x3, 1 -> 64 channels.  K = 9 is no tensor-core shape and the layer is FP32-FMA bound (4.8 GFMA per 32 frames), so the
// inner product runs on Blackwell's packed FFMA2 (fma.rn.f32x2): one instruction = the same input pixel times the weights of two
// adjacent output channels.  That halves the issue slots and leaves the kernel bound by the FMA pipe rather than by instruction issue.
// Thread = (8 pixels along x, group of 8 output channels): consecutive threads write consecutive 16-byte vectors of one pixel.
// =====================================================================================================================
constexpr int kC1Rows = 8;   // image rows per block
__device__ __forceinline__ unsigned long long ffma2_(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long pack2_(float lo, float hi) {
  return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
// PX = pixels along x per thread.  PX = 8 needs 164 registers (three CTAs = 12 warps per SM).  A PX = 4 instantiation (128 registers, four CTAs =
// 16 warps per SM, the weights re-read from shared memory twice as often) was measured in round 2: 0.847 -> 1.088 ms per 94 frames -- the kernel
// is not short of warps, it is bound by the FMA pipe plus the shared-memory weight reads.  Only PX = 8 is instantiated.
template <int PX>
__global__ void __launch_bounds__(128, 3) conv1a_kernel(const __half* __restrict__ x, const __half* __restrict__ w, const float* __restrict__ bias,
                                                                      __half* __restrict__ out, int H, int W, long long total) {
  __shared__ float2 sw2[8 * 4 * 9];   // [channel group][channel pair][tap] = (w of channel 2p, w of channel 2p+1)
  __shared__ float2 sb2[32];
  for (int i = threadIdx.x; i < 288; i += blockDim.x) {
    const int k = i % 9, cp = i / 9;              // cp = channel pair 0..31
    sw2[i] = make_float2(__half2float(w[(2 * cp) * 9 + k]), __half2float(w[(2 * cp + 1) * 9 + k]));
  }
  if (threadIdx.x < 32) sb2[threadIdx.x] = make_float2(bias[2 * threadIdx.x], bias[2 * threadIdx.x + 1]);
  __syncthreads();
  // grid = (threads along a row / 128, H / kC1Rows, batch): no integer divisions on the index path.  A block walks kC1Rows consecutive image
  // rows: the weight staging above and the block launch are paid once per 8 rows instead of once per row (they were ~25 % of a one-row
  // block's life), and the three input rows slide through registers, so each input row is loaded once instead of three times.
  const int tix = blockIdx.x * blockDim.x + threadIdx.x;
  if (tix >= W / PX * 8) return;                   // W / PX pixel groups x 8 channel groups threads per row
  const int cg = tix & 7;
  const int x0 = (tix >> 3) * PX;
  const int y_begin = blockIdx.y * kC1Rows;
  const long long img = blockIdx.z;
  const __half* xi = x + img * (long long)W * H;
  // inputs: per row one aligned 2 * PX-byte vector (x0 is a multiple of PX) plus the two halo pixels; each value duplicated into a float2
  unsigned long long in2[3][PX + 2];
  auto load_row = [&](unsigned long long (&dst)[PX + 2], int y2) {
    const bool rv = (y2 >= 0) && (y2 < H);
    const __half* rp = xi + (long long)(rv ? y2 : 0) * W + x0;
    uint32_t v[PX / 2];
    if constexpr (PX == 8) { const uint4 t = *reinterpret_cast<const uint4*>(rp); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else { const uint2 t = *reinterpret_cast<const uint2*>(rp); v[0] = t.x; v[1] = t.y; }
    const __half hl = (x0 > 0) ? rp[-1] : __float2half(0.f);
    const __half hr = (x0 + PX < W) ? rp[PX] : __float2half(0.f);
    const float fl = rv ? __half2float(hl) : 0.f, fr = rv ? __half2float(hr) : 0.f;
    dst[0] = pack2_(fl, fl);
    dst[PX + 1] = pack2_(fr, fr);
#pragma unroll
    for (int q = 0; q < PX / 2; ++q) {
      const uint32_t vq = rv ? v[q] : 0u;
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&vq));
      dst[1 + 2 * q] = pack2_(f.x, f.x);
      dst[2 + 2 * q] = pack2_(f.y, f.y);
    }
  };
  load_row(in2[0], y_begin - 1);
  load_row(in2[1], y_begin);
#pragma unroll 1
  for (int yy = y_begin; yy < y_begin + kC1Rows && yy < H; ++yy) {
    load_row(in2[2], yy + 1);
    uint32_t packed[PX][4];
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      unsigned long long wp[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) { const float2 t = sw2[(cg * 4 + jp) * 9 + k]; wp[k] = pack2_(t.x, t.y); }
      const float2 bb = sb2[cg * 4 + jp];
      const unsigned long long b2 = pack2_(bb.x, bb.y);
#pragma unroll
      for (int px = 0; px < PX; ++px) {
        unsigned long long acc = b2;                 // bias folded into the accumulator
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) acc = ffma2_(in2[ky][px + kx], wp[ky * 3 + kx], acc);
        const float a0 = fmaxf(__uint_as_float((unsigned)(acc & 0xffffffffu)), 0.f), a1 = fmaxf(__uint_as_float((unsigned)(acc >> 32)), 0.f);
        __half2 h2 = __floats2half2_rn(a0, a1);
        packed[px][jp] = *reinterpret_cast<uint32_t*>(&h2);
      }
    }
    __half* o = out + ((img * H + yy) * (long long)W + x0) * 64 + cg * 8;
#pragma unroll
    for (int px = 0; px < PX; ++px)
      *reinterpret_cast<uint4*>(o + (long long)px * 64) = make_uint4(packed[px][0], packed[px][1], packed[px][2], packed[px][3]);
#pragma unroll
    for (int j = 0; j < PX + 2; ++j) { in2[0][j] = in2[1][j]; in2[1][j] = in2[2][j]; }
  }
}

void launch_conv1a(const __half* x, const __half* w, const float* bias, __half* out, int batch, int H, int W, cudaStream_t st) {
  conv1a_kernel<8><<<dim3((W + 127) / 128, (H + kC1Rows - 1) / kC1Rows, batch), 128, 0, st>>>(x, w, bias, out, H, W, 0);
}

// =====================================================================================================================
// 2x2 max-pool / nearest 2x upsample, NHWC fp16, 8 channels (16 bytes) per thread.
// =====================================================================================================================
__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
  uint4 r;
  __half2* ra = reinterpret_cast<__half2*>(&a);
  __half2* rb = reinterpret_cast<__half2*>(&b);
  __half2* rr = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) rr[i] = __hmax2(ra[i], rb[i]);
  return r;
}

__global__ void maxpool2_kernel(const __half* __restrict__ in, int C8, int H, int W, long long in_ps, __half* __restrict__ out,
                                long long out_ps, long long total) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int c = (int)(gid % C8);
  const long long p = gid / C8;
  const int Wo = W >> 1, Ho = H >> 1;
  const int xo = (int)(p % Wo);
  const int yo = (int)((p / Wo) % Ho);
  const long long b = p / ((long long)Wo * Ho);
  const __half* base = in + ((b * H + 2 * yo) * (long long)W + 2 * xo) * in_ps + c * 8;
  uint4 v00 = *reinterpret_cast<const uint4*>(base);
  uint4 v01 = *reinterpret_cast<const uint4*>(base + in_ps);
  uint4 v10 = *reinterpret_cast<const uint4*>(base + (long long)W * in_ps);
  uint4 v11 = *reinterpret_cast<const uint4*>(base + (long long)W * in_ps + in_ps);
  *reinterpret_cast<uint4*>(out + p * out_ps + c * 8) = hmax8(hmax8(v00, v01), hmax8(v10, v11));
}

void launch_maxpool2(const __half* in, int C, int H, int W, int batch, long long in_ps, __half* out, long long out_ps, cudaStream_t st) {
  const long long total = (long long)batch * (H / 2) * (W / 2) * (C / 8);
  maxpool2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, C / 8, H, W, in_ps, out, out_ps, total);
}

__global__ void upsample2_kernel(const __half* __restrict__ in, int C8, int H, int W, long long in_ps, __half* __restrict__ out,
                                 long long out_ps, long long total) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int c = (int)(gid % C8);
  const long long p = gid / C8;  // output pixel
  const int Wo = W * 2, Ho = H * 2;
  const int xo = (int)(p % Wo);
  const int yo = (int)((p / Wo) % Ho);
  const long long b = p / ((long long)Wo * Ho);
  const uint4 v = *reinterpret_cast<const uint4*>(in + ((b * H + (yo >> 1)) * (long long)W + (xo >> 1)) * in_ps + c * 8);
  *reinterpret_cast<uint4*>(out + p * out_ps + c * 8) = v;
}

void launch_upsample2(const __half* in, int C, int H, int W, int batch, long long in_ps, __half* out, long long out_ps, cudaStream_t st) {
  const long long total = (long long)batch * (H * 2) * (W * 2) * (C / 8);
  upsample2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, C / 8, H, W, in_ps, out, out_ps, total);
}

// =====================================================================================================================
// K4 detector head: softmax over 65 logits per 8x8 cell, drop the dustbin, depth-to-space into the 512x512 heat map.
// One warp per cell: lanes hold logits {l, l+32, (64 on lane 0)}; shuffle max / sum; lane l writes pixels l and l+32.
// =====================================================================================================================
__global__ void softmax_d2s_kernel(const float* __restrict__ logits, int ld, float* __restrict__ heat, int cells) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= cells) return;
  const float* l = logits + (long long)warp * ld;
  const float v0 = l[lane], v1 = l[lane + 32];
  const float v2 = (lane == 0) ? l[64] : -INFINITY;
  float m = fmaxf(fmaxf(v0, v1), v2);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float e0 = expf(v0 - m), e1 = expf(v1 - m), e2 = (lane == 0) ? expf(v2 - m) : 0.f;
  float s = e0 + e1 + e2;
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const int b = warp >> 12, cy = (warp >> 6) & 63, cx = warp & 63;
  float* h = heat + (long long)b * 262144;
  // channel c -> (dy = c / 8, dx = c % 8)
  h[(cy * 8 + (lane >> 3)) * 512 + cx * 8 + (lane & 7)] = e0 / s;
  h[(cy * 8 + 4 + (lane >> 3)) * 512 + cx * 8 + (lane & 7)] = e1 / s;
}

void launch_softmax_d2s(const float* logits, int ld, float* heat, int batch, cudaStream_t st) {
  const int cells = batch * 4096;
  softmax_d2s_kernel<<<(cells * 32 + 255) / 256, 256, 0, st>>>(logits, ld, heat, cells);
}

// =====================================================================================================================
// K5 simple_nms (radius 4).  App. B1:  M = (S == mp9(S));  2x { Sup = mp9(M) > 0; S' = Sup ? 0 : S;
//                                      M |= (S' == mp9(S')) & ~Sup };  out = M ? S : 0.   Padding acts as -inf.
// Round kernel: one CTA computes a 32x32 output tile from a (32+16)^2 input tile of (S, M) held in shared memory;
// 9x9 max = separable row-max then column-max.
// =====================================================================================================================
constexpr int NT = 32;

__global__ void nms_init_kernel(const float* __restrict__ heat, uint8_t* __restrict__ mask) {
  __shared__ float s[NT + 8][NT + 8];
  __shared__ float r[NT + 8][NT];
  const int b = blockIdx.z, x0 = blockIdx.x * NT, y0 = blockIdx.y * NT;
  const float* h = heat + (long long)b * 262144;
  for (int i = threadIdx.x; i < (NT + 8) * (NT + 8); i += blockDim.x) {
    const int yy = i / (NT + 8), xx = i % (NT + 8);
    const int gy = y0 + yy - 4, gx = x0 + xx - 4;
    s[yy][xx] = (gy >= 0 && gy < 512 && gx >= 0 && gx < 512) ? h[gy * 512 + gx] : -INFINITY;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (NT + 8) * NT; i += blockDim.x) {
    const int yy = i / NT, xx = i % NT;
    float m = s[yy][xx];
#pragma unroll
    for (int k = 1; k < 9; ++k) m = fmaxf(m, s[yy][xx + k]);
    r[yy][xx] = m;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NT * NT; i += blockDim.x) {
    const int yy = i / NT, xx = i % NT;
    float m = r[yy][xx];
#pragma unroll
    for (int k = 1; k < 9; ++k) m = fmaxf(m, r[yy + k][xx]);
    mask[(long long)b * 262144 + (y0 + yy) * 512 + x0 + xx] = (s[yy + 4][xx + 4] == m) ? 1 : 0;
  }
}

__global__ void nms_round_kernel(const float* __restrict__ heat, const uint8_t* __restrict__ mask_in, uint8_t* __restrict__ mask_out,
                                 float* __restrict__ scores_out /* non-null on the last round */) {
  constexpr int T16 = NT + 16, T8 = NT + 8;
  __shared__ float s[T16][T16];       // S on tile + 8 halo (-inf outside the image)
  __shared__ uint8_t mk[T16][T16];    // M on tile + 8 halo (0 outside)
  __shared__ float t1[T16][T8];       // row pass scratch
  __shared__ float sp[T8][T8];        // S' on tile + 4 halo (-inf outside the image)
  __shared__ uint8_t sup[T8][T8];
  const int b = blockIdx.z, x0 = blockIdx.x * NT, y0 = blockIdx.y * NT;
  const float* h = heat + (long long)b * 262144;
  const uint8_t* mi = mask_in + (long long)b * 262144;
  for (int i = threadIdx.x; i < T16 * T16; i += blockDim.x) {
    const int yy = i / T16, xx = i % T16;
    const int gy = y0 + yy - 8, gx = x0 + xx - 8;
    const bool in = (gy >= 0 && gy < 512 && gx >= 0 && gx < 512);
    s[yy][xx] = in ? h[gy * 512 + gx] : -INFINITY;
    mk[yy][xx] = in ? mi[gy * 512 + gx] : 0;
  }
  __syncthreads();
  // Sup = mp9(float(M)) > 0 on tile+4: separable OR
  for (int i = threadIdx.x; i < T16 * T8; i += blockDim.x) {
    const int yy = i / T8, xx = i % T8;
    int a = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) a |= mk[yy][xx + k];
    t1[yy][xx] = (float)a;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T8 * T8; i += blockDim.x) {
    const int yy = i / T8, xx = i % T8;
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) a = fmaxf(a, t1[yy + k][xx]);
    const int gy = y0 + yy - 4, gx = x0 + xx - 4;
    const bool in = (gy >= 0 && gy < 512 && gx >= 0 && gx < 512);
    const bool su = a > 0.f;
    sup[yy][xx] = su;
    sp[yy][xx] = in ? (su ? 0.f : s[yy + 4][xx + 4]) : -INFINITY;
  }
  __syncthreads();
  // mp9(S') on the tile
  float (*t2)[T8] = t1;  // reuse: rows 0..T8-1, cols 0..NT-1
  for (int i = threadIdx.x; i < T8 * NT; i += blockDim.x) {
    const int yy = i / NT, xx = i % NT;
    float m = sp[yy][xx];
#pragma unroll
    for (int k = 1; k < 9; ++k) m = fmaxf(m, sp[yy][xx + k]);
    t2[yy][xx] = m;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NT * NT; i += blockDim.x) {
    const int yy = i / NT, xx = i % NT;
    float m = t2[yy][xx];
#pragma unroll
    for (int k = 1; k < 9; ++k) m = fmaxf(m, t2[yy + k][xx]);
    const bool su = sup[yy + 4][xx + 4];
    const bool newm = mk[yy + 8][xx + 8] | ((sp[yy + 4][xx + 4] == m) && !su);
    const long long o = (long long)b * 262144 + (y0 + yy) * 512 + x0 + xx;
    mask_out[o] = newm;
    if (scores_out) scores_out[o] = newm ? s[yy + 8][xx + 8] : 0.f;
  }
}


// ---- K5 fused: the whole simple_nms (initial 9x9 maximum test + two suppression rounds = five max-pool passes in the graph) in ONE kernel ----
// A CTA produces a 64 x 64 output tile from the heat map on tile + 20 pixels (each of the five 9x9 pools eats 4 pixels of halo) held in
// shared memory; every 9-tap max / OR is separable and register-blocked (a thread computes 8 consecutive outputs from 16 inputs with a
// log-step window: 5.6 ops and 2 shared-memory reads per output instead of 9 and 9).  Exactly the arithmetic of the three-kernel version
// (float equality on the unmodified heat values, -inf / 0 outside the image), one read of the heat map and one write of the scores instead
// of three passes over (S, M) -- 0.50 -> see profiles/ for the measured time per 64 frames.
constexpr int kNmsT = 64;                  // output tile
constexpr int kNmsR = 20;                  // halo
constexpr int kNmsW = kNmsT + 2 * kNmsR;   // 104
constexpr int kNmsP = kNmsW + 1;           // float pitch 105 (odd: lanes that walk down a column of rows hit distinct banks)
constexpr int kNmsThreads = 256;

// out[i] = max(in[i .. i+8]) for i = 0..7 from 16 inputs
__device__ __forceinline__ void win9_max(const float (&in)[16], float (&out)[8]) {
  float a[15], b[13], c[9];
#pragma unroll
  for (int i = 0; i < 15; ++i) a[i] = fmaxf(in[i], in[i + 1]);
#pragma unroll
  for (int i = 0; i < 13; ++i) b[i] = fmaxf(a[i], a[i + 2]);
#pragma unroll
  for (int i = 0; i < 9; ++i) c[i] = fmaxf(b[i], b[i + 4]);
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = fmaxf(c[i], in[i + 8]);
}
__device__ __forceinline__ void win9_or(const uint32_t (&in)[16], uint32_t (&out)[8]) {
  uint32_t a[15], b[13], c[9];
#pragma unroll
  for (int i = 0; i < 15; ++i) a[i] = in[i] | in[i + 1];
#pragma unroll
  for (int i = 0; i < 13; ++i) b[i] = a[i] | a[i + 2];
#pragma unroll
  for (int i = 0; i < 9; ++i) c[i] = b[i] | b[i + 4];
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = c[i] | in[i + 8];
}

// Horizontal 9-max of src (rows [y0, y0 + nrows), output columns [x0, x0 + ncols), ncols % 8 == 0; window = columns x .. x+8 of src) -> dst.
// Work item = (row, block of 8 columns); lanes walk rows.  src / dst are region-relative with pitch kNmsP.
template <class LoadF>
__device__ __forceinline__ void row_pass_max(LoadF load, float* dst, int y0, int nrows, int x0, int ncols) {
  const int nblk = ncols >> 3;
  for (int it = threadIdx.x; it < nrows * nblk; it += kNmsThreads) {
    const int r = y0 + it % nrows, c = x0 + (it / nrows) * 8;
    float in[16], out[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) in[i] = load(r, c + i);
    win9_max(in, out);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[r * kNmsP + c + i] = out[i];
  }
}

__global__ void __launch_bounds__(kNmsThreads, 2) nms_fused_kernel(const float* __restrict__ heat, float* __restrict__ scores) {
  extern __shared__ __align__(16) uint8_t nms_smem[];
  float* S = reinterpret_cast<float*>(nms_smem);              // [104][105]  heat on tile + 20 (-inf outside the image)
  float* T = S + kNmsW * kNmsP;                               // [104][105]  row-pass scratch
  uint8_t* M = reinterpret_cast<uint8_t*>(T + kNmsW * kNmsP); // [104][104]  running maximum mask
  uint8_t* U = M + kNmsW * kNmsW;                             // [104][104]  suppression mask of the current round
  uint8_t* B = reinterpret_cast<uint8_t*>(T);                 // [104][104]  row-pass scratch (bytes): never live at the same time as T
  const int b = blockIdx.z, gx0 = blockIdx.x * kNmsT - kNmsR, gy0 = blockIdx.y * kNmsT - kNmsR;
  const float* h = heat + (long long)b * 262144;
  for (int i = threadIdx.x; i < kNmsW * kNmsW; i += kNmsThreads) {
    const int y = i / kNmsW, x = i - y * kNmsW;
    const int gy = gy0 + y, gx = gx0 + x;
    S[y * kNmsP + x] = (gy >= 0 && gy < 512 && gx >= 0 && gx < 512) ? h[gy * 512 + gx] : -INFINITY;
  }
  __syncthreads();
  auto in_img = [&](int y, int x) { const int gy = gy0 + y, gx = gx0 + x; return gy >= 0 && gy < 512 && gx >= 0 && gx < 512; };

  // ---- M0 = (S == mp9(S)) on region [4, 100)^2 : rows 0..103 x output columns 4..99 (window starts at column x - 4)
  row_pass_max([&](int r, int c) { return S[r * kNmsP + c]; }, T, 0, kNmsW, 0, 96);          // T[r][c] = max S[r][c .. c+8]  (centre c + 4)
  __syncthreads();
  for (int it = threadIdx.x; it < 96 * 12; it += kNmsThreads) {                              // column pass: lanes walk columns
    const int c = it % 96, r0 = (it / 96) * 8;                                               // output rows r0+4 .. r0+11, centre column c + 4
    float in[16], out[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) in[i] = T[(r0 + i) * kNmsP + c];
    win9_max(in, out);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int y = r0 + 4 + i, x = c + 4;
      M[y * kNmsW + x] = (in_img(y, x) && S[y * kNmsP + x] == out[i]) ? 1 : 0;               // outside the image the mask is 0
    }
  }
  __syncthreads();

  // ---- two suppression rounds; round k works on the region shrunk by 8 more pixels
#pragma unroll 1
  for (int round = 0; round < 2; ++round) {
    const int m_lo = 4 + 8 * round;                  // M is valid on [m_lo, 104 - m_lo)
    const int u_lo = m_lo + 4, u_n = kNmsW - 2 * u_lo;       // Sup on [u_lo, 104 - u_lo): 88 / 72 wide
    const int o_lo = u_lo + 4, o_n = kNmsW - 2 * o_lo;       // new mask on [o_lo, 104 - o_lo): 80 / 64 wide
    // Sup = mp9(M) > 0 : horizontal OR (rows m_lo .. , output columns u_lo ..), then vertical OR
    {
      const int nrows = kNmsW - 2 * m_lo, nblk = u_n >> 3;
      for (int it = threadIdx.x; it < nrows * nblk; it += kNmsThreads) {
        const int r = m_lo + it % nrows, c = u_lo + (it / nrows) * 8;
        uint32_t in[16], out[8];
#pragma unroll
        for (int i = 0; i < 16; ++i) in[i] = M[r * kNmsW + c - 4 + i];
        win9_or(in, out);
#pragma unroll
        for (int i = 0; i < 8; ++i) B[r * kNmsW + c + i] = (uint8_t)out[i];
      }
    }
    __syncthreads();
    for (int it = threadIdx.x; it < u_n * (u_n >> 3); it += kNmsThreads) {
      const int c = u_lo + it % u_n, r0 = u_lo + (it / u_n) * 8;
      uint32_t in[16], out[8];
#pragma unroll
      for (int i = 0; i < 16; ++i) in[i] = B[(r0 - 4 + i) * kNmsW + c];
      win9_or(in, out);
#pragma unroll
      for (int i = 0; i < 8; ++i) U[(r0 + i) * kNmsW + c] = (uint8_t)out[i];
    }
    __syncthreads();
    // S' = Sup ? 0 : S (-inf outside the image); mp9(S') on [o_lo, ..): horizontal then vertical max; M |= (S' == max) & ~Sup
    auto sprime = [&](int r, int c) { const float v = S[r * kNmsP + c]; return (v == -INFINITY) ? v : (U[r * kNmsW + c] ? 0.f : v); };
    {
      const int nblk = o_n >> 3;
      for (int it = threadIdx.x; it < u_n * nblk; it += kNmsThreads) {
        const int r = u_lo + it % u_n, c = o_lo + (it / u_n) * 8;
        float in[16], out[8];
#pragma unroll
        for (int i = 0; i < 16; ++i) in[i] = sprime(r, c - 4 + i);
        win9_max(in, out);
#pragma unroll
        for (int i = 0; i < 8; ++i) T[r * kNmsP + c + i] = out[i];
      }
    }
    __syncthreads();
    for (int it = threadIdx.x; it < o_n * (o_n >> 3); it += kNmsThreads) {
      const int c = o_lo + it % o_n, r0 = o_lo + (it / o_n) * 8;
      float in[16], out[8];
#pragma unroll
      for (int i = 0; i < 16; ++i) in[i] = T[(r0 - 4 + i) * kNmsP + c];
      win9_max(in, out);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int y = r0 + i;
        const bool su = U[y * kNmsW + c] != 0;
        const bool nm = M[y * kNmsW + c] | ((sprime(y, c) == out[i]) && !su && in_img(y, c));
        M[y * kNmsW + c] = nm;
        if (round == 1) scores[(long long)b * 262144 + (gy0 + y) * 512 + gx0 + c] = nm ? S[y * kNmsP + c] : 0.f;
      }
    }
    __syncthreads();
  }
}

void launch_simple_nms(const float* heat, float* scores, uint8_t* mask_a, uint8_t* mask_b, int batch, cudaStream_t st) {
  static const bool v1 = getenv("AIRFE_NMS_V1") != nullptr;       // the three-kernel version, kept for A/B timing and as a cross-check
  if (!v1) {
    constexpr int smem = 2 * kNmsW * kNmsP * 4 + 2 * kNmsW * kNmsW;      // 109 KB: two CTAs per SM
    static bool attr_set[kMaxDevices] = {};
    const int dev = current_device();
    if (!attr_set[dev]) { cudaFuncSetAttribute(nms_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr_set[dev] = true; }
    nms_fused_kernel<<<dim3(512 / kNmsT, 512 / kNmsT, batch), kNmsThreads, smem, st>>>(heat, scores);
    return;
  }
  dim3 grid(512 / NT, 512 / NT, batch);
  nms_init_kernel<<<grid, 256, 0, st>>>(heat, mask_a);
  nms_round_kernel<<<grid, 256, 0, st>>>(heat, mask_a, mask_b, nullptr);
  nms_round_kernel<<<grid, 256, 0, st>>>(heat, mask_b, mask_a, scores);
}

// =====================================================================================================================
// K6 keypoint selection (detect_point, src/plnet.cpp:309-355): score >= thr, border <= x <= W-border (inclusive), same for y;
// if more than top_k survive: top_k by (score desc, raster index asc) -- the tie order this build defines -- else raster order.
// Stage 1: unordered compaction with one atomic per warp.  Stage 2: O(n^2) rank (n <= ~10k after radius-4 NMS): each
// candidate counts the candidates that precede it in BOTH orders; no sort, deterministic output positions.
// =====================================================================================================================
__global__ void kp_compact_kernel(const float* __restrict__ scores, float thr, int border, int* __restrict__ cand, int cap,
                                  int* __restrict__ count) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 0..262143
  const float s = scores[(long long)b * 262144 + i];
  const int y = i >> 9, x = i & 511;
  const bool ok = (s >= thr) && !(x < border || x > 512 - border || y < border || y > 512 - border);
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if (!m) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(count + b, __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (ok) {
    const int pos = base + __popc(m & ((1u << lane) - 1));
    if (pos < cap) cand[(long long)b * cap + pos] = i;
  }
}

__global__ void kp_rank_kernel(const float* __restrict__ scores, const int* __restrict__ cand, int cap, const int* __restrict__ count,
                               int top_k, float* __restrict__ kp_out, int kp_cap, int* __restrict__ kp_count) {
  const int b = blockIdx.y;
  const int n = min(count[b], cap);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ int s_idx[256];
  __shared__ float s_val[256];
  const float* sc = scores + (long long)b * 262144;
  const int* cd = cand + (long long)b * cap;
  int my_idx = 0;
  float my_val = 0.f;
  if (i < n) { my_idx = cd[i]; my_val = sc[my_idx]; }
  int r_score = 0, r_idx = 0;
  for (int base = 0; base < n; base += 256) {
    const int j = base + threadIdx.x;
    __syncthreads();
    if (j < n) { const int id = cd[j]; s_idx[threadIdx.x] = id; s_val[threadIdx.x] = sc[id]; }
    __syncthreads();
    const int lim = min(256, n - base);
    if (i < n) {
      for (int k = 0; k < lim; ++k) {
        const float v = s_val[k];
        const int id = s_idx[k];
        r_score += (v > my_val) || (v == my_val && id < my_idx);
        r_idx += (id < my_idx);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) kp_count[b] = min(n, min(top_k, kp_cap));
  if (i < n) {
    const int pos = (n > top_k) ? r_score : r_idx;
    if (pos < top_k && pos < kp_cap) {
      float* o = kp_out + ((long long)b * kp_cap + pos) * 3;
      o[0] = my_val;
      o[1] = (float)(my_idx & 511);
      o[2] = (float)(my_idx >> 9);
    }
  }
}

void launch_select_keypoints(const float* scores, int batch, float threshold, int border, int top_k, int* cand_idx, int cand_cap,
                             int* cand_count, float* kp_out, int kp_cap, int* kp_count, cudaStream_t st) {
  cudaMemsetAsync(cand_count, 0, sizeof(int) * batch, st);
  kp_compact_kernel<<<dim3(262144 / 256, batch), 256, 0, st>>>(scores, threshold, border, cand_idx, cand_cap, cand_count);
  kp_rank_kernel<<<dim3((cand_cap + 255) / 256, batch), 256, 0, st>>>(scores, cand_idx, cand_cap, cand_count, top_k, kp_out, kp_cap,
                                                                      kp_count);
}

// =====================================================================================================================
// K7+K8 descriptor sampling (extract_descriptors, src/plnet.cpp:369-417).  One warp per keypoint, 8 channels per lane.
// The dense map is L2-normalised per pixel first (graph: D / clip(||D||, 1e-12)), which is folded into the gather: the
// four neighbours' norms are warp-reduced on the fly, so the normalised dense map is never written to HBM.
// =====================================================================================================================
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void sample_desc_kernel(const float* __restrict__ desc_raw, const float* __restrict__ kp, const int* __restrict__ kp_count,
                                   int kp_cap, float w_scale, float h_scale, float* __restrict__ feat) {
  const int b = blockIdx.y;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= kp_count[b]) return;
  const float* k = kp + ((long long)b * kp_cap + j) * 3;
  const float score = k[0], px = k[1], py = k[2];
  // constants for s = 8, w = h = 64: note `s / 2` is integer division in the reference
  const float den = (float)(64 * 8 - 8 / 2 - 0.5);
  const float sx = 2.f / den;
  const float bx = (float)((1 - 8) / (64 * 8 - 8 / 2 - 0.5) - 1);
  const float nx = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, sx), bx), 1.f), 0.5f);
  const float ny = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(py, sx), bx), 1.f), 0.5f);
  const float ix = __fmul_rn(nx, 63.f), iy = __fmul_rn(ny, 63.f);
  auto clip = [](int v) { return v < 0 ? 0 : (v > 63 ? 63 : v); };
  const int x_nw = clip((int)floorf(ix)), y_nw = clip((int)floorf(iy));
  const int x_ne = clip(x_nw + 1), y_ne = y_nw;
  const int x_sw = x_nw, y_sw = clip(y_nw + 1);
  const int x_se = clip(x_nw + 1), y_se = clip(y_nw + 1);
  const float w_nw = __fmul_rn(__fsub_rn((float)x_se, ix), __fsub_rn((float)y_se, iy));
  const float w_ne = __fmul_rn(__fsub_rn(ix, (float)x_sw), __fsub_rn((float)y_sw, iy));
  const float w_sw = __fmul_rn(__fsub_rn((float)x_ne, ix), __fsub_rn(iy, (float)y_ne));
  const float w_se = __fmul_rn(__fsub_rn(ix, (float)x_nw), __fsub_rn(iy, (float)y_nw));
  const float* base = desc_raw + (long long)b * 4096 * 256;
  const float* p_nw = base + (y_nw * 64 + x_nw) * 256 + lane * 8;
  const float* p_ne = base + (y_ne * 64 + x_ne) * 256 + lane * 8;
  const float* p_sw = base + (y_sw * 64 + x_sw) * 256 + lane * 8;
  const float* p_se = base + (y_se * 64 + x_se) * 256 + lane * 8;
  float v[4][8];
  float nrm[4] = {0.f, 0.f, 0.f, 0.f};
  const float* ps[4] = {p_nw, p_ne, p_sw, p_se};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(ps[q]);
    const float4 c = *reinterpret_cast<const float4*>(ps[q] + 4);
    v[q][0] = a.x; v[q][1] = a.y; v[q][2] = a.z; v[q][3] = a.w; v[q][4] = c.x; v[q][5] = c.y; v[q][6] = c.z; v[q][7] = c.w;
#pragma unroll
    for (int e = 0; e < 8; ++e) nrm[q] = fmaf(v[q][e], v[q][e], nrm[q]);
    nrm[q] = fmaxf(sqrtf(warp_sum(nrm[q])), 1e-12f);
  }
  float o[8];
  float ss = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float acc = __fmul_rn(__fdiv_rn(v[0][e], nrm[0]), w_nw);
    acc = __fadd_rn(acc, __fmul_rn(__fdiv_rn(v[1][e], nrm[1]), w_ne));
    acc = __fadd_rn(acc, __fmul_rn(__fdiv_rn(v[2][e], nrm[2]), w_sw));
    acc = __fadd_rn(acc, __fmul_rn(__fdiv_rn(v[3][e], nrm[3]), w_se));
    o[e] = acc;
    ss = fmaf(acc, acc, ss);
  }
  ss = warp_sum(ss);
  const float inv = ss > 0.f ? 1.f / sqrtf(ss) : 1.f;   // Eigen normalize(): all-zero column stays zero
  float* f = feat + ((long long)b * kp_cap + j) * 259;
  if (lane == 0) { f[0] = score; f[1] = __fmul_rn(px, w_scale); f[2] = __fmul_rn(py, h_scale); }
#pragma unroll
  for (int e = 0; e < 8; ++e) f[3 + lane * 8 + e] = ss > 0.f ? __fmul_rn(o[e], inv) : o[e];
}

void launch_sample_descriptors(const float* desc_raw, const float* kp, const int* kp_count, int kp_cap, int batch, float w_scale,
                               float h_scale, float* feat_out, cudaStream_t st) {
  dim3 grid((kp_cap * 32 + 255) / 256, batch);
  sample_desc_kernel<<<grid, 256, 0, st>>>(desc_raw, kp, kp_count, kp_cap, w_scale, h_scale, feat_out);
}

}  // namespace airfe
