// Launchers of the hand-written non-tensor-core kernels (HBM / shared-memory bound family, SURVEY.md §7.2 "H" rows).
// All pointers are device pointers; all launches are asynchronous on `st`.
#pragma once
#include "common.h"

namespace airfe {

// ---- K1: cv::resize(INTER_LINEAR, u8) -> 512x512 -> /255 -> fp16   (src/plnet.cpp:246-270, src/super_point.cpp:111-165)
struct ResizeTables {        // device arrays of 512 entries each, built on the host by build_resize_tables()
  int* sx; int* a0; int* a1; int* sy; int* b0; int* b1;
};
// ---- §8f rank 1: cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) of Camera::UndistortImage (src/camera.cc:161-182) fused into K1.
// Fixed-point form OpenCV derives from the CV_32F maps of initUndistortRectifyMap: xy = (cvRound(32 x) >> 5, cvRound(32 y) >> 5) saturated to
// int16, a = (fy << 5) | fx with the 5 fractional bits; built on the host by build_remap_host() exactly as OpenCV's remap does.
struct RemapSide { const short* xy; const unsigned short* a; };     // device: [h][w][2], [h][w]
struct RemapMaps {
  RemapSide side[2];     // 0 = left / mono camera, 1 = right camera
  int mode;              // 0 off, 1 = every image uses side 0, 2 = image index parity selects the side (stereo batches: L, R, L, R ...)
};
void build_remap_host(const float* map_x, const float* map_y, int w, int h, short* xy, unsigned short* a);
// standalone rectification (callers that want image_left_rect itself): dst [batch][h][w] u8
void launch_remap_u8(const uint8_t* src, int w, int h, int stride, long long img_stride, int batch, RemapMaps maps, uint8_t* dst, cudaStream_t st);

void build_resize_tables_host(int src_w, int src_h, int* sx, int* a0, int* a1, int* sy, int* b0, int* b1);
void launch_resize_u8_to_f16(const uint8_t* src, int src_w, int src_h, int src_stride, long long src_img_stride, int batch,
                             ResizeTables t, __half* dst /*[B,512,512]*/, uint8_t* dst_u8 /*optional [B,512,512]*/, cudaStream_t st,
                             const RemapMaps* remap = nullptr /* non-null and mode != 0: src holds RAW frames, rectified on the fly */);

// ---- first layer: 3x3 conv, C_in = 1 -> 64, +bias, ReLU, NHWC fp16 out (no tensor-core shape: K = 9)
void launch_conv1a(const __half* x /*[B,512,512]*/, const __half* w /*[64][9]*/, const float* bias, __half* out /*[B,512,512,64]*/,
                   int batch, int H, int W, cudaStream_t st);

// ---- 2x2/2 max-pool and nearest 2x upsample on NHWC fp16 (channel-strided in/out so concat buffers need no copy)
void launch_maxpool2(const __half* in, int C, int H, int W, int batch, long long in_pix_stride, __half* out, long long out_pix_stride,
                     cudaStream_t st);
void launch_upsample2(const __half* in, int C, int H, int W, int batch, long long in_pix_stride, __half* out, long long out_pix_stride,
                      cudaStream_t st);

// ---- K4: 65-way softmax, drop dustbin, 8x8 depth-to-space -> heat [B,512,512] fp32
void launch_softmax_d2s(const float* logits /*[B,64,64,ld]*/, int ld, float* heat, int batch, cudaStream_t st);

// ---- K5: simple_nms radius 4, two refinement rounds, exact float-equality semantics
void launch_simple_nms(const float* heat, float* scores, uint8_t* mask_a, uint8_t* mask_b, int batch, cudaStream_t st);

// ---- K6: threshold + border + top-k.  Output order: if count > top_k then (score desc, raster asc) else raster.
//      kp_out [B][cap][3] (score,x,y) fp32 ; kp_count [B]
void launch_select_keypoints(const float* scores, int batch, float threshold, int border, int top_k,
                             int* cand_idx /*[B][cand_cap]*/, int cand_cap, int* cand_count /*[B]*/, float* kp_out, int kp_cap,
                             int* kp_count, cudaStream_t st);

// ---- K7+K8: per-keypoint bilinear sampling of the (per-pixel L2-normalised) dense descriptor + renormalise.
//      desc_raw [B,64,64,256] fp32 (un-normalised convDb output).  Writes feat [B][cap][259] (score,x*sx,y*sy,desc) fp32.
void launch_sample_descriptors(const float* desc_raw, const float* kp, const int* kp_count, int kp_cap, int batch, float w_scale,
                               float h_scale, float* feat_out, cudaStream_t st);

}  // namespace airfe
