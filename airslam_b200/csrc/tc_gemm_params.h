// Parameter block of the tcgen05 implicit-GEMM kernel (see tc_gemm.cuh for the kernel itself).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace airfe {

struct TcGemmParams {
  CUtensorMap tmA;  // 4-D (C, W, H, B)   box (64, tw, th, tb)   SWIZZLE_128B
  CUtensorMap tmB;  // 4-D (K, N, Hb, Bb) box (64, block_n, 1, 1) SWIZZLE_128B  [MN-major: (N, K, Hb, Bb), box (64, 64, 1, 1)]
  CUtensorMap tmO;  // tma_store only: 5-D (columns of a section, W, H, B, sections) of the fp16 output, box (64, bw, bh, bb, 1) = one warp's 32 rows x 64 columns
  int taps;         // 1 or 9
  int kblocks;      // 64-wide K blocks per tap
  int c_in_pad;     // kblocks * 64 (K offset between taps in Bw)
  int tw, th, tb;   // tile box, tw*th*tb == 128
  int tiles_x, tiles_y, tiles_b;
  int n_tiles, block_n;
  int W, H, B;      // valid output extents
  int b_batched;    // Bw coordinates (Hb, Bb) follow the tile's (y, batch) index: per-head / per-image operand
  int b_batch_xor;  // Bb = tile batch ^ b_batch_xor (cross attention reads the partner image)
  int b_mn_major;   // Bw tile is MN-major (block_n must be 64)
  const float* bias;
  float scale;      // (acc + bias) * scale for columns < scale_cols (the 64^-1/4 / 256^-1/4 factors of the matchers)
  int scale_cols;
  const float* rot;    // optional rotary table [B*W rows][64] = (cos, sin) per pair: columns < rot_cols are rotated pairwise (LightGlue q, k; H == 1 only)
  int rot_cols;
  int out_split;       // > 0: column n is stored at out[(n / out_split) * out_split_stride + row offset + n % out_split]  (q | k | v go to separate matrices)
  long long out_split_stride;
  const float* resid;  // optional fp32 residual, same indexing as `out` (which must be fp32): out = resid + acc + bias
  void* out2;       // optional second store of the same values as fp16 (operand copy for the next GEMM)
  long long out2_sb, out2_sy, out2_sx;
  int relu;
  int out_f32;
  void* out;        // element (b,y,x,n) at out[b*out_sb + y*out_sy + x*out_sx + n]  (ch offset folded into `out`)
  long long out_sb, out_sy, out_sx;
  int n_valid;
  int stages;
  int prewait;      // MMA issuer polls the barriers of unit u + 1 before it issues unit u (see tc_gemm.cuh); 0 with AIRFE_PREWAIT=1 switches it on
  int tma_store;    // epilogue stages fp16 tiles in shared memory and stores them with cp.async.bulk.tensor (plain fp16 outputs with block_n % 128 == 0)
  int wide;         // host only: launch the EW = 4 instantiation (16 epilogue warps)
  int b_resident;   // whole [block_n x K] weight panel of this CTA's (fixed) N tile stays in shared memory; the ring then holds A only
  int dyn_w_stride; // index = tile batch * dyn_w_stride
  const int* dyn_w;  // optional: per-batch-index valid W (rows of a plain GEMM), read from device memory (tb must be 1)
};

constexpr int kTcThreads = 320;   // warp0 TMA, warp1 MMA, warps 2-9 epilogue (two per TMEM lane quarter, splitting the columns)
constexpr int kTcWideThreads = 576;   // EW = 4: warps 2-17 epilogue
constexpr int kTileM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kTileM * kBlockK * 2;  // 16 KiB
constexpr int kTcStoreStage = 8 * 4096;         // tma_store: one 32-row x 64-column fp16 staging tile (4 KiB, SWIZZLE_128B) per epilogue warp

__host__ __device__ inline int tc_b_bytes(int block_n, int mn_major) { return mn_major ? 64 * 128 : ((block_n * 128 + 1023) / 1024) * 1024; }
__host__ __device__ inline int tc_acc_stride(int block_n) { return (block_n + 31) / 32 * 32; }
__host__ __device__ inline int tc_tmem_cols(int block_n) {
  int need = 2 * tc_acc_stride(block_n), c = 32;
  while (c < need) c <<= 1;
  return c;
}

}  // namespace airfe
