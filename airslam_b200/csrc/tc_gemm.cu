// Host side of the tcgen05 implicit-GEMM kernel: TMA tensor-map construction, tiling plan, launch.

#include "common.h"
#include "tc_gemm.cuh"

#include <cudaTypedefs.h>
#include <mutex>
#include <stdlib.h>

namespace airfe {

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
  });
  return fn;
}

// fp16 tensor map, `rank` dims, dims[0] innermost (contiguous); strides_bytes[i] for dims 1..rank-1.
bool make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return false;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: code %d (rank %d dims %llu %llu %llu box %u %u %u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], box[1],
              rank > 2 ? box[2] : 0);
    return false;
  }
  return true;
}

static int num_sms() { return device_sm_count(); }

bool tc_gemm_plan(const TcGemmDesc& d, TcGemmPlan* plan) {
  TcGemmParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  if (d.tw * d.th * d.tb != kTileM) { set_error("tc_gemm: tile box %dx%dx%d != 128 pixels", d.tw, d.th, d.tb); return false; }
  if (d.block_n % 16 || d.block_n < 16 || d.block_n > 256) { set_error("tc_gemm: bad block_n %d", d.block_n); return false; }
  if (d.b_mn_major && d.block_n != 64) { set_error("tc_gemm: MN-major B needs block_n == 64"); return false; }
  {
    uint64_t dims[4] = {(uint64_t)d.a_C, (uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.B};
    uint64_t str[3] = {(uint64_t)d.a_sx * 2, (uint64_t)d.a_sy * 2, (uint64_t)d.a_sb * 2};
    uint32_t box[4] = {64, (uint32_t)d.tw, (uint32_t)d.th, (uint32_t)d.tb};
    if (!make_tmap_f16(&p.tmA, d.a, 4, dims, str, box)) return false;
  }
  {
    const int nb = d.b_batches > 0 ? d.b_batches : 1;
    const int nh = d.b_heads > 0 ? d.b_heads : 1;
    const uint64_t rows = d.b_mn_major ? (uint64_t)d.k_total : (uint64_t)d.n_rows;
    const uint64_t s_head = (uint64_t)(nh > 1 ? d.bw_shead : d.bw_sn * (long long)rows) * 2;
    const uint64_t s_batch = (uint64_t)(nb > 1 ? d.bw_sbatch : d.bw_sn * (long long)rows * nh) * 2;
    if (d.b_mn_major) {
      uint64_t dims[4] = {(uint64_t)d.n_rows, (uint64_t)d.k_total, (uint64_t)nh, (uint64_t)nb};
      uint64_t str[3] = {(uint64_t)d.bw_sn * 2, s_head, s_batch};
      uint32_t box[4] = {64, 64, 1, 1};
      if (!make_tmap_f16(&p.tmB, d.bw, 4, dims, str, box)) return false;
    } else {
      uint64_t dims[4] = {(uint64_t)d.k_total, (uint64_t)d.n_rows, (uint64_t)nh, (uint64_t)nb};
      uint64_t str[3] = {(uint64_t)d.bw_sn * 2, s_head, s_batch};
      uint32_t box[4] = {64, (uint32_t)d.block_n, 1, 1};
      if (!make_tmap_f16(&p.tmB, d.bw, 4, dims, str, box)) return false;
    }
  }
  p.taps = d.taps;
  p.c_in_pad = d.c_in_pad;
  p.kblocks = d.c_in_pad / kBlockK;
  p.tw = d.tw; p.th = d.th; p.tb = d.tb;
  p.tiles_x = (d.W + d.tw - 1) / d.tw;
  p.tiles_y = (d.H + d.th - 1) / d.th;
  p.tiles_b = (d.B + d.tb - 1) / d.tb;
  p.block_n = d.block_n;
  p.n_tiles = (d.n_valid + d.block_n - 1) / d.block_n;
  p.W = d.W; p.H = d.H; p.B = d.B;
  p.b_batched = d.b_batches > 1 || d.b_heads > 1;
  p.b_batch_xor = d.b_batch_xor;
  p.scale = d.scale; p.scale_cols = d.scale_cols;
  p.rot = d.rot; p.rot_cols = d.rot_cols;
  p.out_split = d.out_split; p.out_split_stride = d.out_split_stride;
  if (d.out_split && (d.out_split % d.block_n)) { set_error("tc_gemm: out_split %d must be a multiple of block_n %d", d.out_split, d.block_n); return false; }
  if (d.rot && d.H != 1) { set_error("tc_gemm: fused rotary needs a row GEMM (H == 1)"); return false; }
  p.resid = d.resid;
  p.out2 = d.out2; p.out2_sb = d.out2_sb; p.out2_sy = d.out2_sy; p.out2_sx = d.out2_sx;
  p.b_mn_major = d.b_mn_major;
  p.bias = d.bias;
  p.relu = d.relu;
  p.out_f32 = d.out_f32;
  p.out = d.out;
  p.out_sb = d.out_sb; p.out_sy = d.out_sy; p.out_sx = d.out_sx;
  p.n_valid = d.n_valid;
  p.dyn_w = d.dyn_w;
  p.dyn_w_stride = d.dyn_w_stride;
  const int b_bytes = tc_b_bytes(d.block_n, d.b_mn_major);
  // TMA-store epilogue (tc_gemm.cuh): plain fp16 outputs whose N tile is a whole number of 64-column groups per epilogue warp.  Each warp's
  // 32 tile rows must be one box of the output tensor: the tile box sides are powers of two.  Parity-green (29 operator + 17 pipeline GPU tests ran
  // with it on) but measured SLOWER than the 256-bit STG path on every GEMM it applies to (profiles/r02c_tma_store_ab.txt: Wqkv + rotary 84 -> 88 us,
  // to_qk_v 38 -> 41 us, G3 496->128 137 -> 149 us): the staging tiles cost one A stage and the stores were not what these kernels wait for.
  // Off by default; AIRFE_GEMM_TMA_STORE=1 switches it on.
  static const int tma_store_on = getenv("AIRFE_GEMM_TMA_STORE") ? atoi(getenv("AIRFE_GEMM_TMA_STORE")) : 0;
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  p.tma_store = 0;
  // 16 epilogue warps (EW = 4).  Measured per GEMM on the B200 (profiles/r02e_wide_epilogue_ab.txt): -14 % on the 128->256 1x1 conv (N tile of 256
  // columns), -8 % on Wqkv + fused rotary (the rotary table loads are what the two-warp epilogue waits for), neutral on the other N = 128 row GEMMs,
  // +7 .. 18 % (slower) on narrow N tiles and on the G3 MLP -> default: N tiles of 256 columns and the rotary GEMMs.
  // AIRFE_GEMM_WIDE=0 never, =1 every GEMM (tests, A/B).  The choice depends on the GEMM's shape only, never on the batch.
  const int wide_mode = getenv("AIRFE_GEMM_WIDE") ? atoi(getenv("AIRFE_GEMM_WIDE")) : -1;
  p.wide = wide_mode >= 0 ? (wide_mode == 1 ? 1 : 0) : ((d.rot || d.block_n >= 256) ? 1 : 0);
  if (!p.wide && tma_store_on && !d.out_f32 && !d.resid && !d.out2 && d.block_n % 128 == 0 && pow2(d.tw) && pow2(d.th) && pow2(d.tb) && ((uintptr_t)d.out & 15) == 0 &&
      d.out_sx % 8 == 0 && (d.H == 1 || d.out_sy % 8 == 0) && (d.B == 1 || d.out_sb % 8 == 0) && (!d.out_split || (d.out_split % 64 == 0 && d.out_split_stride % 8 == 0))) {
    const int bw = d.tw < 32 ? d.tw : 32, bh = d.th < 32 / bw ? d.th : 32 / bw, bbx = 32 / (bw * bh);
    if (bbx <= d.tb) {
      const uint64_t cols = d.out_split ? (uint64_t)d.out_split : (uint64_t)d.n_valid;
      const uint64_t nsec = d.out_split ? (uint64_t)((d.n_valid + d.out_split - 1) / d.out_split) : 1;
      const uint64_t sx = (uint64_t)d.out_sx * 2, sy = (d.H == 1 || d.out_sy == 0) ? sx * d.W : (uint64_t)d.out_sy * 2;
      const uint64_t sb = (d.B == 1 || d.out_sb == 0) ? sy * d.H : (uint64_t)d.out_sb * 2;
      const uint64_t ss = d.out_split ? (uint64_t)d.out_split_stride * 2 : sb * d.B;
      uint64_t dims[5] = {cols, (uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.B, nsec};
      uint64_t str[4] = {sx, sy, sb, ss};
      uint32_t box[5] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bbx, 1};
      if (!make_tmap_f16(&p.tmO, d.out, 5, dims, str, box)) return false;
      p.tma_store = 1;
    }
  }
  const int budget = 200 * 1024 - (p.tma_store ? 24 * 1024 : 0);     // 32 KiB of staging tiles: 24 KiB out of the operand budget (a 128 KiB weight panel + 3 A stages still fit), 8 KiB of former slack
  const int total = p.tiles_x * p.tiles_y * p.tiles_b * p.n_tiles;
  const int panel = p.taps * p.kblocks * b_bytes;
  static const bool bres_off = getenv("AIRFE_GEMM_NO_BRES") != nullptr;
  // weights-resident: not for per-head / per-image B operands, and only when >= 3 A stages still fit and every CTA can keep one N tile
  p.b_resident = (!bres_off && !p.b_batched && panel + 3 * kABytes <= budget && p.n_tiles <= num_sms() && total > p.n_tiles) ? 1 : 0;
  int stages;
  if (p.b_resident) {
    stages = (budget - panel) / kABytes;
  } else {
    stages = budget / (kABytes + b_bytes);
  }
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  p.stages = stages;
  static const int prewait = getenv("AIRFE_PREWAIT") ? 1 : 0;   // measured in round 2: no gain (profiles/r02_prewait_ab.txt), off by default
  p.prewait = prewait;
  plan->smem_bytes = stages * (p.b_resident ? kABytes : kABytes + b_bytes) + (p.b_resident ? panel : 0) + 1024 /*align slack*/ + (2 * stages + 5) * 8 + 16 + (p.tma_store ? kTcStoreStage + 1024 + 32 : 0);
  int grid = total < num_sms() ? total : num_sms();
  if (p.b_resident) grid = grid / p.n_tiles * p.n_tiles;   // a CTA's N tile (blockIdx % n_tiles) must never change
  plan->grid = grid;
  plan->flops = 2.0 * (double)d.W * d.H * d.B * (double)d.n_valid * (double)d.taps * (double)d.a_C;
  return true;
}

using TcKernel = void (*)(const TcGemmParams);

bool tc_gemm_launch(const TcGemmPlan& plan, cudaStream_t stream) {
  static bool attr_set[kMaxDevices][8] = {};
  const int dev = current_device();
  const int key = (plan.p.wide ? 4 : 0) | (plan.p.b_mn_major ? 2 : 0) | (plan.p.b_resident ? 1 : 0);
  static const TcKernel kerns[8] = {tc_gemm_kernel<false, false>, tc_gemm_kernel<false, true>, tc_gemm_kernel<true, false>, tc_gemm_kernel<true, true>,
                                    tc_gemm_kernel<false, false, 4>, tc_gemm_kernel<false, true, 4>, tc_gemm_kernel<true, false, 4>, tc_gemm_kernel<true, true, 4>};
  TcKernel kern = kerns[key];
  if (!attr_set[dev][key]) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(tc_gemm_kernel) failed: %s", cudaGetErrorString(cudaGetLastError()));
      return false;
    }
    attr_set[dev][key] = true;
  }
  if (plan.grid <= 0) return true;
  cudaError_t e = launch_pdl(kern, plan.grid, plan.p.wide ? kTcWideThreads : kTcThreads, plan.smem_bytes, stream, plan.p);
  if (e != cudaSuccess) {
    set_error("tc_gemm launch failed: %s", cudaGetErrorString(e));
    return false;
  }
  return true;
}

}  // namespace airfe
