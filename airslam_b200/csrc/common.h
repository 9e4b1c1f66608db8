// Shared host-side declarations for libairfe (error reporting, plans for the tcgen05 GEMM, pointwise kernel launchers).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace airfe {

// Thread-local last-error string surfaced through airfe_last_error() (include/airfe_c.h).
void set_error(const char* fmt, ...);
const char* get_error();

#define AIRFE_CUDA_OK(expr)                                                                              \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) {                                                                              \
      airfe::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);      \
      return false;                                                                                       \
    }                                                                                                     \
  } while (0)

struct TcGemmParams;

// Per-device one-time state (cudaFuncSetAttribute opt-ins, SM counts): the function attribute and the SM count belong to a device, and a
// process may hold contexts on several devices (airfe_create takes a device ordinal), so nothing here may be cached process-wide.
constexpr int kMaxDevices = 64;
inline int current_device() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < kMaxDevices) ? d : 0; }
int device_sm_count();   // multiprocessors of the CURRENT device (cached per device)

// Launch with programmatic dependent launch (PDL) allowed: the kernel may become resident while its predecessor in the stream is
// still draining (it called griddepcontrol.launch_dependents), runs its prologue (barrier init, TMEM allocation, tensor-map
// prefetch, bias staging) and then blocks in griddepcontrol.wait until the predecessor's results are visible.  Every kernel
// launched through here MUST execute ptx::pdl_wait() before it touches data produced by earlier kernels.  AIRFE_NO_PDL=1 disables.
bool pdl_enabled();
template <typename P>
inline cudaError_t launch_pdl(void (*kern)(const P), int grid, int block, size_t smem, cudaStream_t st, const P& params) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, params);
}

// Description of one dense contraction for tc_gemm_plan().
struct TcGemmDesc {
  // A: NHWC fp16 view (element strides), a_C valid channels
  const void* a = nullptr;
  int a_C = 0, W = 0, H = 1, B = 1;
  long long a_sx = 0, a_sy = 0, a_sb = 0;
  // Bw: [n_rows][k_total] K-major fp16 (or MN-major [k_total][n_rows]); optional batch
  const void* bw = nullptr;
  int k_total = 0, n_rows = 0;
  long long bw_sn = 0, bw_sbatch = 0;
  int b_batches = 0, b_mn_major = 0;   // b_batches > 1: Bw has a batch dim (stride bw_sbatch) indexed by the tile batch
  int b_heads = 0; long long bw_shead = 0; // with b_batches: Bw head dim (indexed by the tile's y) and its stride
  int b_batch_xor = 0;
  int taps = 1, c_in_pad = 64;
  int block_n = 64;
  const float* bias = nullptr;
  float scale = 1.f; int scale_cols = 0;
  const float* rot = nullptr; int rot_cols = 0;          // fused rotary embedding (see TcGemmParams)
  int out_split = 0; long long out_split_stride = 0;     // column sections stored to separate matrices
  const float* resid = nullptr;
  void* out2 = nullptr; long long out2_sb = 0, out2_sy = 0, out2_sx = 0;
  int relu = 0, out_f32 = 0;
  void* out = nullptr;
  long long out_sb = 0, out_sy = 0, out_sx = 0;
  int n_valid = 0;
  int tw = 128, th = 1, tb = 1;
  const int* dyn_w = nullptr;
  int dyn_w_stride = 1;
};

}  // namespace airfe

#include "tc_gemm_params.h"

namespace airfe {
struct TcGemmPlan {
  TcGemmParams p;
  int grid = 0;
  int smem_bytes = 0;
  double flops = 0;
};
bool make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                   int swizzle_bytes = 128);
bool tc_gemm_plan(const TcGemmDesc& d, TcGemmPlan* plan);
bool tc_gemm_launch(const TcGemmPlan& plan, cudaStream_t stream);
}  // namespace airfe
