// Low-level operator entry points of the C ABI (device pointers in, device pointers out).  These are what the
// per-kernel parity tests call; the frame-level entry points live in airfe_capi.cu.
#include "common.h"
#include "engine.h"
#include "../../include/airfe_c.h"

using namespace airfe;

extern "C" {

int airfe_op_tc_gemm(const void* a, int a_C, int W, int H, int B, long long a_sx, long long a_sy, long long a_sb,
                     const void* bw, int k_total, int n_rows, long long bw_sn, long long bw_sbatch, int b_batches, int b_mn_major,
                     int taps, int c_in_pad, int block_n, const float* bias, int relu, int out_f32,
                     void* out, long long out_sb, long long out_sy, long long out_sx, int n_valid,
                     int tw, int th, int tb, void* stream) {
  TcGemmDesc d;
  d.a = a; d.a_C = a_C; d.W = W; d.H = H; d.B = B; d.a_sx = a_sx; d.a_sy = a_sy; d.a_sb = a_sb;
  d.bw = bw; d.k_total = k_total; d.n_rows = n_rows; d.bw_sn = bw_sn; d.bw_sbatch = bw_sbatch; d.b_batches = b_batches; if (H > 1 && b_batches > 1) { d.b_heads = 0; }
  d.b_mn_major = b_mn_major; d.taps = taps; d.c_in_pad = c_in_pad; d.block_n = block_n; d.bias = bias; d.relu = relu;
  d.out_f32 = out_f32; d.out = out; d.out_sb = out_sb; d.out_sy = out_sy; d.out_sx = out_sx; d.n_valid = n_valid;
  d.tw = tw; d.th = th; d.tb = tb;
  TcGemmPlan plan;
  if (!tc_gemm_plan(d, &plan)) return AIRFE_ERR_INVALID;
  if (!tc_gemm_launch(plan, (cudaStream_t)stream)) return AIRFE_ERR_CUDA;
  return AIRFE_OK;
}

int airfe_op_conv3x3(const void* in, int C, int W, int H, int B, long long in_ps, const void* w_packed, const float* bias, int n_rows, int c_in,
                     int relu, void* out, long long out_ps, void* pool_out, long long pool_ps, void* stream) {
  Act a; a.p = const_cast<void*>(in); a.C = C; a.W = W; a.H = H; a.ps = in_ps;
  DenseW w; w.w = (__half*)w_packed; w.bias = const_cast<float*>(bias); w.n_rows = n_rows; w.c_in = c_in; w.c_in_pad = (c_in % 32 == 0) ? c_in : (c_in + 63) / 64 * 64;   /* multiples of 32 that are not multiples of 64 run on the K=32 (SWIZZLE_64B) path */ w.taps = 9;
  Act o, po;
  if (out) { o.p = out; o.C = n_rows; o.W = W; o.H = H; o.ps = out_ps; }
  if (pool_out) { po.p = pool_out; po.C = n_rows; po.W = W / 2; po.H = H / 2; po.ps = pool_ps; }
  OpList ol;
  if (!add_conv3x3(&ol, a, w, out ? &o : nullptr, pool_out ? &po : nullptr, B, relu != 0)) return AIRFE_ERR_INVALID;
  if (!ol.run((cudaStream_t)stream)) return AIRFE_ERR_CUDA;
  return AIRFE_OK;
}

void airfe_debug_conv_trace(long long* dev_buf) { conv3x3_set_trace(dev_buf); }
void airfe_debug_match_trace(long long* dev_buf) { match_set_trace(dev_buf); }

}  // extern "C"
