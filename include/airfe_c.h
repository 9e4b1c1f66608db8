/* airfe_c.h -- the extern "C" boundary of the B200-native AirSLAM front-end.
 *
 * Plain pointers and sizes only; no torch / Eigen / OpenCV types.  The C++ class surfaces in
 * include/{feature_detector,plnet,super_point,light_glue,super_glue,point_matcher}.h are thin wrappers over
 * these entry points, exactly where the reference's classes call TensorRT:
 *   PLNet::infer            /root/reference/src/plnet.cpp:221-244   (executeV2 at :233 and :510)
 *   SuperPoint::infer       /root/reference/src/super_point.cpp:103-144 (executeV2 at :133)
 *   SuperPointLightGlue::infer /root/reference/src/light_glue.cpp:120-170 (executeV2 at :159)
 *   SuperGlue::infer        /root/reference/src/super_glue.cpp:137-197 (executeV2 at :185)
 *   PointMatcher::MatchingPoints /root/reference/src/point_matcher.cc:50-108
 * Every function returns AIRFE_OK (0) or a negative error code; airfe_last_error() gives the message.
 * There is no CPU fallback anywhere behind this header: without a CUDA device every compute call fails.
 */
#ifndef AIRFE_C_H_
#define AIRFE_C_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIRFE_OK 0
#define AIRFE_ERR_INVALID (-1)
#define AIRFE_ERR_CUDA (-2)
#define AIRFE_ERR_IO (-3)
#define AIRFE_ERR_CAPACITY (-4)

const char* airfe_last_error(void);

/* ---- low-level operator entry points (device pointers; used by the per-kernel parity tests) ---- */

/* Dense contraction on tcgen05 tensor cores (implicit-GEMM conv / batched GEMM); see csrc/tc_gemm.cuh.
 * Replaces the Conv / MatMul / Gemm / Einsum nodes TensorRT executes for the reference. */
int airfe_op_tc_gemm(const void* a, int a_C, int W, int H, int B, long long a_sx, long long a_sy, long long a_sb,
                     const void* bw, int k_total, int n_rows, long long bw_sn, long long bw_sbatch, int b_batches, int b_mn_major,
                     int taps, int c_in_pad, int block_n, const float* bias, int relu, int out_f32,
                     void* out, long long out_sb, long long out_sy, long long out_sx, int n_valid,
                     int tw, int th, int tb, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AIRFE_C_H_ */
