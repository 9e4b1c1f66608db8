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
 * Every function returns AIRFE_OK (0) or a negative error code; airfe_last_error(ctx) gives the message.
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

typedef struct airfe_ctx airfe_ctx; /* one per (device, model set); NOT thread-safe, like the reference's classes */

/* Message of the last failing call made on `ctx`; with ctx == NULL: of the last failing context-less call (airfe_create, the operator
 * entry points below) on the calling thread. */
const char* airfe_last_error(const airfe_ctx* ctx);

/* ---- low-level operator entry points (device pointers; used by the per-kernel parity tests) ---- */

/* Dense contraction on tcgen05 tensor cores (implicit-GEMM conv / batched GEMM); see csrc/tc_gemm.cuh.
 * Replaces the Conv / MatMul / Gemm / Einsum nodes TensorRT executes for the reference. */
int airfe_op_tc_gemm(const void* a, int a_C, int W, int H, int B, long long a_sx, long long a_sy, long long a_sb,
                     const void* bw, int k_total, int n_rows, long long bw_sn, long long bw_sbatch, int b_batches, int b_mn_major,
                     int taps, int c_in_pad, int block_n, const float* bias, int relu, int out_f32,
                     void* out, long long out_sb, long long out_sy, long long out_sx, int n_valid,
                     int tw, int th, int tb, void* stream);

/* 3x3 convolution (+bias, ReLU, optional fused 2x2 max-pool) on the halo-reuse tcgen05 kernel; see csrc/tc_conv3x3.cuh.
 * in: fp16 NHWC [B,H,W,in_ps]; w_packed: fp16 [n_rows][9 * c_pad] (tap-major, zero padded) with c_pad = c_in when c_in is a multiple of 32,
 * else round_up(c_in,64); out / pool_out may be NULL. */
int airfe_op_conv3x3(const void* in, int C, int W, int H, int B, long long in_ps, const void* w_packed, const float* bias, int n_rows, int c_in,
                     int relu, void* out, long long out_ps, void* pool_out, long long pool_ps, void* stream);

/* ---- frame-level entry points (host buffers in, host buffers out; pinned staging is internal) ---- */

typedef struct airfe_config {
  const char* weights_dir;        /* directory holding *.afw (converted once from the reference's ONNX files) */
  int max_batch;                  /* images per detect call / pairs per match call */
  /* PLNetConfig / SuperPointConfig fields (include/read_configs.h:9-83) */
  int max_keypoints;
  float keypoint_threshold;
  int remove_borders;
  float line_threshold;
  float line_length_threshold;
  /* PointMatcherConfig fields (include/read_configs.h:85-103) */
  int image_width, image_height;
  /* which networks to load */
  int enable_superpoint;          /* G1 */
  int enable_plnet;               /* G2 + G3 */
  int enable_lightglue;           /* G4 */
  int enable_superglue;           /* G5: 0 off, 1 indoor, 2 outdoor */
} airfe_config;

#define AIRFE_NET_SUPERPOINT 0
#define AIRFE_NET_PLNET 1
#define AIRFE_MATCHER_LIGHTGLUE 0
#define AIRFE_MATCHER_SUPERGLUE 1
#define AIRFE_FEAT_DIM 259

/* Page-locked host buffers.  Frame / feature buffers allocated here (or otherwise pinned) are DMA'd directly, skipping the
 * internal staging copy; pageable buffers work too. */
void* airfe_alloc_pinned(long long bytes);
void airfe_free_pinned(void* p);

void airfe_default_config(airfe_config* cfg);
int airfe_create(const airfe_config* cfg, int device, airfe_ctx** out);
void airfe_destroy(airfe_ctx* ctx);

/* Detect on `batch` 8-bit gray images of identical size (image i at gray + i * image_stride_bytes, rows `stride` bytes apart).
 * Replaces SuperPoint::infer / PLNet::infer.  Outputs per image i:
 *   feat    + i*feat_cap*259 : column-major 259 x n_feat[i] (score, x, y, 256-d descriptor), image-pixel coordinates
 *   lines   + i*line_cap*4   : n_lines[i] x (x1,y1,x2,y2) doubles (NULL => no line detection; requires net = PLNET)
 *   junc    + i*junc_cap*259 : junction features, as feat (NULL => junction_detection = false) */
int airfe_detect_batch(airfe_ctx* ctx, int net, int batch, const uint8_t* gray, int width, int height, int stride,
                       long long image_stride_bytes, float* feat, int feat_cap, int* n_feat, double* lines, int line_cap,
                       int* n_lines, float* junc, int junc_cap, int* n_junc);

/* Single image convenience (SURVEY.md 8b signature). */
int airfe_detect(airfe_ctx* ctx, int net, const uint8_t* gray, int width, int height, int stride, float* feat259_colmajor,
                 int feat_cap, int* n_feat, double* lines_xyxy, int line_cap, int* n_lines, float* junc259_colmajor, int junc_cap,
                 int* n_junc);

/* Match `pairs` feature-set pairs.  Replaces PointMatcher::MatchingPoints (without the optional OpenCV RANSAC, which stays a
 * host hook in include/point_matcher.h).  feat0/feat1: pair p at + p*feat_cap*259, column-major 259 x n (row 0 score, 1-2 x,y in
 * image pixels, 3.. descriptor); keypoints are normalised internally exactly like PointMatcher::NormalizeKeypoints with the
 * context's image_width/height.  Outputs per pair p (at + p*match_cap): idx0/idx1 (ascending idx0), score = exp(log-score)
 * (LightGlue) or (mscores0+mscores1)/2 (SuperGlue); DMatch::distance = 1 - score is formed by the C++ wrapper. */
int airfe_match_batch(airfe_ctx* ctx, int matcher, int pairs, const float* feat0, const int* n0, const float* feat1, const int* n1,
                      int feat_cap, int* idx0, int* idx1, float* score, int match_cap, int* n_match);

/* Same, for callers that already applied PointMatcher::NormalizeKeypoints (SuperPointLightGlue::infer / SuperGlue::infer take
 * normalised keypoints, src/point_matcher.cc:59-67). */
int airfe_match_batch_prenormalized(airfe_ctx* ctx, int matcher, int pairs, const float* feat0, const int* n0, const float* feat1,
                                    const int* n1, int feat_cap, int* idx0, int* idx1, float* score, int match_cap, int* n_match);

/* SuperGlue::infer surface (src/super_glue.cpp:137-197): per pair p, indices0 / mscores0 have n0[p] entries (at + p*out_cap),
 * indices1 / mscores1 n1[p] entries; -1 = unmatched.  (The class widens the scores to double.) */
int airfe_superglue_batch(airfe_ctx* ctx, int pairs, const float* feat0, const int* n0, const float* feat1, const int* n1, int feat_cap,
                          int prenormalized, int* indices0, int* indices1, float* mscores0, float* mscores1, int out_cap);

/* The keyframe path of MapBuilder::ExtractFeatureThread (src/map_builder.cc:85-86) for `pairs` stereo pairs at once:
 * Detect(left, right, ...) + MatchingPoints(left, right).  Features never leave the device between detect and match.
 * left/right: pair p at + p*image_stride_bytes.  Feature/line outputs as airfe_detect_batch with index 2*p (left), 2*p+1 (right);
 * junctions are produced for left images only, index p (src/feature_detector.cc:100-101).  lines/junc may be NULL. */
int airfe_detect_match_stereo_batch(airfe_ctx* ctx, int net, int matcher, int pairs, const uint8_t* left, const uint8_t* right,
                                    int width, int height, int stride, long long image_stride_bytes, float* feat, int feat_cap,
                                    int* n_feat, double* lines, int line_cap, int* n_lines, float* junc, int junc_cap, int* n_junc,
                                    int* idx0, int* idx1, float* score, int match_cap, int* n_match);

/* ---- image rectification on the device ----
 * Camera::UndistortImage (src/camera.cc:161-182: cv::remap(image, rect, map1, map2, INTER_LINEAR), called once per frame on the CPU at
 * src/map_builder.cc:43 and src/map_user.cc:108) is folded into the first kernel of the path: with maps installed and rectification
 * enabled, every detect entry point takes the RAW camera frames and the resize kernel gathers through the maps (bit-exact with cv::remap
 * followed by cv::resize; the rectified image is never materialised).
 * map_x / map_y: the CV_32F maps of cv::initUndistortRectifyMap (src/camera.cc:63-66, 71-74), width x height floats each, row-major.
 * side: 0 = left / mono camera, 1 = right camera. */
int airfe_set_rectify_maps(airfe_ctx* ctx, int side, const float* map_x, const float* map_y, int width, int height);
/* mode 0: off (frames are already rectified).  1: every image of a detect call belongs to camera `side 0`.  2: images alternate left, right
 * (what the stereo entry points assume automatically once a right-camera map is installed). */
int airfe_set_rectify(airfe_ctx* ctx, int mode);
/* The rectified image itself, for callers that keep it (MapBuilder::AddInput stores it in the input data): rect = remap(raw). */
int airfe_undistort(airfe_ctx* ctx, int side, const uint8_t* raw, int width, int height, int stride, uint8_t* rect, int rect_stride);

/* ---- point <-> line association and stereo line matching on the device ----
 * The step right after the path in Frame::AddLeftFeatures / AddRightFeatures (src/frame.cc:121-125, 141-187):
 *   AssignPointsToLines (src/line_processor.cc:68-120) for the left and the right image of every pair, and
 *   MatchLines (src/line_processor.cc:122-187) on the stereo point matches that pass the disparity filter of src/frame.cc:141-155
 *   (min_x_diff < |xl - xr| < max_x_diff, |yl - yr| <= max_y_diff: Camera::MinXDiff / MaxXDiff / MaxYDiff).
 * It runs on the DEVICE-RESIDENT results of the last airfe_detect_match_stereo_batch / airfe_stereo_device call of this context (lines,
 * features and matches never come back to the device).  Outputs, image slot s = 2*pair + side:
 *   rel_n   [2*pairs][line_cap]            points on each line (std::map<int,double>::size())
 *   rel_idx [2*pairs][line_cap][rel_cap]   their indices, ascending (map iteration order); rel_dist: double(float distance) as float
 *   line_matches [pairs][line_cap]         index of the matched right line per left line, -1 = none
 * Limits: 256 lines per image, 32 points per line, 8 lines per point -- beyond them the call returns AIRFE_ERR_CAPACITY. */
int airfe_stereo_line_assoc(airfe_ctx* ctx, int pairs, double min_x_diff, double max_x_diff, double max_y_diff, int line_cap, int rel_cap,
                            int* rel_n, int* rel_idx, float* rel_dist, int* line_matches);

/* ---- BoW quantisation of keyframe descriptors on the device ----
 * Database::FrameToBow (src/bow/database.cc:57-89): every 256-d descriptor walks the DBoW2 vocabulary tree (voc/point_voc_L4.bin, k = 10,
 * L = 4) to its word; words accumulate their idf weight into the BowVector, which is L1-normalised (the vocabulary's L1_NORM scoring).
 * airfe_bow_load: `path` is either the converted container weights/point_voc_L4.afw or the reference's Boost binary archive itself
 * (what Database::LoadVocabulary reads, src/bow/database.cc:15-24); the tree stays resident on the device.
 * airfe_bow_transform: feat259 (HOST or DEVICE pointer, column-major 259 x n as everywhere) ->
 *   word_of_feature [n]   word id per keypoint, 0xFFFFFFFF where the word's weight is <= 0 (database.cc:76-78)
 *   bow_ids / bow_vals    the BowVector: ascending word ids with their normalised weights (std::map order), *n_bow entries (<= n) */
int airfe_bow_load(airfe_ctx* ctx, const char* path);
int airfe_bow_transform(airfe_ctx* ctx, const float* feat259, int n, unsigned int* word_of_feature, unsigned int* bow_ids, double* bow_vals, int* n_bow);

/* ---- device-resident keyframe features + batched candidate matching (BASELINE.json config 5) ----
 * The reference re-uploads the same 259 x N keyframe features for every MatchingPoints call of a relocalization query
 * (MapUser::Relocalization, src/map_user.cc:363-376: up to GoodCandidateNum = 3 sequential calls) and of a loop-closure candidate
 * (src/map_refiner.cc:214-230: up to 5).  Here the map's keyframe features are uploaded ONCE (airfe_kf_put, the on-disk form is the
 * column-major 259 x N float block of include/utils.h:206-223) and all (query, candidate) jobs of a batch of queries run as batched
 * matcher launches that read both sides in place (no gather copy, no host round trip). */
int airfe_kf_reserve(airfe_ctx* ctx, int n_keyframes, int feat_cap);          /* (re)allocates the cache: n_keyframes slots of feat_cap columns */
int airfe_kf_put(airfe_ctx* ctx, int slot, const float* feat259_colmajor, int n);   /* host OR device source; n <= feat_cap columns */
int airfe_kf_size(const airfe_ctx* ctx);                                         /* reserved slots (0 = no cache) */
/* Job j matches query job_query[j] (feature set at query_feat + q*feat_cap*259, n = query_n[q]; query_feat may be a HOST or a DEVICE
 * pointer -- the all-gathered queries of the multi-GPU path never touch host memory; query_n is a host array) against keyframe slot
 * job_kf[j].  Outputs per job j: n_match[j] and, when idx0 != NULL, the match list at + j*match_cap exactly as airfe_match_batch
 * (PointMatcher::MatchingPoints without the OpenCV RANSAC hook). */
int airfe_reloc_match(airfe_ctx* ctx, int matcher, const float* query_feat, const int* query_n, int n_queries, int feat_cap, int n_jobs,
                      const int* job_query, const int* job_kf, int* n_match, int* idx0, int* idx1, float* score, int match_cap);
/* The winner rule of src/map_user.cc:370-373: the candidate with strictly more matches than every earlier one.  counts [n_queries][n_cand]
 * (negative = candidate absent); best_cand[q] = winning candidate column or -1, best_count[q] its match count. */
void airfe_reloc_pick(int n_queries, int n_cand, const int* counts, int* best_cand, int* best_count);

/* Same work with the images already resident in device memory (u8, slot s = 2*pair + side at d_images + s*image_stride_bytes) and
 * results left on the device; asynchronous on airfe_stream(ctx).  Used to time the device pipeline without PCIe. */
int airfe_stereo_device(airfe_ctx* ctx, int net, int matcher, int pairs, const void* d_images, int width, int height, int stride,
                        long long image_stride_bytes, int lines, int junctions);

/* One airfe_stereo_device step with CUDA events around every op; writes "name\tflops\tms\n" records into `out`
 * (returns the number of bytes written, <0 on error).  Used by bench.py for the live roofline numbers. */
long long airfe_profile_stereo(airfe_ctx* ctx, int net, int matcher, int pairs, const void* d_images, int width, int height, int stride,
                               long long image_stride_bytes, int lines, int junctions, char* out, long long cap);

/* Algorithmic tensor-core FLOPs and kernel launches of one airfe_stereo_device call (for roofline accounting). */
int airfe_stereo_cost(airfe_ctx* ctx, int net, int matcher, int pairs, int lines, double* tc_flops, int* launches);

/* Parity taps: copy a named intermediate of the last detect call (image `index`) to `dst`; returns bytes, <0 on error. */
long long airfe_debug_read(airfe_ctx* ctx, int net, const char* name, int index, void* dst, long long dst_bytes);

/* Authoring aid: when `dev_buf` (device memory, >= 64*8 int64) is non-NULL, CTA 0 of every following tc_conv3x3 launch writes
 * clock64 stamps of its first 64 tiles there (8 slots per tile: producer issue, MMA start / operands landed / issued,
 * epilogue start / end of the first and last epilogue warp).  NULL switches tracing off.  See tools/trace_conv.py. */
void airfe_debug_conv_trace(long long* dev_buf);

/* Same for the fused matcher kernels: `dev_buf` (>= 2048 int64) receives, from CTA 0 of every following launch, 16 stamps per row tile of
 * tc_ffn at [0, 128) and 16 stamps per query tile of tc_attn at [1024, 1152).  See tools/trace_match.py. */
void airfe_debug_match_trace(long long* dev_buf);

/* The CUDA stream all work of this context is issued on (cudaStream_t), for event timing by the caller. */
void* airfe_stream(airfe_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* AIRFE_C_H_ */
