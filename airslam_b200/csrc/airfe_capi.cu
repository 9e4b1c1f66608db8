// Frame-level C ABI (include/airfe_c.h): context, detect.  Host buffers in/out through pinned staging; all device work
// is asynchronous on the context's stream and synchronised once per call, right before results are handed back.
#include "../../include/airfe_c.h"
#include "detector.h"
#include "matcher.h"
#include "assoc_kernels.h"
#include "bow.h"

#include <map>
#include <memory>
#include <set>
#include <stdlib.h>
#include <string>
#include <vector>

using namespace airfe;

struct airfe_ctx {
  int device = 0;
  char err[1024] = "";                                // message of the last failing call on THIS context (airfe_last_error(ctx))
  airfe_config cfg;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;                 // D2H of the detector results overlaps the matcher (stereo entry point)
  cudaEvent_t ev_det = nullptr, ev_copy = nullptr;
  std::unique_ptr<Detector> sp, pl;
  std::unique_ptr<LightGlue> lg;
  std::unique_ptr<SuperGlue> sg;
  // Contexts sized for the reference's 1024-keypoint TRT profile (max_keypoints > 512: the C++ PointMatcher surface) also hold a 512-row
  // instance: calls whose feature sets fit run on it, i.e. on the FUSED attention kernel (tc_attn.cuh covers <= 512 keys); only genuinely
  // larger sets take the unfused 1024-row path.  lg_use / sg_use = the instance the current call runs on.
  std::unique_ptr<LightGlue> lg_small;
  std::unique_ptr<SuperGlue> sg_small;
  LightGlue* lg_use = nullptr;
  SuperGlue* sg_use = nullptr;
  int* h_sgidx = nullptr; float* h_sgms = nullptr;   // [2][max_batch][1024] each
  float* d_mfeat = nullptr; int* d_mn = nullptr;   // staging for host-provided features [2*max_batch][kKpCap][259]
  float* h_mfeat = nullptr; int* h_mn = nullptr; int* h_midx = nullptr; float* h_mscore = nullptr; int* h_mcount = nullptr;
  // pinned staging
  uint8_t* h_img = nullptr; size_t h_img_bytes = 0;
  float* h_feat = nullptr; float* h_junc = nullptr; float* h_lines = nullptr;
  int* h_counts = nullptr;   // [3][max_batch]
  uint8_t* d_img = nullptr; size_t d_img_bytes = 0;
  // device-resident keyframe features (airfe_kf_*) and the job tables of airfe_reloc_match
  float* kf_feat = nullptr; int* kf_n = nullptr;       // [kf_slots][kf_cap][259] on the device; counts on the host
  int kf_slots = 0, kf_cap = 0;
  int rj_cap = 0, rj_kp_cap = 0;                       // jobs / keypoints per set the tables below are sized for
  const float** h_rptr = nullptr; const float** d_rptr = nullptr;   // [2*rj_cap] per-slot feature pointers
  int* h_rn = nullptr; int* d_rn = nullptr;            // [2*rj_cap] per-slot counts
  int* d_rcount = nullptr; int* d_ridx = nullptr; float* d_rscore = nullptr;   // per job: match count, [cap][2] indices, [cap] scores
  int* h_rcount = nullptr; int* h_ridx = nullptr; float* h_rscore = nullptr;
  float* d_qfeat = nullptr; size_t d_qfeat_bytes = 0;  // staging for host-resident query features
  // rectification maps (airfe_set_rectify_maps): fixed-point form of cv::remap, per camera side
  RemapMaps remap = {};
  short* d_rxy[2] = {nullptr, nullptr}; unsigned short* d_ra[2] = {nullptr, nullptr};
  int remap_w = 0, remap_h = 0;
  uint8_t* d_rect = nullptr; size_t d_rect_bytes = 0;
  std::unique_ptr<BowVocabulary> voc;                 // airfe_bow_load
  float* d_bowfeat = nullptr; size_t d_bowfeat_rows = 0;
  // point <-> line association (airfe_stereo_line_assoc): state of the last stereo call + device scratch
  int last_net = -1, last_matcher = -1, last_pairs = 0, last_w = 0, last_h = 0;
  int la_pairs = 0;
  int *la_rel_n = nullptr, *la_rel_idx = nullptr, *la_cnt = nullptr, *la_row = nullptr, *la_lm = nullptr, *la_ovf = nullptr;
  float* la_rel_dist = nullptr;
  // CUDA graphs of the per-call class-surface paths (one stereo pair per call): key = everything that shapes the launch sequence
  std::map<std::string, cudaGraphExec_t> graphs;
  std::map<std::string, int> graph_seen;
  std::set<std::string> graph_bad;
  int graph_launches = 0;
};

// Per-call (batch <= 2 pairs) paths are launch-latency bound: ~90 (detector) / ~70 (matcher) kernels of a few microseconds each.  The second
// call with a given shape key is stream-captured (copies from / to the context's pinned staging included) and every later one is a single
// cudaGraphLaunch.  The first call stays eager: lazy initialisation (resize tables, function attributes, op lists) must not happen under
// capture.  Anything that goes wrong marks the key as not graphable and the call runs eagerly -- never a different result, only slower.
// AIRFE_NO_GRAPH=1 disables graphs.
template <class F>
static bool run_graphed(airfe_ctx* c, const std::string& key, cudaStream_t st, F&& body) {
  static const bool on = getenv("AIRFE_NO_GRAPH") == nullptr;
  if (!on || profiler().on || c->graph_bad.count(key)) return body();
  auto it = c->graphs.find(key);
  if (it != c->graphs.end()) {
    if (cudaGraphLaunch(it->second, st) == cudaSuccess) { c->graph_launches++; return true; }
    cudaGetLastError();
    c->graph_bad.insert(key);
    return body();
  }
  if (c->graph_seen[key]++ == 0) return body();
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); c->graph_bad.insert(key); return body(); }
  const bool ok = body();
  cudaGraph_t g = nullptr;
  const cudaError_t e = cudaStreamEndCapture(st, &g);
  cudaGraphExec_t ex = nullptr;
  if (!ok || e != cudaSuccess || !g || cudaGraphInstantiate(&ex, g, 0) != cudaSuccess) {
    cudaGetLastError();
    if (g) cudaGraphDestroy(g);
    c->graph_bad.insert(key);
    return body();                         // nothing ran during the failed capture: run it for real
  }
  cudaGraphDestroy(g);
  c->graphs[key] = ex;
  if (cudaGraphLaunch(ex, st) != cudaSuccess) { cudaGetLastError(); c->graph_bad.insert(key); c->graphs.erase(key); cudaGraphExecDestroy(ex); return body(); }
  c->graph_launches++;
  return true;
}

// The maps to hand to the detector for a call on w x h frames, or nullptr (rectification off).  `stereo`: the call interleaves left / right.
static const RemapMaps* remap_for(airfe_ctx* c, int w, int h, bool stereo, RemapMaps* tmp, bool* size_error) {
  *size_error = false;
  if (!c->remap.mode) return nullptr;
  if (w != c->remap_w || h != c->remap_h) { *size_error = true; set_error("rectification maps are %dx%d, frames %dx%d", c->remap_w, c->remap_h, w, h); return nullptr; }
  *tmp = c->remap;
  if (stereo) tmp->mode = c->d_rxy[1] ? 2 : 1;
  if (tmp->mode == 2 && !c->d_rxy[1]) { *size_error = true; set_error("rectification mode 2 needs the right-camera map"); return nullptr; }
  return tmp;
}

// Every failing frame-level call leaves its message in the context it was made on (contexts are per thread / per device; a
// process-global string would let one thread's error overwrite another's).  Failures without a context (airfe_create, the operator
// entry points) are readable through airfe_last_error(NULL) on the calling thread.
static int fail(airfe_ctx* c, int code) {
  if (c) { strncpy(c->err, get_error(), sizeof(c->err) - 1); c->err[sizeof(c->err) - 1] = 0; }
  return code;
}

// Captured graphs hold raw pointers into the staging buffers and rectification maps: drop them whenever one of those is reallocated.
static void drop_graphs(airfe_ctx* c) {
  for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
  c->graphs.clear(); c->graph_seen.clear(); c->graph_bad.clear();
}

// (Re)allocate the pinned + device frame staging.  On failure the context holds NO staging (pointers null, sizes zero), so that neither
// the next call nor airfe_destroy touches freed memory.
static bool grow_image_staging(airfe_ctx* c, size_t need) {
  cudaStreamSynchronize(c->stream);
  drop_graphs(c);
  if (c->h_img) cudaFreeHost(c->h_img);
  if (c->d_img) cudaFree(c->d_img);
  c->h_img = nullptr; c->d_img = nullptr; c->h_img_bytes = c->d_img_bytes = 0;
  if (cudaMallocHost(&c->h_img, need) != cudaSuccess) { c->h_img = nullptr; cudaGetLastError(); set_error("pinned staging allocation (%zu bytes) failed", need); return false; }
  if (cudaMalloc(&c->d_img, need) != cudaSuccess) {
    cudaGetLastError(); cudaFreeHost(c->h_img); c->h_img = nullptr; c->d_img = nullptr;
    set_error("device staging allocation (%zu bytes) failed", need);
    return false;
  }
  c->h_img_bytes = c->d_img_bytes = need;
  return true;
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

extern "C" {

const char* airfe_last_error(const airfe_ctx* ctx) { return ctx ? ctx->err : get_error(); }

void* airfe_alloc_pinned(long long bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) { set_error("cudaMallocHost(%lld) failed", bytes); return nullptr; }
  return p;
}
void airfe_free_pinned(void* p) { if (p) cudaFreeHost(p); }

void airfe_default_config(airfe_config* c) {
  memset(c, 0, sizeof(*c));
  c->weights_dir = "weights";
  c->max_batch = 8;
  c->max_keypoints = 400;          // configs/visual_odometry/vo_euroc.yaml:3-7
  c->keypoint_threshold = 0.004f;
  c->remove_borders = 4;
  c->line_threshold = 0.75f;
  c->line_length_threshold = 50.f;
  c->image_width = 752;            // vo_euroc.yaml:11-12
  c->image_height = 480;
  c->enable_superpoint = 1;
  c->enable_plnet = 1;
  c->enable_lightglue = 1;
  c->enable_superglue = 0;
}

int airfe_create(const airfe_config* cfg, int device, airfe_ctx** out) {
  if (!cfg || !out) { set_error("null argument"); return AIRFE_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    set_error("no CUDA device: libairfe has no CPU path");
    return AIRFE_ERR_CUDA;
  }
  if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", device); return AIRFE_ERR_CUDA; }
  // a context that fails half-way releases everything it already holds (streams, arenas, pinned staging): airfe_destroy copes with null members
  struct Guard { airfe_ctx* p; ~Guard() { if (p) airfe_destroy(p); } airfe_ctx* operator->() const { return p; } airfe_ctx* release() { airfe_ctx* r = p; p = nullptr; return r; } };
  Guard c{new airfe_ctx};
  c->device = device;
  c->cfg = *cfg;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_det, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming) != cudaSuccess) {
    set_error("stream creation failed");
    return AIRFE_ERR_CUDA;
  }
  DetectorConfig dc;
  dc.max_batch = 2 * cfg->max_batch;   // a stereo batch of max_batch pairs = 2*max_batch images
  dc.max_keypoints = cfg->max_keypoints;
  dc.keypoint_threshold = cfg->keypoint_threshold;
  dc.remove_borders = cfg->remove_borders;
  dc.line_threshold = cfg->line_threshold;
  dc.line_length_threshold = cfg->line_length_threshold;
  const std::string wdir = cfg->weights_dir ? cfg->weights_dir : "weights";
  if (cfg->enable_superpoint) {
    c->sp.reset(new Detector);
    dc.enable_lines = false;
    if (!c->sp->init(dc, wdir, false)) return AIRFE_ERR_IO;
  }
  if (cfg->enable_plnet) {
    c->pl.reset(new Detector);
    dc.enable_lines = true;
    if (!c->pl->init(dc, wdir, true)) return AIRFE_ERR_IO;
  }
  if (cfg->enable_superglue) {
    MatcherConfig mc;
    mc.max_pairs = cfg->max_batch;
    mc.cap = cfg->max_keypoints <= 512 ? 512 : 1024;
    mc.image_width = cfg->image_width;
    mc.image_height = cfg->image_height;
    c->sg.reset(new SuperGlue);
    if (!c->sg->init(mc, wdir, cfg->enable_superglue == 2)) return AIRFE_ERR_IO;
    c->sg_use = c->sg.get();
    if (mc.cap > 512) {
      mc.cap = 512;
      c->sg_small.reset(new SuperGlue);
      if (!c->sg_small->init(mc, wdir, cfg->enable_superglue == 2)) return AIRFE_ERR_IO;
    }
    if (cudaMallocHost(&c->h_sgidx, (size_t)2 * cfg->max_batch * 1024 * 4) != cudaSuccess || cudaMallocHost(&c->h_sgms, (size_t)2 * cfg->max_batch * 1024 * 4) != cudaSuccess) {
      set_error("superglue staging allocation failed");
      return AIRFE_ERR_CUDA;
    }
  }
  if (cfg->enable_lightglue || cfg->enable_superglue) {
    const size_t S = 2 * (size_t)cfg->max_batch;
    if (cudaMalloc(&c->d_mfeat, S * kKpCap * 259 * 4) != cudaSuccess || cudaMalloc(&c->d_mn, S * 4) != cudaSuccess ||
        cudaMallocHost(&c->h_mfeat, S * kKpCap * 259 * 4) != cudaSuccess || cudaMallocHost(&c->h_mn, S * 4) != cudaSuccess ||
        cudaMallocHost(&c->h_midx, (size_t)cfg->max_batch * 1024 * 2 * 4) != cudaSuccess ||
        cudaMallocHost(&c->h_mscore, (size_t)cfg->max_batch * 1024 * 4) != cudaSuccess || cudaMallocHost(&c->h_mcount, (size_t)cfg->max_batch * 4) != cudaSuccess) {
      set_error("matcher staging allocation failed");
      return AIRFE_ERR_CUDA;
    }
  }
  if (cfg->enable_lightglue) {
    MatcherConfig mc;
    mc.max_pairs = cfg->max_batch;
    mc.cap = cfg->max_keypoints <= 512 ? 512 : 1024;
    mc.image_width = cfg->image_width;
    mc.image_height = cfg->image_height;
    c->lg.reset(new LightGlue);
    if (!c->lg->init(mc, wdir)) return AIRFE_ERR_IO;
    c->lg_use = c->lg.get();
    if (mc.cap > 512) {
      mc.cap = 512;
      c->lg_small.reset(new LightGlue);
      if (!c->lg_small->init(mc, wdir)) return AIRFE_ERR_IO;
    }
  }
  const int B = cfg->max_batch;
  if (cudaMallocHost(&c->h_feat, (size_t)2 * B * kKpCap * 259 * 4) != cudaSuccess || cudaMallocHost(&c->h_junc, (size_t)2 * B * kKpCap * 259 * 4) != cudaSuccess ||
      cudaMallocHost(&c->h_lines, (size_t)2 * B * kLineCap * 4 * 4) != cudaSuccess || cudaMallocHost(&c->h_counts, (size_t)6 * B * 4) != cudaSuccess) {
    set_error("pinned allocation failed");
    return AIRFE_ERR_CUDA;
  }
  *out = c.release();
  return AIRFE_OK;
}

void airfe_destroy(airfe_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  c->sp.reset();
  c->pl.reset();
  c->lg.reset();
  c->sg.reset();
  c->lg_small.reset();
  c->sg_small.reset();
  if (c->h_sgidx) cudaFreeHost(c->h_sgidx);
  if (c->h_sgms) cudaFreeHost(c->h_sgms);
  if (c->d_mfeat) cudaFree(c->d_mfeat);
  if (c->d_mn) cudaFree(c->d_mn);
  if (c->h_mfeat) cudaFreeHost(c->h_mfeat);
  if (c->h_mn) cudaFreeHost(c->h_mn);
  if (c->h_midx) cudaFreeHost(c->h_midx);
  if (c->h_mscore) cudaFreeHost(c->h_mscore);
  if (c->h_mcount) cudaFreeHost(c->h_mcount);
  if (c->h_img) cudaFreeHost(c->h_img);
  if (c->h_feat) cudaFreeHost(c->h_feat);
  if (c->h_junc) cudaFreeHost(c->h_junc);
  if (c->h_lines) cudaFreeHost(c->h_lines);
  if (c->h_counts) cudaFreeHost(c->h_counts);
  if (c->d_img) cudaFree(c->d_img);
  if (c->kf_feat) cudaFree(c->kf_feat);
  delete[] c->kf_n;
  if (c->h_rptr) cudaFreeHost(c->h_rptr);
  if (c->d_rptr) cudaFree(c->d_rptr);
  if (c->h_rn) cudaFreeHost(c->h_rn);
  if (c->d_rn) cudaFree(c->d_rn);
  if (c->d_rcount) cudaFree(c->d_rcount);
  if (c->d_ridx) cudaFree(c->d_ridx);
  if (c->d_rscore) cudaFree(c->d_rscore);
  if (c->h_rcount) cudaFreeHost(c->h_rcount);
  if (c->h_ridx) cudaFreeHost(c->h_ridx);
  if (c->h_rscore) cudaFreeHost(c->h_rscore);
  if (c->d_qfeat) cudaFree(c->d_qfeat);
  for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
  c->voc.reset();
  if (c->d_bowfeat) cudaFree(c->d_bowfeat);
  if (c->la_rel_n) cudaFree(c->la_rel_n);
  if (c->la_rel_idx) cudaFree(c->la_rel_idx);
  if (c->la_rel_dist) cudaFree(c->la_rel_dist);
  if (c->la_cnt) cudaFree(c->la_cnt);
  if (c->la_row) cudaFree(c->la_row);
  if (c->la_lm) cudaFree(c->la_lm);
  if (c->la_ovf) cudaFree(c->la_ovf);
  for (int k = 0; k < 2; ++k) { if (c->d_rxy[k]) cudaFree(c->d_rxy[k]); if (c->d_ra[k]) cudaFree(c->d_ra[k]); }
  if (c->d_rect) cudaFree(c->d_rect);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->ev_det) cudaEventDestroy(c->ev_det);
  if (c->ev_copy) cudaEventDestroy(c->ev_copy);
  delete c;
}

void* airfe_stream(airfe_ctx* c) { return c ? (void*)c->stream : nullptr; }

static Detector* pick(airfe_ctx* c, int net) {
  Detector* d = net == AIRFE_NET_SUPERPOINT ? c->sp.get() : (net == AIRFE_NET_PLNET ? c->pl.get() : nullptr);
  if (!d) set_error("network %d not enabled in this context", net);
  return d;
}

int airfe_detect_batch(airfe_ctx* c, int net, int batch, const uint8_t* gray, int w, int h, int stride, long long img_stride,
                       float* feat, int feat_cap, int* n_feat, double* lines, int line_cap, int* n_lines, float* junc, int junc_cap,
                       int* n_junc) {
  if (!c || !gray || !feat || !n_feat) { set_error("null argument"); return fail(c, AIRFE_ERR_INVALID); }
  if (w <= 1 || h <= 1 || stride < w) { set_error("empty image"); return fail(c, AIRFE_ERR_INVALID); }   // image.empty() -> false (plnet.cpp:247)
  Detector* d = pick(c, net);
  if (!d) return fail(c, AIRFE_ERR_INVALID);
  if (batch < 1 || batch > 2 * c->cfg.max_batch) { set_error("batch %d outside [1,%d]", batch, 2 * c->cfg.max_batch); return fail(c, AIRFE_ERR_INVALID); }
  if ((lines || junc) && net != AIRFE_NET_PLNET) { set_error("lines / junctions need the PLNet network"); return fail(c, AIRFE_ERR_INVALID); }
  if (junc && !lines) { set_error("junction detection needs line detection"); return fail(c, AIRFE_ERR_INVALID); }
  if ((lines && !n_lines) || (junc && !n_junc)) { set_error("n_lines / n_junc must be given with their buffers"); return fail(c, AIRFE_ERR_INVALID); }
  cudaSetDevice(c->device);
  const size_t one = (size_t)h * stride;
  const size_t need = one * batch;
  if (need > c->h_img_bytes) {
    if (!grow_image_staging(c, need)) return fail(c, AIRFE_ERR_CUDA);
  }
  for (int i = 0; i < batch; ++i) memcpy(c->h_img + one * i, gray + (size_t)img_stride * i, one);
  cudaStream_t st = c->stream;
  RemapMaps rm; bool rm_err;
  const RemapMaps* rmp = remap_for(c, w, h, false, &rm, &rm_err);
  if (rm_err) return fail(c, AIRFE_ERR_INVALID);
  const DetectOutputs& o = d->out();
  const int B = 2 * c->cfg.max_batch;
  int* hc = c->h_counts;
  const int kmax = c->cfg.max_keypoints;
  // Small batches (the class surface: 1 or 2 images per call) run as ONE graph including every copy; the line / junction read-backs are
  // then sized by the caller's capacities instead of by the counts, so that the whole call needs a single synchronisation.
  const bool one_shot = batch <= 4;
  const int lrows = line_cap < kLineCap ? line_cap : kLineCap, jrows = junc_cap < kJunc ? junc_cap : kJunc;
  auto body = [&]() -> bool {
    if (cudaMemcpyAsync(c->d_img, c->h_img, need, cudaMemcpyHostToDevice, st) != cudaSuccess) { set_error("H2D failed"); return false; }
    if (!d->run(c->d_img, batch, w, h, stride, (long long)one, lines != nullptr, junc != nullptr, st, rmp)) return false;
    cudaMemcpyAsync(hc, o.n_feat, 4 * batch, cudaMemcpyDeviceToHost, st);
    if (lines) cudaMemcpyAsync(hc + B, o.n_lines, 4 * batch, cudaMemcpyDeviceToHost, st);
    if (junc) cudaMemcpyAsync(hc + 2 * B, o.n_junc, 4 * batch, cudaMemcpyDeviceToHost, st);
    // counts are bounded by max_keypoints, so copy that many feature columns without waiting for the counts
    for (int i = 0; i < batch; ++i)
      cudaMemcpyAsync(c->h_feat + (size_t)i * kKpCap * 259, o.feat + (size_t)i * kKpCap * 259, (size_t)kmax * 259 * 4, cudaMemcpyDeviceToHost, st);
    if (one_shot && lines)
      for (int i = 0; i < batch; ++i) {
        if (lrows) cudaMemcpyAsync(c->h_lines + (size_t)i * kLineCap * 4, o.lines + (size_t)i * kLineCap * 4, (size_t)lrows * 16, cudaMemcpyDeviceToHost, st);
        if (junc && jrows) cudaMemcpyAsync(c->h_junc + (size_t)i * kKpCap * 259, o.junc + (size_t)i * kKpCap * 259, (size_t)jrows * 259 * 4, cudaMemcpyDeviceToHost, st);
      }
    return cudaGetLastError() == cudaSuccess;
  };
  if (one_shot) {
    char key[160];
    snprintf(key, sizeof(key), "det:%d:%d:%dx%d:%d:%d:%d:%d:%d:%d", net, batch, w, h, stride, lines ? 1 : 0, junc ? 1 : 0, rmp ? rmp->mode : 0, lrows, jrows);
    if (!run_graphed(c, key, st, body)) return fail(c, AIRFE_ERR_CUDA);
  } else if (!body()) {
    return fail(c, AIRFE_ERR_CUDA);
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("detect failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(c, AIRFE_ERR_CUDA); }
  for (int i = 0; i < batch; ++i) {
    const int n = hc[i] < feat_cap ? hc[i] : feat_cap;
    n_feat[i] = n;
    memcpy(feat + (size_t)i * feat_cap * 259, c->h_feat + (size_t)i * kKpCap * 259, (size_t)n * 259 * 4);
  }
  if (lines) {
    const double ws = (double)((float)w / 512.f), hs = (double)((float)h / 512.f);
    for (int i = 0; i < batch; ++i) {
      const int n = hc[B + i] < line_cap ? hc[B + i] : line_cap;
      n_lines[i] = n;
      if (n && !one_shot) cudaMemcpyAsync(c->h_lines + (size_t)i * kLineCap * 4, o.lines + (size_t)i * kLineCap * 4, (size_t)n * 16, cudaMemcpyDeviceToHost, st);
    }
    if (junc)
      for (int i = 0; i < batch; ++i) {
        const int n = hc[2 * B + i] < junc_cap ? hc[2 * B + i] : junc_cap;
        n_junc[i] = n;
        if (n && !one_shot) cudaMemcpyAsync(c->h_junc + (size_t)i * kKpCap * 259, o.junc + (size_t)i * kKpCap * 259, (size_t)n * 259 * 4, cudaMemcpyDeviceToHost, st);
      }
    if (!one_shot) cudaStreamSynchronize(st);
    for (int i = 0; i < batch; ++i) {
      const float* l = c->h_lines + (size_t)i * kLineCap * 4;
      double* dl = lines + (size_t)i * line_cap * 4;
      for (int k = 0; k < n_lines[i]; ++k) {   // Vector4d(x1..y2) then *= w_scale / h_scale in double (plnet.cpp:577-582)
        dl[k * 4 + 0] = (double)l[k * 4 + 0] * ws; dl[k * 4 + 1] = (double)l[k * 4 + 1] * hs;
        dl[k * 4 + 2] = (double)l[k * 4 + 2] * ws; dl[k * 4 + 3] = (double)l[k * 4 + 3] * hs;
      }
      if (junc) memcpy(junc + (size_t)i * junc_cap * 259, c->h_junc + (size_t)i * kKpCap * 259, (size_t)n_junc[i] * 259 * 4);
    }
  }
  return AIRFE_OK;
}

int airfe_detect(airfe_ctx* c, int net, const uint8_t* gray, int w, int h, int stride, float* feat, int feat_cap, int* n_feat,
                 double* lines, int line_cap, int* n_lines, float* junc, int junc_cap, int* n_junc) {
  return airfe_detect_batch(c, net, 1, gray, w, h, stride, 0, feat, feat_cap, n_feat, lines, line_cap, n_lines, junc, junc_cap, n_junc);
}

static void fetch_matches_enqueue(airfe_ctx* c, int pairs, int matcher) {
  const bool sgm = matcher == AIRFE_MATCHER_SUPERGLUE;
  const int cap = sgm ? c->sg_use->cap() : c->lg_use->cap();
  cudaStream_t st = c->stream;
  cudaMemcpyAsync(c->h_mcount, sgm ? c->sg_use->out().m_count : c->lg_use->out().count, 4 * pairs, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(c->h_midx, sgm ? c->sg_use->out().m_idx : c->lg_use->out().idx, (size_t)pairs * cap * 8, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(c->h_mscore, sgm ? c->sg_use->out().m_score : c->lg_use->out().score, (size_t)pairs * cap * 4, cudaMemcpyDeviceToHost, st);
}

static int fetch_matches(airfe_ctx* c, int pairs, int* idx0, int* idx1, float* score, int match_cap, int* n_match, const int* zero_mask,
                         int matcher = AIRFE_MATCHER_LIGHTGLUE, bool enqueued = false) {
  const bool sgm = matcher == AIRFE_MATCHER_SUPERGLUE;
  const int cap = sgm ? c->sg_use->cap() : c->lg_use->cap();
  const int* d_count = sgm ? c->sg_use->out().m_count : c->lg_use->out().count;
  const int* d_idx = sgm ? c->sg_use->out().m_idx : c->lg_use->out().idx;
  const float* d_score = sgm ? c->sg_use->out().m_score : c->lg_use->out().score;
  cudaStream_t st = c->stream;
  if (!enqueued) {
    cudaMemcpyAsync(c->h_mcount, d_count, 4 * pairs, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(c->h_midx, d_idx, (size_t)pairs * cap * 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(c->h_mscore, d_score, (size_t)pairs * cap * 4, cudaMemcpyDeviceToHost, st);
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("match failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(c, AIRFE_ERR_CUDA); }
  for (int p = 0; p < pairs; ++p) {
    int n = c->h_mcount[p];
    if (zero_mask && zero_mask[p]) n = 0;      // features0.cols() < 1 || features1.cols() < 1 -> return 0 (point_matcher.cc:53-55)
    if (n > match_cap) n = match_cap;
    n_match[p] = n;
    for (int k = 0; k < n; ++k) {
      idx0[(size_t)p * match_cap + k] = c->h_midx[((size_t)p * cap + k) * 2];
      idx1[(size_t)p * match_cap + k] = c->h_midx[((size_t)p * cap + k) * 2 + 1];
      score[(size_t)p * match_cap + k] = c->h_mscore[(size_t)p * cap + k];
    }
  }
  return AIRFE_OK;
}

static thread_local bool g_prenorm = false;

int airfe_match_batch(airfe_ctx* c, int matcher, int pairs, const float* feat0, const int* n0, const float* feat1, const int* n1,
                      int feat_cap, int* idx0, int* idx1, float* score, int match_cap, int* n_match);

int airfe_match_batch_prenormalized(airfe_ctx* c, int matcher, int pairs, const float* feat0, const int* n0, const float* feat1, const int* n1,
                                    int feat_cap, int* idx0, int* idx1, float* score, int match_cap, int* n_match) {
  g_prenorm = true;
  int rc = airfe_match_batch(c, matcher, pairs, feat0, n0, feat1, n1, feat_cap, idx0, idx1, score, match_cap, n_match);
  g_prenorm = false;
  return rc;
}

int airfe_match_batch(airfe_ctx* c, int matcher, int pairs, const float* feat0, const int* n0, const float* feat1, const int* n1,
                      int feat_cap, int* idx0, int* idx1, float* score, int match_cap, int* n_match) {
  if (!c || !feat0 || !feat1 || !n0 || !n1 || !idx0 || !idx1 || !score || !n_match) { set_error("null argument"); return fail(c, AIRFE_ERR_INVALID); }
  const bool sgm = matcher == AIRFE_MATCHER_SUPERGLUE;
  if ((sgm && !c->sg) || (!sgm && (matcher != AIRFE_MATCHER_LIGHTGLUE || !c->lg))) { set_error("matcher %d not enabled in this context", matcher); return fail(c, AIRFE_ERR_INVALID); }
  if (pairs < 1 || pairs > c->cfg.max_batch) { set_error("pairs %d outside [1,%d]", pairs, c->cfg.max_batch); return fail(c, AIRFE_ERR_INVALID); }
  cudaSetDevice(c->device);
  int nmax = 0;
  {
    for (int p = 0; p < pairs; ++p) { nmax = n0[p] > nmax ? n0[p] : nmax; nmax = n1[p] > nmax ? n1[p] : nmax; }
    if (sgm) c->sg_use = (c->sg_small && nmax <= 512) ? c->sg_small.get() : c->sg.get();
    else c->lg_use = (c->lg_small && nmax <= 512) ? c->lg_small.get() : c->lg.get();
  }
  const int cap = sgm ? c->sg_use->cap() : c->lg_use->cap();
  std::vector<int> zero(pairs, 0);
  for (int p = 0; p < pairs; ++p) {
    if (n0[p] < 0 || n1[p] < 0) { set_error("pair %d: negative keypoint count %d/%d", p, n0[p], n1[p]); return fail(c, AIRFE_ERR_INVALID); }
    if (n0[p] > cap || n1[p] > cap || n0[p] > feat_cap || n1[p] > feat_cap) { set_error("pair %d: %d/%d keypoints exceed capacity %d", p, n0[p], n1[p], cap); return fail(c, AIRFE_ERR_CAPACITY); }
    zero[p] = (n0[p] < 1 || n1[p] < 1);
    c->h_mn[2 * p] = n0[p];
    c->h_mn[2 * p + 1] = n1[p];
    memcpy(c->h_mfeat + (size_t)(2 * p) * kKpCap * 259, feat0 + (size_t)p * feat_cap * 259, (size_t)n0[p] * 259 * 4);
    memcpy(c->h_mfeat + (size_t)(2 * p + 1) * kKpCap * 259, feat1 + (size_t)p * feat_cap * 259, (size_t)n1[p] * 259 * 4);
  }
  cudaStream_t st = c->stream;
  const bool dense = getenv("AIRFE_DEBUG_DENSE") != nullptr;
  const bool prenorm = g_prenorm;
  // small calls (the class surface: one pair) run as one graph; the H2D sizes are then a bucket (rows rounded up to 64) so that a handful
  // of graphs covers every keypoint count -- rows beyond the count are never read by the kernels
  const bool one_shot = pairs <= 2 && !dense;
  int bucket = 0;
  if (one_shot) {
    bucket = (nmax + 63) / 64 * 64;
    if (nmax <= 415 && bucket > 415) bucket = 415;      // keep 400-keypoint sets inside the 13-column size class of the fused Sinkhorn kernel
    if (bucket > kKpCap) bucket = kKpCap;
  }
  auto body = [&]() -> bool {
    for (int s2 = 0; s2 < 2 * pairs; ++s2) {
      const int rows = one_shot ? bucket : c->h_mn[s2];
      if (rows > 0)
        cudaMemcpyAsync(c->d_mfeat + (size_t)s2 * kKpCap * 259, c->h_mfeat + (size_t)s2 * kKpCap * 259, (size_t)rows * 259 * 4, cudaMemcpyHostToDevice, st);
    }
    cudaMemcpyAsync(c->d_mn, c->h_mn, 8 * pairs, cudaMemcpyHostToDevice, st);
    if (sgm ? !c->sg_use->run(c->d_mfeat, c->d_mn, kKpCap, pairs, dense, st, prenorm, nullptr, one_shot ? bucket : nmax) : !c->lg_use->run(c->d_mfeat, c->d_mn, kKpCap, pairs, dense, st, prenorm))
      return false;
    if (one_shot) fetch_matches_enqueue(c, pairs, matcher);
    return cudaGetLastError() == cudaSuccess;
  };
  if (one_shot) {
    char key[128];
    snprintf(key, sizeof(key), "match:%d:%p:%d:%d:%d", matcher, sgm ? (void*)c->sg_use : (void*)c->lg_use, pairs, prenorm ? 1 : 0, bucket);
    if (!run_graphed(c, key, st, body)) return fail(c, AIRFE_ERR_CUDA);
    return fetch_matches(c, pairs, idx0, idx1, score, match_cap, n_match, zero.data(), matcher, true);
  }
  if (!body()) return fail(c, AIRFE_ERR_CUDA);
  return fetch_matches(c, pairs, idx0, idx1, score, match_cap, n_match, zero.data(), matcher);
}

int airfe_superglue_batch(airfe_ctx* c, int pairs, const float* feat0, const int* n0, const float* feat1, const int* n1, int feat_cap,
                          int prenormalized, int* indices0, int* indices1, float* mscores0, float* mscores1, int out_cap) {
  if (!c || !c->sg) { set_error("superglue not enabled in this context"); return fail(c, AIRFE_ERR_INVALID); }
  if (!indices0 || !indices1 || !mscores0 || !mscores1 || !n0 || !n1) { set_error("null argument"); return fail(c, AIRFE_ERR_INVALID); }
  if (pairs < 1 || pairs > c->cfg.max_batch) { set_error("pairs %d outside [1,%d]", pairs, c->cfg.max_batch); return fail(c, AIRFE_ERR_INVALID); }
  g_prenorm = prenormalized != 0;
  std::vector<int> i0((size_t)pairs * 1024), i1((size_t)pairs * 1024), nm(pairs);
  std::vector<float> sc((size_t)pairs * 1024);
  int rc = airfe_match_batch(c, AIRFE_MATCHER_SUPERGLUE, pairs, feat0, n0, feat1, n1, feat_cap, i0.data(), i1.data(), sc.data(), 1024, nm.data());
  g_prenorm = false;
  if (rc != AIRFE_OK) return fail(c, rc);
  const SuperGlueOutputs& o = c->sg_use->out();
  const int cap = c->sg_use->cap();
  const size_t PC = (size_t)pairs * cap;
  cudaMemcpyAsync(c->h_sgidx, o.idx0, PC * 4, cudaMemcpyDeviceToHost, c->stream);
  cudaMemcpyAsync(c->h_sgidx + PC, o.idx1, PC * 4, cudaMemcpyDeviceToHost, c->stream);
  cudaMemcpyAsync(c->h_sgms, o.ms0, PC * 4, cudaMemcpyDeviceToHost, c->stream);
  cudaMemcpyAsync(c->h_sgms + PC, o.ms1, PC * 4, cudaMemcpyDeviceToHost, c->stream);
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) { set_error("superglue readback failed"); return fail(c, AIRFE_ERR_CUDA); }
  for (int p = 0; p < pairs; ++p) {
    if (n0[p] > out_cap || n1[p] > out_cap) { set_error("output capacity %d too small", out_cap); return fail(c, AIRFE_ERR_CAPACITY); }
    for (int k = 0; k < n0[p]; ++k) { indices0[(size_t)p * out_cap + k] = c->h_sgidx[(size_t)p * cap + k]; mscores0[(size_t)p * out_cap + k] = c->h_sgms[(size_t)p * cap + k]; }
    for (int k = 0; k < n1[p]; ++k) { indices1[(size_t)p * out_cap + k] = c->h_sgidx[PC + (size_t)p * cap + k]; mscores1[(size_t)p * out_cap + k] = c->h_sgms[PC + (size_t)p * cap + k]; }
  }
  return AIRFE_OK;
}

int airfe_stereo_device(airfe_ctx* c, int net, int matcher, int pairs, const void* d_images, int w, int h, int stride, long long img_stride,
                        int lines, int junctions) {
  if (!c || !d_images) { set_error("null argument"); return fail(c, AIRFE_ERR_INVALID); }
  Detector* d = pick(c, net);
  if (!d) return fail(c, AIRFE_ERR_INVALID);
  const bool sgm = matcher == AIRFE_MATCHER_SUPERGLUE;
  if ((sgm && !c->sg) || (!sgm && (matcher != AIRFE_MATCHER_LIGHTGLUE || !c->lg))) { set_error("matcher %d not enabled in this context", matcher); return fail(c, AIRFE_ERR_INVALID); }
  if (pairs < 1 || pairs > c->cfg.max_batch) { set_error("pairs %d outside [1,%d]", pairs, c->cfg.max_batch); return fail(c, AIRFE_ERR_INVALID); }
  cudaSetDevice(c->device);
  RemapMaps rm; bool rm_err;
  const RemapMaps* rmp = remap_for(c, w, h, true, &rm, &rm_err);
  if (rm_err) return fail(c, AIRFE_ERR_INVALID);
  if (!d->run((const uint8_t*)d_images, 2 * pairs, w, h, stride, img_stride, lines != 0, junctions != 0, c->stream, rmp)) return fail(c, AIRFE_ERR_CUDA);
  const DetectOutputs& o = d->out();
  if (sgm) c->sg_use = c->sg.get(); else c->lg_use = c->lg.get();
  if (sgm ? !c->sg->run(o.feat, o.n_feat, kKpCap, pairs, false, c->stream, false, nullptr, c->cfg.max_keypoints) : !c->lg->run(o.feat, o.n_feat, kKpCap, pairs, false, c->stream)) return fail(c, AIRFE_ERR_CUDA);
  c->last_net = net; c->last_matcher = matcher; c->last_pairs = pairs; c->last_w = w; c->last_h = h;
  return AIRFE_OK;
}

long long airfe_profile_stereo(airfe_ctx* c, int net, int matcher, int pairs, const void* d_images, int w, int h, int stride, long long img_stride,
                               int lines, int junctions, char* out, long long cap) {
  if (!c || !out || cap < 1) { set_error("null argument"); return fail(c, AIRFE_ERR_INVALID); }
  cudaStreamSynchronize(c->stream);
  profiler().begin();
  int rc = airfe_stereo_device(c, net, matcher, pairs, d_images, w, h, stride, img_stride, lines, junctions);
  profiler().finish();
  if (rc != AIRFE_OK) return fail(c, rc);
  // FLOPs of the ops whose row counts live on the device were accumulated at slot capacity; scale them to the rows actually processed
  Detector* d = pick(c, net);
  const bool sgm = matcher == AIRFE_MATCHER_SUPERGLUE;
  const int S = 2 * pairs, mcap = sgm ? c->sg->cap() : c->lg->cap();
  std::vector<int> hn(S, 0), hu(S, 0);
  cudaMemcpy(hn.data(), sgm ? c->sg->counts() : c->lg->counts(), 4 * S, cudaMemcpyDeviceToHost);
  if (lines && d->n_unique()) cudaMemcpy(hu.data(), d->n_unique(), 4 * S, cudaMemcpyDeviceToHost);
  double f_rows = 0, f_self = 0, f_cross = 0, f_sim = 0, f_lines = 0;
  for (int s2 = 0; s2 < S; ++s2) {
    const double n = hn[s2] < mcap ? hn[s2] : mcap, m = hn[s2 ^ 1] < mcap ? hn[s2 ^ 1] : mcap;
    f_rows += n; f_self += n * n; f_cross += n * m; f_lines += hu[s2] < kLineCap ? hu[s2] : kLineCap;
    if (!(s2 & 1)) f_sim += n * m;
  }
  f_rows /= (double)S * mcap; f_self /= (double)S * mcap * mcap; f_cross /= (double)S * mcap * mcap; f_sim /= (double)pairs * mcap * mcap;
  f_lines /= (double)S * kLineCap;
  const double factor[6] = {1.0, f_rows, f_self, f_cross, f_sim, f_lines};
  long long off = 0;
  for (auto& r : profiler().recs) {
    char line[256];
    int n = snprintf(line, sizeof(line), "%s\t%.0f\t%.6f\n", r.name.c_str(), r.flops * factor[r.kind >= 0 && r.kind < 6 ? r.kind : 0], r.ms);
    if (off + n >= cap) break;
    memcpy(out + off, line, n);
    off += n;
  }
  if (off < cap) out[off] = 0;
  return off;
}

int airfe_stereo_cost(airfe_ctx* c, int net, int matcher, int pairs, int lines, double* tc_flops, int* launches) {
  if (!c) { set_error("null argument"); return AIRFE_ERR_INVALID; }
  Detector* d = pick(c, net);
  const bool sgm = matcher == AIRFE_MATCHER_SUPERGLUE;
  if (!d || (sgm ? !c->sg : !c->lg)) { set_error("networks not enabled"); return fail(c, AIRFE_ERR_INVALID); }
  if (tc_flops) *tc_flops = d->tc_flops(2 * pairs, lines != 0) + (sgm ? c->sg->tc_flops(pairs) : c->lg->tc_flops(pairs));
  if (launches) *launches = d->launches(2 * pairs, lines != 0) + (sgm ? c->sg->launches(pairs) : c->lg->launches(pairs)) + (lines ? 13 : 4);    // + resize, keypoint selection (2), descriptor sampling; lines: + 9 decode / association / LOI kernels
  return AIRFE_OK;
}

int airfe_detect_match_stereo_batch(airfe_ctx* c, int net, int matcher, int pairs, const uint8_t* left, const uint8_t* right, int w, int h,
                                    int stride, long long img_stride, float* feat, int feat_cap, int* n_feat, double* lines, int line_cap,
                                    int* n_lines, float* junc, int junc_cap, int* n_junc, int* idx0, int* idx1, float* score, int match_cap,
                                    int* n_match) {
  if (!c || !left || !right || !feat || !n_feat || !idx0 || !idx1 || !score || !n_match) { set_error("null argument"); return fail(c, AIRFE_ERR_INVALID); }
  if (w <= 1 || h <= 1 || stride < w) { set_error("empty image"); return fail(c, AIRFE_ERR_INVALID); }
  Detector* d = pick(c, net);
  if (!d) return fail(c, AIRFE_ERR_INVALID);
  const bool sgm = matcher == AIRFE_MATCHER_SUPERGLUE;
  if ((sgm && !c->sg) || (!sgm && (matcher != AIRFE_MATCHER_LIGHTGLUE || !c->lg))) { set_error("matcher %d not enabled in this context", matcher); return fail(c, AIRFE_ERR_INVALID); }
  if (pairs < 1 || pairs > c->cfg.max_batch) { set_error("pairs %d outside [1,%d]", pairs, c->cfg.max_batch); return fail(c, AIRFE_ERR_INVALID); }
  if ((lines || junc) && net != AIRFE_NET_PLNET) { set_error("lines / junctions need the PLNet network"); return fail(c, AIRFE_ERR_INVALID); }
  if (junc && !lines) { set_error("junction detection needs line detection"); return fail(c, AIRFE_ERR_INVALID); }
  if ((lines && !n_lines) || (junc && !n_junc)) { set_error("n_lines / n_junc must be given with their buffers"); return fail(c, AIRFE_ERR_INVALID); }
  cudaSetDevice(c->device);
  const size_t one = (size_t)h * stride, need = one * 2 * pairs;
  if (need > c->h_img_bytes) {
    if (!grow_image_staging(c, need)) return fail(c, AIRFE_ERR_CUDA);
  }
  cudaStream_t st = c->stream;
  if (is_pinned(left) && is_pinned(right)) {        // caller's frames are in pinned memory: DMA straight from them
    for (int p = 0; p < pairs; ++p) {
      cudaMemcpyAsync(c->d_img + one * (2 * p), left + (size_t)img_stride * p, one, cudaMemcpyHostToDevice, st);
      cudaMemcpyAsync(c->d_img + one * (2 * p + 1), right + (size_t)img_stride * p, one, cudaMemcpyHostToDevice, st);
    }
  } else {
    for (int p = 0; p < pairs; ++p) {
      memcpy(c->h_img + one * (2 * p), left + (size_t)img_stride * p, one);
      memcpy(c->h_img + one * (2 * p + 1), right + (size_t)img_stride * p, one);
    }
    cudaMemcpyAsync(c->d_img, c->h_img, need, cudaMemcpyHostToDevice, st);
  }
  RemapMaps rm; bool rm_err;
  const RemapMaps* rmp = remap_for(c, w, h, true, &rm, &rm_err);
  if (rm_err) return fail(c, AIRFE_ERR_INVALID);
  if (!d->run(c->d_img, 2 * pairs, w, h, stride, (long long)one, lines != nullptr, junc != nullptr, st, rmp)) return fail(c, AIRFE_ERR_CUDA);
  const DetectOutputs& o = d->out();
  // detector results go home on a second stream while the matcher runs on the first (13 MB of descriptors per 16 pairs)
  cudaStream_t cs = c->copy_stream;
  cudaEventRecord(c->ev_det, st);
  cudaStreamWaitEvent(cs, c->ev_det, 0);
  const int B = 2 * c->cfg.max_batch, S = 2 * pairs;
  int* hc = c->h_counts;
  cudaMemcpyAsync(hc, o.n_feat, 4 * S, cudaMemcpyDeviceToHost, cs);
  if (lines) cudaMemcpyAsync(hc + B, o.n_lines, 4 * S, cudaMemcpyDeviceToHost, cs);
  if (junc) cudaMemcpyAsync(hc + 2 * B, o.n_junc, 4 * S, cudaMemcpyDeviceToHost, cs);
  const int kmax = c->cfg.max_keypoints;
  const bool feat_direct = is_pinned(feat) && feat_cap >= kmax;   // pinned caller buffer: D2H lands in it directly (columns beyond n_feat are scratch)
  for (int i = 0; i < S; ++i)
    cudaMemcpyAsync(feat_direct ? feat + (size_t)i * feat_cap * 259 : c->h_feat + (size_t)i * kKpCap * 259, o.feat + (size_t)i * kKpCap * 259,
                    (size_t)kmax * 259 * 4, cudaMemcpyDeviceToHost, cs);
  const bool junc_direct = junc && is_pinned(junc) && junc_cap >= kJunc;
  if (junc_direct)
    for (int p = 0; p < pairs; ++p)
      cudaMemcpyAsync(junc + (size_t)p * junc_cap * 259, o.junc + (size_t)(2 * p) * kKpCap * 259, (size_t)kJunc * 259 * 4, cudaMemcpyDeviceToHost, cs);
  cudaEventRecord(c->ev_copy, cs);
  if (sgm) c->sg_use = c->sg.get(); else c->lg_use = c->lg.get();     // the detector's feature sets are bounded by max_keypoints = this instance's sizing
  if (sgm ? !c->sg->run(o.feat, o.n_feat, kKpCap, pairs, false, st, false, nullptr, c->cfg.max_keypoints) : !c->lg->run(o.feat, o.n_feat, kKpCap, pairs, false, st)) {
    cudaStreamSynchronize(cs);
    return fail(c, AIRFE_ERR_CUDA);
  }
  // the detector results land while the matcher is still running: size the line / junction copies from the counts and queue them too
  if (cudaEventSynchronize(c->ev_copy) != cudaSuccess) { set_error("detector readback failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(c, AIRFE_ERR_CUDA); }
  for (int i = 0; i < S; ++i) {
    const int n = hc[i] < feat_cap ? hc[i] : feat_cap;
    n_feat[i] = n;
    if (!feat_direct) memcpy(feat + (size_t)i * feat_cap * 259, c->h_feat + (size_t)i * kKpCap * 259, (size_t)n * 259 * 4);
  }
  if (lines) {
    for (int i = 0; i < S; ++i) {
      const int n = hc[B + i] < line_cap ? hc[B + i] : line_cap;
      n_lines[i] = n;
      if (n) cudaMemcpyAsync(c->h_lines + (size_t)i * kLineCap * 4, o.lines + (size_t)i * kLineCap * 4, (size_t)n * 16, cudaMemcpyDeviceToHost, cs);
    }
    if (junc)
      for (int p = 0; p < pairs; ++p) {
        const int n = hc[2 * B + 2 * p] < junc_cap ? hc[2 * B + 2 * p] : junc_cap;
        n_junc[p] = n;
        if (n && !junc_direct) cudaMemcpyAsync(c->h_junc + (size_t)p * kKpCap * 259, o.junc + (size_t)(2 * p) * kKpCap * 259, (size_t)n * 259 * 4, cudaMemcpyDeviceToHost, cs);
      }
    if (cudaStreamSynchronize(cs) != cudaSuccess) { set_error("line readback failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(c, AIRFE_ERR_CUDA); }
    const double ws = (double)((float)w / 512.f), hs = (double)((float)h / 512.f);
    for (int i = 0; i < S; ++i) {
      const float* l = c->h_lines + (size_t)i * kLineCap * 4;
      double* dl = lines + (size_t)i * line_cap * 4;
      for (int k = 0; k < n_lines[i]; ++k) {
        dl[k * 4 + 0] = (double)l[k * 4 + 0] * ws; dl[k * 4 + 1] = (double)l[k * 4 + 1] * hs;
        dl[k * 4 + 2] = (double)l[k * 4 + 2] * ws; dl[k * 4 + 3] = (double)l[k * 4 + 3] * hs;
      }
    }
    if (junc && !junc_direct)
      for (int p = 0; p < pairs; ++p) memcpy(junc + (size_t)p * junc_cap * 259, c->h_junc + (size_t)p * kKpCap * 259, (size_t)n_junc[p] * 259 * 4);
  }
  int rc = fetch_matches(c, pairs, idx0, idx1, score, match_cap, n_match, nullptr, matcher);   // synchronises the compute stream
  if (rc != AIRFE_OK) return fail(c, rc);
  c->last_net = net; c->last_matcher = matcher; c->last_pairs = pairs; c->last_w = w; c->last_h = h;
  for (int p = 0; p < pairs; ++p)
    if (hc[2 * p] < 1 || hc[2 * p + 1] < 1) n_match[p] = 0;
  return AIRFE_OK;
}

// ---- BoW quantisation (Database::FrameToBow) -------------------------------------------------------------------------------------------
int airfe_bow_load(airfe_ctx* c, const char* path) {
  if (!c || !path) { set_error("null argument"); return fail(c, AIRFE_ERR_INVALID); }
  cudaSetDevice(c->device);
  std::unique_ptr<BowVocabulary> v(new BowVocabulary);
  const size_t len = strlen(path);
  const bool afw = len > 4 && !strcmp(path + len - 4, ".afw");
  if (!(afw ? v->load_afw(path) : v->load_boost_archive(path))) return fail(c, AIRFE_ERR_IO);
  c->voc = std::move(v);
  return AIRFE_OK;
}
int airfe_bow_transform(airfe_ctx* c, const float* feat, int n, unsigned int* word_of_feature, unsigned int* bow_ids, double* bow_vals, int* n_bow) {
  if (!c || !feat || !word_of_feature || !bow_ids || !bow_vals || !n_bow || n < 0) { set_error("null argument"); return fail(c, AIRFE_ERR_INVALID); }
  if (!c->voc) { set_error("no vocabulary: call airfe_bow_load first"); return fail(c, AIRFE_ERR_INVALID); }
  *n_bow = 0;
  if (n == 0) return AIRFE_OK;                                        // database.cc:60
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  BowVocabulary& v = *c->voc;
  if (n > v.leaf_cap) {
    cudaStreamSynchronize(st);
    if (v.d_leaf) cudaFree(v.d_leaf);
    v.d_leaf = nullptr; v.leaf_cap = 0;
    if (cudaMalloc(&v.d_leaf, (size_t)n * 4) != cudaSuccess) { cudaGetLastError(); set_error("bow: allocation failed"); return fail(c, AIRFE_ERR_CUDA); }
    v.leaf_cap = n;
  }
  cudaPointerAttributes pa;
  const bool on_device = cudaPointerGetAttributes(&pa, feat) == cudaSuccess && pa.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  const float* df = feat;
  if (!on_device) {
    if ((size_t)n > c->d_bowfeat_rows) {
      cudaStreamSynchronize(st);
      if (c->d_bowfeat) cudaFree(c->d_bowfeat);
      c->d_bowfeat = nullptr; c->d_bowfeat_rows = 0;
      if (cudaMalloc(&c->d_bowfeat, (size_t)n * 259 * 4) != cudaSuccess) { cudaGetLastError(); set_error("bow: allocation failed"); return fail(c, AIRFE_ERR_CUDA); }
      c->d_bowfeat_rows = n;
    }
    cudaMemcpyAsync(c->d_bowfeat, feat, (size_t)n * 259 * 4, cudaMemcpyHostToDevice, st);
    df = c->d_bowfeat;
  }
  if (!bow_transform_device(v, df, 259, n, v.d_leaf, st)) return fail(c, AIRFE_ERR_CUDA);
  std::vector<int> leaf(n);
  cudaMemcpyAsync(leaf.data(), v.d_leaf, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("bow transform failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(c, AIRFE_ERR_CUDA); }
  // database.cc:70-88 + BowVector::addWeight / normalize(L1): the reference's own double arithmetic, in feature order then map order
  std::map<unsigned int, double> bow;
  for (int i = 0; i < n; ++i) {
    const int node = leaf[i];
    const double w = (node >= 0 && node < v.n_nodes) ? v.weight[node] : 0.0;
    if (w > 0) {
      const unsigned int id = (unsigned int)v.word_id[node];
      auto it = bow.lower_bound(id);
      if (it != bow.end() && it->first == id) it->second += w; else bow.insert(it, std::make_pair(id, w));
      word_of_feature[i] = id;
    } else {
      word_of_feature[i] = 0xFFFFFFFFu;                               // UINT_MAX
    }
  }
  double norm = 0.0;
  for (auto& kv : bow) norm += fabs(kv.second);
  int k = 0;
  for (auto& kv : bow) { bow_ids[k] = kv.first; bow_vals[k] = norm > 0.0 ? kv.second / norm : kv.second; ++k; }
  *n_bow = k;
  return AIRFE_OK;
}

// ---- point <-> line association + stereo line matching on the last stereo call's device-resident results ----------------------------
int airfe_stereo_line_assoc(airfe_ctx* c, int pairs, double min_x_diff, double max_x_diff, double max_y_diff, int line_cap, int rel_cap,
                            int* rel_n, int* rel_idx, float* rel_dist, int* line_matches) {
  constexpr int kML = 256, kRC = 32;
  if (!c || !rel_n || !rel_idx || !rel_dist || !line_matches || line_cap < 1 || rel_cap < 1) { set_error("null argument"); return fail(c, AIRFE_ERR_INVALID); }
  if (c->last_net != AIRFE_NET_PLNET || pairs < 1 || pairs != c->last_pairs) { set_error("line association follows a PLNet stereo call with the same number of pairs"); return fail(c, AIRFE_ERR_INVALID); }
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  if (pairs > c->la_pairs) {
    cudaStreamSynchronize(st);
    int** ip[] = {&c->la_rel_n, &c->la_rel_idx, &c->la_cnt, &c->la_row, &c->la_lm, &c->la_ovf};
    for (auto q : ip) { if (*q) cudaFree(*q); *q = nullptr; }
    if (c->la_rel_dist) cudaFree(c->la_rel_dist);
    c->la_rel_dist = nullptr; c->la_pairs = 0;
    const size_t P = (size_t)pairs;
    if (cudaMalloc(&c->la_rel_n, 2 * P * kML * 4) != cudaSuccess || cudaMalloc(&c->la_rel_idx, 2 * P * kML * kRC * 4) != cudaSuccess ||
        cudaMalloc(&c->la_rel_dist, 2 * P * kML * kRC * 4) != cudaSuccess || cudaMalloc(&c->la_cnt, P * kML * kML * 4) != cudaSuccess ||
        cudaMalloc(&c->la_row, P * kML * 4) != cudaSuccess || cudaMalloc(&c->la_lm, P * kML * 4) != cudaSuccess || cudaMalloc(&c->la_ovf, 4) != cudaSuccess) {
      cudaGetLastError();
      set_error("line association scratch allocation failed");
      return fail(c, AIRFE_ERR_CUDA);
    }
    c->la_pairs = pairs;
  }
  const DetectOutputs& o = c->pl->out();
  const bool sgm = c->last_matcher == AIRFE_MATCHER_SUPERGLUE;
  const int* m_idx = sgm ? c->sg_use->out().m_idx : c->lg_use->out().idx;
  const int* m_count = sgm ? c->sg_use->out().m_count : c->lg_use->out().count;
  const int m_cap = sgm ? c->sg_use->cap() : c->lg_use->cap();
  const double ws = (double)((float)c->last_w / 512.f), hs = (double)((float)c->last_h / 512.f);
  launch_line_assoc(o.lines, o.n_lines, kLineCap, o.feat, o.n_feat, kKpCap, ws, hs, m_idx, m_count, m_cap, pairs, min_x_diff, max_x_diff, max_y_diff, kML, kRC,
                    c->la_rel_n, c->la_rel_idx, c->la_rel_dist, c->la_cnt, c->la_row, c->la_lm, c->la_ovf, st);
  const int S = 2 * pairs;
  std::vector<int> h_nl(S), h_rn((size_t)S * kML), h_lm((size_t)pairs * kML);
  int ovf = 0;
  cudaMemcpyAsync(h_nl.data(), o.n_lines, 4 * S, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(h_rn.data(), c->la_rel_n, (size_t)S * kML * 4, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(h_lm.data(), c->la_lm, (size_t)pairs * kML * 4, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(&ovf, c->la_ovf, 4, cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("line association failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(c, AIRFE_ERR_CUDA); }
  // overflow first: with the flag set a per-line count may exceed the kRC entries the device tables (and h_idx / h_dist below) hold per line
  if (ovf) { set_error("line association overflow (more than %d points on a line or 8 lines through a point)", kRC); return fail(c, AIRFE_ERR_CAPACITY); }
  std::vector<int> h_idx((size_t)kML * kRC);
  std::vector<float> h_dist((size_t)kML * kRC);
  for (int s2 = 0; s2 < S; ++s2) {
    const int nl = h_nl[s2];
    if (nl > kML || nl > line_cap) { set_error("image %d: %d lines exceed the association capacity %d", s2, nl, kML < line_cap ? kML : line_cap); return fail(c, AIRFE_ERR_CAPACITY); }
    if (nl) {
      cudaMemcpy(h_idx.data(), c->la_rel_idx + (size_t)s2 * kML * kRC, (size_t)nl * kRC * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(h_dist.data(), c->la_rel_dist + (size_t)s2 * kML * kRC, (size_t)nl * kRC * 4, cudaMemcpyDeviceToHost);
    }
    for (int i = 0; i < line_cap; ++i) rel_n[(size_t)s2 * line_cap + i] = 0;
    for (int i = 0; i < nl; ++i) {
      const int n = h_rn[(size_t)s2 * kML + i];
      if (n > rel_cap || n > kRC || n < 0) { set_error("line %d of image %d has %d points, rel_cap is %d", i, s2, n, rel_cap < kRC ? rel_cap : kRC); return fail(c, AIRFE_ERR_CAPACITY); }
      rel_n[(size_t)s2 * line_cap + i] = n;
      for (int k = 0; k < n; ++k) {
        rel_idx[((size_t)s2 * line_cap + i) * rel_cap + k] = h_idx[(size_t)i * kRC + k];
        rel_dist[((size_t)s2 * line_cap + i) * rel_cap + k] = h_dist[(size_t)i * kRC + k];
      }
    }
  }
  for (int p = 0; p < pairs; ++p)
    for (int i = 0; i < line_cap; ++i) line_matches[(size_t)p * line_cap + i] = i < kML ? h_lm[(size_t)p * kML + i] : -1;
  return AIRFE_OK;
}

// ---- rectification maps (Camera::UndistortImage on the device) --------------------------------------------------------------------
int airfe_set_rectify_maps(airfe_ctx* c, int side, const float* map_x, const float* map_y, int w, int h) {
  if (!c || side < 0 || side > 1 || !map_x || !map_y || w < 2 || h < 2) { set_error("set_rectify_maps: bad arguments"); return fail(c, AIRFE_ERR_INVALID); }
  if (c->remap_w && (w != c->remap_w || h != c->remap_h) && c->d_rxy[1 - side]) { set_error("both cameras must share one image size"); return fail(c, AIRFE_ERR_INVALID); }
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  drop_graphs(c);
  const size_t n = (size_t)w * h;
  std::vector<short> xy(2 * n);
  std::vector<unsigned short> a(n);
  build_remap_host(map_x, map_y, w, h, xy.data(), a.data());
  if (c->d_rxy[side]) cudaFree(c->d_rxy[side]);
  if (c->d_ra[side]) cudaFree(c->d_ra[side]);
  c->d_rxy[side] = nullptr; c->d_ra[side] = nullptr;
  if (cudaMalloc(&c->d_rxy[side], 4 * n) != cudaSuccess || cudaMalloc(&c->d_ra[side], 2 * n) != cudaSuccess ||
      cudaMemcpy(c->d_rxy[side], xy.data(), 4 * n, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(c->d_ra[side], a.data(), 2 * n, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    set_error("rectification map upload failed");
    return fail(c, AIRFE_ERR_CUDA);
  }
  c->remap.side[side] = RemapSide{c->d_rxy[side], c->d_ra[side]};
  c->remap_w = w; c->remap_h = h;
  return AIRFE_OK;
}
int airfe_set_rectify(airfe_ctx* c, int mode) {
  if (!c || mode < 0 || mode > 2) { set_error("set_rectify: mode must be 0, 1 or 2"); return fail(c, AIRFE_ERR_INVALID); }
  if (mode && !c->d_rxy[0]) { set_error("set_rectify: no left / mono camera map installed"); return fail(c, AIRFE_ERR_INVALID); }
  if (mode == 2 && !c->d_rxy[1]) { set_error("set_rectify: mode 2 needs the right-camera map"); return fail(c, AIRFE_ERR_INVALID); }
  c->remap.mode = mode;
  return AIRFE_OK;
}
int airfe_undistort(airfe_ctx* c, int side, const uint8_t* raw, int w, int h, int stride, uint8_t* rect, int rect_stride) {
  if (!c || side < 0 || side > 1 || !raw || !rect || stride < w || rect_stride < w) { set_error("undistort: bad arguments"); return fail(c, AIRFE_ERR_INVALID); }
  if (!c->d_rxy[side] || w != c->remap_w || h != c->remap_h) { set_error("undistort: no %dx%d map for side %d", w, h, side); return fail(c, AIRFE_ERR_INVALID); }
  cudaSetDevice(c->device);
  const size_t need = (size_t)h * stride;
  if (need > c->h_img_bytes && !grow_image_staging(c, need)) return fail(c, AIRFE_ERR_CUDA);
  if ((size_t)w * h > c->d_rect_bytes) {
    if (c->d_rect) cudaFree(c->d_rect);
    c->d_rect = nullptr; c->d_rect_bytes = 0;
    if (cudaMalloc(&c->d_rect, (size_t)w * h) != cudaSuccess) { cudaGetLastError(); c->d_rect = nullptr; set_error("undistort: allocation failed"); return fail(c, AIRFE_ERR_CUDA); }
    c->d_rect_bytes = (size_t)w * h;
  }
  memcpy(c->h_img, raw, need);
  cudaStream_t st = c->stream;
  cudaMemcpyAsync(c->d_img, c->h_img, need, cudaMemcpyHostToDevice, st);
  RemapMaps m = c->remap;
  m.side[0] = c->remap.side[side];
  m.mode = 1;
  launch_remap_u8(c->d_img, w, h, stride, 0, 1, m, c->d_rect, st);
  if (cudaMemcpy2DAsync(rect, rect_stride, c->d_rect, w, w, h, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
    set_error("undistort failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(c, AIRFE_ERR_CUDA);
  }
  return AIRFE_OK;
}

// ---- device-resident keyframe cache + batched candidate matching -----------------------------------------------------------------
int airfe_kf_reserve(airfe_ctx* c, int n_keyframes, int feat_cap) {
  if (!c || n_keyframes < 1 || feat_cap < 1 || feat_cap > kKpCap) { set_error("kf_reserve: bad arguments"); return fail(c, AIRFE_ERR_INVALID); }
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->kf_feat) cudaFree(c->kf_feat);
  delete[] c->kf_n;
  c->kf_feat = nullptr; c->kf_n = nullptr; c->kf_slots = c->kf_cap = 0;
  const size_t bytes = (size_t)n_keyframes * feat_cap * 259 * 4;
  if (cudaMalloc(&c->kf_feat, bytes) != cudaSuccess) { cudaGetLastError(); c->kf_feat = nullptr; set_error("keyframe cache allocation (%zu bytes) failed", bytes); return fail(c, AIRFE_ERR_CUDA); }
  c->kf_n = new int[n_keyframes]();
  c->kf_slots = n_keyframes; c->kf_cap = feat_cap;
  return AIRFE_OK;
}
int airfe_kf_size(const airfe_ctx* c) { return c ? c->kf_slots : 0; }
int airfe_kf_put(airfe_ctx* c, int slot, const float* feat, int n) {
  if (!c || !c->kf_feat || slot < 0 || slot >= c->kf_slots || n < 0 || n > c->kf_cap || (!feat && n)) { set_error("kf_put: bad slot / count"); return fail(c, AIRFE_ERR_INVALID); }
  cudaSetDevice(c->device);
  // cudaMemcpyDefault: the source may be host (map load) or device memory (features that were just detected)
  if (n && cudaMemcpyAsync(c->kf_feat + (size_t)slot * c->kf_cap * 259, feat, (size_t)n * 259 * 4, cudaMemcpyDefault, c->stream) != cudaSuccess) {
    set_error("kf_put: copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(c, AIRFE_ERR_CUDA);
  }
  if (!is_pinned(feat) && n) cudaStreamSynchronize(c->stream);   // pageable / device sources may be reused by the caller right away
  c->kf_n[slot] = n;
  return AIRFE_OK;
}

static bool grow_reloc_tables(airfe_ctx* c, int n_jobs, int cap) {
  if (n_jobs <= c->rj_cap && cap <= c->rj_kp_cap) return true;     // the match tables are [jobs][cap]: a matcher with a larger capacity needs them regrown too
  if (n_jobs < c->rj_cap) n_jobs = c->rj_cap;
  if (cap < c->rj_kp_cap) cap = c->rj_kp_cap;
  cudaStreamSynchronize(c->stream);
  if (c->h_rptr) cudaFreeHost(c->h_rptr);
  if (c->d_rptr) cudaFree(c->d_rptr);
  if (c->h_rn) cudaFreeHost(c->h_rn);
  if (c->d_rn) cudaFree(c->d_rn);
  if (c->d_rcount) cudaFree(c->d_rcount);
  if (c->d_ridx) cudaFree(c->d_ridx);
  if (c->d_rscore) cudaFree(c->d_rscore);
  if (c->h_rcount) cudaFreeHost(c->h_rcount);
  if (c->h_ridx) cudaFreeHost(c->h_ridx);
  if (c->h_rscore) cudaFreeHost(c->h_rscore);
  c->h_rptr = nullptr; c->d_rptr = nullptr; c->h_rn = nullptr; c->d_rn = nullptr; c->d_rcount = nullptr; c->d_ridx = nullptr; c->d_rscore = nullptr;
  c->h_rcount = nullptr; c->h_ridx = nullptr; c->h_rscore = nullptr; c->rj_cap = 0; c->rj_kp_cap = 0;
  const size_t J = (size_t)n_jobs;
  if (cudaMallocHost(&c->h_rptr, 2 * J * sizeof(float*)) != cudaSuccess || cudaMalloc(&c->d_rptr, 2 * J * sizeof(float*)) != cudaSuccess ||
      cudaMallocHost(&c->h_rn, 2 * J * 4) != cudaSuccess || cudaMalloc(&c->d_rn, 2 * J * 4) != cudaSuccess || cudaMalloc(&c->d_rcount, J * 4) != cudaSuccess ||
      cudaMalloc(&c->d_ridx, J * cap * 8) != cudaSuccess || cudaMalloc(&c->d_rscore, J * cap * 4) != cudaSuccess || cudaMallocHost(&c->h_rcount, J * 4) != cudaSuccess ||
      cudaMallocHost(&c->h_ridx, J * cap * 8) != cudaSuccess || cudaMallocHost(&c->h_rscore, J * cap * 4) != cudaSuccess) {
    cudaGetLastError();
    set_error("reloc job tables (%d jobs) allocation failed", n_jobs);
    return false;
  }
  c->rj_cap = n_jobs; c->rj_kp_cap = cap;
  return true;
}

int airfe_reloc_match(airfe_ctx* c, int matcher, const float* query_feat, const int* query_n, int n_queries, int feat_cap, int n_jobs,
                      const int* job_query, const int* job_kf, int* n_match, int* idx0, int* idx1, float* score, int match_cap) {
  if (!c || !query_feat || !query_n || !job_query || !job_kf || !n_match || n_queries < 1 || n_jobs < 0) { set_error("null argument"); return fail(c, AIRFE_ERR_INVALID); }
  if (idx0 && (!idx1 || !score || match_cap < 1)) { set_error("idx0 needs idx1, score and match_cap"); return fail(c, AIRFE_ERR_INVALID); }
  if (!c->kf_feat) { set_error("no keyframe cache: call airfe_kf_reserve / airfe_kf_put first"); return fail(c, AIRFE_ERR_INVALID); }
  const bool sgm = matcher == AIRFE_MATCHER_SUPERGLUE;
  if ((sgm && !c->sg) || (!sgm && (matcher != AIRFE_MATCHER_LIGHTGLUE || !c->lg))) { set_error("matcher %d not enabled in this context", matcher); return fail(c, AIRFE_ERR_INVALID); }
  if (n_jobs == 0) return AIRFE_OK;
  cudaSetDevice(c->device);
  const int cap = sgm ? c->sg->cap() : c->lg->cap();
  for (int q = 0; q < n_queries; ++q)
    if (query_n[q] < 0 || query_n[q] > cap || query_n[q] > feat_cap) { set_error("query %d: %d keypoints exceed capacity %d", q, query_n[q], cap < feat_cap ? cap : feat_cap); return fail(c, AIRFE_ERR_CAPACITY); }
  if (!grow_reloc_tables(c, n_jobs, cap)) return fail(c, AIRFE_ERR_CUDA);
  cudaStream_t st = c->stream;
  // queries: used in place when they are already on the device (NCCL all-gather output), else staged once
  cudaPointerAttributes pa;
  bool on_device = cudaPointerGetAttributes(&pa, query_feat) == cudaSuccess && pa.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  const float* dq = query_feat;
  if (!on_device) {
    const size_t need = (size_t)n_queries * feat_cap * 259 * 4;
    if (need > c->d_qfeat_bytes) {
      cudaStreamSynchronize(st);
      if (c->d_qfeat) cudaFree(c->d_qfeat);
      c->d_qfeat = nullptr; c->d_qfeat_bytes = 0;
      if (cudaMalloc(&c->d_qfeat, need) != cudaSuccess) { cudaGetLastError(); c->d_qfeat = nullptr; set_error("query staging allocation failed"); return fail(c, AIRFE_ERR_CUDA); }
      c->d_qfeat_bytes = need;
    }
    for (int q = 0; q < n_queries; ++q)
      if (query_n[q]) cudaMemcpyAsync(c->d_qfeat + (size_t)q * feat_cap * 259, query_feat + (size_t)q * feat_cap * 259, (size_t)query_n[q] * 259 * 4, cudaMemcpyHostToDevice, st);
    if (!is_pinned(query_feat)) cudaStreamSynchronize(st);
    dq = c->d_qfeat;
  }
  for (int j = 0; j < n_jobs; ++j) {
    const int q = job_query[j], k = job_kf[j];
    if (q < 0 || q >= n_queries || k < 0 || k >= c->kf_slots) { set_error("job %d: query %d / keyframe %d out of range", j, q, k); return fail(c, AIRFE_ERR_INVALID); }
    if (c->kf_n[k] > cap) { set_error("keyframe %d: %d keypoints exceed capacity %d", k, c->kf_n[k], cap); return fail(c, AIRFE_ERR_CAPACITY); }
    c->h_rptr[2 * j] = dq + (size_t)q * feat_cap * 259;
    c->h_rptr[2 * j + 1] = c->kf_feat + (size_t)k * c->kf_cap * 259;
    c->h_rn[2 * j] = query_n[q];
    c->h_rn[2 * j + 1] = c->kf_n[k];
  }
  cudaMemcpyAsync(c->d_rptr, c->h_rptr, (size_t)2 * n_jobs * sizeof(float*), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(c->d_rn, c->h_rn, (size_t)2 * n_jobs * 4, cudaMemcpyHostToDevice, st);
  const int B = c->cfg.max_batch;
  for (int j0 = 0; j0 < n_jobs; j0 += B) {
    const int P = n_jobs - j0 < B ? n_jobs - j0 : B;
    if (sgm ? !c->sg->run(nullptr, c->d_rn + 2 * j0, 0, P, false, st, false, c->d_rptr + 2 * j0, -1)
            : !c->lg->run(nullptr, c->d_rn + 2 * j0, 0, P, false, st, false, c->d_rptr + 2 * j0))
      return fail(c, AIRFE_ERR_CUDA);
    const int* d_count = sgm ? c->sg->out().m_count : c->lg->out().count;
    cudaMemcpyAsync(c->d_rcount + j0, d_count, (size_t)P * 4, cudaMemcpyDeviceToDevice, st);
    if (idx0) {
      cudaMemcpyAsync(c->d_ridx + (size_t)j0 * cap * 2, sgm ? c->sg->out().m_idx : c->lg->out().idx, (size_t)P * cap * 8, cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(c->d_rscore + (size_t)j0 * cap, sgm ? c->sg->out().m_score : c->lg->out().score, (size_t)P * cap * 4, cudaMemcpyDeviceToDevice, st);
    }
  }
  cudaMemcpyAsync(c->h_rcount, c->d_rcount, (size_t)n_jobs * 4, cudaMemcpyDeviceToHost, st);
  if (idx0) {
    cudaMemcpyAsync(c->h_ridx, c->d_ridx, (size_t)n_jobs * cap * 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(c->h_rscore, c->d_rscore, (size_t)n_jobs * cap * 4, cudaMemcpyDeviceToHost, st);
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("reloc match failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(c, AIRFE_ERR_CUDA); }
  for (int j = 0; j < n_jobs; ++j) {
    int n = c->h_rcount[j];
    if (c->h_rn[2 * j] < 1 || c->h_rn[2 * j + 1] < 1) n = 0;     // features0.cols() < 1 || features1.cols() < 1 -> 0 (point_matcher.cc:53-55)
    if (idx0 && n > match_cap) n = match_cap;
    n_match[j] = n;
    if (idx0)
      for (int k = 0; k < n; ++k) {
        idx0[(size_t)j * match_cap + k] = c->h_ridx[((size_t)j * cap + k) * 2];
        idx1[(size_t)j * match_cap + k] = c->h_ridx[((size_t)j * cap + k) * 2 + 1];
        score[(size_t)j * match_cap + k] = c->h_rscore[(size_t)j * cap + k];
      }
  }
  return AIRFE_OK;
}

void airfe_reloc_pick(int n_queries, int n_cand, const int* counts, int* best_cand, int* best_count) {
  for (int q = 0; q < n_queries; ++q) {
    int best = -1, bc = 0;                       // relocalization_matches starts empty: a candidate must have > 0 matches to win
    for (int k = 0; k < n_cand; ++k) {
      const int v = counts[(size_t)q * n_cand + k];
      if (v > bc) { bc = v; best = k; }          // strictly more than every earlier candidate (map_user.cc:370)
    }
    if (best_cand) best_cand[q] = best;
    if (best_count) best_count[q] = bc;
  }
}

long long airfe_debug_read(airfe_ctx* c, int net, const char* name, int index, void* dst, long long dst_bytes) {
  if (net == 101) {   // SuperGlue tap: "sg_scores" = final [cap+1][cap+1] score matrix of pair `index` (needs AIRFE_DEBUG_DENSE=1)
    if (!c->sg) { set_error("superglue not enabled"); return fail(c, AIRFE_ERR_INVALID); }
    const long long ld = c->sg_use->cap() + 1, nb = ld * ld * 4;
    if (nb > dst_bytes) { set_error("tap needs %lld bytes", nb); return fail(c, AIRFE_ERR_CAPACITY); }
    cudaStreamSynchronize(c->stream);
    if (cudaMemcpy(dst, c->sg_use->out().dense + (size_t)index * ld * ld, (size_t)nb, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("debug read failed"); return fail(c, AIRFE_ERR_CUDA); }
    return nb;
  }
  if (net == 100) {   // LightGlue taps: "lg_scores" = dense log-assignment [cap][cap] of pair `index` (needs AIRFE_DEBUG_DENSE=1)
    if (!c->lg) { set_error("lightglue not enabled"); return fail(c, AIRFE_ERR_INVALID); }
    const long long cap = c->lg_use->cap(), nb = cap * cap * 4;
    if (nb > dst_bytes) { set_error("tap needs %lld bytes", nb); return fail(c, AIRFE_ERR_CAPACITY); }
    cudaStreamSynchronize(c->stream);
    const float* src = !strcmp(name, "lg_scores") ? c->lg_use->out().dense + (size_t)index * cap * cap : nullptr;
    if (!src) { set_error("unknown tap %s", name); return fail(c, AIRFE_ERR_INVALID); }
    if (cudaMemcpy(dst, src, (size_t)nb, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("debug read failed"); return fail(c, AIRFE_ERR_CUDA); }
    return nb;
  }
  Detector* d = pick(c, net);
  if (!d) return fail(c, AIRFE_ERR_INVALID);
  auto it = d->taps.find(name);
  if (it == d->taps.end()) { set_error("unknown tap %s", name); return fail(c, AIRFE_ERR_INVALID); }
  const long long nb = (long long)it->second.second;
  if (nb > dst_bytes) { set_error("tap %s needs %lld bytes", name, nb); return fail(c, AIRFE_ERR_CAPACITY); }
  cudaStreamSynchronize(c->stream);
  if (cudaMemcpy(dst, (const uint8_t*)it->second.first + (size_t)index * nb, (size_t)nb, cudaMemcpyDeviceToHost) != cudaSuccess) {
    set_error("debug read failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(c, AIRFE_ERR_CUDA);
  }
  return nb;
}

}  // extern "C"
