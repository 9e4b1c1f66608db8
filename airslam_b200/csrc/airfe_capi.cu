// Frame-level C ABI (include/airfe_c.h): context, detect.  Host buffers in/out through pinned staging; all device work
// is asynchronous on the context's stream and synchronised once per call, right before results are handed back.
#include "../../include/airfe_c.h"
#include "detector.h"

#include <memory>

using namespace airfe;

struct airfe_ctx {
  int device = 0;
  airfe_config cfg;
  cudaStream_t stream = nullptr;
  std::unique_ptr<Detector> sp, pl;
  // pinned staging
  uint8_t* h_img = nullptr; size_t h_img_bytes = 0;
  float* h_feat = nullptr; float* h_junc = nullptr; float* h_lines = nullptr;
  int* h_counts = nullptr;   // [3][max_batch]
  uint8_t* d_img = nullptr; size_t d_img_bytes = 0;
};

extern "C" {

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
  std::unique_ptr<airfe_ctx> c(new airfe_ctx);
  c->device = device;
  c->cfg = *cfg;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream creation failed"); return AIRFE_ERR_CUDA; }
  DetectorConfig dc;
  dc.max_batch = cfg->max_batch;
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
  const int B = cfg->max_batch;
  if (cudaMallocHost(&c->h_feat, (size_t)B * kKpCap * 259 * 4) != cudaSuccess || cudaMallocHost(&c->h_junc, (size_t)B * kKpCap * 259 * 4) != cudaSuccess ||
      cudaMallocHost(&c->h_lines, (size_t)B * kLineCap * 4 * 4) != cudaSuccess || cudaMallocHost(&c->h_counts, (size_t)3 * B * 4) != cudaSuccess) {
    set_error("pinned allocation failed");
    return AIRFE_ERR_CUDA;
  }
  *out = c.release();
  return AIRFE_OK;
}

void airfe_destroy(airfe_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->sp.reset();
  c->pl.reset();
  if (c->h_img) cudaFreeHost(c->h_img);
  if (c->h_feat) cudaFreeHost(c->h_feat);
  if (c->h_junc) cudaFreeHost(c->h_junc);
  if (c->h_lines) cudaFreeHost(c->h_lines);
  if (c->h_counts) cudaFreeHost(c->h_counts);
  if (c->d_img) cudaFree(c->d_img);
  cudaStreamDestroy(c->stream);
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
  if (!c || !gray || !feat || !n_feat) { set_error("null argument"); return AIRFE_ERR_INVALID; }
  if (w <= 1 || h <= 1 || stride < w) { set_error("empty image"); return AIRFE_ERR_INVALID; }   // image.empty() -> false (plnet.cpp:247)
  Detector* d = pick(c, net);
  if (!d) return AIRFE_ERR_INVALID;
  if (batch < 1 || batch > c->cfg.max_batch) { set_error("batch %d outside [1,%d]", batch, c->cfg.max_batch); return AIRFE_ERR_INVALID; }
  if ((lines || junc) && net != AIRFE_NET_PLNET) { set_error("lines / junctions need the PLNet network"); return AIRFE_ERR_INVALID; }
  if (junc && !lines) { set_error("junction detection needs line detection"); return AIRFE_ERR_INVALID; }
  cudaSetDevice(c->device);
  const size_t one = (size_t)h * stride;
  const size_t need = one * batch;
  if (need > c->h_img_bytes) {
    if (c->h_img) cudaFreeHost(c->h_img);
    if (c->d_img) cudaFree(c->d_img);
    if (cudaMallocHost(&c->h_img, need) != cudaSuccess || cudaMalloc(&c->d_img, need) != cudaSuccess) { set_error("staging allocation failed"); return AIRFE_ERR_CUDA; }
    c->h_img_bytes = c->d_img_bytes = need;
  }
  for (int i = 0; i < batch; ++i) memcpy(c->h_img + one * i, gray + (size_t)img_stride * i, one);
  cudaStream_t st = c->stream;
  if (cudaMemcpyAsync(c->d_img, c->h_img, need, cudaMemcpyHostToDevice, st) != cudaSuccess) { set_error("H2D failed"); return AIRFE_ERR_CUDA; }
  if (!d->run(c->d_img, batch, w, h, stride, (long long)one, lines != nullptr, junc != nullptr, st)) return AIRFE_ERR_CUDA;
  const DetectOutputs& o = d->out();
  const int B = c->cfg.max_batch;
  int* hc = c->h_counts;
  cudaMemcpyAsync(hc, o.n_feat, 4 * batch, cudaMemcpyDeviceToHost, st);
  if (lines) cudaMemcpyAsync(hc + B, o.n_lines, 4 * batch, cudaMemcpyDeviceToHost, st);
  if (junc) cudaMemcpyAsync(hc + 2 * B, o.n_junc, 4 * batch, cudaMemcpyDeviceToHost, st);
  // counts are bounded by max_keypoints, so copy that many feature columns without waiting for the counts
  const int kmax = c->cfg.max_keypoints;
  for (int i = 0; i < batch; ++i)
    cudaMemcpyAsync(c->h_feat + (size_t)i * kKpCap * 259, o.feat + (size_t)i * kKpCap * 259, (size_t)kmax * 259 * 4, cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("detect failed: %s", cudaGetErrorString(cudaGetLastError())); return AIRFE_ERR_CUDA; }
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
      if (n) cudaMemcpyAsync(c->h_lines + (size_t)i * kLineCap * 4, o.lines + (size_t)i * kLineCap * 4, (size_t)n * 16, cudaMemcpyDeviceToHost, st);
    }
    if (junc)
      for (int i = 0; i < batch; ++i) {
        const int n = hc[2 * B + i] < junc_cap ? hc[2 * B + i] : junc_cap;
        n_junc[i] = n;
        if (n) cudaMemcpyAsync(c->h_junc + (size_t)i * kKpCap * 259, o.junc + (size_t)i * kKpCap * 259, (size_t)n * 259 * 4, cudaMemcpyDeviceToHost, st);
      }
    cudaStreamSynchronize(st);
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

long long airfe_debug_read(airfe_ctx* c, int net, const char* name, int index, void* dst, long long dst_bytes) {
  Detector* d = pick(c, net);
  if (!d) return AIRFE_ERR_INVALID;
  auto it = d->taps.find(name);
  if (it == d->taps.end()) { set_error("unknown tap %s", name); return AIRFE_ERR_INVALID; }
  const long long nb = (long long)it->second.second;
  if (nb > dst_bytes) { set_error("tap %s needs %lld bytes", name, nb); return AIRFE_ERR_CAPACITY; }
  cudaStreamSynchronize(c->stream);
  if (cudaMemcpy(dst, (const uint8_t*)it->second.first + (size_t)index * nb, (size_t)nb, cudaMemcpyDeviceToHost) != cudaSuccess) {
    set_error("debug read failed: %s", cudaGetErrorString(cudaGetLastError()));
    return AIRFE_ERR_CUDA;
  }
  return nb;
}

}  // extern "C"
