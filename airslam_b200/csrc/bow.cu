// BoW quantisation of 256-d descriptors (Database::FrameToBow, src/bow/database.cc:57-89) on the device.
// TemplatedVocabulary::transform (3rdparty/DBoW2/include/DBoW2/TemplatedVocabulary.h:1313-1351) walks a k-ary tree (voc/point_voc_L4.bin:
// k = 10, L = 4, 11 111 nodes, 10 000 words) choosing at every level the child with the smallest squared L2 distance
// (FSuperpoint::distance, src/bow/FSuperpoint.cc:46-50; strict '<': the first minimum wins).  One warp per keypoint: the 256-d feature sits
// in registers (8 values per lane), each child descriptor is one coalesced 1 KiB read, the distance is a warp-shuffle reduction.  The tree
// (11.4 MB) stays resident on the device.  The idf accumulation / L1 normalisation of the (<= N entries) BowVector is the reference's own
// std::map arithmetic in double, done on the host in the C wrapper.
#include "bow.h"

#include <stdio.h>
#include <string.h>

#include <fstream>
#include <map>

namespace airfe {

BowVocabulary::~BowVocabulary() {
  if (d_children) cudaFree(d_children);
  if (d_desc) cudaFree(d_desc);
  if (d_leaf) cudaFree(d_leaf);
}

bool BowVocabulary::upload(int k_, int L_, int n_nodes_, const int* children, const float* desc, const int* word_id_, const double* weight_) {
  k = k_; L = L_; n_nodes = n_nodes_;
  if (k < 1 || k > 32 || n_nodes < 1) { set_error("bow: bad vocabulary shape (k = %d, %d nodes)", k, n_nodes); return false; }
  word_id.assign(word_id_, word_id_ + n_nodes);
  weight.assign(weight_, weight_ + n_nodes);
  for (long long i = 0; i < (long long)n_nodes * k; ++i)
    if (children[i] >= n_nodes) { set_error("bow: child id out of range"); return false; }
  AIRFE_CUDA_OK(cudaMalloc(&d_children, (size_t)n_nodes * k * 4));
  AIRFE_CUDA_OK(cudaMalloc(&d_desc, (size_t)n_nodes * 256 * 4));
  AIRFE_CUDA_OK(cudaMemcpy(d_children, children, (size_t)n_nodes * k * 4, cudaMemcpyHostToDevice));
  AIRFE_CUDA_OK(cudaMemcpy(d_desc, desc, (size_t)n_nodes * 256 * 4, cudaMemcpyHostToDevice));
  return true;
}

bool BowVocabulary::load_afw(const std::string& path) {
  WeightFile wf;
  if (!wf.load(path)) return false;
  const WTensor *m = wf.find("voc.meta"), *c = wf.find("voc.children"), *d = wf.find("voc.desc"), *w = wf.find("voc.word_id"), *q = wf.find("voc.weight_f64_bits");
  if (!m || !c || !d || !w || !q || m->dtype != 2 || c->dtype != 2 || d->dtype != 0 || w->dtype != 2 || q->dtype != 2 || m->numel() < 2) { set_error("bow: %s is not a vocabulary container", path.c_str()); return false; }
  const int* meta = (const int*)m->data;
  const int n = c->dims[0];
  if (c->dims[1] != meta[0] || d->numel() != (size_t)n * 256 || w->numel() != (size_t)n || q->numel() != (size_t)n * 2) { set_error("bow: inconsistent vocabulary tensors"); return false; }
  std::vector<double> wt(n);
  memcpy(wt.data(), q->data, (size_t)n * 8);
  return upload(meta[0], meta[1], n, (const int*)c->data, (const float*)d->data, (const int*)w->data, wt.data());
}

// The reference's own file: a Boost binary archive written by the serialize() overloads of include/bow/database.h:35-55 (layout in tools/make_voc.py)
bool BowVocabulary::load_boost_archive(const std::string& path) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) { set_error("bow: cannot open %s", path.c_str()); return false; }
  const size_t size = (size_t)f.tellg();
  std::vector<uint8_t> b(size);
  f.seekg(0);
  f.read((char*)b.data(), size);
  size_t off = 0;
  auto need = [&](size_t n) { return off + n <= size; };
  auto rd = [&](void* dst, size_t n) { memcpy(dst, b.data() + off, n); off += n; };
  uint64_t sl = 0;
  if (!need(8)) { set_error("bow: %s truncated", path.c_str()); return false; }
  rd(&sl, 8);
  if (sl != 22 || !need(22 + 15) || memcmp(b.data() + off, "serialization::archive", 22) != 0) { set_error("bow: %s is not a boost binary archive", path.c_str()); return false; }
  off += 22 + 2 + 4 + 4 + 5;          // signature, library version, sizeof table, endian marker, class header
  int32_t hdr[4];
  if (!need(16 + 5 + 12)) { set_error("bow: %s truncated", path.c_str()); return false; }
  rd(hdr, 16);
  off += 5;
  uint64_t cnt = 0;
  rd(&cnt, 8);
  off += 4;
  const int kk = hdr[0];
  if (kk < 1 || kk > 32 || cnt < 1 || cnt > (1u << 26)) { set_error("bow: implausible vocabulary header (k = %d, %llu nodes)", kk, (unsigned long long)cnt); return false; }
  std::vector<int> children((size_t)cnt * kk, -1), wid(cnt);
  std::vector<float> desc((size_t)cnt * 256);
  std::vector<double> wt(cnt);
  for (uint64_t i = 0; i < cnt; ++i) {
    if (i == 0) off += 5;
    uint32_t ids[2];
    uint64_t nc;
    if (!need(8 + 8 + 8)) { set_error("bow: %s truncated at node %llu", path.c_str(), (unsigned long long)i); return false; }
    rd(ids, 8);
    rd(&wt[i], 8);
    rd(&nc, 8);
    if (ids[1] != i || nc > (uint64_t)kk || !need(4 * nc + 4 + 1024 + 4)) { set_error("bow: %s: node %llu is malformed", path.c_str(), (unsigned long long)i); return false; }
    rd(&children[(size_t)i * kk], 4 * nc);
    off += 4;                           // parent
    rd(&desc[(size_t)i * 256], 1024);
    rd(&wid[i], 4);
  }
  return upload(kk, hdr[1], (int)cnt, children.data(), desc.data(), wid.data(), wt.data());
}

// one warp per keypoint
__global__ void __launch_bounds__(256) bow_transform_kernel(const float* __restrict__ feat, long long feat_stride, int n, const int* __restrict__ children,
                                                            const float* __restrict__ desc, int k, int* __restrict__ leaf) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  const float* f = feat + (long long)i * feat_stride + 3 + lane * 8;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = f[e];
  int node = 0;
  while (true) {
    const int* ch = children + (long long)node * k;
    if (ch[0] < 0) break;
    int best = -1;
    float best_d = 0.f;
    for (int c = 0; c < k; ++c) {
      const int id = ch[c];
      if (id < 0) break;
      const float4 a = *reinterpret_cast<const float4*>(desc + (long long)id * 256 + lane * 8);
      const float4 b = *reinterpret_cast<const float4*>(desc + (long long)id * 256 + lane * 8 + 4);
      const float d0 = v[0] - a.x, d1 = v[1] - a.y, d2 = v[2] - a.z, d3 = v[3] - a.w, d4 = v[4] - b.x, d5 = v[5] - b.y, d6 = v[6] - b.z, d7 = v[7] - b.w;
      float s = d0 * d0;
      s = fmaf(d1, d1, s); s = fmaf(d2, d2, s); s = fmaf(d3, d3, s); s = fmaf(d4, d4, s); s = fmaf(d5, d5, s); s = fmaf(d6, d6, s); s = fmaf(d7, d7, s);
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (best < 0 || s < best_d) { best = id; best_d = s; }       // strict '<': the first minimum wins
    }
    node = best;
  }
  if (lane == 0) leaf[i] = node;
}

bool bow_transform_device(const BowVocabulary& voc, const float* d_feat, long long feat_stride, int n, int* d_leaf, cudaStream_t st) {
  if (n <= 0) return true;
  bow_transform_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(d_feat, feat_stride, n, voc.d_children, voc.d_desc, voc.k, d_leaf);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("bow kernel launch failed: %s", cudaGetErrorString(e)); return false; }
  return true;
}

}  // namespace airfe
