// BoW vocabulary (DBoW2 k-ary tree) resident on the device + the quantisation kernel (bow.cu).
#pragma once
#include "engine.h"

#include <string>
#include <vector>

namespace airfe {

struct BowVocabulary {
  int k = 0, L = 0, n_nodes = 0;
  int* d_children = nullptr;     // [n_nodes][k], -1 padded; children[node][0] < 0 <=> leaf
  float* d_desc = nullptr;       // [n_nodes][256]
  int* d_leaf = nullptr;         // scratch: leaf node per keypoint
  int leaf_cap = 0;
  std::vector<int> word_id;      // host: per node (valid at leaves)
  std::vector<double> weight;    // host: idf weight per node
  ~BowVocabulary();
  bool upload(int k, int L, int n_nodes, const int* children, const float* desc, const int* word_id, const double* weight);
  bool load_afw(const std::string& path);             // weights/point_voc_L4.afw (tools/make_voc.py)
  bool load_boost_archive(const std::string& path);   // the reference's voc/point_voc_L4.bin itself
};

// d_feat: device rows of `feat_stride` floats, descriptor at columns 3..258 (the [n][259] feature layout); d_leaf [n] receives the leaf NODE id
bool bow_transform_device(const BowVocabulary& voc, const float* d_feat, long long feat_stride, int n, int* d_leaf, cudaStream_t st);

}  // namespace airfe
