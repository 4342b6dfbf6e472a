// oracle/bow_oracle.cc — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// Restatement of what Frame::ComputeBoW (reference include/Frame.h:80, lib/libORB_SLAM2.so@0xf84f0) asks of DBoW2:
//   ORBVocabulary::loadFromTextFile   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1362-1448
//   transform(features, BowVector&, FeatureVector&, levelsup)   TemplatedVocabulary.h:1151-1217
//   transform(feature, word_id, weight, nid, levelsup)          TemplatedVocabulary.h:1242-1284
//   BowVector::addWeight / normalize(L1), FeatureVector::addFeature   BowVector.cpp, FeatureVector.cpp
// for the configuration ORBvoc.txt uses (first line "10 6 0 0": L1_NORM scoring, TF_IDF weighting).
// PINNED: tests/test_bow_cpu.py compares it with oracle/_ref/libdbow2_ref.so, which is the reference's own
// DBoW2 source compiled unmodified (oracle/Makefile target `ref`).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace {

struct Node {
  int parent = 0;
  std::vector<int> children;
  uint8_t desc[32];
  double weight = 0;
  int word_id = -1;
};

struct Voc {
  int k = 0, L = 0, scoring = 0, weighting = 0;
  std::vector<Node> nodes;
  int nwords = 0;
};

int popcount_dist(const uint8_t* a, const uint8_t* b) {
  int d = 0;
  for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
  return d;
}

}  // namespace

extern "C" {

void* oracle_voc_load_text(const char* path) {
  std::ifstream f(path);
  if (!f.good()) return nullptr;
  Voc* v = new Voc();
  std::string s;
  std::getline(f, s);
  {
    std::stringstream ss(s);
    ss >> v->k >> v->L >> v->scoring >> v->weighting;
  }
  if (v->k < 0 || v->k > 20 || v->L < 1 || v->L > 10 || v->scoring != 0 || v->weighting != 0) { delete v; return nullptr; }
  v->nodes.resize(1);
  while (!f.eof()) {
    std::string snode;
    std::getline(f, snode);
    std::stringstream ss(snode);
    // the reference appends a node for every line it reads, including a trailing empty one (TemplatedVocabulary.h:1401-1443)
    int nid = (int)v->nodes.size();
    v->nodes.resize(nid + 1);
    // on a trailing empty line the stream sentry fails, so no extraction stores anything: `pid` and the leaf flag
    // (unassigned locals in the reference) keep the previous line's values, the descriptor and weight are never
    // written => a phantom, weightless extra child of the LAST parent (observed with the reference's sources
    // compiled here; its descriptor is uninitialised memory there, zeros here)
    static thread_local int leaf = 0, pid = 0;
    { int v2; if (ss >> v2) pid = v2; }
    v->nodes[nid].parent = pid;
    v->nodes[pid].children.push_back(nid);
    { int lf; if (ss >> lf) leaf = lf; }
    for (int i = 0; i < 32; ++i) {
      int e = 0;
      if (!(ss >> e)) e = 0;
      v->nodes[nid].desc[i] = (uint8_t)e;
    }
    ss >> v->nodes[nid].weight;
    if (leaf > 0) v->nodes[nid].word_id = v->nwords++;
  }
  return v;
}
void oracle_voc_free(void* h) { delete (Voc*)h; }
int oracle_voc_nodes(void* h) { return (int)((Voc*)h)->nodes.size(); }
int oracle_voc_words(void* h) { return ((Voc*)h)->nwords; }

// flat export for building the device vocabulary in tests: parent[n], first_child[n], n_children[n], desc[n*32], weight[n], word_id[n]
void oracle_voc_export(void* h, int* parent, int* first_child, int* n_children, uint8_t* desc, double* weight, int* word_id, int* kL) {
  Voc* v = (Voc*)h;
  kL[0] = v->k; kL[1] = v->L;
  for (size_t i = 0; i < v->nodes.size(); ++i) {
    const Node& n = v->nodes[i];
    parent[i] = n.parent;
    first_child[i] = n.children.empty() ? -1 : n.children[0];
    n_children[i] = (int)n.children.size();
    std::memcpy(desc + 32 * i, n.desc, 32);
    weight[i] = n.weight;
    word_id[i] = n.word_id;
  }
}

// per-feature descent: word id, weight, node id at level L - levelsup
void oracle_voc_transform_features(void* h, const uint8_t* desc, int n, int levelsup, int* word, double* weight, int* node) {
  Voc* v = (Voc*)h;
  const int nid_level = v->L - levelsup;
  for (int i = 0; i < n; ++i) {
    const uint8_t* f = desc + 32 * (size_t)i;
    int nid = 0;
    if (nid_level <= 0) nid = 0;
    int final_id = 0, level = 0;
    do {
      ++level;
      const std::vector<int>& ch = v->nodes[final_id].children;
      final_id = ch[0];
      double best = popcount_dist(f, v->nodes[final_id].desc);
      for (size_t c = 1; c < ch.size(); ++c) {
        const double d = popcount_dist(f, v->nodes[ch[c]].desc);
        if (d < best) { best = d; final_id = ch[c]; }
      }
      if (level == nid_level) nid = final_id;
    } while (!v->nodes[final_id].children.empty());
    word[i] = v->nodes[final_id].word_id;
    weight[i] = v->nodes[final_id].weight;
    node[i] = nid;
  }
}

// full transform -> BowVector (ascending word id, L1-normalised) and FeatureVector (ascending node id, CSR)
int oracle_voc_transform(void* h, const uint8_t* desc, int n, int levelsup, uint32_t* bow_ids, double* bow_vals, int* n_bow,
                         uint32_t* fv_nodes, int* fv_start, uint32_t* fv_idx, int* n_fv) {
  std::vector<int> word(n), node(n);
  std::vector<double> w(n);
  oracle_voc_transform_features(h, desc, n, levelsup, word.data(), w.data(), node.data());
  std::map<uint32_t, double> bv;
  std::map<uint32_t, std::vector<uint32_t>> fv;
  for (int i = 0; i < n; ++i) {
    if (w[i] > 0) {
      auto it = bv.lower_bound((uint32_t)word[i]);
      if (it != bv.end() && !(bv.key_comp()((uint32_t)word[i], it->first))) it->second += w[i];
      else bv.insert(it, std::make_pair((uint32_t)word[i], w[i]));
      fv[(uint32_t)node[i]].push_back((uint32_t)i);
    }
  }
  double norm = 0.0;
  for (auto& e : bv) norm += std::fabs(e.second);
  if (norm > 0.0) for (auto& e : bv) e.second /= norm;
  int k = 0;
  for (auto& e : bv) { bow_ids[k] = e.first; bow_vals[k] = e.second; ++k; }
  *n_bow = k;
  int m = 0, pos = 0;
  for (auto& e : fv) {
    fv_nodes[m] = e.first;
    fv_start[m] = pos;
    for (uint32_t id : e.second) fv_idx[pos++] = id;
    ++m;
  }
  fv_start[m] = pos;
  *n_fv = m;
  return 0;
}

}  // extern "C"
