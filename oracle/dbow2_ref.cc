// oracle/dbow2_ref.cc — thin C wrapper around the reference's OWN DBoW2 sources, compiled from where they lie
// under /root/reference/Thirdparty/DBoW2 (oracle/Makefile target `ref`, output oracle/_ref/libdbow2_ref.so).
// TEST INFRASTRUCTURE: a true reference build of FORB::distance (== ORBmatcher::DescriptorDistance,
// Dependencies.md:16-18) and of ORBVocabulary::transform as called by Frame::ComputeBoW
// (lib/libORB_SLAM2.so@0xf8551: transform(vDesc, mBowVec, mFeatVec, levelsup = 4)).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;  // include/ORBVocabulary.h:31-32

extern "C" {

int dbow2_ref_forb_distance(const uint8_t* a, const uint8_t* b) {
  cv::Mat ma(1, 32, CV_8U), mb(1, 32, CV_8U);
  std::memcpy(ma.data, a, 32);
  std::memcpy(mb.data, b, 32);
  return DBoW2::FORB::distance(ma, mb);
}

void* dbow2_ref_load_text(const char* path) {
  ORBVocabulary* v = new ORBVocabulary();
  if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
  return v;
}
void dbow2_ref_free(void* h) { delete (ORBVocabulary*)h; }
int dbow2_ref_size(void* h) { return (int)((ORBVocabulary*)h)->size(); }

// transform(features, BowVector, FeatureVector, levelsup).  Outputs flattened in map order:
// bow_ids/bow_vals (n_bow entries), fv_nodes + CSR fv_start/fv_idx.  Returns 0 on success.
int dbow2_ref_transform(void* h, const uint8_t* desc, int n, int levelsup, uint32_t* bow_ids, double* bow_vals, int* n_bow,
                        uint32_t* fv_nodes, int* fv_start, uint32_t* fv_idx, int* n_fv) {
  ORBVocabulary* v = (ORBVocabulary*)h;
  std::vector<cv::Mat> feats(n);
  for (int i = 0; i < n; ++i) {
    feats[i].create(1, 32, CV_8U);
    std::memcpy(feats[i].data, desc + 32 * (size_t)i, 32);
  }
  DBoW2::BowVector bv;
  DBoW2::FeatureVector fv;
  v->transform(feats, bv, fv, levelsup);
  int k = 0;
  for (auto& e : bv) { bow_ids[k] = e.first; bow_vals[k] = e.second; ++k; }
  *n_bow = k;
  int m = 0, pos = 0;
  for (auto& e : fv) {
    fv_nodes[m] = e.first;
    fv_start[m] = pos;
    for (unsigned id : e.second) fv_idx[pos++] = id;
    ++m;
  }
  fv_start[m] = pos;
  *n_fv = m;
  return 0;
}

}  // extern "C"
