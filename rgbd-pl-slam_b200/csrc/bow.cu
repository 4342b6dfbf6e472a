// bow.cu — ORB vocabulary on the device: per-feature tree descent of DBoW2's transform()
// (reference Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1242-1284, FORB::distance FORB.cpp:82-102), as
// Frame::ComputeBoW (include/Frame.h:80, lib/libORB_SLAM2.so@0xf84f0) needs it.  One thread per feature walks
// L levels of k children; the node table (~35 MB for ORBvoc) lives in L2.
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"

namespace plslam {

struct VocHeader {  // first bytes of the device blob
  int32_t magic, k, L, n_nodes, n_words, n_child;
  int32_t pad[2];
  uint64_t off_desc, off_child_start, off_child_idx, off_weight, off_word;
  uint64_t bytes;
};

struct VocDev {
  const uint4* desc;
  const int32_t* child_start;
  const int32_t* child_idx;
  const double* weight;
  const int32_t* word;
  int L;
};

namespace {

__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__global__ void __launch_bounds__(128) k_bow_transform(VocDev V, const uint8_t* __restrict__ desc, int n, int levelsup,
                                                       int32_t* __restrict__ word, double* __restrict__ weight,
                                                       int32_t* __restrict__ node) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* F = reinterpret_cast<const uint4*>(desc) + 2 * (size_t)i;
  const uint4 f0 = F[0], f1 = F[1];
  const int nid_level = V.L - levelsup;
  int nid = 0, final_id = 0, level = 0;
  int c0 = V.child_start[0], c1 = V.child_start[1];
  do {
    ++level;
    int best_id = __ldg(V.child_idx + c0);
    int best = hamming256(f0, f1, __ldg(V.desc + 2 * (size_t)best_id), __ldg(V.desc + 2 * (size_t)best_id + 1));
    for (int c = c0 + 1; c < c1; ++c) {
      const int id = __ldg(V.child_idx + c);
      const int d = hamming256(f0, f1, __ldg(V.desc + 2 * (size_t)id), __ldg(V.desc + 2 * (size_t)id + 1));
      if (d < best) { best = d; best_id = id; }
    }
    final_id = best_id;
    if (level == nid_level) nid = final_id;
    c0 = __ldg(V.child_start + final_id);
    c1 = __ldg(V.child_start + final_id + 1);
  } while (c1 > c0);
  word[i] = V.word[final_id];
  weight[i] = V.weight[final_id];
  node[i] = nid;
}


// Batched form on the extractor's block layout [frames][capacity][32]: one launch for every descriptor of a batch.
__global__ void __launch_bounds__(128) k_bow_transform_batch(VocDev V, const uint8_t* __restrict__ desc,
                                                             const int32_t* __restrict__ counts, int capacity, int levelsup,
                                                             int32_t* __restrict__ word, double* __restrict__ weight,
                                                             int32_t* __restrict__ node) {
  const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= min(counts[f], capacity)) return;
  const size_t o = (size_t)f * capacity + i;
  const uint4* F = reinterpret_cast<const uint4*>(desc) + 2 * o;
  const uint4 f0 = F[0], f1 = F[1];
  const int nid_level = V.L - levelsup;
  int nid = 0, final_id = 0, level = 0;
  int c0 = V.child_start[0], c1 = V.child_start[1];
  do {
    ++level;
    int best_id = __ldg(V.child_idx + c0);
    int best = hamming256(f0, f1, __ldg(V.desc + 2 * (size_t)best_id), __ldg(V.desc + 2 * (size_t)best_id + 1));
    for (int c = c0 + 1; c < c1; ++c) {
      const int id = __ldg(V.child_idx + c);
      const int d = hamming256(f0, f1, __ldg(V.desc + 2 * (size_t)id), __ldg(V.desc + 2 * (size_t)id + 1));
      if (d < best) { best = d; best_id = id; }
    }
    final_id = best_id;
    if (level == nid_level) nid = final_id;
    c0 = __ldg(V.child_start + final_id);
    c1 = __ldg(V.child_start + final_id + 1);
  } while (c1 > c0);
  word[o] = V.word[final_id];
  weight[o] = V.weight[final_id];
  node[o] = nid;
}

// FeatureVector assembly (TemplatedVocabulary.h:1196-1213: `if(w > 0) fv.addFeature(nid, i)` in feature order into a
// std::map<NodeId, vector<unsigned>>) as a CSR per frame: nodes ascending, feature indices ascending inside a node.
// One CTA per frame: bitonic sort of (node << 32 | i) keys in shared memory, heads by comparison with the left
// neighbour, ranks by a block scan.
__global__ void __launch_bounds__(256) k_featvec_csr(const int32_t* __restrict__ node, const double* __restrict__ weight,
                                                     const int32_t* __restrict__ counts, int capacity, int npow2,
                                                     int32_t* __restrict__ fv_nodes, int32_t* __restrict__ fv_start,
                                                     int32_t* __restrict__ fv_idx, int32_t* __restrict__ fv_count) {
  extern __shared__ unsigned long long fv_smem[];
  unsigned long long* key = fv_smem;
  int* head = reinterpret_cast<int*>(key + npow2);
  int* warp_tmp = head + npow2;
  const int f = blockIdx.x, t = threadIdx.x;
  const int n = min(counts[f], capacity);
  const size_t o = (size_t)f * capacity;
  for (int i = t; i < npow2; i += 256) {
    unsigned long long k = ~0ull;
    if (i < n && weight[o + i] > 0) k = ((unsigned long long)(unsigned)node[o + i] << 32) | (unsigned)i;
    key[i] = k;
  }
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = t; i < npow2; i += 256) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = key[i], b = key[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            key[i] = b;
            key[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  for (int i = t; i < npow2; i += 256) {
    const unsigned long long k = key[i];
    head[i] = (k != ~0ull && (i == 0 || (unsigned)(key[i - 1] >> 32) != (unsigned)(k >> 32))) ? 1 : 0;
  }
  __syncthreads();
  const int nuniq = block_scan_excl(head, npow2, warp_tmp);  // head[i] = rank of the node of entry i (at head entries)
  int32_t* N = fv_nodes + o;
  int32_t* S = fv_start + (size_t)f * (capacity + 1);
  int32_t* I = fv_idx + o;
  int m = 0;
  for (int i = t; i < npow2; i += 256) {
    const unsigned long long k = key[i];
    if (k == ~0ull) continue;
    I[i] = (int32_t)(unsigned)k;
    const bool isHead = i == 0 || (unsigned)(key[i - 1] >> 32) != (unsigned)(k >> 32);
    if (isHead) {
      N[head[i]] = (int32_t)(unsigned)(k >> 32);
      S[head[i]] = i;
    }
    if (i + 1 == npow2 || key[i + 1] == ~0ull) m = i + 1;  // last valid entry
  }
  if (m) S[nuniq] = m;
  if (t == 0) {
    fv_count[f] = nuniq;
    if (nuniq == 0) S[0] = 0;
  }
}

// BowVector of each frame of a batch (`if (w > 0) v.addWeight(id, w)` in feature order, then v.normalize(L1):
// TemplatedVocabulary.h:1196-1217, BowVector.cpp addWeight / normalize).  One CTA per frame: bitonic sort of
// (word << 32 | i) keys, one thread per distinct word adds its weights in ascending feature order (the order addWeight
// sees them), thread 0 adds the L1 norm in ascending word order (the map's iteration order), everybody divides.
__global__ void __launch_bounds__(256) k_bowvec(const int32_t* __restrict__ word, const double* __restrict__ weight,
                                                const int32_t* __restrict__ counts, int capacity, int npow2,
                                                int32_t* __restrict__ bow_ids, double* __restrict__ bow_vals,
                                                int32_t* __restrict__ bow_count) {
  extern __shared__ unsigned long long fv_smem[];
  unsigned long long* key = fv_smem;
  int* head = reinterpret_cast<int*>(key + npow2);
  int* warp_tmp = head + npow2;
  __shared__ double s_norm;
  const int f = blockIdx.x, t = threadIdx.x;
  const int n = min(counts[f], capacity);
  const size_t o = (size_t)f * capacity;
  for (int i = t; i < npow2; i += 256) {
    unsigned long long k = ~0ull;
    if (i < n && weight[o + i] > 0) k = ((unsigned long long)(unsigned)word[o + i] << 32) | (unsigned)i;
    key[i] = k;
  }
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = t; i < npow2; i += 256) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = key[i], b = key[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            key[i] = b;
            key[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  for (int i = t; i < npow2; i += 256) {
    const unsigned long long k = key[i];
    head[i] = (k != ~0ull && (i == 0 || (unsigned)(key[i - 1] >> 32) != (unsigned)(k >> 32))) ? 1 : 0;
  }
  __syncthreads();
  const int nuniq = block_scan_excl(head, npow2, warp_tmp);  // head[i] = rank of the word of entry i (at head entries)
  int32_t* ID = bow_ids + o;
  double* V = bow_vals + o;
  for (int i = t; i < npow2; i += 256) {
    const unsigned long long k = key[i];
    if (k == ~0ull) continue;
    const unsigned w = (unsigned)(k >> 32);
    if (i != 0 && (unsigned)(key[i - 1] >> 32) == w) continue;
    double sum = 0.0;  // map[word] = w0, then += w1, ... in feature order
    for (int j = i; j < npow2 && key[j] != ~0ull && (unsigned)(key[j] >> 32) == w; ++j) {
      const double wj = weight[o + (unsigned)key[j]];
      sum = j == i ? wj : __dadd_rn(sum, wj);
    }
    ID[head[i]] = (int32_t)w;
    V[head[i]] = sum;
  }
  __syncthreads();
  if (t == 0) {
    double norm = 0.0;
    for (int r = 0; r < nuniq; ++r) norm = __dadd_rn(norm, fabs(V[r]));
    s_norm = norm;
    bow_count[f] = nuniq;
  }
  __syncthreads();
  const double norm = s_norm;
  if (norm > 0.0)
    for (int r = t; r < nuniq; r += 256) V[r] = __ddiv_rn(V[r], norm);
}

__global__ void k_kp_angles(const plslam_keypoint_t* __restrict__ kps, size_t n, float* __restrict__ angle) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) angle[i] = kps[i].angle;
}

// SearchByBoW jobs of the frame pairs (2p = "keyframe", 2p+1 = current frame) of a batch, built on the device
__global__ void k_make_bow_jobs(int npairs, int capacity, const uint8_t* desc, const float* angle, const uint8_t* valid,
                                const int32_t* counts, const int32_t* fv_nodes, const int32_t* fv_start, const int32_t* fv_idx,
                                const int32_t* fv_count, int32_t* match, int32_t* nmatches, float nnratio, int check_ori,
                                plslam_bow_job_t* jobs) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const size_t a = (size_t)(2 * p) * capacity, b = (size_t)(2 * p + 1) * capacity;
  plslam_bow_job_t j;
  j.kf_desc = desc + a * 32;
  j.kf_angle = angle + a;
  j.kf_valid = valid + a;
  j.kf_nodes = fv_nodes + a;
  j.kf_start = fv_start + (size_t)(2 * p) * (capacity + 1);
  j.kf_idx = fv_idx + a;
  j.f_desc = desc + b * 32;
  j.f_angle = angle + b;
  j.f_nodes = fv_nodes + b;
  j.f_start = fv_start + (size_t)(2 * p + 1) * (capacity + 1);
  j.f_idx = fv_idx + b;
  j.match_f = match + (size_t)p * capacity;
  j.nmatches = nmatches + p;
  j.n1 = min(counts[2 * p], capacity);
  j.n2 = min(counts[2 * p + 1], capacity);
  j.n_kf_nodes = fv_count[2 * p];
  j.n_f_nodes = fv_count[2 * p + 1];
  j.nnratio = nnratio;
  j.check_orientation = check_ori;
  j.f_valid = nullptr;
  j.strict_low = 0;
  jobs[p] = j;
}

}  // namespace

struct Vocabulary {
  VocHeader H{};
  void* blob = nullptr;
  VocDev dev{};
  ~Vocabulary() { if (blob) cudaFree(blob); }
  void bind() {
    const char* b = static_cast<const char*>(blob);
    dev.desc = reinterpret_cast<const uint4*>(b + H.off_desc);
    dev.child_start = reinterpret_cast<const int32_t*>(b + H.off_child_start);
    dev.child_idx = reinterpret_cast<const int32_t*>(b + H.off_child_idx);
    dev.weight = reinterpret_cast<const double*>(b + H.off_weight);
    dev.word = reinterpret_cast<const int32_t*>(b + H.off_word);
    dev.L = H.L;
  }
  // host arrays in node order -> device blob
  int build(int k, int L, int n_nodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc, const double* weights) {
    PL_CHECK_ARG(k >= 1 && k <= 20 && L >= 1 && L <= 10 && n_nodes >= 2);
    std::vector<std::vector<int32_t>> children(n_nodes);
    std::vector<int32_t> word(n_nodes, -1);
    int n_words = 0;
    for (int i = 1; i < n_nodes; ++i) {
      PL_CHECK_ARG(parent[i] >= 0 && parent[i] < i);
      children[parent[i]].push_back(i);
      if (is_leaf[i]) word[i] = n_words++;
    }
    PL_CHECK_ARG(!children[0].empty());
    std::vector<int32_t> cstart(n_nodes + 1, 0), cidx;
    cidx.reserve(n_nodes);
    for (int i = 0; i < n_nodes; ++i) {
      cstart[i] = (int32_t)cidx.size();
      cidx.insert(cidx.end(), children[i].begin(), children[i].end());
    }
    cstart[n_nodes] = (int32_t)cidx.size();
    H.magic = 0x564f4332;
    H.k = k; H.L = L; H.n_nodes = n_nodes; H.n_words = n_words; H.n_child = (int32_t)cidx.size();
    size_t off = align_up(sizeof(VocHeader), 256);
    H.off_desc = off; off = align_up(off + (size_t)n_nodes * 32, 256);
    H.off_child_start = off; off = align_up(off + (size_t)(n_nodes + 1) * 4, 256);
    H.off_child_idx = off; off = align_up(off + std::max<size_t>(cidx.size(), 1) * 4, 256);
    H.off_weight = off; off = align_up(off + (size_t)n_nodes * 8, 256);
    H.off_word = off; off = align_up(off + (size_t)n_nodes * 4, 256);
    H.bytes = off;
    std::vector<char> host(off, 0);
    std::memcpy(host.data(), &H, sizeof(H));
    std::memcpy(host.data() + H.off_desc, desc, (size_t)n_nodes * 32);
    std::memset(host.data() + H.off_desc, 0, 32);  // root has no descriptor
    std::memcpy(host.data() + H.off_child_start, cstart.data(), cstart.size() * 4);
    if (!cidx.empty()) std::memcpy(host.data() + H.off_child_idx, cidx.data(), cidx.size() * 4);
    std::memcpy(host.data() + H.off_weight, weights, (size_t)n_nodes * 8);
    std::memcpy(host.data() + H.off_word, word.data(), (size_t)n_nodes * 4);
    PL_CUDA(cudaMalloc(&blob, off));
    PL_CUDA(cudaMemcpy(blob, host.data(), off, cudaMemcpyHostToDevice));
    bind();
    return PLSLAM_OK;
  }
};

}  // namespace plslam

using namespace plslam;

struct plslam_voc {
  Vocabulary v;
};

extern "C" {

int plslam_voc_create(plslam_voc_t** out, int k, int L, int n_nodes, const int32_t* parent, const uint8_t* is_leaf,
                      const uint8_t* descriptors, const double* weights) {
  PL_CHECK_ARG(out && parent && is_leaf && descriptors && weights);
  *out = nullptr;
  plslam_voc* h = new (std::nothrow) plslam_voc();
  PL_CHECK_ARG(h != nullptr);
  int rc = h->v.build(k, L, n_nodes, parent, is_leaf, descriptors, weights);
  if (rc) { delete h; return rc; }
  *out = h;
  return PLSLAM_OK;
}

int plslam_voc_load_text(plslam_voc_t** out, const char* path) {
  PL_CHECK_ARG(out && path);
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) { set_error("cannot open vocabulary %s", path); return PLSLAM_ERR_INVALID; }
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::string buf((size_t)sz, '\0');
  const size_t rd = fread(&buf[0], 1, (size_t)sz, f);
  fclose(f);
  if (rd != (size_t)sz) { set_error("short read on %s", path); return PLSLAM_ERR_INVALID; }
  const char* p = buf.c_str();
  char* e = nullptr;
  const long k = strtol(p, &e, 10); p = e;
  const long L = strtol(p, &e, 10); p = e;
  const long n1 = strtol(p, &e, 10); p = e;
  const long n2 = strtol(p, &e, 10); p = e;
  if (k < 1 || k > 20 || L < 1 || L > 10 || n1 != 0 || n2 != 0) {
    set_error("%s: not an L1_NORM/TF_IDF vocabulary text file (header %ld %ld %ld %ld)", path, k, L, n1, n2);
    return PLSLAM_ERR_INVALID;
  }
  while (*p && *p != '\n') ++p;
  if (*p == '\n') ++p;
  std::vector<int32_t> parent(1, 0);
  std::vector<uint8_t> leaf(1, 0), desc(32, 0);
  std::vector<double> weight(1, 0.0);
  // one node per remaining line; the reference also appends a node for a trailing empty line: the stream sentry
  // fails there and nothing is stored, so the node becomes a phantom extra child of the PREVIOUS line's parent with
  // that line's leaf flag (both unassigned locals), weight 0 and a descriptor that is never written
  // (TemplatedVocabulary.h:1401-1443; behaviour observed with the reference's sources compiled in oracle/_ref) — kept, with zeros
  long prevLeaf = 0, prevPid = 0;
  while (true) {
    const char* eol = p;
    while (*eol && *eol != '\n') ++eol;
    const char* q = p;
    const bool emptyLine = (q == eol);
    const long pid = emptyLine ? prevPid : strtol(q, &e, 10);
    if (!emptyLine) { q = e; prevPid = pid; }
    long lf = prevLeaf;
    if (!emptyLine) { lf = strtol(q, &e, 10); q = e; prevLeaf = lf; }
    uint8_t d[32] = {0};
    for (int i = 0; i < 32 && !emptyLine; ++i) { const long v = strtol(q, &e, 10); if (e == q || e > eol) break; d[i] = (uint8_t)v; q = e; }
    double w = emptyLine ? 0.0 : strtod(q, &e);
    if (!emptyLine && (e == q || e > eol)) w = 0.0;
    if (pid < 0 || pid >= (long)parent.size()) { set_error("%s: bad parent id %ld at node %zu", path, pid, parent.size()); return PLSLAM_ERR_INVALID; }
    parent.push_back((int32_t)pid);
    leaf.push_back(lf > 0);
    desc.insert(desc.end(), d, d + 32);
    weight.push_back(w);
    if (!*eol) break;
    p = eol + 1;
  }
  return plslam_voc_create(out, (int)k, (int)L, (int)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data());
}

void plslam_voc_destroy(plslam_voc_t* h) { delete h; }

int plslam_voc_info(const plslam_voc_t* h, int* k, int* L, int* n_nodes, int* n_words) {
  PL_CHECK_ARG(h);
  if (k) *k = h->v.H.k;
  if (L) *L = h->v.H.L;
  if (n_nodes) *n_nodes = h->v.H.n_nodes;
  if (n_words) *n_words = h->v.H.n_words;
  return PLSLAM_OK;
}

size_t plslam_voc_blob_bytes(const plslam_voc_t* h) { return h ? (size_t)h->v.H.bytes : 0; }

int plslam_voc_export_blob(const plslam_voc_t* h, void* d_out, void* stream) {
  PL_CHECK_ARG(h && d_out);
  PL_CUDA(cudaMemcpyAsync(d_out, h->v.blob, h->v.H.bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return PLSLAM_OK;
}

int plslam_voc_import_blob(plslam_voc_t** out, const void* d_blob, size_t bytes) {
  PL_CHECK_ARG(out && d_blob && bytes >= sizeof(VocHeader));
  *out = nullptr;
  VocHeader H;
  PL_CUDA(cudaMemcpy(&H, d_blob, sizeof(H), cudaMemcpyDeviceToHost));
  PL_CHECK_ARG(H.magic == 0x564f4332 && H.bytes == bytes);
  plslam_voc* h = new (std::nothrow) plslam_voc();
  PL_CHECK_ARG(h != nullptr);
  h->v.H = H;
  if (cudaMalloc(&h->v.blob, bytes) != cudaSuccess || cudaMemcpy(h->v.blob, d_blob, bytes, cudaMemcpyDeviceToDevice) != cudaSuccess) {
    set_error("vocabulary import: %s", cudaGetErrorString(cudaGetLastError()));
    delete h;
    return PLSLAM_ERR_CUDA;
  }
  h->v.bind();
  *out = h;
  return PLSLAM_OK;
}

int plslam_voc_transform_device(const plslam_voc_t* h, const uint8_t* d_descriptors, int n, int levelsup,
                                int32_t* d_word, double* d_weight, int32_t* d_node, void* stream) {
  PL_CHECK_ARG(h && n >= 0);
  if (n == 0) return PLSLAM_OK;
  PL_CHECK_ARG(d_descriptors && d_word && d_weight && d_node);
  PL_CARVEOUT(k_bow_transform);
  k_bow_transform<<<div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(h->v.dev, d_descriptors, n, levelsup, d_word, d_weight, d_node);
  PL_CUDA(cudaGetLastError());
  return PLSLAM_OK;
}

int plslam_voc_transform_host(const plslam_voc_t* h, const uint8_t* descriptors, int n, int levelsup, int32_t* word,
                              double* weight, int32_t* node) {
  PL_CHECK_ARG(h && n >= 0);
  if (n == 0) return PLSLAM_OK;
  PL_CHECK_ARG(descriptors && word && weight && node);
  DevBuf d, w, wt, nd;
  int rc;
  if ((rc = d.ensure((size_t)n * 32)) || (rc = w.ensure((size_t)n * 4)) || (rc = wt.ensure((size_t)n * 8)) || (rc = nd.ensure((size_t)n * 4))) {
    d.release(); w.release(); wt.release(); nd.release();
    return rc;
  }
  cudaError_t e = cudaMemcpy(d.p, descriptors, (size_t)n * 32, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = plslam_voc_transform_device(h, d.as<uint8_t>(), n, levelsup, w.as<int32_t>(), wt.as<double>(), nd.as<int32_t>(), nullptr);
    if (!rc) {
      e = cudaMemcpy(word, w.p, (size_t)n * 4, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemcpy(weight, wt.p, (size_t)n * 8, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemcpy(node, nd.p, (size_t)n * 4, cudaMemcpyDeviceToHost);
    }
  }
  d.release(); w.release(); wt.release(); nd.release();
  if (rc) return rc;
  if (e != cudaSuccess) { set_error("voc transform host path: %s", cudaGetErrorString(e)); return PLSLAM_ERR_CUDA; }
  return PLSLAM_OK;
}

int plslam_voc_featvec_batch_device(const plslam_voc_t* h, const uint8_t* d_descriptors, const int32_t* d_counts, int frames,
                                    int capacity, int levelsup, int32_t* d_word, double* d_weight, int32_t* d_node,
                                    int32_t* d_fv_nodes, int32_t* d_fv_start, int32_t* d_fv_idx, int32_t* d_fv_count,
                                    void* stream) {
  PL_CHECK_ARG(h && d_descriptors && d_counts && d_word && d_weight && d_node && d_fv_nodes && d_fv_start && d_fv_idx && d_fv_count);
  PL_CHECK_ARG(frames >= 1 && frames <= 65535 && capacity >= 1 && capacity <= 16384);
  cudaStream_t st = (cudaStream_t)stream;
  PL_CARVEOUT(k_bow_transform_batch);
  k_bow_transform_batch<<<dim3(div_up(capacity, 128), frames), 128, 0, st>>>(h->v.dev, d_descriptors, d_counts, capacity,
                                                                              levelsup, d_word, d_weight, d_node);
  int npow2 = 32;
  while (npow2 < capacity) npow2 <<= 1;
  const size_t smem = (size_t)npow2 * 12 + 33 * 4;
  static PerDeviceOnce attr;
  if (attr.first()) {
    PL_CUDA(cudaFuncSetAttribute(k_featvec_csr, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  PL_CARVEOUT(k_featvec_csr);
  k_featvec_csr<<<frames, 256, smem, st>>>(d_node, d_weight, d_counts, capacity, npow2, d_fv_nodes, d_fv_start, d_fv_idx,
                                           d_fv_count);
  PL_CUDA(cudaGetLastError());
  return PLSLAM_OK;
}

int plslam_voc_bowvec_batch_device(const int32_t* d_word, const double* d_weight, const int32_t* d_counts, int frames,
                                   int capacity, int32_t* d_bow_ids, double* d_bow_vals, int32_t* d_bow_count, void* stream) {
  PL_CHECK_ARG(d_word && d_weight && d_counts && d_bow_ids && d_bow_vals && d_bow_count);
  PL_CHECK_ARG(frames >= 1 && frames <= 65535 && capacity >= 1 && capacity <= 16384);
  int npow2 = 32;
  while (npow2 < capacity) npow2 <<= 1;
  const size_t smem = (size_t)npow2 * 12 + 33 * 4;
  static PerDeviceOnce attr;
  if (attr.first()) PL_CUDA(cudaFuncSetAttribute(k_bowvec, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  PL_CARVEOUT(k_bowvec);
  k_bowvec<<<frames, 256, smem, (cudaStream_t)stream>>>(d_word, d_weight, d_counts, capacity, npow2, d_bow_ids, d_bow_vals,
                                                        d_bow_count);
  PL_CUDA(cudaGetLastError());
  return PLSLAM_OK;
}

int plslam_voc_compute_bow_host(const plslam_voc_t* h, const uint8_t* descriptors, int n, int levelsup, int32_t* bow_ids,
                                double* bow_vals, int* n_bow, int32_t* fv_nodes, int32_t* fv_start, int32_t* fv_idx, int* n_fv) {
  PL_CHECK_ARG(h && n >= 0 && n <= 16384 && n_bow && n_fv);
  *n_bow = *n_fv = 0;
  if (n == 0) {
    if (fv_start) fv_start[0] = 0;
    return PLSLAM_OK;
  }
  PL_CHECK_ARG(descriptors && bow_ids && bow_vals && fv_nodes && fv_start && fv_idx);
  DevBuf all;
  // one allocation: desc | count | word | node | fv_nodes | fv_idx | bow_ids | fv_start | counts(2) | weight | bow_vals
  const size_t N = (size_t)n;
  const size_t offWord = align_up(N * 32 + 4, 16), offNode = offWord + N * 4, offFvN = offNode + N * 4, offFvI = offFvN + N * 4,
               offBowI = offFvI + N * 4, offFvS = offBowI + N * 4, offCnt = offFvS + (N + 1) * 4,
               offW = align_up(offCnt + 8, 16), offBowV = offW + N * 8, total = offBowV + N * 8;
  int rc = all.ensure(total);
  if (rc) return rc;
  uint8_t* B = all.as<uint8_t>();
  int32_t* dCount = reinterpret_cast<int32_t*>(B + N * 32);
  cudaError_t e = cudaMemcpy(B, descriptors, N * 32, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dCount, &n, 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = plslam_voc_featvec_batch_device(h, B, dCount, 1, n, levelsup, reinterpret_cast<int32_t*>(B + offWord),
                                         reinterpret_cast<double*>(B + offW), reinterpret_cast<int32_t*>(B + offNode),
                                         reinterpret_cast<int32_t*>(B + offFvN), reinterpret_cast<int32_t*>(B + offFvS),
                                         reinterpret_cast<int32_t*>(B + offFvI), reinterpret_cast<int32_t*>(B + offCnt), nullptr);
    if (!rc)
      rc = plslam_voc_bowvec_batch_device(reinterpret_cast<int32_t*>(B + offWord), reinterpret_cast<double*>(B + offW), dCount, 1, n,
                                          reinterpret_cast<int32_t*>(B + offBowI), reinterpret_cast<double*>(B + offBowV),
                                          reinterpret_cast<int32_t*>(B + offCnt) + 1, nullptr);
    int32_t cnt[2] = {0, 0};
    if (!rc) {
      e = cudaMemcpy(cnt, B + offCnt, 8, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemcpy(fv_nodes, B + offFvN, (size_t)cnt[0] * 4, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemcpy(fv_start, B + offFvS, ((size_t)cnt[0] + 1) * 4, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess && cnt[0]) {
        int32_t last = 0;
        e = cudaMemcpy(&last, B + offFvS + (size_t)cnt[0] * 4, 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(fv_idx, B + offFvI, (size_t)last * 4, cudaMemcpyDeviceToHost);
      }
      if (e == cudaSuccess) e = cudaMemcpy(bow_ids, B + offBowI, (size_t)cnt[1] * 4, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemcpy(bow_vals, B + offBowV, (size_t)cnt[1] * 8, cudaMemcpyDeviceToHost);
      *n_fv = cnt[0];
      *n_bow = cnt[1];
    }
  }
  all.release();
  if (rc) return rc;
  if (e != cudaSuccess) { set_error("ComputeBoW host path: %s", cudaGetErrorString(e)); return PLSLAM_ERR_CUDA; }
  return PLSLAM_OK;
}

int plslam_match_bow_pairs_device(const plslam_keypoint_t* d_keypoints, const uint8_t* d_descriptors, const int32_t* d_counts,
                                  int capacity, int npairs, const int32_t* d_fv_nodes, const int32_t* d_fv_start,
                                  const int32_t* d_fv_idx, const int32_t* d_fv_count, const uint8_t* d_kf_valid,
                                  float nnratio, int check_orientation, float* d_angle_scratch, plslam_bow_job_t* d_jobs_scratch,
                                  int32_t* d_match, int32_t* d_nmatches, void* stream) {
  PL_CHECK_ARG(d_keypoints && d_descriptors && d_counts && d_fv_nodes && d_fv_start && d_fv_idx && d_fv_count && d_kf_valid);
  PL_CHECK_ARG(d_angle_scratch && d_jobs_scratch && d_match && d_nmatches && capacity >= 1 && npairs >= 1);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)2 * npairs * capacity;
  PL_CARVEOUT(k_kp_angles);
  k_kp_angles<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_keypoints, n, d_angle_scratch);
  PL_CARVEOUT(k_make_bow_jobs);
  k_make_bow_jobs<<<div_up(npairs, 128), 128, 0, st>>>(npairs, capacity, d_descriptors, d_angle_scratch, d_kf_valid, d_counts,
                                                       d_fv_nodes, d_fv_start, d_fv_idx, d_fv_count, d_match, d_nmatches,
                                                       nnratio, check_orientation, d_jobs_scratch);
  PL_CUDA(cudaGetLastError());
  return plslam_match_bow_batch_device(d_jobs_scratch, npairs, capacity, stream);
}

}  // extern "C"
