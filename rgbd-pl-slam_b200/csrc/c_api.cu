#include <cstdlib>
// c_api.cu — extern "C" entry points declared in include/plslam_b200.h (the drop-in boundary).
#include <new>

#include "common.cuh"
#include "lines.cuh"
#include "orb.cuh"

namespace plslam {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace plslam

using namespace plslam;

struct plslam_orb {
  OrbExtractor impl;
  plslam_orb(int nf, float sf, int nl, int ini, int mn) : impl(nf, sf, nl, ini, mn) {}
};

struct plslam_lines {
  LineExtractor impl;
};

// The pipelined front-end keeps two streams per batch in flight.  CUDA maps streams onto CUDA_DEVICE_MAX_CONNECTIONS
// hardware queues (default 8); streams that share a queue serialise behind each other's 50 ms region-growing kernels.
// Ask for the maximum (32) when the library is loaded, unless the host application already chose a value.  This only
// takes effect if the CUDA context is created after the library is loaded.
__attribute__((constructor)) static void plslam_default_connections() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

namespace plslam {
int carveout_pct() {
  static const int v = [] {
    const char* e = std::getenv("PLSLAM_CARVEOUT");
    return e ? std::atoi(e) : 65;
  }();
  return v;
}
}  // namespace plslam

extern "C" {

const char* plslam_last_error(void) { return g_err; }
const char* plslam_version(void) { return "plslam_b200 0.1 (sm_100a)"; }

int plslam_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int plslam_orb_create(plslam_orb_t** out, int nfeatures, float scaleFactor, int nlevels, int iniThFAST,
                      int minThFAST) {
  PL_CHECK_ARG(out != nullptr);
  *out = nullptr;
  PL_CHECK_ARG(nfeatures > 0 && nlevels >= 1 && nlevels <= ORB_MAXL && scaleFactor > 1.0f);
  PL_CHECK_ARG(iniThFAST >= minThFAST && minThFAST >= 1 && iniThFAST < 255);
  plslam_orb* h = new (std::nothrow) plslam_orb(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
  if (!h) {
    set_error("out of host memory");
    return PLSLAM_ERR_INVALID;
  }
  *out = h;
  return PLSLAM_OK;
}

void plslam_orb_destroy(plslam_orb_t* h) { delete h; }

int plslam_orb_set_blur_kernel(plslam_orb_t* h, const int32_t k[7]) {
  PL_CHECK_ARG(h && k);
  return h->impl.set_blur_kernel(k);
}

int plslam_orb_levels(const plslam_orb_t* h) { return h ? h->impl.nlevels : 0; }

int plslam_orb_tables(const plslam_orb_t* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                      int32_t* features_per_level, int32_t* umax16) {
  PL_CHECK_ARG(h);
  const OrbExtractor& o = h->impl;
  for (int i = 0; i < o.nlevels; ++i) {
    if (scale) scale[i] = o.mvScaleFactor[i];
    if (inv_scale) inv_scale[i] = o.mvInvScaleFactor[i];
    if (sigma2) sigma2[i] = o.mvLevelSigma2[i];
    if (inv_sigma2) inv_sigma2[i] = o.mvInvLevelSigma2[i];
    if (features_per_level) features_per_level[i] = o.mnFeaturesPerLevel[i];
  }
  if (umax16)
    for (int i = 0; i < 16; ++i) umax16[i] = o.umax[i];
  return PLSLAM_OK;
}

int plslam_orb_max_keypoints(const plslam_orb_t* h) { return h ? h->impl.max_keypoints() : 0; }

int plslam_orb_extract(plslam_orb_t* h, const uint8_t* image, int width, int height, int pitch,
                       plslam_keypoint_t* keypoints, uint8_t* descriptors, int capacity, int* n_out) {
  PL_CHECK_ARG(h && n_out);
  *n_out = 0;
  if (!image || width <= 0 || height <= 0) return PLSLAM_OK;  // empty image: silent return (@0x76dda)
  int32_t cnt = 0;
  int rc = h->impl.extract_host(image, 1, width, height, pitch, (size_t)pitch * height, keypoints, descriptors,
                                capacity, &cnt);
  *n_out = cnt;
  return rc;
}

int plslam_orb_extract_batch_host(plslam_orb_t* h, const uint8_t* images, int batch, int width, int height,
                                  int pitch, size_t frame_stride, plslam_keypoint_t* keypoints,
                                  uint8_t* descriptors, int capacity, int32_t* counts) {
  PL_CHECK_ARG(h);
  return h->impl.extract_host(images, batch, width, height, pitch, frame_stride, keypoints, descriptors, capacity,
                              counts);
}

int plslam_orb_extract_batch_device(plslam_orb_t* h, const uint8_t* d_images, int batch, int width, int height,
                                    int pitch, size_t frame_stride, plslam_keypoint_t* d_keypoints,
                                    uint8_t* d_descriptors, int capacity, int32_t* d_counts, void* stream) {
  PL_CHECK_ARG(h);
  return h->impl.extract_device(d_images, batch, width, height, pitch, frame_stride, d_keypoints, d_descriptors,
                                capacity, d_counts, (cudaStream_t)stream);
}

int plslam_orb_check_status(plslam_orb_t* h, void* stream) {
  PL_CHECK_ARG(h);
  return h->impl.check_status((cudaStream_t)stream);
}

int plslam_orb_level_size(const plslam_orb_t* h, int level, int* width, int* height) {
  PL_CHECK_ARG(h && width && height);
  return h->impl.level_size(level, width, height);
}

int plslam_orb_copy_level(plslam_orb_t* h, int frame, int level, int which, uint8_t* out, size_t out_bytes) {
  PL_CHECK_ARG(h);
  return h->impl.copy_level(frame, level, which, out, out_bytes);
}

int plslam_orb_copy_candidates(plslam_orb_t* h, int frame, int level, int32_t* xyr, int capacity, int* n_out) {
  PL_CHECK_ARG(h);
  return h->impl.copy_candidates(frame, level, xyr, capacity, n_out);
}

// ---- lines ----
int plslam_lines_create(plslam_lines_t** out) {
  PL_CHECK_ARG(out != nullptr);
  *out = new (std::nothrow) plslam_lines();
  if (!*out) {
    set_error("out of host memory");
    return PLSLAM_ERR_INVALID;
  }
  return PLSLAM_OK;
}
void plslam_lines_destroy(plslam_lines_t* h) { delete h; }
int plslam_lines_set_max_lines(plslam_lines_t* h, int max_lines) {
  PL_CHECK_ARG(h && max_lines >= 0 && max_lines <= 256);
  h->impl.set_max_lines(max_lines);
  return PLSLAM_OK;
}
int plslam_lines_capacity(const plslam_lines_t* h) { return h ? h->impl.out_capacity() : 0; }
int plslam_lines_extract(plslam_lines_t* h, const uint8_t* image, int width, int height, int pitch,
                         plslam_keyline_t* keylines, uint8_t* descriptors, double* line_functions, int capacity,
                         int* n_out) {
  PL_CHECK_ARG(h && n_out);
  *n_out = 0;
  if (!image || width <= 0 || height <= 0) return PLSLAM_OK;
  int32_t cnt = 0;
  int rc = h->impl.extract_host(image, 1, width, height, pitch, (size_t)pitch * height, keylines, descriptors,
                                line_functions, capacity, &cnt);
  *n_out = cnt;
  return rc;
}
int plslam_lines_extract_batch_host(plslam_lines_t* h, const uint8_t* images, int batch, int width, int height,
                                    int pitch, size_t frame_stride, plslam_keyline_t* keylines, uint8_t* descriptors,
                                    double* line_functions, int capacity, int32_t* counts) {
  PL_CHECK_ARG(h);
  return h->impl.extract_host(images, batch, width, height, pitch, frame_stride, keylines, descriptors, line_functions,
                              capacity, counts);
}
int plslam_lines_extract_batch_device(plslam_lines_t* h, const uint8_t* d_images, int batch, int width, int height,
                                      int pitch, size_t frame_stride, plslam_keyline_t* d_keylines,
                                      uint8_t* d_descriptors, double* d_line_functions, int capacity,
                                      int32_t* d_counts, void* stream) {
  PL_CHECK_ARG(h);
  return h->impl.extract_device(d_images, batch, width, height, pitch, frame_stride, d_keylines, d_descriptors,
                                d_line_functions, capacity, d_counts, (cudaStream_t)stream);
}
int plslam_lines_check_status(plslam_lines_t* h, void* stream) {
  PL_CHECK_ARG(h);
  return h->impl.check_status((cudaStream_t)stream);
}
int plslam_lines_scaled_size(const plslam_lines_t* h, int* width, int* height) {
  PL_CHECK_ARG(h && width && height);
  return h->impl.scaled_size(width, height);
}
int plslam_lines_copy_scaled(plslam_lines_t* h, int frame, uint8_t* out, size_t out_bytes) {
  PL_CHECK_ARG(h);
  return h->impl.copy_scaled(frame, out, out_bytes);
}
int plslam_lines_copy_level_lines(plslam_lines_t* h, int frame, float* degrees, int32_t* grad2, size_t count) {
  PL_CHECK_ARG(h);
  return h->impl.copy_angles(frame, degrees, grad2, count);
}
int plslam_lines_compute_lbd(plslam_lines_t* h, const uint8_t* image, int width, int height, int pitch,
                             const plslam_keyline_t* keylines, int n, uint8_t* descriptors) {
  PL_CHECK_ARG(h);
  return h->impl.compute_lbd_host(image, width, height, pitch, keylines, n, descriptors);
}
int plslam_lines_copy_segments(plslam_lines_t* h, int frame, double* seg7, int capacity, int* n_out) {
  PL_CHECK_ARG(h && n_out);
  std::vector<LsdSegment> v(capacity > 0 ? capacity : 1);
  int rc = h->impl.copy_segments(frame, v.data(), capacity, n_out);
  if (rc) return rc;
  for (int i = 0; i < *n_out; ++i) {
    double* o = seg7 + 7 * i;
    o[0] = v[i].x1; o[1] = v[i].y1; o[2] = v[i].x2; o[3] = v[i].y2;
    o[4] = v[i].width; o[5] = v[i].prec; o[6] = v[i].nfa;
  }
  return PLSLAM_OK;
}

}  // extern "C"
