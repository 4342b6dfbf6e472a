// lines.cuh — parameter block and host context of the LSD + LBD line extractor.
#pragma once
#include <vector>

#include "common.cuh"

namespace plslam {

constexpr int LSD_BINS = 1024;
constexpr int LBD_W = 7, LBD_BANDS = 9, LBD_ROWS = LBD_W * LBD_BANDS;

struct LsdRect {  // lsd.cpp `rect`
  double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p;
};

struct LsdSegment {  // one accepted line segment (full-resolution coordinates)
  float x1, y1, x2, y2;
  double width, prec, nfa;
};

struct LineParams {
  int W, H;            // input frame
  int sw, sh, spitch;  // 0.8-scaled image
  int P;               // sw * sh
  int ksize;           // Gaussian pre-blur taps (7 for sigma 0.75)
  int blurk[9];
  double scale;        // 0.8
  double rho, prec, p, log_nt, density_th, log_eps;
  int min_reg_size;
  int g2_min;          // smallest gx^2 + gy^2 whose gradient norm exceeds rho
  int max_lines;       // keep the strongest max_lines by response (0 = keep all)
  int rect_cap;        // capacity of the per-frame rectangle / segment lists
  int out_cap;         // capacity of the per-frame output (keylines kept)
  int batch;           // frames of the current launch
  int grow_variant;    // k_lsd_grow: bit 0 in-batch speculation, bit 1 L1 prefetch of a new point's rows, bit 2 bitmap-guided prefetch
  float gaussL[LBD_W * 3], gaussG[LBD_ROWS];
};

class LineExtractor {
 public:
  LineExtractor();
  ~LineExtractor();
  void set_max_lines(int n) { max_lines = n; cfgW = 0; }
  int out_capacity() const { return max_lines > 0 ? max_lines : 8192; }  // keep-all mode: fixed bound, overflow reported

  int extract_device(const uint8_t* d_images, int batch, int W, int H, int pitch, size_t frame_stride,
                     plslam_keyline_t* d_keylines, uint8_t* d_desc, double* d_funcs, int capacity, int32_t* d_counts,
                     cudaStream_t st);
  int extract_host(const uint8_t* images, int batch, int W, int H, int pitch, size_t frame_stride,
                   plslam_keyline_t* keylines, uint8_t* desc, double* funcs, int capacity, int32_t* counts);
  int check_status(cudaStream_t st);
  // parity accessors (last batch)
  int scaled_size(int* w, int* h) const;
  int copy_scaled(int frame, uint8_t* out, size_t bytes);
  int copy_angles(int frame, float* deg_out, int32_t* g2_out, size_t n);
  int copy_segments(int frame, LsdSegment* out, int capacity, int* n_out);
  int compute_lbd_host(const uint8_t* image, int W, int H, int pitch, const plslam_keyline_t* keylines, int n, uint8_t* desc);

  StageTimer* timer = nullptr;
  const int* device_status() const { return status.as<int>(); }
  // optional marker for the caller's scheduling: recorded on the extraction stream just before (1) or just after (2) the
  // region-growing kernel is enqueued
  cudaEvent_t mark_event = nullptr;
  int mark_where = 0;
  int batches_in_flight = 1;  // how many batches like this one the caller keeps on the GPU at once (pipeline depth)
  int max_lines = 40;  // lsdNFeatures of the PL-SLAM fork family
  int rect_cap = 4096;

 private:
  int configure(int W, int H, int batch);
  int device = -1, numSMs = 148, cfgW = 0, cfgH = 0, cfgB = 0, last_batch = 0;
  LineParams P{};
  DevBuf scaled, pix, degp, g2p, ubm, owner, recttmp, listpool, rectstage, coef, rowhist, binstart, maxg2, seeds, nseeds, regbuf, rects, nrects, nfaq, rectout, segs, nsegs, resp,
      rowsum, status;
  DevBuf stageIn, stageKl, stageDesc, stageFuncs, stageCnt;
  cudaStream_t ownStream = nullptr;
  void* pinnedStatus = nullptr;
  bool statusArmed = false;
};

}  // namespace plslam
