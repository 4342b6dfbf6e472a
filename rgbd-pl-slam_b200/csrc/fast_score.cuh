// fast_score.cuh — FAST-9/16 arc score of one pixel (cv::FAST cornerScore<16>).
#pragma once
#include <cstdint>

namespace plslam {

// s = max over the 16 arcs of 9 contiguous circle pixels of min|v - p| for the better polarity
// (0 if neither polarity has an arc of constant sign).  corner(th) <=> s > th; response = s - 1.
//
// Works on the raw pixel values: min over an arc of (v - p) = v - max(p), min of (p - v) = min(p) - v.
// Sliding 9-windows are built from 2-, 4- and 8-windows (log steps).
// NOTE: an earlier form that took min/max of the differences d = v - p and combined the polarities as
// max(mn, -mx) was miscompiled by nvcc 12.9 for sm_100a (the negation was dropped: tools/dbg/mm2.cu
// reproduces it), so keep the polarities in this explicit form; tests/test_orb_gpu.py pins the result.
__device__ __forceinline__ int fast_arc_score(const uint8_t* p, int pp) {
  const int v = p[0];
  int q[16];
  q[0] = p[3 * pp];      q[1] = p[3 * pp + 1];   q[2] = p[2 * pp + 2];   q[3] = p[pp + 3];
  q[4] = p[3];           q[5] = p[-pp + 3];      q[6] = p[-2 * pp + 2];  q[7] = p[-3 * pp + 1];
  q[8] = p[-3 * pp];     q[9] = p[-3 * pp - 1];  q[10] = p[-2 * pp - 2]; q[11] = p[-pp - 3];
  q[12] = p[-3];         q[13] = p[pp - 3];      q[14] = p[2 * pp - 2];  q[15] = p[3 * pp - 1];
  int mn2[16], mx2[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    mn2[k] = min(q[k], q[(k + 1) & 15]);
    mx2[k] = max(q[k], q[(k + 1) & 15]);
  }
  int mn4[16], mx4[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    mn4[k] = min(mn2[k], mn2[(k + 2) & 15]);
    mx4[k] = max(mx2[k], mx2[(k + 2) & 15]);
  }
  int lo = 255, hi = 0;  // smallest window maximum, largest window minimum
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int wmn = min(min(mn4[k], mn4[(k + 4) & 15]), q[(k + 8) & 15]);
    const int wmx = max(max(mx4[k], mx4[(k + 4) & 15]), q[(k + 8) & 15]);
    lo = min(lo, wmx);
    hi = max(hi, wmn);
  }
  const int darker = v - lo;    // all nine pixels of the best arc are below v by at least this much
  const int brighter = hi - v;  // ... above v by at least this much
  int best = 0;
  if (darker > best) best = darker;
  if (brighter > best) best = brighter;
  return best;
}

}  // namespace plslam
