// fast_score.cuh — FAST-9/16 arc score of one pixel (cv::FAST cornerScore<16>).
#pragma once
#include <cstdint>

namespace plslam {

// s = max over the 16 arcs of 9 contiguous circle pixels of min|v - p| for the better polarity
// (0 if neither polarity has an arc of constant sign).  corner(th) <=> s > th; response = s - 1.
//
// Works on the raw pixel values: min over an arc of (v - p) = v - max(p), min of (p - v) = min(p) - v.
// Sliding 9-windows are built from 2- and 4-windows (log steps) and one three-input step.
// NOTE: an earlier form that took min/max of the differences d = v - p and combined the polarities as
// max(mn, -mx) was miscompiled by nvcc 12.9 for sm_100a (the negation was dropped: tools/dbg/mm2.cu
// reproduces it), so keep the polarities in this explicit form; tests/test_orb_gpu.py pins the result.
// The 16 circle pixels are packed two per register as 16-bit lanes (Q[j] = q[2j] | q[2j+1] << 16): sm_100a has single
// instructions for 2 x 16-bit min / max (VIMNMX.U16x2) and for three-input min / max (VIMNMX3.U16x2), so the 2-, 4- and
// 9-windows of all 16 arcs cost 8 instructions per stage and polarity instead of 16 to 32 scalar ones.
__device__ __forceinline__ int fast_arc_score(const uint8_t* p, int pp) {
  const int v = p[0];
  unsigned Q[8];
  Q[0] = p[3 * pp] | ((unsigned)p[3 * pp + 1] << 16);
  Q[1] = p[2 * pp + 2] | ((unsigned)p[pp + 3] << 16);
  Q[2] = p[3] | ((unsigned)p[-pp + 3] << 16);
  Q[3] = p[-2 * pp + 2] | ((unsigned)p[-3 * pp + 1] << 16);
  Q[4] = p[-3 * pp] | ((unsigned)p[-3 * pp - 1] << 16);
  Q[5] = p[-2 * pp - 2] | ((unsigned)p[-pp - 3] << 16);
  Q[6] = p[-3] | ((unsigned)p[pp - 3] << 16);
  Q[7] = p[2 * pp - 2] | ((unsigned)p[3 * pp - 1] << 16);
  unsigned mn2[8], mx2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const unsigned r1 = __funnelshift_r(Q[j], Q[(j + 1) & 7], 16);  // (q[2j+1], q[2j+2])
    mn2[j] = __vminu2(Q[j], r1);
    mx2[j] = __vmaxu2(Q[j], r1);
  }
  unsigned mn4[8], mx4[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mn4[j] = __vminu2(mn2[j], mn2[(j + 1) & 7]);
    mx4[j] = __vmaxu2(mx2[j], mx2[(j + 1) & 7]);
  }
  unsigned wmn[8], wmx[8];  // minimum / maximum of the nine pixels of arc k, two arcs per register
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    wmn[j] = __vimin3_u16x2(mn4[j], mn4[(j + 2) & 7], Q[(j + 4) & 7]);
    wmx[j] = __vimax3_u16x2(mx4[j], mx4[(j + 2) & 7], Q[(j + 4) & 7]);
  }
  // smallest window maximum, largest window minimum
  const unsigned lo2 = __vimin3_u16x2(__vimin3_u16x2(wmx[0], wmx[1], wmx[2]), __vimin3_u16x2(wmx[3], wmx[4], wmx[5]), __vminu2(wmx[6], wmx[7]));
  const unsigned hi2 = __vimax3_u16x2(__vimax3_u16x2(wmn[0], wmn[1], wmn[2]), __vimax3_u16x2(wmn[3], wmn[4], wmn[5]), __vmaxu2(wmn[6], wmn[7]));
  const int lo = (int)min(lo2 & 0xffffu, lo2 >> 16), hi = (int)max(hi2 & 0xffffu, hi2 >> 16);
  const int darker = v - lo;    // all nine pixels of the best arc are below v by at least this much
  const int brighter = hi - v;  // ... above v by at least this much
  int best = 0;
  if (darker > best) best = darker;
  if (brighter > best) best = brighter;
  return best;
}

}  // namespace plslam
