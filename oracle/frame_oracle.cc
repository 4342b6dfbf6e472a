// frame_oracle.cc — CPU restatement of the per-frame steps that follow extraction in ORB_SLAM2::Frame::Frame
// (reference include/Frame.h:60; machine code lib/libORB_SLAM2.so@0xf9370): UndistortKeyPoints (Frame.h:266,
// call @0xfa0db), ComputeStereoFromRGBD (Frame.h:120, @0xfa0ea), ComputeImageBounds (Frame.h:270, @0xfa27e),
// AssignFeaturesToGrid / PosInGrid (Frame.h:273,110, @0xfa382, @0xf5fa0-0xf600f).
//
// TEST INFRASTRUCTURE ONLY: never linked into the product.
//
// Parity status: UndistortKeyPoints, ComputeImageBounds, ComputeStereoFromRGBD, AssignFeaturesToGrid / PosInGrid are PINNED
// AGAINST THE REFERENCE'S OWN CODE, executed from lib/libORB_SLAM2.so on a faked Frame (tests/golden/reference_code.py;
// fixtures un*, st*, fg* of tests/golden/reference_library.npz; tests/test_golden_cpu.py).
//
// cv::undistortPoints (called by UndistortKeyPoints and ComputeImageBounds with R = empty, P = mK) lives in OpenCV, which is
// not vendored; its published algorithm is restated here (normalise, 5 fixed-point iterations of the Brown model in double,
// re-project with K) and pinned bit-for-bit against cv2 4.13 in tests/test_frame_cpu.py.  Built with -ffp-contract=off.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct Calib {
  float fx, fy, cx, cy;
  float dist[5];  // k1 k2 p1 p2 k3 (mDistCoef, CV_32F)
  float bf;
};

// one point of cv::undistortPoints(src, dst, K, distCoeffs, noArray(), K), float in / float out
void undistort_point(const Calib& c, float xf, float yf, float* ox, float* oy) {
  const double fx = c.fx, fy = c.fy, cx = c.cx, cy = c.cy;
  const double ifx = 1. / fx, ify = 1. / fy;
  double k[12] = {c.dist[0], c.dist[1], c.dist[2], c.dist[3], c.dist[4], 0, 0, 0, 0, 0, 0, 0};
  double x = xf, y = yf;
  const double u = x, v = y;
  x = (x - cx) * ifx;
  y = (y - cy) * ify;
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; ++j) {
    const double r2 = x * x + y * y;
    const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
    if (icdist < 0) {
      x = (u - cx) * ifx;
      y = (v - cy) * ify;
      break;
    }
    const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
    const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
    x = (x0 - deltaX) * icdist;
    y = (y0 - deltaY) * icdist;
  }
  // RR = K * I
  const double xx = fx * x + 0.0 * y + cx;
  const double yy = 0.0 * x + fy * y + cy;
  const double ww = 1. / (0.0 * x + 0.0 * y + 1.0);
  *ox = (float)(xx * ww);
  *oy = (float)(yy * ww);
}

}  // namespace

extern "C" {

// calib: fx fy cx cy k1 k2 p1 p2 k3 bf (10 floats)
static Calib make_calib(const float* p) {
  Calib c;
  c.fx = p[0]; c.fy = p[1]; c.cx = p[2]; c.cy = p[3];
  for (int i = 0; i < 5; ++i) c.dist[i] = p[4 + i];
  c.bf = p[9];
  return c;
}

void oracle_undistort_points(const float* calib10, const float* xy, int n, float* out_xy) {
  const Calib c = make_calib(calib10);
  for (int i = 0; i < n; ++i) {
    if (c.dist[0] == 0.0f) {  // UndistortKeyPoints: mDistCoef.at<float>(0)==0.0 -> copy
      out_xy[2 * i] = xy[2 * i];
      out_xy[2 * i + 1] = xy[2 * i + 1];
    } else {
      undistort_point(c, xy[2 * i], xy[2 * i + 1], &out_xy[2 * i], &out_xy[2 * i + 1]);
    }
  }
}

// ComputeImageBounds: bounds4 = mnMinX, mnMaxX, mnMinY, mnMaxY
void oracle_image_bounds(const float* calib10, int cols, int rows, float* bounds4) {
  const Calib c = make_calib(calib10);
  if (c.dist[0] != 0.0f) {
    const float corners[8] = {0.f, 0.f, (float)cols, 0.f, 0.f, (float)rows, (float)cols, (float)rows};
    float m[8];
    for (int i = 0; i < 4; ++i) undistort_point(c, corners[2 * i], corners[2 * i + 1], &m[2 * i], &m[2 * i + 1]);
    bounds4[0] = std::min(m[0], m[4]);
    bounds4[1] = std::max(m[2], m[6]);
    bounds4[2] = std::min(m[1], m[3]);
    bounds4[3] = std::max(m[5], m[7]);
  } else {
    bounds4[0] = 0.f; bounds4[1] = (float)cols; bounds4[2] = 0.f; bounds4[3] = (float)rows;
  }
}

// One frame: keypoints (x, y) -> undistorted (x, y), mvuRight, mvDepth, grid CSR ([ix][iy] order, items ascending).
// depth: rows x cols float (imDepth after convertTo(CV_32F, mDepthMapFactor)), pitch in floats.
void oracle_frame_post(const float* calib10, const float* bounds4, const float* xy, int n, const float* depth, int cols,
                       int rows, int dpitch, float* un_xy, float* uright, float* zdepth, int32_t* grid_start,
                       int32_t* grid_items) {
  const Calib c = make_calib(calib10);
  const int GC = 64, GR = 48;
  oracle_undistort_points(calib10, xy, n, un_xy);
  for (int i = 0; i < n; ++i) {
    uright[i] = -1.f;
    zdepth[i] = -1.f;
    const int v = (int)xy[2 * i + 1], u = (int)xy[2 * i];  // truncating casts (@0xf6cd0-0xf6cd6), distorted keypoint
    const float d = depth[(size_t)v * dpitch + u];
    if (d > 0) {
      zdepth[i] = d;
      uright[i] = un_xy[2 * i] - c.bf / d;
    }
  }
  const float wInv = (float)GC / (bounds4[1] - bounds4[0]);
  const float hInv = (float)GR / (bounds4[3] - bounds4[2]);
  std::vector<std::vector<int32_t>> cells(GC * GR);
  for (int i = 0; i < n; ++i) {
    const int px = (int)roundf((un_xy[2 * i] - bounds4[0]) * wInv);
    const int py = (int)roundf((un_xy[2 * i + 1] - bounds4[2]) * hInv);
    if (px < 0 || px >= GC || py < 0 || py >= GR) continue;
    cells[px * GR + py].push_back(i);
  }
  int32_t off = 0;
  for (int cidx = 0; cidx < GC * GR; ++cidx) {
    grid_start[cidx] = off;
    for (int32_t i : cells[cidx]) grid_items[off++] = i;
  }
  grid_start[GC * GR] = off;
}

// ---------------------------------------------------------------------------------------------------------------------
// The line analogues of the Frame steps (reference include/Frame.h:107 isInFrustum(MapLine*, float), :116 GetLinesInArea, :267
// UndistortKeyLines).  PARITY UNPINNED: the reference ships these as declarations only (no source, and lib/libORB_SLAM2.so is
// stock ORB-SLAM2 without any line code).  The definitions below follow the header contracts (member names of
// include/MapLine.h:113-129, the comments at Frame.h:106,115) and the public PL-SLAM fork family the reference derives from
// (ORB-SLAM2_with_line: Frame.cc), restated from memory, with the point versions' arithmetic where they share a step.
// ---------------------------------------------------------------------------------------------------------------------
// Frame::UndistortKeyLines: both end points of every key line through cv::undistortPoints (the pinned primitive above); a zero
// k1 copies.  xy: n x 4 (startPointX, startPointY, endPointX, endPointY).
void oracle_undistort_keylines(const float* calib10, const float* xy4, int n, float* out_xy4) {
  oracle_undistort_points(calib10, xy4, 2 * n, out_xy4);
}
// Frame::GetLinesInArea(x1, y1, x2, y2, r, minLevel, maxLevel): a key line is a candidate when the squared distance between
// its mid point (KeyLine::pt) and the query's mid point is <= r * r, when (y1 - y2) / (x1 - x2) - keyline.angle <= r * 0.01
// (the family's slope test, kept as written) and when its octave passes the level filter of GetFeaturesInArea.  lines: n x 4
// (pt.x, pt.y, angle, octave as float).  Returns the number of candidates; out receives their indices in ascending order.
int oracle_get_lines_in_area(float x1, float y1, float x2, float y2, float r, int minLevel, int maxLevel, const float* lines, int n,
                             int* out, int cap) {
  const bool bCheckLevels = (minLevel > 0) || (maxLevel > 0);
  const float mx = 0.5f * (x1 + x2), my = 0.5f * (y1 + y2);
  int cnt = 0;
  for (int i = 0; i < n; ++i) {
    const float px = lines[4 * i], py = lines[4 * i + 1], ang = lines[4 * i + 2];
    const int oct = (int)lines[4 * i + 3];
    const float dx = mx - px, dy = my - py;
    const float distance = dx * dx + dy * dy;
    if (distance > r * r) continue;
    const float slope = (y1 - y2) / (x1 - x2) - ang;
    if (slope > r * 0.01f) continue;
    if (bCheckLevels) {
      if (oct < minLevel) continue;
      if (maxLevel >= 0 && oct > maxLevel) continue;
    }
    if (cnt < cap) out[cnt] = i;
    ++cnt;
  }
  return cnt;
}
// Frame::isInFrustum(MapLine* pML, float viewingCosLimit): both end points in front of the camera and inside the image bounds,
// the mid point inside the scale-invariance range and seen within the viewing-angle limit; fills mTrackProjX1/Y1/X1R, X2/Y2/X2R,
// mnTrackScaleLevel, mTrackViewCos, mbTrackInView.  Arithmetic of the point version (oracle_is_in_frustum in match_oracle.cc):
// gemm small path, float division, fused projections, norm and dot in double.  PredictScale(dist, logScaleFactor) =
// ceil(log(mfMaxDistance / dist) / logScaleFactor) clamped to [0, nLevels - 1] with the restated logf.
// sp_ep: M x 6 world end points; out_proj: M x 6 (u1, v1, u1r, u2, v2, u2r).
extern "C" int oracle_predict_scale(float maxDistance, float dist, float logScaleFactor, int nLevels);
void oracle_line_in_frustum(int M, const float* sp_ep, const float* normal, const float* distRange, const float* cam8,
                            const float* Tcw, const float* Ow, float mbf, float logScaleFactor, int nLevels, float cosLimit,
                            uint8_t* inView, float* proj6, int* level, float* viewCos) {
  const float fx = cam8[0], fy = cam8[1], cx = cam8[2], cy = cam8[3], mnMinX = cam8[4], mnMaxX = cam8[5], mnMinY = cam8[6], mnMaxY = cam8[7];
  for (int i = 0; i < M; ++i) {
    inView[i] = 0;
    for (int k = 0; k < 6; ++k) proj6[6 * i + k] = 0.f;
    level[i] = 0;
    viewCos[i] = 0.f;
    float uv[2][3];
    bool ok = true;
    for (int e = 0; e < 2 && ok; ++e) {
      const float* X = sp_ep + 6 * i + 3 * e;
      float pc[3];
      for (int r = 0; r < 3; ++r) {
        const float p0 = Tcw[r * 4] * X[0], p1 = Tcw[r * 4 + 1] * X[1], p2 = Tcw[r * 4 + 2] * X[2];
        pc[r] = (float)((double)((p0 + p1) + p2) + (double)Tcw[r * 4 + 3]);
      }
      if (pc[2] < 0.0f) { ok = false; break; }
      const float invz = 1.0f / pc[2];
      const float u = std::fmaf(pc[0] * fx, invz, cx), v = std::fmaf(pc[1] * fy, invz, cy);
      if (u < mnMinX || u > mnMaxX || v < mnMinY || v > mnMaxY) { ok = false; break; }
      uv[e][0] = u; uv[e][1] = v; uv[e][2] = std::fmaf(-invz, mbf, u);
    }
    if (!ok) continue;
    float OM[3];
    double n2 = 0;
    for (int r = 0; r < 3; ++r) {
      OM[r] = 0.5f * (sp_ep[6 * i + r] + sp_ep[6 * i + 3 + r]) - Ow[r];
      n2 += (double)OM[r] * (double)OM[r];
    }
    const float dist = (float)std::sqrt(n2);
    if (dist < 0.8f * distRange[2 * i] || dist > 1.2f * distRange[2 * i + 1]) continue;
    double dot = 0;
    for (int r = 0; r < 3; ++r) dot += (double)OM[r] * (double)normal[3 * i + r];
    const float vc = (float)(dot / (double)dist);
    if (vc < cosLimit) continue;
    level[i] = oracle_predict_scale(distRange[2 * i + 1], dist, logScaleFactor, nLevels);
    inView[i] = 1;
    for (int e = 0; e < 2; ++e)
      for (int k = 0; k < 3; ++k) proj6[6 * i + 3 * e + k] = uv[e][k];
    viewCos[i] = vc;
  }
}

}  // extern "C"
