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

}  // extern "C"
