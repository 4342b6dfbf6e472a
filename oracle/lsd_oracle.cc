// oracle/lsd_oracle.cc — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// Restatement of the line side of the reference's front-end:
//   ORB_SLAM2::LineSegment::ExtractLineSegment (include/ExtractLineSegment.h:38)
// which the reference declares but ships neither as source nor as machine code.
// Its arithmetic lives in un-vendored third-party code: OpenCV-contrib `line_descriptor`
// (LSDDetector::detect, BinaryDescriptor::compute — version matching OpenCV 3.3,
// CMakeLists.txt:20) which wraps imgproc's LineSegmentDetector (LSD_REFINE_ADV).  This file
// restates those published algorithms:
//   * LSD  (von Gioi et al.; OpenCV imgproc lsd.cpp structure): PINNED here against
//     cv2 4.13 `createLineSegmentDetector(LSD_REFINE_ADV).detect` — in COMPAT mode (libm
//     trigonometry) tests/test_oracle_cv2.py requires identical segments, widths and NFA values.
//   * LBD  (Zhang & Koch; contrib binary_descriptor.cpp structure): PARITY UNPINNED — no
//     executable copy of line_descriptor exists in this environment; the restatement follows
//     the published algorithm and is only checked for self-consistency (oracle vs CUDA).
//   * LSDDetector KeyLine filling and the fork's "keep the lsdNFeatures strongest lines by
//     response" step (comparator include/auxiliar.h:67-72): UNPINNED (header evidence only).
//
// Two modes (LsdParams.compat):
//   compat = 1  libm sincos/sincosf/atan2: what cv2 runs; bit-identical to cv2 4.13 (tests pin it).
//   compat = 0  PINNED definitions the CUDA path reproduces bit for bit: sin/cos from the shared
//               double-precision definition pl_sincos (same as the ORB oracle), pl_atan2f for
//               KeyLine::angle.  Everything else is common to both modes, including the seed
//               order (cv2 4.13 uses std::stable_sort: bin descending, then raster order) and
//               rect_nfa's row scan, both recovered from the cv2 4.13 binary because they differ
//               from the OpenCV 3.x sources (unstable sort, integer-slope polygon scan).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

const double PI = 3.14159265358979323846;
const double M_3_2_PI_ = (3 * PI) / 2;
const double M_2__PI_ = 2 * PI;
const double NOTDEF = -1024.0;
const double DEG_TO_RADS = PI / 180;
const double RELATIVE_ERROR_FACTOR = 100.0;

inline int cvRoundf(float v) { return (int)lrintf(v); }
inline int cvRoundd(double v) { return (int)lrint(v); }

// shared sin/cos definition (see oracle/orb_oracle.cc orb_sincos): double Cody-Waite + fdlibm kernels
void pl_sincos(double xd, double* s_out, double* c_out) {
  double kf = std::rint(xd * 0.63661977236758134308);
  int k = (int)kf;
  double r = xd - kf * 1.57079632673412561417e+00;
  r = r - kf * 6.07710050650619224932e-11;
  double z = r * r;
  double ps = -1.66666666666666324348e-01 +
              z * (8.33333333332248946124e-03 +
                   z * (-1.98412698298579493134e-04 +
                        z * (2.75573137070700676789e-06 +
                             z * (-2.50507602534068634195e-08 + z * 1.58969099521155010221e-10))));
  double sn = r + (r * z) * ps;
  double pc = 4.16666666666666019037e-02 +
              z * (-1.38888888888741095749e-03 +
                   z * (2.48015872894767294178e-05 +
                        z * (-2.75573143513906633035e-07 +
                             z * (2.08757232129817482790e-09 + z * -1.13596475577881948265e-11))));
  double cs = (1.0 - 0.5 * z) + (z * z) * pc;
  switch (k & 3) {
    case 0: *s_out = sn; *c_out = cs; break;
    case 1: *s_out = cs; *c_out = -sn; break;
    case 2: *s_out = -sn; *c_out = -cs; break;
    default: *s_out = -cs; *c_out = sn; break;
  }
}

// Pinned atan2 (replaces libm atan2 in KeyLine::angle): fdlibm-style atan in double (breakpoints
// 7/16, 11/16, 19/16, 39/16; odd/even split polynomial), every operation individually rounded,
// quadrant fix-up, one rounding to float.  The CUDA path evaluates the same sequence.
double pl_atan(double x) {  // x >= 0
  static const double atanhi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01, 1.57079632679489655800e+00};
  static const double atanlo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17, 6.12323399573676603587e-17};
  static const double aT[11] = {3.33333333333329318027e-01, -1.99999999998764832476e-01, 1.42857142725034663711e-01, -1.11111104054623557880e-01,
                                9.09088713343650656196e-02, -7.69187620504482999495e-02, 6.66107313738753120669e-02, -5.83357013379057348645e-02,
                                4.97687799461593236017e-02, -3.65315727442169155270e-02, 1.62858201153657823623e-02};
  int id;
  if (x > 1e300) return atanhi[3] + atanlo[3];
  if (x < 0.4375) { id = -1; }
  else if (x < 1.1875) {
    if (x < 0.6875) { id = 0; x = (2.0 * x - 1.0) / (2.0 + x); }
    else { id = 1; x = (x - 1.0) / (x + 1.0); }
  } else {
    if (x < 2.4375) { id = 2; x = (x - 1.5) / (1.0 + 1.5 * x); }
    else { id = 3; x = -1.0 / x; }
  }
  double z = x * x, w = z * z;
  double s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
  double s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
  if (id < 0) return x - x * (s1 + s2);
  return atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
}
float pl_atan2f(float yf, float xf) {
  const double y = yf, x = xf;
  if (x == 0.0 && y == 0.0) return 0.f;
  const double ax = std::fabs(x), ay = std::fabs(y);
  double a;
  if (ax == 0.0) a = 1.57079632679489655800e+00;
  else a = pl_atan(ay / ax);
  if (x < 0) a = 3.14159265358979323846 - a;
  if (y < 0) a = -a;
  return (float)a;
}

float fast_atan2(float y, float x) {  // cv::fastAtan2
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
              p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  const float eps = (float)2.2204460492503131e-16;
  float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + eps);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + eps);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    else i = 2 * (n - 1) - i;
  }
  return i;
}

// OpenCV's 8-bit fixed-point Gaussian kernel (getGaussianKernelFixedPoint_ED): k*256 rounded with
// error diffusion from the borders inwards, centre = 256 - 2*sum(side).
void gauss_table_u8(double sigma, int ksize, std::vector<int>& out) {
  std::vector<double> k(ksize);
  double sum = 0;
  const double scale2X = -0.5 / (sigma * sigma);
  for (int i = 0; i < ksize; ++i) {
    double x = i - (ksize - 1) * 0.5;
    k[i] = std::exp(scale2X * x * x);
    sum += k[i];
  }
  for (int i = 0; i < ksize; ++i) k[i] = k[i] / sum * 256.0;
  out.assign(ksize, 0);
  double err = 0;
  int side = 0;
  for (int i = 0; i < ksize / 2; ++i) {
    double v = k[i] + err;
    int r = (int)std::lrint(v);
    err = v - r;
    out[i] = out[ksize - 1 - i] = r;
    side += r;
  }
  out[ksize / 2] = 256 - 2 * side;
}

// cv::GaussianBlur on 8UC1, fixed-point path: (sum_y sum_x k[y]k[x]p + 32768) >> 16, REFLECT_101.
void gauss_blur_u8(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep, const std::vector<int>& k) {
  const int n = (int)k.size(), r = n / 2;
  std::vector<int> tmp((size_t)w * h);
  for (int y = 0; y < h; ++y) {
    const uint8_t* S = src + (size_t)y * sstep;
    int* T = tmp.data() + (size_t)y * w;
    for (int x = 0; x < w; ++x) {
      int a = 0;
      if (x >= r && x < w - r) for (int i = 0; i < n; ++i) a += k[i] * S[x + i - r];
      else for (int i = 0; i < n; ++i) a += k[i] * S[reflect101(x + i - r, w)];
      T[x] = a;
    }
  }
  std::vector<const int*> R(n);
  for (int y = 0; y < h; ++y) {
    for (int i = 0; i < n; ++i) R[i] = tmp.data() + (size_t)reflect101(y + i - r, h) * w;
    uint8_t* D = dst + (size_t)y * dstep;
    for (int x = 0; x < w; ++x) {
      int a = 32768;
      for (int i = 0; i < n; ++i) a += k[i] * R[i][x];
      D[x] = (uint8_t)(a >> 16);
    }
  }
}

// cv::resize(..., INTER_LINEAR_EXACT) on 8UC1: 8.8 fixed-point coefficients, exact products,
// one rounding at the end.
// inv_scale = the fx/fy passed to cv::resize (LSD passes SCALE = 0.8 with dsize = Size(), so the
// sampling step is exactly 1/0.8 whatever the rounded destination size is); 0 = derive from the sizes.
void linear_exact_coeffs(int ssize, int dsize, double inv_scale, std::vector<int>& ofs, std::vector<int>& c1) {
  ofs.resize(dsize);
  c1.resize(dsize);
  if (inv_scale <= 0) inv_scale = (double)dsize / ssize;
  const double scale = 1.0 / inv_scale;
  for (int d = 0; d < dsize; ++d) {
    double f = scale * (d + 0.5) - 0.5;
    int i = (int)std::floor(f);
    if (i >= 0 && ssize > 1) {
      if (i < ssize - 1) {
        ofs[d] = i;
        c1[d] = (int)std::lrint((f - i) * 256.0);
      } else {
        ofs[d] = ssize - 1;
        c1[d] = 0;
      }
    } else {
      ofs[d] = 0;
      c1[d] = 0;
    }
  }
}

void resize_linear_exact_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh, int dstep,
                            double inv_scale = 0) {
  std::vector<int> xo, xc, yo, yc;
  linear_exact_coeffs(sw, dw, inv_scale, xo, xc);
  linear_exact_coeffs(sh, dh, inv_scale, yo, yc);
  for (int y = 0; y < dh; ++y) {
    const uint8_t* S0 = src + (size_t)yo[y] * sstep;
    const uint8_t* S1 = src + (size_t)std::min(yo[y] + 1, sh - 1) * sstep;
    const int b1 = yc[y], b0 = 256 - b1;
    uint8_t* D = dst + (size_t)y * dstep;
    for (int x = 0; x < dw; ++x) {
      const int sx = xo[x], sx1 = std::min(sx + 1, sw - 1);
      const int a1 = xc[x], a0 = 256 - a1;
      const int t0 = S0[sx] * a0 + S0[sx1] * a1;
      const int t1 = S1[sx] * a0 + S1[sx1] * a1;
      D[x] = (uint8_t)((t0 * b0 + t1 * b1 + 32768) >> 16);
    }
  }
}

// ---------------------------------------------------------------------------
// LSD (imgproc LineSegmentDetector, LSD_REFINE_ADV, default parameters)
// ---------------------------------------------------------------------------
struct LsdParams {
  double scale = 0.8, sigma_scale = 0.6, quant = 2.0, ang_th = 22.5, log_eps = 0.0, density_th = 0.7;
  int n_bins = 1024;
  int compat = 0;
};

struct RegionPoint {
  int x, y;
  double angle, modgrad;
};

struct Rect {
  double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p;
};

struct NormPoint {
  int x, y, norm;
};

struct Segment {
  float x1, y1, x2, y2;
  double width, prec, nfa;
  double rect[12];  // debug: final Rect in scaled-image coordinates (x1,y1,x2,y2,width,x,y,theta,dx,dy,prec,p)
};

inline double distSq(double x1, double y1, double x2, double y2) { return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1); }
inline double dist(double x1, double y1, double x2, double y2) { return std::sqrt(distSq(x1, y1, x2, y2)); }
inline double angle_diff_signed(double a, double b) {
  double diff = a - b;
  while (diff <= -PI) diff += M_2__PI_;
  while (diff > PI) diff -= M_2__PI_;
  return diff;
}
inline double angle_diff(double a, double b) { return std::fabs(angle_diff_signed(a, b)); }
inline bool double_equal(double a, double b) {
  if (a == b) return true;
  double abs_diff = std::fabs(a - b), aa = std::fabs(a), bb = std::fabs(b);
  double abs_max = aa > bb ? aa : bb;
  if (abs_max < DBL_MIN) abs_max = DBL_MIN;
  return (abs_diff / abs_max) <= (RELATIVE_ERROR_FACTOR * DBL_EPSILON);
}
inline double log_gamma_windschitl(double x) {
  return 0.918938533204673 + (x - 0.5) * std::log(x) - x + 0.5 * x * std::log(x * std::sinh(1 / x) + 1 / (810.0 * std::pow(x, 6.0)));
}
inline double log_gamma_lanczos(double x) {
  static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
  double a = (x + 0.5) * std::log(x + 5.5) - (x + 5.5);
  double b = 0;
  for (int n = 0; n < 7; ++n) {
    a -= std::log(x + double(n));
    b += q[n] * std::pow(x, double(n));
  }
  return a + std::log(b);
}
inline double log_gamma(double x) { return x > 15.0 ? log_gamma_windschitl(x) : log_gamma_lanczos(x); }

class Lsd {
 public:
  LsdParams P;
  int img_width = 0, img_height = 0;
  double LOG_NT = 0;
  std::vector<uint8_t> scaled;
  std::vector<double> angles, modgrad;
  std::vector<uint8_t> used;
  std::vector<NormPoint> ordered_points;
  // statistics for planning / reporting
  long stat_regions = 0, stat_region_points = 0, stat_rects = 0, stat_defined = 0;

  void sc(double a, double* s, double* c) const {
    if (P.compat) { ::sincos(a, s, c); }
    else pl_sincos(a, s, c);
  }
  void scf(float a, float* s, float* c) const {
    if (P.compat) { ::sincosf(a, s, c); }
    else { double ds, dc; pl_sincos((double)a, &ds, &dc); *s = (float)ds; *c = (float)dc; }
  }

  void ll_angle(double threshold, int n_bins) {
    const int W = img_width, H = img_height;
    angles.assign((size_t)W * H, NOTDEF);
    modgrad.assign((size_t)W * H, 0.0);
    double max_grad = -1;
    for (int y = 0; y < H - 1; ++y) {
      const uint8_t* r0 = &scaled[(size_t)y * W];
      const uint8_t* r1 = &scaled[(size_t)(y + 1) * W];
      for (int x = 0; x < W - 1; ++x) {
        int DA = r1[x + 1] - r0[x];
        int BC = r0[x + 1] - r1[x];
        int gx = DA + BC, gy = DA - BC;
        double norm = std::sqrt((gx * gx + gy * gy) / 4.0);
        modgrad[(size_t)y * W + x] = norm;
        if (norm <= threshold) {
          angles[(size_t)y * W + x] = NOTDEF;
        } else {
          angles[(size_t)y * W + x] = fast_atan2(float(gx), float(-gy)) * DEG_TO_RADS;
          if (norm > max_grad) max_grad = norm;
          ++stat_defined;
        }
      }
    }
    double bin_coef = (max_grad > 0) ? double(n_bins - 1) / max_grad : 0;
    ordered_points.clear();
    ordered_points.reserve((size_t)W * H);
    for (int y = 0; y < H - 1; ++y)
      for (int x = 0; x < W - 1; ++x) {
        NormPoint p;
        p.x = x; p.y = y;
        p.norm = int(modgrad[(size_t)y * W + x] * bin_coef);
        ordered_points.push_back(p);
      }
    // cv2 4.13 sorts with std::stable_sort (merge sort with a temporary buffer, cv2.abi3.so@0xb9a7c8-0xb9a8d1),
    // i.e. bin descending then raster order: the seed order is deterministic and is the pinned order.
    std::stable_sort(ordered_points.begin(), ordered_points.end(), [](const NormPoint& a, const NormPoint& b) { return a.norm > b.norm; });
  }

  bool isAligned(int x, int y, double theta, double prec) const {
    if (x < 0 || y < 0 || x >= img_width || y >= img_height) return false;
    const double a = angles[(size_t)y * img_width + x];
    if (a == NOTDEF) return false;
    double n_theta = theta - a;
    if (n_theta < 0) n_theta = -n_theta;
    if (n_theta > M_3_2_PI_) {
      n_theta -= M_2__PI_;
      if (n_theta < 0) n_theta = -n_theta;
    }
    return n_theta <= prec;
  }

  void region_grow(int sx, int sy, std::vector<RegionPoint>& reg, double& reg_angle, double prec) {
    reg.clear();
    const int W = img_width;
    RegionPoint seed;
    seed.x = sx; seed.y = sy;
    reg_angle = angles[(size_t)sy * W + sx];
    seed.angle = reg_angle;
    seed.modgrad = modgrad[(size_t)sy * W + sx];
    reg.push_back(seed);
    float s, c;
    scf((float)reg_angle, &s, &c);  // float(std::cos(reg_angle)) in the 4.x source uses the double overload for the seed
    if (P.compat) { double ds, dc; ::sincos(reg_angle, &ds, &dc); c = (float)dc; s = (float)ds; }
    else { double ds, dc; pl_sincos(reg_angle, &ds, &dc); c = (float)dc; s = (float)ds; }
    float sumdx = c, sumdy = s;
    used[(size_t)sy * W + sx] = 1;
    for (size_t i = 0; i < reg.size(); ++i) {
      const int px = reg[i].x, py = reg[i].y;
      int xx_min = std::max(px - 1, 0), xx_max = std::min(px + 1, img_width - 1);
      int yy_min = std::max(py - 1, 0), yy_max = std::min(py + 1, img_height - 1);
      for (int yy = yy_min; yy <= yy_max; ++yy)
        for (int xx = xx_min; xx <= xx_max; ++xx) {
          uint8_t& is_used = used[(size_t)yy * W + xx];
          if (is_used != 1 && isAligned(xx, yy, reg_angle, prec)) {
            const double angle = angles[(size_t)yy * W + xx];
            is_used = 1;
            RegionPoint rp;
            rp.x = xx; rp.y = yy;
            rp.modgrad = modgrad[(size_t)yy * W + xx];
            rp.angle = angle;
            reg.push_back(rp);
            float cs, sn;
            scf(float(angle), &sn, &cs);
            sumdx += cs;
            sumdy += sn;
            reg_angle = fast_atan2(sumdy, sumdx) * DEG_TO_RADS;
          }
        }
    }
    ++stat_regions;
    stat_region_points += (long)reg.size();
  }

  double get_theta(const std::vector<RegionPoint>& reg, double x, double y, double reg_angle, double prec) const {
    double Ixx = 0.0, Iyy = 0.0, Ixy = 0.0;
    for (size_t i = 0; i < reg.size(); ++i) {
      const double regx = reg[i].x, regy = reg[i].y, weight = reg[i].modgrad;
      double dx = regx - x, dy = regy - y;
      Ixx += dy * dy * weight;
      Iyy += dx * dx * weight;
      Ixy -= dx * dy * weight;
    }
    double lambda = 0.5 * (Ixx + Iyy - std::sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
    double theta = (std::fabs(Ixx) > std::fabs(Iyy)) ? double(fast_atan2(float(lambda - Ixx), float(Ixy)))
                                                     : double(fast_atan2(float(Ixy), float(lambda - Iyy)));
    theta *= DEG_TO_RADS;
    if (angle_diff(theta, reg_angle) > prec) theta += PI;
    return theta;
  }

  void region2rect(const std::vector<RegionPoint>& reg, double reg_angle, double prec, double p, Rect& rec) const {
    double x = 0, y = 0, sum = 0;
    for (size_t i = 0; i < reg.size(); ++i) {
      const double weight = reg[i].modgrad;
      x += double(reg[i].x) * weight;
      y += double(reg[i].y) * weight;
      sum += weight;
    }
    x /= sum;
    y /= sum;
    double theta = get_theta(reg, x, y, reg_angle, prec);
    double dx, dy;
    sc(theta, &dy, &dx);
    double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
    for (size_t i = 0; i < reg.size(); ++i) {
      double regdx = double(reg[i].x) - x, regdy = double(reg[i].y) - y;
      double l = regdx * dx + regdy * dy;
      double w = -regdx * dy + regdy * dx;
      if (l > l_max) l_max = l; else if (l < l_min) l_min = l;
      if (w > w_max) w_max = w; else if (w < w_min) w_min = w;
    }
    rec.x1 = x + l_min * dx; rec.y1 = y + l_min * dy;
    rec.x2 = x + l_max * dx; rec.y2 = y + l_max * dy;
    rec.width = w_max - w_min;
    rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy; rec.prec = prec; rec.p = p;
    if (rec.width < 1.0) rec.width = 1.0;
  }

  bool reduce_region_radius(std::vector<RegionPoint>& reg, double reg_angle, double prec, double p, Rect& rec,
                            double density, double density_th) {
    double xc = double(reg[0].x), yc = double(reg[0].y);
    double radSq1 = distSq(xc, yc, rec.x1, rec.y1), radSq2 = distSq(xc, yc, rec.x2, rec.y2);
    double radSq = radSq1 > radSq2 ? radSq1 : radSq2;
    while (density < density_th) {
      radSq *= 0.75 * 0.75;
      for (size_t i = 0; i < reg.size(); ++i) {
        if (distSq(xc, yc, double(reg[i].x), double(reg[i].y)) > radSq) {
          used[(size_t)reg[i].y * img_width + reg[i].x] = 0;
          std::swap(reg[i], reg[reg.size() - 1]);
          reg.pop_back();
          --i;
        }
      }
      if (reg.size() < 2) return false;
      region2rect(reg, reg_angle, prec, p, rec);
      density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
    }
    return true;
  }

  bool refine(std::vector<RegionPoint>& reg, double reg_angle, double prec, double p, Rect& rec, double density_th) {
    double density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
    if (density >= density_th) return true;
    double xc = double(reg[0].x), yc = double(reg[0].y);
    const double ang_c = reg[0].angle;
    double sum = 0, s_sum = 0;
    int n = 0;
    for (size_t i = 0; i < reg.size(); ++i) {
      used[(size_t)reg[i].y * img_width + reg[i].x] = 0;
      if (dist(xc, yc, reg[i].x, reg[i].y) < rec.width) {
        const double angle = reg[i].angle;
        double ang_d = angle_diff_signed(angle, ang_c);
        sum += ang_d;
        s_sum += ang_d * ang_d;
        ++n;
      }
    }
    double mean_angle = sum / double(n);
    double tau = 2.0 * std::sqrt((s_sum - 2.0 * mean_angle * sum) / double(n) + mean_angle * mean_angle);
    region_grow(reg[0].x, reg[0].y, reg, reg_angle, tau);
    if (reg.size() < 2) return false;
    region2rect(reg, reg_angle, prec, p, rec);
    density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
    if (density < density_th) return reduce_region_radius(reg, reg_angle, prec, p, rec, density, density_th);
    return true;
  }

  double nfa(int n, int k, double p) const {
    if (n == 0 || k == 0) return -LOG_NT;
    if (n == k) return -LOG_NT - double(n) * std::log10(p);
    double p_term = p / (1 - p);
    double log1term = log_gamma(double(n) + 1) - log_gamma(double(k) + 1) - log_gamma(double(n - k) + 1) +
                      double(k) * std::log(p) + double(n - k) * std::log(1.0 - p);
    double term = std::exp(log1term);
    if (double_equal(term, 0)) {
      if (k > n * p) return -log1term / M_LN10 - LOG_NT;
      else return -LOG_NT;
    }
    double bin_tail = term;
    double tolerance = 0.1;
    for (int i = k + 1; i <= n; ++i) {
      double bin_term = double(n - i + 1) / double(i);
      double mult_term = bin_term * p_term;
      term *= mult_term;
      bin_tail += term;
      if (bin_term < 1) {
        double err = term * ((1 - std::pow(mult_term, double(n - i + 1))) / (1 - mult_term) - 1);
        if (err < tolerance * std::fabs(-std::log10(bin_tail) - LOG_NT) * bin_tail) break;
      }
    }
    return -std::log10(bin_tail) - LOG_NT;
  }

  // rect_nfa as compiled into cv2 4.13 (row scan between the two corner chains; recovered from the
  // binary, cv2.abi3.so@0xb96390): corners rotated so that v0 is the (min y, then min x) corner;
  // rows ceil(v0.y)..ceil(v2.y) inclusive; per row x from ceil(left chain) to trunc(right chain);
  // chain switch tests use the ceil'ed corner rows (note the < / <= asymmetry).
  double rect_nfa_rows(const Rect& r) const {
    int total_pts = 0, alg_pts = 0;
    const double hw = 0.5 * r.width;
    const double dyhw = r.dy * hw, dxhw = r.dx * hw;
    double ux[4] = {r.x1 - dyhw, r.x2 - dyhw, r.x2 + dyhw, r.x1 + dyhw};
    double uy[4] = {r.y1 + dxhw, r.y2 + dxhw, r.y2 - dxhw, r.y1 - dxhw};
    int off = 0;
    for (int i = 1; i < 4; ++i)
      if (uy[i] < uy[off] || (uy[i] == uy[off] && ux[i] < ux[off])) off = i;
    double vx[4], vy[4];
    for (int n = 0; n < 4; ++n) { vx[n] = ux[(off + n) & 3]; vy[n] = uy[(off + n) & 3]; }
    const int iy0 = (int)std::ceil(vy[0]), iy1 = (int)std::ceil(vy[1]), iy2 = (int)std::ceil(vy[2]), iy3 = (int)std::ceil(vy[3]);
    const double s01 = (iy1 == iy0) ? 0.0 : (vx[1] - vx[0]) / (vy[1] - vy[0]);
    const double s12 = (iy2 == iy1) ? 0.0 : (vx[2] - vx[1]) / (vy[2] - vy[1]);
    const double s03 = (iy3 == iy0) ? 0.0 : (vx[3] - vx[0]) / (vy[3] - vy[0]);
    const double s32 = (iy3 == iy2) ? 0.0 : (vx[2] - vx[3]) / (vy[2] - vy[3]);
    for (int y = iy0; y <= iy2; ++y) {
      if (y < 0 || y >= img_height) continue;
      const double yd = (double)y;
      const double xa = (iy1 < y) ? (yd - vy[1]) * s12 + vx[1] : (yd - vy[0]) * s01 + vx[0];
      const double xb = (iy3 <= y) ? (yd - vy[3]) * s32 + vx[3] : (yd - vy[0]) * s03 + vx[0];
      int xs = (int)std::ceil(xa);
      const int xe = (int)xb;
      if (xs < 0) xs = 0;
      for (int x = xs; x <= xe; ++x) {
        if (x >= img_width) break;
        ++total_pts;
        if (isAligned(x, y, r.theta, r.prec)) ++alg_pts;
      }
    }
    return nfa(total_pts, alg_pts, r.p);
  }

  double rect_nfa(const Rect& rec) const { return rect_nfa_rows(rec); }

  double rect_improve(Rect& rec) const {
    const double LOG_EPS = P.log_eps;
    double delta = 0.5, delta_2 = delta / 2.0;
    double log_nfa = rect_nfa(rec);
    if (log_nfa > LOG_EPS) return log_nfa;
    Rect r = rec;
    for (int n = 0; n < 5; ++n) {
      r.p /= 2;
      r.prec = r.p * PI;
      double v = rect_nfa(r);
      if (v > log_nfa) { log_nfa = v; rec = r; }
    }
    if (log_nfa > LOG_EPS) return log_nfa;
    r = rec;
    for (int n = 0; n < 5; ++n) {
      if ((r.width - delta) >= 0.5) {
        r.width -= delta;
        double v = rect_nfa(r);
        if (v > log_nfa) { rec = r; log_nfa = v; }
      }
    }
    if (log_nfa > LOG_EPS) return log_nfa;
    r = rec;
    for (int n = 0; n < 5; ++n) {
      if ((r.width - delta) >= 0.5) {
        r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2;
        r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2;
        r.width -= delta;
        double v = rect_nfa(r);
        if (v > log_nfa) { rec = r; log_nfa = v; }
      }
    }
    if (log_nfa > LOG_EPS) return log_nfa;
    r = rec;
    for (int n = 0; n < 5; ++n) {
      if ((r.width - delta) >= 0.5) {
        r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2;
        r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2;
        r.width -= delta;
        double v = rect_nfa(r);
        if (v > log_nfa) { rec = r; log_nfa = v; }
      }
    }
    if (log_nfa > LOG_EPS) return log_nfa;
    r = rec;
    for (int n = 0; n < 5; ++n) {
      if ((r.width - delta) >= 0.5) {
        r.p /= 2;
        r.prec = r.p * PI;
        double v = rect_nfa(r);
        if (v > log_nfa) { rec = r; log_nfa = v; }
      }
    }
    return log_nfa;
  }

  void detect(const uint8_t* img, int W, int H, int pitch, std::vector<Segment>& lines) {
    lines.clear();
    const double prec = PI * P.ang_th / 180;
    const double p = P.ang_th / 180;
    const double rho = P.quant / std::sin(prec);
    if (P.scale != 1) {
      const double sigma = (P.scale < 1) ? (P.sigma_scale / P.scale) : P.sigma_scale;
      const double sprec = 3;
      const unsigned int h = (unsigned int)(std::ceil(sigma * std::sqrt(2 * sprec * std::log(10.0))));
      std::vector<int> k;
      gauss_table_u8(sigma, 1 + 2 * (int)h, k);
      std::vector<uint8_t> g((size_t)W * H);
      gauss_blur_u8(img, W, H, pitch, g.data(), W, k);
      img_width = cvRoundd(W * P.scale);
      img_height = cvRoundd(H * P.scale);
      scaled.resize((size_t)img_width * img_height);
      resize_linear_exact_u8(g.data(), W, H, W, scaled.data(), img_width, img_height, img_width, P.scale);
    } else {
      img_width = W; img_height = H;
      scaled.resize((size_t)W * H);
      for (int y = 0; y < H; ++y) std::memcpy(&scaled[(size_t)y * W], img + (size_t)y * pitch, W);
    }
    ll_angle(rho, P.n_bins);
    LOG_NT = 5 * (std::log10(double(img_width)) + std::log10(double(img_height))) / 2 + std::log10(11.0);
    const size_t min_reg_size = size_t(-LOG_NT / std::log10(p));
    used.assign((size_t)img_width * img_height, 0);
    std::vector<RegionPoint> reg;
    for (size_t i = 0, n = ordered_points.size(); i < n; ++i) {
      const int px = ordered_points[i].x, py = ordered_points[i].y;
      const size_t idx = (size_t)py * img_width + px;
      if (used[idx] == 0 && angles[idx] != NOTDEF) {
        double reg_angle;
        region_grow(px, py, reg, reg_angle, prec);
        if (reg.size() < min_reg_size) continue;
        Rect rec;
        region2rect(reg, reg_angle, prec, p, rec);
        if (!refine(reg, reg_angle, prec, p, rec, P.density_th)) continue;
        ++stat_rects;
        double log_nfa = rect_improve(rec);
        if (log_nfa <= P.log_eps) continue;
        Segment s;
        { const double rr[12] = {rec.x1, rec.y1, rec.x2, rec.y2, rec.width, rec.x, rec.y, rec.theta, rec.dx, rec.dy, rec.prec, rec.p};
          std::memcpy(s.rect, rr, sizeof(rr)); }
        rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
        if (P.scale != 1) {
          rec.x1 /= P.scale; rec.y1 /= P.scale; rec.x2 /= P.scale; rec.y2 /= P.scale;
          rec.width /= P.scale;
        }
        s.x1 = float(rec.x1); s.y1 = float(rec.y1); s.x2 = float(rec.x2); s.y2 = float(rec.y2);
        s.width = rec.width; s.prec = rec.p; s.nfa = log_nfa;
        lines.push_back(s);
      }
    }
  }
};

// ---------------------------------------------------------------------------
// contrib line_descriptor: KeyLine filling (LSDDetector::detect) and LBD (BinaryDescriptor::compute)
// ---------------------------------------------------------------------------
struct KeyLine {  // cv::line_descriptor::KeyLine field order
  float angle;
  int class_id, octave;
  float pt_x, pt_y, response, size;
  float startPointX, startPointY, endPointX, endPointY;
  float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
  float lineLength;
  int numOfPixels;
};

void fill_keylines(const std::vector<Segment>& segs, int W, int H, int compat, std::vector<KeyLine>& out) {
  out.clear();
  int class_counter = -1;
  for (const Segment& s : segs) {
    float e[4] = {s.x1, s.y1, s.x2, s.y2};
    // checkLineExtremes
    if (e[0] < 0) e[0] = 0; if (e[0] >= W) e[0] = (float)W - 1.0f;
    if (e[2] < 0) e[2] = 0; if (e[2] >= W) e[2] = (float)W - 1.0f;
    if (e[1] < 0) e[1] = 0; if (e[1] >= H) e[1] = (float)H - 1.0f;
    if (e[3] < 0) e[3] = 0; if (e[3] >= H) e[3] = (float)H - 1.0f;
    KeyLine kl;
    const float octaveScale = 1.0f;  // pow((float)scale, 0): single octave (numOctaves = 1)
    kl.startPointX = e[0] * octaveScale; kl.startPointY = e[1] * octaveScale;
    kl.endPointX = e[2] * octaveScale; kl.endPointY = e[3] * octaveScale;
    kl.sPointInOctaveX = e[0]; kl.sPointInOctaveY = e[1];
    kl.ePointInOctaveX = e[2]; kl.ePointInOctaveY = e[3];
    // (float) sqrt(pow(e0-e2, 2) + pow(e1-e3, 2)) evaluated in double on float differences
    const double dx = (double)(e[0] - e[2]), dy = (double)(e[1] - e[3]);
    kl.lineLength = (float)std::sqrt(dx * dx + dy * dy);
    // LineIterator (8-connected) between the rounded end points: count = max(|dx|,|dy|) + 1
    const int x0 = cvRoundf(e[0]), y0 = cvRoundf(e[1]), x1 = cvRoundf(e[2]), y1 = cvRoundf(e[3]);
    kl.numOfPixels = std::max(std::abs(x1 - x0), std::abs(y1 - y0)) + 1;
    if (compat) kl.angle = (float)std::atan2((double)(kl.endPointY - kl.startPointY), (double)(kl.endPointX - kl.startPointX));
    else kl.angle = pl_atan2f(kl.endPointY - kl.startPointY, kl.endPointX - kl.startPointX);
    kl.class_id = ++class_counter;
    kl.octave = 0;
    kl.size = (kl.endPointX - kl.startPointX) * (kl.endPointY - kl.startPointY);
    kl.response = kl.lineLength / (float)std::max(W, H);
    kl.pt_x = (kl.endPointX + kl.startPointX) / 2;
    kl.pt_y = (kl.endPointY + kl.startPointY) / 2;
    out.push_back(kl);
  }
}

const int kCombinations[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6},
                                  {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8},
                                  {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};
constexpr int LBD_W = 7, LBD_BANDS = 9;

struct LbdTables {
  double gaussCoefL[LBD_W * 3], gaussCoefG[LBD_W * LBD_BANDS];
  LbdTables() {
    double u = (LBD_W * 3 - 1) / 2;
    double sigma = (LBD_W * 2 + 1) / 2;  // integer division as published: 7
    double invsigma2 = -1 / (2 * sigma * sigma);
    for (int i = 0; i < LBD_W * 3; ++i) { double dis = i - u; gaussCoefL[i] = std::exp(dis * dis * invsigma2); }
    u = (LBD_BANDS * LBD_W - 1) / 2;
    sigma = u;
    invsigma2 = -1 / (2 * sigma * sigma);
    for (int i = 0; i < LBD_BANDS * LBD_W; ++i) { double dis = i - u; gaussCoefG[i] = std::exp(dis * dis * invsigma2); }
  }
};

// cv::Sobel(img, CV_16S, 1, 0, 3) / (0, 1, 3), BORDER_REFLECT_101, at one pixel
inline void sobel_at(const uint8_t* img, int W, int H, int pitch, int x, int y, int& dx, int& dy) {
  const int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W), ym = reflect101(y - 1, H), yp = reflect101(y + 1, H);
  const uint8_t* r0 = img + (size_t)ym * pitch;
  const uint8_t* r1 = img + (size_t)y * pitch;
  const uint8_t* r2 = img + (size_t)yp * pitch;
  dx = (r0[xp] - r0[xm]) + 2 * (r1[xp] - r1[xm]) + (r2[xp] - r2[xm]);
  dy = (r2[xm] - r0[xm]) + 2 * (r2[x] - r0[x]) + (r2[xp] - r0[xp]);
}

// BinaryDescriptor::computeLBD for one line + binaryConversion over the 32 band pairs
void lbd_one(const uint8_t* img, int W, int H, int pitch, const KeyLine& kl, int compat, uint8_t* desc32, float* desc72) {
  static const LbdTables T;
  const short heightOfLSP = LBD_W * LBD_BANDS;
  float pgdLBandSum[LBD_BANDS] = {0}, ngdLBandSum[LBD_BANDS] = {0}, pgdL2BandSum[LBD_BANDS] = {0}, ngdL2BandSum[LBD_BANDS] = {0};
  float pgdOBandSum[LBD_BANDS] = {0}, ngdOBandSum[LBD_BANDS] = {0}, pgdO2BandSum[LBD_BANDS] = {0}, ngdO2BandSum[LBD_BANDS] = {0};
  const short halfHeight = (heightOfLSP - 1) / 2;
  const short imageWidth = (short)(W - 1), imageHeight = (short)(H - 1);
  const short lengthOfLSP = (short)kl.numOfPixels;
  const short halfWidth = (lengthOfLSP - 1) / 2;
  const float lineMiddlePointX = (float)(0.5 * (kl.sPointInOctaveX + kl.ePointInOctaveX));
  const float lineMiddlePointY = (float)(0.5 * (kl.sPointInOctaveY + kl.ePointInOctaveY));
  float dL[2], dO[2];
  {
    double s, c;
    if (compat) { c = std::cos((double)kl.angle); s = std::sin((double)kl.angle); }
    else pl_sincos((double)kl.angle, &s, &c);
    dL[0] = (float)c; dL[1] = (float)s;
  }
  dO[0] = -dL[1]; dO[1] = dL[0];
  float sCorX0 = -dL[0] * halfWidth + dL[1] * halfHeight + lineMiddlePointX;
  float sCorY0 = -dL[1] * halfWidth - dL[0] * halfHeight + lineMiddlePointY;
  for (short hID = 0; hID < heightOfLSP; ++hID) {
    float sCorX = sCorX0, sCorY = sCorY0;
    float pgdLRowSum = 0, ngdLRowSum = 0, pgdORowSum = 0, ngdORowSum = 0;
    for (short wID = 0; wID < lengthOfLSP; ++wID) {
      short tempCor = (short)std::round(sCorX);
      short xCor = (tempCor < 0) ? 0 : (tempCor > imageWidth) ? imageWidth : tempCor;
      tempCor = (short)std::round(sCorY);
      short yCor = (tempCor < 0) ? 0 : (tempCor > imageHeight) ? imageHeight : tempCor;
      int dx, dy;
      sobel_at(img, W, H, pitch, xCor, yCor, dx, dy);
      float gDL = (float)dx * dL[0] + (float)dy * dL[1];
      float gDO = (float)dx * dO[0] + (float)dy * dO[1];
      if (gDL > 0) pgdLRowSum += gDL; else ngdLRowSum -= gDL;
      if (gDO > 0) pgdORowSum += gDO; else ngdORowSum -= gDO;
      sCorX += dL[0];
      sCorY += dL[1];
    }
    sCorX0 -= dL[1];
    sCorY0 += dL[0];
    float coefInGaussion = (float)T.gaussCoefG[hID];
    pgdLRowSum = coefInGaussion * pgdLRowSum;
    ngdLRowSum = coefInGaussion * ngdLRowSum;
    float pgdL2RowSum = pgdLRowSum * pgdLRowSum, ngdL2RowSum = ngdLRowSum * ngdLRowSum;
    pgdORowSum = coefInGaussion * pgdORowSum;
    ngdORowSum = coefInGaussion * ngdORowSum;
    float pgdO2RowSum = pgdORowSum * pgdORowSum, ngdO2RowSum = ngdORowSum * ngdORowSum;
    auto acc = [&](int band, float coef) {
      pgdLBandSum[band] += coef * pgdLRowSum;
      ngdLBandSum[band] += coef * ngdLRowSum;
      pgdL2BandSum[band] += coef * coef * pgdL2RowSum;
      ngdL2BandSum[band] += coef * coef * ngdL2RowSum;
      pgdOBandSum[band] += coef * pgdORowSum;
      ngdOBandSum[band] += coef * ngdORowSum;
      pgdO2BandSum[band] += coef * coef * pgdO2RowSum;
      ngdO2BandSum[band] += coef * coef * ngdO2RowSum;
    };
    short bandID = (short)(hID / LBD_W);
    acc(bandID, (float)T.gaussCoefL[hID % LBD_W + LBD_W]);
    bandID--;
    if (bandID >= 0) acc(bandID, (float)T.gaussCoefL[hID % LBD_W + 2 * LBD_W]);
    bandID = bandID + 2;
    if (bandID < LBD_BANDS) acc(bandID, (float)T.gaussCoefL[hID % LBD_W]);
  }
  float desVec[LBD_BANDS * 8];
  const float invN2 = (float)(1.0 / (LBD_W * 2.0)), invN3 = (float)(1.0 / (LBD_W * 3.0));
  for (short bandID = 0; bandID < LBD_BANDS; ++bandID) {
    const float invN = (bandID == 0 || bandID == LBD_BANDS - 1) ? invN2 : invN3;
    const int desID = bandID * 8;
    float temp = pgdLBandSum[bandID] * invN;
    desVec[desID] = temp;
    desVec[desID + 4] = std::sqrt(pgdL2BandSum[bandID] * invN - temp * temp);
    temp = ngdLBandSum[bandID] * invN;
    desVec[desID + 1] = temp;
    desVec[desID + 5] = std::sqrt(ngdL2BandSum[bandID] * invN - temp * temp);
    temp = pgdOBandSum[bandID] * invN;
    desVec[desID + 2] = temp;
    desVec[desID + 6] = std::sqrt(pgdO2BandSum[bandID] * invN - temp * temp);
    temp = ngdOBandSum[bandID] * invN;
    desVec[desID + 3] = temp;
    desVec[desID + 7] = std::sqrt(ngdO2BandSum[bandID] * invN - temp * temp);
  }
  float tempM = 0, tempS = 0;
  for (int b = 0; b < LBD_BANDS; ++b) {
    const float* d = desVec + 8 * b;
    tempM += d[0] * d[0]; tempM += d[1] * d[1]; tempM += d[2] * d[2]; tempM += d[3] * d[3];
    tempS += d[4] * d[4]; tempS += d[5] * d[5]; tempS += d[6] * d[6]; tempS += d[7] * d[7];
  }
  tempM = 1 / std::sqrt(tempM);
  tempS = 1 / std::sqrt(tempS);
  for (int b = 0; b < LBD_BANDS; ++b) {
    float* d = desVec + 8 * b;
    d[0] *= tempM; d[1] *= tempM; d[2] *= tempM; d[3] *= tempM;
    d[4] *= tempS; d[5] *= tempS; d[6] *= tempS; d[7] *= tempS;
  }
  for (int i = 0; i < LBD_BANDS * 8; ++i) if (desVec[i] > 0.4) desVec[i] = (float)0.4;
  float temp = 0;
  for (int i = 0; i < LBD_BANDS * 8; ++i) temp += desVec[i] * desVec[i];
  temp = 1 / std::sqrt(temp);
  for (int i = 0; i < LBD_BANDS * 8; ++i) desVec[i] = desVec[i] * temp;
  if (desc72) std::memcpy(desc72, desVec, sizeof(desVec));
  for (int comb = 0; comb < 32; ++comb) {
    const float* f1 = &desVec[8 * kCombinations[comb][0]];
    const float* f2 = &desVec[8 * kCombinations[comb][1]];
    uint8_t result = 0;
    for (int i = 0; i < 8; ++i) if (f1[i] > f2[i]) result += (uint8_t)(1 << i);
    desc32[comb] = result;
  }
}

}  // namespace

// ---------------------------------------------------------------------------
// C entry points (ctypes)
// ---------------------------------------------------------------------------
extern "C" {

void oracle_gauss_table_u8(double sigma, int ksize, int* out) {
  std::vector<int> k;
  gauss_table_u8(sigma, ksize, k);
  for (int i = 0; i < ksize; ++i) out[i] = k[i];
}
void oracle_gauss_blur_u8(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep, const int* k, int ksize) {
  std::vector<int> kk(k, k + ksize);
  gauss_blur_u8(src, w, h, sstep, dst, dstep, kk);
}
void oracle_resize_linear_exact_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh, int dstep, double inv_scale) {
  resize_linear_exact_u8(src, sw, sh, sstep, dst, dw, dh, dstep, inv_scale);
}
void oracle_pl_sincos(double x, double* s, double* c) { pl_sincos(x, s, c); }
float oracle_pl_atan2f(float y, float x) { return pl_atan2f(y, x); }
double oracle_lsd_nfa(int n, int k, double p, double log_nt) { Lsd d; d.LOG_NT = log_nt; return d.nfa(n, k, p); }
// debug: LSD with the final rectangles (n x 19 doubles: 7 as oracle_lsd_detect + 12 rect fields)
int oracle_lsd_detect_dbg(const uint8_t* img, int W, int H, int pitch, int compat, double* out, int cap) {
  Lsd d;
  d.P.compat = compat;
  std::vector<Segment> segs;
  d.detect(img, W, H, pitch, segs);
  int n = (int)std::min<size_t>(segs.size(), cap);
  for (int i = 0; i < n; ++i) {
    double* o = out + 19 * i;
    o[0] = segs[i].x1; o[1] = segs[i].y1; o[2] = segs[i].x2; o[3] = segs[i].y2;
    o[4] = segs[i].width; o[5] = segs[i].prec; o[6] = segs[i].nfa;
    std::memcpy(o + 7, segs[i].rect, 12 * sizeof(double));
  }
  return (int)segs.size();
}

// LSD detect: out = n x 7 doubles (x1, y1, x2, y2 as float values, width, prec, nfa). stats: 4 longs (may be NULL).
int oracle_lsd_detect(const uint8_t* img, int W, int H, int pitch, int compat, double* out, int cap, long* stats) {
  Lsd d;
  d.P.compat = compat;
  std::vector<Segment> segs;
  d.detect(img, W, H, pitch, segs);
  int n = (int)std::min<size_t>(segs.size(), cap);
  for (int i = 0; i < n; ++i) {
    double* o = out + 7 * i;
    o[0] = segs[i].x1; o[1] = segs[i].y1; o[2] = segs[i].x2; o[3] = segs[i].y2;
    o[4] = segs[i].width; o[5] = segs[i].prec; o[6] = segs[i].nfa;
  }
  if (stats) { stats[0] = d.stat_regions; stats[1] = d.stat_region_points; stats[2] = d.stat_rects; stats[3] = d.stat_defined; }
  return (int)segs.size();
}

// scaled image / level-line angles (degrees as float, -1024 = NOTDEF) of the LSD front half, for stage parity
int oracle_lsd_stage(const uint8_t* img, int W, int H, int pitch, uint8_t* scaled_out, double* angles_out, double* modgrad_out, int* wh) {
  Lsd d;
  std::vector<Segment> segs;
  d.detect(img, W, H, pitch, segs);
  wh[0] = d.img_width; wh[1] = d.img_height;
  size_t n = (size_t)d.img_width * d.img_height;
  if (scaled_out) std::memcpy(scaled_out, d.scaled.data(), n);
  if (angles_out) std::memcpy(angles_out, d.angles.data(), n * sizeof(double));
  if (modgrad_out) std::memcpy(modgrad_out, d.modgrad.data(), n * sizeof(double));
  return (int)segs.size();
}

// LineSegment::ExtractLineSegment (ExtractLineSegment.h:38): LSD -> KeyLines -> keep the `max_lines`
// strongest by response (stable) -> LBD -> line functions.  keylines: cap x 68 B, desc: cap x 32,
// funcs: cap x 3 doubles.  Returns the number of lines kept; *n_detected = LSD segment count.
int oracle_extract_lines(const uint8_t* img, int W, int H, int pitch, int compat, int max_lines, void* keylines,
                         uint8_t* desc, double* funcs, int cap, int* n_detected) {
  Lsd d;
  d.P.compat = compat;
  std::vector<Segment> segs;
  d.detect(img, W, H, pitch, segs);
  std::vector<KeyLine> kls;
  fill_keylines(segs, W, H, compat, kls);
  if (n_detected) *n_detected = (int)kls.size();
  if (max_lines > 0 && (int)kls.size() > max_lines) {
    // sort(keylines, sort_lines_by_response()) (auxiliar.h:67-72) + resize + re-index; ties pinned to detection order
    std::stable_sort(kls.begin(), kls.end(), [](const KeyLine& a, const KeyLine& b) { return a.response > b.response; });
    kls.resize(max_lines);
    for (int i = 0; i < max_lines; ++i) kls[i].class_id = i;
  }
  int n = (int)kls.size();
  if (n > cap) return -n;
  for (int i = 0; i < n; ++i) {
    std::memcpy((char*)keylines + (size_t)i * sizeof(KeyLine), &kls[i], sizeof(KeyLine));
    lbd_one(img, W, H, pitch, kls[i], compat, desc + (size_t)i * 32, nullptr);
    // lineF = sp x ep / sqrt(l0^2 + l1^2)  (Eigen Vector3d, doubles)
    const double x1 = kls[i].startPointX, y1 = kls[i].startPointY, x2 = kls[i].endPointX, y2 = kls[i].endPointY;
    double l0 = y1 * 1.0 - 1.0 * y2, l1 = 1.0 * x2 - x1 * 1.0, l2 = x1 * y2 - y1 * x2;
    const double nrm = std::sqrt(l0 * l0 + l1 * l1);
    funcs[3 * i] = l0 / nrm; funcs[3 * i + 1] = l1 / nrm; funcs[3 * i + 2] = l2 / nrm;
  }
  return n;
}

// LBD of given keylines (72 floats + 32 bytes each)
void oracle_lbd(const uint8_t* img, int W, int H, int pitch, const void* keylines, int n, int compat, uint8_t* desc32, float* desc72) {
  const KeyLine* k = (const KeyLine*)keylines;
  for (int i = 0; i < n; ++i) lbd_one(img, W, H, pitch, k[i], compat, desc32 + (size_t)i * 32, desc72 ? desc72 + (size_t)i * 72 : nullptr);
}

}  // extern "C"
