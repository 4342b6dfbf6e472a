// pl_math.cuh — device versions of the pinned scalar math shared by the ORB and line kernels.
// Every function is the operation-by-operation twin of the oracle's definition (oracle/orb_oracle.cc,
// oracle/lsd_oracle.cc): no FMA contraction (explicit _rn intrinsics; the TUs are also built with
// --fmad=false), IEEE division and square root.
#pragma once
#include <cstdint>

namespace plslam {

constexpr double PL_PI = 3.14159265358979323846;
constexpr double PL_DEG_TO_RADS = PL_PI / 180;

// cv::fastAtan2 (reference call site lib/libORB_SLAM2.so@0x7022a; arithmetic SURVEY B.4): degrees in [0, 360)
__device__ __forceinline__ float fast_atan2_dev(float y, float x) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
              p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  const float eps = (float)2.2204460492503131e-16;
  const float ax = fabsf(x), ay = fabsf(y);
  // ax >= ay: c = ay / (ax + eps), a = poly(c); else c = ax / (ay + eps), a = 90 - poly(c): written without a branch
  // (min / (max + eps) is the same quotient in both cases, also when ax == ay)
  const float c = __fdiv_rn(fminf(ax, ay), __fadd_rn(fmaxf(ax, ay), eps));
  const float c2 = __fmul_rn(c, c);
  float a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  if (ax < ay) a = __fsub_rn(90.f, a);
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

// pinned sin/cos: double Cody-Waite reduction by pi/2 + fdlibm kernel polynomials
__device__ __forceinline__ void pl_sincos_dev(double xd, double* s_out, double* c_out) {
  const double kf = rint(__dmul_rn(xd, 0.63661977236758134308));
  const int k = (int)kf;
  double r = __dsub_rn(xd, __dmul_rn(kf, 1.57079632673412561417e+00));
  r = __dsub_rn(r, __dmul_rn(kf, 6.07710050650619224932e-11));
  const double z = __dmul_rn(r, r);
  double ps = __dadd_rn(-2.50507602534068634195e-08, __dmul_rn(z, 1.58969099521155010221e-10));
  ps = __dadd_rn(2.75573137070700676789e-06, __dmul_rn(z, ps));
  ps = __dadd_rn(-1.98412698298579493134e-04, __dmul_rn(z, ps));
  ps = __dadd_rn(8.33333333332248946124e-03, __dmul_rn(z, ps));
  ps = __dadd_rn(-1.66666666666666324348e-01, __dmul_rn(z, ps));
  const double sn = __dadd_rn(r, __dmul_rn(__dmul_rn(r, z), ps));
  double pc = __dadd_rn(2.08757232129817482790e-09, __dmul_rn(z, -1.13596475577881948265e-11));
  pc = __dadd_rn(-2.75573143513906633035e-07, __dmul_rn(z, pc));
  pc = __dadd_rn(2.48015872894767294178e-05, __dmul_rn(z, pc));
  pc = __dadd_rn(-1.38888888888741095749e-03, __dmul_rn(z, pc));
  pc = __dadd_rn(4.16666666666666019037e-02, __dmul_rn(z, pc));
  const double cs = __dadd_rn(__dsub_rn(1.0, __dmul_rn(0.5, z)), __dmul_rn(__dmul_rn(z, z), pc));
  switch (k & 3) {
    case 0: *s_out = sn; *c_out = cs; break;
    case 1: *s_out = cs; *c_out = -sn; break;
    case 2: *s_out = -sn; *c_out = -cs; break;
    default: *s_out = -cs; *c_out = sn; break;
  }
}

__device__ __forceinline__ void pl_sincosf_dev(float x, float* s_out, float* c_out) {
  double s, c;
  pl_sincos_dev((double)x, &s, &c);
  *s_out = (float)s;
  *c_out = (float)c;
}

// pinned atan (x >= 0) / atan2f: fdlibm-style, see oracle/lsd_oracle.cc pl_atan
__device__ __forceinline__ double pl_atan_dev(double x) {
  const double atanhi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01,
                            1.57079632679489655800e+00};
  const double atanlo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17,
                            6.12323399573676603587e-17};
  const double aT[11] = {3.33333333333329318027e-01,  -1.99999999998764832476e-01, 1.42857142725034663711e-01,
                         -1.11111104054623557880e-01, 9.09088713343650656196e-02,  -7.69187620504482999495e-02,
                         6.66107313738753120669e-02,  -5.83357013379057348645e-02, 4.97687799461593236017e-02,
                         -3.65315727442169155270e-02, 1.62858201153657823623e-02};
  int id;
  if (x > 1e300) return __dadd_rn(atanhi[3], atanlo[3]);
  if (x < 0.4375) {
    id = -1;
  } else if (x < 1.1875) {
    if (x < 0.6875) {
      id = 0;
      x = __ddiv_rn(__dsub_rn(__dmul_rn(2.0, x), 1.0), __dadd_rn(2.0, x));
    } else {
      id = 1;
      x = __ddiv_rn(__dsub_rn(x, 1.0), __dadd_rn(x, 1.0));
    }
  } else {
    if (x < 2.4375) {
      id = 2;
      x = __ddiv_rn(__dsub_rn(x, 1.5), __dadd_rn(1.0, __dmul_rn(1.5, x)));
    } else {
      id = 3;
      x = __ddiv_rn(-1.0, x);
    }
  }
  const double z = __dmul_rn(x, x), w = __dmul_rn(z, z);
  double a = __dadd_rn(aT[8], __dmul_rn(w, aT[10]));
  a = __dadd_rn(aT[6], __dmul_rn(w, a));
  a = __dadd_rn(aT[4], __dmul_rn(w, a));
  a = __dadd_rn(aT[2], __dmul_rn(w, a));
  a = __dadd_rn(aT[0], __dmul_rn(w, a));
  const double s1 = __dmul_rn(z, a);
  double b = __dadd_rn(aT[7], __dmul_rn(w, aT[9]));
  b = __dadd_rn(aT[5], __dmul_rn(w, b));
  b = __dadd_rn(aT[3], __dmul_rn(w, b));
  b = __dadd_rn(aT[1], __dmul_rn(w, b));
  const double s2 = __dmul_rn(w, b);
  if (id < 0) return __dsub_rn(x, __dmul_rn(x, __dadd_rn(s1, s2)));
  return __dsub_rn(atanhi[id], __dsub_rn(__dsub_rn(__dmul_rn(x, __dadd_rn(s1, s2)), atanlo[id]), x));
}

__device__ __forceinline__ float pl_atan2f_dev(float yf, float xf) {
  const double y = yf, x = xf;
  if (x == 0.0 && y == 0.0) return 0.f;
  const double ax = fabs(x), ay = fabs(y);
  double a;
  if (ax == 0.0) a = 1.57079632679489655800e+00;
  else a = pl_atan_dev(__ddiv_rn(ay, ax));
  if (x < 0) a = __dsub_rn(3.14159265358979323846, a);
  if (y < 0) a = -a;
  return (float)a;
}

// logf as glibc >= 2.27 computes it (table of 1/c and log(c), degree-3 polynomial in double, one rounding to float): what
// MapPoint::PredictScale (reference lib/libORB_SLAM2.so@0x8fc7b) gets from the C library; twin of oracle/match_oracle.cc pl_logf.
__device__ __forceinline__ float pl_logf_dev(float x) {
  const double T[16][2] = {
      {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
      {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
      {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
      {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
      {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
      {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
      {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
      {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
  unsigned ix = __float_as_uint(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2 == 0) return __uint_as_float(0xff800000u);
    if (ix == 0x7f800000u) return x;
    if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return __uint_as_float(0x7fc00000u);
    ix = __float_as_uint(__fmul_rn(x, 8388608.0f));
    ix -= 23u << 23;
  }
  const unsigned tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) & 15u);
  const int k = (int)tmp >> 23;
  const unsigned iz = ix - (tmp & (0x1ffu << 23));
  const double z = (double)__uint_as_float(iz);
  const double r = __dsub_rn(__dmul_rn(z, T[i][0]), 1.0);
  const double y0 = __dadd_rn(T[i][1], __dmul_rn((double)k, 0x1.62e42fefa39efp-1));
  const double r2 = __dmul_rn(r, r);
  double y = __dadd_rn(__dmul_rn(0x1.5575b0be00b6ap-2, r), -0x1.ffffef20a4123p-2);
  y = __dadd_rn(__dmul_rn(-0x1.00ea348b88334p-2, r2), y);
  y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
  return (float)y;
}

// MapPoint::PredictScale (@0x8fc20 / @0x8fb60): ceilf(logf(mfMaxDistance / dist) / mfLogScaleFactor) clamped to the pyramid
__device__ __forceinline__ int predict_scale_dev(float maxDistance, float dist, float logScaleFactor, int nLevels) {
  int nScale = (int)ceilf(__fdiv_rn(pl_logf_dev(__fdiv_rn(maxDistance, dist)), logScaleFactor));
  if (nScale < 0) nScale = 0;
  else if (nScale >= nLevels) nScale = nLevels - 1;
  return nScale;
}

__device__ __forceinline__ int reflect101_dev(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

}  // namespace plslam
