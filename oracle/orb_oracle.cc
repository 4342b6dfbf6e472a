// oracle/orb_oracle.cc — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A scalar C++ restatement of ORB_SLAM2::ORBextractor as shipped by the reference
// (maxee1900/RGBD-PL-SLAM).  The reference has no source for this class: only the
// declaration (include/ORBextractor.h:45-111) and machine code in lib/libORB_SLAM2.so.
// Every function below cites the header line and/or the binary address it follows
// (addresses as listed in SURVEY.md Appendix A), and for the OpenCV primitives the
// model of SURVEY.md Appendix B, which tests/test_oracle_cv2.py pins against
// cv2 4.13 (resize, GaussianBlur, FAST, fastAtan2) bit for bit.
//
// Parity status: PINNED AGAINST THE REFERENCE ITSELF.  ORB_SLAM2::ORBextractor::operator() is executed from the shipped
// lib/libORB_SLAM2.so (tests/golden/reference_code.py: the library is dlopen'ed over generated stub dependencies, its OpenCV
// entry points served by ABI-exact shims over the cv2-pinned primitives) and this restatement reproduces every keypoint
// field and descriptor byte, order included (fixtures tests/golden/reference_library.npz, tests/test_golden_cpu.py); the
// constructor tables, DistributeOctTree and ComputeKeyPointsOctTree are pinned the same way on their own.  The primitives are
// pinned against cv2 4.13.  Choices the reference leaves to its environment (DESIGN.md "Pinned choices"): blur table (runtime
// parameter, default = cv2 4.13 table), sin/cos definition (orb_sincos below; coincides with this container's glibc on every
// argument that occurred), quad-tree tie-break (size, creation sequence) = the reference's heap-address order under a
// monotonic allocator.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.  The product (rgbd-pl-slam_b200/) never does.
//
// Build: see oracle/Makefile (g++ -O3 -march=native -ffp-contract=off).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <vector>

namespace {

static const int kPattern[1024] = {
#include "orb_pattern.inc"
};

constexpr int PATCH_SIZE = 31;       // rodata 31.0f @0x1269c0
constexpr int HALF_PATCH_SIZE = 15;  // hp^2 = 225.0 @0x1269d0
constexpr int EDGE_THRESHOLD = 19;   // @0x705e1
constexpr int MIN_BORDER = EDGE_THRESHOLD - 3;  // 16, @0x760c6

inline int cvRoundf(float v) { return (int)lrintf(v); }    // vcvtss2si: round-half-even
inline int cvRoundd(double v) { return (int)lrint(v); }

struct KeyPoint {  // cv::KeyPoint layout, 28 B (stores @0x7656d-0x7659d)
  float x, y, size, angle, response;
  int octave, class_id;
};

// ---------------------------------------------------------------------------
// OpenCV primitive models
// ---------------------------------------------------------------------------

// B.1  cv::resize(..., INTER_LINEAR) on 8UC1 (call site lib/libORB_SLAM2.so@0x70b07).
void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh,
                      int dstep) {
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<short> xa(2 * dw), ya(2 * dh);
  auto coeffs = [](int d, int ssize, int dsize, int& ofs, short& c0, short& c1) {
    double inv_scale = (double)dsize / ssize;
    double scale = 1.0 / inv_scale;
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    ofs = s;
    c0 = (short)cvRoundf((1.f - f) * 2048.f);
    c1 = (short)cvRoundf(f * 2048.f);
  };
  for (int x = 0; x < dw; ++x) coeffs(x, sw, dw, xofs[x], xa[2 * x], xa[2 * x + 1]);
  for (int y = 0; y < dh; ++y) coeffs(y, sh, dh, yofs[y], ya[2 * y], ya[2 * y + 1]);
  std::vector<int> r0(dw), r1(dw);
  auto hrow = [&](int sy, std::vector<int>& out) {
    const uint8_t* S = src + (size_t)sy * sstep;
    for (int x = 0; x < dw; ++x) {
      int sx = xofs[x];
      int a = S[sx] * xa[2 * x];
      if (xa[2 * x + 1]) a += S[sx + 1] * xa[2 * x + 1];
      out[x] = a;
    }
  };
  for (int y = 0; y < dh; ++y) {
    int sy = yofs[y];
    hrow(sy, r0);
    hrow(std::min(sy + 1, sh - 1), r1);
    int b0 = ya[2 * y], b1 = ya[2 * y + 1];
    uint8_t* D = dst + (size_t)y * dstep;
    for (int x = 0; x < dw; ++x)
      D[x] = (uint8_t)((((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2);
  }
}

inline int reflect101(int i, int n) {
  if (i < 0) return -i;
  if (i >= n) return 2 * (n - 1) - i;
  return i;
}

// B.2  cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on 8UC1 (call @0x77487):
// dst = (sum_y sum_x k[y] k[x] p + 32768) >> 16 with an integer 7-tap table of sum 256.
void blur7_u8(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep, const int k[7]) {
  std::vector<int> tmp((size_t)w * h);
  for (int y = 0; y < h; ++y) {
    const uint8_t* S = src + (size_t)y * sstep;
    int* T = tmp.data() + (size_t)y * w;
    for (int x = 0; x < w; ++x) {
      int a = 0;
      if (x >= 3 && x < w - 3) {
        for (int i = 0; i < 7; ++i) a += k[i] * S[x + i - 3];
      } else {
        for (int i = 0; i < 7; ++i) a += k[i] * S[reflect101(x + i - 3, w)];
      }
      T[x] = a;
    }
  }
  for (int y = 0; y < h; ++y) {
    const int* R[7];
    for (int i = 0; i < 7; ++i) R[i] = tmp.data() + (size_t)reflect101(y + i - 3, h) * w;
    uint8_t* D = dst + (size_t)y * dstep;
    for (int x = 0; x < w; ++x) {
      int a = 32768;
      for (int i = 0; i < 7; ++i) a += k[i] * R[i][x];
      D[x] = (uint8_t)(a >> 16);
    }
  }
}

// B.3  cv::FAST(img, kps, th, nonmaxSuppression=true), TYPE_9_16, on a sub-image
// (call sites @0x763d4 / @0x76753).  Returns keypoints row-major; pt in sub-image coords.
static const int kCircle[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},  {3, 0},  {3, -1}, {2, -2}, {1, -3},
                                   {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

// max over the 16 arcs of 9 contiguous circle pixels of min |v - p| for one polarity, both
// polarities; a pixel is a corner at threshold t iff fast_arc_score > t, response = score-1.
inline int fast_arc_score(const uint8_t* p, const int off[16]) {
  int v = p[0];
  int d[25];
  for (int k = 0; k < 16; ++k) d[k] = v - p[off[k]];
  for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
  int best = 0;
  for (int k = 0; k < 16; ++k) {
    int mn = d[k], mx = d[k];
    for (int i = 1; i < 9; ++i) {
      mn = std::min(mn, d[k + i]);
      mx = std::max(mx, d[k + i]);
    }
    best = std::max(best, std::max(mn, -mx));
  }
  return best;
}

struct FastPt { int x, y, score; };

void fast9_nms(const uint8_t* img, int w, int h, int step, int th, std::vector<FastPt>& out) {
  out.clear();
  if (w < 7 || h < 7) return;
  int off[16];
  for (int k = 0; k < 16; ++k) off[k] = kCircle[k][1] * step + kCircle[k][0];
  std::vector<int> sc((size_t)w * h, 0);
  bool any = false;
  for (int y = 3; y < h - 3; ++y) {
    const uint8_t* row = img + (size_t)y * step;
    for (int x = 3; x < w - 3; ++x) {
      const uint8_t* p = row + x;
      int v = p[0];
      // quick reject: every 9-arc contains circle pixel 0 or 8 (and 4 or 12)
      int a0 = std::abs(v - p[off[0]]), a8 = std::abs(v - p[off[8]]);
      if (a0 <= th && a8 <= th) continue;
      int a4 = std::abs(v - p[off[4]]), a12 = std::abs(v - p[off[12]]);
      if (a4 <= th && a12 <= th) continue;
      int s = fast_arc_score(p, off);
      if (s > th) { sc[(size_t)y * w + x] = s - 1 + 1; any = true; }  // store score+1 > 0 marks a corner
    }
  }
  if (!any) return;
  // NMS: keep iff response strictly greater than all 8 neighbours' responses (non-corner = 0).
  // Stored value is response+1 for corners, 0 otherwise; response >= th >= 1 so ordering is preserved
  // (a corner with response r beats a non-corner iff r > 0).
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; ++x) {
      int s = sc[(size_t)y * w + x];
      if (!s) continue;
      bool keep = true;
      for (int dy = -1; dy <= 1 && keep; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          if (!dx && !dy) continue;
          if (sc[(size_t)(y + dy) * w + x + dx] >= s) { keep = false; break; }
        }
      if (keep) out.push_back({x, y, s - 1});
    }
}

// B.4  cv::fastAtan2(y, x) (call @0x7022a): float polynomial, no FMA.
float fast_atan2(float y, float x) {
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

// Pinned sin/cos definition (replaces glibc sincosf @0x77803, whose result is libm-version
// dependent).  Double-precision Cody-Waite reduction by pi/2 plus the fdlibm
// kernel polynomials, every operation individually rounded (no FMA; this TU is built with
// -ffp-contract=off), result rounded once to float.  The CUDA path evaluates the same
// sequence with __dmul_rn/__dadd_rn.
void orb_sincos(float x, float* s_out, float* c_out) {
  double xd = (double)x;
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
  double s, c;
  switch (k & 3) {
    case 0: s = sn; c = cs; break;
    case 1: s = cs; c = -sn; break;
    case 2: s = -sn; c = -cs; break;
    default: s = -cs; c = sn; break;
  }
  *s_out = (float)s;
  *c_out = (float)c;
}

// ---------------------------------------------------------------------------
// ORBextractor
// ---------------------------------------------------------------------------

struct Level {
  int w = 0, h = 0;
  std::vector<uint8_t> img, blurred;
};

struct ExtractorNode {  // include/ORBextractor.h:32-43
  std::vector<KeyPoint> vKeys;
  int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
  std::list<ExtractorNode>::iterator lit;
  bool bNoMore = false;
  // DivideNode, lib/libORB_SLAM2.so@0x70c60
  void DivideNode(ExtractorNode& n1, ExtractorNode& n2, ExtractorNode& n3, ExtractorNode& n4) const {
    const int halfX = (int)std::ceil((float)(URx - ULx) / 2);
    const int halfY = (int)std::ceil((float)(BRy - ULy) / 2);
    n1.ULx = ULx; n1.ULy = ULy; n1.URx = ULx + halfX; n1.URy = ULy;
    n1.BLx = ULx; n1.BLy = ULy + halfY; n1.BRx = ULx + halfX; n1.BRy = ULy + halfY;
    n2.ULx = n1.URx; n2.ULy = n1.URy; n2.URx = URx; n2.URy = URy;
    n2.BLx = n1.BRx; n2.BLy = n1.BRy; n2.BRx = URx; n2.BRy = ULy + halfY;
    n3.ULx = n1.BLx; n3.ULy = n1.BLy; n3.URx = n1.BRx; n3.URy = n1.BRy;
    n3.BLx = BLx; n3.BLy = BLy; n3.BRx = n1.BRx; n3.BRy = BLy;
    n4.ULx = n3.URx; n4.ULy = n3.URy; n4.URx = n2.BRx; n4.URy = n2.BRy;
    n4.BLx = n3.BRx; n4.BLy = n3.BRy; n4.BRx = BRx; n4.BRy = BRy;
    for (const KeyPoint& kp : vKeys) {
      if (kp.x < (float)n1.URx) {
        if (kp.y < (float)n1.BRy) n1.vKeys.push_back(kp); else n3.vKeys.push_back(kp);
      } else if (kp.y < (float)n1.BRy) n2.vKeys.push_back(kp);
      else n4.vKeys.push_back(kp);
    }
    if (n1.vKeys.size() == 1) n1.bNoMore = true;
    if (n2.vKeys.size() == 1) n2.bNoMore = true;
    if (n3.vKeys.size() == 1) n3.bNoMore = true;
    if (n4.vKeys.size() == 1) n4.bNoMore = true;
  }
};

struct SizeNode {  // pair<int, ExtractorNode*> with the pinned tie-break (creation sequence)
  int size;
  long seq;
  ExtractorNode* node;
  bool operator<(const SizeNode& o) const { return size != o.size ? size < o.size : seq < o.seq; }
};

class OrbOracle {
 public:
  int nfeatures, nlevels, iniThFAST, minThFAST;
  double scaleFactor;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
  std::vector<int> mnFeaturesPerLevel, umax;
  int blurk[7] = {18, 34, 48, 56, 48, 34, 18};
  std::vector<Level> pyr;
  std::vector<std::vector<KeyPoint>> candidates;  // per level, pre-distribution, border-local coords

  // ORBextractor::ORBextractor, include/ORBextractor.h:51-52, lib/libORB_SLAM2.so@0x73050
  OrbOracle(int nf, float sf, int nl, int ini, int mn)
      : nfeatures(nf), nlevels(nl), iniThFAST(ini), minThFAST(mn), scaleFactor((double)sf) {
    mvScaleFactor.resize(nl); mvLevelSigma2.resize(nl);
    mvInvScaleFactor.resize(nl); mvInvLevelSigma2.resize(nl);
    mvScaleFactor[0] = 1.f; mvLevelSigma2[0] = 1.f;
    for (int i = 1; i < nl; ++i) {
      mvScaleFactor[i] = (float)((double)mvScaleFactor[i - 1] * scaleFactor);
      mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
    }
    for (int i = 0; i < nl; ++i) {
      mvInvScaleFactor[i] = 1.f / mvScaleFactor[i];
      mvInvLevelSigma2[i] = 1.f / mvLevelSigma2[i];
    }
    mnFeaturesPerLevel.resize(nl);
    float factor = (float)(1.0 / scaleFactor);
    float nDesired = (float)nf * (1.f - factor) / (1.f - (float)std::pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; ++l) {
      mnFeaturesPerLevel[l] = cvRoundf(nDesired);
      sum += mnFeaturesPerLevel[l];
      nDesired *= factor;
    }
    mnFeaturesPerLevel[nl - 1] = std::max(nf - sum, 0);
    umax.assign(HALF_PATCH_SIZE + 1, 0);
    int v, v0;
    int vmax = (int)std::floor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) umax[v] = cvRoundd(std::sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
  }

  // ORBextractor::ComputePyramid, ORBextractor.h:89, @0x70430.  The 19-px reflected
  // border the reference materialises is never read downstream and is not modelled.
  void ComputePyramid(const uint8_t* img, int W, int H, int pitch) {
    pyr.assign(nlevels, Level());
    for (int l = 0; l < nlevels; ++l) {
      float inv = mvInvScaleFactor[l];
      Level& L = pyr[l];
      L.w = cvRoundf((float)W * inv);
      L.h = cvRoundf((float)H * inv);
      L.img.resize((size_t)L.w * L.h);
      if (l == 0) {
        for (int y = 0; y < H; ++y) std::memcpy(&L.img[(size_t)y * W], img + (size_t)y * pitch, W);
      } else {
        Level& P = pyr[l - 1];
        resize_linear_u8(P.img.data(), P.w, P.h, P.w, L.img.data(), L.w, L.h, L.w);
      }
    }
  }

  // ORBextractor::DistributeOctTree, ORBextractor.h:91-92, @0x73c60
  std::vector<KeyPoint> DistributeOctTree(const std::vector<KeyPoint>& keys, int minX, int maxX, int minY,
                                          int maxY, int N) {
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    const float hX = (float)(maxX - minX) / nIni;
    std::list<ExtractorNode> lNodes;
    std::vector<ExtractorNode*> vpIniNodes(nIni);
    long seq = 0;
    for (int i = 0; i < nIni; ++i) {
      ExtractorNode ni;
      ni.ULx = (int)(hX * (float)i); ni.ULy = 0;
      ni.URx = (int)(hX * (float)(i + 1)); ni.URy = 0;
      ni.BLx = ni.ULx; ni.BLy = maxY - minY;
      ni.BRx = ni.URx; ni.BRy = maxY - minY;
      ni.vKeys.reserve(keys.size());
      lNodes.push_back(ni);
      vpIniNodes[i] = &lNodes.back();
    }
    for (const KeyPoint& kp : keys) vpIniNodes[(int)(kp.x / hX)]->vKeys.push_back(kp);
    for (auto lit = lNodes.begin(); lit != lNodes.end();) {
      if (lit->vKeys.size() == 1) { lit->bNoMore = true; ++lit; }
      else if (lit->vKeys.empty()) lit = lNodes.erase(lit);
      else ++lit;
    }
    bool bFinish = false;
    std::vector<SizeNode> vSizeAndPointerToNode;
    vSizeAndPointerToNode.reserve(lNodes.size() * 4);
    auto add_children = [&](ExtractorNode* ch[4], int& nToExpand) {
      for (int c = 0; c < 4; ++c) {
        ExtractorNode& n = *ch[c];
        if (n.vKeys.size() > 0) {
          lNodes.push_front(n);
          if (n.vKeys.size() > 1) {
            ++nToExpand;
            vSizeAndPointerToNode.push_back({(int)n.vKeys.size(), seq++, &lNodes.front()});
            lNodes.front().lit = lNodes.begin();
          }
        }
      }
    };
    while (!bFinish) {
      int prevSize = (int)lNodes.size();
      auto lit = lNodes.begin();
      int nToExpand = 0;
      vSizeAndPointerToNode.clear();
      while (lit != lNodes.end()) {
        if (lit->bNoMore) { ++lit; continue; }
        ExtractorNode n1, n2, n3, n4;
        lit->DivideNode(n1, n2, n3, n4);
        ExtractorNode* ch[4] = {&n1, &n2, &n3, &n4};
        add_children(ch, nToExpand);
        lit = lNodes.erase(lit);
      }
      if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) {
        bFinish = true;
      } else if ((int)lNodes.size() + nToExpand * 3 > N) {
        while (!bFinish) {
          prevSize = (int)lNodes.size();
          std::vector<SizeNode> vPrev = vSizeAndPointerToNode;
          vSizeAndPointerToNode.clear();
          std::sort(vPrev.begin(), vPrev.end());
          for (int j = (int)vPrev.size() - 1; j >= 0; --j) {
            ExtractorNode n1, n2, n3, n4;
            vPrev[j].node->DivideNode(n1, n2, n3, n4);
            ExtractorNode* ch[4] = {&n1, &n2, &n3, &n4};
            int dummy = 0;
            add_children(ch, dummy);
            lNodes.erase(vPrev[j].node->lit);
            if ((int)lNodes.size() >= N) break;
          }
          if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) bFinish = true;
        }
      }
    }
    std::vector<KeyPoint> out;
    out.reserve(nfeatures);
    for (auto& n : lNodes) {
      const KeyPoint* best = &n.vKeys[0];
      float maxR = best->response;
      for (size_t k = 1; k < n.vKeys.size(); ++k)
        if (n.vKeys[k].response > maxR) { best = &n.vKeys[k]; maxR = best->response; }
      out.push_back(*best);
    }
    return out;
  }

  // IC_Angle, @0x6fb10; pt in level coordinates.
  float IC_Angle(const Level& L, float px, float py) const {
    int m01 = 0, m10 = 0;
    const int step = L.w;
    const uint8_t* center = &L.img[(size_t)cvRoundf(py) * step + cvRoundf(px)];
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
      int v_sum = 0, d = umax[v];
      for (int u = -d; u <= d; ++u) {
        int vp = center[u + v * step], vm = center[u - v * step];
        v_sum += vp - vm;
        m10 += u * (vp + vm);
      }
      m01 += v * v_sum;
    }
    return fast_atan2((float)m01, (float)m10);
  }

  // ComputeKeyPointsOctTree, ORBextractor.h:90, @0x75fa0
  void ComputeKeyPointsOctTree(std::vector<std::vector<KeyPoint>>& all) {
    all.assign(nlevels, {});
    candidates.assign(nlevels, {});
    const float W = 30;
    std::vector<FastPt> cell;
    for (int level = 0; level < nlevels; ++level) {
      const Level& L = pyr[level];
      const int minBorderX = MIN_BORDER, minBorderY = MIN_BORDER;
      const int maxBorderX = L.w - EDGE_THRESHOLD + 3, maxBorderY = L.h - EDGE_THRESHOLD + 3;
      std::vector<KeyPoint>& vToDistributeKeys = candidates[level];
      const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
      const int nCols = (int)(width / W), nRows = (int)(height / W);
      if (nCols <= 0 || nRows <= 0) continue;
      const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
      for (int i = 0; i < nRows; ++i) {
        const int iniY = minBorderY + i * hCell;
        int maxY = iniY + hCell + 6;
        if (iniY >= maxBorderY - 3) continue;
        if (maxY > maxBorderY) maxY = maxBorderY;
        for (int j = 0; j < nCols; ++j) {
          const int iniX = minBorderX + j * wCell;
          int maxX = iniX + wCell + 6;
          if (iniX >= maxBorderX - 6) continue;
          if (maxX > maxBorderX) maxX = maxBorderX;
          const uint8_t* sub = &L.img[(size_t)iniY * L.w + iniX];
          fast9_nms(sub, maxX - iniX, maxY - iniY, L.w, iniThFAST, cell);
          if (cell.empty()) fast9_nms(sub, maxX - iniX, maxY - iniY, L.w, minThFAST, cell);
          for (const FastPt& p : cell) {
            KeyPoint kp;
            kp.x = (float)(p.x + j * wCell); kp.y = (float)(p.y + i * hCell);
            kp.size = 7.f; kp.angle = -1.f; kp.response = (float)p.score; kp.octave = 0; kp.class_id = -1;
            vToDistributeKeys.push_back(kp);
          }
        }
      }
      std::vector<KeyPoint>& keypoints = all[level];
      if (!vToDistributeKeys.empty())
        keypoints = DistributeOctTree(vToDistributeKeys, minBorderX, maxBorderX, minBorderY, maxBorderY,
                                      mnFeaturesPerLevel[level]);
      const int scaledPatchSize = (int)((float)PATCH_SIZE * mvScaleFactor[level]);
      for (KeyPoint& kp : keypoints) {
        kp.x += minBorderX; kp.y += minBorderY;
        kp.octave = level; kp.size = (float)scaledPatchSize;
      }
    }
    for (int level = 0; level < nlevels; ++level)
      for (KeyPoint& kp : all[level]) kp.angle = IC_Angle(pyr[level], kp.x, kp.y);
  }

  // computeOrbDescriptor (inlined @0x777a8-0x77c72)
  void computeOrbDescriptor(const KeyPoint& kp, const Level& L, uint8_t* desc) const {
    float angle = kp.angle * 0.017453292f;  // factorPI @0x1269cc
    float a, b;
    orb_sincos(angle, &b, &a);  // b = sin, a = cos
    const int step = L.w;
    const uint8_t* center = &L.blurred[(size_t)cvRoundf(kp.y) * step + cvRoundf(kp.x)];
    auto GET = [&](int idx) -> int {
      float px = (float)kPattern[2 * idx], py = (float)kPattern[2 * idx + 1];
      int r = cvRoundf(std::fmaf(px, b, py * a));
      int c = cvRoundf(std::fmaf(px, a, -(py * b)));
      return center[r * step + c];
    };
    for (int i = 0; i < 32; ++i) {
      int val = 0;
      for (int k = 0; k < 8; ++k) {
        int t0 = GET(16 * i + 2 * k), t1 = GET(16 * i + 2 * k + 1);
        val |= (t0 < t1) << k;
      }
      desc[i] = (uint8_t)val;
    }
  }

  // ORBextractor::operator(), ORBextractor.h:59-61, @0x76da0
  int Extract(const uint8_t* img, int W, int H, int pitch, std::vector<KeyPoint>& kps, std::vector<uint8_t>& desc) {
    kps.clear(); desc.clear();
    if (!img || W <= 0 || H <= 0) return 0;
    ComputePyramid(img, W, H, pitch);
    std::vector<std::vector<KeyPoint>> all;
    ComputeKeyPointsOctTree(all);
    int n = 0;
    for (auto& v : all) n += (int)v.size();
    desc.assign((size_t)n * 32, 0);
    int offset = 0;
    for (int level = 0; level < nlevels; ++level) {
      std::vector<KeyPoint>& keypoints = all[level];
      Level& L = pyr[level];
      if (keypoints.empty()) { L.blurred.clear(); continue; }
      L.blurred.resize(L.img.size());
      blur7_u8(L.img.data(), L.w, L.h, L.w, L.blurred.data(), L.w, blurk);
      for (size_t i = 0; i < keypoints.size(); ++i)
        computeOrbDescriptor(keypoints[i], L, &desc[(size_t)(offset + i) * 32]);
      offset += (int)keypoints.size();
      if (level != 0) {
        float scale = mvScaleFactor[level];
        for (KeyPoint& kp : keypoints) { kp.x *= scale; kp.y *= scale; }
      }
      kps.insert(kps.end(), keypoints.begin(), keypoints.end());
    }
    return n;
  }
};

}  // namespace

// ---------------------------------------------------------------------------
// C entry points for ctypes (tests / bench cpu_baseline only)
// ---------------------------------------------------------------------------
extern "C" {

void oracle_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh, int dstep) {
  resize_linear_u8(src, sw, sh, sstep, dst, dw, dh, dstep);
}
void oracle_blur7_u8(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep, const int* k) {
  blur7_u8(src, w, h, sstep, dst, dstep, k);
}
// out: n x 3 ints (x, y, response)
int oracle_fast9(const uint8_t* img, int w, int h, int step, int th, int* out, int cap) {
  std::vector<FastPt> v;
  fast9_nms(img, w, h, step, th, v);
  int n = (int)std::min<size_t>(v.size(), cap);
  for (int i = 0; i < n; ++i) { out[3 * i] = v[i].x; out[3 * i + 1] = v[i].y; out[3 * i + 2] = v[i].score; }
  return (int)v.size();
}
float oracle_fast_atan2(float y, float x) { return fast_atan2(y, x); }
void oracle_sincos(float x, float* s, float* c) { orb_sincos(x, s, c); }
const int* oracle_orb_pattern() { return kPattern; }

void* oracle_orb_create(int nf, float sf, int nl, int ini, int mn) { return new OrbOracle(nf, sf, nl, ini, mn); }
void oracle_orb_destroy(void* h) { delete (OrbOracle*)h; }
void oracle_orb_set_blur_kernel(void* h, const int* k) { std::memcpy(((OrbOracle*)h)->blurk, k, 7 * sizeof(int)); }
void oracle_orb_tables(void* h, float* sf, float* isf, float* s2, float* is2, int* quota, int* umax16) {
  OrbOracle* o = (OrbOracle*)h;
  for (int i = 0; i < o->nlevels; ++i) {
    sf[i] = o->mvScaleFactor[i]; isf[i] = o->mvInvScaleFactor[i];
    s2[i] = o->mvLevelSigma2[i]; is2[i] = o->mvInvLevelSigma2[i];
    quota[i] = o->mnFeaturesPerLevel[i];
  }
  for (int i = 0; i < 16; ++i) umax16[i] = o->umax[i];
}
// kps: cap x 28 B, desc: cap x 32 B.  Returns the keypoint count (may exceed cap; then nothing is copied).
int oracle_orb_extract(void* h, const uint8_t* img, int W, int H, int pitch, void* kps, uint8_t* desc, int cap) {
  OrbOracle* o = (OrbOracle*)h;
  std::vector<KeyPoint> k;
  std::vector<uint8_t> d;
  int n = o->Extract(img, W, H, pitch, k, d);
  if (n <= cap) {
    if (n) { std::memcpy(kps, k.data(), (size_t)n * sizeof(KeyPoint)); std::memcpy(desc, d.data(), (size_t)n * 32); }
  }
  return n;
}
int oracle_orb_level_size(void* h, int level, int* w, int* hh) {
  OrbOracle* o = (OrbOracle*)h;
  if (level < 0 || level >= (int)o->pyr.size()) return -1;
  *w = o->pyr[level].w; *hh = o->pyr[level].h;
  return 0;
}
// which: 0 = pyramid level, 1 = blurred level (empty if the level had no keypoints). Dense rows of width w.
int oracle_orb_level_copy(void* h, int level, int which, uint8_t* out) {
  OrbOracle* o = (OrbOracle*)h;
  const std::vector<uint8_t>& v = which ? o->pyr[level].blurred : o->pyr[level].img;
  if (v.empty()) return 0;
  std::memcpy(out, v.data(), v.size());
  return (int)v.size();
}
// pre-distribution candidate list of a level (border-local coords): cap x 3 ints (x, y, response)
int oracle_orb_candidates(void* h, int level, int* out, int cap) {
  OrbOracle* o = (OrbOracle*)h;
  const auto& v = o->candidates[level];
  int n = (int)std::min<size_t>(v.size(), cap);
  for (int i = 0; i < n; ++i) { out[3 * i] = (int)v[i].x; out[3 * i + 1] = (int)v[i].y; out[3 * i + 2] = (int)v[i].response; }
  return (int)v.size();
}
// Stand-alone DistributeOctTree on an (x, y, response) list; out = indices into the input in output order.
int oracle_orb_distribute(void* h, const int* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int* out_idx, int cap) {
  OrbOracle* o = (OrbOracle*)h;
  std::vector<KeyPoint> keys(n);
  for (int i = 0; i < n; ++i) {
    keys[i].x = (float)xyr[3 * i]; keys[i].y = (float)xyr[3 * i + 1]; keys[i].response = (float)xyr[3 * i + 2];
    keys[i].class_id = i; keys[i].octave = 0; keys[i].size = 7; keys[i].angle = -1;
  }
  std::vector<KeyPoint> r = o->DistributeOctTree(keys, minX, maxX, minY, maxY, N);
  int m = (int)std::min<size_t>(r.size(), cap);
  for (int i = 0; i < m; ++i) out_idx[i] = r[i].class_id;
  return (int)r.size();
}

}  // extern "C"
