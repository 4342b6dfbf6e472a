// oracle/match_oracle.cc — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// Array-form restatement of the Hamming matching cores of the reference:
//   ORBmatcher::DescriptorDistance            include/ORBmatcher.h:44, lib/libORB_SLAM2.so@0x79d20;
//                                             same source as Thirdparty/DBoW2/DBoW2/FORB.cpp:82-102 (golden)
//   ORBmatcher::SearchByBoW(KeyFrame*,Frame&) include/ORBmatcher.h:104, @0x80150
//   ORBmatcher::SearchByProjection(Frame&,const Frame&,th,bMono)   include/ORBmatcher.h:78, @0x80d00
//   Frame::GetFeaturesInArea                  include/Frame.h:113
//   ORBmatcher::ComputeThreeMaxima            @0x79c40
//   LSDmatcher / LineSegment::LineSegmentMathch: cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) semantics
//                                             (header evidence include/auxiliar.h:30-51; UNPINNED beyond that)
// Parity status: DescriptorDistance, ComputeThreeMaxima, RadiusByViewingCos, CheckDistEpipolarLine, Frame::GetFeaturesInArea and
// the whole functions SearchByProjection (both overloads), SearchByBoW, SearchForTriangulation and SearchForInitialization are
// PINNED AGAINST THE REFERENCE'S OWN CODE, executed from lib/libORB_SLAM2.so on hand-laid-out Frame / KeyFrame / MapPoint
// objects (tests/golden/reference_code.py; fixtures reference_code.npz / reference_library.npz; tests/test_golden_cpu.py).
// Constants from the binary: TH_LOW=50, TH_HIGH=100, HISTO_LENGTH=30 (@0x1269e0-e8), histogram factor
// HISTO_LENGTH/360 (@0x1269f8), secondary-bin cut 0.1 (@0x1269f0).
// The pointer-based containers (MapPoint*, std::map FeatureVector, mGrid vectors) are flattened to
// arrays/CSR; iteration orders are preserved.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

constexpr int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;
constexpr int FRAME_GRID_ROWS = 48, FRAME_GRID_COLS = 64;  // include/Frame.h:41-42

// FORB::distance / ORBmatcher::DescriptorDistance: SWAR popcount over 8 x uint32
int descriptor_distance(const uint8_t* a, const uint8_t* b) {
  const int32_t* pa = (const int32_t*)a;
  const int32_t* pb = (const int32_t*)b;
  int dist = 0;
  for (int i = 0; i < 8; i++, pa++, pb++) {
    unsigned int v = *pa ^ *pb;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

void compute_three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// ORBmatcher::RadiusByViewingCos (@0x79b60): the float is widened and compared with the DOUBLE 0.998 (@0x126a00), so
// 0.998f itself (= 0.99800002...) already gets the small radius.  Pinned by executing the reference's machine code
// (tests/golden/reference_code.py).
inline float radius_by_viewing_cos(float viewCos) { return (double)viewCos > 0.998 ? 2.5f : 4.0f; }

// Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel) (include/Frame.h:113; machine code @0xfbc60): cell range by
// floorf / ceilf of (x - mnMinX -/+ r) * mfGridElementWidthInv clamped to the 64 x 48 grid (@0xfbcc0-0xfbdd4), cells walked
// [ix][iy], items in insertion order; level filter when minLevel > 0 or maxLevel >= 0 (@0xfbdda-0xfbdf9, @0xfc05d-0xfc06d);
// a feature is a candidate when |dx| < r and |dy| < r (@0xfbf87-0xfbfb9).  Calls f(index) for each candidate in the
// reference's order.  Pinned by running the reference's own function (tests/golden/reference_code.py).
template <class F>
inline void for_features_in_area(float x, float y, float r, int minLevel, int maxLevel, float mnMinX, float mnMinY, float gwi,
                                 float ghi, const int* gridStart, const int* gridItems, const float* xy, const int* octave, F&& f) {
  const int nMinCellX = std::max(0, (int)floorf((x - mnMinX - r) * gwi));
  if (nMinCellX >= FRAME_GRID_COLS) return;
  const int nMaxCellX = std::min((int)FRAME_GRID_COLS - 1, (int)ceilf((x - mnMinX + r) * gwi));
  if (nMaxCellX < 0) return;
  const int nMinCellY = std::max(0, (int)floorf((y - mnMinY - r) * ghi));
  if (nMinCellY >= FRAME_GRID_ROWS) return;
  const int nMaxCellY = std::min((int)FRAME_GRID_ROWS - 1, (int)ceilf((y - mnMinY + r) * ghi));
  if (nMaxCellY < 0) return;
  const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
    for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
      const int c = ix * FRAME_GRID_ROWS + iy;
      for (int j = gridStart[c]; j < gridStart[c + 1]; ++j) {
        const int i2 = gridItems[j];
        if (bCheckLevels) {
          if (octave[i2] < minLevel) continue;
          if (maxLevel >= 0 && octave[i2] > maxLevel) continue;
        }
        const float distx = xy[2 * i2] - x, disty = xy[2 * i2 + 1] - y;
        if (!(fabsf(distx) < r && fabsf(disty) < r)) continue;
        f(i2);
      }
    }
}

inline int rot_bin(float a1, float a2) {
  const float factor = (float)HISTO_LENGTH / 360.0f;  // 0.0833333 @0x1269f8
  float rot = a1 - a2;
  if (rot < 0.0) rot += 360.0f;
  int bin = (int)roundf(rot * factor);
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}

// logf as glibc >= 2.27 computes it (sysdeps/ieee754/flt-32/e_logf.c: 16-entry table of 1/c and log(c), degree-3 polynomial in
// double, one rounding to float).  MapPoint::PredictScale (@0x8fb60 / @0x8fc20) calls logf from the C library the binary is
// loaded with; the restatement is pinned against this container's glibc 2.39 bit for bit (tests/test_oracle_match_cpu.py,
// table read back from libm.so.6).  Special cases (zero, negative, inf, nan, subnormal) follow the same source.
static const double LOGF_TAB[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
inline float pl_logf(float x) {
  uint32_t ix;
  std::memcpy(&ix, &x, 4);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2 == 0) return -INFINITY;
    if (ix == 0x7f800000u) return x;
    if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return NAN;
    const float xs = x * 0x1p23f;  // subnormal: normalise
    std::memcpy(&ix, &xs, 4);
    ix -= 23u << 23;
  }
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) % 16);
  const int k = (int32_t)tmp >> 23;
  const uint32_t iz = ix - (tmp & (0x1ffu << 23));
  float zf;
  std::memcpy(&zf, &iz, 4);
  const double z = (double)zf;
  const double r = z * LOGF_TAB[i][0] - 1;
  const double y0 = LOGF_TAB[i][1] + (double)k * 0x1.62e42fefa39efp-1;
  const double r2 = r * r;
  double y = 0x1.5575b0be00b6ap-2 * r + -0x1.ffffef20a4123p-2;
  y = -0x1.00ea348b88334p-2 * r2 + y;
  y = y * r2 + (y0 + r);
  return (float)y;
}

// MapPoint::PredictScale(const float& currentDist, Frame* / KeyFrame*) (@0x8fc20 / @0x8fb60): ratio = mfMaxDistance / dist
// (vdivss), ceilf(logf(ratio) / mfLogScaleFactor) truncated to int, clamped to [0, mnScaleLevels - 1].
inline int predict_scale(float maxDistance, float dist, float logScaleFactor, int nLevels) {
  const float ratio = maxDistance / dist;
  int nScale = (int)ceilf(pl_logf(ratio) / logScaleFactor);
  if (nScale < 0) nScale = 0;
  else if (nScale >= nLevels) nScale = nLevels - 1;
  return nScale;
}

}  // namespace

extern "C" {

int oracle_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }
float oracle_radius_by_viewing_cos(float v) { return radius_by_viewing_cos(v); }
// Frame::GetFeaturesInArea on a CSR grid; cam4 = mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv.  Returns the
// number of candidates, the first `cap` of them in out.
int oracle_get_features_in_area(float x, float y, float r, int minLevel, int maxLevel, const float* cam4, const int* gridStart,
                                const int* gridItems, const float* xy, const int* octave, int* out, int cap) {
  int n = 0;
  for_features_in_area(x, y, r, minLevel, maxLevel, cam4[0], cam4[1], cam4[2], cam4[3], gridStart, gridItems, xy, octave, [&](int i) {
    if (n < cap) out[n] = i;
    ++n;
  });
  return n;
}
// ComputeThreeMaxima on bin populations; out3 must hold the caller's initial values (-1 in every reference call site)
void oracle_three_maxima(const int* sizes, int L, int* out3) {
  std::vector<std::vector<int>> h(L);
  for (int i = 0; i < L; ++i) h[i].resize(sizes[i]);
  compute_three_maxima(h.data(), L, out3[0], out3[1], out3[2]);
}

// BFMatcher(NORM_HAMMING).knnMatch(query, train, k=2): per query the two nearest train rows, ties by lower
// train index.  out: nq x 4 ints (idx1, dist1, idx2, dist2); missing entries are -1.
void oracle_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int* out) {
  for (int i = 0; i < nq; ++i) {
    int b1 = -1, d1 = 1 << 30, b2 = -1, d2 = 1 << 30;
    for (int j = 0; j < nt; ++j) {
      const int d = descriptor_distance(q + 32 * i, t + 32 * j);
      if (d < d1) { b2 = b1; d2 = d1; b1 = j; d1 = d; }
      else if (d < d2) { b2 = j; d2 = d; }
    }
    out[4 * i] = b1; out[4 * i + 1] = b1 < 0 ? -1 : d1; out[4 * i + 2] = b2; out[4 * i + 3] = b2 < 0 ? -1 : d2;
  }
}

// SearchByBoW(KeyFrame*, Frame&, matches).  FeatureVectors as CSR sorted by node id:
//   nodes[k], start[k]..start[k+1] into idx[].  kfValid[i] = (pMP && !pMP->isBad()).
// matchF[N2] receives the KF feature index matched to each F feature (-1 = none).  Returns nmatches.
int oracle_search_by_bow(const uint8_t* dKF, const float* angKF, const uint8_t* kfValid, int nNodesKF, const int* nodesKF,
                         const int* startKF, const int* idxKF, const uint8_t* dF, const float* angF, int N2, int nNodesF,
                         const int* nodesF, const int* startF, const int* idxF, float nnratio, int checkOri, int* matchF) {
  for (int i = 0; i < N2; ++i) matchF[i] = -1;
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  int a = 0, b = 0;
  while (a < nNodesKF && b < nNodesF) {
    if (nodesKF[a] == nodesF[b]) {
      for (int iKF = startKF[a]; iKF < startKF[a + 1]; ++iKF) {
        const int realIdxKF = idxKF[iKF];
        if (!kfValid[realIdxKF]) continue;
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (int iF = startF[b]; iF < startF[b + 1]; ++iF) {
          const int realIdxF = idxF[iF];
          if (matchF[realIdxF] >= 0) continue;
          const int dist = descriptor_distance(dKF + 32 * realIdxKF, dF + 32 * realIdxF);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
          else if (dist < bestDist2) { bestDist2 = dist; }
        }
        if (bestDist1 <= TH_LOW) {
          if ((float)bestDist1 < nnratio * (float)bestDist2) {
            matchF[bestIdxF] = realIdxKF;
            if (checkOri) rotHist[rot_bin(angKF[realIdxKF], angF[bestIdxF])].push_back(bestIdxF);
            nmatches++;
          }
        }
      }
      ++a; ++b;
    } else if (nodesKF[a] < nodesF[b]) {
      a = (int)(std::lower_bound(nodesKF, nodesKF + nNodesKF, nodesF[b]) - nodesKF);
    } else {
      b = (int)(std::lower_bound(nodesF, nodesF + nNodesF, nodesKF[a]) - nodesF);
    }
  }
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    compute_three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int j : rotHist[i]) { matchF[j] = -1; nmatches--; }
    }
  }
  return nmatches;
}

// SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) (ORBmatcher.h:105; @0x82cc0, loop closing).
// Same walk over the shared vocabulary nodes as the key-frame / frame form above, with three differences read from the
// binary: KF2 features need a good map point as well (valid2) and are skipped once matched (vbMatched2), the distance
// test is bestDist1 < TH_LOW (cmpl $0x31 @0x83490, against $0x32 @0x808e0 in the other form), and the result is indexed
// by KF1: match12[i1] = KF2 feature whose map point is assigned to KF1 feature i1 (-1 = none).  Returns nmatches.
int oracle_search_by_bow_kfkf(const uint8_t* d1, const float* ang1, const uint8_t* valid1, int N1, int nNodes1, const int* nodes1,
                              const int* start1, const int* idx1, const uint8_t* d2, const float* ang2, const uint8_t* valid2,
                              int N2, int nNodes2, const int* nodes2, const int* start2, const int* idx2, float nnratio,
                              int checkOri, int* match12) {
  for (int i = 0; i < N1; ++i) match12[i] = -1;
  std::vector<uint8_t> matched2(N2, 0);
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  int a = 0, b = 0;
  while (a < nNodes1 && b < nNodes2) {
    if (nodes1[a] == nodes2[b]) {
      for (int i1 = start1[a]; i1 < start1[a + 1]; ++i1) {
        const int r1 = idx1[i1];
        if (!valid1[r1]) continue;
        int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
        for (int i2 = start2[b]; i2 < start2[b + 1]; ++i2) {
          const int r2 = idx2[i2];
          if (matched2[r2] || !valid2[r2]) continue;
          const int dist = descriptor_distance(d1 + 32 * r1, d2 + 32 * r2);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = r2; }
          else if (dist < bestDist2) { bestDist2 = dist; }
        }
        if (bestDist1 < TH_LOW) {
          if ((float)bestDist1 < nnratio * (float)bestDist2) {
            match12[r1] = bestIdx2;
            matched2[bestIdx2] = 1;
            if (checkOri) rotHist[rot_bin(ang1[r1], ang2[bestIdx2])].push_back(r1);
            nmatches++;
          }
        }
      }
      ++a; ++b;
    } else if (nodes1[a] < nodes2[b]) {
      a = (int)(std::lower_bound(nodes1, nodes1 + nNodes1, nodes2[b]) - nodes1);
    } else {
      b = (int)(std::lower_bound(nodes2, nodes2 + nNodes2, nodes1[a]) - nodes2);
    }
  }
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    compute_three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int j : rotHist[i]) { match12[j] = -1; nmatches--; }
    }
  }
  return nmatches;
}

// SearchByProjection(CurrentFrame, LastFrame, th, bMono).
// Last frame: N1 points with lastValid[i] = (pMP && !outlier), world position (x,y,z), representative
// descriptor, mvKeys[i].octave, mvKeysUn[i].angle, lastObs[i] = (pMP->Observations() > 0).
// Current frame: N2 undistorted keypoints (x, y, octave, angle), descriptors, mvuRight, curTaken[i] =
// (mvpMapPoints[i] && Observations() > 0) on entry, the 64x48 grid as CSR in [ix][iy] order, bounds,
// intrinsics and pose.  cam = {fx, fy, cx, cy, mbf, mb, mnMinX, mnMaxX, mnMinY, mnMaxY, gridWInv, gridHInv}.
// matchCur[N2] receives the last-frame index assigned to each current keypoint (-1 = none). Returns nmatches.
int oracle_search_by_projection(int N1, const uint8_t* lastValid, const float* lastXYZ, const uint8_t* lastDesc,
                                const int* lastOctave, const float* lastAngle, const uint8_t* lastObs, int N2,
                                const float* curXY, const int* curOctave, const float* curAngle, const uint8_t* curDesc,
                                const float* curURight, const uint8_t* curTaken, const int* gridStart,
                                const int* gridItems, const float* cam, const float* scaleFactors, const float* TcwCur,
                                const float* TcwLast, float th, int bMono, int checkOri, int* matchCur) {
  const float fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3], mbf = cam[4], mb = cam[5];
  const float mnMinX = cam[6], mnMaxX = cam[7], mnMinY = cam[8], mnMaxY = cam[9], gwi = cam[10], ghi = cam[11];
  std::vector<uint8_t> taken(curTaken, curTaken + N2);
  for (int i = 0; i < N2; ++i) matchCur[i] = -1;
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  // Rcw, tcw (row-major 3x4); twc = -Rcw^T tcw; tlc = Rlw twc + tlw.  cv::Mat products of CV_32F matrices go
  // through cv::gemm.
  const float* Rc = TcwCur;  // Rc[r*4+c], t = Rc[r*4+3]
  float twc[3];
  for (int r = 0; r < 3; ++r) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += (double)Rc[k * 4 + r] * (double)Rc[k * 4 + 3];
    twc[r] = (float)(-1.0 * s);
  }
  // Rlw * twc + tlw and Rcw * x3Dw + tcw carry no transpose flag and have an inner dimension of 3: cv::gemm's small-matrix
  // path (float products summed left to right in float, then (double)sum + (double)c rounded to float), as pinned for the
  // epipole below; -Rcw.t() * tcw above carries a transpose flag: GEMMSingleMul, double accumulator.
  float tlc2;
  {
    const float p0 = TcwLast[8] * twc[0], p1 = TcwLast[9] * twc[1], p2 = TcwLast[10] * twc[2];
    const float s = (p0 + p1) + p2;
    tlc2 = (float)((double)s + (double)TcwLast[2 * 4 + 3]);
  }
  const bool bForward = tlc2 > mb && !bMono;
  const bool bBackward = -tlc2 > mb && !bMono;
  for (int i = 0; i < N1; i++) {
    if (!lastValid[i]) continue;
    const float* X = lastXYZ + 3 * i;
    float pc[3];
    for (int r = 0; r < 3; ++r) {
      const float p0 = Rc[r * 4] * X[0], p1 = Rc[r * 4 + 1] * X[1], p2 = Rc[r * 4 + 2] * X[2];
      const float s = (p0 + p1) + p2;
      pc[r] = (float)((double)s + (double)Rc[r * 4 + 3]);
    }
    const float xc = pc[0], yc = pc[1];
    // the binary's sequence (it was compiled with FMA): vdivss @0x81c92, vmulss + vfmadd213ss @0x81caf-0x81cba and
    // @0x81cce-0x81cd9, vfnmadd132ss @0x81eb5 for ur below
    const float invzc = 1.0f / pc[2];
    if (invzc < 0) continue;
    float u = std::fmaf(xc * fx, invzc, cx);
    float v = std::fmaf(yc * fy, invzc, cy);
    if (u < mnMinX || u > mnMaxX) continue;
    if (v < mnMinY || v > mnMaxY) continue;
    const int nLastOctave = lastOctave[i];
    const float radius = th * scaleFactors[nLastOctave];
    int minLevel, maxLevel;
    if (bForward) { minLevel = nLastOctave; maxLevel = -1; }
    else if (bBackward) { minLevel = 0; maxLevel = nLastOctave; }
    else { minLevel = nLastOctave - 1; maxLevel = nLastOctave + 1; }
    int bestDist = 256, bestIdx2 = -1;
    for_features_in_area(u, v, radius, minLevel, maxLevel, mnMinX, mnMinY, gwi, ghi, gridStart, gridItems, curXY, curOctave, [&](int i2) {
      if (taken[i2]) return;
      if (curURight[i2] > 0) {
        const float ur = std::fmaf(-invzc, mbf, u);
        const float er = fabsf(ur - curURight[i2]);
        if (er > radius) return;
      }
      const int dist = descriptor_distance(lastDesc + 32 * i, curDesc + 32 * i2);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    });
    if (bestDist <= TH_HIGH) {
      matchCur[bestIdx2] = i;
      if (lastObs[i]) taken[bestIdx2] = 1;
      nmatches++;
      if (checkOri) rotHist[rot_bin(lastAngle[i], curAngle[bestIdx2])].push_back(bestIdx2);
    }
  }
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    compute_three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i != ind1 && i != ind2 && i != ind3)
        for (int j : rotHist[i]) { matchCur[j] = -1; nmatches--; }
    }
  }
  return nmatches;
}

float oracle_logf(float x) { return pl_logf(x); }
// number of floats in [first, first + count * stride) (bit patterns) on which the restated logf differs from the C library's
int oracle_logf_mismatches(uint32_t first, uint32_t count, uint32_t stride) {
  int bad = 0;
  for (uint32_t k = 0; k < count; ++k) {
    const uint32_t b = first + k * stride;
    float x;
    std::memcpy(&x, &b, 4);
    const float a = pl_logf(x), c = ::logf(x);
    if (std::memcmp(&a, &c, 4) != 0 && !(a != a && c != c)) ++bad;
  }
  return bad;
}
int oracle_predict_scale(float maxDistance, float dist, float logScaleFactor, int nLevels) {
  return predict_scale(maxDistance, dist, logScaleFactor, nLevels);
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, float th, int ORBdist)
// (ORBmatcher.h:82; @0x7e8c0, Tracking::Relocalization).  Read from the binary: Ow = -Rcw.t() * tcw (@0x7ebec-0x7ec2b);
// per map point of the key frame that exists, is not bad and is not in sAlreadyFound (@0x7ef24-0x7ef84): x3Dc = Rcw * x3Dw +
// tcw (gemm, @0x7efaa-0x7efc0), invzc = 1.0f / zc (vdivss @0x7f4cd, no sign test), u = fma(xc * fx, invzc, cx), v likewise
// (@0x7f4d1-0x7f4fb), image-bounds test (@0x7f513-0x7f542); PO = x3Dw - Ow, dist3D = (float)cv::norm(PO) (@0x7f96e-0x7fa81:
// squares summed in double in element order, sqrt), rejected when 0.8f * mfMinDistance > dist3D or dist3D > 1.2f *
// mfMaxDistance (@0x7faaa-0x7fab8); level = PredictScale(dist3D, &CurrentFrame) (@0x7fc3f); radius = th * mvScaleFactors[level]
// (@0x7fc6f); GetFeaturesInArea(u, v, radius, level - 1, level + 1) (@0x7fc8e); a candidate is skipped when the frame's feature
// already has a map point (@0x7fd5c); best = strictly smaller distance (@0x7fdba); accepted when bestDist <= ORBdist
// (@0x7fdf2); rotation histogram of pKF->mvKeysUn[i].angle - CurrentFrame.mvKeysUn[best].angle (@0x7fe65-0x7fe98).
// mpValid[i] = pMP && !pMP->isBad() && !sAlreadyFound.count(pMP); mpDistRange = (mfMinDistance, mfMaxDistance) per point;
// curTaken[i2] = CurrentFrame.mvpMapPoints[i2] != NULL on entry; cam = {fx, fy, cx, cy, mnMinX, mnMaxX, mnMinY, mnMaxY, gwi, ghi}.
// matchCur[N2] receives the key-frame index assigned to each current keypoint (-1 = none).  Returns nmatches.
int oracle_search_by_projection_kf(int M, const uint8_t* mpValid, const float* mpXYZ, const uint8_t* mpDesc,
                                   const float* mpDistRange, const float* kfAngle, int N2, const float* curXY,
                                   const int* curOctave, const float* curAngle, const uint8_t* curDesc, const uint8_t* curTaken,
                                   const int* gridStart, const int* gridItems, const float* cam, const float* scaleFactors,
                                   int nLevels, float logScaleFactor, const float* TcwCur, float th, int ORBdist, int checkOri,
                                   int* matchCur) {
  const float fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
  const float mnMinX = cam[4], mnMaxX = cam[5], mnMinY = cam[6], mnMaxY = cam[7], gwi = cam[8], ghi = cam[9];
  std::vector<uint8_t> taken(curTaken, curTaken + N2);
  for (int i = 0; i < N2; ++i) matchCur[i] = -1;
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float* Rc = TcwCur;
  float Ow[3];
  for (int r = 0; r < 3; ++r) {  // transpose flag: GEMMSingleMul, double accumulator
    double s = 0;
    for (int k = 0; k < 3; ++k) s += (double)Rc[k * 4 + r] * (double)Rc[k * 4 + 3];
    Ow[r] = (float)(-1.0 * s);
  }
  for (int i = 0; i < M; i++) {
    if (!mpValid[i]) continue;
    const float* X = mpXYZ + 3 * i;
    float pc[3];
    for (int r = 0; r < 3; ++r) {
      const float p0 = Rc[r * 4] * X[0], p1 = Rc[r * 4 + 1] * X[1], p2 = Rc[r * 4 + 2] * X[2];
      const float s = (p0 + p1) + p2;
      pc[r] = (float)((double)s + (double)Rc[r * 4 + 3]);
    }
    const float invzc = 1.0f / pc[2];
    const float u = std::fmaf(pc[0] * fx, invzc, cx);
    const float v = std::fmaf(pc[1] * fy, invzc, cy);
    if (u < mnMinX || u > mnMaxX) continue;
    if (v < mnMinY || v > mnMaxY) continue;
    double n2 = 0;
    for (int r = 0; r < 3; ++r) {
      const float d = X[r] - Ow[r];
      n2 += (double)d * (double)d;
    }
    const float dist3D = (float)std::sqrt(n2);
    const float maxDistance = 1.2f * mpDistRange[2 * i + 1], minDistance = 0.8f * mpDistRange[2 * i];
    if (minDistance > dist3D || dist3D > maxDistance) continue;
    const int level = predict_scale(mpDistRange[2 * i + 1], dist3D, logScaleFactor, nLevels);
    const float radius = th * scaleFactors[level];
    int bestDist = 256, bestIdx2 = -1;
    for_features_in_area(u, v, radius, level - 1, level + 1, mnMinX, mnMinY, gwi, ghi, gridStart, gridItems, curXY, curOctave, [&](int i2) {
      if (taken[i2]) return;
      const int dist = descriptor_distance(mpDesc + 32 * i, curDesc + 32 * i2);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    });
    if (bestIdx2 >= 0 && bestDist <= ORBdist) {
      matchCur[bestIdx2] = i;
      taken[bestIdx2] = 1;
      nmatches++;
      if (checkOri) rotHist[rot_bin(kfAngle[i], curAngle[bestIdx2])].push_back(bestIdx2);
    }
  }
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    compute_three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i != ind1 && i != ind2 && i != ind3)
        for (int j : rotHist[i]) { matchCur[j] = -1; nmatches--; }
    }
  }
  return nmatches;
}

}  // extern "C"
namespace {
// KeyFrame::GetFeaturesInArea(x, y, r) (@0x96fe0): the key frame's grid (mnGridCols x mnGridRows, CSR in [ix][iy] order), integer
// bounds converted to float, no level filter; a feature is a candidate when |dx| < r and |dy| < r.
template <class F>
inline void for_kf_features_in_area(float x, float y, float r, int mnMinX, int mnMinY, float gwi, float ghi, int cols, int rows,
                                    const int* gridStart, const int* gridItems, const float* xy, F&& f) {
  const int nMinCellX = std::max(0, (int)floorf((x - (float)mnMinX - r) * gwi));
  if (nMinCellX >= cols) return;
  const int nMaxCellX = std::min(cols - 1, (int)ceilf((x - (float)mnMinX + r) * gwi));
  if (nMaxCellX < 0) return;
  const int nMinCellY = std::max(0, (int)floorf((y - (float)mnMinY - r) * ghi));
  if (nMinCellY >= rows) return;
  const int nMaxCellY = std::min(rows - 1, (int)ceilf((y - (float)mnMinY + r) * ghi));
  if (nMaxCellY < 0) return;
  for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
    for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
      const int c = ix * rows + iy;
      for (int j = gridStart[c]; j < gridStart[c + 1]; ++j) {
        const int i2 = gridItems[j];
        const float distx = xy[2 * i2] - x, disty = xy[2 * i2 + 1] - y;
        if (fabsf(distx) < r && fabsf(disty) < r) f(i2);
      }
    }
}

}  // namespace
extern "C" {

// ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th)
// (ORBmatcher.h:86; @0x880f0, LoopClosing).  Read from the binary: scw = (float)sqrt(sRcw.row(0).dot(sRcw.row(0))) (dot in double,
// vsqrtsd @0x8836b); Rcw = sRcw / scw and tcw = Scw.col(3) / scw are cv::operator/(Mat, double) (@0x884fb, @0x886ab), which OpenCV
// evaluates as a scaled conversion: every element TIMES (float)(1.0 / (double)scw); Ow = -Rcw.t() * tcw; per map point that is
// not bad and not already in vpMatched: p3Dc = Rcw * p3Dw + tcw (gemm), z < 0 rejects (@0x890d9), invz = 1.0f / z, x = X * invz,
// u = fma(x, fx, cx), v likewise (@0x890f0-0x89160), KeyFrame::IsInImage (u >= mnMinX && u < mnMaxX ..., int bounds), the
// scale-invariance range, PO.dot(Pn) < 0.5 * dist rejects (doubles, @0x89acc), level = PredictScale(dist, pKF), radius =
// (float)th * mvScaleFactors[level] (@0x89b22), KeyFrame::GetFeaturesInArea(u, v, radius); a candidate is skipped when already
// matched or when its octave is outside [level - 1, level] (@0x89c1c-0x89c54); best = strictly smaller distance; accepted when
// bestDist <= TH_LOW (@0x89daf).  mpValid[i] = !isBad() && pMP not in vpMatched on entry; kfMatched[i] = vpMatched[i] != NULL on entry.
// matchKF[N] receives the map-point index newly assigned to each key-frame feature (-1 = none).  Returns nmatches.
int oracle_search_by_projection_sim3(int M, const uint8_t* mpValid, const float* mpXYZ, const float* mpNormal,
                                     const float* mpDistRange, const uint8_t* mpDesc, int N, const float* kfXY, const int* kfOctave,
                                     const uint8_t* kfDesc, const uint8_t* kfMatched, const int* gridStart, const int* gridItems,
                                     int gridCols, int gridRows, const float* Scw, const float* cam4, const int* bounds4, float gwi,
                                     float ghi, const float* scaleFactors, int nLevels, float logScaleFactor, int th, int* matchKF) {
  const float fx = cam4[0], fy = cam4[1], cx = cam4[2], cy = cam4[3];
  const int mnMinX = bounds4[0], mnMinY = bounds4[1], mnMaxX = bounds4[2], mnMaxY = bounds4[3];
  std::vector<uint8_t> matched(kfMatched, kfMatched + N);
  for (int i = 0; i < N; ++i) matchKF[i] = -1;
  double dot0 = 0;
  for (int k = 0; k < 3; ++k) dot0 += (double)Scw[k] * (double)Scw[k];
  const float scw = (float)std::sqrt(dot0);
  const float inv = (float)(1.0 / (double)scw);
  float T[12];  // Rcw | tcw
  for (int k = 0; k < 12; ++k) T[k] = Scw[k] * inv + 0.0f;
  float Ow[3];
  for (int r = 0; r < 3; ++r) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += (double)T[k * 4 + r] * (double)T[k * 4 + 3];
    Ow[r] = (float)(-1.0 * s);
  }
  int nmatches = 0;
  for (int i = 0; i < M; ++i) {
    if (!mpValid[i]) continue;
    const float* X = mpXYZ + 3 * i;
    float pc[3];
    for (int r = 0; r < 3; ++r) {
      const float p0 = T[r * 4] * X[0], p1 = T[r * 4 + 1] * X[1], p2 = T[r * 4 + 2] * X[2];
      const float s = (p0 + p1) + p2;
      pc[r] = (float)((double)s + (double)T[r * 4 + 3]);
    }
    if (pc[2] < 0.0f) continue;
    const float invz = 1.0f / pc[2];
    const float x = pc[0] * invz, y = pc[1] * invz;
    const float u = std::fmaf(x, fx, cx), v = std::fmaf(y, fy, cy);
    if (!(u >= (float)mnMinX && u < (float)mnMaxX && v >= (float)mnMinY && v < (float)mnMaxY)) continue;
    float PO[3];
    double n2 = 0;
    for (int r = 0; r < 3; ++r) {
      PO[r] = X[r] - Ow[r];
      n2 += (double)PO[r] * (double)PO[r];
    }
    const float dist = (float)std::sqrt(n2);
    if (dist < 0.8f * mpDistRange[2 * i] || dist > 1.2f * mpDistRange[2 * i + 1]) continue;
    double dot = 0;
    for (int r = 0; r < 3; ++r) dot += (double)PO[r] * (double)mpNormal[3 * i + r];
    if (dot < 0.5 * (double)dist) continue;
    const int level = predict_scale(mpDistRange[2 * i + 1], dist, logScaleFactor, nLevels);
    const float radius = (float)th * scaleFactors[level];
    int bestDist = 256, bestIdx = -1;
    for_kf_features_in_area(u, v, radius, mnMinX, mnMinY, gwi, ghi, gridCols, gridRows, gridStart, gridItems, kfXY, [&](int idx) {
      if (matched[idx]) return;
      const int kpLevel = kfOctave[idx];
      if (kpLevel < level - 1 || kpLevel > level) return;
      const int d = descriptor_distance(mpDesc + 32 * i, kfDesc + 32 * idx);
      if (d < bestDist) { bestDist = d; bestIdx = idx; }
    });
    if (bestDist <= TH_LOW) {
      matchKF[bestIdx] = i;
      matched[bestIdx] = 1;
      nmatches++;
    }
  }
  return nmatches;
}

// The matching core of ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th) (ORBmatcher.h:119;
// @0x7a500, LocalMapping::SearchInNeighbors): for every candidate map point the key-frame feature it would be fused into.  The
// points are independent here (Fuse keeps no matched flags); what depends on the order is the map-graph bookkeeping that follows
// each match (Replace / AddObservation / AddMapPoint), which stays with the caller.  Read from the binary: p3Dc = Rcw * p3Dw +
// tcw (gemm), z < 0 rejects, invz = 1.0f / z, x = X * invz, u = fma(x, fx, cx), v = fma(fy, y, cy) (@0x7ab66-0x7abf0),
// KeyFrame::IsInImage, ur = fma(-bf, invz, u) (@0x7b63e), the invariance range, PO.dot(Pn) < 0.5 * dist rejects (@0x7b4c6), level =
// PredictScale(dist, pKF), radius = th * mvScaleFactors[level] (@0x7b515), KeyFrame::GetFeaturesInArea; a candidate needs its octave
// in [level - 1, level] (@0x7b5e1-0x7b5ef) and a reprojection error below the chi-square bound: stereo (mvuRight >= 0) e2 =
// fma(er, er, fma(ex, ex, ey * ey)) with e2 * mvInvLevelSigma2[octave] > 7.8 rejecting, mono e2 = fma(ex, ex, ey * ey) against 5.99
// (products in float, comparison in double, @0x7b645-0x7b668, @0x7b961-0x7b97b); best = strictly smaller distance; a match needs
// bestDist <= TH_LOW.  mpValid[i] = pMP && !isBad() && !IsInKeyFrame(pKF).  Tcw = Rcw | tcw (3x4), cam5 = fx fy cx cy mbf.
void oracle_fuse_search(int M, const uint8_t* mpValid, const float* mpXYZ, const float* mpNormal, const float* mpDistRange,
                        const uint8_t* mpDesc, int N, const float* kfXY, const int* kfOctave, const float* kfURight,
                        const uint8_t* kfDesc, const int* gridStart, const int* gridItems, int gridCols, int gridRows,
                        const float* Tcw, const float* Ow, const float* cam5, const int* bounds4, float gwi, float ghi,
                        const float* scaleFactors, const float* invLevelSigma2, int nLevels, float logScaleFactor, float th,
                        int* bestIdxOut) {
  (void)N;
  const float fx = cam5[0], fy = cam5[1], cx = cam5[2], cy = cam5[3], bf = cam5[4];
  const int mnMinX = bounds4[0], mnMinY = bounds4[1], mnMaxX = bounds4[2], mnMaxY = bounds4[3];
  for (int i = 0; i < M; ++i) {
    bestIdxOut[i] = -1;
    if (!mpValid[i]) continue;
    const float* X = mpXYZ + 3 * i;
    float pc[3];
    for (int r = 0; r < 3; ++r) {
      const float p0 = Tcw[r * 4] * X[0], p1 = Tcw[r * 4 + 1] * X[1], p2 = Tcw[r * 4 + 2] * X[2];
      const float s = (p0 + p1) + p2;
      pc[r] = (float)((double)s + (double)Tcw[r * 4 + 3]);
    }
    if (pc[2] < 0.0f) continue;
    const float invz = 1.0f / pc[2];
    const float x = pc[0] * invz, y = pc[1] * invz;
    const float u = std::fmaf(x, fx, cx), v = std::fmaf(fy, y, cy);
    if (!(u >= (float)mnMinX && u < (float)mnMaxX && v >= (float)mnMinY && v < (float)mnMaxY)) continue;
    const float ur = std::fmaf(-bf, invz, u);
    float PO[3];
    double n2 = 0;
    for (int r = 0; r < 3; ++r) {
      PO[r] = X[r] - Ow[r];
      n2 += (double)PO[r] * (double)PO[r];
    }
    const float dist = (float)std::sqrt(n2);
    if (dist < 0.8f * mpDistRange[2 * i] || dist > 1.2f * mpDistRange[2 * i + 1]) continue;
    double dot = 0;
    for (int r = 0; r < 3; ++r) dot += (double)PO[r] * (double)mpNormal[3 * i + r];
    if (dot < 0.5 * (double)dist) continue;
    const int level = predict_scale(mpDistRange[2 * i + 1], dist, logScaleFactor, nLevels);
    const float radius = th * scaleFactors[level];
    int bestDist = 256, bestIdx = -1;
    for_kf_features_in_area(u, v, radius, mnMinX, mnMinY, gwi, ghi, gridCols, gridRows, gridStart, gridItems, kfXY, [&](int idx) {
      const int kpLevel = kfOctave[idx];
      if (kpLevel < level - 1 || kpLevel > level) return;
      const float ex = u - kfXY[2 * idx], ey = v - kfXY[2 * idx + 1];
      if (kfURight[idx] >= 0) {
        const float er = ur - kfURight[idx];
        const float e2 = std::fmaf(er, er, std::fmaf(ex, ex, ey * ey));
        if ((double)(e2 * invLevelSigma2[kpLevel]) > 7.8) return;
      } else {
        const float e2 = std::fmaf(ex, ex, ey * ey);
        if ((double)(e2 * invLevelSigma2[kpLevel]) > 5.99) return;
      }
      const int d = descriptor_distance(mpDesc + 32 * i, kfDesc + 32 * idx);
      if (d < bestDist) { bestDist = d; bestIdx = idx; }
    });
    if (bestDist <= TH_LOW) bestIdxOut[i] = bestIdx;
  }
}

// The matching core of ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th,
// vector<MapPoint*>& vpReplacePoint) (ORBmatcher.h:122; @0x7bb20, LoopClosing::SearchAndFuse): the similarity is taken apart as in
// SearchByProjection(KeyFrame*, Scw, ...) above, the per-point tests are those of the rigid Fuse without the chi-square test and
// without ur (u = fma(x, fx, cx), v = fma(y, fy, cy) @0x7caac-0x7cabe; 0.5 * dist @0x7d429; th * mvScaleFactors[level] @0x7d47a;
// bestDist <= TH_LOW @0x7d6fc).  mpValid[i] = !isBad() && pMP not among pKF->GetMapPoints().
void oracle_fuse_search_sim3(int M, const uint8_t* mpValid, const float* mpXYZ, const float* mpNormal, const float* mpDistRange,
                             const uint8_t* mpDesc, int N, const float* kfXY, const int* kfOctave, const uint8_t* kfDesc,
                             const int* gridStart, const int* gridItems, int gridCols, int gridRows, const float* Scw,
                             const float* cam4, const int* bounds4, float gwi, float ghi, const float* scaleFactors, int nLevels,
                             float logScaleFactor, float th, int* bestIdxOut) {
  (void)N;
  const float fx = cam4[0], fy = cam4[1], cx = cam4[2], cy = cam4[3];
  const int mnMinX = bounds4[0], mnMinY = bounds4[1], mnMaxX = bounds4[2], mnMaxY = bounds4[3];
  double dot0 = 0;
  for (int k = 0; k < 3; ++k) dot0 += (double)Scw[k] * (double)Scw[k];
  const float inv = (float)(1.0 / (double)(float)std::sqrt(dot0));
  float T[12], Ow[3];
  for (int k = 0; k < 12; ++k) T[k] = Scw[k] * inv + 0.0f;
  for (int r = 0; r < 3; ++r) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += (double)T[k * 4 + r] * (double)T[k * 4 + 3];
    Ow[r] = (float)(-1.0 * s);
  }
  for (int i = 0; i < M; ++i) {
    bestIdxOut[i] = -1;
    if (!mpValid[i]) continue;
    const float* X = mpXYZ + 3 * i;
    float pc[3];
    for (int r = 0; r < 3; ++r) {
      const float p0 = T[r * 4] * X[0], p1 = T[r * 4 + 1] * X[1], p2 = T[r * 4 + 2] * X[2];
      const float s = (p0 + p1) + p2;
      pc[r] = (float)((double)s + (double)T[r * 4 + 3]);
    }
    if (pc[2] < 0.0f) continue;
    const float invz = 1.0f / pc[2];
    const float u = std::fmaf(pc[0] * invz, fx, cx), v = std::fmaf(pc[1] * invz, fy, cy);
    if (!(u >= (float)mnMinX && u < (float)mnMaxX && v >= (float)mnMinY && v < (float)mnMaxY)) continue;
    float PO[3];
    double n2 = 0;
    for (int r = 0; r < 3; ++r) {
      PO[r] = X[r] - Ow[r];
      n2 += (double)PO[r] * (double)PO[r];
    }
    const float dist = (float)std::sqrt(n2);
    if (dist < 0.8f * mpDistRange[2 * i] || dist > 1.2f * mpDistRange[2 * i + 1]) continue;
    double dot = 0;
    for (int r = 0; r < 3; ++r) dot += (double)PO[r] * (double)mpNormal[3 * i + r];
    if (dot < 0.5 * (double)dist) continue;
    const int level = predict_scale(mpDistRange[2 * i + 1], dist, logScaleFactor, nLevels);
    const float radius = th * scaleFactors[level];
    int bestDist = 256, bestIdx = -1;
    for_kf_features_in_area(u, v, radius, mnMinX, mnMinY, gwi, ghi, gridCols, gridRows, gridStart, gridItems, kfXY, [&](int idx) {
      const int kpLevel = kfOctave[idx];
      if (kpLevel < level - 1 || kpLevel > level) return;
      const int d = descriptor_distance(mpDesc + 32 * i, kfDesc + 32 * idx);
      if (d < bestDist) { bestDist = d; bestIdx = idx; }
    });
    if (bestDist <= TH_LOW) bestIdxOut[i] = bestIdx;
  }
}

// ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
// const cv::Mat& t12, const float th) (ORBmatcher.h:116; @0x838b0, LoopClosing::ComputeSim3).  Read from the binary: sR12 = s12 *
// R12 and sR21 = (1.0 / s12) * R12.t() are scaled conversions (every element times (float)alpha; 1.0 / s12 in double, @0x83a8e),
// t21 = -sR21 * t12 is a gemm with scale -1; per map point of one key frame that exists, is not bad and is not matched on entry:
// p3Dc = R?w * p3Dw + t?w, then into the other camera with sR21 / t21 (or sR12 / t12), z < 0 rejects, u = fma(x, fx, cx) with the
// OTHER key frame's intrinsics, IsInImage, dist3D = (float)cv::norm(p3Dc in the other camera) inside the invariance range,
// PredictScale, radius = th * mvScaleFactors[level], GetFeaturesInArea, octave in [level - 1, level], best = strictly smaller
// distance starting from INT_MAX, accepted when bestDist <= TH_HIGH (@0x86692, @0x866e9); a pair is kept when both directions
// agree.  matched1[i] = vpMatches12[i] != NULL on entry, matched2 = the KF2 features those points are observed at.
// Returns nFound; match12[N1] = KF2 feature newly matched to KF1 feature i (-1 = none).
static void sim3_direction(int N, const uint8_t* valid, const uint8_t* already, const float* xyz, const float* distRange,
                           const uint8_t* desc, const float* Tw, const float* sR, const float* tt, const float* camO,
                           const int* boundsO, float gwi, float ghi, int cols, int rows, const int* gridStart, const int* gridItems,
                           const float* xyO, const int* octO, const uint8_t* descO, const float* scaleFactors, int nLevels,
                           float logScaleFactor, float th, int* match) {
  const float fx = camO[0], fy = camO[1], cx = camO[2], cy = camO[3];
  for (int i = 0; i < N; ++i) {
    match[i] = -1;
    if (!valid[i] || already[i]) continue;
    const float* X = xyz + 3 * i;
    float pa[3], pb[3];
    for (int r = 0; r < 3; ++r) {
      const float p0 = Tw[r * 4] * X[0], p1 = Tw[r * 4 + 1] * X[1], p2 = Tw[r * 4 + 2] * X[2];
      pa[r] = (float)((double)((p0 + p1) + p2) + (double)Tw[r * 4 + 3]);
    }
    for (int r = 0; r < 3; ++r) {
      const float p0 = sR[r * 3] * pa[0], p1 = sR[r * 3 + 1] * pa[1], p2 = sR[r * 3 + 2] * pa[2];
      pb[r] = (float)((double)((p0 + p1) + p2) + (double)tt[r]);
    }
    if (pb[2] < 0.0f) continue;
    const float invz = 1.0f / pb[2];
    const float u = std::fmaf(pb[0] * invz, fx, cx), v = std::fmaf(pb[1] * invz, fy, cy);
    if (!(u >= (float)boundsO[0] && u < (float)boundsO[2] && v >= (float)boundsO[1] && v < (float)boundsO[3])) continue;
    double n2 = 0;
    for (int r = 0; r < 3; ++r) n2 += (double)pb[r] * (double)pb[r];
    const float dist = (float)std::sqrt(n2);
    if (dist < 0.8f * distRange[2 * i] || dist > 1.2f * distRange[2 * i + 1]) continue;
    const int level = predict_scale(distRange[2 * i + 1], dist, logScaleFactor, nLevels);
    const float radius = th * scaleFactors[level];
    int bestDist = INT_MAX, bestIdx = -1;
    for_kf_features_in_area(u, v, radius, boundsO[0], boundsO[1], gwi, ghi, cols, rows, gridStart, gridItems, xyO, [&](int idx) {
      if (octO[idx] < level - 1 || octO[idx] > level) return;
      const int d = descriptor_distance(desc + 32 * i, descO + 32 * idx);
      if (d < bestDist) { bestDist = d; bestIdx = idx; }
    });
    if (bestDist <= TH_HIGH) match[i] = bestIdx;
  }
}
// the three derived transforms, as OpenCV evaluates the reference's expressions (see above)
void oracle_sim3_transforms(float s12, const float* R12, const float* t12, float* sR12, float* sR21, float* t21) {
  const float a12 = (float)(double)s12, a21 = (float)(1.0 / (double)s12);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      sR12[3 * i + j] = R12[3 * i + j] * a12 + 0.0f;
      sR21[3 * i + j] = R12[3 * j + i] * a21 + 0.0f;
    }
  for (int i = 0; i < 3; ++i) {
    const float p0 = sR21[3 * i] * t12[0], p1 = sR21[3 * i + 1] * t12[1], p2 = sR21[3 * i + 2] * t12[2];
    t21[i] = (float)((double)((p0 + p1) + p2) * -1.0);
  }
}
int oracle_search_by_sim3(int N1, const uint8_t* valid1, const uint8_t* matched1, const float* xyz1, const float* range1,
                          const uint8_t* mdesc1, const float* kxy1, const int* koct1, const uint8_t* kdesc1, const int* gs1,
                          const int* gi1, const float* T1w, const float* cam1, const int* bounds1, int N2, const uint8_t* valid2,
                          const uint8_t* matched2, const float* xyz2, const float* range2, const uint8_t* mdesc2, const float* kxy2,
                          const int* koct2, const uint8_t* kdesc2, const int* gs2, const int* gi2, const float* T2w,
                          const float* cam2, const int* bounds2, float gwi, float ghi, int cols, int rows, const float* scaleFactors,
                          int nLevels, float logScaleFactor, float s12, const float* R12, const float* t12, float th, int* match12) {
  float sR12[9], sR21[9], t21[3];
  oracle_sim3_transforms(s12, R12, t12, sR12, sR21, t21);
  std::vector<int> m1(std::max(N1, 1)), m2(std::max(N2, 1));
  sim3_direction(N1, valid1, matched1, xyz1, range1, mdesc1, T1w, sR21, t21, cam2, bounds2, gwi, ghi, cols, rows, gs2, gi2, kxy2, koct2,
                 kdesc2, scaleFactors, nLevels, logScaleFactor, th, m1.data());
  sim3_direction(N2, valid2, matched2, xyz2, range2, mdesc2, T2w, sR12, t12, cam1, bounds1, gwi, ghi, cols, rows, gs1, gi1, kxy1, koct1,
                 kdesc1, scaleFactors, nLevels, logScaleFactor, th, m2.data());
  int nFound = 0;
  for (int i1 = 0; i1 < N1; ++i1) {
    match12[i1] = -1;
    const int idx2 = m1[i1];
    if (idx2 >= 0 && m2[idx2] == i1) {
      match12[i1] = idx2;
      nFound++;
    }
  }
  return nFound;
}

// Frame::isInFrustum(MapPoint* pMP, float viewingCosLimit) (include/Frame.h:107-ish "isInFrustum"; @0xf5190), what
// Tracking::SearchLocalPoints runs on every local map point before ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th)
// (it fills the mTrack* fields that matcher reads).  Read from the binary: Pc = mRcw * P + mtcw (gemm small path); PcZ < 0
// rejects (@0xf5740; zero passes); invz = 1.0f / PcZ (vdivss @0xf5759); u = fma(PcX * fx, invz, cx), v likewise (@0xf5761-
// 0xf57c0); bounds tests; PO = P - mOw; dist = (float)cv::norm(PO); 0.8f * mfMinDistance > dist or dist > 1.2f * mfMaxDistance
// rejects (@0xf5b15-0xf5b21); viewCos = (float)(PO.dot(Pn) / (double)dist) (dot in double, vdivsd @0xf5da6); viewingCosLimit >
// viewCos rejects (@0xf5dae); level = PredictScale(dist, this); mTrackProjXR = fma(-invz, mbf, u) (vfnmadd132ss @0xf5dec).
// cam = {fx, fy, cx, cy, mnMinX, mnMaxX, mnMinY, mnMaxY}; Tcw 3x4 row-major (mRcw | mtcw); Ow = mOw.  Rejected points keep
// inView = 0 and zeros in the other outputs.
void oracle_is_in_frustum(int M, const float* xyz, const float* normal, const float* distRange, const float* cam,
                          const float* Tcw, const float* Ow, float mbf, float logScaleFactor, int nLevels, float cosLimit,
                          uint8_t* inView, float* proj, int* level, float* viewCos) {
  const float fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3], mnMinX = cam[4], mnMaxX = cam[5], mnMinY = cam[6], mnMaxY = cam[7];
  for (int i = 0; i < M; ++i) {
    inView[i] = 0;
    proj[3 * i] = proj[3 * i + 1] = proj[3 * i + 2] = 0.f;
    level[i] = 0;
    viewCos[i] = 0.f;
    const float* X = xyz + 3 * i;
    float pc[3];
    for (int r = 0; r < 3; ++r) {
      const float p0 = Tcw[r * 4] * X[0], p1 = Tcw[r * 4 + 1] * X[1], p2 = Tcw[r * 4 + 2] * X[2];
      const float s = (p0 + p1) + p2;
      pc[r] = (float)((double)s + (double)Tcw[r * 4 + 3]);
    }
    if (pc[2] < 0.0f) continue;
    const float invz = 1.0f / pc[2];
    const float u = std::fmaf(pc[0] * fx, invz, cx);
    if (u < mnMinX || u > mnMaxX) continue;
    const float v = std::fmaf(pc[1] * fy, invz, cy);
    if (v < mnMinY || v > mnMaxY) continue;
    float PO[3];
    double n2 = 0;
    for (int r = 0; r < 3; ++r) {
      PO[r] = X[r] - Ow[r];
      n2 += (double)PO[r] * (double)PO[r];
    }
    const float dist = (float)std::sqrt(n2);
    if (0.8f * distRange[2 * i] > dist || dist > 1.2f * distRange[2 * i + 1]) continue;
    double dot = 0;
    for (int r = 0; r < 3; ++r) dot += (double)PO[r] * (double)normal[3 * i + r];
    const float vc = (float)(dot / (double)dist);
    if (cosLimit > vc) continue;
    level[i] = predict_scale(distRange[2 * i + 1], dist, logScaleFactor, nLevels);
    inView[i] = 1;
    proj[3 * i] = u;
    proj[3 * i + 1] = v;
    proj[3 * i + 2] = std::fmaf(-invz, mbf, u);
    viewCos[i] = vc;
  }
}

// ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th) (reference
// include/ORBmatcher.h:61; lib/libORB_SLAM2.so@0x79f10, called by Tracking::SearchLocalPoints).  Read from the binary:
// mbTrackInView (+0x30) / isBad() gate @0x79fab-0x79fbc; RadiusByViewingCos(&mTrackViewCos) @0x79fd9 (2.5 / 4.0, @0x79b60);
// `th != 1.0 => r *= th` @0x79fe4-0x7a002; GetFeaturesInArea(mTrackProjX, mTrackProjY, r * mvScaleFactors[level],
// level - 1, level) @0x7a051; bestDist = bestDist2 = 256 @0x7a09a; Observations() > 0 skip @0x7a0d8; mvuRight > 0 =>
// |mTrackProjXR - uR| > r * sf skips @0x7a113-0x7a146; dist < bestDist / else dist < bestDist2 @0x7a192,0x7a368;
// bestDist <= TH_HIGH (100) @0x7a27b; same level && bestDist > mfNNratio * bestDist2 rejects @0x7a286,0x7a39c.
// Map points flattened: mpValid = mbTrackInView && !isBad(), mpProj = (mTrackProjX, mTrackProjY, mTrackProjXR),
// mpLevel = mnTrackScaleLevel, mpViewCos = mTrackViewCos, mpDesc = GetDescriptor(), mpObs = Observations() > 0.
// cam4 = {mnMinX, mnMinY, gridWInv, gridHInv}.  matchF[N] = map point index assigned to each frame keypoint (-1 = none).
int oracle_search_local_points(int M, const uint8_t* mpValid, const float* mpProj, const int* mpLevel, const float* mpViewCos,
                               const uint8_t* mpDesc, const uint8_t* mpObs, int N, const float* fXY, const int* fOctave,
                               const uint8_t* fDesc, const float* fURight, const uint8_t* fTaken, const int* gridStart,
                               const int* gridItems, const float* cam4, const float* scaleFactors, float th, float nnratio,
                               int* matchF) {
  const float mnMinX = cam4[0], mnMinY = cam4[1], gwi = cam4[2], ghi = cam4[3];
  std::vector<uint8_t> taken(fTaken, fTaken + N);
  for (int i = 0; i < N; ++i) matchF[i] = -1;
  int nmatches = 0;
  const bool bFactor = th != 1.0f;
  for (int iMP = 0; iMP < M; ++iMP) {
    if (!mpValid[iMP]) continue;
    const int nPredictedLevel = mpLevel[iMP];
    float r = radius_by_viewing_cos(mpViewCos[iMP]);
    if (bFactor) r *= th;
    const float x = mpProj[3 * iMP], y = mpProj[3 * iMP + 1], xr = mpProj[3 * iMP + 2];
    const float radius = r * scaleFactors[nPredictedLevel];
    const int minLevel = nPredictedLevel - 1, maxLevel = nPredictedLevel;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for_features_in_area(x, y, radius, minLevel, maxLevel, mnMinX, mnMinY, gwi, ghi, gridStart, gridItems, fXY, fOctave, [&](int idx) {
      if (taken[idx]) return;
      if (fURight[idx] > 0) {
        const float er = fabsf(xr - fURight[idx]);
        if (er > radius) return;
      }
      const int dist = descriptor_distance(mpDesc + 32 * iMP, fDesc + 32 * idx);
      if (dist < bestDist) {
        bestDist2 = bestDist;
        bestDist = dist;
        bestLevel2 = bestLevel;
        bestLevel = fOctave[idx];
        bestIdx = idx;
      } else if (dist < bestDist2) {
        bestLevel2 = fOctave[idx];
        bestDist2 = dist;
      }
    });
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue;
      matchF[bestIdx] = iMP;
      if (mpObs[iMP]) taken[bestIdx] = 1;
      nmatches++;
    }
  }
  return nmatches;
}

// ORBmatcher::CheckDistEpipolarLine(kp1, kp2, F12, pKF2) (ORBmatcher.h:89; machine code @0x79b90-0x79c32, compiled with FMA:
// the fused operations below are the binary's — vfmadd231ss @0x79bb3, @0x79bbf, @0x79bcd, @0x79bf3, vfmadd132ss @0x79c00).
// F: row-major 3x3 float.  3.84 is a double (@0x126a08), the comparison is made in double (@0x79c1b-0x79c2f).
static bool check_dist_epipolar_line(float x1, float y1, float x2, float y2, const float* F, float sigma2_kp2) {
  const float b = std::fmaf(x1, F[1], y1 * F[4]) + F[7];
  const float a = std::fmaf(x1, F[0], y1 * F[3]) + F[6];
  const float den = std::fmaf(a, a, b * b);
  if (den == 0.0f) return false;
  const float c = std::fmaf(y1, F[5], x1 * F[2]) + F[8];
  const float num = c + std::fmaf(b, y2, a * x2);
  const float dsqr = (num * num) / den;
  return 3.84 * (double)sigma2_kp2 > (double)dsqr;
}
int oracle_check_dist_epipolar_line(float x1, float y1, float x2, float y2, const float* F, float sigma2_kp2) {
  return check_dist_epipolar_line(x1, y1, x2, y2, F, sigma2_kp2) ? 1 : 0;
}

// The epipole of KF1's camera centre in KF2, as SearchForTriangulation computes it before its loops (@0x86b9c-0x86f8b):
// C2 = R2w * Cw + t2w is one cv::gemm (3x3 by 3x1 CV_32F: small-matrix path, float sum, (double)s * 1 + (double)t * 1),
// invz = 1.0f / C2z (@0x86ee0), ex = fma(invz, fx * C2x, cx) (@0x86eee-0x86ef2), ey likewise (@0x86f4a-0x86f8b).
void oracle_epipole(const float* R2w, const float* t2w, const float* Cw, float fx, float fy, float cx, float cy, float* ex, float* ey) {
  float C2[3];
  for (int r = 0; r < 3; ++r) {
    const float p0 = R2w[3 * r] * Cw[0], p1 = R2w[3 * r + 1] * Cw[1], p2 = R2w[3 * r + 2] * Cw[2];
    const float s = (p0 + p1) + p2;
    C2[r] = (float)((double)s * 1.0 + (double)t2w[r] * 1.0);
  }
  const float invz = 1.0f / C2[2];
  *ex = std::fmaf(invz, fx * C2[0], cx);
  *ey = std::fmaf(invz, fy * C2[1], cy);
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (ORBmatcher.h:86; @0x86b30).
// Read from the binary: candidates only inside a shared vocabulary node (merge of the two FeatureVectors, lower_bound on
// mismatch); KF1 features with a map point are skipped (@0x87836-0x8783e), with bOnlyStereo also those with mvuRight < 0
// (@0x87869-0x87876); bestDist starts at TH_LOW = 50 (@0x87915); a KF2 feature is skipped when already matched
// (vbMatched2 bit test @0x87977 — this build DOES set the bit on acceptance, @0x87bc9) or when it has a map point (@0x8797d),
// with bOnlyStereo when mvuRight < 0 (@0x879a0-0x879a8); dist > 50 or dist > bestDist skips (@0x87a01-0x87a13: an equal
// distance replaces the earlier candidate); when neither feature is stereo the candidate must be at least
// sqrt(100 * mvScaleFactors[octave2]) away from the epipole: fma(dx, dx, dy * dy) against 100.0f * scale (@0x87a52-0x87a91);
// CheckDistEpipolarLine decides (@0x87aaf).  Accepted: vMatches12[idx1] = bestIdx2, bit set, rotation histogram with
// kp1.angle - kp2.angle (@0x87bf1-0x87c2d), ComputeThreeMaxima, pairs collected in idx1 order.
// hasMP*: GetMapPoint(i) != NULL.  match12[N1] receives vMatches12.  Returns nmatches.
int oracle_search_for_triangulation(int N1, const uint8_t* d1, const float* xy1, const float* ang1, const float* uright1,
                                    const uint8_t* hasMP1, int nNodes1, const int* nodes1, const int* start1, const int* idx1v,
                                    int N2, const uint8_t* d2, const float* xy2, const float* ang2, const int* oct2,
                                    const float* uright2, const uint8_t* hasMP2, int nNodes2, const int* nodes2, const int* start2,
                                    const int* idx2v, const float* scaleFactors2, const float* levelSigma2_2, const float* F12,
                                    float ex, float ey, int onlyStereo, int checkOri, int* match12) {
  for (int i = 0; i < N1; ++i) match12[i] = -1;
  std::vector<uint8_t> matched2(N2, 0);
  std::vector<int> rotHist[HISTO_LENGTH];
  int nmatches = 0;
  int a = 0, b = 0;
  while (a < nNodes1 && b < nNodes2) {
    if (nodes1[a] == nodes2[b]) {
      for (int i1 = start1[a]; i1 < start1[a + 1]; ++i1) {
        const int idx1 = idx1v[i1];
        if (hasMP1[idx1]) continue;
        const bool bStereo1 = uright1[idx1] >= 0.0f;
        if (onlyStereo && !bStereo1) continue;
        int bestDist = TH_LOW, bestIdx2 = -1;
        for (int i2 = start2[b]; i2 < start2[b + 1]; ++i2) {
          const int idx2 = idx2v[i2];
          if (matched2[idx2] || hasMP2[idx2]) continue;
          const bool bStereo2 = uright2[idx2] >= 0.0f;
          if (onlyStereo && !bStereo2) continue;
          const int dist = descriptor_distance(d1 + 32 * idx1, d2 + 32 * idx2);
          if (dist > TH_LOW || dist > bestDist) continue;
          if (!bStereo1 && !bStereo2) {
            const float distex = ex - xy2[2 * idx2], distey = ey - xy2[2 * idx2 + 1];
            if (100.0f * scaleFactors2[oct2[idx2]] > std::fmaf(distex, distex, distey * distey)) continue;
          }
          if (check_dist_epipolar_line(xy1[2 * idx1], xy1[2 * idx1 + 1], xy2[2 * idx2], xy2[2 * idx2 + 1], F12,
                                       levelSigma2_2[oct2[idx2]])) {
            bestIdx2 = idx2;
            bestDist = dist;
          }
        }
        if (bestIdx2 >= 0) {
          match12[idx1] = bestIdx2;
          matched2[bestIdx2] = 1;
          nmatches++;
          if (checkOri) rotHist[rot_bin(ang1[idx1], ang2[bestIdx2])].push_back(idx1);
        }
      }
      ++a; ++b;
    } else if (nodes1[a] < nodes2[b]) {
      a = (int)(std::lower_bound(nodes1, nodes1 + nNodes1, nodes2[b]) - nodes1);
    } else {
      b = (int)(std::lower_bound(nodes2, nodes2 + nNodes2, nodes1[a]) - nodes2);
    }
  }
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    compute_three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int j : rotHist[i]) { match12[j] = -1; nmatches--; }
    }
  }
  return nmatches;
}

// ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched, vector<int>& vnMatches12,
// int windowSize) (ORBmatcher.h:82; @0x7db00) — the monocular-initialisation matcher.  Only level-0 key points of F1 are used;
// candidates are F2's level-0 key points inside the window around vbPrevMatched[i1] (Frame::GetFeaturesInArea(x, y, windowSize,
// 0, 0)); a candidate already matched with a distance <= the present one is skipped; best / second best by strict '<';
// accepted when bestDist <= TH_LOW and bestDist < bestDist2 * mfNNratio; a re-matched F2 key point releases its earlier F1
// partner; rotation histogram over accepted i1 (angle of F1 minus angle of F2), three-maxima filter; finally
// vbPrevMatched[i1] = F2 position of every surviving match.  Order dependent (vMatchedDistance / vnMatches21), restated
// sequentially.  Pinned by running the reference's own function (tests/golden/reference_code.py, fixture si*).
// xy1/oct1/ang1: F1.mvKeysUn; F2 likewise with its grid as CSR; prevMatched: N1 x 2 floats, updated in place.
int oracle_search_for_initialization(int N1, const float* xy1, const int* oct1, const float* ang1, const uint8_t* d1, int N2,
                                     const float* xy2, const int* oct2, const float* ang2, const uint8_t* d2, const int* gridStart,
                                     const int* gridItems, const float* cam4, float* prevMatched, int windowSize, float nnratio,
                                     int checkOri, int* matches12) {
  (void)xy1;
  int nmatches = 0;
  for (int i = 0; i < N1; ++i) matches12[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  std::vector<int> vMatchedDistance(N2, INT_MAX), vnMatches21(N2, -1);
  for (int i1 = 0; i1 < N1; ++i1) {
    const int level1 = oct1[i1];
    if (level1 > 0) continue;
    int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
    bool any = false;
    for_features_in_area(prevMatched[2 * i1], prevMatched[2 * i1 + 1], (float)windowSize, level1, level1, cam4[0], cam4[1], cam4[2],
                         cam4[3], gridStart, gridItems, xy2, oct2, [&](int i2) {
                           any = true;
                           const int dist = descriptor_distance(d1 + 32 * i1, d2 + 32 * i2);
                           if (vMatchedDistance[i2] <= dist) return;
                           if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
                           else if (dist < bestDist2) { bestDist2 = dist; }
                         });
    if (!any) continue;
    if (bestDist <= TH_LOW) {
      if (bestDist < (float)bestDist2 * nnratio) {
        if (vnMatches21[bestIdx2] >= 0) { matches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
        matches12[i1] = bestIdx2;
        vnMatches21[bestIdx2] = i1;
        vMatchedDistance[bestIdx2] = bestDist;
        nmatches++;
        if (checkOri) rotHist[rot_bin(ang1[i1], ang2[bestIdx2])].push_back(i1);
      }
    }
  }
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    compute_three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i])
        if (matches12[idx1] >= 0) { matches12[idx1] = -1; nmatches--; }
    }
  }
  for (int i1 = 0; i1 < N1; ++i1)
    if (matches12[i1] >= 0) { prevMatched[2 * i1] = xy2[2 * matches12[i1]]; prevMatched[2 * i1 + 1] = xy2[2 * matches12[i1] + 1]; }
  return nmatches;
}

}  // extern "C"
