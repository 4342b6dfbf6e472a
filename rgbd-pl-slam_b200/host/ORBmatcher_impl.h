// ORBmatcher_impl.h — the reference-signature member templates of ORB_SLAM2::ORBmatcher (declared in ORBmatcher.h):
// each flattens the members the reference's function reads into a FrameView / MapPointsView, calls the flattened form
// (one C-ABI call, one kernel) and writes the result where the reference writes it.  Included by ORBmatcher.h.
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>

namespace ORB_SLAM2 {
namespace dropin {

// DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned int>>) -> CSR, keys ascending as the map iterates
template <class FeatVec>
inline void flatten_featvec(const FeatVec& fv, FeatureVectorView* out) {
  out->nodes.clear();
  out->start.clear();
  out->idx.clear();
  out->start.push_back(0);
  for (typename FeatVec::const_iterator it = fv.begin(); it != fv.end(); ++it) {
    out->nodes.push_back((int32_t)it->first);
    for (size_t k = 0; k < it->second.size(); ++k) out->idx.push_back((int32_t)it->second[k]);
    out->start.push_back((int32_t)out->idx.size());
  }
}

// every array of a view must cover its N features: an unfilled member would be read out of bounds by the upload
inline void require(bool ok, const char* what) {
  if (!ok) throw std::invalid_argument(std::string("ORBmatcher: ") + what);
}

// The features, grid, bounds, intrinsics and pose of a Frame.  Map-point state is filled by the callers (it differs).
template <class FrameT>
inline void view_of_frame(const FrameT& F, FrameView* V, bool withGrid) {
  V->mvKeys = F.mvKeys;
  V->mvKeysUn = F.mvKeysUn;
  V->mDescriptors = F.mDescriptors;
  V->mvuRight = F.mvuRight;
  V->mvScaleFactors = F.mvScaleFactors;
  const size_t n = F.mvKeysUn.size();
  require(F.mvKeys.size() == n && F.mvuRight.size() == n && (size_t)F.mDescriptors.rows == n && F.mvpMapPoints.size() == n,
          "Frame members (mvKeys, mvKeysUn, mvuRight, mDescriptors, mvpMapPoints) differ in length");
  if (withGrid) {
    V->gridStart.assign(1, 0);
    V->gridItems.clear();
    for (int ix = 0; ix < 64; ++ix)
      for (int iy = 0; iy < 48; ++iy) {
        for (size_t k = 0; k < F.mGrid[ix][iy].size(); ++k) V->gridItems.push_back((int32_t)F.mGrid[ix][iy][k]);
        V->gridStart.push_back((int32_t)V->gridItems.size());
      }
    V->mnMinX = FrameT::mnMinX; V->mnMaxX = FrameT::mnMaxX; V->mnMinY = FrameT::mnMinY; V->mnMaxY = FrameT::mnMaxY;
    V->mfGridElementWidthInv = FrameT::mfGridElementWidthInv;
    V->mfGridElementHeightInv = FrameT::mfGridElementHeightInv;
  }
  V->fx = FrameT::fx; V->fy = FrameT::fy; V->cx = FrameT::cx; V->cy = FrameT::cy;
  V->mbf = F.mbf; V->mb = F.mb;
  if (!F.mTcw.empty())
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) V->mTcw[4 * r + c] = F.mTcw.template at<float>(r, c);
}

// What the key-frame projection matchers read of a KeyFrame (pKF->mGrid is protected in the reference's KeyFrame.h: befriend
// ORBmatcher there or add a const accessor).
template <class KeyFrameT>
inline void view_of_keyframe_grid(KeyFrameT* pKF, KeyFrameGridView* kv) {
  kv->mvKeysUn = pKF->mvKeysUn;
  kv->mDescriptors = pKF->mDescriptors;
  kv->mvScaleFactors = pKF->mvScaleFactors;
  kv->mnGridCols = pKF->mnGridCols; kv->mnGridRows = pKF->mnGridRows;
  kv->mfGridElementWidthInv = pKF->mfGridElementWidthInv; kv->mfGridElementHeightInv = pKF->mfGridElementHeightInv;
  kv->mnMinX = pKF->mnMinX; kv->mnMinY = pKF->mnMinY; kv->mnMaxX = pKF->mnMaxX; kv->mnMaxY = pKF->mnMaxY;
  kv->fx = pKF->fx; kv->fy = pKF->fy; kv->cx = pKF->cx; kv->cy = pKF->cy;
  require((size_t)pKF->mDescriptors.rows == kv->mvKeysUn.size(), "KeyFrame members differ in length");
  kv->gridStart.assign(1, 0);
  kv->gridItems.clear();
  for (int ix = 0; ix < kv->mnGridCols; ++ix)
    for (int iy = 0; iy < kv->mnGridRows; ++iy) {
      for (size_t k = 0; k < pKF->mGrid[ix][iy].size(); ++k) kv->gridItems.push_back((int32_t)pKF->mGrid[ix][iy][k]);
      kv->gridStart.push_back((int32_t)kv->gridItems.size());
    }
}

// Scw = [s R | s t] taken apart with the reference's arithmetic: s from the first row (dot in double), every element times
// (float)(1.0 / s) as cv::operator/(Mat, double) evaluates, Ow = -Rcw.t() * tcw through the double-accumulating product
inline void decompose_sim3(const float S[12], float T[12], float Ow[3]) {
  double d0 = 0;
  for (int k = 0; k < 3; ++k) d0 += (double)S[k] * (double)S[k];
  const float inv = (float)(1.0 / (double)(float)std::sqrt(d0));
  for (int k = 0; k < 12; ++k) { volatile float p = S[k] * inv; T[k] = p + 0.0f; }
  for (int r = 0; r < 3; ++r) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += (double)T[4 * k + r] * (double)T[4 * k + 3];
    Ow[r] = (float)(-1.0 * s);
  }
}

// One candidate map point of the key-frame projection matchers: the distance from the camera centre with the reference's
// arithmetic, the scale-invariance and viewing-angle tests, pMP->PredictScale; fills row i of the view when the point passes
template <class KeyFrameT, class MapPointT>
inline void loop_point(MapPointT* pMP, KeyFrameT* pKF, const float Ow[3], int i, LoopPointsView* pv) {
  const cv::Mat p3Dw = pMP->GetWorldPos();
  float PO[3];
  double n2 = 0;
  for (int k = 0; k < 3; ++k) {
    const float x = p3Dw.template at<float>(k);
    pv->worldPos[3 * (size_t)i + k] = x;
    PO[k] = x - Ow[k];
    n2 += (double)PO[k] * (double)PO[k];
  }
  const float dist = (float)std::sqrt(n2);
  if (dist < pMP->GetMinDistanceInvariance() || dist > pMP->GetMaxDistanceInvariance()) return;
  const cv::Mat Pn = pMP->GetNormal();
  double dot = 0;
  for (int k = 0; k < 3; ++k) dot += (double)PO[k] * (double)Pn.template at<float>(k);
  if (dot < 0.5 * (double)dist) return;
  pv->level[i] = pMP->PredictScale(dist, pKF);
  const cv::Mat d = pMP->GetDescriptor();
  std::memcpy(pv->descriptors.ptr(i), d.ptr(0), 32);
  pv->valid[i] = 1;
}
inline void init_loop_points(int m, LoopPointsView* pv) {
  pv->valid.assign(m, 0);
  pv->worldPos.assign((size_t)m * 3, 0.f);
  pv->descriptors.create(m > 0 ? m : 1, 32, CV_8U);
  pv->level.assign(m, 0);
}

}  // namespace dropin

// ORBmatcher.h:78 (@0x80d00)
template <class FrameT>
int ORBmatcher::SearchByProjection(FrameT& CurrentFrame, const FrameT& LastFrame, const float th, const bool bMono) {
  FrameView cur, last;
  dropin::view_of_frame(CurrentFrame, &cur, true);
  dropin::view_of_frame(LastFrame, &last, false);
  const int n1 = (int)last.mvKeysUn.size(), n2 = (int)cur.mvKeysUn.size();
  dropin::require(LastFrame.mvbOutlier.size() == (size_t)n1, "LastFrame.mvbOutlier differs in length");
  last.hasMapPoint.assign(n1, 0);
  last.mapPointObserved.assign(n1, 0);
  last.mvbOutlier.assign(n1, 0);
  last.mapPointWorldPos.assign((size_t)n1 * 3, 0.f);
  last.mapPointDescriptor.create(n1 > 0 ? n1 : 1, 32, CV_8U);
  for (int i = 0; i < n1; ++i) {
    auto* pMP = LastFrame.mvpMapPoints[i];
    last.mvbOutlier[i] = LastFrame.mvbOutlier[i] ? 1 : 0;
    if (!pMP) continue;
    last.hasMapPoint[i] = 1;
    last.mapPointObserved[i] = pMP->Observations() > 0;
    const cv::Mat x3Dw = pMP->GetWorldPos();
    for (int k = 0; k < 3; ++k) last.mapPointWorldPos[3 * (size_t)i + k] = x3Dw.template at<float>(k);
    const cv::Mat d = pMP->GetDescriptor();
    std::memcpy(last.mapPointDescriptor.ptr(i), d.ptr(0), 32);
  }
  cur.mapPointObserved.assign(n2, 0);
  for (int i = 0; i < n2; ++i)
    if (CurrentFrame.mvpMapPoints[i] && CurrentFrame.mvpMapPoints[i]->Observations() > 0) cur.mapPointObserved[i] = 1;
  std::vector<int> m;
  const int n = SearchByProjection(cur, last, th, bMono, m, true);
  for (int i2 = 0; i2 < n2; ++i2) {
    if (m[i2] >= 0) CurrentFrame.mvpMapPoints[i2] = LastFrame.mvpMapPoints[m[i2]];
    else if (m[i2] == -2) CurrentFrame.mvpMapPoints[i2] = nullptr;  // assigned, then removed by the rotation check
  }
  return n;
}

// ORBmatcher.h:61 (@0x79f10)
template <class FrameT, class MapPointT>
int ORBmatcher::SearchByProjection(FrameT& F, const std::vector<MapPointT*>& vpMapPoints, const float th) {
  FrameView f;
  dropin::view_of_frame(F, &f, true);
  const int n = (int)f.mvKeysUn.size(), m = (int)vpMapPoints.size();
  f.mapPointObserved.assign(n, 0);
  for (int i = 0; i < n; ++i)
    if (F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0) f.mapPointObserved[i] = 1;
  MapPointsView mp;
  mp.inViewAndGood.assign(m, 0);
  mp.trackProj.assign((size_t)m * 3, 0.f);
  mp.trackScaleLevel.assign(m, 0);
  mp.trackViewCos.assign(m, 0.f);
  mp.observed.assign(m, 0);
  mp.descriptors.create(m > 0 ? m : 1, 32, CV_8U);
  for (int i = 0; i < m; ++i) {
    MapPointT* pMP = vpMapPoints[i];
    if (!pMP || !pMP->mbTrackInView || pMP->isBad()) continue;
    mp.inViewAndGood[i] = 1;
    mp.trackProj[3 * (size_t)i] = pMP->mTrackProjX;
    mp.trackProj[3 * (size_t)i + 1] = pMP->mTrackProjY;
    mp.trackProj[3 * (size_t)i + 2] = pMP->mTrackProjXR;
    mp.trackScaleLevel[i] = pMP->mnTrackScaleLevel;
    mp.trackViewCos[i] = pMP->mTrackViewCos;
    mp.observed[i] = pMP->Observations() > 0;
    const cv::Mat d = pMP->GetDescriptor();
    std::memcpy(mp.descriptors.ptr(i), d.ptr(0), 32);
  }
  std::vector<int> match;
  const int nm = SearchByProjection(f, mp, th, match);
  for (int i = 0; i < n; ++i)
    if (match[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[match[i]];
  return nm;
}

// ORBmatcher.h:82 (@0x7e8c0).  The camera centre and the distance of a point from it are evaluated here with the arithmetic of
// the reference (Ow = -Rcw.t() * tcw: gemm with a transpose flag, double accumulator; PO = x3Dw - Ow in float; cv::norm:
// squares summed in double, sqrt), so that the invariance test and pMP->PredictScale() see the value the reference gives them.
template <class FrameT, class KeyFrameT, class MapPointT>
int ORBmatcher::SearchByProjection(FrameT& CurrentFrame, KeyFrameT* pKF, const std::set<MapPointT*>& sAlreadyFound, const float th,
                                   const int ORBdist) {
  FrameView cur;
  dropin::view_of_frame(CurrentFrame, &cur, true);
  const int n2 = (int)cur.mvKeysUn.size();
  cur.hasMapPoint.assign(n2, 0);
  for (int i = 0; i < n2; ++i) cur.hasMapPoint[i] = CurrentFrame.mvpMapPoints[i] != nullptr;
  const std::vector<MapPointT*> vpMPs = pKF->GetMapPointMatches();
  const int m = (int)vpMPs.size();
  dropin::require(pKF->mvKeysUn.size() == (size_t)m, "KeyFrame members differ in length");
  float Ow[3];
  for (int r = 0; r < 3; ++r) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += (double)cur.mTcw[4 * k + r] * (double)cur.mTcw[4 * k + 3];
    Ow[r] = (float)(-1.0 * s);
  }
  KeyFramePointsView kv;
  kv.valid.assign(m, 0);
  kv.worldPos.assign((size_t)m * 3, 0.f);
  kv.descriptors.create(m > 0 ? m : 1, 32, CV_8U);
  kv.angle.assign(m, 0.f);
  kv.level.assign(m, 0);
  for (int i = 0; i < m; ++i) {
    MapPointT* pMP = vpMPs[i];
    kv.angle[i] = pKF->mvKeysUn[i].angle;
    if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
    const cv::Mat x3Dw = pMP->GetWorldPos();
    double n2sum = 0;
    for (int k = 0; k < 3; ++k) {
      const float x = x3Dw.template at<float>(k);
      kv.worldPos[3 * (size_t)i + k] = x;
      const float d = x - Ow[k];
      n2sum += (double)d * (double)d;
    }
    const float dist3D = (float)std::sqrt(n2sum);
    const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
    if (dist3D < minDistance || dist3D > maxDistance) continue;
    kv.level[i] = pMP->PredictScale(dist3D, &CurrentFrame);
    const cv::Mat d = pMP->GetDescriptor();
    std::memcpy(kv.descriptors.ptr(i), d.ptr(0), 32);
    kv.valid[i] = 1;
  }
  std::vector<int> match;
  const int nm = SearchByProjection(cur, kv, th, ORBdist, match);
  for (int i2 = 0; i2 < n2; ++i2)
    if (match[i2] >= 0) CurrentFrame.mvpMapPoints[i2] = vpMPs[match[i2]];
  return nm;
}

// ORBmatcher.h:86 (@0x880f0).  The similarity is taken apart here with the reference's arithmetic (scale from the first row of
// sRcw, dot in double; every element times (float)(1.0 / scw), as cv::operator/(Mat, double) evaluates; Ow through the
// double-accumulating product), so that the invariance test, the viewing-angle test and pMP->PredictScale() see the reference's
// distance; the kernel repeats the same decomposition for the projections.
template <class KeyFrameT, class MapPointT>
int ORBmatcher::SearchByProjection(KeyFrameT* pKF, cv::Mat Scw, const std::vector<MapPointT*>& vpPoints,
                                   std::vector<MapPointT*>& vpMatched, int th) {
  KeyFrameGridView kv;
  dropin::view_of_keyframe_grid(pKF, &kv);
  const int n = (int)kv.mvKeysUn.size(), m = (int)vpPoints.size();
  dropin::require((int)vpMatched.size() == n, "vpMatched and the key frame's features differ in length");
  float S[12], T[12], Ow[3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) S[4 * r + c] = Scw.template at<float>(r, c);
  dropin::decompose_sim3(S, T, Ow);
  std::set<MapPointT*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
  spAlreadyFound.erase(static_cast<MapPointT*>(nullptr));
  std::vector<uint8_t> matchedOnEntry(n, 0);
  for (int i = 0; i < n; ++i) matchedOnEntry[i] = vpMatched[i] != nullptr;
  LoopPointsView pv;
  dropin::init_loop_points(m, &pv);
  for (int i = 0; i < m; ++i) {
    MapPointT* pMP = vpPoints[i];
    if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
    dropin::loop_point(pMP, pKF, Ow, i, &pv);
  }
  std::vector<int> match;
  const int nm = SearchByProjection(kv, S, pv, matchedOnEntry, th, match);
  for (int i = 0; i < n; ++i)
    if (match[i] >= 0) vpMatched[i] = vpPoints[match[i]];
  return nm;
}

// ORBmatcher.h:119 (@0x7a500)
template <class KeyFrameT, class MapPointT>
int ORBmatcher::Fuse(KeyFrameT* pKF, const std::vector<MapPointT*>& vpMapPoints, const float th) {
  KeyFrameGridView kv;
  dropin::view_of_keyframe_grid(pKF, &kv);
  kv.mvuRight = pKF->mvuRight;
  kv.mvInvLevelSigma2 = pKF->mvInvLevelSigma2;
  kv.mbf = pKF->mbf;
  const cv::Mat Rcw = pKF->GetRotation(), tcw = pKF->GetTranslation(), OwM = pKF->GetCameraCenter();
  float T[12], Ow[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T[4 * r + c] = Rcw.template at<float>(r, c);
    T[4 * r + 3] = tcw.template at<float>(r);
    Ow[r] = OwM.template at<float>(r);
  }
  const int m = (int)vpMapPoints.size();
  LoopPointsView pv;
  dropin::init_loop_points(m, &pv);
  for (int i = 0; i < m; ++i) {
    MapPointT* pMP = vpMapPoints[i];
    if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
    dropin::loop_point(pMP, pKF, Ow, i, &pv);
  }
  std::vector<int> best;
  FuseSearch(kv, T, Ow, false, pv, th, best);
  int nFused = 0;
  for (int i = 0; i < m; ++i) {
    if (best[i] < 0) continue;
    MapPointT* pMP = vpMapPoints[i];
    if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;  // the bookkeeping of an earlier point may have changed it
    const size_t bestIdx = (size_t)best[i];
    MapPointT* pMPinKF = pKF->GetMapPoint(bestIdx);
    if (pMPinKF) {
      if (!pMPinKF->isBad()) {
        if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
        else pMPinKF->Replace(pMP);
      }
    } else {
      pMP->AddObservation(pKF, bestIdx);
      pKF->AddMapPoint(pMP, bestIdx);
    }
    nFused++;
  }
  return nFused;
}

// ORBmatcher.h:122 (@0x7bb20)
template <class KeyFrameT, class MapPointT>
int ORBmatcher::Fuse(KeyFrameT* pKF, cv::Mat Scw, const std::vector<MapPointT*>& vpPoints, float th,
                     std::vector<MapPointT*>& vpReplacePoint) {
  KeyFrameGridView kv;
  dropin::view_of_keyframe_grid(pKF, &kv);
  float S[12], T[12], Ow[3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) S[4 * r + c] = Scw.template at<float>(r, c);
  dropin::decompose_sim3(S, T, Ow);
  const std::set<MapPointT*> spAlreadyFound = pKF->GetMapPoints();
  const int m = (int)vpPoints.size();
  dropin::require((int)vpReplacePoint.size() == m, "vpReplacePoint and vpPoints differ in length");
  LoopPointsView pv;
  dropin::init_loop_points(m, &pv);
  for (int i = 0; i < m; ++i) {
    MapPointT* pMP = vpPoints[i];
    if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
    dropin::loop_point(pMP, pKF, Ow, i, &pv);
  }
  std::vector<int> best;
  FuseSearch(kv, S, nullptr, true, pv, th, best);
  int nFused = 0;
  for (int i = 0; i < m; ++i) {
    if (best[i] < 0) continue;
    MapPointT* pMP = vpPoints[i];
    const size_t bestIdx = (size_t)best[i];
    MapPointT* pMPinKF = pKF->GetMapPoint(bestIdx);
    if (pMPinKF) {
      if (!pMPinKF->isBad()) vpReplacePoint[i] = pMPinKF;
    } else {
      pMP->AddObservation(pKF, bestIdx);
      pKF->AddMapPoint(pMP, bestIdx);
    }
    nFused++;
  }
  return nFused;
}

// ORBmatcher.h:116 (@0x838b0)
namespace dropin {
// a * x + b for a 3x4 [A | b] and a 3-vector, as cv::gemm's small-matrix path evaluates it (float products summed left to right,
// then (double)sum + (double)b rounded to float)
inline void affine3(const float T[12], const float x[3], float y[3]) {
  for (int r = 0; r < 3; ++r) {
    volatile float p0 = T[4 * r] * x[0], p1 = T[4 * r + 1] * x[1], p2 = T[4 * r + 2] * x[2];
    volatile float s01 = p0 + p1;
    volatile float s = s01 + p2;
    y[r] = (float)((double)s + (double)T[4 * r + 3]);
  }
}
template <class KeyFrameT, class MapPointT>
inline void sim3_points(const std::vector<MapPointT*>& vp, const std::vector<uint8_t>& already, const float Tw[12], const float T2[12],
                        KeyFrameT* pKFother, LoopPointsView* pv) {
  const int n = (int)vp.size();
  init_loop_points(n, pv);
  for (int i = 0; i < n; ++i) {
    MapPointT* pMP = vp[i];
    if (!pMP || already[i] || pMP->isBad()) continue;
    const cv::Mat p3Dw = pMP->GetWorldPos();
    float X[3], pa[3], pb[3];
    for (int k = 0; k < 3; ++k) X[k] = pv->worldPos[3 * (size_t)i + k] = p3Dw.template at<float>(k);
    affine3(Tw, X, pa);
    affine3(T2, pa, pb);
    double n2 = 0;
    for (int k = 0; k < 3; ++k) n2 += (double)pb[k] * (double)pb[k];
    const float dist3D = (float)std::sqrt(n2);
    if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
    pv->level[i] = pMP->PredictScale(dist3D, pKFother);
    const cv::Mat d = pMP->GetDescriptor();
    std::memcpy(pv->descriptors.ptr(i), d.ptr(0), 32);
    pv->valid[i] = 1;
  }
}
}  // namespace dropin

template <class KeyFrameT, class MapPointT>
int ORBmatcher::SearchBySim3(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12, const float& s12,
                             const cv::Mat& R12, const cv::Mat& t12, const float th) {
  KeyFrameGridView kv1, kv2;
  dropin::view_of_keyframe_grid(pKF1, &kv1);
  dropin::view_of_keyframe_grid(pKF2, &kv2);
  KeyFrameT* kfs[2] = {pKF1, pKF2};
  float Tw[2][12];
  for (int s = 0; s < 2; ++s) {
    const cv::Mat R = kfs[s]->GetRotation(), t = kfs[s]->GetTranslation();
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) Tw[s][4 * r + c] = R.template at<float>(r, c);
      Tw[s][4 * r + 3] = t.template at<float>(r);
    }
  }
  float R9[9], t3[3], sR12[9], sR21[9], t21[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) R9[3 * r + c] = R12.template at<float>(r, c);
    t3[r] = t12.template at<float>(r);
  }
  Sim3Transforms(s12, R9, t3, sR12, sR21, t21);
  float T21[12], T12[12];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) { T21[4 * r + c] = sR21[3 * r + c]; T12[4 * r + c] = sR12[3 * r + c]; }
    T21[4 * r + 3] = t21[r];
    T12[4 * r + 3] = t3[r];
  }
  const std::vector<MapPointT*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
  const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
  dropin::require((int)vpMatches12.size() == N1 && (int)kv1.mvKeysUn.size() == N1 && (int)kv2.mvKeysUn.size() == N2,
                  "vpMatches12 / key-frame members differ in length");
  std::vector<uint8_t> vbAlreadyMatched1(N1, 0), vbAlreadyMatched2(N2, 0);
  for (int i = 0; i < N1; ++i) {
    MapPointT* pMP = vpMatches12[i];
    if (pMP) {
      vbAlreadyMatched1[i] = 1;
      const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
      if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = 1;
    }
  }
  LoopPointsView p1, p2;
  dropin::sim3_points(vpMapPoints1, vbAlreadyMatched1, Tw[0], T21, pKF2, &p1);
  dropin::sim3_points(vpMapPoints2, vbAlreadyMatched2, Tw[1], T12, pKF1, &p2);
  std::vector<int> vnMatch1, vnMatch2;
  Sim3Search(kv2, Tw[0], T21, p1, th, vnMatch1);
  Sim3Search(kv1, Tw[1], T12, p2, th, vnMatch2);
  int nFound = 0;
  for (int i1 = 0; i1 < N1; ++i1) {
    const int idx2 = vnMatch1[i1];
    if (idx2 >= 0 && vnMatch2[idx2] == i1) {
      vpMatches12[i1] = vpMapPoints2[idx2];
      nFound++;
    }
  }
  return nFound;
}

// ORBmatcher.h:104 (@0x80150)
template <class KeyFrameT, class FrameT, class MapPointT>
int ORBmatcher::SearchByBoW(KeyFrameT* pKF, FrameT& F, std::vector<MapPointT*>& vpMapPointMatches) {
  const std::vector<MapPointT*> vpMapPointsKF = pKF->GetMapPointMatches();
  vpMapPointMatches = std::vector<MapPointT*>(F.N, static_cast<MapPointT*>(nullptr));
  FrameView kf, f;
  kf.mvKeysUn = pKF->mvKeysUn;
  kf.mDescriptors = pKF->mDescriptors;
  const size_t n1 = kf.mvKeysUn.size();
  dropin::require(vpMapPointsKF.size() == n1 && (size_t)pKF->mDescriptors.rows == n1, "KeyFrame members differ in length");
  kf.hasMapPoint.assign(n1, 0);
  for (size_t i = 0; i < n1; ++i) kf.hasMapPoint[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();
  dropin::flatten_featvec(pKF->mFeatVec, &kf.mFeatVec);
  f.mvKeys = F.mvKeys;
  f.mDescriptors = F.mDescriptors;
  dropin::require((size_t)F.mDescriptors.rows == F.mvKeys.size(), "Frame members differ in length");
  dropin::flatten_featvec(F.mFeatVec, &f.mFeatVec);
  std::vector<int> match;
  const int nm = SearchByBoW(kf, f, match);
  for (size_t i = 0; i < match.size(); ++i)
    if (match[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[match[i]];
  return nm;
}

// ORBmatcher.h:105 (@0x82cc0)
template <class KeyFrameT, class MapPointT>
int ORBmatcher::SearchByBoW(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12) {
  FrameView v[2];
  KeyFrameT* kfs[2] = {pKF1, pKF2};
  std::vector<MapPointT*> mps[2];
  for (int s = 0; s < 2; ++s) {
    mps[s] = kfs[s]->GetMapPointMatches();
    v[s].mvKeysUn = kfs[s]->mvKeysUn;
    v[s].mDescriptors = kfs[s]->mDescriptors;
    const size_t n = v[s].mvKeysUn.size();
    dropin::require(mps[s].size() == n && (size_t)kfs[s]->mDescriptors.rows == n, "KeyFrame members differ in length");
    v[s].hasMapPoint.assign(n, 0);
    for (size_t i = 0; i < n; ++i) v[s].hasMapPoint[i] = mps[s][i] && !mps[s][i]->isBad();
    dropin::flatten_featvec(kfs[s]->mFeatVec, &v[s].mFeatVec);
  }
  vpMatches12 = std::vector<MapPointT*>(mps[0].size(), static_cast<MapPointT*>(nullptr));
  std::vector<int> m12;
  const int nm = SearchByBoW(v[0], v[1], m12, 0);
  for (size_t i = 0; i < m12.size(); ++i)
    if (m12[i] >= 0) vpMatches12[i] = mps[1][m12[i]];
  return nm;
}

// ORBmatcher.h:108 (@0x7db00)
template <class FrameT>
int ORBmatcher::SearchForInitialization(FrameT& F1, FrameT& F2, std::vector<cv::Point2f>& vbPrevMatched,
                                        std::vector<int>& vnMatches12, int windowSize) {
  FrameView f1, f2;
  f1.mvKeysUn = F1.mvKeysUn;
  f1.mDescriptors = F1.mDescriptors;
  dropin::require((size_t)F1.mDescriptors.rows == F1.mvKeysUn.size(), "F1 members differ in length");
  dropin::view_of_frame(F2, &f2, true);
  return SearchForInitialization(f1, f2, vbPrevMatched, vnMatches12, windowSize);
}

// ORBmatcher.h:111 (@0x86b30)
template <class KeyFrameT>
int ORBmatcher::SearchForTriangulation(KeyFrameT* pKF1, KeyFrameT* pKF2, cv::Mat F12,
                                       std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo) {
  FrameView v[2];
  KeyFrameT* kfs[2] = {pKF1, pKF2};
  for (int s = 0; s < 2; ++s) {
    KeyFrameT* kf = kfs[s];
    v[s].mvKeysUn = kf->mvKeysUn;
    v[s].mvuRight = kf->mvuRight;
    v[s].mDescriptors = kf->mDescriptors;
    const size_t n = v[s].mvKeysUn.size();
    dropin::require(kf->mvuRight.size() == n && (size_t)kf->mDescriptors.rows == n, "KeyFrame members differ in length");
    v[s].hasMapPoint.assign(n, 0);
    for (size_t i = 0; i < n; ++i) v[s].hasMapPoint[i] = kf->GetMapPoint(i) != nullptr;
    dropin::flatten_featvec(kf->mFeatVec, &v[s].mFeatVec);
  }
  v[1].mvScaleFactors = pKF2->mvScaleFactors;
  v[1].mvLevelSigma2 = pKF2->mvLevelSigma2;
  // epipole in the second image: C2 = R2w * Cw + t2w
  const cv::Mat Cw = pKF1->GetCameraCenter(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
  float R[9], t[3], C[3], F[9], ex, ey;
  for (int r = 0; r < 3; ++r) {
    t[r] = t2w.template at<float>(r);
    C[r] = Cw.template at<float>(r);
    for (int c = 0; c < 3; ++c) {
      R[3 * r + c] = R2w.template at<float>(r, c);
      F[3 * r + c] = F12.template at<float>(r, c);
    }
  }
  Epipole(R, t, C, pKF2->fx, pKF2->fy, pKF2->cx, pKF2->cy, &ex, &ey);
  std::vector<int> m12;
  const int nm = SearchForTriangulation(v[0], v[1], F, ex, ey, bOnlyStereo, m12);
  vMatchedPairs.clear();
  vMatchedPairs.reserve(nm);
  for (size_t i = 0; i < m12.size(); ++i)
    if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)m12[i]));
  return nm;
}

}  // namespace ORB_SLAM2
