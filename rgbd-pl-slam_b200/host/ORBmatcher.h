// ORBmatcher.h — drop-in surface of the reference's include/ORBmatcher.h:37-141 for the Hamming cores on the hot
// path: DescriptorDistance (:44), SearchByProjection(Frame&, const Frame&, th, bMono) (:78), SearchByProjection(Frame&, const
// vector<MapPoint*>&, th) (:61), SearchByBoW(KeyFrame*, Frame&, matches) (:104) and SearchForTriangulation (:111).
// Two forms of each: the reference's EXACT signatures as member templates over the caller's Frame / KeyFrame / MapPoint
// classes (with the reference's own headers included, Tracking / LocalMapping call them unchanged; definitions in
// ORBmatcher_impl.h), and the flattened FrameView forms they are built on (see FrameView.h).
#ifndef PLSLAM_ORBMATCHER_H
#define PLSLAM_ORBMATCHER_H

#include <cstddef>
#include <utility>
#include <set>
#include <vector>

#include "FrameView.h"
#include "cv_compat.h"

namespace ORB_SLAM2 {

class ORBmatcher {
 public:
  ORBmatcher(float nnratio = 0.6, bool checkOri = true);

  // Computes the Hamming distance between two ORB descriptors
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);

  // ---- the reference's signatures (ORBmatcher.h:61, :78, :104, :111).  FrameT / KeyFrameT / MapPointT are the
  //      caller's ORB_SLAM2::Frame / KeyFrame / MapPoint: the members read and written are the ones the reference's
  //      functions read and write (Frame: N, mvKeys, mvKeysUn, mvuRight, mDescriptors, mvpMapPoints, mvbOutlier, mGrid,
  //      fx fy cx cy, mnMinX.., mfGridElement*Inv, mbf, mb, mTcw, mvScaleFactors, mFeatVec; MapPoint: isBad, Observations,
  //      GetWorldPos, GetDescriptor, mbTrackInView, mTrackProjX/Y/XR, mnTrackScaleLevel, mTrackViewCos; KeyFrame:
  //      GetMapPointMatches, GetMapPoint, GetRotation, GetTranslation, GetCameraCenter, fx fy cx cy, mvKeysUn, mvuRight,
  //      mDescriptors, mFeatVec, mvScaleFactors, mvLevelSigma2) ----
  // Search matches between Frame keypoints and projected MapPoints. Returns number of matches (Tracking::SearchLocalPoints)
  template <class FrameT, class MapPointT>
  int SearchByProjection(FrameT& F, const std::vector<MapPointT*>& vpMapPoints, const float th = 3);
  // Project MapPoints tracked in last frame into the current frame and search matches (Tracking::TrackWithMotionModel)
  template <class FrameT>
  int SearchByProjection(FrameT& CurrentFrame, const FrameT& LastFrame, const float th, const bool bMono);
  // Project MapPoints seen in KeyFrame into the Frame and search matches (Tracking::Relocalization) (ORBmatcher.h:82).
  // Further members read: Frame::mfLogScaleFactor / mnScaleLevels through pMP->PredictScale(dist, &CurrentFrame),
  // MapPoint::GetMinDistanceInvariance / GetMaxDistanceInvariance.
  template <class FrameT, class KeyFrameT, class MapPointT>
  int SearchByProjection(FrameT& CurrentFrame, KeyFrameT* pKF, const std::set<MapPointT*>& sAlreadyFound, const float th,
                         const int ORBdist);
  // Project MapPoints using a Similarity Transformation and search matches (LoopClosing) (ORBmatcher.h:86).  pKF->mGrid is
  // protected in the reference's KeyFrame.h: the template reads it through `pKF->mGrid`, i.e. it needs ORBmatcher befriended
  // there or a public accessor.
  template <class KeyFrameT, class MapPointT>
  int SearchByProjection(KeyFrameT* pKF, cv::Mat Scw, const std::vector<MapPointT*>& vpPoints, std::vector<MapPointT*>& vpMatched,
                         int th);
  // Project MapPoints into KeyFrame and search for duplicated MapPoints (LocalMapping::SearchInNeighbors) (ORBmatcher.h:119), and
  // the variant with a Similarity Transformation (LoopClosing::SearchAndFuse) (ORBmatcher.h:122).  The matching of all candidates
  // is ONE kernel (the points are independent); the map-graph bookkeeping that follows each match — pMPinKF->Replace(pMP) /
  // pMP->Replace(pMPinKF) / AddObservation + AddMapPoint, or vpReplacePoint[i] = pMPinKF — is replayed here in the reference's
  // order on the caller's objects.  Further members read: pKF->mvuRight, mvInvLevelSigma2, mbf, GetRotation / GetTranslation /
  // GetCameraCenter, GetMapPoint, GetMapPoints; pMP->IsInKeyFrame, Observations.
  template <class KeyFrameT, class MapPointT>
  int Fuse(KeyFrameT* pKF, const std::vector<MapPointT*>& vpMapPoints, const float th = 3.0);
  template <class KeyFrameT, class MapPointT>
  int Fuse(KeyFrameT* pKF, cv::Mat Scw, const std::vector<MapPointT*>& vpPoints, float th, std::vector<MapPointT*>& vpReplacePoint);
  // Search matches between MapPoints seen in KF1 and KF2 transforming by a Sim3 [s12*R12|t12] (LoopClosing::ComputeSim3)
  // (ORBmatcher.h:116).  Both directions run on the device; the transforms, the invariance test, PredictScale and the
  // agreement pass are evaluated here.  Further members read: pMP->GetIndexInKeyFrame(pKF2).
  template <class KeyFrameT, class MapPointT>
  int SearchBySim3(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12, const float& s12, const cv::Mat& R12,
                   const cv::Mat& t12, const float th);
  // Search matches between MapPoints in a KeyFrame and ORB in a Frame (Relocalisation, TrackReferenceKeyFrame)
  template <class KeyFrameT, class FrameT, class MapPointT>
  int SearchByBoW(KeyFrameT* pKF, FrameT& F, std::vector<MapPointT*>& vpMapPointMatches);
  // Search matches between the MapPoints of two key frames by vocabulary node (LoopClosing) (ORBmatcher.h:105)
  template <class KeyFrameT, class MapPointT>
  int SearchByBoW(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12);
  // Matching for the Map Initialization (only used in the monocular case) (ORBmatcher.h:108)
  template <class FrameT>
  int SearchForInitialization(FrameT& F1, FrameT& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                              int windowSize = 10);
  // Matching to triangulate new MapPoints. Check Epipolar Constraint (LocalMapping::CreateNewMapPoints)
  template <class KeyFrameT>
  int SearchForTriangulation(KeyFrameT* pKF1, KeyFrameT* pKF2, cv::Mat F12,
                             std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo);

  // ---- flattened forms ----
  // Project MapPoints tracked in last frame into the current frame and search matches (Tracking).
  // vnMatches[i2] = index in LastFrame matched to current keypoint i2, or -1 (-2 with reportRemoved: assigned during the
  // scan and taken away again by the rotation-consistency check, where the reference leaves NULL).  Returns the number of matches.
  int SearchByProjection(FrameView& CurrentFrame, const FrameView& LastFrame, const float th, const bool bMono,
                         std::vector<int>& vnMatches, bool reportRemoved = false);

  // Search matches between Frame keypoints and projected MapPoints (Tracking::SearchLocalPoints; ORBmatcher.h:61).
  // vnMatches[i] = index in vpMapPoints assigned to F's keypoint i, or -1.  Returns the number of matches.
  int SearchByProjection(FrameView& F, const MapPointsView& vpMapPoints, const float th, std::vector<int>& vnMatches);

  // Key-frame form (ORBmatcher.h:82) on views: KF.valid = map point present, good and not already found; KF.level = the
  // search level of each point (PredictScale), points outside their scale-invariance range marked invalid; CurrentFrame needs
  // mvKeysUn, mDescriptors, hasMapPoint (mvpMapPoints[i] != NULL), grid, bounds, intrinsics, pose, mvScaleFactors.
  // vnMatches[i2] = key-frame index assigned to current keypoint i2, or -1.
  int SearchByProjection(FrameView& CurrentFrame, const KeyFramePointsView& KF, const float th, const int ORBdist,
                         std::vector<int>& vnMatches);

  // Loop-closing form (ORBmatcher.h:86) on views: Scw = rows 0..2 of the 4x4 similarity; matchedOnEntry[i] = vpMatched[i] != NULL;
  // vnMatches[i] = index into the points newly assigned to key-frame feature i, or -1.
  int SearchByProjection(const KeyFrameGridView& KF, const float Scw[12], const LoopPointsView& P,
                         const std::vector<uint8_t>& matchedOnEntry, int th, std::vector<int>& vnMatches);

  // Matching core of Fuse on views: pose = Rcw | tcw with ow = camera centre (useScw = false, KF needs mvuRight / mvInvLevelSigma2 /
  // mbf) or rows 0..2 of Scw (useScw = true).  vnBestIdx[i] = key-frame feature map point i would be fused into, or -1.
  void FuseSearch(const KeyFrameGridView& KF, const float pose[12], const float ow[3], bool useScw, const LoopPointsView& P, float th,
                  std::vector<int>& vnBestIdx);

  // One direction of SearchBySim3 on views: the points (world coordinates) go into their own key frame's camera with poseOwn
  // (R?w | t?w) and into the other camera with pose2 (sR | t); KFother is the key frame they are searched in.
  void Sim3Search(const KeyFrameGridView& KFother, const float poseOwn[12], const float pose2[12], const LoopPointsView& P, float th,
                  std::vector<int>& vnBestIdx);

  // Brute force constrained to ORB that belong to the same vocabulary node (Relocalisation / TrackReferenceKeyFrame).
  // vnMatches[iF] = index in the KeyFrame matched to F's feature iF, or -1.
  int SearchByBoW(const FrameView& KF, FrameView& F, std::vector<int>& vnMatches);

  // Key frame / key frame form (ORBmatcher.h:105): both views need mvKeysUn, mDescriptors, mFeatVec and hasMapPoint
  // (map point present and not bad).  vnMatches12[i1] = KF2 feature whose map point goes to KF1's feature i1, or -1.
  int SearchByBoW(const FrameView& KF1, const FrameView& KF2, std::vector<int>& vnMatches12, int /*tag: KF-KF*/);

  // SearchForInitialization on views: F1 needs mvKeysUn and mDescriptors, F2 also its grid and bounds
  int SearchForInitialization(const FrameView& F1, const FrameView& F2, std::vector<cv::Point2f>& vbPrevMatched,
                              std::vector<int>& vnMatches12, int windowSize);

  // SearchForTriangulation on views: KF1 / KF2 need mvKeysUn, mvuRight, mDescriptors, mFeatVec, hasMapPoint; KF2 also
  // mvScaleFactors and mvLevelSigma2; (ex, ey) = the epipole of KF1's centre in KF2 (Epipole below).
  // vnMatches12[i1] = KF2 feature matched to KF1 feature i1, or -1.
  int SearchForTriangulation(const FrameView& KF1, const FrameView& KF2, const float F12[9], float ex, float ey,
                             const bool bOnlyStereo, std::vector<int>& vnMatches12);
  // sR12 = s12 * R12, sR21 = (1.0 / s12) * R12.t(), t21 = -sR21 * t12 as the reference's cv::Mat expressions evaluate (@0x83a8e)
  static void Sim3Transforms(float s12, const float R12[9], const float t12[3], float sR12[9], float sR21[9], float t21[3]);
  // C2 = R2w * Cw + t2w projected with KF2's intrinsics, with the binary's rounding sequence (@0x86b9c-0x86f8b)
  static void Epipole(const float R2w[9], const float t2w[3], const float Cw[3], float fx, float fy, float cx, float cy,
                      float* ex, float* ey);

 public:
  static const int TH_LOW;
  static const int TH_HIGH;
  static const int HISTO_LENGTH;

 protected:
  float RadiusByViewingCos(const float& viewCos);
  float mfNNratio;
  bool mbCheckOrientation;
};

}  // namespace ORB_SLAM2

#include "ORBmatcher_impl.h"
#endif
