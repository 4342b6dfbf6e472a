// FrameView.h — the slice of ORB_SLAM2::Frame / KeyFrame state the Hamming matchers read, as plain arrays.
// Frame, KeyFrame and MapPoint themselves (pointer graphs guarded by mutexes) are outside the hot path
// (SURVEY.md section 2, rows 6/11); INTEGRATION.md shows the ~20 lines that fill these views from the real classes.
#pragma once
#include <cstdint>
#include <vector>

#include "cv_compat.h"

namespace ORB_SLAM2 {

struct FeatureVectorView {         // DBoW2::FeatureVector = std::map<NodeId, std::vector<unsigned>> flattened
  std::vector<int32_t> nodes;      // keys, ascending
  std::vector<int32_t> start;      // nodes.size() + 1 offsets into idx
  std::vector<int32_t> idx;        // feature indices, in the vectors' order
};

// The fields of the local map points that ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) reads
// (reference include/MapPoint.h: mTrackProjX/Y/XR, mbTrackInView, mnTrackScaleLevel, mTrackViewCos, set by
// Frame::isInFrustum for Tracking::SearchLocalPoints), flattened in the order of vpMapPoints.
struct MapPointsView {
  std::vector<uint8_t> inViewAndGood;   // pMP->mbTrackInView && !pMP->isBad()
  std::vector<float> trackProj;         // M x 3: mTrackProjX, mTrackProjY, mTrackProjXR
  std::vector<int32_t> trackScaleLevel; // mnTrackScaleLevel
  std::vector<float> trackViewCos;      // mTrackViewCos
  cv::Mat descriptors;                  // M x 32: GetDescriptor()
  std::vector<uint8_t> observed;        // Observations() > 0
};

// The map points of a key frame as ORBmatcher::SearchByProjection(Frame&, KeyFrame*, set<MapPoint*>&, th, ORBdist) reads them
// (reference include/ORBmatcher.h:82), in the order of pKF->GetMapPointMatches().
struct KeyFramePointsView {
  std::vector<uint8_t> valid;      // pMP && !pMP->isBad() && !sAlreadyFound.count(pMP) && dist3D inside the invariance range
  std::vector<float> worldPos;     // M x 3: GetWorldPos()
  cv::Mat descriptors;             // M x 32: GetDescriptor()
  std::vector<float> angle;        // pKF->mvKeysUn[i].angle
  std::vector<int32_t> level;      // pMP->PredictScale(dist3D, &CurrentFrame)
};

// Key frame and map points as ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th) reads them
// (reference include/ORBmatcher.h:86).
struct KeyFrameGridView {
  std::vector<cv::KeyPoint> mvKeysUn;
  cv::Mat mDescriptors;                        // N x 32
  std::vector<float> mvScaleFactors;
  std::vector<int32_t> gridStart, gridItems;   // KeyFrame::mGrid[mnGridCols][mnGridRows] as CSR in [ix][iy] order
  int mnGridCols = 64, mnGridRows = 48;
  float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
  int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
  float fx = 0, fy = 0, cx = 0, cy = 0;
  // read by ORBmatcher::Fuse(pKF, vpMapPoints, th) on top (reprojection-error test): empty otherwise
  std::vector<float> mvuRight, mvInvLevelSigma2;
  float mbf = 0;
};
struct LoopPointsView {
  std::vector<uint8_t> valid;      // !isBad() && not in vpMatched on entry && inside the invariance range && PO.dot(Pn) >= 0.5 * dist
  std::vector<float> worldPos;     // M x 3
  cv::Mat descriptors;             // M x 32
  std::vector<int32_t> level;      // pMP->PredictScale(dist, pKF)
};

struct FrameView {
  // features
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
  cv::Mat mDescriptors;                        // N x 32
  std::vector<float> mvuRight;
  std::vector<float> mvScaleFactors, mvLevelSigma2;
  FeatureVectorView mFeatVec;
  // per-feature map point state
  std::vector<uint8_t> hasMapPoint;            // mvpMapPoints[i] != NULL (KeyFrame: && !isBad())
  std::vector<uint8_t> mapPointObserved;       // mvpMapPoints[i] && Observations() > 0
  std::vector<uint8_t> mvbOutlier;
  std::vector<float> mapPointWorldPos;         // N x 3, GetWorldPos()
  cv::Mat mapPointDescriptor;                  // N x 32, pMP->GetDescriptor()
  // grid (Frame::mGrid[64][48]) as CSR in [ix][iy] order, and bounds
  std::vector<int32_t> gridStart, gridItems;
  float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0, mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
  // camera
  float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mb = 0;
  float mTcw[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};  // rows 0..2 of the 4x4 pose
};

}  // namespace ORB_SLAM2
