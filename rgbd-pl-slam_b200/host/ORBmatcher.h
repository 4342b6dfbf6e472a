// ORBmatcher.h — drop-in surface of the reference's include/ORBmatcher.h:37-141 for the Hamming cores on the hot
// path: DescriptorDistance (:44), SearchByProjection(Frame&, const Frame&, th, bMono) (:78), SearchByProjection(Frame&, const vector<MapPoint*>&, th) (:61) and
// SearchByBoW(KeyFrame*, Frame&, matches) (:104).  Frames are passed as FrameView (see FrameView.h).
#ifndef PLSLAM_ORBMATCHER_H
#define PLSLAM_ORBMATCHER_H

#include <vector>

#include "FrameView.h"
#include "cv_compat.h"

namespace ORB_SLAM2 {

class ORBmatcher {
 public:
  ORBmatcher(float nnratio = 0.6, bool checkOri = true);

  // Computes the Hamming distance between two ORB descriptors
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);

  // Project MapPoints tracked in last frame into the current frame and search matches (Tracking).
  // vnMatches[i2] = index in LastFrame matched to current keypoint i2, or -1.  Returns the number of matches.
  int SearchByProjection(FrameView& CurrentFrame, const FrameView& LastFrame, const float th, const bool bMono,
                         std::vector<int>& vnMatches);

  // Search matches between Frame keypoints and projected MapPoints (Tracking::SearchLocalPoints; ORBmatcher.h:61).
  // vnMatches[i] = index in vpMapPoints assigned to F's keypoint i, or -1.  Returns the number of matches.
  int SearchByProjection(FrameView& F, const MapPointsView& vpMapPoints, const float th, std::vector<int>& vnMatches);

  // Brute force constrained to ORB that belong to the same vocabulary node (Relocalisation / TrackReferenceKeyFrame).
  // vnMatches[iF] = index in the KeyFrame matched to F's feature iF, or -1.
  int SearchByBoW(const FrameView& KF, FrameView& F, std::vector<int>& vnMatches);

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
#endif
