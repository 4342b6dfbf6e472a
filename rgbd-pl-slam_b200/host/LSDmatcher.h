// LSDmatcher.h — drop-in surface of the reference's include/LSDmatcher.h:25-78 for what is recoverable of it:
// the reference ships this class as a header only (no source, no machine code), and the header evidence
// (include/auxiliar.h:30-51, include/ExtractLineSegment.h:41-44) shows brute-force Hamming kNN (k = 2) on LBD rows
// with MAD-based thresholds.  The projection-window overloads (LSDmatcher.h:32-38) depend on Frame/MapLine geometry
// that cannot be recovered; they are listed as out of scope in DESIGN.md.
#ifndef PLSLAM_LSDMATCHER_H
#define PLSLAM_LSDMATCHER_H

#include <vector>

#include "cv_compat.h"

namespace ORB_SLAM2 {

class LSDmatcher {
 public:
  LSDmatcher(float nnratio = 0.6, bool checkOri = true);
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
  // knnMatch(ldesc1, ldesc2, k = 2) + nearest/second-nearest ratio test; vnMatches12[i] = row of ldesc2 or -1
  int MatchKNN(const cv::Mat& ldesc1, const cv::Mat& ldesc2, std::vector<int>& vnMatches12,
               std::vector<std::vector<cv::DMatch> >* knn = nullptr);

 public:
  static const int TH_LOW;
  static const int TH_HIGH;
  static const int HISTO_LENGTH;

 protected:
  float mfNNratio;
  bool mbCheckOrientation;
};

}  // namespace ORB_SLAM2
#endif
