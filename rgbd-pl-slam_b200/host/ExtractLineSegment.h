// ExtractLineSegment.h — drop-in for the reference's include/ExtractLineSegment.h:30-55 (class ORB_SLAM2::LineSegment).
#ifndef PLSLAM_EXTRACTLINESEGMENT_H
#define PLSLAM_EXTRACTLINESEGMENT_H

#include <vector>

#include "auxiliar.h"
#include "cv_compat.h"

struct plslam_lines;

namespace ORB_SLAM2 {

using cv::line_descriptor::KeyLine;
using Eigen::Vector3d;

class LineSegment {
 public:
  LineSegment();
  ~LineSegment();

  // LSD detection + LBD description + line equations (reference ExtractLineSegment.h:38; the header defaults
  // `int scale = 1.2` (truncated to 1) and numOctaves = 1 mean one full-resolution octave, the only mode implemented)
  void ExtractLineSegment(const cv::Mat& img, std::vector<KeyLine>& vkeyLines, cv::Mat& ldesc,
                          std::vector<Vector3d>& vkeylineFunctions, int scale = 1, int numOctaves = 1);

  // kNN (k = 2) Hamming matching of two LBD descriptor blocks into mvlineMatches (ExtractLineSegment.h:41)
  void LineSegmentMathch(cv::Mat& ldesc1, cv::Mat& ldesc2);

  // median-absolute-deviation statistics of the last LineSegmentMathch (ExtractLineSegment.h:44; auxiliar.h:30-51)
  void LineDescriptorMAD();

  // overlap ratio between an observed and a projected segment along the line (ExtractLineSegment.h:47)
  double LineSegmentOverlap(double spl_obs, double epl_obs, double spl_proj, double epl_proj);

  void SetMaxLines(int n);  // lsdNFeatures (default 40)
  const std::vector<std::vector<cv::DMatch> >& Matches() const { return mvlineMatches; }
  double NNMad() const { return mnnMad; }
  double NN12Mad() const { return mnn12Mad; }

 protected:
  std::vector<std::vector<cv::DMatch> > mvlineMatches;
  double mnnMad = 0, mnn12Mad = 0;

 private:
  plslam_lines* mpImpl = nullptr;
};

}  // namespace ORB_SLAM2
#endif
