// ORBextractor.h — drop-in for the reference's include/ORBextractor.h:45-111 (class ORB_SLAM2::ORBextractor):
// same constructor, operator(), getters and public mvImagePyramid; the implementation calls the
// B200 C-ABI (include/plslam_b200.h).  There is no CPU fallback: failures throw std::runtime_error.
#ifndef PLSLAM_ORBEXTRACTOR_H
#define PLSLAM_ORBEXTRACTOR_H

#include <vector>

#include "cv_compat.h"

struct plslam_orb;

namespace ORB_SLAM2 {

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
  ~ORBextractor();

  // Compute the ORB features and descriptors on an image.  Mask is ignored (as in the reference).
  void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
                  cv::OutputArray descriptors);

  int inline GetLevels() { return nlevels; }
  float inline GetScaleFactor() { return (float)scaleFactor; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  // Public member of the reference (ORBextractor.h:85).  Only the stock stereo matcher reads it; filling it
  // costs a device->host copy of the pyramid per frame, so it is populated only when enabled here.
  std::vector<cv::Mat> mvImagePyramid;
  void SetPopulatePyramid(bool on) { mbPopulatePyramid = on; }

  // 7-tap table of the 7x7 sigma=2 blur (OpenCV-version dependent, see include/plslam_b200.h)
  void SetBlurKernel(const int k[7]);

 protected:
  int nfeatures;
  double scaleFactor;
  int nlevels;
  int iniThFAST;
  int minThFAST;
  std::vector<int> mnFeaturesPerLevel;
  std::vector<int> umax;
  std::vector<float> mvScaleFactor;
  std::vector<float> mvInvScaleFactor;
  std::vector<float> mvLevelSigma2;
  std::vector<float> mvInvLevelSigma2;

 private:
  plslam_orb* mpImpl = nullptr;
  bool mbPopulatePyramid = false;
  std::vector<cv::KeyPoint> mvScratch;
};

}  // namespace ORB_SLAM2
#endif
