// ORBextractor.cc — ORB_SLAM2::ORBextractor over the C-ABI (reference interface include/ORBextractor.h:45-111).
#include "ORBextractor.h"

#include <stdexcept>
#include <string>

#include "../../include/plslam_b200.h"

namespace ORB_SLAM2 {

static void check(int rc, const char* what) {
  if (rc != PLSLAM_OK) throw std::runtime_error(std::string(what) + ": " + plslam_last_error());
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
  check(plslam_orb_create(&mpImpl, nfeatures, _scaleFactor, nlevels, iniThFAST, minThFAST), "ORBextractor");
  mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
  mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
  mnFeaturesPerLevel.resize(nlevels); umax.resize(16);
  check(plslam_orb_tables(mpImpl, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                          mvInvLevelSigma2.data(), mnFeaturesPerLevel.data(), umax.data()), "ORBextractor tables");
  mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() { plslam_orb_destroy(mpImpl); }

void ORBextractor::SetBlurKernel(const int k[7]) { check(plslam_orb_set_blur_kernel(mpImpl, k), "SetBlurKernel"); }

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors) {
  if (_image.empty()) return;  // silent return, lib/libORB_SLAM2.so@0x76dda
  cv::Mat image = _image.getMat();
  if (image.type() != CV_8UC1) throw std::runtime_error("ORBextractor: image must be CV_8UC1");
  const int cap = plslam_orb_max_keypoints(mpImpl);
  static_assert(sizeof(cv::KeyPoint) == sizeof(plslam_keypoint_t), "KeyPoint layout");
  mvScratch.resize(cap);
  cv::Mat all(cap, 32, CV_8U);
  int n = 0;
  check(plslam_orb_extract(mpImpl, image.data, image.cols, image.rows, (int)image.step,
                           reinterpret_cast<plslam_keypoint_t*>(mvScratch.data()), all.data, cap, &n), "ORBextractor()");
  _keypoints.assign(mvScratch.begin(), mvScratch.begin() + n);
  if (n == 0) {
    _descriptors.release();  // @0x78001
  } else {
    _descriptors.create(n, 32, CV_8U);
    cv::Mat d = _descriptors.getMat();
    for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), all.ptr(i), 32);
  }
  if (mbPopulatePyramid) {
    for (int l = 0; l < nlevels; ++l) {
      int w = 0, h = 0;
      check(plslam_orb_level_size(mpImpl, l, &w, &h), "level size");
      mvImagePyramid[l].create(h, w, CV_8U);
      check(plslam_orb_copy_level(mpImpl, 0, l, 0, mvImagePyramid[l].data, (size_t)w * h), "copy level");
    }
  }
}

}  // namespace ORB_SLAM2
