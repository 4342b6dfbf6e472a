// TumIO.h — the two on-disk formats of the RGB-D example, with the reference's signatures, over the C-ABI
// (include/plslam_b200.h, last section).  Header only.
//   LoadImages         Examples/RGB-D/rgbd_tum.cc:22-23,151-176 (a free function of the example program)
//   SaveTrajectoryTUM  the file System::SaveTrajectoryTUM writes (include/System.h:104), for per-frame camera poses Tcw the
//                      caller has already chained (the reference walks mlRelativeFramePoses / reference key frames to get them)
#ifndef PLSLAM_HOST_TUMIO_H
#define PLSLAM_HOST_TUMIO_H
#include <stdexcept>
#include <string>
#include <vector>

#include "plslam_b200.h"

inline void LoadImages(const std::string& strAssociationFilename, std::vector<std::string>& vstrImageFilenamesRGB,
                       std::vector<std::string>& vstrImageFilenamesD, std::vector<double>& vTimestamps) {
  const int kStride = 512;
  int n = 0;
  if (plslam_tum_load_associations(strAssociationFilename.c_str(), nullptr, nullptr, nullptr, kStride, 0, &n) != PLSLAM_OK)
    throw std::runtime_error(plslam_last_error());
  std::vector<double> t(n);
  std::vector<char> rgb((size_t)n * kStride + 1), depth((size_t)n * kStride + 1);
  if (n && plslam_tum_load_associations(strAssociationFilename.c_str(), t.data(), rgb.data(), depth.data(), kStride, n, &n) != PLSLAM_OK)
    throw std::runtime_error(plslam_last_error());
  for (int i = 0; i < n; ++i) {  // appended, as the reference's push_back does
    vTimestamps.push_back(t[i]);
    vstrImageFilenamesRGB.emplace_back(rgb.data() + (size_t)i * kStride);
    vstrImageFilenamesD.emplace_back(depth.data() + (size_t)i * kStride);
  }
}

namespace ORB_SLAM2 {
struct TrajectoryPose {
  double timestamp;
  float Tcw[12];  // rows 0..2 of the 4x4 CV_32F camera pose, row-major
};
inline void SaveTrajectoryTUM(const std::string& filename, const std::vector<TrajectoryPose>& poses) {
  std::vector<double> t(poses.size());
  std::vector<float> T(poses.size() * 12);
  for (size_t i = 0; i < poses.size(); ++i) {
    t[i] = poses[i].timestamp;
    for (int k = 0; k < 12; ++k) T[i * 12 + k] = poses[i].Tcw[k];
  }
  if (plslam_tum_save_trajectory(filename.c_str(), t.data(), T.data(), (int)poses.size()) != PLSLAM_OK)
    throw std::runtime_error(plslam_last_error());
}
}  // namespace ORB_SLAM2
#endif
