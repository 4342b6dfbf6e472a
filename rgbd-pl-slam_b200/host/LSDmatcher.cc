// LSDmatcher.cc — ORB_SLAM2::LSDmatcher kNN core over the C-ABI (reference include/LSDmatcher.h:25-78; header-only there).
#include "LSDmatcher.h"

#include <stdexcept>
#include <string>

#include "../../include/plslam_b200.h"

namespace ORB_SLAM2 {

// The header declares the constants without values (LSDmatcher.h:61-63); ORBmatcher's are used.
const int LSDmatcher::TH_HIGH = PLSLAM_TH_HIGH;
const int LSDmatcher::TH_LOW = PLSLAM_TH_LOW;
const int LSDmatcher::HISTO_LENGTH = PLSLAM_HISTO_LENGTH;

LSDmatcher::LSDmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

int LSDmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return plslam_descriptor_distance(a.data, b.data); }

int LSDmatcher::MatchKNN(const cv::Mat& ldesc1, const cv::Mat& ldesc2, std::vector<int>& vnMatches12,
                         std::vector<std::vector<cv::DMatch> >* knn) {
  const int nq = ldesc1.rows, nt = ldesc2.rows;
  vnMatches12.assign(nq, -1);
  if (knn) knn->assign(nq, std::vector<cv::DMatch>());
  if (nq == 0) return 0;
  cv::Mat q = ldesc1.step == 32 ? ldesc1 : ldesc1.clone(), t = (nt == 0 || ldesc2.step == 32) ? ldesc2 : ldesc2.clone();
  std::vector<int32_t> out((size_t)nq * 4);
  if (plslam_match_knn2_host(q.data, nq, t.data, nt, out.data()) != PLSLAM_OK)
    throw std::runtime_error(std::string("LSDmatcher::MatchKNN: ") + plslam_last_error());
  int n = 0;
  for (int i = 0; i < nq; ++i) {
    const int i1 = out[4 * i], d1 = out[4 * i + 1], i2 = out[4 * i + 2], d2 = out[4 * i + 3];
    if (knn) {
      if (i1 >= 0) (*knn)[i].push_back(cv::DMatch(i, i1, (float)d1));
      if (i2 >= 0) (*knn)[i].push_back(cv::DMatch(i, i2, (float)d2));
    }
    if (i1 >= 0 && d1 <= TH_HIGH && (i2 < 0 || (float)d1 < mfNNratio * (float)d2)) {
      vnMatches12[i] = i1;
      ++n;
    }
  }
  return n;
}

}  // namespace ORB_SLAM2
