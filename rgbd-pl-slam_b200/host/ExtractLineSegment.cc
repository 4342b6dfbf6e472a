// ExtractLineSegment.cc — ORB_SLAM2::LineSegment over the C-ABI (reference interface include/ExtractLineSegment.h:30-55).
#include "ExtractLineSegment.h"

#include <stdexcept>
#include <string>

#include "../../include/plslam_b200.h"

namespace ORB_SLAM2 {

static void check(int rc, const char* what) {
  if (rc != PLSLAM_OK) throw std::runtime_error(std::string(what) + ": " + plslam_last_error());
}

LineSegment::LineSegment() { check(plslam_lines_create(&mpImpl), "LineSegment"); }
LineSegment::~LineSegment() { plslam_lines_destroy(mpImpl); }
void LineSegment::SetMaxLines(int n) { check(plslam_lines_set_max_lines(mpImpl, n), "SetMaxLines"); }

void LineSegment::ExtractLineSegment(const cv::Mat& img, std::vector<KeyLine>& vkeyLines, cv::Mat& ldesc,
                                     std::vector<Vector3d>& vkeylineFunctions, int scale, int numOctaves) {
  if (scale != 1 || numOctaves != 1) throw std::runtime_error("LineSegment: only scale = 1, numOctaves = 1 (the header defaults)");
  vkeyLines.clear();
  vkeylineFunctions.clear();
  if (img.empty()) { ldesc.release(); return; }
  const int cap = plslam_lines_capacity(mpImpl);
  static_assert(sizeof(KeyLine) == sizeof(plslam_keyline_t), "KeyLine layout");
  std::vector<KeyLine> kl(cap);
  std::vector<double> fn((size_t)cap * 3);
  cv::Mat all(cap, 32, CV_8U);
  int n = 0;
  check(plslam_lines_extract(mpImpl, img.data, img.cols, img.rows, (int)img.step, reinterpret_cast<plslam_keyline_t*>(kl.data()),
                             all.data, fn.data(), cap, &n), "ExtractLineSegment");
  vkeyLines.assign(kl.begin(), kl.begin() + n);
  if (n == 0) ldesc.release();
  else {
    ldesc.create(n, 32, CV_8U);
    for (int i = 0; i < n; ++i) std::memcpy(ldesc.ptr(i), all.ptr(i), 32);
  }
  vkeylineFunctions.resize(n);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) vkeylineFunctions[i](k) = fn[3 * (size_t)i + k];
}

void LineSegment::LineSegmentMathch(cv::Mat& ldesc1, cv::Mat& ldesc2) {
  mvlineMatches.clear();
  const int nq = ldesc1.rows, nt = ldesc2.rows;
  if (nq == 0) return;
  cv::Mat q = ldesc1.step == 32 ? ldesc1 : ldesc1.clone(), t = (nt == 0 || ldesc2.step == 32) ? ldesc2 : ldesc2.clone();
  std::vector<int32_t> out((size_t)nq * 4);
  check(plslam_match_knn2_host(q.data, nq, t.data, nt, out.data()), "LineSegmentMathch");
  mvlineMatches.resize(nq);
  for (int i = 0; i < nq; ++i) {
    if (out[4 * i] >= 0) mvlineMatches[i].push_back(cv::DMatch(i, out[4 * i], (float)out[4 * i + 1]));
    if (out[4 * i + 2] >= 0) mvlineMatches[i].push_back(cv::DMatch(i, out[4 * i + 2], (float)out[4 * i + 3]));
  }
}

void LineSegment::LineDescriptorMAD() {
  std::vector<double> nn, nn12;
  for (const auto& m : mvlineMatches)
    if (m.size() >= 2) {
      nn.push_back(m[0].distance);
      nn12.push_back(m[1].distance - m[0].distance);
    }
  mnnMad = vector_mad(nn);
  mnn12Mad = vector_mad(nn12);
}

double LineSegment::LineSegmentOverlap(double spl_obs, double epl_obs, double spl_proj, double epl_proj) {
  const double sln = std::min(spl_obs, epl_obs), eln = std::max(spl_obs, epl_obs);
  const double spn = std::min(spl_proj, epl_proj), epn = std::max(spl_proj, epl_proj);
  const double length = eln - spn;
  double overlap;
  if (epn < sln || spn > eln) overlap = 0.0;
  else if (epn > eln && spn < sln) overlap = eln - sln;
  else overlap = std::min(eln, epn) - std::max(sln, spn);
  return length > 0.01 ? overlap / length : 0.0;
}

}  // namespace ORB_SLAM2
