// ORBmatcher.cc — ORB_SLAM2::ORBmatcher Hamming cores over the C-ABI (reference include/ORBmatcher.h:37-141).
#include "ORBmatcher.h"

#include <stdexcept>
#include <string>

#include "../../include/plslam_b200.h"

namespace ORB_SLAM2 {

const int ORBmatcher::TH_HIGH = PLSLAM_TH_HIGH;            // lib/libORB_SLAM2.so@0x1269e8
const int ORBmatcher::TH_LOW = PLSLAM_TH_LOW;              // @0x1269e4
const int ORBmatcher::HISTO_LENGTH = PLSLAM_HISTO_LENGTH;  // @0x1269e0

static void check(int rc, const char* what) {
  if (rc != PLSLAM_OK) throw std::runtime_error(std::string(what) + ": " + plslam_last_error());
}

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return plslam_descriptor_distance(a.data, b.data); }

float ORBmatcher::RadiusByViewingCos(const float& viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }  // @0x79b60

int ORBmatcher::SearchByProjection(FrameView& Cur, const FrameView& Last, const float th, const bool bMono,
                                   std::vector<int>& vnMatches) {
  const int n1 = (int)Last.mvKeysUn.size(), n2 = (int)Cur.mvKeysUn.size();
  vnMatches.assign(n2, -1);
  if (n1 == 0 || n2 == 0) return 0;
  std::vector<uint8_t> valid(n1);
  std::vector<int32_t> loct(n1), coct(n2);
  std::vector<float> lang(n1), cang(n2), cxy((size_t)n2 * 2);
  for (int i = 0; i < n1; ++i) {
    valid[i] = Last.hasMapPoint[i] && !(Last.mvbOutlier.size() ? Last.mvbOutlier[i] : 0);
    loct[i] = Last.mvKeys[i].octave;
    lang[i] = Last.mvKeysUn[i].angle;
  }
  for (int i = 0; i < n2; ++i) {
    coct[i] = Cur.mvKeysUn[i].octave;
    cang[i] = Cur.mvKeysUn[i].angle;
    cxy[2 * i] = Cur.mvKeysUn[i].pt.x;
    cxy[2 * i + 1] = Cur.mvKeysUn[i].pt.y;
  }
  int32_t nm = 0;
  plslam_proj_job_t j{};
  j.last_valid = valid.data(); j.last_xyz = Last.mapPointWorldPos.data(); j.last_desc = Last.mapPointDescriptor.data;
  j.last_octave = loct.data(); j.last_angle = lang.data(); j.last_obs = Last.mapPointObserved.data();
  j.cur_xy = cxy.data(); j.cur_octave = coct.data(); j.cur_angle = cang.data(); j.cur_desc = Cur.mDescriptors.data;
  j.cur_uright = Cur.mvuRight.data(); j.cur_taken = Cur.mapPointObserved.data();
  j.grid_start = Cur.gridStart.data(); j.grid_items = Cur.gridItems.data(); j.scale_factors = Cur.mvScaleFactors.data();
  j.match_cur = vnMatches.data(); j.nmatches = &nm;
  const float cam[12] = {Cur.fx, Cur.fy, Cur.cx, Cur.cy, Cur.mbf, Cur.mb, Cur.mnMinX, Cur.mnMaxX, Cur.mnMinY, Cur.mnMaxY,
                         Cur.mfGridElementWidthInv, Cur.mfGridElementHeightInv};
  std::memcpy(j.cam, cam, sizeof(cam));
  std::memcpy(j.tcw_cur, Cur.mTcw, sizeof(j.tcw_cur));
  std::memcpy(j.tcw_last, Last.mTcw, sizeof(j.tcw_last));
  j.th = th; j.n1 = n1; j.n2 = n2; j.mono = bMono; j.check_orientation = mbCheckOrientation;
  check(plslam_match_projection_host(&j, (int)Cur.mvScaleFactors.size()), "SearchByProjection");
  return nm;
}

int ORBmatcher::SearchByProjection(FrameView& F, const MapPointsView& MP, const float th, std::vector<int>& vnMatches) {
  const int m = (int)MP.inViewAndGood.size(), n = (int)F.mvKeysUn.size();
  vnMatches.assign(n, -1);
  if (m == 0 || n == 0) return 0;
  std::vector<int32_t> oct(n);
  std::vector<float> xy((size_t)n * 2);
  for (int i = 0; i < n; ++i) {
    oct[i] = F.mvKeysUn[i].octave;
    xy[2 * i] = F.mvKeysUn[i].pt.x;
    xy[2 * i + 1] = F.mvKeysUn[i].pt.y;
  }
  int32_t nm = 0;
  plslam_local_job_t j{};
  j.mp_valid = MP.inViewAndGood.data(); j.mp_proj = MP.trackProj.data(); j.mp_level = MP.trackScaleLevel.data();
  j.mp_viewcos = MP.trackViewCos.data(); j.mp_desc = MP.descriptors.data; j.mp_obs = MP.observed.data();
  j.f_xy = xy.data(); j.f_octave = oct.data(); j.f_desc = F.mDescriptors.data; j.f_uright = F.mvuRight.data();
  j.f_taken = F.mapPointObserved.data(); j.grid_start = F.gridStart.data(); j.grid_items = F.gridItems.data();
  j.scale_factors = F.mvScaleFactors.data(); j.match_f = vnMatches.data(); j.nmatches = &nm;
  j.cam[0] = F.mnMinX; j.cam[1] = F.mnMinY; j.cam[2] = F.mfGridElementWidthInv; j.cam[3] = F.mfGridElementHeightInv;
  j.th = th; j.nnratio = mfNNratio; j.m = m; j.n = n;
  check(plslam_match_local_points_host(&j, (int)F.mvScaleFactors.size()), "SearchByProjection(local map)");
  return nm;
}

int ORBmatcher::SearchByBoW(const FrameView& KF, FrameView& F, std::vector<int>& vnMatches) {
  const int n1 = (int)KF.mvKeysUn.size(), n2 = (int)F.mvKeys.size();
  vnMatches.assign(n2, -1);
  if (n1 == 0 || n2 == 0) return 0;
  std::vector<float> a1(n1), a2(n2);
  for (int i = 0; i < n1; ++i) a1[i] = KF.mvKeysUn[i].angle;
  for (int i = 0; i < n2; ++i) a2[i] = F.mvKeys[i].angle;
  int32_t nm = 0;
  plslam_bow_job_t j{};
  j.kf_desc = KF.mDescriptors.data; j.kf_angle = a1.data(); j.kf_valid = KF.hasMapPoint.data();
  j.kf_nodes = KF.mFeatVec.nodes.data(); j.kf_start = KF.mFeatVec.start.data(); j.kf_idx = KF.mFeatVec.idx.data();
  j.f_desc = F.mDescriptors.data; j.f_angle = a2.data();
  j.f_nodes = F.mFeatVec.nodes.data(); j.f_start = F.mFeatVec.start.data(); j.f_idx = F.mFeatVec.idx.data();
  j.match_f = vnMatches.data(); j.nmatches = &nm;
  j.n1 = n1; j.n2 = n2; j.n_kf_nodes = (int)KF.mFeatVec.nodes.size(); j.n_f_nodes = (int)F.mFeatVec.nodes.size();
  j.nnratio = mfNNratio; j.check_orientation = mbCheckOrientation;
  check(plslam_match_bow_host(&j), "SearchByBoW");
  return nm;
}

}  // namespace ORB_SLAM2
