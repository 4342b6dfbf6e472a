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

// every array of a view must cover its features: a short or unfilled vector would be read out of bounds by the upload
static void need(bool ok, const char* what) {
  if (!ok) throw std::invalid_argument(std::string("ORBmatcher: ") + what);
}

int ORBmatcher::SearchByProjection(FrameView& Cur, const FrameView& Last, const float th, const bool bMono,
                                   std::vector<int>& vnMatches, bool reportRemoved) {
  const int n1 = (int)Last.mvKeysUn.size(), n2 = (int)Cur.mvKeysUn.size();
  vnMatches.assign(n2, -1);
  if (n1 == 0 || n2 == 0) return 0;
  need((int)Last.mvKeys.size() == n1 && (int)Last.hasMapPoint.size() == n1 && (int)Last.mapPointObserved.size() == n1 &&
           (Last.mvbOutlier.empty() || (int)Last.mvbOutlier.size() == n1) && Last.mapPointWorldPos.size() == (size_t)n1 * 3 &&
           Last.mapPointDescriptor.rows >= n1 && Last.mapPointDescriptor.cols == 32,
       "LastFrame view: mvKeys / hasMapPoint / mapPointObserved / mvbOutlier / mapPointWorldPos / mapPointDescriptor do not cover its N features");
  need((int)Cur.mvuRight.size() == n2 && (int)Cur.mapPointObserved.size() == n2 && Cur.mDescriptors.rows >= n2 &&
           Cur.mDescriptors.cols == 32 && Cur.gridStart.size() == 64 * 48 + 1 &&
           (int)Cur.gridItems.size() == Cur.gridStart.back() && !Cur.mvScaleFactors.empty(),
       "CurrentFrame view: mvuRight (fill with -1 for monocular frames) / mapPointObserved / mDescriptors / grid / mvScaleFactors incomplete");
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
  j.report_removed = reportRemoved ? 1 : 0;
  check(plslam_match_projection_host(&j, (int)Cur.mvScaleFactors.size()), "SearchByProjection");
  return nm;
}

int ORBmatcher::SearchByProjection(FrameView& Cur, const KeyFramePointsView& KF, const float th, const int ORBdist,
                                   std::vector<int>& vnMatches) {
  const int m = (int)KF.valid.size(), n2 = (int)Cur.mvKeysUn.size();
  vnMatches.assign(n2, -1);
  if (m == 0 || n2 == 0) return 0;
  need(KF.worldPos.size() == (size_t)m * 3 && (int)KF.angle.size() == m && (int)KF.level.size() == m && KF.descriptors.rows >= m &&
           KF.descriptors.cols == 32,
       "KeyFramePointsView arrays do not cover its M map points");
  need((int)Cur.hasMapPoint.size() == n2 && Cur.mDescriptors.rows >= n2 && Cur.mDescriptors.cols == 32 &&
           Cur.gridStart.size() == 64 * 48 + 1 && (int)Cur.gridItems.size() == Cur.gridStart.back() && !Cur.mvScaleFactors.empty(),
       "Frame view: hasMapPoint / mDescriptors / grid / mvScaleFactors incomplete");
  for (int i = 0; i < m; ++i)
    need(!KF.valid[i] || (KF.level[i] >= 0 && KF.level[i] < (int)Cur.mvScaleFactors.size()), "KeyFramePointsView level out of range");
  std::vector<float> cxy((size_t)n2 * 2), cang(n2);
  std::vector<int32_t> coct(n2);
  for (int i = 0; i < n2; ++i) {
    cxy[2 * i] = Cur.mvKeysUn[i].pt.x; cxy[2 * i + 1] = Cur.mvKeysUn[i].pt.y;
    coct[i] = Cur.mvKeysUn[i].octave; cang[i] = Cur.mvKeysUn[i].angle;
  }
  int32_t nm = 0;
  plslam_proj_job_t j{};
  j.last_valid = KF.valid.data(); j.last_xyz = KF.worldPos.data(); j.last_desc = KF.descriptors.data;
  j.last_octave = KF.level.data(); j.last_angle = KF.angle.data();
  j.cur_xy = cxy.data(); j.cur_octave = coct.data(); j.cur_angle = cang.data(); j.cur_desc = Cur.mDescriptors.data;
  j.cur_taken = Cur.hasMapPoint.data();
  j.grid_start = Cur.gridStart.data(); j.grid_items = Cur.gridItems.data(); j.scale_factors = Cur.mvScaleFactors.data();
  j.match_cur = vnMatches.data(); j.nmatches = &nm;
  const float cam[12] = {Cur.fx, Cur.fy, Cur.cx, Cur.cy, 0.f, 0.f, Cur.mnMinX, Cur.mnMaxX, Cur.mnMinY, Cur.mnMaxY,
                         Cur.mfGridElementWidthInv, Cur.mfGridElementHeightInv};
  std::memcpy(j.cam, cam, sizeof(cam));
  std::memcpy(j.tcw_cur, Cur.mTcw, sizeof(j.tcw_cur));
  j.th = th; j.n1 = m; j.n2 = n2; j.check_orientation = mbCheckOrientation;
  j.mode = 1; j.orb_dist = ORBdist; j.n_levels = (int)Cur.mvScaleFactors.size();
  check(plslam_match_projection_host(&j, (int)Cur.mvScaleFactors.size()), "SearchByProjection(Frame, KeyFrame)");
  return nm;
}

int ORBmatcher::SearchByProjection(const KeyFrameGridView& KF, const float Scw[12], const LoopPointsView& P,
                                   const std::vector<uint8_t>& matchedOnEntry, int th, std::vector<int>& vnMatches) {
  const int m = (int)P.valid.size(), n = (int)KF.mvKeysUn.size();
  vnMatches.assign(n, -1);
  if (m == 0 || n == 0) return 0;
  need(P.worldPos.size() == (size_t)m * 3 && (int)P.level.size() == m && P.descriptors.rows >= m && P.descriptors.cols == 32,
       "LoopPointsView arrays do not cover its M map points");
  need((int)matchedOnEntry.size() == n && KF.mDescriptors.rows >= n && KF.mDescriptors.cols == 32 && !KF.mvScaleFactors.empty() &&
           KF.gridStart.size() == (size_t)KF.mnGridCols * KF.mnGridRows + 1 && (int)KF.gridItems.size() == KF.gridStart.back(),
       "KeyFrameGridView: vpMatched / mDescriptors / grid / mvScaleFactors incomplete");
  for (int i = 0; i < m; ++i)
    need(!P.valid[i] || (P.level[i] >= 0 && P.level[i] < (int)KF.mvScaleFactors.size()), "LoopPointsView level out of range");
  std::vector<float> xy((size_t)n * 2);
  std::vector<int32_t> oct(n);
  for (int i = 0; i < n; ++i) {
    xy[2 * i] = KF.mvKeysUn[i].pt.x; xy[2 * i + 1] = KF.mvKeysUn[i].pt.y;
    oct[i] = KF.mvKeysUn[i].octave;
  }
  int32_t nm = 0;
  plslam_kfproj_job_t j{};
  j.mp_valid = P.valid.data(); j.mp_xyz = P.worldPos.data(); j.mp_desc = P.descriptors.data; j.mp_level = P.level.data();
  j.kf_xy = xy.data(); j.kf_octave = oct.data(); j.kf_desc = KF.mDescriptors.data; j.kf_matched = matchedOnEntry.data();
  j.grid_start = KF.gridStart.data(); j.grid_items = KF.gridItems.data(); j.scale_factors = KF.mvScaleFactors.data();
  j.match_kf = vnMatches.data(); j.nmatches = &nm;
  std::memcpy(j.scw, Scw, sizeof(j.scw));
  j.cam[0] = KF.fx; j.cam[1] = KF.fy; j.cam[2] = KF.cx; j.cam[3] = KF.cy;
  j.bounds[0] = KF.mnMinX; j.bounds[1] = KF.mnMinY; j.bounds[2] = KF.mnMaxX; j.bounds[3] = KF.mnMaxY;
  j.grid_width_inv = KF.mfGridElementWidthInv; j.grid_height_inv = KF.mfGridElementHeightInv;
  j.grid_cols = KF.mnGridCols; j.grid_rows = KF.mnGridRows; j.n_levels = (int)KF.mvScaleFactors.size(); j.th = th; j.m = m; j.n = n;
  check(plslam_match_kf_projection_host(&j), "SearchByProjection(KeyFrame, Scw)");
  return nm;
}

void ORBmatcher::FuseSearch(const KeyFrameGridView& KF, const float pose[12], const float ow[3], bool useScw,
                            const LoopPointsView& P, float th, std::vector<int>& vnBestIdx) {
  const int m = (int)P.valid.size(), n = (int)KF.mvKeysUn.size();
  vnBestIdx.assign(m, -1);
  if (m == 0 || n == 0) return;
  need(P.worldPos.size() == (size_t)m * 3 && (int)P.level.size() == m && P.descriptors.rows >= m && P.descriptors.cols == 32,
       "LoopPointsView arrays do not cover its M map points");
  need(KF.mDescriptors.rows >= n && KF.mDescriptors.cols == 32 && !KF.mvScaleFactors.empty() &&
           KF.gridStart.size() == (size_t)KF.mnGridCols * KF.mnGridRows + 1 && (int)KF.gridItems.size() == KF.gridStart.back() &&
           (useScw || ((int)KF.mvuRight.size() == n && KF.mvInvLevelSigma2.size() == KF.mvScaleFactors.size())),
       "KeyFrameGridView: mDescriptors / grid / mvScaleFactors / mvuRight / mvInvLevelSigma2 incomplete");
  for (int i = 0; i < m; ++i)
    need(!P.valid[i] || (P.level[i] >= 0 && P.level[i] < (int)KF.mvScaleFactors.size()), "LoopPointsView level out of range");
  std::vector<float> xy((size_t)n * 2);
  std::vector<int32_t> oct(n);
  for (int i = 0; i < n; ++i) {
    xy[2 * i] = KF.mvKeysUn[i].pt.x; xy[2 * i + 1] = KF.mvKeysUn[i].pt.y;
    oct[i] = KF.mvKeysUn[i].octave;
  }
  plslam_fuse_job_t j{};
  j.mp_valid = P.valid.data(); j.mp_xyz = P.worldPos.data(); j.mp_desc = P.descriptors.data; j.mp_level = P.level.data();
  j.kf_xy = xy.data(); j.kf_octave = oct.data(); j.kf_desc = KF.mDescriptors.data;
  j.kf_uright = useScw ? nullptr : KF.mvuRight.data(); j.inv_level_sigma2 = useScw ? nullptr : KF.mvInvLevelSigma2.data();
  j.grid_start = KF.gridStart.data(); j.grid_items = KF.gridItems.data(); j.scale_factors = KF.mvScaleFactors.data();
  j.best_idx = vnBestIdx.data();
  std::memcpy(j.pose, pose, sizeof(j.pose));
  if (ow) std::memcpy(j.ow, ow, sizeof(j.ow));
  j.cam[0] = KF.fx; j.cam[1] = KF.fy; j.cam[2] = KF.cx; j.cam[3] = KF.cy; j.cam[4] = KF.mbf;
  j.bounds[0] = KF.mnMinX; j.bounds[1] = KF.mnMinY; j.bounds[2] = KF.mnMaxX; j.bounds[3] = KF.mnMaxY;
  j.grid_width_inv = KF.mfGridElementWidthInv; j.grid_height_inv = KF.mfGridElementHeightInv; j.th = th;
  j.grid_cols = KF.mnGridCols; j.grid_rows = KF.mnGridRows; j.n_levels = (int)KF.mvScaleFactors.size(); j.use_scw = useScw ? 1 : 0;
  j.m = m; j.n = n;
  check(plslam_match_fuse_search_host(&j), "Fuse");
}

void ORBmatcher::Sim3Transforms(float s12, const float R12[9], const float t12[3], float sR12[9], float sR21[9], float t21[3]) {
  plslam_sim3_transforms(s12, R12, t12, sR12, sR21, t21);
}

void ORBmatcher::Sim3Search(const KeyFrameGridView& KF, const float poseOwn[12], const float pose2[12], const LoopPointsView& P,
                            float th, std::vector<int>& vnBestIdx) {
  const int m = (int)P.valid.size(), n = (int)KF.mvKeysUn.size();
  vnBestIdx.assign(m, -1);
  if (m == 0 || n == 0) return;
  need(P.worldPos.size() == (size_t)m * 3 && (int)P.level.size() == m && P.descriptors.rows >= m && P.descriptors.cols == 32,
       "LoopPointsView arrays do not cover its M map points");
  need(KF.mDescriptors.rows >= n && KF.mDescriptors.cols == 32 && !KF.mvScaleFactors.empty() &&
           KF.gridStart.size() == (size_t)KF.mnGridCols * KF.mnGridRows + 1 && (int)KF.gridItems.size() == KF.gridStart.back(),
       "KeyFrameGridView: mDescriptors / grid / mvScaleFactors incomplete");
  for (int i = 0; i < m; ++i)
    need(!P.valid[i] || (P.level[i] >= 0 && P.level[i] < (int)KF.mvScaleFactors.size()), "LoopPointsView level out of range");
  std::vector<float> xy((size_t)n * 2);
  std::vector<int32_t> oct(n);
  for (int i = 0; i < n; ++i) {
    xy[2 * i] = KF.mvKeysUn[i].pt.x; xy[2 * i + 1] = KF.mvKeysUn[i].pt.y;
    oct[i] = KF.mvKeysUn[i].octave;
  }
  plslam_fuse_job_t j{};
  j.mp_valid = P.valid.data(); j.mp_xyz = P.worldPos.data(); j.mp_desc = P.descriptors.data; j.mp_level = P.level.data();
  j.kf_xy = xy.data(); j.kf_octave = oct.data(); j.kf_desc = KF.mDescriptors.data;
  j.grid_start = KF.gridStart.data(); j.grid_items = KF.gridItems.data(); j.scale_factors = KF.mvScaleFactors.data();
  j.best_idx = vnBestIdx.data();
  std::memcpy(j.pose, poseOwn, sizeof(j.pose));
  std::memcpy(j.pose2, pose2, sizeof(j.pose2));
  j.cam[0] = KF.fx; j.cam[1] = KF.fy; j.cam[2] = KF.cx; j.cam[3] = KF.cy;
  j.bounds[0] = KF.mnMinX; j.bounds[1] = KF.mnMinY; j.bounds[2] = KF.mnMaxX; j.bounds[3] = KF.mnMaxY;
  j.grid_width_inv = KF.mfGridElementWidthInv; j.grid_height_inv = KF.mfGridElementHeightInv; j.th = th;
  j.grid_cols = KF.mnGridCols; j.grid_rows = KF.mnGridRows; j.n_levels = (int)KF.mvScaleFactors.size(); j.use_scw = 2;
  j.m = m; j.n = n;
  check(plslam_match_fuse_search_host(&j), "SearchBySim3");
}

int ORBmatcher::SearchByProjection(FrameView& F, const MapPointsView& MP, const float th, std::vector<int>& vnMatches) {
  const int m = (int)MP.inViewAndGood.size(), n = (int)F.mvKeysUn.size();
  vnMatches.assign(n, -1);
  if (m == 0 || n == 0) return 0;
  need(MP.trackProj.size() == (size_t)m * 3 && (int)MP.trackScaleLevel.size() == m && (int)MP.trackViewCos.size() == m &&
           (int)MP.observed.size() == m && MP.descriptors.rows >= m && MP.descriptors.cols == 32,
       "MapPointsView arrays do not cover its M map points");
  need((int)F.mvuRight.size() == n && (int)F.mapPointObserved.size() == n && F.mDescriptors.rows >= n && F.mDescriptors.cols == 32 &&
           F.gridStart.size() == 64 * 48 + 1 && (int)F.gridItems.size() == F.gridStart.back() && !F.mvScaleFactors.empty(),
       "Frame view: mvuRight (fill with -1 for monocular frames) / mapPointObserved / mDescriptors / grid / mvScaleFactors incomplete");
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
  need((int)KF.hasMapPoint.size() == n1 && KF.mDescriptors.rows >= n1 && KF.mDescriptors.cols == 32 && F.mDescriptors.rows >= n2 &&
           F.mDescriptors.cols == 32 && KF.mFeatVec.start.size() == KF.mFeatVec.nodes.size() + 1 &&
           F.mFeatVec.start.size() == F.mFeatVec.nodes.size() + 1,
       "SearchByBoW views: hasMapPoint / mDescriptors / mFeatVec incomplete");
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

int ORBmatcher::SearchByBoW(const FrameView& KF1, const FrameView& KF2, std::vector<int>& vnMatches12, int) {
  const int n1 = (int)KF1.mvKeysUn.size(), n2 = (int)KF2.mvKeysUn.size();
  vnMatches12.assign(n1, -1);
  if (n1 == 0 || n2 == 0) return 0;
  need((int)KF1.hasMapPoint.size() == n1 && (int)KF2.hasMapPoint.size() == n2 && KF1.mDescriptors.rows >= n1 &&
           KF1.mDescriptors.cols == 32 && KF2.mDescriptors.rows >= n2 && KF2.mDescriptors.cols == 32 &&
           KF1.mFeatVec.start.size() == KF1.mFeatVec.nodes.size() + 1 && KF2.mFeatVec.start.size() == KF2.mFeatVec.nodes.size() + 1,
       "SearchByBoW(KF, KF) views: hasMapPoint / mDescriptors / mFeatVec incomplete");
  std::vector<float> a1(n1), a2(n2);
  for (int i = 0; i < n1; ++i) a1[i] = KF1.mvKeysUn[i].angle;
  for (int i = 0; i < n2; ++i) a2[i] = KF2.mvKeysUn[i].angle;
  std::vector<int32_t> scratch(n2);
  int32_t nm = 0;
  plslam_bow_job_t j{};
  j.kf_desc = KF1.mDescriptors.data; j.kf_angle = a1.data(); j.kf_valid = KF1.hasMapPoint.data();
  j.kf_nodes = KF1.mFeatVec.nodes.data(); j.kf_start = KF1.mFeatVec.start.data(); j.kf_idx = KF1.mFeatVec.idx.data();
  j.f_desc = KF2.mDescriptors.data; j.f_angle = a2.data(); j.f_valid = KF2.hasMapPoint.data();
  j.f_nodes = KF2.mFeatVec.nodes.data(); j.f_start = KF2.mFeatVec.start.data(); j.f_idx = KF2.mFeatVec.idx.data();
  j.match_f = scratch.data(); j.nmatches = &nm;
  j.n1 = n1; j.n2 = n2; j.n_kf_nodes = (int)KF1.mFeatVec.nodes.size(); j.n_f_nodes = (int)KF2.mFeatVec.nodes.size();
  j.nnratio = mfNNratio; j.check_orientation = mbCheckOrientation;
  check(plslam_match_bow_kfkf_host(&j, vnMatches12.data()), "SearchByBoW(KF, KF)");
  return nm;
}

int ORBmatcher::SearchForInitialization(const FrameView& F1, const FrameView& F2, std::vector<cv::Point2f>& vbPrevMatched,
                                        std::vector<int>& vnMatches12, int windowSize) {
  const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
  vnMatches12.assign(n1, -1);
  if (n1 == 0 || n2 == 0) return 0;
  need((int)vbPrevMatched.size() == n1 && F1.mDescriptors.rows >= n1 && F2.mDescriptors.rows >= n2 &&
           F2.gridStart.size() == 64 * 48 + 1 && (int)F2.gridItems.size() == F2.gridStart.back(),
       "SearchForInitialization views: vbPrevMatched / mDescriptors / grid incomplete");
  std::vector<int32_t> o1(n1), o2(n2);
  std::vector<float> a1(n1), a2(n2), xy2((size_t)n2 * 2), prev((size_t)n1 * 2);
  for (int i = 0; i < n1; ++i) {
    o1[i] = F1.mvKeysUn[i].octave; a1[i] = F1.mvKeysUn[i].angle;
    prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y;
  }
  for (int i = 0; i < n2; ++i) {
    o2[i] = F2.mvKeysUn[i].octave; a2[i] = F2.mvKeysUn[i].angle;
    xy2[2 * i] = F2.mvKeysUn[i].pt.x; xy2[2 * i + 1] = F2.mvKeysUn[i].pt.y;
  }
  int32_t nm = 0;
  plslam_init_job_t j{};
  j.f1_octave = o1.data(); j.f1_angle = a1.data(); j.f1_desc = F1.mDescriptors.data;
  j.f2_xy = xy2.data(); j.f2_angle = a2.data(); j.f2_octave = o2.data(); j.f2_desc = F2.mDescriptors.data;
  j.grid_start = F2.gridStart.data(); j.grid_items = F2.gridItems.data();
  j.prev_matched = prev.data(); j.match12 = vnMatches12.data(); j.nmatches = &nm;
  j.cam[0] = F2.mnMinX; j.cam[1] = F2.mnMinY; j.cam[2] = F2.mfGridElementWidthInv; j.cam[3] = F2.mfGridElementHeightInv;
  j.nnratio = mfNNratio; j.window_size = windowSize; j.n1 = n1; j.n2 = n2; j.check_orientation = mbCheckOrientation;
  check(plslam_match_initialization_host(&j), "SearchForInitialization");
  for (int i = 0; i < n1; ++i) vbPrevMatched[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]);
  return nm;
}

void ORBmatcher::Epipole(const float R2w[9], const float t2w[3], const float Cw[3], float fx, float fy, float cx, float cy,
                         float* ex, float* ey) {
  check(plslam_match_epipole(R2w, t2w, Cw, fx, fy, cx, cy, ex, ey), "Epipole");
}

int ORBmatcher::SearchForTriangulation(const FrameView& KF1, const FrameView& KF2, const float F12[9], float ex, float ey,
                                       const bool bOnlyStereo, std::vector<int>& vnMatches12) {
  const int n1 = (int)KF1.mvKeysUn.size(), n2 = (int)KF2.mvKeysUn.size();
  vnMatches12.assign(n1, -1);
  if (n1 == 0 || n2 == 0) return 0;
  need((int)KF1.mvuRight.size() == n1 && (int)KF1.hasMapPoint.size() == n1 && KF1.mDescriptors.rows >= n1 &&
           (int)KF2.mvuRight.size() == n2 && (int)KF2.hasMapPoint.size() == n2 && KF2.mDescriptors.rows >= n2 &&
           KF1.mFeatVec.start.size() == KF1.mFeatVec.nodes.size() + 1 && KF2.mFeatVec.start.size() == KF2.mFeatVec.nodes.size() + 1 &&
           !KF2.mvScaleFactors.empty() && KF2.mvLevelSigma2.size() == KF2.mvScaleFactors.size(),
       "SearchForTriangulation views: mvuRight / hasMapPoint / mDescriptors / mFeatVec / mvScaleFactors / mvLevelSigma2 incomplete");
  std::vector<float> xy1((size_t)n1 * 2), a1(n1), xy2((size_t)n2 * 2), a2(n2);
  std::vector<int32_t> o2(n2);
  for (int i = 0; i < n1; ++i) {
    xy1[2 * i] = KF1.mvKeysUn[i].pt.x; xy1[2 * i + 1] = KF1.mvKeysUn[i].pt.y; a1[i] = KF1.mvKeysUn[i].angle;
  }
  for (int i = 0; i < n2; ++i) {
    xy2[2 * i] = KF2.mvKeysUn[i].pt.x; xy2[2 * i + 1] = KF2.mvKeysUn[i].pt.y; a2[i] = KF2.mvKeysUn[i].angle;
    o2[i] = KF2.mvKeysUn[i].octave;
  }
  int32_t nm = 0;
  plslam_tri_job_t j{};
  j.kf1_desc = KF1.mDescriptors.data; j.kf1_xy = xy1.data(); j.kf1_angle = a1.data(); j.kf1_uright = KF1.mvuRight.data();
  j.kf1_has_mp = KF1.hasMapPoint.data(); j.kf1_nodes = KF1.mFeatVec.nodes.data(); j.kf1_start = KF1.mFeatVec.start.data();
  j.kf1_idx = KF1.mFeatVec.idx.data();
  j.kf2_desc = KF2.mDescriptors.data; j.kf2_xy = xy2.data(); j.kf2_angle = a2.data(); j.kf2_octave = o2.data();
  j.kf2_uright = KF2.mvuRight.data(); j.kf2_has_mp = KF2.hasMapPoint.data(); j.kf2_nodes = KF2.mFeatVec.nodes.data();
  j.kf2_start = KF2.mFeatVec.start.data(); j.kf2_idx = KF2.mFeatVec.idx.data();
  j.scale_factors = KF2.mvScaleFactors.data(); j.level_sigma2 = KF2.mvLevelSigma2.data();
  j.match12 = vnMatches12.data(); j.nmatches = &nm;
  std::memcpy(j.F12, F12, sizeof(j.F12));
  j.ex = ex; j.ey = ey;
  j.n1 = n1; j.n2 = n2; j.n1_nodes = (int)KF1.mFeatVec.nodes.size(); j.n2_nodes = (int)KF2.mFeatVec.nodes.size();
  j.only_stereo = bOnlyStereo; j.check_orientation = mbCheckOrientation;
  check(plslam_match_triangulation_host(&j, (int)KF2.mvScaleFactors.size()), "SearchForTriangulation");
  return nm;
}

}  // namespace ORB_SLAM2
