// dropin_matchers_demo.cc — ORB_SLAM2::ORBmatcher called through the REFERENCE'S signatures (include/ORBmatcher.h:61, :78,
// :104, :111) on mock Frame / KeyFrame / MapPoint classes that carry the member names of the reference's
// include/Frame.h, KeyFrame.h and MapPoint.h (the real classes are pointer graphs guarded by mutexes, outside the hot
// path).  tests/test_dropin_gpu.py writes the array form of four matcher cases into a directory, this program builds the
// object form, calls the matcher the way Tracking / LocalMapping do, and writes what the calls left in the objects;
// the test compares that with the flattened C-ABI calls on the same arrays.
//   dropin_matchers_demo <dir>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/plslam_b200.h"
#include "ORBmatcher.h"

namespace DBoW2 {
typedef std::map<unsigned int, std::vector<unsigned int> > FeatureVector;  // DBoW2/FeatureVector.h
}

namespace ORB_SLAM2 {

class Frame;
class KeyFrame;
class MapPoint {  // include/MapPoint.h: the members the matchers touch
 public:
  cv::Mat GetNormal() { return mNormalVector.clone(); }
  // map-graph bookkeeping called by ORBmatcher::Fuse: the demo's stand-ins log the call and apply the minimum effect a later
  // iteration can observe, exactly like the stand-ins of tests/golden/reference_code.py under which the reference's Fuse was run
  bool IsInKeyFrame(KeyFrame*) { return inKF; }
  int GetIndexInKeyFrame(KeyFrame*) { return idxInKF2; }
  int idxInKF2 = -1;  // demo: where the second key frame of case_sim3 observes this point
  void AddObservation(KeyFrame* pKF, size_t idx);
  void Replace(MapPoint* pMP);
  bool inKF = false;
  int PredictScale(const float& currentDist, KeyFrame* pKF);
  cv::Mat mNormalVector;  // (protected in the reference)
  float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }
  float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
  int PredictScale(const float& currentDist, Frame* pF);
  float mfMinDistance = 0, mfMaxDistance = 0;  // (protected in the reference)
  cv::Mat GetWorldPos() { return mWorldPos.clone(); }
  cv::Mat GetDescriptor() { return mDescriptor.clone(); }
  int Observations() { return nObs; }
  bool isBad() { return mbBad; }
  // Variables used by the tracking
  float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0;
  bool mbTrackInView = false;
  int mnTrackScaleLevel = 0;
  float mTrackViewCos = 0;
  // (protected in the reference)
  cv::Mat mWorldPos, mDescriptor;
  int nObs = 0;
  bool mbBad = false;
  int id = -1;  // demo only: position in the vector it came from
};

class Frame {  // include/Frame.h
 public:
  static float fx, fy, cx, cy;
  float mbf = 0, mb = 0;
  int N = 0;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
  std::vector<float> mvuRight;
  DBoW2::FeatureVector mFeatVec;
  cv::Mat mDescriptors;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  static float mfGridElementWidthInv, mfGridElementHeightInv;
  std::vector<std::size_t> mGrid[64][48];
  cv::Mat mTcw;
  int mnScaleLevels = 0;
  float mfLogScaleFactor = 0;
  std::vector<float> mvScaleFactors;
  static float mnMinX, mnMaxX, mnMinY, mnMaxY;
};
// MapPoint.cc of the reference stays on the CPU; the demo's stand-in evaluates the same formula (@0x8fc20)
int MapPoint::PredictScale(const float& currentDist, Frame* pF) {
  return plslam_predict_scale(mfMaxDistance, currentDist, pF->mfLogScaleFactor, pF->mnScaleLevels);
}
float Frame::fx, Frame::fy, Frame::cx, Frame::cy, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv, Frame::mnMinX,
    Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;

class KeyFrame {  // include/KeyFrame.h
 public:
  std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
  cv::Mat GetRotation() { return R.clone(); }
  cv::Mat GetTranslation() { return t.clone(); }
  cv::Mat GetCameraCenter() { return Ow.clone(); }
  float fx = 0, fy = 0, cx = 0, cy = 0;
  std::vector<cv::KeyPoint> mvKeysUn;
  std::vector<float> mvuRight;
  cv::Mat mDescriptors;
  DBoW2::FeatureVector mFeatVec;
  std::vector<float> mvScaleFactors, mvLevelSigma2;
  std::vector<MapPoint*> mvpMapPoints;
  cv::Mat R, t, Ow;
  int mnGridCols = 64, mnGridRows = 48;
  float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
  int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
  int mnScaleLevels = 0;
  float mfLogScaleFactor = 0;
  std::vector<std::vector<std::vector<size_t> > > mGrid;  // (protected in the reference: see ORBmatcher.h)
  std::vector<float> mvInvLevelSigma2;
  float mbf = 0;
  void AddMapPoint(MapPoint* pMP, const size_t& idx);
  std::set<MapPoint*> GetMapPoints() {
    std::set<MapPoint*> s;
    for (MapPoint* p : mvpMapPoints)
      if (p && !p->isBad()) s.insert(p);
    return s;
  }
};
static std::vector<int32_t> g_log;
static KeyFrame* g_fuse_kf = nullptr;
void MapPoint::AddObservation(KeyFrame* pKF, size_t idx) {
  g_log.push_back(1); g_log.push_back(id); g_log.push_back((int32_t)idx);
  nObs += pKF->mvuRight[idx] >= 0 ? 2 : 1;
  inKF = true;
}
void MapPoint::Replace(MapPoint* pMP) {
  g_log.push_back(3); g_log.push_back(id); g_log.push_back(pMP->id);
  mbBad = true;
  for (MapPoint*& q : g_fuse_kf->mvpMapPoints)
    if (q == this) q = pMP;
}
void KeyFrame::AddMapPoint(MapPoint* pMP, const size_t& idx) {
  g_log.push_back(2); g_log.push_back(pMP->id); g_log.push_back((int32_t)idx);
  mvpMapPoints[idx] = pMP;
}
int MapPoint::PredictScale(const float& currentDist, KeyFrame* pKF) {
  return plslam_predict_scale(mfMaxDistance, currentDist, pKF->mfLogScaleFactor, pKF->mnScaleLevels);
}

}  // namespace ORB_SLAM2

using namespace ORB_SLAM2;

static std::string g_dir;
template <class T>
static std::vector<T> load(const std::string& name) {
  std::ifstream f(g_dir + "/" + name, std::ios::binary | std::ios::ate);
  if (!f) {
    std::fprintf(stderr, "missing %s\n", name.c_str());
    std::exit(2);
  }
  const size_t bytes = (size_t)f.tellg();
  std::vector<T> v(bytes / sizeof(T));
  f.seekg(0);
  f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
  return v;
}
template <class T>
static void save(const std::string& name, const std::vector<T>& v) {
  std::ofstream f(g_dir + "/" + name, std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
}
static cv::Mat mat_u8(const std::vector<uint8_t>& d, int rows) {
  cv::Mat m(rows > 0 ? rows : 1, 32, CV_8U);
  if (rows) std::memcpy(m.data, d.data(), (size_t)rows * 32);
  m.rows = rows;
  return m;
}
static cv::Mat mat_f(const float* p, int r, int c) {
  cv::Mat m(r, c, CV_32F);
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) m.at<float>(i, j) = p[i * c + j];
  return m;
}
static cv::Mat desc_row(const std::vector<uint8_t>& d, int i) {
  cv::Mat m(1, 32, CV_8U);
  std::memcpy(m.data, d.data() + (size_t)i * 32, 32);
  return m;
}
static DBoW2::FeatureVector featvec(const std::vector<int32_t>& nodes, const std::vector<int32_t>& start, const std::vector<int32_t>& idx) {
  DBoW2::FeatureVector fv;
  for (size_t k = 0; k < nodes.size(); ++k)
    for (int j = start[k]; j < start[k + 1]; ++j) fv[(unsigned)nodes[k]].push_back((unsigned)idx[j]);
  return fv;
}
static void fill_grid(Frame& F, const std::vector<int32_t>& gs, const std::vector<int32_t>& gi) {
  for (int ix = 0; ix < 64; ++ix)
    for (int iy = 0; iy < 48; ++iy) {
      F.mGrid[ix][iy].clear();
      for (int j = gs[ix * 48 + iy]; j < gs[ix * 48 + iy + 1]; ++j) F.mGrid[ix][iy].push_back((size_t)gi[j]);
    }
}
static std::vector<MapPoint*> g_pool;
static MapPoint* new_mp() {
  g_pool.push_back(new MapPoint());
  return g_pool.back();
}

// ---- TrackWithMotionModel: matcher.SearchByProjection(mCurrentFrame, mLastFrame, th, mSensor == MONOCULAR) ----
static void case_projection() {
  const auto valid = load<uint8_t>("pj_last_valid"), obs = load<uint8_t>("pj_last_obs"), ldesc = load<uint8_t>("pj_last_desc");
  const auto xyz = load<float>("pj_last_xyz"), lang = load<float>("pj_last_angle");
  const auto loct = load<int32_t>("pj_last_octave");
  const auto cxy = load<float>("pj_cur_xy"), cang = load<float>("pj_cur_angle"), cur = load<float>("pj_cur_uright");
  const auto coct = load<int32_t>("pj_cur_octave"), gs = load<int32_t>("pj_cur_grid_start"), gi = load<int32_t>("pj_cur_grid_items");
  const auto cdesc = load<uint8_t>("pj_cur_desc"), taken = load<uint8_t>("pj_cur_taken");
  const auto cam = load<float>("pj_cam"), sf = load<float>("pj_sf"), tc = load<float>("pj_tc"), tl = load<float>("pj_tl"),
             par = load<float>("pj_par");  // th, mono
  const int n1 = (int)valid.size(), n2 = (int)taken.size();
  Frame::fx = cam[0]; Frame::fy = cam[1]; Frame::cx = cam[2]; Frame::cy = cam[3];
  Frame::mnMinX = cam[6]; Frame::mnMaxX = cam[7]; Frame::mnMinY = cam[8]; Frame::mnMaxY = cam[9];
  Frame::mfGridElementWidthInv = cam[10]; Frame::mfGridElementHeightInv = cam[11];
  Frame Last, Cur;
  Last.N = n1; Last.mbf = Cur.mbf = cam[4]; Last.mb = Cur.mb = cam[5];
  Last.mvKeys.resize(n1); Last.mvKeysUn.resize(n1); Last.mvuRight.assign(n1, -1.f); Last.mvpMapPoints.assign(n1, nullptr);
  Last.mvbOutlier.assign(n1, false);
  Last.mDescriptors = mat_u8(ldesc, n1);
  for (int i = 0; i < n1; ++i) {
    Last.mvKeys[i].octave = Last.mvKeysUn[i].octave = loct[i];
    Last.mvKeys[i].angle = Last.mvKeysUn[i].angle = lang[i];
    // valid = mvpMapPoints[i] && !mvbOutlier[i]: invalid ones alternate between "no map point" and "outlier"
    if (valid[i] || (i & 1)) {
      MapPoint* p = new_mp();
      p->id = i;
      p->mWorldPos = mat_f(&xyz[3 * (size_t)i], 3, 1);
      p->mDescriptor = desc_row(ldesc, i);
      p->nObs = obs[i] ? 2 : 0;
      Last.mvpMapPoints[i] = p;
      if (!valid[i]) Last.mvbOutlier[i] = true;
    }
  }
  float T[16] = {0};
  std::memcpy(T, tl.data(), 12 * sizeof(float)); T[15] = 1;
  Last.mTcw = mat_f(T, 4, 4);
  Last.mvScaleFactors = sf;
  Cur.N = n2;
  Cur.mvKeys.resize(n2); Cur.mvKeysUn.resize(n2); Cur.mvuRight = cur; Cur.mvpMapPoints.assign(n2, nullptr); Cur.mvbOutlier.assign(n2, false);
  Cur.mDescriptors = mat_u8(cdesc, n2);
  std::vector<MapPoint*> before(n2, nullptr);
  for (int i = 0; i < n2; ++i) {
    Cur.mvKeysUn[i].pt = cv::Point2f(cxy[2 * i], cxy[2 * i + 1]);
    Cur.mvKeys[i].pt = Cur.mvKeysUn[i].pt;
    Cur.mvKeys[i].octave = Cur.mvKeysUn[i].octave = coct[i];
    Cur.mvKeys[i].angle = Cur.mvKeysUn[i].angle = cang[i];
    // taken = mvpMapPoints[i] && Observations() > 0; a few more carry a map point nobody observes yet (may be replaced)
    if (taken[i] || i % 7 == 0) {
      MapPoint* p = new_mp();
      p->nObs = taken[i] ? 1 : 0;
      Cur.mvpMapPoints[i] = before[i] = p;
    }
  }
  std::memcpy(T, tc.data(), 12 * sizeof(float));
  Cur.mTcw = mat_f(T, 4, 4);
  Cur.mvScaleFactors = sf;
  fill_grid(Cur, gs, gi);

  ORBmatcher matcher(0.9f, true);
  const int nmatches = matcher.SearchByProjection(Cur, Last, par[0], par[1] != 0.f);

  std::vector<int32_t> out(n2 + 1);
  for (int i = 0; i < n2; ++i) {
    MapPoint* p = Cur.mvpMapPoints[i];
    out[i] = !p ? (before[i] ? -2 : -1) : (p == before[i] ? -1 : p->id);  // -2: had a map point, NULL now
  }
  out[n2] = nmatches;
  save("pj_out", out);
}

// ---- Tracking::SearchLocalPoints: matcher.SearchByProjection(mCurrentFrame, mvpLocalMapPoints, th) ----
static void case_local() {
  const auto valid = load<uint8_t>("lp_mp_valid"), obs = load<uint8_t>("lp_mp_obs"), mdesc = load<uint8_t>("lp_mp_desc");
  const auto proj = load<float>("lp_mp_proj"), vcos = load<float>("lp_mp_viewcos");
  const auto level = load<int32_t>("lp_mp_level");
  const auto xy = load<float>("lp_fr_xy"), ur = load<float>("lp_fr_uright");
  const auto oct = load<int32_t>("lp_fr_octave"), gs = load<int32_t>("lp_fr_grid_start"), gi = load<int32_t>("lp_fr_grid_items");
  const auto fdesc = load<uint8_t>("lp_fr_desc"), taken = load<uint8_t>("lp_fr_taken");
  const auto cam4 = load<float>("lp_cam4"), sf = load<float>("lp_sf"), par = load<float>("lp_par");  // th, nnratio
  const int m = (int)valid.size(), n = (int)taken.size();
  Frame::mnMinX = cam4[0]; Frame::mnMinY = cam4[1]; Frame::mfGridElementWidthInv = cam4[2]; Frame::mfGridElementHeightInv = cam4[3];
  Frame F;
  F.N = n;
  F.mvKeys.resize(n); F.mvKeysUn.resize(n); F.mvuRight = ur; F.mvpMapPoints.assign(n, nullptr); F.mvbOutlier.assign(n, false);
  F.mDescriptors = mat_u8(fdesc, n);
  F.mvScaleFactors = sf;
  std::vector<MapPoint*> before(n, nullptr);
  for (int i = 0; i < n; ++i) {
    F.mvKeysUn[i].pt = cv::Point2f(xy[2 * i], xy[2 * i + 1]);
    F.mvKeysUn[i].octave = oct[i];
    F.mvKeys[i] = F.mvKeysUn[i];
    if (taken[i]) {
      MapPoint* p = new_mp();
      p->nObs = 1;
      F.mvpMapPoints[i] = before[i] = p;
    }
  }
  fill_grid(F, gs, gi);
  std::vector<MapPoint*> vp(m, nullptr);
  for (int i = 0; i < m; ++i) {
    MapPoint* p = new_mp();
    p->id = i;
    // valid = mbTrackInView && !isBad(): invalid ones alternate between the two reasons
    p->mbTrackInView = valid[i] || (i & 1);
    p->mbBad = !valid[i] && (i & 1);
    p->mTrackProjX = proj[3 * (size_t)i]; p->mTrackProjY = proj[3 * (size_t)i + 1]; p->mTrackProjXR = proj[3 * (size_t)i + 2];
    p->mnTrackScaleLevel = level[i];
    p->mTrackViewCos = vcos[i];
    p->mDescriptor = desc_row(mdesc, i);
    p->nObs = obs[i] ? 3 : 0;
    vp[i] = p;
  }
  ORBmatcher matcher(par[1], true);
  const int nmatches = matcher.SearchByProjection(F, vp, par[0]);
  std::vector<int32_t> out(n + 1);
  for (int i = 0; i < n; ++i) out[i] = (F.mvpMapPoints[i] && F.mvpMapPoints[i] != before[i]) ? F.mvpMapPoints[i]->id : -1;
  out[n] = nmatches;
  save("lp_out", out);
}

// ---- Tracking::TrackReferenceKeyFrame: matcher.SearchByBoW(mpReferenceKF, mCurrentFrame, vpMapPointMatches) ----
static void case_bow() {
  const auto kdesc = load<uint8_t>("bw_kf_desc"), kvalid = load<uint8_t>("bw_kf_valid"), fdesc = load<uint8_t>("bw_f_desc");
  const auto kang = load<float>("bw_kf_angle"), fang = load<float>("bw_f_angle"), par = load<float>("bw_par");  // nnratio, ori
  const int n1 = (int)kvalid.size(), n2 = (int)fang.size();
  KeyFrame KF;
  KF.mvKeysUn.resize(n1); KF.mvuRight.assign(n1, -1.f); KF.mvpMapPoints.assign(n1, nullptr);
  KF.mDescriptors = mat_u8(kdesc, n1);
  for (int i = 0; i < n1; ++i) {
    KF.mvKeysUn[i].angle = kang[i];
    // valid = map point exists and is not bad: invalid ones alternate between NULL and a bad point
    if (kvalid[i] || (i & 1)) {
      MapPoint* p = new_mp();
      p->id = i;
      p->mbBad = !kvalid[i];
      KF.mvpMapPoints[i] = p;
    }
  }
  KF.mFeatVec = featvec(load<int32_t>("bw_kf_nodes"), load<int32_t>("bw_kf_start"), load<int32_t>("bw_kf_idx"));
  Frame F;
  F.N = n2;
  F.mvKeys.resize(n2); F.mvKeysUn.resize(n2); F.mvuRight.assign(n2, -1.f); F.mvpMapPoints.assign(n2, nullptr);
  F.mDescriptors = mat_u8(fdesc, n2);
  for (int i = 0; i < n2; ++i) F.mvKeys[i].angle = F.mvKeysUn[i].angle = fang[i];
  F.mFeatVec = featvec(load<int32_t>("bw_f_nodes"), load<int32_t>("bw_f_start"), load<int32_t>("bw_f_idx"));
  ORBmatcher matcher(par[0], par[1] != 0.f);
  std::vector<MapPoint*> vpMapPointMatches;
  const int nmatches = matcher.SearchByBoW(&KF, F, vpMapPointMatches);
  std::vector<int32_t> out(n2 + 1);
  for (int i = 0; i < n2; ++i) out[i] = vpMapPointMatches[i] ? vpMapPointMatches[i]->id : -1;
  out[n2] = nmatches;
  save("bw_out", out);
}

// ---- Tracking::Relocalization: matcher2.SearchByProjection(mCurrentFrame, vpCandidateKFs[i], sFound, 10, 100) ----
static void case_relocalisation() {
  const auto state = load<uint8_t>("rk_kf_state"), kdesc = load<uint8_t>("rk_kf_desc"), cdesc = load<uint8_t>("rk_cur_desc");
  const auto xyz = load<float>("rk_kf_xyz"), rng = load<float>("rk_kf_dist_range"), kang = load<float>("rk_kf_angle");
  const auto cxy = load<float>("rk_cur_xy"), cang = load<float>("rk_cur_angle"), cam = load<float>("rk_cam"), sf = load<float>("rk_sf");
  const auto coct = load<int32_t>("rk_cur_octave"), gs = load<int32_t>("rk_cur_grid_start"), gi = load<int32_t>("rk_cur_grid_items");
  const auto taken = load<uint8_t>("rk_cur_taken");
  const auto tc = load<float>("rk_tc"), par = load<float>("rk_par");  // th, ORBdist, logScaleFactor
  const int m = (int)state.size(), n2 = (int)taken.size();
  KeyFrame KF;
  KF.mvKeysUn.resize(m); KF.mvpMapPoints.assign(m, nullptr);
  std::set<MapPoint*> sFound;
  for (int i = 0; i < m; ++i) {
    KF.mvKeysUn[i].angle = kang[i];
    if (!state[i]) continue;
    MapPoint* p = new_mp();
    p->id = i;
    p->mbBad = state[i] == 2;
    p->mWorldPos = mat_f(&xyz[3 * (size_t)i], 3, 1);
    p->mDescriptor = desc_row(kdesc, i);
    p->mfMinDistance = rng[2 * i];
    p->mfMaxDistance = rng[2 * i + 1];
    KF.mvpMapPoints[i] = p;
    if (state[i] == 3) sFound.insert(p);
  }
  Frame F;
  F.N = n2;
  Frame::fx = cam[0]; Frame::fy = cam[1]; Frame::cx = cam[2]; Frame::cy = cam[3];
  Frame::mnMinX = cam[4]; Frame::mnMaxX = cam[5]; Frame::mnMinY = cam[6]; Frame::mnMaxY = cam[7];
  Frame::mfGridElementWidthInv = cam[8]; Frame::mfGridElementHeightInv = cam[9];
  F.mvKeys.resize(n2); F.mvKeysUn.resize(n2); F.mvuRight.assign(n2, -1.f); F.mvpMapPoints.assign(n2, nullptr);
  F.mDescriptors = mat_u8(cdesc, n2);
  MapPoint* other = new_mp();
  for (int i = 0; i < n2; ++i) {
    F.mvKeysUn[i].pt = cv::Point2f(cxy[2 * i], cxy[2 * i + 1]);
    F.mvKeysUn[i].octave = coct[i];
    F.mvKeysUn[i].angle = cang[i];
    F.mvKeys[i] = F.mvKeysUn[i];
    if (taken[i]) F.mvpMapPoints[i] = other;
  }
  for (int ix = 0; ix < 64; ++ix)
    for (int iy = 0; iy < 48; ++iy)
      for (int k = gs[ix * 48 + iy]; k < gs[ix * 48 + iy + 1]; ++k) F.mGrid[ix][iy].push_back((size_t)gi[k]);
  float T[16] = {0};
  for (int k = 0; k < 12; ++k) T[k] = tc[k];
  T[15] = 1.f;
  F.mTcw = mat_f(T, 4, 4);
  F.mvScaleFactors = sf;
  F.mnScaleLevels = (int)sf.size();
  F.mfLogScaleFactor = par[2];
  ORBmatcher matcher2(0.9f, true);
  const int nmatches = matcher2.SearchByProjection(F, &KF, sFound, par[0], (int)par[1]);
  std::vector<int32_t> out(n2 + 1);
  for (int i = 0; i < n2; ++i) out[i] = (F.mvpMapPoints[i] && F.mvpMapPoints[i] != other) ? F.mvpMapPoints[i]->id : -1;
  out[n2] = nmatches;
  save("rk_out", out);
}

// ---- LoopClosing::ComputeSim3: matcher.SearchByProjection(mpCurrentKF, mScw, mvpLoopMapPoints, mvpCurrentMatchedPoints, 10) ----
static void case_loop_projection() {
  const auto state = load<uint8_t>("lc_mp_state"), mdesc = load<uint8_t>("lc_mp_desc"), kdesc = load<uint8_t>("lc_kf_desc");
  const auto xyz = load<float>("lc_mp_xyz"), nrm = load<float>("lc_mp_normal"), rng = load<float>("lc_mp_dist_range");
  const auto kxy = load<float>("lc_kf_xy"), cam4 = load<float>("lc_kf_cam4"), sf = load<float>("lc_kf_scale_factors"), scw = load<float>("lc_scw");
  const auto koct = load<int32_t>("lc_kf_octave"), gs = load<int32_t>("lc_kf_grid_start"), gi = load<int32_t>("lc_kf_grid_items"),
             bounds = load<int32_t>("lc_kf_bounds4"), mi = load<int32_t>("lc_matched_in");
  const auto par = load<float>("lc_par");  // th, gwi, ghi, logScaleFactor
  const int m = (int)state.size(), n = (int)koct.size();
  KeyFrame KF;
  KF.mvKeysUn.resize(n);
  for (int i = 0; i < n; ++i) {
    KF.mvKeysUn[i].pt = cv::Point2f(kxy[2 * i], kxy[2 * i + 1]);
    KF.mvKeysUn[i].octave = koct[i];
  }
  KF.mDescriptors = mat_u8(kdesc, n);
  KF.mvScaleFactors = sf;
  KF.mnScaleLevels = (int)sf.size();
  KF.mfLogScaleFactor = par[3];
  KF.fx = cam4[0]; KF.fy = cam4[1]; KF.cx = cam4[2]; KF.cy = cam4[3];
  KF.mnMinX = bounds[0]; KF.mnMinY = bounds[1]; KF.mnMaxX = bounds[2]; KF.mnMaxY = bounds[3];
  KF.mfGridElementWidthInv = par[1]; KF.mfGridElementHeightInv = par[2];
  KF.mGrid.assign(64, std::vector<std::vector<size_t> >(48));
  for (int ix = 0; ix < 64; ++ix)
    for (int iy = 0; iy < 48; ++iy)
      for (int k = gs[ix * 48 + iy]; k < gs[ix * 48 + iy + 1]; ++k) KF.mGrid[ix][iy].push_back((size_t)gi[k]);
  std::vector<MapPoint*> vpPoints(m);
  for (int i = 0; i < m; ++i) {
    MapPoint* p = new_mp();
    p->id = i;
    p->mbBad = state[i] == 2;
    p->mWorldPos = mat_f(&xyz[3 * (size_t)i], 3, 1);
    p->mNormalVector = mat_f(&nrm[3 * (size_t)i], 3, 1);
    p->mDescriptor = desc_row(mdesc, i);
    p->mfMinDistance = rng[2 * i];
    p->mfMaxDistance = rng[2 * i + 1];
    vpPoints[i] = p;
  }
  std::vector<MapPoint*> vpMatched(n, nullptr);
  for (int i = 0; i < n; ++i)
    if (mi[i] >= 0) vpMatched[i] = vpPoints[mi[i]];
  float S[16] = {0};
  for (int k = 0; k < 12; ++k) S[k] = scw[k];
  S[15] = 1.f;
  ORBmatcher matcher(0.75f, true);
  const int nmatches = matcher.SearchByProjection(&KF, mat_f(S, 4, 4), vpPoints, vpMatched, (int)par[0]);
  std::vector<int32_t> out(n + 1);
  for (int i = 0; i < n; ++i) out[i] = (vpMatched[i] && mi[i] < 0) ? vpMatched[i]->id : -1;
  out[n] = nmatches;
  save("lc_out", out);
}

// ---- LocalMapping::SearchInNeighbors: matcher.Fuse(pKFi, vpMapPointMatches) / LoopClosing::SearchAndFuse: matcher.Fuse(pKF, cvScw,
//      mvpLoopMapPoints, 4, vpReplacePoints) ----
static void case_fuse(const std::string& pfx, bool sim3) {
  const auto state = load<uint8_t>(pfx + "mp_state"), mdesc = load<uint8_t>(pfx + "mp_desc"), kdesc = load<uint8_t>(pfx + "kf_desc");
  const auto xyz = load<float>(pfx + "mp_xyz"), nrm = load<float>(pfx + "mp_normal"), rng = load<float>(pfx + "mp_dist_range");
  const auto kxy = load<float>(pfx + "kf_xy"), cam4 = load<float>(pfx + "kf_cam4"), sf = load<float>(pfx + "kf_scale_factors");
  const auto ur = load<float>(pfx + "kf_uright"), inv = load<float>(pfx + "kf_inv_level_sigma2"), tcw = load<float>(pfx + "kf_tcw"),
             ow = load<float>(pfx + "kf_ow");
  const auto koct = load<int32_t>(pfx + "kf_octave"), gs = load<int32_t>(pfx + "kf_grid_start"), gi = load<int32_t>(pfx + "kf_grid_items"),
             bounds = load<int32_t>(pfx + "kf_bounds4"), mnobs = load<int32_t>(pfx + "mp_nobs"), knobs = load<int32_t>(pfx + "kp_nobs");
  const auto khas = load<uint8_t>(pfx + "kp_has"), kbad = load<uint8_t>(pfx + "kp_bad");
  const auto par = load<float>(pfx + "par");  // th, gwi, ghi, logScaleFactor, mbf
  const int m = (int)state.size(), n = (int)koct.size();
  KeyFrame KF;
  g_fuse_kf = &KF;
  g_log.clear();
  KF.mvKeysUn.resize(n);
  for (int i = 0; i < n; ++i) {
    KF.mvKeysUn[i].pt = cv::Point2f(kxy[2 * i], kxy[2 * i + 1]);
    KF.mvKeysUn[i].octave = koct[i];
  }
  KF.mDescriptors = mat_u8(kdesc, n);
  KF.mvScaleFactors = sf; KF.mvInvLevelSigma2 = inv; KF.mvuRight = ur; KF.mbf = par[4];
  KF.mnScaleLevels = (int)sf.size(); KF.mfLogScaleFactor = par[3];
  KF.fx = cam4[0]; KF.fy = cam4[1]; KF.cx = cam4[2]; KF.cy = cam4[3];
  KF.mnMinX = bounds[0]; KF.mnMinY = bounds[1]; KF.mnMaxX = bounds[2]; KF.mnMaxY = bounds[3];
  KF.mfGridElementWidthInv = par[1]; KF.mfGridElementHeightInv = par[2];
  KF.mGrid.assign(64, std::vector<std::vector<size_t> >(48));
  for (int ix = 0; ix < 64; ++ix)
    for (int iy = 0; iy < 48; ++iy)
      for (int k = gs[ix * 48 + iy]; k < gs[ix * 48 + iy + 1]; ++k) KF.mGrid[ix][iy].push_back((size_t)gi[k]);
  float R9[9], t3[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) R9[3 * r + c] = tcw[4 * r + c];
    t3[r] = tcw[4 * r + 3];
  }
  KF.R = mat_f(R9, 3, 3); KF.t = mat_f(t3, 3, 1); KF.Ow = mat_f(ow.data(), 3, 1);
  KF.mvpMapPoints.assign(n, nullptr);
  for (int i = 0; i < n; ++i)
    if (khas[i]) {
      MapPoint* p = new_mp();
      p->id = m + i; p->mbBad = kbad[i] != 0; p->nObs = knobs[i];
      KF.mvpMapPoints[i] = p;
    }
  std::vector<int32_t> alias;
  if (sim3) alias = load<int32_t>(pfx + "mp_alias");
  std::vector<MapPoint*> vp(m, nullptr);
  for (int i = 0; i < m; ++i) {
    if (state[i] == 0) continue;
    if (state[i] == 4) { vp[i] = KF.mvpMapPoints[alias[i]]; continue; }
    MapPoint* p = new_mp();
    p->id = i; p->mbBad = state[i] == 2; p->inKF = state[i] == 3; p->nObs = mnobs[i];
    p->mWorldPos = mat_f(&xyz[3 * (size_t)i], 3, 1);
    p->mNormalVector = mat_f(&nrm[3 * (size_t)i], 3, 1);
    p->mDescriptor = desc_row(mdesc, i);
    p->mfMinDistance = rng[2 * i]; p->mfMaxDistance = rng[2 * i + 1];
    vp[i] = p;
  }
  ORBmatcher matcher(0.6f, true);
  int nFused;
  std::vector<int32_t> out;
  if (sim3) {
    const auto scw = load<float>(pfx + "kf_scw");
    float S[16] = {0};
    for (int k = 0; k < 12; ++k) S[k] = scw[k];
    S[15] = 1.f;
    std::vector<MapPoint*> vpReplace(m, nullptr);
    nFused = matcher.Fuse(&KF, mat_f(S, 4, 4), vp, par[0], vpReplace);
    for (int i = 0; i < m; ++i) out.push_back(vpReplace[i] ? vpReplace[i]->id : -1);
  } else {
    nFused = matcher.Fuse(&KF, vp, par[0]);
  }
  out.push_back(nFused);
  save(pfx + "out", out);
  save(pfx + "log", g_log);
}

// ---- LoopClosing::ComputeSim3: matcher.SearchBySim3(mpCurrentKF, pKF, vpMapPointMatches, s, R, t, 7.5) ----
static void case_sim3() {
  const auto par = load<float>("s3_par");  // th, gwi, ghi, logScaleFactor, s12
  const auto R12 = load<float>("s3_R12"), t12 = load<float>("s3_t12");
  const auto mi = load<int32_t>("s3_matched_in");
  KeyFrame K[2];
  for (int s = 0; s < 2; ++s) {
    const std::string p = s ? "s3_kf2_" : "s3_kf1_", q = s ? "s3_mp2_" : "s3_mp1_";
    const auto kdesc = load<uint8_t>(p + "desc"), state = load<uint8_t>(q + "state"), mdesc = load<uint8_t>(q + "desc");
    const auto kxy = load<float>(p + "xy"), cam4 = load<float>(p + "cam4"), sf = load<float>(p + "scale_factors"), tcw = load<float>(p + "tcw");
    const auto xyz = load<float>(q + "xyz"), rng = load<float>(q + "dist_range");
    const auto koct = load<int32_t>(p + "octave"), gs = load<int32_t>(p + "grid_start"), gi = load<int32_t>(p + "grid_items"),
               bounds = load<int32_t>(p + "bounds4");
    const int n = (int)koct.size();
    KeyFrame& KF = K[s];
    KF.mvKeysUn.resize(n);
    for (int i = 0; i < n; ++i) {
      KF.mvKeysUn[i].pt = cv::Point2f(kxy[2 * i], kxy[2 * i + 1]);
      KF.mvKeysUn[i].octave = koct[i];
    }
    KF.mDescriptors = mat_u8(kdesc, n);
    KF.mvScaleFactors = sf; KF.mnScaleLevels = (int)sf.size(); KF.mfLogScaleFactor = par[3];
    KF.fx = cam4[0]; KF.fy = cam4[1]; KF.cx = cam4[2]; KF.cy = cam4[3];
    KF.mnMinX = bounds[0]; KF.mnMinY = bounds[1]; KF.mnMaxX = bounds[2]; KF.mnMaxY = bounds[3];
    KF.mfGridElementWidthInv = par[1]; KF.mfGridElementHeightInv = par[2];
    KF.mGrid.assign(64, std::vector<std::vector<size_t> >(48));
    for (int ix = 0; ix < 64; ++ix)
      for (int iy = 0; iy < 48; ++iy)
        for (int k = gs[ix * 48 + iy]; k < gs[ix * 48 + iy + 1]; ++k) KF.mGrid[ix][iy].push_back((size_t)gi[k]);
    float R9[9], t3[3];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) R9[3 * r + c] = tcw[4 * r + c];
      t3[r] = tcw[4 * r + 3];
    }
    KF.R = mat_f(R9, 3, 3); KF.t = mat_f(t3, 3, 1);
    KF.mvpMapPoints.assign(n, nullptr);
    for (int i = 0; i < n; ++i) {
      if (!state[i]) continue;
      MapPoint* mp = new_mp();
      mp->id = i; mp->mbBad = state[i] == 2;
      mp->mWorldPos = mat_f(&xyz[3 * (size_t)i], 3, 1);
      mp->mDescriptor = desc_row(mdesc, i);
      mp->mfMinDistance = rng[2 * i]; mp->mfMaxDistance = rng[2 * i + 1];
      if (s) mp->idxInKF2 = i;
      KF.mvpMapPoints[i] = mp;
    }
  }
  const int n1 = (int)K[0].mvKeysUn.size();
  std::vector<MapPoint*> vpMatches12(n1, nullptr);
  for (int i = 0; i < n1; ++i)
    if (mi[i] >= 0) vpMatches12[i] = K[1].mvpMapPoints[mi[i]];
  ORBmatcher matcher(0.75f, true);
  const float s12 = par[4];
  const int nFound = matcher.SearchBySim3(&K[0], &K[1], vpMatches12, s12, mat_f(R12.data(), 3, 3), mat_f(t12.data(), 3, 1), par[0]);
  std::vector<int32_t> out(n1 + 1);
  for (int i = 0; i < n1; ++i) out[i] = (vpMatches12[i] && mi[i] < 0) ? vpMatches12[i]->id : -1;
  out[n1] = nFound;
  save("s3_out", out);
}

// ---- LoopClosing::ComputeSim3: matcher.SearchByBoW(mpCurrentKF, pKF, vvpMapPointMatches[i]) ----
static void case_bow_keyframes() {
  const auto par = load<float>("bk_par");  // nnratio, ori
  KeyFrame K[2];
  for (int s = 0; s < 2; ++s) {
    const std::string p = s ? "bk_kf2_" : "bk_kf1_";
    const auto desc = load<uint8_t>(p + "desc"), valid = load<uint8_t>(p + "valid");
    const auto ang = load<float>(p + "angle");
    const int n = (int)valid.size();
    K[s].mvKeysUn.resize(n); K[s].mvuRight.assign(n, -1.f); K[s].mvpMapPoints.assign(n, nullptr);
    K[s].mDescriptors = mat_u8(desc, n);
    for (int i = 0; i < n; ++i) {
      K[s].mvKeysUn[i].angle = ang[i];
      if (valid[i] || (i & 1)) {  // invalid = NULL or bad, alternating
        MapPoint* mp = new_mp();
        mp->id = i;
        mp->mbBad = !valid[i];
        K[s].mvpMapPoints[i] = mp;
      }
    }
    K[s].mFeatVec = featvec(load<int32_t>(p + "nodes"), load<int32_t>(p + "start"), load<int32_t>(p + "idx"));
  }
  ORBmatcher matcher(par[0], par[1] != 0.f);
  std::vector<MapPoint*> vpMatches12;
  const int nmatches = matcher.SearchByBoW(&K[0], &K[1], vpMatches12);
  const int n1 = (int)K[0].mvKeysUn.size();
  std::vector<int32_t> out(n1 + 1);
  for (int i = 0; i < n1; ++i) out[i] = vpMatches12[i] ? vpMatches12[i]->id : -1;
  out[n1] = nmatches;
  save("bk_out", out);
}

// ---- LocalMapping::CreateNewMapPoints: matcher.SearchForTriangulation(mpCurrentKeyFrame, pKF2, F12, vMatchedIndices, false) ----
static void case_triangulation() {
  const auto par = load<float>("tr_par");  // fx fy cx cy only_stereo ori
  const auto F12 = load<float>("tr_F12"), R2w = load<float>("tr_R2w"), t2w = load<float>("tr_t2w"), Cw = load<float>("tr_Cw");
  const auto sf = load<float>("tr_sf"), sg = load<float>("tr_sg");
  KeyFrame K[2];
  for (int s = 0; s < 2; ++s) {
    const std::string p = s ? "tr_kf2_" : "tr_kf1_";
    const auto desc = load<uint8_t>(p + "desc"), has = load<uint8_t>(p + "has_mp");
    const auto xy = load<float>(p + "xy"), ang = load<float>(p + "angle"), ur = load<float>(p + "uright");
    const int n = (int)has.size();
    K[s].mvKeysUn.resize(n);
    K[s].mvuRight = ur;
    K[s].mDescriptors = mat_u8(desc, n);
    K[s].mvpMapPoints.assign(n, nullptr);
    std::vector<int32_t> oct;
    if (s) oct = load<int32_t>(p + "octave");
    for (int i = 0; i < n; ++i) {
      K[s].mvKeysUn[i].pt = cv::Point2f(xy[2 * i], xy[2 * i + 1]);
      K[s].mvKeysUn[i].angle = ang[i];
      if (s) K[s].mvKeysUn[i].octave = oct[i];
      if (has[i]) K[s].mvpMapPoints[i] = new_mp();
    }
    K[s].mFeatVec = featvec(load<int32_t>(p + "nodes"), load<int32_t>(p + "start"), load<int32_t>(p + "idx"));
    K[s].fx = par[0]; K[s].fy = par[1]; K[s].cx = par[2]; K[s].cy = par[3];
    K[s].mvScaleFactors = sf;
    K[s].mvLevelSigma2 = sg;
  }
  K[0].Ow = mat_f(Cw.data(), 3, 1);
  K[1].R = mat_f(R2w.data(), 3, 3);
  K[1].t = mat_f(t2w.data(), 3, 1);
  ORBmatcher matcher(0.6f, par[5] != 0.f);
  std::vector<std::pair<size_t, size_t> > vMatchedIndices;
  const int nmatches = matcher.SearchForTriangulation(&K[0], &K[1], mat_f(F12.data(), 3, 3), vMatchedIndices, par[4] != 0.f);
  std::vector<int32_t> out;
  for (auto& pr : vMatchedIndices) {
    out.push_back((int32_t)pr.first);
    out.push_back((int32_t)pr.second);
  }
  out.push_back(nmatches);
  save("tr_out", out);
}

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <dir>\n", argv[0]);
    return 2;
  }
  g_dir = argv[1];
  try {
    case_projection();
    case_local();
    case_bow();
    case_bow_keyframes();
    case_relocalisation();
    case_loop_projection();
    case_fuse("fu_", false);
    case_fuse("fs_", true);
    case_sim3();
    case_triangulation();
    // an unfilled member must be reported, not read out of bounds
    Frame bad, last;
    bad.mvKeysUn.resize(4);
    bool thrown = false;
    try {
      ORBmatcher(0.9f, true).SearchByProjection(bad, last, 7.f, false);
    } catch (const std::invalid_argument&) {
      thrown = true;
    }
    if (!thrown) {
      std::fprintf(stderr, "missing length check\n");
      return 1;
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  for (MapPoint* p : g_pool) delete p;
  std::printf("ok\n");
  return 0;
}
