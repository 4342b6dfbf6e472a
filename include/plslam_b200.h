/* plslam_b200.h — C-ABI of the B200-native RGBD-PL-SLAM front-end (ORB + LSD/LBD + matchers).
 *
 * This is the drop-in boundary: plain pointers, sizes, int status codes, opaque handles and an
 * explicit cudaStream_t (passed as void*).  No torch / OpenCV types appear here.  The C++ classes
 * in rgbd-pl-slam_b200/host/ (ORB_SLAM2::ORBextractor, LineSegment, ORBmatcher, LSDmatcher) are
 * thin veneers over these entry points with the reference's exact signatures.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference
 * repo maxee1900/RGBD-PL-SLAM; "@0x…" = address in its prebuilt lib/libORB_SLAM2.so, the only
 * form in which the reference ships ORBextractor / ORBmatcher).
 *
 * There is NO CPU fallback: every compute entry point returns PLSLAM_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef PLSLAM_B200_H
#define PLSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLSLAM_OK 0
#define PLSLAM_ERR_INVALID 1   /* bad argument */
#define PLSLAM_ERR_CUDA 2      /* CUDA runtime failure (see plslam_last_error) */
#define PLSLAM_ERR_CAPACITY 3  /* caller-provided output capacity too small */
#define PLSLAM_ERR_OVERFLOW 4  /* an internal fixed-capacity buffer overflowed (reported, never truncated silently) */
#define PLSLAM_ERR_INTERNAL 5  /* a device-side watchdog fired (speculative region-growing scheduler made no progress) */

/* cv::KeyPoint layout (28 B; stores @0x7656d-0x7659d): pt.x, pt.y, size, angle, response, octave, class_id */
typedef struct plslam_keypoint {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} plslam_keypoint_t;

/* cv::line_descriptor::KeyLine layout (OpenCV-contrib, used by include/ExtractLineSegment.h:38) */
typedef struct plslam_keyline {
  float angle;
  int32_t class_id;
  int32_t octave;
  float pt_x, pt_y;
  float response;
  float size;
  float startPointX, startPointY, endPointX, endPointY;
  float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
  float lineLength;
  int32_t numOfPixels;
} plslam_keyline_t;

const char* plslam_last_error(void);      /* thread-local message of the last failing call */
const char* plslam_version(void);
int plslam_device_count(void);            /* number of usable CUDA devices (0 => every compute call fails) */

/* ------------------------------------------------------------------------------------------
 * ORB extractor  — replaces ORB_SLAM2::ORBextractor (include/ORBextractor.h:45-111)
 * ---------------------------------------------------------------------------------------- */
typedef struct plslam_orb plslam_orb_t;

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (ORBextractor.h:51-52, @0x73050).  Binds to the current CUDA device. */
int plslam_orb_create(plslam_orb_t** out, int nfeatures, float scaleFactor, int nlevels, int iniThFAST,
                      int minThFAST);
void plslam_orb_destroy(plslam_orb_t* h);

/* 7-tap integer Gaussian table (sum 256) used for the 7x7 sigma=2 blur (@0x77487).  Default is
 * OpenCV 4.13's {18,34,48,56,48,34,18}; OpenCV 3.3 builds may want {18,34,49,55,49,34,18}. */
int plslam_orb_set_blur_kernel(plslam_orb_t* h, const int32_t k[7]);

/* GetLevels / GetScaleFactor(s) / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (ORBextractor.h:63-83) plus the per-level quotas and umax table.
 * Any output pointer may be NULL.  Arrays hold nlevels entries (umax: 16). */
int plslam_orb_levels(const plslam_orb_t* h);
int plslam_orb_tables(const plslam_orb_t* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                      int32_t* features_per_level, int32_t* umax16);
/* Upper bound of keypoints per frame: sum over levels of (quota + 3) (DistributeOctTree may overshoot by 3). */
int plslam_orb_max_keypoints(const plslam_orb_t* h);

/* ORBextractor::operator()(image, mask, keypoints, descriptors) (ORBextractor.h:59-61, @0x76da0)
 * on ONE host image (CV_8UC1, `pitch` bytes per row).  Keypoints are written level-major, in
 * the reference's order, coordinates rescaled to level 0; descriptors are n x 32 bytes row-major.
 * An empty image (NULL / zero size) returns PLSLAM_OK with *n_out = 0 (@0x76dda). */
int plslam_orb_extract(plslam_orb_t* h, const uint8_t* image, int width, int height, int pitch,
                       plslam_keypoint_t* keypoints, uint8_t* descriptors, int capacity, int* n_out);

/* Batched extension (additive; the reference is single-frame): `batch` frames of identical size in
 * host memory, `frame_stride` bytes apart.  Output for frame f starts at keypoints + f*capacity
 * and descriptors + f*capacity*32; counts[f] receives its keypoint count.  Host<->device copies
 * happen inside the call. */
int plslam_orb_extract_batch_host(plslam_orb_t* h, const uint8_t* images, int batch, int width, int height,
                                  int pitch, size_t frame_stride, plslam_keypoint_t* keypoints,
                                  uint8_t* descriptors, int capacity, int32_t* counts);

/* Same with device-resident inputs and outputs; asynchronous on `stream` (a cudaStream_t).
 * capacity must be >= plslam_orb_max_keypoints().  d_status (device int32, may be NULL) receives
 * PLSLAM_OK or PLSLAM_ERR_OVERFLOW. */
int plslam_orb_extract_batch_device(plslam_orb_t* h, const uint8_t* d_images, int batch, int width, int height,
                                    int pitch, size_t frame_stride, plslam_keypoint_t* d_keypoints,
                                    uint8_t* d_descriptors, int capacity, int32_t* d_counts, void* stream);

/* ORBextractor::mvImagePyramid (public member, ORBextractor.h:85): copy level `level` of frame `frame`
 * of the last batch into `out` (dense rows, width*height bytes).  which: 0 = pyramid level,
 * 1 = 7x7-blurred level (the reference's workingMat, @0x773ff-0x77487). */
int plslam_orb_level_size(const plslam_orb_t* h, int level, int* width, int* height);
int plslam_orb_copy_level(plslam_orb_t* h, int frame, int level, int which, uint8_t* out, size_t out_bytes);

/* Parity/debug: FAST candidates of (frame, level) of the last batch before DistributeOctTree, as
 * (x, y, response) int32 triples in border-local coordinates, sorted in the reference's list order
 * (cell row, cell col, y, x) (@0x765b8-0x765e3).  Returns the count in *n_out. */
int plslam_orb_copy_candidates(plslam_orb_t* h, int frame, int level, int32_t* xyr, int capacity, int* n_out);

/* ------------------------------------------------------------------------------------------
 * Line extractor — replaces ORB_SLAM2::LineSegment (include/ExtractLineSegment.h:30-55)
 * ---------------------------------------------------------------------------------------- */
typedef struct plslam_lines plslam_lines_t;

/* LineSegment::LineSegment() (ExtractLineSegment.h:33).  Binds to the current CUDA device. */
int plslam_lines_create(plslam_lines_t** out);
void plslam_lines_destroy(plslam_lines_t* h);

/* Number of strongest lines (by KeyLine::response, comparator include/auxiliar.h:67-72) kept per
 * frame before LBD; the PL-SLAM fork family hard-codes 40.  0 keeps every detected line. */
int plslam_lines_set_max_lines(plslam_lines_t* h, int max_lines);
/* Per-frame output capacity the batched device entry point requires. */
int plslam_lines_capacity(const plslam_lines_t* h);

/* LineSegment::ExtractLineSegment(img, keylines, ldesc, keylineFunctions, scale=1, numOctaves=1)
 * (ExtractLineSegment.h:38; scale/numOctaves are the header defaults: one full-resolution octave) on
 * ONE host image.  keylines: n x 68 B (cv::line_descriptor::KeyLine layout), descriptors: n x 32 B
 * LBD rows (CV_8U), line_functions: n x 3 doubles (Eigen::Vector3d, l = sp x ep / |(l0,l1)|). */
int plslam_lines_extract(plslam_lines_t* h, const uint8_t* image, int width, int height, int pitch,
                         plslam_keyline_t* keylines, uint8_t* descriptors, double* line_functions, int capacity,
                         int* n_out);

/* Batched extensions (additive), same conventions as the ORB batch entry points. */
int plslam_lines_extract_batch_host(plslam_lines_t* h, const uint8_t* images, int batch, int width, int height,
                                    int pitch, size_t frame_stride, plslam_keyline_t* keylines, uint8_t* descriptors,
                                    double* line_functions, int capacity, int32_t* counts);
int plslam_lines_extract_batch_device(plslam_lines_t* h, const uint8_t* d_images, int batch, int width, int height,
                                      int pitch, size_t frame_stride, plslam_keyline_t* d_keylines,
                                      uint8_t* d_descriptors, double* d_line_functions, int capacity,
                                      int32_t* d_counts, void* stream);
/* Synchronises `stream` and returns PLSLAM_ERR_OVERFLOW if an internal list overflowed in the last batch. */
int plslam_lines_check_status(plslam_lines_t* h, void* stream);
int plslam_orb_check_status(plslam_orb_t* h, void* stream);

/* Parity/debug accessors on the last batch: the x0.8 scaled image LSD works on, the level-line
 * field (degrees, -1024 = undefined; gx^2+gy^2), and every accepted LSD segment in detection order
 * as 5 doubles x1,y1,x2,y2 (float32 values),width + prec + log-NFA = 7 doubles per segment. */
int plslam_lines_scaled_size(const plslam_lines_t* h, int* width, int* height);
int plslam_lines_copy_scaled(plslam_lines_t* h, int frame, uint8_t* out, size_t out_bytes);
int plslam_lines_copy_level_lines(plslam_lines_t* h, int frame, float* degrees, int32_t* grad2, size_t count);
int plslam_lines_copy_segments(plslam_lines_t* h, int frame, double* seg7, int capacity, int* n_out);
/* BinaryDescriptor::compute(image, keylines, descriptors) (OpenCV-contrib line_descriptor; the second half of
 * LineSegment::ExtractLineSegment, include/ExtractLineSegment.h:38): the 32 LBD bytes of each of n GIVEN key lines of a
 * host image (only the KeyLine fields the descriptor reads matter: sPointInOctave / ePointInOctave, angle, numOfPixels).
 * Host pointers; synchronous. */
int plslam_lines_compute_lbd(plslam_lines_t* h, const uint8_t* image, int width, int height, int pitch,
                             const plslam_keyline_t* keylines, int n, uint8_t* descriptors);
/* Profiling aid: reads and clears the 32 device-side counters of the speculative region-growing scheduler
 * (k_lsd_grow_aw: regions issued / void / squashed while running / squashed after finishing / inserted, rectangles,
 * validation chunks, scheduler idle polls, worker idle polls, blocked seeds, pick chunks, frames, then cycle
 * accounting; names in tools/prof_aw.py).  Synchronises the device.  Counters never influence a result. */
int plslam_debug_grow_stats(unsigned long long* out32);

/* ------------------------------------------------------------------------------------------
 * Matchers — Hamming cores of ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:37-141) and
 * ORB_SLAM2::LSDmatcher / LineSegment::LineSegmentMathch (include/LSDmatcher.h:25-78,
 * include/ExtractLineSegment.h:41).  Pointer-based containers (MapPoint*, DBoW2::FeatureVector,
 * Frame::mGrid) are passed flattened; all pointers inside a job are DEVICE pointers.
 * ---------------------------------------------------------------------------------------- */
#define PLSLAM_TH_LOW 50        /* ORBmatcher::TH_LOW       (@0x1269e4) */
#define PLSLAM_TH_HIGH 100      /* ORBmatcher::TH_HIGH      (@0x1269e8) */
#define PLSLAM_HISTO_LENGTH 30  /* ORBmatcher::HISTO_LENGTH (@0x1269e0) */
#define PLSLAM_GRID_COLS 64     /* FRAME_GRID_COLS (include/Frame.h:42) */
#define PLSLAM_GRID_ROWS 48     /* FRAME_GRID_ROWS (include/Frame.h:41) */

/* ORBmatcher::DescriptorDistance(a, b) (ORBmatcher.h:44, @0x79d20; FORB.cpp:82-102) on two 32-byte
 * host rows.  Scalar, host-side: one pair is 8 popcounts; the batched forms below are the GPU path. */
int plslam_descriptor_distance(const uint8_t* a, const uint8_t* b);

/* cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, k=2) — what LineSegmentMathch / LSDmatcher run on
 * LBD rows (include/auxiliar.h:30-51) and the all-pairs ORB case.  out: nq x 4 int32
 * (trainIdx1, dist1, trainIdx2, dist2), -1 where fewer than 2 train rows exist; ties -> lower index. */
typedef struct plslam_knn_job {
  const uint8_t* query; /* nq x 32 */
  const uint8_t* train; /* nt x 32 */
  int32_t* out;         /* nq x 4 */
  int32_t nq, nt;
} plslam_knn_job_t;
int plslam_match_knn2_batch_device(const plslam_knn_job_t* d_jobs, int njobs, int max_nq, void* stream);
int plslam_match_knn2_host(const uint8_t* query, int nq, const uint8_t* train, int nt, int32_t* out);

/* ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& matches) (ORBmatcher.h:104, @0x80150).
 * FeatureVectors (std::map<NodeId, vector<unsigned>>) as CSR sorted by node id. */
typedef struct plslam_bow_job {
  const uint8_t* kf_desc;   /* N1 x 32 : pKF->mDescriptors */
  const float* kf_angle;    /* N1      : pKF->mvKeysUn[i].angle */
  const uint8_t* kf_valid;  /* N1      : pMP && !pMP->isBad() */
  const int32_t* kf_nodes;  /* n_kf_nodes   : FeatureVector keys, ascending */
  const int32_t* kf_start;  /* n_kf_nodes+1 : offsets into kf_idx */
  const int32_t* kf_idx;    /* feature indices per node */
  const uint8_t* f_desc;    /* N2 x 32 : F.mDescriptors */
  const float* f_angle;     /* N2      : F.mvKeys[i].angle */
  const int32_t* f_nodes;
  const int32_t* f_start;
  const int32_t* f_idx;
  int32_t* match_f;         /* N2 : KF feature index matched to each F feature, -1 = none (vpMapPointMatches) */
  int32_t* nmatches;        /* 1  : return value */
  int32_t n1, n2, n_kf_nodes, n_f_nodes;
  float nnratio;            /* mfNNratio */
  int32_t check_orientation;/* mbCheckOrientation */
  /* ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) (ORBmatcher.h:105, @0x82cc0)
   * is the same walk with the second key frame in the role of F: its features need a good map point too (f_valid, NULL
   * = all valid) and the distance test is strict (strict_low: bestDist1 < TH_LOW, cmpl $0x31 @0x83490).  match_f is
   * then indexed by KF2; vpMatches12[match_f[i2]] = KF2's map point i2 (plslam_match_bow_kfkf_host inverts it). */
  const uint8_t* f_valid;
  int32_t strict_low;
} plslam_bow_job_t;
int plslam_match_bow_batch_device(const plslam_bow_job_t* d_jobs, int njobs, int max_n, void* stream); /* max_n >= every job's n1 and n2 */

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, float th, bool bMono)
 * (ORBmatcher.h:78, @0x80d00) including Frame::GetFeaturesInArea (Frame.h:113). */
typedef struct plslam_proj_job {
  /* LastFrame */
  const uint8_t* last_valid;   /* N1 : mvpMapPoints[i] && !mvbOutlier[i] */
  const float* last_xyz;       /* N1 x 3 : pMP->GetWorldPos() */
  const uint8_t* last_desc;    /* N1 x 32 : pMP->GetDescriptor() */
  const int32_t* last_octave;  /* N1 : mvKeys[i].octave */
  const float* last_angle;     /* N1 : mvKeysUn[i].angle */
  const uint8_t* last_obs;     /* N1 : pMP->Observations() > 0 */
  /* CurrentFrame */
  const float* cur_xy;         /* N2 x 2 : mvKeysUn[i].pt */
  const int32_t* cur_octave;   /* N2 */
  const float* cur_angle;      /* N2 */
  const uint8_t* cur_desc;     /* N2 x 32 */
  const float* cur_uright;     /* N2 : mvuRight */
  const uint8_t* cur_taken;    /* N2 : mvpMapPoints[i] && Observations() > 0 on entry */
  const int32_t* grid_start;   /* 64*48+1 : CSR of mGrid in [ix][iy] order */
  const int32_t* grid_items;
  const float* scale_factors;  /* mvScaleFactors */
  int32_t* match_cur;          /* N2 : LastFrame index assigned to each current keypoint, -1 = none (and, with
                                  report_removed, -2 = assigned during the scan and taken away again by the rotation
                                  check: the reference leaves CurrentFrame.mvpMapPoints[i] NULL there) */
  int32_t* nmatches;           /* 1 */
  float cam[12];               /* fx, fy, cx, cy, mbf, mb, mnMinX, mnMaxX, mnMinY, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv */
  float tcw_cur[12];           /* CurrentFrame.mTcw rows 0..2 (3x4 row-major) */
  float tcw_last[12];          /* LastFrame.mTcw */
  float th;
  int32_t n1, n2, mono, check_orientation;
  int32_t report_removed;      /* 0: removed entries read -1 like never-assigned ones; 1: they read -2 */
  /* mode 1 = ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, float th,
   * int ORBdist) (ORBmatcher.h:82, @0x7e8c0; Tracking::Relocalization): the "last" arrays describe the KEY FRAME's map points
   * (last_valid = pMP && !isBad() && !sAlreadyFound.count(pMP); last_angle = pKF->mvKeysUn[i].angle; last_octave, last_obs,
   * cur_uright, tcw_last, mono unused), last_dist_range = (mfMinDistance, mfMaxDistance) per point, cur_taken =
   * CurrentFrame.mvpMapPoints[i] != NULL.  The search level is MapPoint::PredictScale(dist3D, &CurrentFrame) (@0x8fc20) from
   * log_scale_factor = mfLogScaleFactor and n_levels = mnScaleLevels, the window spans levels level-1 .. level+1, every
   * assigned key point is closed to later points, and the acceptance threshold is orb_dist.  mode 0 = the frame / frame form. */
  const float* last_dist_range; /* N1 x 2, mode 1 only; NULL = the caller made the distance test and passes the predicted levels
                                   (pMP->PredictScale(dist3D, &CurrentFrame)) in last_octave */
  int32_t mode, orb_dist, n_levels;
  float log_scale_factor;
} plslam_proj_job_t;
int plslam_match_projection_batch_device(const plslam_proj_job_t* d_jobs, int njobs, int max_n1, int max_n2, void* stream);

/* ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched,
 * int th) (ORBmatcher.h:86, @0x880f0; LoopClosing::ComputeSim3 / SearchAndFuse) including KeyFrame::GetFeaturesInArea (@0x96fe0),
 * KeyFrame::IsInImage (@0x97480) and MapPoint::PredictScale(dist, pKF) (@0x8fb60): map points projected into a key frame with a
 * similarity transform.  One warp per job walks the points in order (a key-frame feature assigned to an earlier point is closed
 * to later ones). */
typedef struct plslam_kfproj_job {
  /* vpPoints */
  const uint8_t* mp_valid;      /* M : !pMP->isBad() && pMP is not in vpMatched on entry */
  const float* mp_xyz;          /* M x 3 : GetWorldPos() */
  const float* mp_normal;       /* M x 3 : GetNormal() */
  const float* mp_dist_range;   /* M x 2 : mfMinDistance, mfMaxDistance */
  const uint8_t* mp_desc;       /* M x 32 : GetDescriptor() */
  const int32_t* mp_level;      /* M or NULL.  Not NULL: the caller made the scale-invariance and viewing-angle tests itself
                                   (folded into mp_valid) and passes pMP->PredictScale(dist, pKF); mp_normal / mp_dist_range unused */
  /* pKF */
  const float* kf_xy;           /* N x 2 : mvKeysUn[i].pt */
  const int32_t* kf_octave;     /* N */
  const uint8_t* kf_desc;       /* N x 32 */
  const uint8_t* kf_matched;    /* N : vpMatched[i] != NULL on entry */
  const int32_t* grid_start;    /* grid_cols * grid_rows + 1 : CSR of KeyFrame::mGrid in [ix][iy] order */
  const int32_t* grid_items;
  const float* scale_factors;   /* pKF->mvScaleFactors */
  int32_t* match_kf;            /* N : index into vpPoints newly assigned to each key-frame feature, -1 = none */
  int32_t* nmatches;            /* 1 */
  float scw[12];                /* rows 0..2 of Scw = [s R | s t] */
  float cam[4];                 /* pKF->fx, fy, cx, cy */
  int32_t bounds[4];            /* pKF->mnMinX, mnMinY, mnMaxX, mnMaxY (int members of KeyFrame) */
  float grid_width_inv, grid_height_inv, log_scale_factor;
  int32_t grid_cols, grid_rows, n_levels, th, m, n;
} plslam_kfproj_job_t;
int plslam_match_kf_projection_batch_device(const plslam_kfproj_job_t* d_jobs, int njobs, int max_n, void* stream); /* max_n >= every job's n */
int plslam_match_kf_projection_host(const plslam_kfproj_job_t* job); /* HOST pointers inside *job */

/* The matching core of ORBmatcher::Fuse (ORBmatcher.h:119 Fuse(pKF, vpMapPoints, th), @0x7a500, LocalMapping::SearchInNeighbors;
 * ORBmatcher.h:122 Fuse(pKF, Scw, vpPoints, th, vpReplacePoint), @0x7bb20, LoopClosing::SearchAndFuse): for every candidate map
 * point the key-frame feature it would be fused into.  Fuse keeps no matched flags, so the points are independent: one warp per
 * map point.  What depends on the order is the map-graph bookkeeping after each match (pMPinKF->Replace / AddObservation /
 * AddMapPoint, or vpReplacePoint[i] = pMPinKF), which stays with the caller (host veneer), replayed in order over best_idx. */
typedef struct plslam_fuse_job {
  const uint8_t* mp_valid;      /* M : pMP && !isBad() && !IsInKeyFrame(pKF) (with Scw: && not in pKF->GetMapPoints()) */
  const float* mp_xyz;          /* M x 3 */
  const float* mp_normal;       /* M x 3 */
  const float* mp_dist_range;   /* M x 2 : mfMinDistance, mfMaxDistance */
  const uint8_t* mp_desc;       /* M x 32 */
  const int32_t* mp_level;      /* M or NULL, as in plslam_kfproj_job_t */
  const float* kf_xy;           /* N x 2 */
  const int32_t* kf_octave;     /* N */
  const float* kf_uright;       /* N : pKF->mvuRight (chi-square test of the rigid form; unused with Scw) */
  const uint8_t* kf_desc;       /* N x 32 */
  const int32_t* grid_start;    /* grid_cols * grid_rows + 1 */
  const int32_t* grid_items;
  const float* scale_factors;   /* pKF->mvScaleFactors */
  const float* inv_level_sigma2;/* pKF->mvInvLevelSigma2 (rigid form) */
  int32_t* best_idx;            /* M : key-frame feature, -1 = no match */
  float pose[12];               /* use_scw = 0: Rcw | tcw (pKF->GetRotation(), GetTranslation()); 1: rows 0..2 of Scw; 2: see pose2 */
  float pose2[12];              /* use_scw = 2 (one direction of ORBmatcher::SearchBySim3, ORBmatcher.h:116, @0x838b0): the points are
                                   taken into their own key frame's camera with pose (R?w | t?w), then into the OTHER camera with
                                   pose2 (sR21 | t21 or sR12 | t12); the distance is the norm of that last vector, there is no
                                   viewing-angle test, candidates compete from INT_MAX and a match needs bestDist <= TH_HIGH */
  float ow[3];                  /* use_scw = 0: pKF->GetCameraCenter(); 1: ignored (derived from Scw as the reference does) */
  float cam[5];                 /* fx, fy, cx, cy, mbf */
  int32_t bounds[4];            /* mnMinX, mnMinY, mnMaxX, mnMaxY */
  float grid_width_inv, grid_height_inv, log_scale_factor, th;
  int32_t grid_cols, grid_rows, n_levels, use_scw, m, n;
} plslam_fuse_job_t;
int plslam_match_fuse_search_batch_device(const plslam_fuse_job_t* d_jobs, int njobs, int max_m, void* stream);
int plslam_match_fuse_search_host(const plslam_fuse_job_t* job); /* HOST pointers inside *job */
/* sR12 = s12 * R12, sR21 = (1.0 / s12) * R12.t(), t21 = -sR21 * t12 as the reference's cv::Mat expressions evaluate (@0x83a8e:
 * every element times (float)alpha, the product through gemm with scale -1); R12 row-major 3x3.  Host arithmetic, no GPU. */
void plslam_sim3_transforms(float s12, const float* R12, const float* t12, float* sR12, float* sR21, float* t21);

/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th) (ORBmatcher.h:61,
 * @0x79f10) — the local-map search of Tracking::SearchLocalPoints, including Frame::GetFeaturesInArea and
 * RadiusByViewingCos (@0x79b60).  The map-point fields it reads (filled by Frame::isInFrustum in the reference) are
 * passed flattened. */
typedef struct plslam_local_job {
  const uint8_t* mp_valid;     /* M : pMP->mbTrackInView && !pMP->isBad() */
  const float* mp_proj;        /* M x 3 : mTrackProjX, mTrackProjY, mTrackProjXR */
  const int32_t* mp_level;     /* M : mnTrackScaleLevel */
  const float* mp_viewcos;     /* M : mTrackViewCos */
  const uint8_t* mp_desc;      /* M x 32 : pMP->GetDescriptor() */
  const uint8_t* mp_obs;       /* M : pMP->Observations() > 0 */
  const float* f_xy;           /* N x 2 : F.mvKeysUn[i].pt */
  const int32_t* f_octave;     /* N */
  const uint8_t* f_desc;       /* N x 32 */
  const float* f_uright;       /* N : F.mvuRight */
  const uint8_t* f_taken;      /* N : F.mvpMapPoints[i] && Observations() > 0 on entry */
  const int32_t* grid_start;   /* 64*48+1 : CSR of F.mGrid in [ix][iy] order */
  const int32_t* grid_items;
  const float* scale_factors;  /* F.mvScaleFactors */
  int32_t* match_f;            /* N : map-point index assigned to each keypoint, -1 = none */
  int32_t* nmatches;           /* 1 */
  float cam[4];                /* mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv */
  float th, nnratio;           /* th; mfNNratio */
  int32_t m, n;
} plslam_local_job_t;
int plslam_match_local_points_batch_device(const plslam_local_job_t* d_jobs, int njobs, int max_n, void* stream); /* max_n >= every job's n */
int plslam_match_local_points_host(const plslam_local_job_t* job, int n_scale_levels);

/* ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched, vector<int>& vnMatches12,
 * int windowSize) (ORBmatcher.h:108, @0x7db00) — the monocular-initialisation matcher, including
 * Frame::GetFeaturesInArea(x, y, windowSize, 0, 0) on F2's grid.  Only level-0 key points of F1 take part. */
typedef struct plslam_init_job {
  const int32_t* f1_octave;    /* N1 : F1.mvKeysUn[i].octave */
  const float* f1_angle;       /* N1 : F1.mvKeysUn[i].angle */
  const uint8_t* f1_desc;      /* N1 x 32 */
  const float* f2_xy;          /* N2 x 2 : F2.mvKeysUn[i].pt */
  const float* f2_angle;       /* N2 */
  const int32_t* f2_octave;    /* N2 : F2.mvKeysUn[i].octave (only level 0 is searched) */
  const uint8_t* f2_desc;      /* N2 x 32 */
  const int32_t* grid_start;   /* 64*48+1 : CSR of F2.mGrid in [ix][iy] order */
  const int32_t* grid_items;
  float* prev_matched;         /* N1 x 2 : vbPrevMatched, read as the window centres and updated in place for the matches */
  int32_t* match12;            /* N1 : vnMatches12 (F2 key point matched to each F1 key point, -1 = none) */
  int32_t* nmatches;           /* 1  : return value */
  float cam[4];                /* F2: mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv */
  float nnratio;               /* mfNNratio */
  int32_t window_size, n1, n2, check_orientation;
} plslam_init_job_t;
int plslam_match_initialization_batch_device(const plslam_init_job_t* d_jobs, int njobs, int max_n1, int max_n2, void* stream);
int plslam_match_initialization_host(const plslam_init_job_t* job); /* HOST pointers inside *job */

/* ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, vector<pair<size_t,size_t>>&
 * vMatchedPairs, bool bOnlyStereo) (ORBmatcher.h:86, @0x86b30) with CheckDistEpipolarLine (ORBmatcher.h:89, @0x79b90).
 * FeatureVectors as CSR sorted by node id, key-frame fields flattened.  vMatchedPairs = {(i, match12[i]) : match12[i] >= 0}
 * in ascending i, as the reference collects them. */
typedef struct plslam_tri_job {
  const uint8_t* kf1_desc;    /* N1 x 32 : pKF1->mDescriptors */
  const float* kf1_xy;        /* N1 x 2  : pKF1->mvKeysUn[i].pt */
  const float* kf1_angle;     /* N1      : pKF1->mvKeysUn[i].angle */
  const float* kf1_uright;    /* N1      : pKF1->mvuRight (>= 0 means stereo) */
  const uint8_t* kf1_has_mp;  /* N1      : pKF1->GetMapPoint(i) != NULL */
  const int32_t* kf1_nodes;   /* n1_nodes   : FeatureVector keys, ascending */
  const int32_t* kf1_start;   /* n1_nodes+1 */
  const int32_t* kf1_idx;
  const uint8_t* kf2_desc;    /* N2 x 32 */
  const float* kf2_xy;        /* N2 x 2 */
  const float* kf2_angle;     /* N2 */
  const int32_t* kf2_octave;  /* N2 : pKF2->mvKeysUn[i].octave */
  const float* kf2_uright;    /* N2 */
  const uint8_t* kf2_has_mp;  /* N2 */
  const int32_t* kf2_nodes;
  const int32_t* kf2_start;
  const int32_t* kf2_idx;
  const float* scale_factors; /* pKF2->mvScaleFactors */
  const float* level_sigma2;  /* pKF2->mvLevelSigma2 */
  int32_t* match12;           /* N1 : vMatches12 (KF2 feature matched to each KF1 feature, -1 = none) */
  int32_t* nmatches;          /* 1  : return value */
  float F12[9];               /* fundamental matrix, row-major CV_32F */
  float ex, ey;               /* epipole of KF1's centre in KF2: plslam_match_epipole */
  int32_t n1, n2, n1_nodes, n2_nodes;
  int32_t only_stereo;        /* bOnlyStereo */
  int32_t check_orientation;  /* mbCheckOrientation */
} plslam_tri_job_t;
int plslam_match_triangulation_batch_device(const plslam_tri_job_t* d_jobs, int njobs, int max_n, void* stream); /* max_n >= every job's n1 and n2 */
int plslam_match_triangulation_host(const plslam_tri_job_t* job, int n_scale_levels); /* HOST pointers inside *job */
/* The epipole as SearchForTriangulation computes it (@0x86b9c-0x86f8b): C2 = R2w * Cw + t2w, ex = fx * C2x / C2z + cx with
 * the binary's rounding sequence.  R2w row-major 3x3 (pKF2->GetRotation()), t2w (GetTranslation()), Cw (pKF1->GetCameraCenter()). */
int plslam_match_epipole(const float* R2w_3x3, const float* t2w, const float* Cw, float fx, float fy, float cx, float cy,
                         float* ex, float* ey);

/* Single-job convenience forms for the class veneers: every pointer inside *job is a HOST pointer; the call
 * uploads the arrays, runs the kernel and writes match_* / nmatches back.  n_grid_items = grid_start[64*48]. */
int plslam_match_bow_host(const plslam_bow_job_t* job);
/* SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12): kf_* = pKF1, f_* = pKF2 (f_valid required); match12 [n1] receives the
 * KF2 feature assigned to each KF1 feature (-1 = none).  job->match_f is scratch [n2]. */
int plslam_match_bow_kfkf_host(const plslam_bow_job_t* job, int32_t* match12);
int plslam_match_projection_host(const plslam_proj_job_t* job, int n_scale_levels);
/* MapPoint::PredictScale(currentDist, pF / pKF) (@0x8fc20 / @0x8fb60) as the matcher kernels evaluate it (a one-thread kernel
 * on the current device: the arithmetic is the device's, so the two agree by construction); -1 on a CUDA error */
int plslam_predict_scale(float max_distance, float current_dist, float log_scale_factor, int n_levels);

/* Pair matching on the batched extractor outputs without a host round trip: for p in [0, npairs)
 * query = frame (2p), train = frame (2p+1) of a [frames][capacity][32] descriptor block whose valid
 * row counts live in the device array d_counts.  d_out: [npairs][capacity][4] int32 as knn2.
 * d_jobs_scratch: npairs job descriptors of device scratch. */
int plslam_match_knn2_pairs_device(const uint8_t* d_desc, const int32_t* d_counts, int capacity, int npairs,
                                   int32_t* d_out, plslam_knn_job_t* d_jobs_scratch, void* stream);

/* ------------------------------------------------------------------------------------------
 * Front-end = what Frame::Frame (include/Frame.h:60) runs per RGB-D frame on the hot path:
 * ExtractORB (Frame.h:67) + ExtractLSD (Frame.h:70), batched, plus optional frame-pair kNN matching
 * of both descriptor sets (ORB all-pairs and LineSegmentMathch).  ORB and line stages run on two
 * internal streams forked from / joined to the caller's stream.
 * ---------------------------------------------------------------------------------------- */
typedef struct plslam_frontend plslam_frontend_t;
typedef struct plslam_frontend_io {
  plslam_keypoint_t* keypoints;   /* [batch][kp_capacity] */
  uint8_t* descriptors;           /* [batch][kp_capacity][32] */
  int32_t* kp_counts;             /* [batch] */
  plslam_keyline_t* keylines;     /* [batch][line_capacity] */
  uint8_t* line_descriptors;      /* [batch][line_capacity][32] */
  double* line_functions;         /* [batch][line_capacity][3] */
  int32_t* line_counts;           /* [batch] */
  int32_t* orb_matches;           /* [batch/2][kp_capacity][4] or NULL */
  int32_t* line_matches;          /* [batch/2][line_capacity][4] or NULL */
} plslam_frontend_io_t;

int plslam_frontend_create(plslam_frontend_t** out, int nfeatures, float scaleFactor, int nlevels, int iniThFAST,
                           int minThFAST, int max_lines);
/* Same with `depth` independent pipeline slots (each a full workspace): successive process/submit calls rotate
 * over the slots, so up to `depth` batches are in flight on the GPU when the caller uses separate streams (device
 * path) or the asynchronous host path below.  The line detector's region growing is sequential per frame (one warp
 * per frame), so batches in flight are what fills the machine. */
int plslam_frontend_create_pipelined(plslam_frontend_t** out, int nfeatures, float scaleFactor, int nlevels,
                                     int iniThFAST, int minThFAST, int max_lines, int depth);
int plslam_frontend_depth(const plslam_frontend_t* h);
void plslam_frontend_destroy(plslam_frontend_t* h);
int plslam_frontend_capacities(const plslam_frontend_t* h, int* kp_capacity, int* line_capacity);
/* All pointers in `io` are device pointers; asynchronous on `stream`. match_pairs != 0 also fills the match blocks. */
int plslam_frontend_process_device(plslam_frontend_t* h, const uint8_t* d_images, int batch, int width, int height,
                                   int pitch, size_t frame_stride, const plslam_frontend_io_t* io, int match_pairs,
                                   void* stream);
/* All pointers are HOST pointers (pinned memory makes the copies asynchronous); returns after the results landed. */
int plslam_frontend_process_host(plslam_frontend_t* h, const uint8_t* images, int batch, int width, int height,
                                 int pitch, size_t frame_stride, const plslam_frontend_io_t* io, int match_pairs);
/* Asynchronous host path: submit enqueues H2D + kernels + D2H on the next slot and returns; the host output
 * buffers of a submit are valid after plslam_frontend_wait_host(), which waits for every pending slot and reports
 * PLSLAM_ERR_OVERFLOW if any batch overflowed an internal list. */
int plslam_frontend_submit_host(plslam_frontend_t* h, const uint8_t* images, int batch, int width, int height, int pitch,
                                size_t frame_stride, const plslam_frontend_io_t* io, int match_pairs);
int plslam_frontend_wait_host(plslam_frontend_t* h);
/* Completion-ordered variant for deep pipelines: acquire_slot blocks until some slot's previous batch has finished and
 * returns its index (0 .. depth-1); submit_host_slot enqueues on that slot.  The caller can keep one set of host output
 * buffers per slot.  Work enqueued this way is ready to run at once, which keeps the copy engines' in-order queues from
 * coupling the slots to each other. */
int plslam_frontend_acquire_slot(plslam_frontend_t* h);
int plslam_frontend_submit_host_slot(plslam_frontend_t* h, int slot, const uint8_t* images, int batch, int width, int height,
                                     int pitch, size_t frame_stride, const plslam_frontend_io_t* io, int match_pairs);
/* Wave submission (the throughput form of the host path): `n_batches` (1 .. depth) batches of `batch` frames each enter the
 * pipeline together.  images[i] / ios[i] are the host frames and host result buffers of batch i (pinned memory for
 * asynchronous copies).  All uploads of the wave run back to back on an upload stream into input buffers the previous wave
 * is not reading, every batch starts after the wave's last upload (the slots then run in phase, which is where the
 * pipeline is fastest), results leave on a download stream.  Returns once everything is enqueued; the uploads of a wave
 * overlap the kernels of the wave before it.  plslam_frontend_wait_host() returns when all results have landed.  The host
 * buffers of a wave must stay untouched until then.  Successive waves rotate over the pipeline slots (a wave of n batches
 * takes the next n slots), so waves smaller than `depth` keep the whole pipeline busy while only the first wave's upload is
 * exposed: depth / 2 batches per call is what bench.py uses (profiles/r02_wave_sweep.log). */
int plslam_frontend_submit_host_wave(plslam_frontend_t* h, const uint8_t* const* images, int n_batches, int batch, int width,
                                     int height, int pitch, size_t frame_stride, const plslam_frontend_io_t* ios, int match_pairs);
/* Per-stage device times (ms, CUDA events on the launching streams) of the last process call made after
 * plslam_frontend_enable_timing(h, 1).  names/ms hold up to `capacity` entries; returns the number of stages. */
int plslam_frontend_enable_timing(plslam_frontend_t* h, int enable);
int plslam_frontend_stage_times(plslam_frontend_t* h, const char** names, float* ms, int capacity);
/* Number of kernel launches one process call issues for `batch` frames (for launch accounting). */
int plslam_frontend_launches_per_call(const plslam_frontend_t* h, int match_pairs);
/* Synchronises `stream` and reports PLSLAM_ERR_OVERFLOW if an internal list overflowed in the last call. */
int plslam_frontend_check_status(plslam_frontend_t* h, void* stream);

/* ------------------------------------------------------------------------------------------
 * ORB vocabulary — the part of DBoW2 (reference Thirdparty/DBoW2) that Frame::ComputeBoW
 * (include/Frame.h:80, @0xf84f0) uses: ORBVocabulary::loadFromTextFile (TemplatedVocabulary.h:1362-1448)
 * and the per-feature tree descent of transform() (TemplatedVocabulary.h:1242-1284).  The descent
 * (k Hamming distances per level) runs on the GPU; BowVector / FeatureVector assembly (std::map
 * accumulation and L1 normalisation, TemplatedVocabulary.h:1151-1217): plslam_voc_bowvec_batch_device /
 * plslam_voc_compute_bow_host.
 * ---------------------------------------------------------------------------------------- */
typedef struct plslam_voc plslam_voc_t;

/* loadFromTextFile(path) for ORBvoc.txt-style files ("k L scoring weighting" then "parent isLeaf d0..d31 weight"
 * per node; only L1_NORM/TF_IDF = "0 0" is supported). */
int plslam_voc_load_text(plslam_voc_t** out, const char* path);
/* Same from arrays in node order 1..n_nodes-1 (entry 0 = root, ignored): parent id, leaf flag, 32-byte descriptor,
 * weight.  Children lists are built in node order, as loadFromTextFile does. */
int plslam_voc_create(plslam_voc_t** out, int k, int L, int n_nodes, const int32_t* parent, const uint8_t* is_leaf,
                      const uint8_t* descriptors, const double* weights);
void plslam_voc_destroy(plslam_voc_t* h);
int plslam_voc_info(const plslam_voc_t* h, int* k, int* L, int* n_nodes, int* n_words);

/* Flat device image of the vocabulary, for the one-off broadcast at start-up (NCCL over NVLink in bench/tests:
 * rank 0 loads, exports, every rank imports).  export copies blob_bytes to d_out (device memory of the caller). */
size_t plslam_voc_blob_bytes(const plslam_voc_t* h);
int plslam_voc_export_blob(const plslam_voc_t* h, void* d_out, void* stream);
int plslam_voc_import_blob(plslam_voc_t** out, const void* d_blob, size_t bytes);

/* transform(feature, word_id, weight, nid, levelsup) for n descriptors (n x 32 bytes).  d_node receives the node
 * at level L - levelsup (0 when that level is not reached).  Features with weight 0 are "stopped" words. */
int plslam_voc_transform_device(const plslam_voc_t* h, const uint8_t* d_descriptors, int n, int levelsup,
                                int32_t* d_word, double* d_weight, int32_t* d_node, void* stream);
int plslam_voc_transform_host(const plslam_voc_t* h, const uint8_t* descriptors, int n, int levelsup, int32_t* word,
                              double* weight, int32_t* node);

/* Frame::ComputeBoW for a whole batch on the extractor's block layout ([frames][capacity][32] descriptors with
 * d_counts[f] valid rows): the descent of every descriptor plus the assembly of each frame's FeatureVector
 * (`if (w > 0) fv.addFeature(nid, i)`, TemplatedVocabulary.h:1196-1213) as a CSR — d_fv_nodes [frames][capacity]
 * ascending node ids, d_fv_start [frames][capacity+1], d_fv_idx [frames][capacity] feature indices ascending inside
 * a node, d_fv_count [frames] number of nodes — i.e. exactly the arrays plslam_bow_job_t takes.  d_word / d_weight /
 * d_node are [frames][capacity] (BowVector accumulation stays with the caller: plslam_voc_transform_* semantics). */
int plslam_voc_featvec_batch_device(const plslam_voc_t* h, const uint8_t* d_descriptors, const int32_t* d_counts, int frames,
                                    int capacity, int levelsup, int32_t* d_word, double* d_weight, int32_t* d_node,
                                    int32_t* d_fv_nodes, int32_t* d_fv_start, int32_t* d_fv_idx, int32_t* d_fv_count,
                                    void* stream);
/* The other half of Frame::ComputeBoW (Frame.h:80,189: mBowVec): `if (w > 0) v.addWeight(id, w)` in feature order, then
 * v.normalize(L1) (TemplatedVocabulary.h:1196-1217), for every frame of a batch on the outputs of
 * plslam_voc_featvec_batch_device: d_bow_ids / d_bow_vals [frames][capacity] word ids ascending and their normalised
 * weights (std::map iteration order), d_bow_count [frames]. */
int plslam_voc_bowvec_batch_device(const int32_t* d_word, const double* d_weight, const int32_t* d_counts, int frames,
                                   int capacity, int32_t* d_bow_ids, double* d_bow_vals, int32_t* d_bow_count, void* stream);
/* Frame::ComputeBoW of one frame on host arrays (n x 32 descriptor bytes, n <= 16384): mBowVec as (bow_ids, bow_vals)
 * [<= n] and mFeatVec as the CSR (fv_nodes [<= n], fv_start [<= n + 1], fv_idx [<= n]); descent, sort, accumulation and
 * normalisation all run on the device. */
int plslam_voc_compute_bow_host(const plslam_voc_t* h, const uint8_t* descriptors, int n, int levelsup, int32_t* bow_ids,
                                double* bow_vals, int* n_bow, int32_t* fv_nodes, int32_t* fv_start, int32_t* fv_idx, int* n_fv);
/* ORBmatcher::SearchByBoW on the frame pairs of a batch without a host round trip (config C4): pair p matches
 * "keyframe" = frame 2p against frame 2p+1.  d_kf_valid [frames][capacity] = pMP && !pMP->isBad() per keyframe
 * feature.  d_match [npairs][capacity] (index into frame 2p per feature of frame 2p+1, -1 = none), d_nmatches [npairs].
 * Scratch: d_angle_scratch 2*npairs*capacity floats, d_jobs_scratch npairs jobs. */
int plslam_match_bow_pairs_device(const plslam_keypoint_t* d_keypoints, const uint8_t* d_descriptors, const int32_t* d_counts,
                                  int capacity, int npairs, const int32_t* d_fv_nodes, const int32_t* d_fv_start,
                                  const int32_t* d_fv_idx, const int32_t* d_fv_count, const uint8_t* d_kf_valid,
                                  float nnratio, int check_orientation, float* d_angle_scratch, plslam_bow_job_t* d_jobs_scratch,
                                  int32_t* d_match, int32_t* d_nmatches, void* stream);

/* ------------------------------------------------------------------------------------------
 * Frame post-extraction steps — what ORB_SLAM2::Frame::Frame (include/Frame.h:60, @0xf9370) runs on the
 * extractor's output before any matcher can use it: UndistortKeyPoints (Frame.h:266, @0xfa0db),
 * ComputeStereoFromRGBD (Frame.h:120, @0xfa0ea), AssignFeaturesToGrid / PosInGrid (Frame.h:273,110, @0xfa382)
 * and, once per run, ComputeImageBounds (Frame.h:270, @0xfa27e).  Batched on the device so keypoints stay in HBM
 * between extraction and plslam_match_projection_* (whose grid_start / grid_items layout this produces).
 * ---------------------------------------------------------------------------------------- */
typedef struct plslam_frame_calib {
  float fx, fy, cx, cy;     /* mK (Camera.fx ... of the settings file, Examples/RGB-D/TUM1.yaml:8-11) */
  float k1, k2, p1, p2, k3; /* mDistCoef (TUM1.yaml:13-17); k1 == 0 means "not distorted" as in UndistortKeyPoints */
  float bf;                 /* mbf (Camera.bf, TUM1.yaml:26) */
} plslam_frame_calib_t;

/* ComputeImageBounds: bounds4 = mnMinX, mnMaxX, mnMinY, mnMaxY (undistorted image corners; the image itself if
 * k1 == 0).  Four points once per run: host scalar arithmetic, the same double sequence as the device code. */
int plslam_frame_image_bounds(const plslam_frame_calib_t* calib, int cols, int rows, float bounds4[4]);

/* Per frame f of the batch, for its d_counts[f] keypoints (block layout [batch][kp_capacity] as produced by
 * plslam_orb_extract_batch_device / plslam_frontend_process_device):
 *   d_un_xy      [batch][kp_capacity][2]  mvKeysUn[i].pt
 *   d_uright     [batch][kp_capacity]     mvuRight (-1 without depth)
 *   d_depth_out  [batch][kp_capacity]     mvDepth  (-1 without depth)
 *   d_grid_start [batch][64*48+1], d_grid_items [batch][kp_capacity]  CSR of mGrid in [ix][iy] order
 * d_depth: float depth maps (imDepth after convertTo(CV_32F, mDepthMapFactor)), rows x depth_pitch floats per
 * frame, frames depth_frame_stride floats apart (0 = one map shared by all frames). */
int plslam_frame_post_batch_device(const plslam_frame_calib_t* calib, const float bounds4[4],
                                   const plslam_keypoint_t* d_keypoints, const int32_t* d_counts, int batch,
                                   int kp_capacity, const float* d_depth, int cols, int rows, int depth_pitch,
                                   size_t depth_frame_stride, float* d_un_xy, float* d_uright, float* d_depth_out,
                                   int32_t* d_grid_start, int32_t* d_grid_items, void* stream);
/* One frame, HOST pointers (uploads, runs the kernel, copies the results back). */
int plslam_frame_post_host(const plslam_frame_calib_t* calib, const float bounds4[4], const plslam_keypoint_t* keypoints,
                           int n, const float* depth, int cols, int rows, int depth_pitch, float* un_xy, float* uright,
                           float* depth_out, int32_t* grid_start, int32_t* grid_items);

/* Frame::isInFrustum(MapPoint* pMP, float viewingCosLimit) (reference include/Frame.h "isInFrustum", lib/libORB_SLAM2.so@0xf5190)
 * for all local map points of a frame at once: what Tracking::SearchLocalPoints runs per point before
 * ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th); the outputs are the mTrack* fields that matcher reads
 * (plslam_local_job_t: mp_valid = in_view && !isBad(), mp_proj, mp_level, mp_viewcos), one thread per map point. */
typedef struct plslam_frustum_job {
  const float* mp_xyz;         /* M x 3 : pMP->GetWorldPos() */
  const float* mp_normal;      /* M x 3 : pMP->GetNormal() */
  const float* mp_dist_range;  /* M x 2 : mfMinDistance, mfMaxDistance (the invariance range is 0.8f * min .. 1.2f * max) */
  uint8_t* in_view;            /* M : mbTrackInView (the return value) */
  float* proj;                 /* M x 3 : mTrackProjX, mTrackProjY, mTrackProjXR (0 where not in view) */
  int32_t* level;              /* M : mnTrackScaleLevel */
  float* viewcos;              /* M : mTrackViewCos */
  float cam[8];                /* fx, fy, cx, cy, mnMinX, mnMaxX, mnMinY, mnMaxY */
  float tcw[12];               /* mRcw | mtcw, rows 0..2 of the pose */
  float ow[3];                 /* mOw */
  float mbf, log_scale_factor, viewing_cos_limit;
  int32_t n_levels, m;
} plslam_frustum_job_t;
int plslam_frame_is_in_frustum_batch_device(const plslam_frustum_job_t* d_jobs, int njobs, int max_m, void* stream);
int plslam_frame_is_in_frustum_host(const plslam_frustum_job_t* job); /* HOST pointers inside *job */

/* The line analogues of the Frame steps (reference include/Frame.h:267 UndistortKeyLines, :116 GetLinesInArea, :107
 * isInFrustum(MapLine*, float)).  PARITY UNPINNED: the reference declares them and ships no definition in any form (no source;
 * the prebuilt binary has no line code); the definitions follow the header contracts (include/MapLine.h:113-129) and the public
 * PL-SLAM fork family, see oracle/frame_oracle.cc.  Host pointers, one frame per call. */
/* both end points of n key lines through the pinned undistortPoints; xy4 / out_xy4: n x 4 (startX, startY, endX, endY) */
int plslam_frame_undistort_keylines_host(const plslam_frame_calib_t* calib, const float* xy4, int n, float* out_xy4);
/* nq window queries (x1, y1, x2, y2, r, minLevel, maxLevel as 7 floats each) over n key lines (pt.x, pt.y, angle, octave as 4
 * floats each): out_start [nq + 1] offsets into out_items (capacity item_cap, indices ascending inside a query).  Returns
 * PLSLAM_ERR_CAPACITY with out_start[nq] = the needed capacity when item_cap is too small. */
int plslam_frame_lines_in_area_host(const float* queries7, int nq, const float* lines4, int n, int32_t* out_start,
                                    int32_t* out_items, int item_cap);
typedef struct plslam_line_frustum_job {
  const float* ml_sp_ep;       /* M x 6 : MapLine::GetWorldPos() (start point, end point) */
  const float* ml_normal;      /* M x 3 : GetNormal() */
  const float* ml_dist_range;  /* M x 2 : mfMinDistance, mfMaxDistance */
  uint8_t* in_view;            /* M : mbTrackInView */
  float* proj;                 /* M x 6 : mTrackProjX1, Y1, X1R, X2, Y2, X2R */
  int32_t* level;              /* M : mnTrackScaleLevel */
  float* viewcos;              /* M : mTrackViewCos */
  float cam[8];                /* fx, fy, cx, cy, mnMinX, mnMaxX, mnMinY, mnMaxY */
  float tcw[12];               /* mRcw | mtcw */
  float ow[3];                 /* mOw */
  float mbf, log_scale_factor, viewing_cos_limit;
  int32_t n_levels, m;
} plslam_line_frustum_job_t;
int plslam_frame_line_in_frustum_host(const plslam_line_frustum_job_t* job);

/* ------------------------------------------------------------------------------------------------
 * On-disk formats either side of the path (host only, no device work).
 * ------------------------------------------------------------------------------------------------ */
/* LoadImages (Examples/RGB-D/rgbd_tum.cc:151-176): one entry per non-empty line "t_rgb rgb_file t_depth depth_file"; the RGB
 * time stamp is the frame's.  A line that does not parse still yields an entry (0 / empty strings), as in the reference.
 * Names are written as NUL-terminated strings `name_stride` bytes apart.  capacity == 0: only *count is returned. */
int plslam_tum_load_associations(const char* path, double* timestamps, char* rgb_names, char* depth_names, int name_stride,
                                 int capacity, int* count);
/* One line of System::SaveTrajectoryTUM (include/System.h:104, lib/libORB_SLAM2.so@0x3df90): "t tx ty tz qx qy qz qw\n" for the
 * camera pose Tcw (rows 0..2 of the 4x4 CV_32F matrix, row-major 3x4): Rwc = Rcw^T, twc = -Rwc tcw, quaternion of Rwc as
 * Converter::toQuaternion (Eigen, double); fixed notation, 6 decimals for t, 9 for the rest. */
int plslam_tum_pose_to_line(double timestamp, const float* Tcw_3x4, char* line, int capacity, int* length);
int plslam_tum_save_trajectory(const char* path, const double* timestamps, const float* Tcw_3x4, int n);

#ifdef __cplusplus
}
#endif
#endif /* PLSLAM_B200_H */
