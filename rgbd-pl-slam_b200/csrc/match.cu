// match.cu — Hamming matching cores: brute-force kNN(k=2), SearchByBoW, SearchByProjection.
//
// Replaces the inner loops of ORB_SLAM2::ORBmatcher (reference include/ORBmatcher.h:37-141, machine
// code lib/libORB_SLAM2.so@0x79b50-0x8a086) and the kNN matching LSDmatcher / LineSegmentMathch run
// on LBD rows (include/LSDmatcher.h, include/auxiliar.h:30-51).  All-pairs Hamming moves ~64 B per
// descriptor and does 8 popc per pair: it is integer-ALU (popc issue) bound, not HBM bound.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "pl_math.cuh"

namespace plslam {
namespace {

__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ------------------------------------------------------------------------------------------
// kNN (k = 2): one CTA per (job, 256-query slab); the train set streams through shared memory in
// chunks, every thread keeps its query in registers and reads train rows as broadcast LDS.128.
// ------------------------------------------------------------------------------------------
constexpr int KNN_CHUNK = 1024;

__global__ void __launch_bounds__(256) k_knn2(const plslam_knn_job_t* __restrict__ jobs) {
  __shared__ uint4 tr[KNN_CHUNK * 2];
  const plslam_knn_job_t J = jobs[blockIdx.y];
  const int q0 = blockIdx.x * 256;
  if (q0 >= J.nq) return;
  const int q = q0 + threadIdx.x;
  const bool active = q < J.nq;
  uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
  if (active) {
    const uint4* Q = reinterpret_cast<const uint4*>(J.query) + (size_t)q * 2;
    a0 = Q[0];
    a1 = Q[1];
  }
  int b1 = -1, d1 = 1 << 30, b2 = -1, d2 = 1 << 30;
  const uint4* T = reinterpret_cast<const uint4*>(J.train);
  for (int c0 = 0; c0 < J.nt; c0 += KNN_CHUNK) {
    const int cn = min(KNN_CHUNK, J.nt - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < cn * 2; i += 256) tr[i] = T[(size_t)c0 * 2 + i];
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int j = 0; j < cn; ++j) {
        const int d = hamming256(a0, a1, tr[2 * j], tr[2 * j + 1]);
        if (d < d1) { b2 = b1; d2 = d1; b1 = c0 + j; d1 = d; }
        else if (d < d2) { b2 = c0 + j; d2 = d; }
      }
    }
  }
  if (active) {
    int4 o = make_int4(b1, b1 < 0 ? -1 : d1, b2, b2 < 0 ? -1 : d2);
    reinterpret_cast<int4*>(J.out)[q] = o;
  }
}

// ------------------------------------------------------------------------------------------
// rotation-consistency helpers (ComputeThreeMaxima @0x79c40; factor HISTO_LENGTH/360 @0x1269f8)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int rot_bin_dev(float a1, float a2) {
  const float factor = (float)PLSLAM_HISTO_LENGTH / 360.0f;
  float rot = __fsub_rn(a1, a2);
  if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
  int bin = (int)roundf(__fmul_rn(rot, factor));
  if (bin == PLSLAM_HISTO_LENGTH) bin = 0;
  return bin;
}

__device__ void three_maxima_dev(const int* hist, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  ind1 = ind2 = ind3 = -1;
  for (int i = 0; i < PLSLAM_HISTO_LENGTH; i++) {
    const int s = hist[i];
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// warp-wide (dist, order) minimum: returns the lane holding the smallest key; key = dist << 20 | order
__device__ __forceinline__ unsigned warp_min_u32(unsigned v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

// ------------------------------------------------------------------------------------------
// SearchByBoW: greedy and order dependent (an F feature claimed by an earlier KF feature is skipped
// by later ones), so one warp walks the KF features of a job in FeatureVector order and the lanes
// split the F candidates of the shared vocabulary node.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_bow(const plslam_bow_job_t* __restrict__ jobs) {
  extern __shared__ int smem_i[];
  const plslam_bow_job_t J = jobs[blockIdx.x];
  const int lane = threadIdx.x;
  int* matchF = smem_i;                   // [n2]
  int* entryIdx = smem_i + J.n2;          // [n1] matched F index per accepted KF feature, in order
  int* entryBin = entryIdx + J.n1;        // [n1]
  __shared__ int hist[PLSLAM_HISTO_LENGTH];
  for (int i = lane; i < J.n2; i += 32) matchF[i] = -1;
  if (lane < PLSLAM_HISTO_LENGTH) hist[lane] = 0;
  __syncwarp();
  int nmatches = 0, nentries = 0;
  int a = 0, b = 0;
  const uint4* DK = reinterpret_cast<const uint4*>(J.kf_desc);
  const uint4* DF = reinterpret_cast<const uint4*>(J.f_desc);
  while (a < J.n_kf_nodes && b < J.n_f_nodes) {
    const int na = J.kf_nodes[a], nb = J.f_nodes[b];
    if (na == nb) {
      const int fs = J.f_start[b], fe = J.f_start[b + 1];
      for (int iKF = J.kf_start[a]; iKF < J.kf_start[a + 1]; ++iKF) {
        const int realIdxKF = J.kf_idx[iKF];
        if (!J.kf_valid[realIdxKF]) continue;
        const uint4 a0 = DK[2 * realIdxKF], a1 = DK[2 * realIdxKF + 1];
        // per lane: best and second best over its candidates; order = position in the node list
        unsigned k1 = 0xffffffffu, k2 = 0xffffffffu;  // keys dist << 20 | order (order < 2^20)
        for (int iF = fs + lane; iF < fe; iF += 32) {
          const int realIdxF = J.f_idx[iF];
          if (matchF[realIdxF] >= 0) continue;
          if (J.f_valid && !J.f_valid[realIdxF]) continue;  // key frame / key frame form: KF2's map point must be good
          const int dist = hamming256(a0, a1, DF[2 * realIdxF], DF[2 * realIdxF + 1]);
          const unsigned key = ((unsigned)dist << 20) | (unsigned)(iF - fs);
          if (key < k1) { k2 = k1; k1 = key; }
          else if (key < k2) { k2 = key; }
        }
        // global best = min key; second best distance = min over the remaining keys
        const unsigned g1 = warp_min_u32(k1);
        const unsigned mine2 = (k1 == g1) ? k2 : k1;
        const unsigned g2 = warp_min_u32(mine2);
        if (g1 == 0xffffffffu) continue;
        const int bestDist1 = (int)(g1 >> 20);
        const int bestDist2 = g2 == 0xffffffffu ? 256 : (int)(g2 >> 20);
        if (bestDist1 <= PLSLAM_TH_LOW - (J.strict_low ? 1 : 0) && (float)bestDist1 < __fmul_rn(J.nnratio, (float)bestDist2)) {
          const int bestIdxF = J.f_idx[fs + (int)(g1 & 0xfffffu)];
          if (lane == 0) {
            matchF[bestIdxF] = realIdxKF;
            if (J.check_orientation) {
              const int bin = rot_bin_dev(J.kf_angle[realIdxKF], J.f_angle[bestIdxF]);
              hist[bin]++;
              entryIdx[nentries] = bestIdxF;
              entryBin[nentries] = bin;
            }
          }
          ++nentries;
          ++nmatches;
          __syncwarp();
        }
      }
      ++a;
      ++b;
    } else if (na < nb) {
      while (a < J.n_kf_nodes && J.kf_nodes[a] < nb) ++a;  // lower_bound
    } else {
      while (b < J.n_f_nodes && J.f_nodes[b] < na) ++b;
    }
  }
  __syncwarp();
  if (J.check_orientation) {
    int i1, i2, i3;
    three_maxima_dev(hist, i1, i2, i3);
    int removed = 0;
    for (int e = lane; e < nentries; e += 32) {
      const int bin = entryBin[e];
      if (bin != i1 && bin != i2 && bin != i3) {
        matchF[entryIdx[e]] = -1;
        ++removed;
      }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, d);
    nmatches -= removed;
  }
  __syncwarp();
  for (int i = lane; i < J.n2; i += 32) J.match_f[i] = matchF[i];
  if (lane == 0) *J.nmatches = nmatches;
}

// ------------------------------------------------------------------------------------------
// SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (@0x86b30): greedy like SearchByBoW (a KF2 feature
// matched by an earlier KF1 feature is skipped later: the vbMatched2 bit IS set in this build, @0x87bc9), so one warp walks
// the KF1 features of a job in FeatureVector order and the lanes split the KF2 candidates of the shared node.  Within one
// KF1 feature the sequential rule "dist <= bestDist replaces" has a closed form: every test but that one is independent of
// the running state, so the winner is the smallest distance among the candidates passing them, the LAST one on ties.
// CheckDistEpipolarLine (@0x79b90) with the binary's fused operations (__fmaf_rn where it has vfmadd).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool check_dist_epipolar_line_dev(float x1, float y1, float x2, float y2, const float* F, float sigma2) {
  const float b = __fadd_rn(__fmaf_rn(x1, F[1], __fmul_rn(y1, F[4])), F[7]);
  const float a = __fadd_rn(__fmaf_rn(x1, F[0], __fmul_rn(y1, F[3])), F[6]);
  const float den = __fmaf_rn(a, a, __fmul_rn(b, b));
  if (den == 0.0f) return false;
  const float c = __fadd_rn(__fmaf_rn(y1, F[5], __fmul_rn(x1, F[2])), F[8]);
  const float num = __fadd_rn(c, __fmaf_rn(b, y2, __fmul_rn(a, x2)));
  const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
  return __dmul_rn(3.84, (double)sigma2) > (double)dsqr;
}

__global__ void __launch_bounds__(32) k_triangulation(const plslam_tri_job_t* __restrict__ jobs) {
  extern __shared__ int smem_i[];
  const plslam_tri_job_t& J = jobs[blockIdx.x];
  const int lane = threadIdx.x;
  const int N1 = J.n1, N2 = J.n2;
  int* matched2 = smem_i;            // [N2]
  int* entryIdx = matched2 + N2;     // [N1] accepted KF1 feature indices, in order
  int* entryBin = entryIdx + N1;     // [N1]
  __shared__ int hist[PLSLAM_HISTO_LENGTH];
  __shared__ float F[9];
  for (int i = lane; i < N2; i += 32) matched2[i] = 0;
  for (int i = lane; i < N1; i += 32) J.match12[i] = -1;
  if (lane < PLSLAM_HISTO_LENGTH) hist[lane] = 0;
  if (lane < 9) F[lane] = J.F12[lane];
  __syncwarp();
  const float ex = J.ex, ey = J.ey;
  const bool onlyStereo = J.only_stereo != 0;
  int nmatches = 0, nentries = 0;
  int a = 0, b = 0;
  const uint4* D1 = reinterpret_cast<const uint4*>(J.kf1_desc);
  const uint4* D2 = reinterpret_cast<const uint4*>(J.kf2_desc);
  while (a < J.n1_nodes && b < J.n2_nodes) {
    const int na = J.kf1_nodes[a], nb = J.kf2_nodes[b];
    if (na == nb) {
      const int fs = J.kf2_start[b], fe = J.kf2_start[b + 1];
      for (int i1 = J.kf1_start[a]; i1 < J.kf1_start[a + 1]; ++i1) {
        const int idx1 = J.kf1_idx[i1];
        if (J.kf1_has_mp[idx1]) continue;
        const bool bStereo1 = J.kf1_uright[idx1] >= 0.0f;
        if (onlyStereo && !bStereo1) continue;
        const uint4 a0 = D1[2 * idx1], a1 = D1[2 * idx1 + 1];
        const float x1 = J.kf1_xy[2 * idx1], y1 = J.kf1_xy[2 * idx1 + 1];
        // key = dist << 20 | (0xfffff - order): the minimum is the smallest distance, the last candidate on ties
        unsigned k = 0xffffffffu;
        for (int i2 = fs + lane; i2 < fe; i2 += 32) {
          const int idx2 = J.kf2_idx[i2];
          if (matched2[idx2] || J.kf2_has_mp[idx2]) continue;
          const bool bStereo2 = J.kf2_uright[idx2] >= 0.0f;
          if (onlyStereo && !bStereo2) continue;
          const int dist = hamming256(a0, a1, D2[2 * idx2], D2[2 * idx2 + 1]);
          if (dist > PLSLAM_TH_LOW) continue;
          const float x2 = J.kf2_xy[2 * idx2], y2 = J.kf2_xy[2 * idx2 + 1];
          const int oct2 = J.kf2_octave[idx2];
          if (!bStereo1 && !bStereo2) {
            const float dx = __fsub_rn(ex, x2), dy = __fsub_rn(ey, y2);
            if (__fmul_rn(100.0f, J.scale_factors[oct2]) > __fmaf_rn(dx, dx, __fmul_rn(dy, dy))) continue;
          }
          if (!check_dist_epipolar_line_dev(x1, y1, x2, y2, F, J.level_sigma2[oct2])) continue;
          const unsigned key = ((unsigned)dist << 20) | (0xfffffu - (unsigned)(i2 - fs));
          k = min(k, key);
        }
        const unsigned g = warp_min_u32(k);
        if (g == 0xffffffffu) continue;
        const int bestIdx2 = J.kf2_idx[fs + (int)(0xfffffu - (g & 0xfffffu))];
        if (lane == 0) {
          J.match12[idx1] = bestIdx2;
          matched2[bestIdx2] = 1;
          if (J.check_orientation) {
            const int bin = rot_bin_dev(J.kf1_angle[idx1], J.kf2_angle[bestIdx2]);
            hist[bin]++;
            entryIdx[nentries] = idx1;
            entryBin[nentries] = bin;
          }
        }
        ++nentries;
        ++nmatches;
        __syncwarp();
      }
      ++a;
      ++b;
    } else if (na < nb) {
      while (a < J.n1_nodes && J.kf1_nodes[a] < nb) ++a;  // lower_bound
    } else {
      while (b < J.n2_nodes && J.kf2_nodes[b] < na) ++b;
    }
  }
  __syncwarp();
  if (J.check_orientation) {
    int i1, i2, i3;
    three_maxima_dev(hist, i1, i2, i3);
    int removed = 0;
    for (int e = lane; e < nentries; e += 32) {
      const int bin = entryBin[e];
      if (bin != i1 && bin != i2 && bin != i3) {
        J.match12[entryIdx[e]] = -1;
        ++removed;
      }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, d);
    nmatches -= removed;
  }
  if (lane == 0) *J.nmatches = nmatches;
}

// ------------------------------------------------------------------------------------------
// SearchByProjection(CurrentFrame, LastFrame, th, bMono): one warp per frame pair walks the last
// frame's map points in order (a current keypoint claimed by an earlier point is skipped later);
// lanes split the grid cells of the search window, candidate order = [ix][iy][position in cell].
// ------------------------------------------------------------------------------------------
__global__ void k_predict_scale(float maxDistance, float dist, float logScaleFactor, int nLevels, int* out) {
  *out = predict_scale_dev(maxDistance, dist, logScaleFactor, nLevels);
}

__global__ void __launch_bounds__(32) k_projection(const plslam_proj_job_t* __restrict__ jobs) {
  extern __shared__ int smem_i[];
  const plslam_proj_job_t& J = jobs[blockIdx.x];
  const int lane = threadIdx.x;
  const int N1 = J.n1, N2 = J.n2;
  int* matchCur = smem_i;            // [N2]
  int* entryIdx = matchCur + N2;     // [N1]
  int* entryBin = entryIdx + N1;     // [N1]
  uint8_t* taken = reinterpret_cast<uint8_t*>(entryBin + N1);  // [N2]
  __shared__ int hist[PLSLAM_HISTO_LENGTH];
  for (int i = lane; i < N2; i += 32) { matchCur[i] = -1; taken[i] = J.cur_taken[i]; }
  if (lane < PLSLAM_HISTO_LENGTH) hist[lane] = 0;
  __syncwarp();
  const float fx = J.cam[0], fy = J.cam[1], cx = J.cam[2], cy = J.cam[3], mbf = J.cam[4], mb = J.cam[5];
  const float mnMinX = J.cam[6], mnMaxX = J.cam[7], mnMinY = J.cam[8], mnMaxY = J.cam[9], gwi = J.cam[10], ghi = J.cam[11];
  const float* Rc = J.tcw_cur;
  float twc[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double s = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) s = __dadd_rn(s, __dmul_rn((double)Rc[k * 4 + r], (double)Rc[k * 4 + 3]));
    twc[r] = (float)__dmul_rn(-1.0, s);
  }
  // Rlw * twc + tlw and Rcw * x3Dw + tcw are cv::gemm calls without a transpose flag and an inner dimension of 3: the
  // small-matrix path (float products summed left to right in float, then (double)sum + (double)c rounded to float).
  // -Rcw.t() * tcw above carries a transpose flag: GEMMSingleMul, double accumulator.
  float tlc2;
  {
    const float p0 = __fmul_rn(J.tcw_last[8], twc[0]), p1 = __fmul_rn(J.tcw_last[9], twc[1]), p2 = __fmul_rn(J.tcw_last[10], twc[2]);
    const float s = __fadd_rn(__fadd_rn(p0, p1), p2);
    tlc2 = (float)__dadd_rn((double)s, (double)J.tcw_last[11]);
  }
  const bool kfMode = J.mode == 1;  // SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist): see the job struct
  const bool bForward = tlc2 > mb && !J.mono && !kfMode;
  const bool bBackward = -tlc2 > mb && !J.mono && !kfMode;
  const int accept = kfMode ? J.orb_dist : PLSLAM_TH_HIGH;
  int nmatches = 0, nentries = 0;
  const uint4* DL = reinterpret_cast<const uint4*>(J.last_desc);
  const uint4* DC = reinterpret_cast<const uint4*>(J.cur_desc);
  for (int i = 0; i < N1; ++i) {
    if (!J.last_valid[i]) continue;
    const float* X = J.last_xyz + 3 * i;
    float pc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float p0 = __fmul_rn(Rc[r * 4], X[0]), p1 = __fmul_rn(Rc[r * 4 + 1], X[1]), p2 = __fmul_rn(Rc[r * 4 + 2], X[2]);
      const float s = __fadd_rn(__fadd_rn(p0, p1), p2);
      pc[r] = (float)__dadd_rn((double)s, (double)Rc[r * 4 + 3]);
    }
    // the binary's sequence: vdivss @0x81c92 (float division), vmulss + vfmadd213ss @0x81caf-0x81cba / @0x81cce-0x81cd9
    const float invzc = __fdiv_rn(1.0f, pc[2]);
    if (!kfMode && invzc < 0) continue;  // (the key-frame form has no sign test: @0x7f4cd)
    const float u = __fmaf_rn(__fmul_rn(pc[0], fx), invzc, cx);
    const float v = __fmaf_rn(__fmul_rn(pc[1], fy), invzc, cy);
    if (u < mnMinX || u > mnMaxX) continue;
    if (v < mnMinY || v > mnMaxY) continue;
    int oct;
    if (kfMode && !J.last_dist_range) {
      oct = J.last_octave[i];  // the caller ran the distance test and MapPoint::PredictScale itself (reference-signature veneer)
    } else if (kfMode) {
      // PO = x3Dw - Ow (float), dist3D = (float)cv::norm(PO): squares summed in double in element order (@0x7f96e-0x7fa81)
      double n2 = 0;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double d = (double)__fsub_rn(X[r], twc[r]);
        n2 = __dadd_rn(n2, __dmul_rn(d, d));
      }
      const float dist3D = (float)sqrt(n2);
      const float dMin = J.last_dist_range[2 * i], dMax = J.last_dist_range[2 * i + 1];
      if (__fmul_rn(0.8f, dMin) > dist3D || dist3D > __fmul_rn(1.2f, dMax)) continue;
      oct = predict_scale_dev(dMax, dist3D, J.log_scale_factor, J.n_levels);
    } else {
      oct = J.last_octave[i];
    }
    const float radius = __fmul_rn(J.th, J.scale_factors[oct]);
    int minLevel, maxLevel;
    if (bForward) { minLevel = oct; maxLevel = -1; }
    else if (bBackward) { minLevel = 0; maxLevel = oct; }
    else { minLevel = oct - 1; maxLevel = oct + 1; }
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(u, mnMinX), radius), gwi)));
    if (nMinCellX >= PLSLAM_GRID_COLS) continue;
    const int nMaxCellX = min(PLSLAM_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(u, mnMinX), radius), gwi)));
    if (nMaxCellX < 0) continue;
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(v, mnMinY), radius), ghi)));
    if (nMinCellY >= PLSLAM_GRID_ROWS) continue;
    const int nMaxCellY = min(PLSLAM_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(v, mnMinY), radius), ghi)));
    if (nMaxCellY < 0) continue;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    const int ncy = nMaxCellY - nMinCellY + 1, ncells = (nMaxCellX - nMinCellX + 1) * ncy;
    const uint4 a0 = DL[2 * i], a1 = DL[2 * i + 1];
    const float ur = __fmaf_rn(-invzc, mbf, u);  // vfnmadd132ss @0x81eb5
    unsigned best = 0xffffffffu;  // dist << 22 | cellRank << 8 | pos  (cellRank < 2^14, pos < 2^8)
    int bestI2 = -1;
    for (int c = lane; c < ncells; c += 32) {
      const int ix = nMinCellX + c / ncy, iy = nMinCellY + c % ncy;
      const int cell = ix * PLSLAM_GRID_ROWS + iy;
      const int s0 = J.grid_start[cell], s1 = J.grid_start[cell + 1];
      for (int j = s0; j < s1; ++j) {
        const int i2 = J.grid_items[j];
        if (bCheckLevels) {
          const int o2 = J.cur_octave[i2];
          if (o2 < minLevel) continue;
          if (maxLevel >= 0 && o2 > maxLevel) continue;
        }
        const float distx = __fsub_rn(J.cur_xy[2 * i2], u), disty = __fsub_rn(J.cur_xy[2 * i2 + 1], v);
        if (!(fabsf(distx) < radius && fabsf(disty) < radius)) continue;
        if (taken[i2]) continue;
        if (!kfMode) {
          const float ur2 = J.cur_uright[i2];
          if (ur2 > 0) {
            if (fabsf(__fsub_rn(ur, ur2)) > radius) continue;
          }
        }
        const int dist = hamming256(a0, a1, DC[2 * i2], DC[2 * i2 + 1]);
        const unsigned key = ((unsigned)dist << 22) | ((unsigned)c << 8) | (unsigned)min(j - s0, 255);
        if (key < best) { best = key; bestI2 = i2; }
      }
    }
    const unsigned g = warp_min_u32(best);
    if (g == 0xffffffffu) continue;
    const int bestDist = (int)(g >> 22);
    if (bestDist <= accept) {
      const unsigned src = __ballot_sync(0xffffffffu, best == g);
      const int bestIdx2 = __shfl_sync(0xffffffffu, bestI2, __ffs(src) - 1);
      if (lane == 0) {
        matchCur[bestIdx2] = i;
        if (kfMode || J.last_obs[i]) taken[bestIdx2] = 1;
        if (J.check_orientation) {
          const int bin = rot_bin_dev(J.last_angle[i], J.cur_angle[bestIdx2]);
          hist[bin]++;
          entryIdx[nentries] = bestIdx2;
          entryBin[nentries] = bin;
        }
      }
      ++nentries;
      ++nmatches;
      __syncwarp();
    }
  }
  __syncwarp();
  if (J.check_orientation) {
    int i1, i2, i3;
    three_maxima_dev(hist, i1, i2, i3);
    int removed = 0;
    for (int e = lane; e < nentries; e += 32) {
      const int bin = entryBin[e];
      if (bin != i1 && bin != i2 && bin != i3) {
        matchCur[entryIdx[e]] = J.report_removed ? -2 : -1;
        ++removed;
      }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, d);
    nmatches -= removed;
  }
  __syncwarp();
  for (int i = lane; i < N2; i += 32) J.match_cur[i] = matchCur[i];
  if (lane == 0) *J.nmatches = nmatches;
}


// ------------------------------------------------------------------------------------------
// SearchByProjection(KeyFrame* pKF, cv::Mat Scw, vpPoints, vpMatched, th) (@0x880f0, loop closing): one warp per job walks the
// map points in order; lanes split the cells of KeyFrame::GetFeaturesInArea's window (candidate order [ix][iy][position]).
// The similarity is taken apart as the binary does it (see oracle/match_oracle.cc oracle_search_by_projection_sim3): scale from
// the first row of sRcw (dot in double), every element times (float)(1.0 / scw), Ow through the double-accumulating gemm.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_kf_projection(const plslam_kfproj_job_t* __restrict__ jobs) {
  extern __shared__ int smem_i[];
  const plslam_kfproj_job_t& J = jobs[blockIdx.x];
  const int lane = threadIdx.x;
  const int M = J.m, N = J.n;
  int* matchKF = smem_i;                                        // [N]
  uint8_t* matched = reinterpret_cast<uint8_t*>(matchKF + N);   // [N]
  for (int i = lane; i < N; i += 32) { matchKF[i] = -1; matched[i] = J.kf_matched[i]; }
  __syncwarp();
  const float fx = J.cam[0], fy = J.cam[1], cx = J.cam[2], cy = J.cam[3];
  const float mnMinX = (float)J.bounds[0], mnMinY = (float)J.bounds[1], mnMaxX = (float)J.bounds[2], mnMaxY = (float)J.bounds[3];
  const float gwi = J.grid_width_inv, ghi = J.grid_height_inv;
  const int cols = J.grid_cols, rows = J.grid_rows;
  float T[12];
  {
    double d0 = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) d0 = __dadd_rn(d0, __dmul_rn((double)J.scw[k], (double)J.scw[k]));
    const float scw = (float)sqrt(d0);
    const float inv = (float)__ddiv_rn(1.0, (double)scw);
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = __fadd_rn(__fmul_rn(J.scw[k], inv), 0.0f);
  }
  float Ow[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double s = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) s = __dadd_rn(s, __dmul_rn((double)T[k * 4 + r], (double)T[k * 4 + 3]));
    Ow[r] = (float)__dmul_rn(-1.0, s);
  }
  const uint4* DM = reinterpret_cast<const uint4*>(J.mp_desc);
  const uint4* DK = reinterpret_cast<const uint4*>(J.kf_desc);
  int nmatches = 0;
  for (int i = 0; i < M; ++i) {
    if (!J.mp_valid[i]) continue;
    const float* X = J.mp_xyz + 3 * (size_t)i;
    float pc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float p0 = __fmul_rn(T[r * 4], X[0]), p1 = __fmul_rn(T[r * 4 + 1], X[1]), p2 = __fmul_rn(T[r * 4 + 2], X[2]);
      pc[r] = (float)__dadd_rn((double)__fadd_rn(__fadd_rn(p0, p1), p2), (double)T[r * 4 + 3]);
    }
    if (pc[2] < 0.0f) continue;
    const float invz = __fdiv_rn(1.0f, pc[2]);
    const float u = __fmaf_rn(__fmul_rn(pc[0], invz), fx, cx), v = __fmaf_rn(__fmul_rn(pc[1], invz), fy, cy);
    if (!(u >= mnMinX && u < mnMaxX && v >= mnMinY && v < mnMaxY)) continue;
    int level;
    if (J.mp_level) {
      level = J.mp_level[i];
    } else {
      float PO[3];
      double n2 = 0;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        PO[r] = __fsub_rn(X[r], Ow[r]);
        n2 = __dadd_rn(n2, __dmul_rn((double)PO[r], (double)PO[r]));
      }
      const float dist = (float)sqrt(n2);
      const float dMin = J.mp_dist_range[2 * (size_t)i], dMax = J.mp_dist_range[2 * (size_t)i + 1];
      if (dist < __fmul_rn(0.8f, dMin) || dist > __fmul_rn(1.2f, dMax)) continue;
      double dot = 0;
#pragma unroll
      for (int r = 0; r < 3; ++r) dot = __dadd_rn(dot, __dmul_rn((double)PO[r], (double)J.mp_normal[3 * (size_t)i + r]));
      if (dot < __dmul_rn(0.5, (double)dist)) continue;
      level = predict_scale_dev(dMax, dist, J.log_scale_factor, J.n_levels);
    }
    const float radius = __fmul_rn((float)J.th, J.scale_factors[level]);
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(u, mnMinX), radius), gwi)));
    if (nMinCellX >= cols) continue;
    const int nMaxCellX = min(cols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(u, mnMinX), radius), gwi)));
    if (nMaxCellX < 0) continue;
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(v, mnMinY), radius), ghi)));
    if (nMinCellY >= rows) continue;
    const int nMaxCellY = min(rows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(v, mnMinY), radius), ghi)));
    if (nMaxCellY < 0) continue;
    const int ncy = nMaxCellY - nMinCellY + 1, ncells = (nMaxCellX - nMinCellX + 1) * ncy;
    const uint4 a0 = DM[2 * (size_t)i], a1 = DM[2 * (size_t)i + 1];
    unsigned best = 0xffffffffu;  // dist << 22 | cellRank << 8 | pos
    int bestI = -1;
    for (int c = lane; c < ncells; c += 32) {
      const int ix = nMinCellX + c / ncy, iy = nMinCellY + c % ncy;
      const int cell = ix * rows + iy;
      const int s0 = J.grid_start[cell], s1 = J.grid_start[cell + 1];
      for (int j = s0; j < s1; ++j) {
        const int idx = J.grid_items[j];
        const float distx = __fsub_rn(J.kf_xy[2 * idx], u), disty = __fsub_rn(J.kf_xy[2 * idx + 1], v);
        if (!(fabsf(distx) < radius && fabsf(disty) < radius)) continue;
        if (matched[idx]) continue;
        const int kpLevel = J.kf_octave[idx];
        if (kpLevel < level - 1 || kpLevel > level) continue;
        const int d = hamming256(a0, a1, DK[2 * idx], DK[2 * idx + 1]);
        const unsigned key = ((unsigned)d << 22) | ((unsigned)c << 8) | (unsigned)min(j - s0, 255);
        if (key < best) { best = key; bestI = idx; }
      }
    }
    const unsigned g = warp_min_u32(best);
    if (g == 0xffffffffu) continue;
    if ((int)(g >> 22) <= PLSLAM_TH_LOW) {
      const unsigned src = __ballot_sync(0xffffffffu, best == g);
      const int bestIdx = __shfl_sync(0xffffffffu, bestI, __ffs(src) - 1);
      if (lane == 0) {
        matchKF[bestIdx] = i;
        matched[bestIdx] = 1;
      }
      ++nmatches;
      __syncwarp();
    }
  }
  __syncwarp();
  for (int i = lane; i < N; i += 32) J.match_kf[i] = matchKF[i];
  if (lane == 0) *J.nmatches = nmatches;
}

// ------------------------------------------------------------------------------------------
// The matching core of ORBmatcher::Fuse (@0x7a500 rigid pose with the chi-square tests, @0x7bb20 similarity): the map points are
// independent, so one WARP per map point (8 per CTA, blockIdx.y = job); lanes split the cells of KeyFrame::GetFeaturesInArea.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fuse_search(const plslam_fuse_job_t* __restrict__ jobs) {
  const plslam_fuse_job_t& J = jobs[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= J.m) return;
  int result = -1;
  do {
    if (!J.mp_valid[i]) break;
    const float fx = J.cam[0], fy = J.cam[1], cx = J.cam[2], cy = J.cam[3], bf = J.cam[4];
    const float mnMinX = (float)J.bounds[0], mnMinY = (float)J.bounds[1], mnMaxX = (float)J.bounds[2], mnMaxY = (float)J.bounds[3];
    const float gwi = J.grid_width_inv, ghi = J.grid_height_inv;
    const int cols = J.grid_cols, rows = J.grid_rows;
    float T[12], Ow[3];
    if (J.use_scw == 1) {
      double d0 = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) d0 = __dadd_rn(d0, __dmul_rn((double)J.pose[k], (double)J.pose[k]));
      const float inv = (float)__ddiv_rn(1.0, (double)(float)sqrt(d0));
#pragma unroll
      for (int k = 0; k < 12; ++k) T[k] = __fadd_rn(__fmul_rn(J.pose[k], inv), 0.0f);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) s = __dadd_rn(s, __dmul_rn((double)T[k * 4 + r], (double)T[k * 4 + 3]));
        Ow[r] = (float)__dmul_rn(-1.0, s);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 12; ++k) T[k] = J.pose[k];
#pragma unroll
      for (int r = 0; r < 3; ++r) Ow[r] = J.ow[r];
    }
    const bool sim3 = J.use_scw == 2;
    const int accept = sim3 ? PLSLAM_TH_HIGH : PLSLAM_TH_LOW;
    const float* X = J.mp_xyz + 3 * (size_t)i;
    float pc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float p0 = __fmul_rn(T[r * 4], X[0]), p1 = __fmul_rn(T[r * 4 + 1], X[1]), p2 = __fmul_rn(T[r * 4 + 2], X[2]);
      pc[r] = (float)__dadd_rn((double)__fadd_rn(__fadd_rn(p0, p1), p2), (double)T[r * 4 + 3]);
    }
    if (sim3) {  // into the other key frame's camera
      float pb[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float p0 = __fmul_rn(J.pose2[r * 4], pc[0]), p1 = __fmul_rn(J.pose2[r * 4 + 1], pc[1]), p2 = __fmul_rn(J.pose2[r * 4 + 2], pc[2]);
        pb[r] = (float)__dadd_rn((double)__fadd_rn(__fadd_rn(p0, p1), p2), (double)J.pose2[r * 4 + 3]);
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) pc[r] = pb[r];
    }
    if (pc[2] < 0.0f) break;
    const float invz = __fdiv_rn(1.0f, pc[2]);
    const float u = __fmaf_rn(__fmul_rn(pc[0], invz), fx, cx), v = __fmaf_rn(fy, __fmul_rn(pc[1], invz), cy);
    if (!(u >= mnMinX && u < mnMaxX && v >= mnMinY && v < mnMaxY)) break;
    const float ur = __fmaf_rn(-bf, invz, u);
    int level;
    if (J.mp_level) {
      level = J.mp_level[i];
    } else if (sim3) {
      double n2 = 0;
#pragma unroll
      for (int r = 0; r < 3; ++r) n2 = __dadd_rn(n2, __dmul_rn((double)pc[r], (double)pc[r]));
      const float dist = (float)sqrt(n2);
      const float dMin = J.mp_dist_range[2 * (size_t)i], dMax = J.mp_dist_range[2 * (size_t)i + 1];
      if (dist < __fmul_rn(0.8f, dMin) || dist > __fmul_rn(1.2f, dMax)) break;
      level = predict_scale_dev(dMax, dist, J.log_scale_factor, J.n_levels);
    } else {
      float PO[3];
      double n2 = 0;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        PO[r] = __fsub_rn(X[r], Ow[r]);
        n2 = __dadd_rn(n2, __dmul_rn((double)PO[r], (double)PO[r]));
      }
      const float dist = (float)sqrt(n2);
      const float dMin = J.mp_dist_range[2 * (size_t)i], dMax = J.mp_dist_range[2 * (size_t)i + 1];
      if (dist < __fmul_rn(0.8f, dMin) || dist > __fmul_rn(1.2f, dMax)) break;
      double dot = 0;
#pragma unroll
      for (int r = 0; r < 3; ++r) dot = __dadd_rn(dot, __dmul_rn((double)PO[r], (double)J.mp_normal[3 * (size_t)i + r]));
      if (dot < __dmul_rn(0.5, (double)dist)) break;
      level = predict_scale_dev(dMax, dist, J.log_scale_factor, J.n_levels);
    }
    const float radius = __fmul_rn(J.th, J.scale_factors[level]);
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(u, mnMinX), radius), gwi)));
    if (nMinCellX >= cols) break;
    const int nMaxCellX = min(cols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(u, mnMinX), radius), gwi)));
    if (nMaxCellX < 0) break;
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(v, mnMinY), radius), ghi)));
    if (nMinCellY >= rows) break;
    const int nMaxCellY = min(rows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(v, mnMinY), radius), ghi)));
    if (nMaxCellY < 0) break;
    const int ncy = nMaxCellY - nMinCellY + 1, ncells = (nMaxCellX - nMinCellX + 1) * ncy;
    const uint4* DM = reinterpret_cast<const uint4*>(J.mp_desc);
    const uint4* DK = reinterpret_cast<const uint4*>(J.kf_desc);
    const uint4 a0 = DM[2 * (size_t)i], a1 = DM[2 * (size_t)i + 1];
    unsigned best = 0xffffffffu;  // dist << 22 | cellRank << 8 | pos
    int bestI = -1;
    for (int c = lane; c < ncells; c += 32) {
      const int ix = nMinCellX + c / ncy, iy = nMinCellY + c % ncy;
      const int cell = ix * rows + iy;
      const int s0 = J.grid_start[cell], s1 = J.grid_start[cell + 1];
      for (int j = s0; j < s1; ++j) {
        const int idx = J.grid_items[j];
        const float kx = J.kf_xy[2 * idx], ky = J.kf_xy[2 * idx + 1];
        if (!(fabsf(__fsub_rn(kx, u)) < radius && fabsf(__fsub_rn(ky, v)) < radius)) continue;
        const int kpLevel = J.kf_octave[idx];
        if (kpLevel < level - 1 || kpLevel > level) continue;
        if (J.use_scw == 0) {
          const float ex = __fsub_rn(u, kx), ey = __fsub_rn(v, ky);
          const float kr = J.kf_uright[idx];
          float e2 = __fmaf_rn(ex, ex, __fmul_rn(ey, ey));
          double bound = 5.99;
          if (kr >= 0) {
            const float er = __fsub_rn(ur, kr);
            e2 = __fmaf_rn(er, er, e2);
            bound = 7.8;
          }
          if ((double)__fmul_rn(e2, J.inv_level_sigma2[kpLevel]) > bound) continue;
        }
        const int d = hamming256(a0, a1, DK[2 * idx], DK[2 * idx + 1]);
        const unsigned key = ((unsigned)d << 22) | ((unsigned)c << 8) | (unsigned)min(j - s0, 255);
        if (key < best) { best = key; bestI = idx; }
      }
    }
    const unsigned g = warp_min_u32(best);
    if (g == 0xffffffffu || (int)(g >> 22) > accept) break;
    const unsigned src = __ballot_sync(0xffffffffu, best == g);
    result = __shfl_sync(0xffffffffu, bestI, __ffs(src) - 1);
  } while (false);
  if (lane == 0) J.best_idx[i] = result;
}

// ------------------------------------------------------------------------------------------
// SearchByProjection(Frame &F, const vector<MapPoint*>&, th) (@0x79f10, Tracking::SearchLocalPoints): one warp per
// frame walks the local map points in order (a keypoint assigned to an observed map point is skipped later); lanes
// split the grid cells of the search window.  The reference's best / second-best update over the candidates in
// [ix][iy][position] order ends with the two smallest (distance, order) keys: the best is the first occurrence of
// the minimum, the second-best slot holds the first occurrence of the smallest remaining value (see
// oracle/match_oracle.cc for the sequential form the result is compared with).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_local_points(const plslam_local_job_t* __restrict__ jobs) {
  extern __shared__ int smem_i[];
  const plslam_local_job_t& J = jobs[blockIdx.x];
  const int lane = threadIdx.x;
  const int M = J.m, N = J.n;
  int* matchF = smem_i;                                       // [N]
  uint8_t* taken = reinterpret_cast<uint8_t*>(matchF + N);    // [N]
  for (int i = lane; i < N; i += 32) { matchF[i] = -1; taken[i] = J.f_taken[i]; }
  __syncwarp();
  const float mnMinX = J.cam[0], mnMinY = J.cam[1], gwi = J.cam[2], ghi = J.cam[3];
  const bool bFactor = J.th != 1.0f;
  const uint4* DM = reinterpret_cast<const uint4*>(J.mp_desc);
  const uint4* DF = reinterpret_cast<const uint4*>(J.f_desc);
  int nmatches = 0;
  for (int i = 0; i < M; ++i) {
    if (!J.mp_valid[i]) continue;
    const int level = J.mp_level[i];
    float r = (double)J.mp_viewcos[i] > 0.998 ? 2.5f : 4.0f;  // RadiusByViewingCos compares in double (@0x79b64-0x79b70)
    if (bFactor) r = __fmul_rn(r, J.th);
    const float x = J.mp_proj[3 * i], y = J.mp_proj[3 * i + 1], xr = J.mp_proj[3 * i + 2];
    const float radius = __fmul_rn(r, J.scale_factors[level]);
    const int minLevel = level - 1, maxLevel = level;
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, mnMinX), radius), gwi)));
    if (nMinCellX >= PLSLAM_GRID_COLS) continue;
    const int nMaxCellX = min(PLSLAM_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, mnMinX), radius), gwi)));
    if (nMaxCellX < 0) continue;
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, mnMinY), radius), ghi)));
    if (nMinCellY >= PLSLAM_GRID_ROWS) continue;
    const int nMaxCellY = min(PLSLAM_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, mnMinY), radius), ghi)));
    if (nMaxCellY < 0) continue;
    const int ncy = nMaxCellY - nMinCellY + 1, ncells = (nMaxCellX - nMinCellX + 1) * ncy;
    const uint4 a0 = DM[2 * i], a1 = DM[2 * i + 1];
    // per lane: the two smallest keys dist << 22 | cellRank << 8 | pos, with keypoint index and octave
    unsigned k1 = 0xffffffffu, k2 = 0xffffffffu;
    int i1 = -1, o1 = -1, o2 = -1;
    for (int c = lane; c < ncells; c += 32) {
      const int ix = nMinCellX + c / ncy, iy = nMinCellY + c % ncy;
      const int cell = ix * PLSLAM_GRID_ROWS + iy;
      const int s0 = J.grid_start[cell], s1 = J.grid_start[cell + 1];
      for (int j = s0; j < s1; ++j) {
        const int idx = J.grid_items[j];
        const int oc = J.f_octave[idx];
        if (oc < minLevel || oc > maxLevel) continue;  // bCheckLevels is always true here (maxLevel = level >= 0)
        const float distx = __fsub_rn(J.f_xy[2 * idx], x), disty = __fsub_rn(J.f_xy[2 * idx + 1], y);
        if (!(fabsf(distx) < radius && fabsf(disty) < radius)) continue;
        if (taken[idx]) continue;
        const float ur = J.f_uright[idx];
        if (ur > 0 && fabsf(__fsub_rn(xr, ur)) > radius) continue;
        const int dist = hamming256(a0, a1, DF[2 * idx], DF[2 * idx + 1]);
        const unsigned key = ((unsigned)dist << 22) | ((unsigned)c << 8) | (unsigned)min(j - s0, 255);
        if (key < k1) { k2 = k1; o2 = o1; k1 = key; i1 = idx; o1 = oc; }
        else if (key < k2) { k2 = key; o2 = oc; }
      }
    }
    const unsigned g1 = warp_min_u32(k1);
    if (g1 == 0xffffffffu) continue;
    const int bestDist = (int)(g1 >> 22);
    if (bestDist > PLSLAM_TH_HIGH) continue;
    const int src1 = __ffs(__ballot_sync(0xffffffffu, k1 == g1)) - 1;
    const int bestIdx = __shfl_sync(0xffffffffu, i1, src1), bestLevel = __shfl_sync(0xffffffffu, o1, src1);
    // runner-up over all lanes: the owner of the best contributes its own second key
    const unsigned mine2 = lane == src1 ? k2 : k1;
    const int mine2o = lane == src1 ? o2 : o1;
    const unsigned g2 = warp_min_u32(mine2);
    int bestDist2 = 256, bestLevel2 = -1;
    if (g2 != 0xffffffffu) {
      const int src2 = __ffs(__ballot_sync(0xffffffffu, mine2 == g2)) - 1;
      bestDist2 = (int)(g2 >> 22);
      bestLevel2 = __shfl_sync(0xffffffffu, mine2o, src2);
    }
    if (bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(J.nnratio, (float)bestDist2)) continue;
    if (lane == 0) {
      matchF[bestIdx] = i;
      if (J.mp_obs[i]) taken[bestIdx] = 1;
    }
    ++nmatches;
    __syncwarp();
  }
  __syncwarp();
  for (int i = lane; i < N; i += 32) J.match_f[i] = matchF[i];
  if (lane == 0) *J.nmatches = nmatches;
}

// ------------------------------------------------------------------------------------------
// SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (ORBmatcher.h:108, @0x7db00), the monocular
// initialisation matcher: one warp per frame pair walks F1's level-0 key points in order (a re-matched F2 key point
// releases its earlier partner, so the scan is order dependent); lanes split the grid cells of the window around
// vbPrevMatched[i1].  Sequential best / second-best with strict '<' over the candidates in [ix][iy][position] order =
// the two smallest (distance, order) keys; a candidate whose recorded distance is <= the present one is skipped.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_search_init(const plslam_init_job_t* __restrict__ jobs) {
  extern __shared__ int smem_i[];
  const plslam_init_job_t& J = jobs[blockIdx.x];
  const int lane = threadIdx.x;
  const int N1 = J.n1, N2 = J.n2;
  int* match12 = smem_i;             // [N1]
  int* match21 = match12 + N1;       // [N2]
  int* matchedDist = match21 + N2;   // [N2]
  int* entryBin = matchedDist + N2;  // [N1] rotation-histogram bin of i1 (-1: never accepted)
  __shared__ int hist[PLSLAM_HISTO_LENGTH];
  for (int i = lane; i < N1; i += 32) { match12[i] = -1; entryBin[i] = -1; }
  for (int i = lane; i < N2; i += 32) { match21[i] = -1; matchedDist[i] = 0x7fffffff; }
  if (lane < PLSLAM_HISTO_LENGTH) hist[lane] = 0;
  __syncwarp();
  const float mnMinX = J.cam[0], mnMinY = J.cam[1], gwi = J.cam[2], ghi = J.cam[3];
  const float radius = (float)J.window_size;
  const uint4* D1 = reinterpret_cast<const uint4*>(J.f1_desc);
  const uint4* D2 = reinterpret_cast<const uint4*>(J.f2_desc);
  int nmatches = 0;
  for (int i1 = 0; i1 < N1; ++i1) {
    if (J.f1_octave[i1] > 0) continue;
    const float x = J.prev_matched[2 * i1], y = J.prev_matched[2 * i1 + 1];
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, mnMinX), radius), gwi)));
    if (nMinCellX >= PLSLAM_GRID_COLS) continue;
    const int nMaxCellX = min(PLSLAM_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, mnMinX), radius), gwi)));
    if (nMaxCellX < 0) continue;
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, mnMinY), radius), ghi)));
    if (nMinCellY >= PLSLAM_GRID_ROWS) continue;
    const int nMaxCellY = min(PLSLAM_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, mnMinY), radius), ghi)));
    if (nMaxCellY < 0) continue;
    const int ncy = nMaxCellY - nMinCellY + 1, ncells = (nMaxCellX - nMinCellX + 1) * ncy;
    const uint4 a0 = D1[2 * i1], a1 = D1[2 * i1 + 1];
    // GetFeaturesInArea(x, y, windowSize, 0, 0): maxLevel = 0 >= 0 => bCheckLevels, only level-0 key points of F2 pass
    unsigned k1 = 0xffffffffu, k2 = 0xffffffffu;
    int c1 = -1;
    for (int c = lane; c < ncells; c += 32) {
      const int ix = nMinCellX + c / ncy, iy = nMinCellY + c % ncy;
      const int cell = ix * PLSLAM_GRID_ROWS + iy;
      const int s0 = J.grid_start[cell], s1 = J.grid_start[cell + 1];
      for (int j = s0; j < s1; ++j) {
        const int i2 = J.grid_items[j];
        if (J.f2_octave[i2] != 0) continue;
        const float distx = __fsub_rn(J.f2_xy[2 * i2], x), disty = __fsub_rn(J.f2_xy[2 * i2 + 1], y);
        if (!(fabsf(distx) < radius && fabsf(disty) < radius)) continue;
        const int dist = hamming256(a0, a1, D2[2 * i2], D2[2 * i2 + 1]);
        if (matchedDist[i2] <= dist) continue;
        const unsigned key = ((unsigned)dist << 22) | ((unsigned)c << 8) | (unsigned)min(j - s0, 255);
        if (key < k1) { k2 = k1; k1 = key; c1 = i2; }
        else if (key < k2) { k2 = key; }
      }
    }
    const unsigned g1 = warp_min_u32(k1);
    if (g1 == 0xffffffffu) continue;  // (an empty window, or only candidates that are matched better already)
    const int bestDist = (int)(g1 >> 22);
    if (bestDist > PLSLAM_TH_LOW) continue;
    const int src1 = __ffs(__ballot_sync(0xffffffffu, k1 == g1)) - 1;
    const int bestIdx2 = __shfl_sync(0xffffffffu, c1, src1);
    const unsigned mine2 = lane == src1 ? k2 : k1;
    const unsigned g2 = warp_min_u32(mine2);
    // bestDist2 stays INT_MAX without a runner-up: (float)INT_MAX * nnratio is far above any distance
    const float lim = g2 == 0xffffffffu ? __fmul_rn(2147483648.f, J.nnratio) : __fmul_rn((float)(int)(g2 >> 22), J.nnratio);
    if (!((float)bestDist < lim)) continue;
    if (lane == 0) {
      const int old = match21[bestIdx2];
      if (old >= 0) match12[old] = -1;
      match12[i1] = bestIdx2;
      match21[bestIdx2] = i1;
      matchedDist[bestIdx2] = bestDist;
      if (J.check_orientation) {
        const int bin = rot_bin_dev(J.f1_angle[i1], J.f2_angle[bestIdx2]);
        hist[bin]++;
        entryBin[i1] = bin;
      }
    }
    __syncwarp();
  }
  __syncwarp();
  if (J.check_orientation) {
    int b1, b2, b3;
    three_maxima_dev(hist, b1, b2, b3);
    for (int i = lane; i < N1; i += 32) {
      const int bin = entryBin[i];
      if (bin >= 0 && bin != b1 && bin != b2 && bin != b3) match12[i] = -1;
    }
    __syncwarp();
  }
  for (int i = lane; i < N1; i += 32) {
    const int m = match12[i];
    J.match12[i] = m;
    if (m >= 0) {
      ++nmatches;
      J.prev_matched[2 * i] = J.f2_xy[2 * m];
      J.prev_matched[2 * i + 1] = J.f2_xy[2 * m + 1];
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) nmatches += __shfl_xor_sync(0xffffffffu, nmatches, d);
  if (lane == 0) *J.nmatches = nmatches;
}

}  // namespace
}  // namespace plslam

using namespace plslam;

extern "C" {

int plslam_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  // Bit-count of a XOR b over 8 x 32-bit words, as FORB::distance (reference Thirdparty/DBoW2/DBoW2/FORB.cpp:82-102)
  int dist = 0;
  for (int i = 0; i < 8; ++i) {
    uint32_t x, y;
    std::memcpy(&x, a + 4 * i, 4);
    std::memcpy(&y, b + 4 * i, 4);
    dist += __builtin_popcount(x ^ y);
  }
  return dist;
}

int plslam_match_knn2_batch_device(const plslam_knn_job_t* d_jobs, int njobs, int max_nq, void* stream) {
  PL_CHECK_ARG(d_jobs && njobs >= 1 && njobs <= 65535 && max_nq >= 1);
  PL_CARVEOUT(k_knn2);
  k_knn2<<<dim3(div_up(max_nq, 256), njobs), 256, 0, (cudaStream_t)stream>>>(d_jobs);
  PL_CUDA(cudaGetLastError());
  return PLSLAM_OK;
}

int plslam_match_knn2_host(const uint8_t* query, int nq, const uint8_t* train, int nt, int32_t* out) {
  PL_CHECK_ARG(nq >= 0 && nt >= 0 && (nq == 0 || (query && out)) && (nt == 0 || train));
  if (nq == 0) return PLSLAM_OK;
  DevBuf dq, dt, dout, djob;
  int rc;
  if ((rc = dq.ensure((size_t)nq * 32)) || (rc = dt.ensure((size_t)std::max(nt, 1) * 32)) ||
      (rc = dout.ensure((size_t)nq * 16)) || (rc = djob.ensure(sizeof(plslam_knn_job_t)))) {
    dq.release(); dt.release(); dout.release(); djob.release();
    return rc;
  }
  auto cleanup = [&]() { dq.release(); dt.release(); dout.release(); djob.release(); };
  cudaError_t e = cudaMemcpy(dq.p, query, (size_t)nq * 32, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && nt) e = cudaMemcpy(dt.p, train, (size_t)nt * 32, cudaMemcpyHostToDevice);
  plslam_knn_job_t job{dq.as<uint8_t>(), dt.as<uint8_t>(), dout.as<int32_t>(), nq, nt};
  if (e == cudaSuccess) e = cudaMemcpy(djob.p, &job, sizeof(job), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = plslam_match_knn2_batch_device(djob.as<plslam_knn_job_t>(), 1, nq, nullptr);
    if (rc) { cleanup(); return rc; }
    e = cudaMemcpy(out, dout.p, (size_t)nq * 16, cudaMemcpyDeviceToHost);
  }
  cleanup();
  if (e != cudaSuccess) {
    set_error("knn2 host path: %s", cudaGetErrorString(e));
    return PLSLAM_ERR_CUDA;
  }
  return PLSLAM_OK;
}

int plslam_match_bow_batch_device(const plslam_bow_job_t* d_jobs, int njobs, int max_n, void* stream) {
  const int max_n2 = max_n;
  PL_CHECK_ARG(d_jobs && njobs >= 1 && max_n2 >= 0);
  // shared: matchF[n2] + entryIdx[n1] + entryBin[n1]; n1 is bounded by the caller's max_n2-style capacity
  const size_t smem = (size_t)max_n2 * 4 * 3;
  PL_CHECK_ARG(smem <= 200 * 1024);
  static PerDeviceOnce attr;
  if (attr.first()) {
    PL_CUDA(cudaFuncSetAttribute(k_bow, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    PL_CUDA(cudaFuncSetAttribute(k_projection, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  PL_CARVEOUT(k_bow);
  k_bow<<<njobs, 32, smem, (cudaStream_t)stream>>>(d_jobs);
  PL_CUDA(cudaGetLastError());
  return PLSLAM_OK;
}

int plslam_match_projection_batch_device(const plslam_proj_job_t* d_jobs, int njobs, int max_n1, int max_n2, void* stream) {
  PL_CHECK_ARG(d_jobs && njobs >= 1 && max_n1 >= 0 && max_n2 >= 0);
  const size_t smem = (size_t)max_n2 * 4 + (size_t)max_n1 * 8 + (size_t)max_n2 + 16;
  PL_CHECK_ARG(smem <= 200 * 1024);
  static PerDeviceOnce attr;
  if (attr.first()) {
    PL_CUDA(cudaFuncSetAttribute(k_bow, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    PL_CUDA(cudaFuncSetAttribute(k_projection, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  PL_CARVEOUT(k_projection);
  k_projection<<<njobs, 32, smem, (cudaStream_t)stream>>>(d_jobs);
  PL_CUDA(cudaGetLastError());
  return PLSLAM_OK;
}

}  // extern "C"

namespace {
// tiny RAII helper for the single-job host paths
struct Uploader {
  std::vector<void*> bufs;
  cudaError_t err = cudaSuccess;
  ~Uploader() { for (void* p : bufs) cudaFree(p); }
  template <typename T>
  const T* up(const T* host, size_t n) {
    void* d = nullptr;
    if (err == cudaSuccess) err = cudaMalloc(&d, std::max<size_t>(n * sizeof(T), 16));
    if (err == cudaSuccess) bufs.push_back(d);
    if (err == cudaSuccess && n) err = cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice);
    return static_cast<const T*>(d);
  }
  template <typename T>
  T* out(size_t n) {
    void* d = nullptr;
    if (err == cudaSuccess) err = cudaMalloc(&d, std::max<size_t>(n * sizeof(T), 16));
    if (err == cudaSuccess) bufs.push_back(d);
    return static_cast<T*>(d);
  }
};
}  // namespace

extern "C" {

int plslam_match_bow_host(const plslam_bow_job_t* job) {
  PL_CHECK_ARG(job && job->match_f && job->nmatches && job->n1 >= 0 && job->n2 >= 0);
  Uploader U;
  plslam_bow_job_t d = *job;
  const int nk = job->n_kf_nodes, nf = job->n_f_nodes;
  const int lenK = nk ? job->kf_start[nk] : 0, lenF = nf ? job->f_start[nf] : 0;
  d.kf_desc = U.up(job->kf_desc, (size_t)job->n1 * 32);
  d.kf_angle = U.up(job->kf_angle, job->n1);
  d.kf_valid = U.up(job->kf_valid, job->n1);
  d.kf_nodes = U.up(job->kf_nodes, nk);
  d.kf_start = U.up(job->kf_start, nk + 1);
  d.kf_idx = U.up(job->kf_idx, lenK);
  d.f_desc = U.up(job->f_desc, (size_t)job->n2 * 32);
  d.f_angle = U.up(job->f_angle, job->n2);
  d.f_nodes = U.up(job->f_nodes, nf);
  d.f_start = U.up(job->f_start, nf + 1);
  d.f_idx = U.up(job->f_idx, lenF);
  if (job->f_valid) d.f_valid = U.up(job->f_valid, job->n2);
  d.match_f = U.out<int32_t>(job->n2);
  d.nmatches = U.out<int32_t>(1);
  const plslam_bow_job_t* dj = U.up(&d, 1);
  if (U.err != cudaSuccess) { set_error("bow host path: %s", cudaGetErrorString(U.err)); return PLSLAM_ERR_CUDA; }
  int rc = plslam_match_bow_batch_device(dj, 1, std::max(job->n1, job->n2), nullptr);
  if (rc) return rc;
  PL_CUDA(cudaMemcpy(job->match_f, d.match_f, (size_t)job->n2 * 4, cudaMemcpyDeviceToHost));
  PL_CUDA(cudaMemcpy(job->nmatches, d.nmatches, 4, cudaMemcpyDeviceToHost));
  return PLSLAM_OK;
}

int plslam_match_bow_kfkf_host(const plslam_bow_job_t* job, int32_t* match12) {
  PL_CHECK_ARG(job && match12 && job->f_valid && job->match_f);
  plslam_bow_job_t j = *job;
  j.strict_low = 1;
  int rc = plslam_match_bow_host(&j);
  if (rc) return rc;
  for (int i = 0; i < job->n1; ++i) match12[i] = -1;
  for (int i2 = 0; i2 < job->n2; ++i2)
    if (job->match_f[i2] >= 0) match12[job->match_f[i2]] = i2;  // a KF2 feature is matched at most once, a KF1 feature too
  return PLSLAM_OK;
}

int plslam_match_triangulation_batch_device(const plslam_tri_job_t* d_jobs, int njobs, int max_n, void* stream) {
  PL_CHECK_ARG(d_jobs && njobs >= 1 && max_n >= 0);
  const size_t smem = (size_t)max_n * 4 * 3;  // matched2[n2] + entryIdx[n1] + entryBin[n1]
  PL_CHECK_ARG(smem <= 200 * 1024);
  static PerDeviceOnce attr;
  if (attr.first()) {
    PL_CUDA(cudaFuncSetAttribute(k_triangulation, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  PL_CARVEOUT(k_triangulation);
  k_triangulation<<<njobs, 32, smem, (cudaStream_t)stream>>>(d_jobs);
  PL_CUDA(cudaGetLastError());
  return PLSLAM_OK;
}

// C2 = R2w * Cw + t2w through cv::gemm's small-matrix path (float sum, scaled and added in double), invz = 1 / C2z,
// ex = fma(invz, fx * C2x, cx) — the sequence of @0x86b9c-0x86f8b.  Host scalar work, once per key-frame pair.
int plslam_match_epipole(const float* R2w_3x3, const float* t2w, const float* Cw, float fx, float fy, float cx, float cy,
                         float* ex, float* ey) {
  PL_CHECK_ARG(R2w_3x3 && t2w && Cw && ex && ey);
  float C2[3];
  for (int r = 0; r < 3; ++r) {
    float s = R2w_3x3[3 * r] * Cw[0];
    s = s + R2w_3x3[3 * r + 1] * Cw[1];
    s = s + R2w_3x3[3 * r + 2] * Cw[2];
    C2[r] = (float)((double)s * 1.0 + (double)t2w[r] * 1.0);
  }
  const float invz = 1.0f / C2[2];
  *ex = fmaf(invz, fx * C2[0], cx);
  *ey = fmaf(invz, fy * C2[1], cy);
  return PLSLAM_OK;
}

int plslam_match_triangulation_host(const plslam_tri_job_t* job, int n_scale_levels) {
  PL_CHECK_ARG(job && job->match12 && job->nmatches && job->n1 >= 0 && job->n2 >= 0 && n_scale_levels >= 1);
  Uploader U;
  plslam_tri_job_t d = *job;
  const int n1 = job->n1, n2 = job->n2, na = job->n1_nodes, nb = job->n2_nodes;
  const int len1 = na ? job->kf1_start[na] : 0, len2 = nb ? job->kf2_start[nb] : 0;
  d.kf1_desc = U.up(job->kf1_desc, (size_t)n1 * 32);
  d.kf1_xy = U.up(job->kf1_xy, (size_t)n1 * 2);
  d.kf1_angle = U.up(job->kf1_angle, n1);
  d.kf1_uright = U.up(job->kf1_uright, n1);
  d.kf1_has_mp = U.up(job->kf1_has_mp, n1);
  d.kf1_nodes = U.up(job->kf1_nodes, na);
  d.kf1_start = U.up(job->kf1_start, na + 1);
  d.kf1_idx = U.up(job->kf1_idx, len1);
  d.kf2_desc = U.up(job->kf2_desc, (size_t)n2 * 32);
  d.kf2_xy = U.up(job->kf2_xy, (size_t)n2 * 2);
  d.kf2_angle = U.up(job->kf2_angle, n2);
  d.kf2_octave = U.up(job->kf2_octave, n2);
  d.kf2_uright = U.up(job->kf2_uright, n2);
  d.kf2_has_mp = U.up(job->kf2_has_mp, n2);
  d.kf2_nodes = U.up(job->kf2_nodes, nb);
  d.kf2_start = U.up(job->kf2_start, nb + 1);
  d.kf2_idx = U.up(job->kf2_idx, len2);
  d.scale_factors = U.up(job->scale_factors, n_scale_levels);
  d.level_sigma2 = U.up(job->level_sigma2, n_scale_levels);
  d.match12 = U.out<int32_t>(std::max(n1, 1));
  d.nmatches = U.out<int32_t>(1);
  const plslam_tri_job_t* dj = U.up(&d, 1);
  if (U.err != cudaSuccess) { set_error("triangulation host path: %s", cudaGetErrorString(U.err)); return PLSLAM_ERR_CUDA; }
  int rc = plslam_match_triangulation_batch_device(dj, 1, std::max(n1, n2), nullptr);
  if (rc) return rc;
  if (n1) PL_CUDA(cudaMemcpy(job->match12, d.match12, (size_t)n1 * 4, cudaMemcpyDeviceToHost));
  PL_CUDA(cudaMemcpy(job->nmatches, d.nmatches, 4, cudaMemcpyDeviceToHost));
  return PLSLAM_OK;
}

int plslam_match_projection_host(const plslam_proj_job_t* job, int n_scale_levels) {
  PL_CHECK_ARG(job && job->match_cur && job->nmatches && job->n1 >= 0 && job->n2 >= 0 && n_scale_levels >= 1);
  Uploader U;
  plslam_proj_job_t d = *job;
  const int n1 = job->n1, n2 = job->n2, ncell = PLSLAM_GRID_COLS * PLSLAM_GRID_ROWS;
  const int nitems = job->grid_start[ncell];
  d.last_valid = U.up(job->last_valid, n1);
  d.last_xyz = U.up(job->last_xyz, (size_t)n1 * 3);
  d.last_desc = U.up(job->last_desc, (size_t)n1 * 32);
  const bool kfMode = job->mode == 1;
  PL_CHECK_ARG(!kfMode || job->last_octave || (job->last_dist_range && job->n_levels >= 1 && job->n_levels <= n_scale_levels));
  d.last_octave = job->last_octave ? U.up(job->last_octave, n1) : nullptr;
  d.last_angle = U.up(job->last_angle, n1);
  d.last_obs = kfMode ? nullptr : U.up(job->last_obs, n1);
  d.last_dist_range = kfMode && job->last_dist_range ? U.up(job->last_dist_range, (size_t)n1 * 2) : nullptr;
  d.cur_xy = U.up(job->cur_xy, (size_t)n2 * 2);
  d.cur_octave = U.up(job->cur_octave, n2);
  d.cur_angle = U.up(job->cur_angle, n2);
  d.cur_desc = U.up(job->cur_desc, (size_t)n2 * 32);
  d.cur_uright = kfMode ? nullptr : U.up(job->cur_uright, n2);
  d.cur_taken = U.up(job->cur_taken, n2);
  d.grid_start = U.up(job->grid_start, ncell + 1);
  d.grid_items = U.up(job->grid_items, nitems);
  d.scale_factors = U.up(job->scale_factors, n_scale_levels);
  d.match_cur = U.out<int32_t>(n2);
  d.nmatches = U.out<int32_t>(1);
  const plslam_proj_job_t* dj = U.up(&d, 1);
  if (U.err != cudaSuccess) { set_error("projection host path: %s", cudaGetErrorString(U.err)); return PLSLAM_ERR_CUDA; }
  int rc = plslam_match_projection_batch_device(dj, 1, n1, n2, nullptr);
  if (rc) return rc;
  PL_CUDA(cudaMemcpy(job->match_cur, d.match_cur, (size_t)n2 * 4, cudaMemcpyDeviceToHost));
  PL_CUDA(cudaMemcpy(job->nmatches, d.nmatches, 4, cudaMemcpyDeviceToHost));
  return PLSLAM_OK;
}

int plslam_match_kf_projection_batch_device(const plslam_kfproj_job_t* d_jobs, int njobs, int max_n, void* stream) {
  PL_CHECK_ARG(d_jobs && njobs >= 1 && max_n >= 0);
  const size_t smem = (size_t)max_n * 5 + 16;  // matchKF[n] ints + matched[n] bytes
  PL_CHECK_ARG(smem <= 200 * 1024);
  static PerDeviceOnce attr;
  if (attr.first()) {
    PL_CUDA(cudaFuncSetAttribute(k_kf_projection, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  PL_CARVEOUT(k_kf_projection);
  k_kf_projection<<<njobs, 32, smem, (cudaStream_t)stream>>>(d_jobs);
  PL_CUDA(cudaGetLastError());
  return PLSLAM_OK;
}

int plslam_match_kf_projection_host(const plslam_kfproj_job_t* job) {
  PL_CHECK_ARG(job && job->match_kf && job->nmatches && job->m >= 0 && job->n >= 0 && job->n_levels >= 1 && job->grid_cols >= 1 &&
               job->grid_rows >= 1 && job->grid_start);
  Uploader U;
  plslam_kfproj_job_t d = *job;
  const int m = job->m, n = job->n, ncell = job->grid_cols * job->grid_rows;
  const int nitems = job->grid_start[ncell];
  d.mp_valid = U.up(job->mp_valid, m);
  d.mp_xyz = U.up(job->mp_xyz, (size_t)m * 3);
  PL_CHECK_ARG(job->mp_level || (job->mp_normal && job->mp_dist_range));
  d.mp_normal = job->mp_level ? nullptr : U.up(job->mp_normal, (size_t)m * 3);
  d.mp_dist_range = job->mp_level ? nullptr : U.up(job->mp_dist_range, (size_t)m * 2);
  d.mp_level = job->mp_level ? U.up(job->mp_level, m) : nullptr;
  d.mp_desc = U.up(job->mp_desc, (size_t)m * 32);
  d.kf_xy = U.up(job->kf_xy, (size_t)n * 2);
  d.kf_octave = U.up(job->kf_octave, n);
  d.kf_desc = U.up(job->kf_desc, (size_t)n * 32);
  d.kf_matched = U.up(job->kf_matched, n);
  d.grid_start = U.up(job->grid_start, ncell + 1);
  d.grid_items = U.up(job->grid_items, nitems);
  d.scale_factors = U.up(job->scale_factors, job->n_levels);
  d.match_kf = U.out<int32_t>(std::max(n, 1));
  d.nmatches = U.out<int32_t>(1);
  const plslam_kfproj_job_t* dj = U.up(&d, 1);
  if (U.err != cudaSuccess) { set_error("kf projection host path: %s", cudaGetErrorString(U.err)); return PLSLAM_ERR_CUDA; }
  int rc = plslam_match_kf_projection_batch_device(dj, 1, n, nullptr);
  if (rc) return rc;
  if (n) PL_CUDA(cudaMemcpy(job->match_kf, d.match_kf, (size_t)n * 4, cudaMemcpyDeviceToHost));
  PL_CUDA(cudaMemcpy(job->nmatches, d.nmatches, 4, cudaMemcpyDeviceToHost));
  return PLSLAM_OK;
}

int plslam_match_fuse_search_batch_device(const plslam_fuse_job_t* d_jobs, int njobs, int max_m, void* stream) {
  PL_CHECK_ARG(d_jobs && njobs >= 1 && njobs <= 65535 && max_m >= 0);
  if (max_m == 0) return PLSLAM_OK;
  PL_CARVEOUT(k_fuse_search);
  k_fuse_search<<<dim3(div_up(max_m, 8), njobs), 256, 0, (cudaStream_t)stream>>>(d_jobs);
  PL_CUDA(cudaGetLastError());
  return PLSLAM_OK;
}

int plslam_match_fuse_search_host(const plslam_fuse_job_t* job) {
  PL_CHECK_ARG(job && job->best_idx && job->m >= 0 && job->n >= 0 && job->n_levels >= 1 && job->grid_cols >= 1 && job->grid_rows >= 1 &&
               job->grid_start && (job->mp_level || (job->mp_dist_range && (job->mp_normal || job->use_scw == 2))) &&
               (job->use_scw || (job->kf_uright && job->inv_level_sigma2)));
  Uploader U;
  plslam_fuse_job_t d = *job;
  const int m = job->m, n = job->n, ncell = job->grid_cols * job->grid_rows;
  if (m == 0) return PLSLAM_OK;
  const int nitems = job->grid_start[ncell];
  d.mp_valid = U.up(job->mp_valid, m);
  d.mp_xyz = U.up(job->mp_xyz, (size_t)m * 3);
  d.mp_normal = (job->mp_level || !job->mp_normal) ? nullptr : U.up(job->mp_normal, (size_t)m * 3);
  d.mp_dist_range = job->mp_level ? nullptr : U.up(job->mp_dist_range, (size_t)m * 2);
  d.mp_level = job->mp_level ? U.up(job->mp_level, m) : nullptr;
  d.mp_desc = U.up(job->mp_desc, (size_t)m * 32);
  d.kf_xy = U.up(job->kf_xy, (size_t)n * 2);
  d.kf_octave = U.up(job->kf_octave, n);
  d.kf_uright = job->kf_uright ? U.up(job->kf_uright, n) : nullptr;
  d.kf_desc = U.up(job->kf_desc, (size_t)n * 32);
  d.grid_start = U.up(job->grid_start, ncell + 1);
  d.grid_items = U.up(job->grid_items, nitems);
  d.scale_factors = U.up(job->scale_factors, job->n_levels);
  d.inv_level_sigma2 = job->inv_level_sigma2 ? U.up(job->inv_level_sigma2, job->n_levels) : nullptr;
  d.best_idx = U.out<int32_t>(m);
  const plslam_fuse_job_t* dj = U.up(&d, 1);
  if (U.err != cudaSuccess) { set_error("fuse host path: %s", cudaGetErrorString(U.err)); return PLSLAM_ERR_CUDA; }
  int rc = plslam_match_fuse_search_batch_device(dj, 1, m, nullptr);
  if (rc) return rc;
  PL_CUDA(cudaMemcpy(job->best_idx, d.best_idx, (size_t)m * 4, cudaMemcpyDeviceToHost));
  return PLSLAM_OK;
}

void plslam_sim3_transforms(float s12, const float* R12, const float* t12, float* sR12, float* sR21, float* t21) {
  const float a12 = (float)(double)s12, a21 = (float)(1.0 / (double)s12);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      volatile float p = R12[3 * i + j] * a12, q = R12[3 * j + i] * a21;  // two roundings each: product, then + 0.0f
      sR12[3 * i + j] = p + 0.0f;
      sR21[3 * i + j] = q + 0.0f;
    }
  for (int i = 0; i < 3; ++i) {
    volatile float p0 = sR21[3 * i] * t12[0], p1 = sR21[3 * i + 1] * t12[1], p2 = sR21[3 * i + 2] * t12[2];
    volatile float s01 = p0 + p1;
    volatile float s = s01 + p2;
    t21[i] = (float)((double)s * -1.0);
  }
}

int plslam_predict_scale(float max_distance, float current_dist, float log_scale_factor, int n_levels) {
  // one-thread kernel: the arithmetic is the device's (pl_logf_dev), so the veneer and the matcher kernels agree by construction
  int* d = nullptr;
  int h = -1;
  if (cudaMalloc(&d, sizeof(int)) != cudaSuccess) return -1;
  k_predict_scale<<<1, 1>>>(max_distance, current_dist, log_scale_factor, n_levels, d);
  if (cudaMemcpy(&h, d, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) h = -1;
  cudaFree(d);
  return h;
}

int plslam_match_local_points_batch_device(const plslam_local_job_t* d_jobs, int njobs, int max_n, void* stream) {
  PL_CHECK_ARG(d_jobs && njobs >= 1 && max_n >= 0);
  const size_t smem = (size_t)max_n * 5 + 16;  // matchF[n] ints + taken[n] bytes
  PL_CHECK_ARG(smem <= 200 * 1024);
  static PerDeviceOnce attr;
  if (attr.first()) {
    PL_CUDA(cudaFuncSetAttribute(k_local_points, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  PL_CARVEOUT(k_local_points);
  k_local_points<<<njobs, 32, smem, (cudaStream_t)stream>>>(d_jobs);
  PL_CUDA(cudaGetLastError());
  return PLSLAM_OK;
}

int plslam_match_local_points_host(const plslam_local_job_t* job, int n_scale_levels) {
  PL_CHECK_ARG(job && job->match_f && job->nmatches && job->m >= 0 && job->n >= 0 && n_scale_levels >= 1);
  Uploader U;
  plslam_local_job_t d = *job;
  const int m = job->m, n = job->n, ncell = PLSLAM_GRID_COLS * PLSLAM_GRID_ROWS;
  const int nitems = job->grid_start[ncell];
  d.mp_valid = U.up(job->mp_valid, m);
  d.mp_proj = U.up(job->mp_proj, (size_t)m * 3);
  d.mp_level = U.up(job->mp_level, m);
  d.mp_viewcos = U.up(job->mp_viewcos, m);
  d.mp_desc = U.up(job->mp_desc, (size_t)m * 32);
  d.mp_obs = U.up(job->mp_obs, m);
  d.f_xy = U.up(job->f_xy, (size_t)n * 2);
  d.f_octave = U.up(job->f_octave, n);
  d.f_desc = U.up(job->f_desc, (size_t)n * 32);
  d.f_uright = U.up(job->f_uright, n);
  d.f_taken = U.up(job->f_taken, n);
  d.grid_start = U.up(job->grid_start, ncell + 1);
  d.grid_items = U.up(job->grid_items, nitems);
  d.scale_factors = U.up(job->scale_factors, n_scale_levels);
  d.match_f = U.out<int32_t>(n);
  d.nmatches = U.out<int32_t>(1);
  const plslam_local_job_t* dj = U.up(&d, 1);
  if (U.err != cudaSuccess) { set_error("local points host path: %s", cudaGetErrorString(U.err)); return PLSLAM_ERR_CUDA; }
  int rc = plslam_match_local_points_batch_device(dj, 1, n, nullptr);
  if (rc) return rc;
  PL_CUDA(cudaMemcpy(job->match_f, d.match_f, (size_t)n * 4, cudaMemcpyDeviceToHost));
  PL_CUDA(cudaMemcpy(job->nmatches, d.nmatches, 4, cudaMemcpyDeviceToHost));
  return PLSLAM_OK;
}

int plslam_match_initialization_batch_device(const plslam_init_job_t* d_jobs, int njobs, int max_n1, int max_n2, void* stream) {
  PL_CHECK_ARG(d_jobs && njobs >= 1 && njobs <= 65535 && max_n1 >= 0 && max_n2 >= 0);
  const size_t smem = ((size_t)2 * max_n1 + 2 * (size_t)max_n2) * 4;
  PL_CHECK_ARG(smem <= 200 * 1024);
  static PerDeviceOnce attr;
  if (attr.first()) PL_CUDA(cudaFuncSetAttribute(k_search_init, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  PL_CARVEOUT(k_search_init);
  k_search_init<<<njobs, 32, smem, (cudaStream_t)stream>>>(d_jobs);
  PL_CUDA(cudaGetLastError());
  return PLSLAM_OK;
}

int plslam_match_initialization_host(const plslam_init_job_t* job) {
  PL_CHECK_ARG(job && job->match12 && job->nmatches && job->prev_matched && job->n1 >= 0 && job->n2 >= 0);
  Uploader U;
  plslam_init_job_t d = *job;
  const int n1 = job->n1, n2 = job->n2, ncell = PLSLAM_GRID_COLS * PLSLAM_GRID_ROWS;
  const int nitems = job->grid_start[ncell];
  d.f1_octave = U.up(job->f1_octave, n1);
  d.f1_angle = U.up(job->f1_angle, n1);
  d.f1_desc = U.up(job->f1_desc, (size_t)n1 * 32);
  d.f2_xy = U.up(job->f2_xy, (size_t)n2 * 2);
  d.f2_angle = U.up(job->f2_angle, n2);
  d.f2_octave = U.up(job->f2_octave, n2);
  d.f2_desc = U.up(job->f2_desc, (size_t)n2 * 32);
  d.grid_start = U.up(job->grid_start, ncell + 1);
  d.grid_items = U.up(job->grid_items, nitems);
  d.prev_matched = const_cast<float*>(U.up(job->prev_matched, (size_t)n1 * 2));
  d.match12 = U.out<int32_t>(n1);
  d.nmatches = U.out<int32_t>(1);
  const plslam_init_job_t* dj = U.up(&d, 1);
  if (U.err != cudaSuccess) { set_error("initialization host path: %s", cudaGetErrorString(U.err)); return PLSLAM_ERR_CUDA; }
  int rc = plslam_match_initialization_batch_device(dj, 1, n1, n2, nullptr);
  if (rc) return rc;
  PL_CUDA(cudaMemcpy(job->match12, d.match12, (size_t)n1 * 4, cudaMemcpyDeviceToHost));
  PL_CUDA(cudaMemcpy(job->prev_matched, d.prev_matched, (size_t)n1 * 8, cudaMemcpyDeviceToHost));
  PL_CUDA(cudaMemcpy(job->nmatches, d.nmatches, 4, cudaMemcpyDeviceToHost));
  return PLSLAM_OK;
}

}  // extern "C"
