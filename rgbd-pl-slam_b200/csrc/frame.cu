// frame.cu — the per-frame steps that follow extraction in ORB_SLAM2::Frame::Frame (reference include/Frame.h:60,
// lib/libORB_SLAM2.so@0xf9370), batched on the device so the extractor's keypoints never leave HBM before matching:
//   UndistortKeyPoints        Frame.h:266, call @0xfa0db   cv::undistortPoints(mat, mat, mK, mDistCoef, Mat(), mK)
//   ComputeStereoFromRGBD     Frame.h:120, call @0xfa0ea   d = imDepth.at<float>((int)kp.y, (int)kp.x) on the DISTORTED
//                                                           keypoint (@0xf6cd0-0xf6cd6); d > 0 => mvDepth = d,
//                                                           mvuRight = kpUn.x - mbf / d (@0xf6ceb-0xf6d14)
//   AssignFeaturesToGrid      Frame.h:273, call @0xfa382   PosInGrid: roundf((pt - min) * inv), 64 x 48 cells (@0xf5fa0)
//   ComputeImageBounds        Frame.h:270, call @0xfa27e   (once per run: host scalar code on 4 corner points)
// The grid is produced as the CSR ([ix][iy] order, keypoint indices ascending inside a cell, i.e. push_back order) that
// plslam_match_projection_* consumes.  One CTA per frame; undistortion in double with individually rounded operations
// (twin of oracle/frame_oracle.cc, which is pinned bit for bit against cv2 4.13's undistortPoints).
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "pl_math.cuh"

namespace plslam {
namespace {

constexpr int GC = PLSLAM_GRID_COLS, GR = PLSLAM_GRID_ROWS, NCELL = GC * GR;

struct FrameParams {
  plslam_frame_calib_t c;
  float minX, maxX, minY, maxY, wInv, hInv;
  int cols, rows, dpitch;
  size_t dstride;
  int cap;
};

// one point of cv::undistortPoints(src, dst, K, D, noArray(), K): float in, float out, double inside
__host__ __device__ inline void undistort_point(const plslam_frame_calib_t& c, float xf, float yf, float* ox, float* oy) {
#ifdef __CUDA_ARCH__
#define DM(a, b) __dmul_rn(a, b)
#define DA(a, b) __dadd_rn(a, b)
#define DS(a, b) __dsub_rn(a, b)
#define DD(a, b) __ddiv_rn(a, b)
#else
#define DM(a, b) ((a) * (b))
#define DA(a, b) ((a) + (b))
#define DS(a, b) ((a) - (b))
#define DD(a, b) ((a) / (b))
#endif
  const double fx = c.fx, fy = c.fy, cx = c.cx, cy = c.cy;
  const double ifx = DD(1., fx), ify = DD(1., fy);
  const double k0 = c.k1, k1 = c.k2, k2 = c.p1, k3 = c.p2, k4 = c.k3;
  double x = xf, y = yf;
  const double u = x, v = y;
  x = DM(DS(x, cx), ifx);
  y = DM(DS(y, cy), ify);
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; ++j) {
    const double r2 = DA(DM(x, x), DM(y, y));
    // k[5..11] are zero for the 4/5-coefficient model: the numerator is 1 + 0 and the thin-prism terms add +0.0
    const double den = DA(1., DM(DA(DM(DA(DM(k4, r2), k1), r2), k0), r2));
    const double icdist = DD(1., den);
    if (icdist < 0) {
      x = DM(DS(u, cx), ifx);
      y = DM(DS(v, cy), ify);
      break;
    }
    const double deltaX = DA(DM(DM(DM(2., k2), x), y), DM(k3, DA(r2, DM(DM(2., x), x))));
    const double deltaY = DA(DM(k2, DA(r2, DM(DM(2., y), y))), DM(DM(DM(2., k3), x), y));
    x = DM(DS(x0, deltaX), icdist);
    y = DM(DS(y0, deltaY), icdist);
  }
  *ox = (float)DA(DM(fx, x), cx);
  *oy = (float)DA(DM(fy, y), cy);
#undef DM
#undef DA
#undef DS
#undef DD
}

__device__ __forceinline__ int cell_of(const FrameParams& P, float x, float y) {
  const int px = (int)roundf(__fmul_rn(__fsub_rn(x, P.minX), P.wInv));
  const int py = (int)roundf(__fmul_rn(__fsub_rn(y, P.minY), P.hInv));
  if (px < 0 || px >= GC || py < 0 || py >= GR) return -1;
  return px * GR + py;
}

__global__ void __launch_bounds__(256) k_frame_post(const __grid_constant__ FrameParams P,
                                                    const plslam_keypoint_t* __restrict__ kps,
                                                    const int32_t* __restrict__ counts, const float* __restrict__ depth,
                                                    float2* __restrict__ un_xy, float* __restrict__ uright,
                                                    float* __restrict__ zdepth, int32_t* __restrict__ grid_start,
                                                    int32_t* __restrict__ grid_items) {
  __shared__ int cnt[NCELL];
  __shared__ int fill[NCELL];
  __shared__ int warp_tmp[33];
  const int f = blockIdx.x, t = threadIdx.x;
  const int n = min(counts[f], P.cap);
  const plslam_keypoint_t* K = kps + (size_t)f * P.cap;
  float2* U = un_xy + (size_t)f * P.cap;
  float* R = uright + (size_t)f * P.cap;
  float* Z = zdepth + (size_t)f * P.cap;
  const float* D = depth + (size_t)f * P.dstride;
  int32_t* GS = grid_start + (size_t)f * (NCELL + 1);
  int32_t* GI = grid_items + (size_t)f * P.cap;
  for (int c = t; c < NCELL; c += 256) {
    cnt[c] = 0;
    fill[c] = 0;
  }
  __syncthreads();
  const bool distorted = P.c.k1 != 0.0f;  // mDistCoef.at<float>(0) == 0.0 -> mvKeysUn = mvKeys
  for (int i = t; i < n; i += 256) {
    const float x = K[i].x, y = K[i].y;
    float ux = x, uy = y;
    if (distorted) undistort_point(P.c, x, y, &ux, &uy);
    U[i] = make_float2(ux, uy);
    const int v = (int)y, u = (int)x;
    float d = 0.f;
    if (u >= 0 && v >= 0 && u < P.cols && v < P.rows) d = __ldg(D + (size_t)v * P.dpitch + u);
    float z = -1.f, r = -1.f;
    if (d > 0) {
      z = d;
      r = __fsub_rn(ux, __fdiv_rn(P.c.bf, d));
    }
    Z[i] = z;
    R[i] = r;
    const int c = cell_of(P, ux, uy);
    if (c >= 0) atomicAdd(&cnt[c], 1);
  }
  __syncthreads();
  const int total = block_scan_excl(cnt, NCELL, warp_tmp);  // cnt[c] = start of cell c
  for (int c = t; c < NCELL; c += 256) GS[c] = cnt[c];
  if (t == 0) GS[NCELL] = total;
  __syncthreads();
  for (int i = t; i < n; i += 256) {
    const float2 p = U[i];
    const int c = cell_of(P, p.x, p.y);
    if (c >= 0) GI[cnt[c] + atomicAdd(&fill[c], 1)] = i;
  }
  __syncthreads();
  // push_back order inside a cell = ascending keypoint index
  for (int c = t; c < NCELL; c += 256) {
    const int b = cnt[c], e = b + fill[c];
    for (int i = b + 1; i < e; ++i) {
      const int v = GI[i];
      int j = i - 1;
      while (j >= b && GI[j] > v) {
        GI[j + 1] = GI[j];
        --j;
      }
      GI[j + 1] = v;
    }
  }
}

int make_params(const plslam_frame_calib_t* calib, const float* bounds4, int cols, int rows, int dpitch, size_t dstride,
                int cap, FrameParams* P) {
  PL_CHECK_ARG(calib && bounds4 && cols > 0 && rows > 0 && dpitch >= cols && cap > 0);
  PL_CHECK_ARG(bounds4[1] > bounds4[0] && bounds4[3] > bounds4[2]);
  P->c = *calib;
  P->minX = bounds4[0];
  P->maxX = bounds4[1];
  P->minY = bounds4[2];
  P->maxY = bounds4[3];
  // Frame::Frame: mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (mnMaxX - mnMinX)
  P->wInv = (float)GC / (bounds4[1] - bounds4[0]);
  P->hInv = (float)GR / (bounds4[3] - bounds4[2]);
  P->cols = cols;
  P->rows = rows;
  P->dpitch = dpitch;
  P->dstride = dstride;
  P->cap = cap;
  return PLSLAM_OK;
}

// ------------------------------------------------------------------------------------------
// Frame::isInFrustum (@0xf5190) over the local map points of a frame: thread = map point.  Arithmetic as read from the binary
// (see oracle/match_oracle.cc oracle_is_in_frustum): gemm small path for Pc, float division, fused projections, cv::norm and
// Mat::dot in double, PredictScale with the C library's logf.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_is_in_frustum(const plslam_frustum_job_t* __restrict__ jobs) {
  const plslam_frustum_job_t& J = jobs[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= J.m) return;
  const float fx = J.cam[0], fy = J.cam[1], cx = J.cam[2], cy = J.cam[3];
  const float mnMinX = J.cam[4], mnMaxX = J.cam[5], mnMinY = J.cam[6], mnMaxY = J.cam[7];
  uint8_t inView = 0;
  float u = 0.f, v = 0.f, ur = 0.f, vc = 0.f;
  int level = 0;
  do {
    const float* X = J.mp_xyz + 3 * (size_t)i;
    float pc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float p0 = __fmul_rn(J.tcw[r * 4], X[0]), p1 = __fmul_rn(J.tcw[r * 4 + 1], X[1]), p2 = __fmul_rn(J.tcw[r * 4 + 2], X[2]);
      pc[r] = (float)__dadd_rn((double)__fadd_rn(__fadd_rn(p0, p1), p2), (double)J.tcw[r * 4 + 3]);
    }
    if (pc[2] < 0.0f) break;
    const float invz = __fdiv_rn(1.0f, pc[2]);
    const float uu = __fmaf_rn(__fmul_rn(pc[0], fx), invz, cx);
    if (uu < mnMinX || uu > mnMaxX) break;
    const float vv = __fmaf_rn(__fmul_rn(pc[1], fy), invz, cy);
    if (vv < mnMinY || vv > mnMaxY) break;
    float PO[3];
    double n2 = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      PO[r] = __fsub_rn(X[r], J.ow[r]);
      n2 = __dadd_rn(n2, __dmul_rn((double)PO[r], (double)PO[r]));
    }
    const float dist = (float)sqrt(n2);
    const float dMin = J.mp_dist_range[2 * (size_t)i], dMax = J.mp_dist_range[2 * (size_t)i + 1];
    if (__fmul_rn(0.8f, dMin) > dist || dist > __fmul_rn(1.2f, dMax)) break;
    double dot = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) dot = __dadd_rn(dot, __dmul_rn((double)PO[r], (double)J.mp_normal[3 * (size_t)i + r]));
    const float c = (float)__ddiv_rn(dot, (double)dist);
    if (J.viewing_cos_limit > c) break;
    level = predict_scale_dev(dMax, dist, J.log_scale_factor, J.n_levels);
    inView = 1;
    u = uu;
    v = vv;
    ur = __fmaf_rn(-invz, J.mbf, uu);
    vc = c;
  } while (false);
  J.in_view[i] = inView;
  J.proj[3 * (size_t)i] = u;
  J.proj[3 * (size_t)i + 1] = v;
  J.proj[3 * (size_t)i + 2] = ur;
  J.level[i] = level;
  J.viewcos[i] = vc;
}

// ------------------------------------------------------------------------------------------
// Line analogues of the Frame steps (Frame.h:267, :116, :107; parity unpinned, see the header and oracle/frame_oracle.cc).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_undistort_points(const __grid_constant__ plslam_frame_calib_t c, const float* __restrict__ xy,
                                                          int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (c.k1 == 0.0f) {
    out[2 * i] = xy[2 * i];
    out[2 * i + 1] = xy[2 * i + 1];
  } else {
    undistort_point(c, xy[2 * i], xy[2 * i + 1], out + 2 * i, out + 2 * i + 1);
  }
}

// GetLinesInArea: warp per query, lanes over the key lines 32 at a time, candidates appended in index order (ballot ranks).
// Pass 0 counts (out_items == nullptr), pass 1 writes at out_start[q].
__global__ void __launch_bounds__(256) k_lines_in_area(const float* __restrict__ queries7, int nq, const float* __restrict__ lines4, int n,
                                                       int* __restrict__ counts, const int* __restrict__ out_start,
                                                       int* __restrict__ out_items) {
  const int lane = threadIdx.x & 31, q = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (q >= nq) return;
  const float* Q = queries7 + 7 * (size_t)q;
  const float x1 = Q[0], y1 = Q[1], x2 = Q[2], y2 = Q[3], r = Q[4];
  const int minLevel = (int)Q[5], maxLevel = (int)Q[6];
  const bool bCheckLevels = (minLevel > 0) || (maxLevel > 0);
  const float mx = __fmul_rn(0.5f, __fadd_rn(x1, x2)), my = __fmul_rn(0.5f, __fadd_rn(y1, y2));
  const float r2 = __fmul_rn(r, r), slopeMax = __fmul_rn(r, 0.01f);
  const float slopeQ = __fdiv_rn(__fsub_rn(y1, y2), __fsub_rn(x1, x2));
  int cnt = 0;
  for (int b = 0; b < n; b += 32) {
    const int i = b + lane;
    bool ok = false;
    if (i < n) {
      const float4 L = reinterpret_cast<const float4*>(lines4)[i];
      const float dx = __fsub_rn(mx, L.x), dy = __fsub_rn(my, L.y);
      const float distance = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
      ok = !(distance > r2) && !(__fsub_rn(slopeQ, L.z) > slopeMax);
      if (ok && bCheckLevels) {
        const int oct = (int)L.w;
        if (oct < minLevel || (maxLevel >= 0 && oct > maxLevel)) ok = false;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (ok && out_items) out_items[out_start[q] + cnt + __popc(m & ((1u << lane) - 1u))] = i;
    cnt += __popc(m);
  }
  if (lane == 0 && counts) counts[q] = cnt;
}

__global__ void __launch_bounds__(256) k_line_in_frustum(const __grid_constant__ plslam_line_frustum_job_t J) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= J.m) return;
  const float fx = J.cam[0], fy = J.cam[1], cx = J.cam[2], cy = J.cam[3];
  const float mnMinX = J.cam[4], mnMaxX = J.cam[5], mnMinY = J.cam[6], mnMaxY = J.cam[7];
  uint8_t inView = 0;
  float pr[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, vc = 0.f;
  int level = 0;
  do {
    float uv[6];
    bool ok = true;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float* X = J.ml_sp_ep + 6 * (size_t)i + 3 * e;
      float pc[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float p0 = __fmul_rn(J.tcw[r * 4], X[0]), p1 = __fmul_rn(J.tcw[r * 4 + 1], X[1]), p2 = __fmul_rn(J.tcw[r * 4 + 2], X[2]);
        pc[r] = (float)__dadd_rn((double)__fadd_rn(__fadd_rn(p0, p1), p2), (double)J.tcw[r * 4 + 3]);
      }
      if (pc[2] < 0.0f) { ok = false; break; }
      const float invz = __fdiv_rn(1.0f, pc[2]);
      const float u = __fmaf_rn(__fmul_rn(pc[0], fx), invz, cx), v = __fmaf_rn(__fmul_rn(pc[1], fy), invz, cy);
      if (u < mnMinX || u > mnMaxX || v < mnMinY || v > mnMaxY) { ok = false; break; }
      uv[3 * e] = u; uv[3 * e + 1] = v; uv[3 * e + 2] = __fmaf_rn(-invz, J.mbf, u);
    }
    if (!ok) break;
    float OM[3];
    double n2 = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      OM[r] = __fsub_rn(__fmul_rn(0.5f, __fadd_rn(J.ml_sp_ep[6 * (size_t)i + r], J.ml_sp_ep[6 * (size_t)i + 3 + r])), J.ow[r]);
      n2 = __dadd_rn(n2, __dmul_rn((double)OM[r], (double)OM[r]));
    }
    const float dist = (float)sqrt(n2);
    const float dMin = J.ml_dist_range[2 * (size_t)i], dMax = J.ml_dist_range[2 * (size_t)i + 1];
    if (dist < __fmul_rn(0.8f, dMin) || dist > __fmul_rn(1.2f, dMax)) break;
    double dot = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) dot = __dadd_rn(dot, __dmul_rn((double)OM[r], (double)J.ml_normal[3 * (size_t)i + r]));
    const float c = (float)__ddiv_rn(dot, (double)dist);
    if (c < J.viewing_cos_limit) break;
    level = predict_scale_dev(dMax, dist, J.log_scale_factor, J.n_levels);
    inView = 1;
#pragma unroll
    for (int k = 0; k < 6; ++k) pr[k] = uv[k];
    vc = c;
  } while (false);
  J.in_view[i] = inView;
#pragma unroll
  for (int k = 0; k < 6; ++k) J.proj[6 * (size_t)i + k] = pr[k];
  J.level[i] = level;
  J.viewcos[i] = vc;
}

}  // namespace
}  // namespace plslam

using namespace plslam;

extern "C" {

int plslam_frame_image_bounds(const plslam_frame_calib_t* calib, int cols, int rows, float bounds4[4]) {
  PL_CHECK_ARG(calib && bounds4 && cols > 0 && rows > 0);
  if (calib->k1 != 0.0f) {
    const float corners[8] = {0.f, 0.f, (float)cols, 0.f, 0.f, (float)rows, (float)cols, (float)rows};
    float m[8];
    for (int i = 0; i < 4; ++i) undistort_point(*calib, corners[2 * i], corners[2 * i + 1], &m[2 * i], &m[2 * i + 1]);
    bounds4[0] = std::min(m[0], m[4]);
    bounds4[1] = std::max(m[2], m[6]);
    bounds4[2] = std::min(m[1], m[3]);
    bounds4[3] = std::max(m[5], m[7]);
  } else {
    bounds4[0] = 0.f;
    bounds4[1] = (float)cols;
    bounds4[2] = 0.f;
    bounds4[3] = (float)rows;
  }
  return PLSLAM_OK;
}

int plslam_frame_post_batch_device(const plslam_frame_calib_t* calib, const float bounds4[4],
                                   const plslam_keypoint_t* d_keypoints, const int32_t* d_counts, int batch,
                                   int kp_capacity, const float* d_depth, int cols, int rows, int depth_pitch,
                                   size_t depth_frame_stride, float* d_un_xy, float* d_uright, float* d_depth_out,
                                   int32_t* d_grid_start, int32_t* d_grid_items, void* stream) {
  PL_CHECK_ARG(d_keypoints && d_counts && d_depth && d_un_xy && d_uright && d_depth_out && d_grid_start && d_grid_items);
  PL_CHECK_ARG(batch >= 1 && batch <= 65535);
  FrameParams P;
  int rc = make_params(calib, bounds4, cols, rows, depth_pitch, depth_frame_stride, kp_capacity, &P);
  if (rc) return rc;
  PL_CARVEOUT(k_frame_post);
  k_frame_post<<<batch, 256, 0, (cudaStream_t)stream>>>(P, d_keypoints, d_counts, d_depth,
                                                        reinterpret_cast<float2*>(d_un_xy), d_uright, d_depth_out,
                                                        d_grid_start, d_grid_items);
  PL_CUDA(cudaGetLastError());
  return PLSLAM_OK;
}

int plslam_frame_post_host(const plslam_frame_calib_t* calib, const float bounds4[4], const plslam_keypoint_t* keypoints,
                           int n, const float* depth, int cols, int rows, int depth_pitch, float* un_xy, float* uright,
                           float* depth_out, int32_t* grid_start, int32_t* grid_items) {
  PL_CHECK_ARG(keypoints && depth && un_xy && uright && depth_out && grid_start && grid_items && n >= 0);
  const int cap = std::max(n, 1);
  struct Scoped : DevBuf {
    ~Scoped() { release(); }
  } dk, dc, dd, du, dr, dz, gs, gi;
  int rc;
  if ((rc = dk.ensure((size_t)cap * sizeof(plslam_keypoint_t))) || (rc = dc.ensure(4)) ||
      (rc = dd.ensure((size_t)rows * depth_pitch * 4)) || (rc = du.ensure((size_t)cap * 8)) ||
      (rc = dr.ensure((size_t)cap * 4)) || (rc = dz.ensure((size_t)cap * 4)) ||
      (rc = gs.ensure((size_t)(NCELL + 1) * 4)) || (rc = gi.ensure((size_t)cap * 4)))
    return rc;
  const int32_t cnt = n;
  if (n) PL_CUDA(cudaMemcpy(dk.p, keypoints, (size_t)n * sizeof(plslam_keypoint_t), cudaMemcpyHostToDevice));
  PL_CUDA(cudaMemcpy(dc.p, &cnt, 4, cudaMemcpyHostToDevice));
  PL_CUDA(cudaMemcpy(dd.p, depth, (size_t)rows * depth_pitch * 4, cudaMemcpyHostToDevice));
  rc = plslam_frame_post_batch_device(calib, bounds4, dk.as<plslam_keypoint_t>(), dc.as<int32_t>(), 1, cap, dd.as<float>(),
                                      cols, rows, depth_pitch, 0, du.as<float>(), dr.as<float>(), dz.as<float>(),
                                      gs.as<int32_t>(), gi.as<int32_t>(), nullptr);
  if (rc) return rc;
  PL_CUDA(cudaDeviceSynchronize());
  if (n) {
    PL_CUDA(cudaMemcpy(un_xy, du.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    PL_CUDA(cudaMemcpy(uright, dr.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    PL_CUDA(cudaMemcpy(depth_out, dz.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  }
  PL_CUDA(cudaMemcpy(grid_start, gs.p, (size_t)(NCELL + 1) * 4, cudaMemcpyDeviceToHost));
  if (n) PL_CUDA(cudaMemcpy(grid_items, gi.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return PLSLAM_OK;
}

int plslam_frame_undistort_keylines_host(const plslam_frame_calib_t* calib, const float* xy4, int n, float* out_xy4) {
  PL_CHECK_ARG(calib && n >= 0);
  if (n == 0) return PLSLAM_OK;
  PL_CHECK_ARG(xy4 && out_xy4);
  DevBuf in, out;
  int rc;
  if ((rc = in.ensure((size_t)n * 16)) || (rc = out.ensure((size_t)n * 16))) { in.release(); out.release(); return rc; }
  cudaError_t e = cudaMemcpy(in.p, xy4, (size_t)n * 16, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    k_undistort_points<<<div_up(2 * n, 256), 256>>>(*calib, in.as<float>(), 2 * n, out.as<float>());
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out_xy4, out.p, (size_t)n * 16, cudaMemcpyDeviceToHost);
  in.release(); out.release();
  if (e != cudaSuccess) { set_error("undistort_keylines: %s", cudaGetErrorString(e)); return PLSLAM_ERR_CUDA; }
  return PLSLAM_OK;
}

int plslam_frame_lines_in_area_host(const float* queries7, int nq, const float* lines4, int n, int32_t* out_start,
                                    int32_t* out_items, int item_cap) {
  PL_CHECK_ARG(nq >= 0 && n >= 0 && out_start && item_cap >= 0);
  out_start[0] = 0;
  if (nq == 0) return PLSLAM_OK;
  PL_CHECK_ARG(queries7 && (n == 0 || lines4) && (item_cap == 0 || out_items));
  DevBuf dq, dl, dc, ds, di;
  int rc = PLSLAM_OK;
  cudaError_t e = cudaSuccess;
  std::vector<int> cnt(nq, 0);
  if ((rc = dq.ensure((size_t)nq * 28)) || (rc = dl.ensure((size_t)std::max(n, 1) * 16)) || (rc = dc.ensure((size_t)nq * 4)) ||
      (rc = ds.ensure((size_t)(nq + 1) * 4))) goto done;
  e = cudaMemcpy(dq.p, queries7, (size_t)nq * 28, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && n) e = cudaMemcpy(dl.p, lines4, (size_t)n * 16, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) goto done;
  k_lines_in_area<<<div_up(nq, 8), 256>>>(dq.as<float>(), nq, dl.as<float>(), n, dc.as<int>(), nullptr, nullptr);
  e = cudaMemcpy(cnt.data(), dc.p, (size_t)nq * 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) goto done;
  for (int q = 0; q < nq; ++q) out_start[q + 1] = out_start[q] + cnt[q];
  if (out_start[nq] > item_cap) { rc = PLSLAM_ERR_CAPACITY; set_error("lines_in_area: %d candidates, capacity %d", out_start[nq], item_cap); goto done; }
  if (out_start[nq] > 0) {
    if ((rc = di.ensure((size_t)out_start[nq] * 4))) goto done;
    e = cudaMemcpy(ds.p, out_start, (size_t)(nq + 1) * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) goto done;
    k_lines_in_area<<<div_up(nq, 8), 256>>>(dq.as<float>(), nq, dl.as<float>(), n, nullptr, ds.as<int>(), di.as<int>());
    e = cudaMemcpy(out_items, di.p, (size_t)out_start[nq] * 4, cudaMemcpyDeviceToHost);
  }
done:
  dq.release(); dl.release(); dc.release(); ds.release(); di.release();
  if (e != cudaSuccess) { set_error("lines_in_area: %s", cudaGetErrorString(e)); return PLSLAM_ERR_CUDA; }
  return rc;
}

int plslam_frame_line_in_frustum_host(const plslam_line_frustum_job_t* job) {
  PL_CHECK_ARG(job && job->m >= 0 && job->n_levels >= 1);
  const int m = job->m;
  if (m == 0) return PLSLAM_OK;
  PL_CHECK_ARG(job->ml_sp_ep && job->ml_normal && job->ml_dist_range && job->in_view && job->proj && job->level && job->viewcos);
  DevBuf in, out;
  int rc;
  if ((rc = in.ensure((size_t)m * 11 * 4)) || (rc = out.ensure((size_t)m * (24 + 4 + 4 + 1) + 64))) { in.release(); out.release(); return rc; }
  float* dIn = in.as<float>();
  plslam_line_frustum_job_t d = *job;
  d.ml_sp_ep = dIn; d.ml_normal = dIn + 6 * (size_t)m; d.ml_dist_range = dIn + 9 * (size_t)m;
  d.proj = out.as<float>(); d.viewcos = d.proj + 6 * (size_t)m;
  d.level = reinterpret_cast<int32_t*>(d.viewcos + m); d.in_view = reinterpret_cast<uint8_t*>(d.level + m);
  cudaError_t e = cudaMemcpy(dIn, job->ml_sp_ep, (size_t)m * 24, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dIn + 6 * (size_t)m, job->ml_normal, (size_t)m * 12, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dIn + 9 * (size_t)m, job->ml_dist_range, (size_t)m * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    k_line_in_frustum<<<div_up(m, 256), 256>>>(d);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(job->proj, d.proj, (size_t)m * 24, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(job->viewcos, d.viewcos, (size_t)m * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(job->level, d.level, (size_t)m * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(job->in_view, d.in_view, (size_t)m, cudaMemcpyDeviceToHost);
  in.release(); out.release();
  if (e != cudaSuccess) { set_error("line_in_frustum: %s", cudaGetErrorString(e)); return PLSLAM_ERR_CUDA; }
  return PLSLAM_OK;
}

int plslam_frame_is_in_frustum_batch_device(const plslam_frustum_job_t* d_jobs, int njobs, int max_m, void* stream) {
  PL_CHECK_ARG(d_jobs && njobs >= 1 && njobs <= 65535 && max_m >= 0);
  if (max_m == 0) return PLSLAM_OK;
  PL_CARVEOUT(k_is_in_frustum);
  k_is_in_frustum<<<dim3(div_up(max_m, 256), njobs), 256, 0, (cudaStream_t)stream>>>(d_jobs);
  PL_CUDA(cudaGetLastError());
  return PLSLAM_OK;
}

int plslam_frame_is_in_frustum_host(const plslam_frustum_job_t* job) {
  PL_CHECK_ARG(job && job->m >= 0 && job->n_levels >= 1);
  const int m = job->m;
  if (m == 0) return PLSLAM_OK;
  PL_CHECK_ARG(job->mp_xyz && job->mp_normal && job->mp_dist_range && job->in_view && job->proj && job->level && job->viewcos);
  DevBuf in, out, dj;
  const size_t inBytes = (size_t)m * 8 * sizeof(float), outBytes = (size_t)m * (1 + 12 + 4 + 4);
  int rc;
  if ((rc = in.ensure(inBytes)) || (rc = out.ensure(align_up(outBytes, 16) + 64)) || (rc = dj.ensure(sizeof(plslam_frustum_job_t)))) {
    in.release(); out.release(); dj.release();
    return rc;
  }
  float* dIn = in.as<float>();
  plslam_frustum_job_t d = *job;
  d.mp_xyz = dIn; d.mp_normal = dIn + 3 * (size_t)m; d.mp_dist_range = dIn + 6 * (size_t)m;
  d.proj = out.as<float>(); d.viewcos = d.proj + 3 * (size_t)m;
  d.level = reinterpret_cast<int32_t*>(d.viewcos + m); d.in_view = reinterpret_cast<uint8_t*>(d.level + m);
  cudaError_t e = cudaMemcpy(dIn, job->mp_xyz, (size_t)m * 12, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dIn + 3 * (size_t)m, job->mp_normal, (size_t)m * 12, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dIn + 6 * (size_t)m, job->mp_dist_range, (size_t)m * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dj.p, &d, sizeof(d), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = plslam_frame_is_in_frustum_batch_device(dj.as<plslam_frustum_job_t>(), 1, m, nullptr);
    if (rc) { in.release(); out.release(); dj.release(); return rc; }
    e = cudaMemcpy(job->proj, d.proj, (size_t)m * 12, cudaMemcpyDeviceToHost);
  }
  if (e == cudaSuccess) e = cudaMemcpy(job->viewcos, d.viewcos, (size_t)m * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(job->level, d.level, (size_t)m * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(job->in_view, d.in_view, (size_t)m, cudaMemcpyDeviceToHost);
  in.release(); out.release(); dj.release();
  if (e != cudaSuccess) {
    set_error("is_in_frustum host path: %s", cudaGetErrorString(e));
    return PLSLAM_ERR_CUDA;
  }
  return PLSLAM_OK;
}

}  // extern "C"
