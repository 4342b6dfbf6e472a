// lines.cu — B200-native line front-end: LSD (LSD_REFINE_ADV) + KeyLine filling + LBD descriptors.
//
// Replaces ORB_SLAM2::LineSegment::ExtractLineSegment (reference include/ExtractLineSegment.h:38),
// whose arithmetic lives in OpenCV-contrib line_descriptor (LSDDetector::detect ->
// imgproc LineSegmentDetector, BinaryDescriptor::compute).  The algorithm restated here is the one
// pinned by oracle/lsd_oracle.cc (bit-identical to cv2 4.13 for LSD in its libm mode); this file
// reproduces the oracle's PINNED mode bit for bit.
//
// Stage map (one launch per stage for the whole batch):
//   k_lsd_scale    7x7 sigma-0.75 Gaussian (dp4a) + x0.8 INTER_LINEAR_EXACT, fused through shared memory
//   k_lsd_grad     2x2 gradient, level-line angle (fastAtan2), per-pixel record {deg, cos, sin, g2} + 4-byte angle plane
//   k_lsd_rowhist / k_lsd_colscan / k_lsd_scatter   stable counting sort of the seeds (bin desc, raster)
//   k_lsd_grow     region growing + rectangle fit + density refinement, one warp per frame: the greedy growth is
//                  sequential inside a frame (running region angle, shared `used` map).  Lanes = 4 frontier points x 8
//                  neighbours with in-batch speculation; the `used` flag lives in the pixel records.  This is the mode
//                  for large batches (many batches in flight fill the machine).
//   k_lsd_grow_sw  the same result with one CTA of 8 / 16 / 32 warps per frame: regions grown speculatively in parallel
//                  over an owner plane and retired in seed order through a window (no round barriers)
//   k_lsd_nfa      rect_improve / rect_nfa: persistent grid, one warp per rectangle (independent of `used`)
//   k_lsd_finish   ordered compaction, KeyLine fields, strongest-N selection, line equations
//   k_lbd          LBD band descriptors, one CTA per kept line
#include "lines.cuh"

#include <algorithm>
#include <cmath>

#include "pl_math.cuh"

namespace plslam {

namespace {

constexpr float NOTDEF_F = -1024.f;
constexpr double M_3_2_PI_D = (3 * PL_PI) / 2;
constexpr double M_2__PI_D = 2 * PL_PI;

// ------------------------------------------------------------------------------------------
// k_lsd_scale: GaussianBlur(7x7, sigma 0.75) then resize(x0.8, INTER_LINEAR_EXACT) (lsd.cpp flsd()).
// Both are OpenCV's 8-bit fixed-point paths: blur = (sum k k p + 32768) >> 16 with the table
// {0,4,56,136,56,4,0}; resize uses 8.8 coefficients, exact products and one final rounding.
// Tile = 64x16 scaled pixels; the 82x22 blurred source patch never leaves shared memory.
// ------------------------------------------------------------------------------------------
constexpr int ST_W = 64, ST_H = 16, SB_W = 84, SB_H = 24;
constexpr int RAW_P = 96;  // bytes per row of the raw patch (word aligned; >= SB_W + 6 + 3 bytes of over-read)
constexpr int HS_P = 88;   // u16 per row of the horizontally filtered patch

__global__ void __launch_bounds__(256) k_lsd_scale(const __grid_constant__ LineParams L, const uint8_t* __restrict__ img,
                                                   int pitch, size_t frame_stride, const int* __restrict__ coef,
                                                   uint8_t* __restrict__ scaled) {
  __shared__ __align__(16) uint8_t raw[SB_H + 6][RAW_P];
  __shared__ __align__(16) unsigned short hs[SB_H + 6][HS_P];
  __shared__ uint8_t bl[SB_H][SB_W];
  const int f = blockIdx.z;
  const int ox0 = blockIdx.x * ST_W, oy0 = blockIdx.y * ST_H;
  const int* xofs = coef;
  const int* xc1 = coef + L.sw;
  const int* yofs = xc1 + L.sw;
  const int* yc1 = yofs + L.sh;
  const int oxl = min(ox0 + ST_W, L.sw) - 1, oyl = min(oy0 + ST_H, L.sh) - 1;
  const int bx0 = xofs[ox0], bx1 = min(xofs[oxl] + 1, L.W - 1);
  const int by0 = yofs[oy0], by1 = min(yofs[oyl] + 1, L.H - 1);
  const int nbx = bx1 - bx0 + 1, nby = by1 - by0 + 1;
  const uint8_t* S = img + (size_t)f * frame_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // raw patch (3-px halo, BORDER_REFLECT_101 at the image edges): warp = row, lanes along x.  Tiles whose columns lie inside
  // the image (6 of 8 per row at 640x480) take the rows as aligned words and realign them with a funnel shift; the others, and
  // images whose rows are not word aligned, go byte by byte through the reflection.
  const int cx0 = bx0 - 3;
  if (cx0 >= 0 && cx0 + nbx + 6 <= L.W && ((pitch | (int)(frame_stride & 3) | (int)(reinterpret_cast<uintptr_t>(img) & 3)) & 3) == 0) {
    const int xa = cx0 & ~3, sft = 8 * (cx0 - xa);
    for (int r = warp; r < nby + 6; r += 8) {
      const uint8_t* row = S + (size_t)reflect101_dev(by0 - 3 + r, L.H) * pitch;
      // 25 source words cover the 24 words of the patch row at any shift; words past the pitch are never consumed
      const uint32_t w0 = (lane <= RAW_P / 4 && xa + 4 * lane + 3 < pitch) ? __ldg(reinterpret_cast<const uint32_t*>(row + xa) + lane) : 0u;
      const uint32_t w1 = __shfl_down_sync(0xffffffffu, w0, 1);
      if (lane < RAW_P / 4) reinterpret_cast<uint32_t*>(raw[r])[lane] = __funnelshift_r(w0, w1, sft);
    }
  } else {
    for (int r = warp; r < nby + 6; r += 8) {
      const uint8_t* row = S + (size_t)reflect101_dev(by0 - 3 + r, L.H) * pitch;
      for (int c = lane; c < RAW_P; c += 32)
        raw[r][c] = c < nbx + 6 ? row[reflect101_dev(bx0 - 3 + c, L.W)] : (uint8_t)0;
    }
  }
  __syncthreads();
  // horizontal 7-tap pass, 4 outputs per thread: aligned words, funnel shifts for the byte windows, dp4a for the taps
  const unsigned K0 = (unsigned)L.blurk[0] | ((unsigned)L.blurk[1] << 8) | ((unsigned)L.blurk[2] << 16) | ((unsigned)L.blurk[3] << 24);
  const unsigned K1 = (unsigned)L.blurk[4] | ((unsigned)L.blurk[5] << 8) | ((unsigned)L.blurk[6] << 16);
  const int nwq = (nbx + 3) >> 2;
  for (int r = warp; r < nby + 6; r += 8)
    for (int wq = lane; wq < nwq; wq += 32) {
      const uint32_t* R = reinterpret_cast<const uint32_t*>(raw[r]) + wq;
      const unsigned w0 = R[0], w1 = R[1], w2 = R[2];
      const unsigned h0 = __dp4a(w0, K0, __dp4a(w1, K1, 0u));
      const unsigned h1 = __dp4a(__funnelshift_r(w0, w1, 8), K0, __dp4a(__funnelshift_r(w1, w2, 8), K1, 0u));
      const unsigned h2 = __dp4a(__funnelshift_r(w0, w1, 16), K0, __dp4a(__funnelshift_r(w1, w2, 16), K1, 0u));
      const unsigned h3 = __dp4a(__funnelshift_r(w0, w1, 24), K0, __dp4a(__funnelshift_r(w1, w2, 24), K1, 0u));
      uint2 o;
      o.x = h0 | (h1 << 16);
      o.y = h2 | (h3 << 16);
      *reinterpret_cast<uint2*>(&hs[r][4 * wq]) = o;
    }
  __syncthreads();
  // vertical pass: thread = (column, segment of 8 rows), 14-value register window
  {
    const int k0 = L.blurk[0], k1 = L.blurk[1], k2 = L.blurk[2], k3 = L.blurk[3], k4 = L.blurk[4], k5 = L.blurk[5], k6 = L.blurk[6];
    for (int it = threadIdx.x; it < 3 * SB_W; it += 256) {
      const int seg = it / SB_W, c = it - seg * SB_W;
      const int r0 = seg * 8;
      if (c >= nbx || r0 >= nby) continue;
      int v[14];
#pragma unroll
      for (int j = 0; j < 14; ++j) v[j] = r0 + j < nby + 6 ? hs[r0 + j][c] : 0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (r0 + j < nby) {
          const int a = 32768 + k0 * v[j] + k1 * v[j + 1] + k2 * v[j + 2] + k3 * v[j + 3] + k4 * v[j + 4] + k5 * v[j + 5] + k6 * v[j + 6];
          bl[r0 + j][c] = (uint8_t)(a >> 16);
        }
    }
  }
  __syncthreads();
  uint8_t* D = scaled + (size_t)f * L.spitch * L.sh;
  // thread = 4 consecutive output pixels of one row (16 rows x 16 quads = the CTA), one word store when the quad is complete
  {
    const int ty = threadIdx.x >> 4, tx = (threadIdx.x & 15) * 4;
    const int oy = oy0 + ty;
    if (oy < L.sh && ox0 + tx < L.sw) {
      const int sy = __ldg(yofs + oy), sy1 = min(sy + 1, L.H - 1);
      const int b1 = __ldg(yc1 + oy), b0 = 256 - b1;
      const uint8_t* r0 = bl[sy - by0];
      const uint8_t* r1 = bl[sy1 - by0];
      uint32_t out = 0;
      const int nq = min(4, L.sw - (ox0 + tx));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ox = min(ox0 + tx + j, L.sw - 1);
        const int sx = __ldg(xofs + ox), sx1 = min(sx + 1, L.W - 1);
        const int a1 = __ldg(xc1 + ox), a0 = 256 - a1;
        const int t0 = r0[sx - bx0] * a0 + r0[sx1 - bx0] * a1;
        const int t1 = r1[sx - bx0] * a0 + r1[sx1 - bx0] * a1;
        out |= (uint32_t)(uint8_t)((t0 * b0 + t1 * b1 + 32768) >> 16) << (8 * j);
      }
      uint8_t* dp = D + (size_t)oy * L.spitch + ox0 + tx;
      if (nq == 4) {
        *reinterpret_cast<uint32_t*>(dp) = out;  // spitch is a multiple of 32, ox0 + tx of 4
      } else {
        for (int j = 0; j < nq; ++j) dp[j] = (uint8_t)(out >> (8 * j));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// k_lsd_grad: ll_angle() of lsd.cpp.  Per scaled pixel a 16-byte record:
//   .x level-line angle in degrees (fastAtan2(gx, -gy)) or NOTDEF, .y/.z cos/sin of float(angle in rad)
//   (what region_grow adds to its running direction), .w gx^2 + gy^2 (modgrad = sqrt(.w / 4)).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double modgrad_of(int g2) { return sqrt(__dmul_rn((double)g2, 0.25)); }

constexpr int GRAD_TW = 128, GRAD_TH = 8;  // pixels of a CTA's tile: warp = row, lane = 4 pixels 32 apart
__global__ void __launch_bounds__(256) k_lsd_grad(const __grid_constant__ LineParams L, const uint8_t* __restrict__ scaled,
                                                  uint4* __restrict__ pix, float* __restrict__ degPlane,
                                                  unsigned* __restrict__ g2Plane, int* __restrict__ maxg2,
                                                  unsigned* __restrict__ bmAll) {
  // Only ~1 pixel in 4 has a gradient above rho and needs the angle and its double-precision sin/cos.  The CTA stages its
  // 128x8 tile (+1 row, +1 column) in shared memory with word loads, writes the records of the undefined pixels and queues
  // the defined ones (one shared-memory atomic per warp and 32 pixels); the queue is then processed by full warps, so the
  // double-precision pipe is not spent on mostly idle lanes.  (Round-1 form: one pixel per thread, 196 k CTAs per batch with
  // two barriers each and four byte loads per pixel - 0.90 ms per 256 frames, bound by barrier and shared-atomic latency.)
  __shared__ uint32_t tile[GRAD_TH + 1][GRAD_TW / 4 + 1];
  __shared__ uint2 q[GRAD_TW * GRAD_TH];  // (pixel index in frame, gx & 0xffff | gy << 16)
  __shared__ int qn, cmax;
  const int f = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * GRAD_TW, y0 = blockIdx.y * GRAD_TH;
  if (threadIdx.x == 0) { qn = 0; cmax = 0; }
  {
    const uint8_t* S = scaled + (size_t)f * L.spitch * L.sh;
    for (int i = threadIdx.x; i < (GRAD_TH + 1) * (GRAD_TW / 4 + 1); i += 256) {
      const int r = i / (GRAD_TW / 4 + 1), w = i - r * (GRAD_TW / 4 + 1);
      const int yy = y0 + r, xb = x0 + 4 * w;
      tile[r][w] = (yy < L.sh && xb < L.spitch) ? __ldg(reinterpret_cast<const uint32_t*>(S + (size_t)yy * L.spitch + xb)) : 0u;
    }
  }
  __syncthreads();
  uint4* P = pix + (size_t)f * L.P;
  float* DP = degPlane + (size_t)f * L.P;  // the angles alone, 4 B per pixel: what the rectangle scans of k_lsd_nfa read
  unsigned* G2 = g2Plane + (size_t)f * L.P;  // gx^2 + gy^2 of the defined pixels, 0 elsewhere: what the seed sort reads
  const uint8_t* t0 = reinterpret_cast<const uint8_t*>(tile[warp]);
  const uint8_t* t1 = reinterpret_cast<const uint8_t*>(tile[warp + 1]);
  const int y = y0 + warp;
  int m = 0;
#pragma unroll
  for (int j = 0; j < GRAD_TW / 32; ++j) {
    const int xl = j * 32 + lane, x = x0 + xl;
    int g2 = 0, gx = 0, gy = 0;
    bool defined = false;
    const bool inside = x < L.sw && y < L.sh;
    if (inside && x < L.sw - 1 && y < L.sh - 1) {
      const int DA = (int)t1[xl + 1] - (int)t0[xl], BC = (int)t0[xl + 1] - (int)t1[xl];
      gx = DA + BC;
      gy = DA - BC;
      g2 = gx * gx + gy * gy;
      defined = g2 >= L.g2_min;  // <=> sqrt(g2 / 4) > rho, threshold found on the host with the same double operations
    }
    const int idx = y * L.sw + x;
    if (inside) G2[idx] = defined ? (unsigned)g2 : 0u;
    const unsigned dm = __ballot_sync(0xffffffffu, defined);
    int base = 0;
    if (lane == 0 && dm) base = atomicAdd(&qn, __popc(dm));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (defined) {
      q[base + __popc(dm & ((1u << lane) - 1u))] = make_uint2((unsigned)idx, ((unsigned)gx & 0xffffu) | ((unsigned)gy << 16));
      m = max(m, g2);
    } else if (inside) {
      P[idx] = make_uint4(__float_as_uint(NOTDEF_F), 0u, 0u, (unsigned)g2);
      DP[idx] = NOTDEF_F;
    }
    if (bmAll) {
      // bitmap of the pixels that can never join a region (no defined angle): bit (idx & 31) of word (idx >> 5), zeroed by the
      // caller.  A warp holds 32 consecutive pixels of one row = 32 consecutive bits, which straddle two words unless the row
      // starts word-aligned.
      const unsigned und = __ballot_sync(0xffffffffu, inside && !defined);
      if (lane == 0 && y < L.sh) {
        const int bmWords = (L.P + 31) / 32;
        unsigned* bm = bmAll + (size_t)f * ((bmWords + 3) / 4 * 4);
        const int i0 = y * L.sw + x0 + j * 32, sft = i0 & 31;
        if ((L.sw & 31) == 0) {
          if (x0 + j * 32 < L.sw) bm[i0 >> 5] = und;  // rows are whole words: this warp is the only writer (no memset needed)
        } else if (und) {
          atomicOr(bm + (i0 >> 5), und << sft);
          if (sft && (und >> (32 - sft))) atomicOr(bm + (i0 >> 5) + 1, und >> (32 - sft));
        }
      }
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if (lane == 0 && m > 0) atomicMax(&cmax, m);
  __syncthreads();
  if (threadIdx.x == 0 && cmax > 0) atomicMax(maxg2 + f, cmax);
  const int n = qn;
  for (int i = threadIdx.x; i < n; i += 256) {
    const uint2 e = q[i];
    const int gx = (int)(short)(e.y & 0xffffu), gy = (int)e.y >> 16;
    const float deg = fast_atan2_dev((float)gx, (float)(-gy));
    const float af = (float)__dmul_rn((double)deg, PL_DEG_TO_RADS);
    float sn, cs;
    pl_sincosf_dev(af, &sn, &cs);
    P[e.x] = make_uint4(__float_as_uint(deg), __float_as_uint(cs), __float_as_uint(sn), (unsigned)(gx * gx + gy * gy));
    DP[e.x] = deg;
  }
}

// ------------------------------------------------------------------------------------------
// Seed ordering: cv2 4.13 stable-sorts all pixels by bin = int(modgrad * 1023 / max_grad) descending
// (stable => raster order inside a bin).  Only pixels with a defined angle can seed a region, so
// only those are sorted.  Counting sort: per-row bin histograms, a column scan over rows, a scan
// over bins (descending), and a per-row stable scatter.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int lsd_bin(int g2, int mg2) {
  const double max_grad = modgrad_of(mg2);
  const double bin_coef = __ddiv_rn((double)(LSD_BINS - 1), max_grad);
  return (int)__dmul_rn(modgrad_of(g2), bin_coef);
}

// The three kernels read the 4-byte g2 plane of k_lsd_grad (0 = no defined angle), not the 16-byte records, and the
// histogram table has one row per GROUP of 8 image rows (the CTA of k_lsd_scatter rebuilds the split of a group's counts over
// its 8 rows in shared memory): 0.85 GB of traffic per 256 frames of 640x480 instead of 3.0 GB with per-row tables.
constexpr int SORT_ROWS = 8;  // image rows per CTA = per histogram row
__global__ void __launch_bounds__(256) k_lsd_rowhist(const __grid_constant__ LineParams L, const unsigned* __restrict__ g2Plane,
                                                     const int* __restrict__ maxg2, unsigned* __restrict__ grouphist) {
  __shared__ unsigned hist[LSD_BINS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.y, y = blockIdx.x * SORT_ROWS + warp;
  for (int i = threadIdx.x; i < LSD_BINS; i += 256) hist[i] = 0;
  __syncthreads();
  const int mg2 = maxg2[f];
  if (mg2 > 0 && y < L.sh) {
    const unsigned* row = g2Plane + (size_t)f * L.P + (size_t)y * L.sw;
    const double bin_coef = __ddiv_rn((double)(LSD_BINS - 1), modgrad_of(mg2));  // lsd_bin(), its frame constant hoisted
    for (int xb = 0; xb < L.sw; xb += 8 * 32) {
      unsigned g2v[8];  // eight loads in flight per lane
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int x = xb + k * 32 + lane;
        g2v[k] = x < L.sw ? __ldg(row + x) : 0u;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (g2v[k]) atomicAdd(&hist[(int)__dmul_rn(modgrad_of((int)g2v[k]), bin_coef)], 1u);
    }
  }
  __syncthreads();
  unsigned* out = grouphist + ((size_t)f * gridDim.x + blockIdx.x) * LSD_BINS;
  for (int i = threadIdx.x; i < LSD_BINS; i += 256) out[i] = hist[i];
}

__global__ void __launch_bounds__(LSD_BINS) k_lsd_colscan(const __grid_constant__ LineParams L, int ngroups,
                                                          unsigned* __restrict__ grouphist, unsigned* __restrict__ binstart,
                                                          int* __restrict__ nseeds) {
  __shared__ int tot[LSD_BINS];
  __shared__ int warpTmp[33];
  const int f = blockIdx.x, b = threadIdx.x;
  unsigned* col = grouphist + (size_t)f * ngroups * LSD_BINS + b;
  unsigned run = 0;
  for (int g = 0; g < ngroups; ++g) {
    const unsigned v = col[(size_t)g * LSD_BINS];
    col[(size_t)g * LSD_BINS] = run;
    run += v;
  }
  tot[LSD_BINS - 1 - b] = (int)run;  // descending bin order
  __syncthreads();
  const int total = block_scan_excl(tot, LSD_BINS, warpTmp);
  binstart[(size_t)f * LSD_BINS + b] = (unsigned)tot[LSD_BINS - 1 - b];
  if (b == 0) nseeds[f] = total;
}

__global__ void __launch_bounds__(256) k_lsd_scatter(const __grid_constant__ LineParams L, const unsigned* __restrict__ g2Plane,
                                                     const int* __restrict__ maxg2, const unsigned* __restrict__ grouphist,
                                                     const unsigned* __restrict__ binstart, unsigned* __restrict__ seeds) {
  __shared__ unsigned cnt[SORT_ROWS][LSD_BINS];
  extern __shared__ unsigned sortList[];  // [SORT_ROWS][sw]: the seeds of each row in x order, x | bin << 16
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int f = blockIdx.y, y = blockIdx.x * SORT_ROWS + warp;
  const int mg2 = maxg2[f];
  if (mg2 <= 0) return;
  for (int i = lane; i < LSD_BINS; i += 32) cnt[warp][i] = 0;
  __syncwarp();
  const unsigned* row = g2Plane + (size_t)f * L.P + (size_t)y * L.sw;
  unsigned* myList = sortList + (size_t)warp * L.sw;
  // Pass 1: the row's seeds compacted in x order (~1 pixel in 4) with their bins (a double-precision square root each), and
  // the row's count per bin; then, over the 8 warps, the offset of the row inside its group's share of each bin.
  int nd = 0;
  if (y < L.sh) {
    const double bin_coef = __ddiv_rn((double)(LSD_BINS - 1), modgrad_of(mg2));  // lsd_bin(), its frame constant hoisted
    for (int xb = 0; xb < L.sw; xb += 8 * 32) {
      unsigned g2v[8];  // eight loads in flight per lane
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int x = xb + k * 32 + lane;
        g2v[k] = x < L.sw ? __ldg(row + x) : 0u;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int x = xb + k * 32 + lane;
        const unsigned g2 = g2v[k];
        const unsigned dm = __ballot_sync(0xffffffffu, g2 != 0u);
        if (g2) {
          const int bin = (int)__dmul_rn(modgrad_of((int)g2), bin_coef);
          atomicAdd(&cnt[warp][bin], 1u);
          myList[nd + __popc(dm & lt)] = (unsigned)x | ((unsigned)bin << 16);
        }
        nd += __popc(dm);
      }
    }
  }
  __syncthreads();
  {
    // cnt[w][b] becomes the position in the frame's seed list of the first seed of row w in bin b: start of the bin +
    // the groups above + the rows of this group above
    const unsigned* goff = grouphist + ((size_t)f * gridDim.x + blockIdx.x) * LSD_BINS;
    const unsigned* bs = binstart + (size_t)f * LSD_BINS;
    for (int b = threadIdx.x; b < LSD_BINS; b += 256) {
      unsigned run = bs[b] + goff[b];
#pragma unroll
      for (int w = 0; w < SORT_ROWS; ++w) {
        const unsigned v = cnt[w][b];
        cnt[w][b] = run;
        run += v;
      }
    }
  }
  __syncthreads();
  if (y >= L.sh) return;
  unsigned* out = seeds + (size_t)f * L.P;
  // Pass 2: stable scatter, 32 seeds per step
  for (int i0 = 0; i0 < nd; i0 += 32) {
    const int i = i0 + lane;
    const bool def = i < nd;
    int bin = -1 - lane, x = 0;  // unique non-matching key for idle lanes
    if (def) {
      const unsigned e = myList[i];
      bin = (int)(e >> 16);
      x = (int)(e & 0xffffu);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (def) {
      const int rank = __popc(peers & lt);
      const unsigned base = cnt[warp][bin];
      out[base + rank] = (unsigned)(y * L.sw + x);
    }
    __syncwarp();
    if (def && (peers >> lane) == 1u) cnt[warp][bin] += __popc(peers);  // highest lane of each peer group
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// k_lsd_grow: the sequential heart of LSD, one warp per frame.
// ------------------------------------------------------------------------------------------
constexpr int REG_SMEM = 1024;  // region-list entries kept in shared memory (larger regions spill to global)
constexpr int SW_G = 32;        // k_lsd_grow_sw: seeds per group (one warp-wide load of the sorted seed list)
// where region growing keeps the `used` map: in the pixel records (first version, kept for comparison), in a per-frame bitmap
// in shared or global memory (one warp per frame), or in the owner plane of the CTA-per-frame mode
enum { GM_REC = 0, GM_BMS = 1, GM_BMG = 2, GM_MW = 3 };
constexpr unsigned USED_BIT = 0x80000000u;  // `used` flag of a pixel: bit 31 of its record's .w (gx^2+gy^2 < 2^20)

// The `used` map lives in the pixel records themselves (global memory, L1-resident around the growing region),
// so a frame's CTA needs ~5 KB of shared memory and 32 frames fit on one SM.  Only this warp touches the frame's
// records while the kernel runs; __syncwarp() orders its stores and loads.
#ifdef PLSLAM_GROW_PROF
// debug build only (make PROF=1): cycle counters of k_lsd_grow, lane 0 of every frame adds its totals
__device__ unsigned long long g_grow_prof[16];
#define GP_DECL long long gp_t0 = 0
#define GP_START() gp_t0 = clock64()
#define GP_ADD(slot) C.prof[slot] += clock64() - gp_t0
#define GP_CNT(slot, v) C.prof[slot] += (v)
#else
#define GP_DECL
#define GP_START()
#define GP_ADD(slot)
#define GP_CNT(slot, v)
#endif
struct GrowCtx {
#ifdef PLSLAM_GROW_PROF
  long long* prof;
#endif
  uint4* pix;         // per-pixel records of this frame (read-write: used flags)
  unsigned* regS;     // region list ((y << 16) | x): first REG_SMEM entries in shared memory ...
  unsigned* regG;     // ... the rest in this frame's global scratch
  double* stage;      // 3 x 32 doubles of shared staging
  int sw, sh, P;
  int lane;
  bool prefetch;
  bool prefetch2;     // bitmap modes: prefetch only the record sectors of neighbours that are still available
  // ---- bitmap modes (GM_BMS / GM_BMG): one bit per pixel, bit (idx & 31) of word (idx >> 5), set = the pixel cannot join a
  // region any more (no defined angle, or used).  k_lsd_grad writes the undefined bits; the growing warp sets and clears the
  // rest.  GM_BMS keeps the frame's bitmap in shared memory (24 KB at 640x480), GM_BMG in global memory (read through L2).
  unsigned* bm;
  // ---- CTA-per-frame mode (k_lsd_grow_sw): ordered speculative regions, see the kernel's header ----
  // Owner plane: one word per pixel.  MW_FREE, or (tag << 1) | released with tag = seed rank + 1 of the region that
  // marked the pixel.  released = 1 is a TOMBSTONE: the region gave the pixel back (refine() un-marks, reduce_region_radius()
  // drops far points) but its result still depends on having held it, so an older region that takes the pixel while the
  // releasing region is in flight must still poison it.  Tombstones of retired regions read as free.
  // Window slot of a tag = (tag - 1) % nslots (one slot per seed of the sorted list).
  unsigned* own;
  unsigned myVal;         // this region's mark: tag << 1
  const volatile unsigned* headTag;  // shared memory: smallest tag that is not final yet (only ever grows)
  int* poison;            // [nslots] poison flags of the window's seeds (shared memory); mine is poison[slot]
  volatile int* blocker;  // [nslots] tag of the older region that poisoned the seed's region
  int* plist;             // [nslots] ring of poisoned slots (slot + 1) for the retirement warp, 0 = not written yet
  unsigned* plTail;
  unsigned long long* pstat;  // 3 counters: poisoned by an older holder / an older tombstone / robbed a younger region
  int slot, nslots;
  int cap;                // entries of this worker's list area (list + scratch of refine())
  __device__ __forceinline__ bool mw_poisoned() const { return *reinterpret_cast<volatile int*>(poison + slot) != 0; }
  __device__ __forceinline__ void mw_mark_poison(int j, unsigned tag) const {
    blocker[j] = (int)tag;
    if (atomicExch(poison + j, 1) == 0) {  // first poison of this run: tell the retirement warp
      const unsigned e = atomicAdd(plTail, 1u);
      *reinterpret_cast<volatile int*>(plist + e % (unsigned)nslots) = j + 1;
    }
  }
  // May this region still test the pixel?  Not if a retired region or this region itself holds it.  A pixel held (or
  // tombstoned) by another in-flight region stays a candidate: whether it is used only matters if it passes the alignment
  // test, and then mw_take() settles it (a younger region is robbed and poisoned, an older one poisons me).
  // `h0` = the head tag read BEFORE the owner word was loaded: a holder that was still running when the word was read, and
  // may un-mark the pixel later, must not pass for retired because it retired meanwhile.
  __device__ __forceinline__ bool mw_available(unsigned ov, unsigned h0) const {
    if (ov == 0xffffffffu) return true;
    if ((ov | 1u) == (myVal | 1u)) return (ov & 1u) != 0;  // mine: only what I released myself
    if (ov & 1u) return true;                                // tombstone: free if its region retired, else mw_take decides
    return (ov >> 1) >= h0;
  }
  // Take pixel `id`, last seen holding `ov`.
  __device__ __forceinline__ void mw_take(int id, unsigned ov) const {
    unsigned exp = ov;
    for (int it = 0; it < 64; ++it) {
      if (exp != 0xffffffffu && exp < myVal) {  // a lower word is in place
        const unsigned t = exp >> 1;
        if ((exp & 1u) && t < *headTag) {  // tombstone of a retired region: free, but atomicMin cannot replace it
          const unsigned old = atomicCAS(own + id, exp, myVal);
          if (old == exp) return;
          exp = old;
          continue;
        }
        atomicAdd(pstat + (exp & 1u), 1ull);
        mw_mark_poison(slot, t);  // an older region holds the pixel (or gave it back and is still in flight)
        return;
      }
      const unsigned old = atomicMin(own + id, myVal);
      if (old == 0xffffffffu || (old | 1u) == (myVal | 1u)) return;
      if (old > myVal) {  // robbed a younger in-flight region (or its tombstone): it must not retire with this result
        atomicAdd(pstat + 2, 1ull);
        mw_mark_poison((int)(((old >> 1) - 1u) % (unsigned)nslots), myVal >> 1);
        return;
      }
      exp = old;  // an older word got there first
    }
    mw_mark_poison(slot, 0);  // never seen: give up on this run, the region is grown again
  }
  // refine() / reduce_region_radius(): the pixel is free again, the tombstone keeps the dependence visible
  __device__ __forceinline__ void mw_release(int id) const { atomicCAS(own + id, myVal, myVal | 1u); }
  __device__ __forceinline__ unsigned reg_get(int i) const { return i < REG_SMEM ? regS[i] : regG[i]; }
  __device__ __forceinline__ void reg_set(int i, unsigned v) const {
    if (i < REG_SMEM) regS[i] = v; else regG[i] = v;
  }
  __device__ __forceinline__ unsigned* wptr(int idx) const { return reinterpret_cast<unsigned*>(pix + idx) + 3; }
  template <int M>
  __device__ __forceinline__ unsigned bm_word(int w) const {
    if (M == GM_BMS) {
      unsigned v;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(bm + w)) : "memory");
      return v;
    }
    return __ldcg(bm + w);
  }
  template <int M>
  __device__ __forceinline__ bool bm_test(int idx) const { return (bm_word<M>(idx >> 5) >> (idx & 31)) & 1u; }
  template <int M>
  __device__ __forceinline__ void bm_set(int idx) const {
    if (M == GM_BMS)
      asm volatile("red.shared.or.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bm + (idx >> 5))), "r"(1u << (idx & 31)) : "memory");
    else
      atomicOr(bm + (idx >> 5), 1u << (idx & 31));
  }
  template <int M>
  __device__ __forceinline__ void bm_clear(int idx) const {
    if (M == GM_BMS)
      asm volatile("red.shared.and.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bm + (idx >> 5))), "r"(~(1u << (idx & 31))) : "memory");
    else
      atomicAnd(bm + (idx >> 5), ~(1u << (idx & 31)));
  }
  // Bitmap modes: pull the record sectors (32 B = two 16-byte records) of the still available neighbours of a new region point
  // towards L1; they are loaded when the point reaches the scan front, a batch or more from now.
  template <int M>
  __device__ __forceinline__ void bm_prefetch_around(int idx, unsigned xy) const {
    const int y = (int)(xy >> 16);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      if (y + dy < 0 || y + dy >= sh) continue;
      int id0 = idx + dy * sw - 1;  // records id0, id0 + 1, id0 + 2 (at the image border one of them belongs to another row: harmless)
      id0 = max(0, min(id0, P - 3));
      const unsigned w0 = bm_word<M>(id0 >> 5), w1 = bm_word<M>((id0 + 2) >> 5);
      const unsigned av = ~__funnelshift_r(w0, w1, id0 & 31) & 7u;
      const unsigned inA = (id0 & 1) ? 1u : 3u;  // which of the three records share the first sector
      const uint4* q = pix + id0;
      if (av & inA) asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
      if (av & (7u ^ inA)) asm volatile("prefetch.global.L1 [%0];" ::"l"(q + 2));
    }
  }
};

__device__ __forceinline__ bool lsd_aligned(double theta, float deg, double prec) {
  const double a = __dmul_rn((double)deg, PL_DEG_TO_RADS);
  double n_theta = __dsub_rn(theta, a);
  if (n_theta < 0) n_theta = -n_theta;
  if (n_theta > M_3_2_PI_D) {
    n_theta = __dsub_rn(n_theta, M_2__PI_D);
    if (n_theta < 0) n_theta = -n_theta;
  }
  return n_theta <= prec;
}

__device__ __forceinline__ bool lsd_aligned_rad(double theta, double a, double prec) {
  double n_theta = __dsub_rn(theta, a);
  if (n_theta < 0) n_theta = -n_theta;
  if (n_theta > M_3_2_PI_D) {
    n_theta = __dsub_rn(n_theta, M_2__PI_D);
    if (n_theta < 0) n_theta = -n_theta;
  }
  return n_theta <= prec;
}

// region_grow(): returns the region size; reg[0] must already hold the seed pixel.
template <int M>
__device__ int lsd_region_grow(const GrowCtx& C, double prec, double* reg_angle_out) {
  constexpr bool BM = M == GM_BMS || M == GM_BMG;
  const int lane = C.lane;
  const unsigned seedxy = C.reg_get(0);
  const int seed = (int)(seedxy >> 16) * C.sw + (int)(seedxy & 0xffff);
  const uint4 srec = C.pix[seed];
  double reg_angle = __dmul_rn((double)__uint_as_float(srec.x), PL_DEG_TO_RADS);
  float sumdx, sumdy;
  {
    double s, c;
    pl_sincos_dev(reg_angle, &s, &c);  // the seed uses the double-precision angle (sincos() in lsd.cpp)
    sumdx = (float)c;
    sumdy = (float)s;
  }
  if (lane == 0) {
    if (BM) C.bm_set<M>(seed);
    else *C.wptr(seed) = srec.w | USED_BIT;
  }
  __syncwarp();
  int n = 1;
  for (int i = 0; i < n;) {
    const int m = min(4, n - i);
    // 32 lanes: lane = 8 * point + neighbour; neighbours in the 3x3 row-major order of lsd.cpp (yy outer, xx
    // inner) without the centre, which is always already used
    const int pt = lane >> 3, nb8 = lane & 7, nb = nb8 < 4 ? nb8 : nb8 + 1;
    int nidx = -1;
    unsigned nxy = 0, w = 0;
    float deg = NOTDEF_F, cs = 0.f, sn = 0.f;
    if (pt < m) {
      const unsigned p = C.reg_get(i + pt);
      const int nx = (int)(p & 0xffff) + (nb % 3) - 1, ny = (int)(p >> 16) + (nb / 3) - 1;
      if (nx >= 0 && ny >= 0 && nx < C.sw && ny < C.sh) {
        nidx = ny * C.sw + nx;
        nxy = ((unsigned)ny << 16) | (unsigned)nx;
        if (!BM || !C.bm_test<M>(nidx)) {
          const uint4 r = C.pix[nidx];
          w = r.w;
          if (BM || !(w & USED_BIT)) {
            deg = __uint_as_float(r.x);
            cs = __uint_as_float(r.y);
            sn = __uint_as_float(r.z);
          }
        }
      }
    }
    unsigned todo = 0xffffffffu;  // lanes not yet passed by the sequential scan
    while (true) {
      const bool cand = deg != NOTDEF_F && lsd_aligned(reg_angle, deg, prec);
      const unsigned mask = __ballot_sync(0xffffffffu, cand) & todo;
      if (!mask) break;
      const int j = __ffs(mask) - 1;
      const int aidx = __shfl_sync(0xffffffffu, nidx, j);
      if (lane == j) {
        if (BM) C.bm_set<M>(nidx);
        else *C.wptr(nidx) = w | USED_BIT;
        C.reg_set(n, nxy);
      }
      ++n;
      sumdx = __fadd_rn(sumdx, __shfl_sync(0xffffffffu, cs, j));
      sumdy = __fadd_rn(sumdy, __shfl_sync(0xffffffffu, sn, j));
      reg_angle = __dmul_rn((double)fast_atan2_dev(sumdy, sumdx), PL_DEG_TO_RADS);
      if (nidx == aidx) deg = NOTDEF_F;  // the same pixel seen from another frontier point is now used
      todo = j == 31 ? 0u : (0xffffffffu << (j + 1));
    }
    i += m;
    __syncwarp();
  }
  *reg_angle_out = reg_angle;
  return n;
}

// region_grow() with in-batch speculation.  The reference examines the 8 neighbours of each region point in
// order and updates the region angle after every accepted pixel, so later tests see the new angle.  Here the
// 32 neighbour slots of 4 frontier points are first tested against the angle at batch entry (set m); every lane
// then rebuilds, with the same float additions in the same order, the sums it would have seen had all earlier
// lanes of m been accepted, re-tests itself against that exact pre-state and the warp commits the longest prefix
// on which the two tests agree (plus the corrected decision of the first disagreeing lane).  A batch whose
// decisions do not depend on the drift of the angle costs one round instead of one round per accepted pixel;
// the result is identical to the sequential scan by construction.
template <int M>
__device__ int lsd_region_grow_spec(const GrowCtx& C, double prec, double* reg_angle_out) {
  constexpr bool MW = M == GM_MW, BM = M == GM_BMS || M == GM_BMG;
  const int lane = C.lane;
  const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;
  const unsigned seedxy = C.reg_get(0);
  const int seed = (int)(seedxy >> 16) * C.sw + (int)(seedxy & 0xffff);
  const uint4 srec = C.pix[seed];
  double reg_angle = __dmul_rn((double)__uint_as_float(srec.x), PL_DEG_TO_RADS);
  float sumdx, sumdy;
  {
    double s, c;
    pl_sincos_dev(reg_angle, &s, &c);
    sumdx = (float)c;
    sumdy = (float)s;
  }
  if (lane == 0) {
    if (MW) C.mw_take(seed, __ldcg(C.own + seed));
    else if (BM) C.bm_set<M>(seed);
    else *C.wptr(seed) = srec.w | USED_BIT;
  }
  __syncwarp();
  int n = 1;
  GP_DECL;
  GP_CNT(8, 1);
  const int pt = lane >> 3, nb8 = lane & 7, nb = nb8 < 4 ? nb8 : nb8 + 1;
  const int ox = (nb % 3) - 1, oy = (nb / 3) - 1;
  for (int i = 0; i < n;) {
    if (MW && C.mw_poisoned()) break;  // a poisoned region is discarded anyway: stop early
    const int m4 = min(4, n - i);
    GP_START();
    GP_CNT(9, 1);
    int nidx = -1;
    unsigned nxy = 0, w = 0, ovk = 0;
    float deg = NOTDEF_F, cs = 0.f, sn = 0.f;
    unsigned h0 = 0;
    if (MW) h0 = *C.headTag;
    if (pt < m4) {
      const unsigned p = i + 4 <= REG_SMEM ? C.regS[i + pt] : C.reg_get(i + pt);  // warp-uniform fast path
      const int nx = (int)(p & 0xffff) + ox, ny = (int)(p >> 16) + oy;
      if (nx >= 0 && ny >= 0 && nx < C.sw && ny < C.sh) {
        const int id = ny * C.sw + nx;
        if (BM) {
          if (!C.bm_test<M>(id)) {  // defined and not used: only these records are loaded at all
            const uint4 r = C.pix[id];
            nidx = id;
            nxy = ((unsigned)ny << 16) | (unsigned)nx;
            deg = __uint_as_float(r.x);
            cs = __uint_as_float(r.y);
            sn = __uint_as_float(r.z);
          }
        } else {
          unsigned ov = 0;
          if (MW) ov = __ldcg(C.own + id);  // owner words live in L2 (atomics), never in a possibly stale L1 line
          const uint4 r = C.pix[id];
          const bool avail = MW ? (__uint_as_float(r.x) != NOTDEF_F && C.mw_available(ov, h0)) : !(r.w & USED_BIT);
          if (avail && __uint_as_float(r.x) != NOTDEF_F) {
            nidx = id;
            nxy = ((unsigned)ny << 16) | (unsigned)nx;
            w = r.w;
            ovk = ov;
            deg = __uint_as_float(r.x);
            cs = __uint_as_float(r.y);
            sn = __uint_as_float(r.z);
          }
        }
      }
    }
    // the candidate's level-line angle in radians, once per batch (both alignment tests of every round use it)
    const double arad = __dmul_rn((double)deg, PL_DEG_TO_RADS);
    const bool anyc = __any_sync(FULL, deg != NOTDEF_F);
    GP_ADD(0);
    GP_CNT(6, m4);
    GP_CNT(7, anyc ? 1 : 0);
    GP_START();
    if (anyc) {
      // lanes looking at the same pixel (MATCH.ANY costs a step per distinct value: only the candidate lanes take part)
      const unsigned candm = __ballot_sync(FULL, deg != NOTDEF_F);
      const unsigned peers = deg != NOTDEF_F ? __match_any_sync(candm, nidx) : (1u << lane);
      unsigned todo = FULL;                                 // lanes the sequential scan has not passed yet
      while (true) {
        const bool in = (todo >> lane) & 1u;
        const bool a0 = in && deg != NOTDEF_F && lsd_aligned_rad(reg_angle, arad, prec) && !(peers & todo & lt);
        const unsigned m = __ballot_sync(FULL, a0);
        if (!m) break;
        GP_CNT(10, 1);
        {
          // common case of thin regions: the accepted lane is the last lane that holds a candidate pixel at all, so no
          // later decision can depend on the new angle and the earlier lanes were tested against the exact state
          const unsigned defm = __ballot_sync(FULL, in && deg != NOTDEF_F);
          const int j0 = __ffs(m) - 1;
          if ((defm >> j0) == 1u) {
            if (lane == j0) {
              if (MW) C.mw_take(nidx, ovk);
              else if (BM) C.bm_set<M>(nidx);
              else *C.wptr(nidx) = w | USED_BIT;
              C.reg_set(n, nxy);
              if (BM && C.prefetch2) {
                C.bm_prefetch_around<M>(nidx, nxy);
              } else if (C.prefetch) {
                const int up = (nxy >> 16) > 0 ? nidx - C.sw : nidx, dn = (int)(nxy >> 16) < C.sh - 1 ? nidx + C.sw : nidx;
                const uint4* q = C.pix + nidx;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(q + (up - nidx)));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(q + (dn - nidx)));
              }
            }
            ++n;
            sumdx = __fadd_rn(sumdx, __shfl_sync(FULL, cs, j0));
            sumdy = __fadd_rn(sumdy, __shfl_sync(FULL, sn, j0));
            reg_angle = __dmul_rn((double)fast_atan2_dev(sumdy, sumdx), PL_DEG_TO_RADS);
            break;
          }
        }
        // sums before this lane, assuming every earlier lane of m is accepted (additions in scan order)
        float sx = sumdx, sy = sumdy;
        for (unsigned r = m; r; r &= r - 1u) {
          const int j = __ffs(r) - 1;
          const float cj = __shfl_sync(FULL, cs, j), sj = __shfl_sync(FULL, sn, j);
          if (lane > j) {
            sx = __fadd_rn(sx, cj);
            sy = __fadd_rn(sy, sj);
          }
        }
        const float px = a0 ? __fadd_rn(sx, cs) : sx, py = a0 ? __fadd_rn(sy, sn) : sy;  // sums after this lane
        const double ra_post = __dmul_rn((double)fast_atan2_dev(py, px), PL_DEG_TO_RADS);
        const double ra_up = __shfl_up_sync(FULL, ra_post, 1);
        const double ra = (m & lt) ? ra_up : reg_angle;  // exact region angle this lane is tested against
        const bool a1 = in && deg != NOTDEF_F && lsd_aligned_rad(ra, arad, prec) && !(peers & m & lt);
        const unsigned m1 = __ballot_sync(FULL, a1);
        const unsigned bad = (m ^ m1) & todo;
        unsigned commit;
        if (!bad) {
          commit = m;
        } else {
          const int b = __ffs(bad) - 1;
          commit = (m & ((1u << b) - 1u)) | (m1 & (1u << b));
        }
        if ((commit >> lane) & 1u) {
          if (MW) C.mw_take(nidx, ovk);
          else if (BM) C.bm_set<M>(nidx);
          else *C.wptr(nidx) = w | USED_BIT;
          C.reg_set(n + __popc(commit & lt), nxy);
          if (BM && C.prefetch2) {
            C.bm_prefetch_around<M>(nidx, nxy);
          } else if (C.prefetch) {
            // the 3x3 neighbourhood of the new region point is examined when the point reaches the scan front, a
            // few batches from now: pull the lines holding its three record rows towards L1
            const int up = (nxy >> 16) > 0 ? nidx - C.sw : nidx, dn = (int)(nxy >> 16) < C.sh - 1 ? nidx + C.sw : nidx;
            const uint4* q = C.pix + nidx;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(q + (up - nidx)));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(q + (dn - nidx)));
          }
        }
        n += __popc(commit);
        if (!bad) {
          sumdx = __shfl_sync(FULL, px, 31);
          sumdy = __shfl_sync(FULL, py, 31);
          reg_angle = __shfl_sync(FULL, ra_post, 31);
          break;
        }
        GP_CNT(11, 1);
        const int b = __ffs(bad) - 1;
        const float bx = __shfl_sync(FULL, sx, b), by = __shfl_sync(FULL, sy, b);
        if ((m1 >> b) & 1u) {  // lane b is accepted after all: its own contribution enters the sums
          sumdx = __fadd_rn(bx, __shfl_sync(FULL, cs, b));
          sumdy = __fadd_rn(by, __shfl_sync(FULL, sn, b));
          reg_angle = __dmul_rn((double)fast_atan2_dev(sumdy, sumdx), PL_DEG_TO_RADS);
        } else {  // lane b is rejected under its exact pre-state: the state stays that pre-state
          sumdx = bx;
          sumdy = by;
          reg_angle = __shfl_sync(FULL, ra, b);
        }
        if (peers & commit) deg = NOTDEF_F;  // the pixel is used now, whichever frontier point looks at it
        todo = b == 31 ? 0u : (FULL << (b + 1));
      }
    }
    i += m4;
    __syncwarp();
    GP_ADD(1);
  }
  GP_CNT(12, n);
  *reg_angle_out = reg_angle;
  return n;
}

__device__ __forceinline__ double angle_diff_signed_dev(double a, double b) {
  double diff = __dsub_rn(a, b);
  while (diff <= -PL_PI) diff = __dadd_rn(diff, M_2__PI_D);
  while (diff > PL_PI) diff = __dsub_rn(diff, M_2__PI_D);
  return diff;
}

// region2rect() + get_theta(); sums run in region order (staged through shared memory, accumulated
// redundantly by every lane so the results are warp-uniform).
__device__ void lsd_region2rect(const GrowCtx& C, int n, double reg_angle, double prec, double p, LsdRect* rec) {
  const int lane = C.lane;
  double* sA = C.stage;
  double* sB = C.stage + 32;
  double* sC = C.stage + 64;
  // The three running sums are independent chains that must add in region order: lanes 0..2 own one chain each (lane k
  // reads its own staging row), so a chunk costs one load + one add per point instead of three of each.
  const double* myRow = C.stage + 32 * min(lane, 2);
  double acc = 0;
  for (int c = 0; c < n; c += 32) {
    const int i = c + lane;
    if (i < n) {
      const unsigned pxy = C.reg_get(i);
      const int py = (int)(pxy >> 16), px = (int)(pxy & 0xffff);
      const double w = modgrad_of((int)(*C.wptr(py * C.sw + px) & ~USED_BIT));
      sA[lane] = __dmul_rn((double)px, w);
      sB[lane] = __dmul_rn((double)py, w);
      sC[lane] = w;
    }
    __syncwarp();
    const int cnt = min(32, n - c);
    if (lane < 3)
      for (int j = 0; j < cnt; ++j) acc = __dadd_rn(acc, myRow[j]);
    __syncwarp();
  }
  const double sum = __shfl_sync(0xffffffffu, acc, 2);
  const double x = __ddiv_rn(__shfl_sync(0xffffffffu, acc, 0), sum);
  const double y = __ddiv_rn(__shfl_sync(0xffffffffu, acc, 1), sum);
  acc = 0;
  for (int c = 0; c < n; c += 32) {
    const int i = c + lane;
    if (i < n) {
      const unsigned pxy = C.reg_get(i);
      const int py = (int)(pxy >> 16), px = (int)(pxy & 0xffff);
      const double w = modgrad_of((int)(*C.wptr(py * C.sw + px) & ~USED_BIT));
      const double dx = __dsub_rn((double)px, x), dy = __dsub_rn((double)py, y);
      sA[lane] = __dmul_rn(__dmul_rn(dy, dy), w);
      sB[lane] = __dmul_rn(__dmul_rn(dx, dx), w);
      sC[lane] = -__dmul_rn(__dmul_rn(dx, dy), w);  // Ixy -= v  ==  Ixy += -v exactly
    }
    __syncwarp();
    const int cnt = min(32, n - c);
    if (lane < 3)
      for (int j = 0; j < cnt; ++j) acc = __dadd_rn(acc, myRow[j]);
    __syncwarp();
  }
  const double Ixx = __shfl_sync(0xffffffffu, acc, 0), Iyy = __shfl_sync(0xffffffffu, acc, 1),
               Ixy = __shfl_sync(0xffffffffu, acc, 2);
  const double dI = __dsub_rn(Ixx, Iyy);
  const double disc = __dadd_rn(__dmul_rn(dI, dI), __dmul_rn(__dmul_rn(4.0, Ixy), Ixy));
  const double lambda = __dmul_rn(0.5, __dsub_rn(__dadd_rn(Ixx, Iyy), sqrt(disc)));
  double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2_dev((float)__dsub_rn(lambda, Ixx), (float)Ixy)
                                         : (double)fast_atan2_dev((float)Ixy, (float)__dsub_rn(lambda, Iyy));
  theta = __dmul_rn(theta, PL_DEG_TO_RADS);
  if (fabs(angle_diff_signed_dev(theta, reg_angle)) > prec) theta = __dadd_rn(theta, PL_PI);
  double dx, dy;
  pl_sincos_dev(theta, &dy, &dx);
  double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
  for (int i = lane; i < n; i += 32) {
    const unsigned pxy = C.reg_get(i);
    const int py = (int)(pxy >> 16), px = (int)(pxy & 0xffff);
    const double rdx = __dsub_rn((double)px, x), rdy = __dsub_rn((double)py, y);
    const double l = __dadd_rn(__dmul_rn(rdx, dx), __dmul_rn(rdy, dy));
    const double w = __dadd_rn(__dmul_rn(-rdx, dy), __dmul_rn(rdy, dx));
    l_max = fmax(l_max, l);
    l_min = fmin(l_min, l);
    w_max = fmax(w_max, w);
    w_min = fmin(w_min, w);
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    l_max = fmax(l_max, __shfl_xor_sync(0xffffffffu, l_max, d));
    l_min = fmin(l_min, __shfl_xor_sync(0xffffffffu, l_min, d));
    w_max = fmax(w_max, __shfl_xor_sync(0xffffffffu, w_max, d));
    w_min = fmin(w_min, __shfl_xor_sync(0xffffffffu, w_min, d));
  }
  rec->x1 = __dadd_rn(x, __dmul_rn(l_min, dx));
  rec->y1 = __dadd_rn(y, __dmul_rn(l_min, dy));
  rec->x2 = __dadd_rn(x, __dmul_rn(l_max, dx));
  rec->y2 = __dadd_rn(y, __dmul_rn(l_max, dy));
  rec->width = __dsub_rn(w_max, w_min);
  rec->x = x;
  rec->y = y;
  rec->theta = theta;
  rec->dx = dx;
  rec->dy = dy;
  rec->prec = prec;
  rec->p = p;
  if (rec->width < 1.0) rec->width = 1.0;
}

__device__ __forceinline__ double dist_sq_dev(double x1, double y1, double x2, double y2) {
  const double dx = __dsub_rn(x2, x1), dy = __dsub_rn(y2, y1);
  return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}
__device__ __forceinline__ double rect_density(int n, const LsdRect& r) {
  return __ddiv_rn((double)n, __dmul_rn(sqrt(dist_sq_dev(r.x1, r.y1, r.x2, r.y2)), r.width));
}

// refine() + reduce_region_radius() in their straight-line form (the CTA-per-frame kernel keeps it: its warps share one CTA's
// instruction stream, and this is the form its inter-warp protocol was validated with); returns false when the region must be
// dropped. *n_io = region size.
template <int M>
__device__ bool lsd_refine(const GrowCtx& C, int* n_io, double reg_angle, double prec, double p, LsdRect* rec,
                           double density_th, int variant) {
  constexpr bool MW = M == GM_MW, BM = M == GM_BMS || M == GM_BMG;
  const int lane = C.lane;
  int n = *n_io;
  double density = rect_density(n, *rec);
  if (density >= density_th) return true;
  const unsigned seedxy = C.reg_get(0);
  const int sy = (int)(seedxy >> 16), sx = (int)(seedxy & 0xffff);
  const double xc = (double)sx, yc = (double)sy;
  const double ang_c = __dmul_rn((double)__uint_as_float(C.pix[sy * C.sw + sx].x), PL_DEG_TO_RADS);
  double* sA = C.stage;
  double* sF = C.stage + 32;
  double* sV2 = C.stage + 64;
  double acc = 0;
  int cntN = 0;
  for (int c = 0; c < n; c += 32) {
    const int i = c + lane;
    bool inside = false;
    if (i < n) {
      const unsigned pxy = C.reg_get(i);
      const int py = (int)(pxy >> 16), px = (int)(pxy & 0xffff);
      const int idx = py * C.sw + px;
      const uint4 r = C.pix[idx];
      if (MW) C.mw_release(idx);
      else if (BM) C.bm_clear<M>(idx);
      else *C.wptr(idx) = r.w & ~USED_BIT;
      double flag = 0.0, v = 0.0;
      if (sqrt(dist_sq_dev(xc, yc, (double)px, (double)py)) < rec->width) {
        const double ang = __dmul_rn((double)__uint_as_float(r.x), PL_DEG_TO_RADS);
        v = angle_diff_signed_dev(ang, ang_c);
        flag = 1.0;
      }
      sA[lane] = v;
      sF[lane] = flag;
      sV2[lane] = __dmul_rn(v, v);
      inside = flag != 0.0;
    }
    cntN += __popc(__ballot_sync(0xffffffffu, inside));
    __syncwarp();
    const int cnt = min(32, n - c);
    // two ordered chains (sum of v, sum of v*v over the flagged points): lane 0 and lane 1 own one each
    if (lane < 2) {
      const double* row = lane == 0 ? sA : sV2;
      for (int j = 0; j < cnt; ++j)
        if (sF[j] != 0.0) acc = __dadd_rn(acc, row[j]);
    }
    __syncwarp();
  }
  const double sum = __shfl_sync(0xffffffffu, acc, 0), s_sum = __shfl_sync(0xffffffffu, acc, 1);
  const double mean_angle = __ddiv_rn(sum, (double)cntN);
  const double tau = __dmul_rn(
      2.0, sqrt(__dadd_rn(__ddiv_rn(__dsub_rn(s_sum, __dmul_rn(__dmul_rn(2.0, mean_angle), sum)), (double)cntN),
                          __dmul_rn(mean_angle, mean_angle))));
  if (BM) __syncwarp();  // the cleared bits are read by the growth that follows
  n = (MW || (variant & 1)) ? lsd_region_grow_spec<M>(C, tau, &reg_angle) : lsd_region_grow<M>(C, tau, &reg_angle);
  if (MW && C.mw_poisoned()) {
    *n_io = n;
    return false;
  }
  *n_io = n;
  if (n < 2) return false;
  lsd_region2rect(C, n, reg_angle, prec, p, rec);
  density = rect_density(n, *rec);
  if (density >= density_th) return true;
  // reduce_region_radius()
  const double radSq1 = dist_sq_dev(xc, yc, rec->x1, rec->y1), radSq2 = dist_sq_dev(xc, yc, rec->x2, rec->y2);
  double radSq = radSq1 > radSq2 ? radSq1 : radSq2;
  while (density < density_th) {
    radSq = __dmul_rn(radSq, 0.75 * 0.75);
    // The reference removes far points by swapping each with the current last element (that order defines the later
    // sums).  Equivalent closed form (checked exhaustively against the sequential loop): with n' points kept, the far
    // positions below n' in increasing order receive the kept points at or above n' in decreasing order.
    if (n + (n >> 1) + 32 > (MW ? C.cap : C.P)) {  // no room for the scratch list behind the region (rare; kept for safety)
      if (lane == 0) {
        for (int i = 0; i < n; ++i) {
          const unsigned pxy = C.reg_get(i);
          const int py = (int)(pxy >> 16), px = (int)(pxy & 0xffff);
          if (dist_sq_dev(xc, yc, (double)px, (double)py) > radSq) {
            if (MW) {
              C.mw_release(py * C.sw + px);
            } else if (BM) {
              C.bm_clear<M>(py * C.sw + px);
            } else {
              unsigned* wp = C.wptr(py * C.sw + px);
              *wp = *wp & ~USED_BIT;
            }
            C.reg_set(i, C.reg_get(n - 1));
            C.reg_set(n - 1, pxy);
            --n;
            --i;
          }
        }
      }
      n = __shfl_sync(0xffffffffu, n, 0);
    } else {
      const unsigned FAR = 0x80000000u, FULL = 0xffffffffu, lt = (1u << lane) - 1u;  // row index < 2^15: bit 31 is free
      int nfar = 0;
      for (int c = 0; c < n; c += 32) {
        const int i = c + lane;
        bool far = false;
        if (i < n) {
          const unsigned pxy = C.reg_get(i);
          const int py = (int)(pxy >> 16), px = (int)(pxy & 0xffff);
          far = dist_sq_dev(xc, yc, (double)px, (double)py) > radSq;
          if (far) {
            if (MW) {
              C.mw_release(py * C.sw + px);
            } else if (BM) {
              C.bm_clear<M>(py * C.sw + px);
            } else {
              unsigned* wp = C.wptr(py * C.sw + px);
              *wp = *wp & ~USED_BIT;
            }
            C.reg_set(i, pxy | FAR);
          }
        }
        nfar += __popc(__ballot_sync(FULL, far));
      }
      __syncwarp();
      const int nk = n - nfar;
      // holes (far positions below nk, ascending) into scratch: the global list beyond n is free
      unsigned* holes = C.regG + n;
      int nh = 0;
      for (int c = 0; c < nk; c += 32) {
        const int i = c + lane;
        const bool hole = i < nk && (C.reg_get(i) & FAR);
        const unsigned b = __ballot_sync(FULL, hole);
        if (hole) holes[nh + __popc(b & lt)] = (unsigned)i;
        nh += __popc(b);
      }
      __syncwarp();
      // fillers (kept positions at or above nk, descending): the k-th goes to the k-th hole
      int nf = 0;
      for (int c = n - 1; c >= nk; c -= 32) {
        const int i = c - lane;
        unsigned v = 0;
        const bool fill = i >= nk && !((v = C.reg_get(i)) & FAR);
        const unsigned b = __ballot_sync(FULL, fill);
        if (fill) C.reg_set((int)holes[nf + __popc(b & lt)], v);
        nf += __popc(b);
      }
      n = nk;
    }
    __syncwarp();
    *n_io = n;
    if (n < 2) return false;
    lsd_region2rect(C, n, reg_angle, prec, p, rec);
    density = rect_density(n, *rec);
  }
  return true;
}

// One seed of the LSD main loop: region_grow() -> region2rect() -> refine() (re-growing with the tolerance tau of the region's
// first part) -> reduce_region_radius() (repeatedly), written as ONE loop around a single region-growing site and a single
// rectangle-fitting site.  The straight-line form (grow; rect; refine{grow; rect; while{reduce; rect}}) inlines the growth twice and
// the rectangle fit three times: 148 KB of code for k_lsd_grow, and with ~28 warps per SM at different places in it the warps
// wait for instructions more than for anything else (ncu at saturation, 4096 frames per launch: stall_no_instruction 8.1 of 17
// cycles per issued instruction, profiles/r02_kernels_ncu.md).  Same operations in the same order, hence the same result.
// reg[0] must hold the seed pixel.  Returns whether the region yields a rectangle; *n_out = size of the region list left behind.
template <int M>
__device__ __forceinline__ bool lsd_seed_region(const GrowCtx& C, double prec, double p, double density_th, int min_reg_size,
                                                int variant, LsdRect* rec, int* n_out) {
  constexpr bool MW = M == GM_MW, BM = M == GM_BMS || M == GM_BMG;
  const int lane = C.lane;
  int phase = 0;          // 0 first growth, 1 re-grown by refine(), 2 inside reduce_region_radius()
  double tol = prec;      // angular tolerance of the next growth
  double reg_angle = 0, xc = 0, yc = 0, radSq = 0;
  int n = 0;
  while (true) {
    if (BM && phase) __syncwarp();  // the bits refine() cleared are read by the growth that follows
    (void)variant;  // (bit 0 once selected the one-accept-per-round scan, lsd_region_grow: dropped from this kernel, it doubled
                    //  the growth code for an experiment switch; the function stays for reference)
    n = lsd_region_grow_spec<M>(C, tol, &reg_angle);
    *n_out = n;
    if (MW && C.mw_poisoned()) return false;
    GP_CNT(15, (phase == 0 && n < min_reg_size) ? 1 : 0);
    GP_CNT(5, (phase == 0 && n == 1) ? 1 : 0);
    if (n < (phase ? 2 : min_reg_size)) return false;
    while (true) {
      lsd_region2rect(C, n, reg_angle, prec, p, rec);
      const double density = rect_density(n, *rec);
      if (density >= density_th) return true;
      if (phase == 0) break;  // -> refine(): un-mark the region, estimate tau, grow again
      if (phase == 1) {       // reduce_region_radius(): start from the farther end of the rectangle
        const double radSq1 = dist_sq_dev(xc, yc, rec->x1, rec->y1), radSq2 = dist_sq_dev(xc, yc, rec->x2, rec->y2);
        radSq = radSq1 > radSq2 ? radSq1 : radSq2;
        phase = 2;
      }
      radSq = __dmul_rn(radSq, 0.75 * 0.75);
      // The reference removes far points by swapping each with the current last element (that order defines the later
      // sums).  Equivalent closed form (checked exhaustively against the sequential loop): with n' points kept, the far
      // positions below n' in increasing order receive the kept points at or above n' in decreasing order.
      if (n + (n >> 1) + 32 > (MW ? C.cap : C.P)) {  // no room for the scratch list behind the region (rare; kept for safety)
        if (lane == 0) {
          for (int i = 0; i < n; ++i) {
            const unsigned pxy = C.reg_get(i);
            const int py = (int)(pxy >> 16), px = (int)(pxy & 0xffff);
            if (dist_sq_dev(xc, yc, (double)px, (double)py) > radSq) {
              if (MW) {
                C.mw_release(py * C.sw + px);
              } else if (BM) {
                C.bm_clear<M>(py * C.sw + px);
              } else {
                unsigned* wp = C.wptr(py * C.sw + px);
                *wp = *wp & ~USED_BIT;
              }
              C.reg_set(i, C.reg_get(n - 1));
              C.reg_set(n - 1, pxy);
              --n;
              --i;
            }
          }
        }
        n = __shfl_sync(0xffffffffu, n, 0);
      } else {
        const unsigned FAR = 0x80000000u, FULL = 0xffffffffu, lt = (1u << lane) - 1u;  // row index < 2^15: bit 31 is free
        int nfar = 0;
        for (int c = 0; c < n; c += 32) {
          const int i = c + lane;
          bool far = false;
          if (i < n) {
            const unsigned pxy = C.reg_get(i);
            const int py = (int)(pxy >> 16), px = (int)(pxy & 0xffff);
            far = dist_sq_dev(xc, yc, (double)px, (double)py) > radSq;
            if (far) {
              if (MW) {
                C.mw_release(py * C.sw + px);
              } else if (BM) {
                C.bm_clear<M>(py * C.sw + px);
              } else {
                unsigned* wp = C.wptr(py * C.sw + px);
                *wp = *wp & ~USED_BIT;
              }
              C.reg_set(i, pxy | FAR);
            }
          }
          nfar += __popc(__ballot_sync(FULL, far));
        }
        __syncwarp();
        const int nk = n - nfar;
        // holes (far positions below nk, ascending) into scratch: the global list beyond n is free
        unsigned* holes = C.regG + n;
        int nh = 0;
        for (int c = 0; c < nk; c += 32) {
          const int i = c + lane;
          const bool hole = i < nk && (C.reg_get(i) & FAR);
          const unsigned b = __ballot_sync(FULL, hole);
          if (hole) holes[nh + __popc(b & lt)] = (unsigned)i;
          nh += __popc(b);
        }
        __syncwarp();
        // fillers (kept positions at or above nk, descending): the k-th goes to the k-th hole
        int nf = 0;
        for (int c = n - 1; c >= nk; c -= 32) {
          const int i = c - lane;
          unsigned v = 0;
          const bool fill = i >= nk && !((v = C.reg_get(i)) & FAR);
          const unsigned b = __ballot_sync(FULL, fill);
          if (fill) C.reg_set((int)holes[nf + __popc(b & lt)], v);
          nf += __popc(b);
        }
        n = nk;
      }
      __syncwarp();
      *n_out = n;
      if (n < 2) return false;
    }
    // ---- refine(), first half: tau = 2 * standard deviation of the level-line angles near the seed
    const unsigned seedxy = C.reg_get(0);
    const int sy = (int)(seedxy >> 16), sx = (int)(seedxy & 0xffff);
    xc = (double)sx;
    yc = (double)sy;
    const double ang_c = __dmul_rn((double)__uint_as_float(C.pix[sy * C.sw + sx].x), PL_DEG_TO_RADS);
    double* sA = C.stage;
    double* sF = C.stage + 32;
    double* sV2 = C.stage + 64;
    double acc = 0;
    int cntN = 0;
    for (int c = 0; c < n; c += 32) {
      const int i = c + lane;
      bool inside = false;
      if (i < n) {
        const unsigned pxy = C.reg_get(i);
        const int py = (int)(pxy >> 16), px = (int)(pxy & 0xffff);
        const int idx = py * C.sw + px;
        const uint4 r = C.pix[idx];
        if (MW) C.mw_release(idx);
        else if (BM) C.bm_clear<M>(idx);
        else *C.wptr(idx) = r.w & ~USED_BIT;
        double flag = 0.0, v = 0.0;
        if (sqrt(dist_sq_dev(xc, yc, (double)px, (double)py)) < rec->width) {
          const double ang = __dmul_rn((double)__uint_as_float(r.x), PL_DEG_TO_RADS);
          v = angle_diff_signed_dev(ang, ang_c);
          flag = 1.0;
        }
        sA[lane] = v;
        sF[lane] = flag;
        sV2[lane] = __dmul_rn(v, v);
        inside = flag != 0.0;
      }
      cntN += __popc(__ballot_sync(0xffffffffu, inside));
      __syncwarp();
      const int cnt = min(32, n - c);
      // two ordered chains (sum of v, sum of v*v over the flagged points): lane 0 and lane 1 own one each
      if (lane < 2) {
        const double* row = lane == 0 ? sA : sV2;
        for (int j = 0; j < cnt; ++j)
          if (sF[j] != 0.0) acc = __dadd_rn(acc, row[j]);
      }
      __syncwarp();
    }
    const double sum = __shfl_sync(0xffffffffu, acc, 0), s_sum = __shfl_sync(0xffffffffu, acc, 1);
    const double mean_angle = __ddiv_rn(sum, (double)cntN);
    const double tau = __dmul_rn(
        2.0, sqrt(__dadd_rn(__ddiv_rn(__dsub_rn(s_sum, __dmul_rn(__dmul_rn(2.0, mean_angle), sum)), (double)cntN),
                            __dmul_rn(mean_angle, mean_angle))));
    tol = tau;
    phase = 1;
  }
}

constexpr int GROW_WARPS = 4;  // frames per CTA at most (one warp each): keeps the long-running kernel from holding every CTA slot of an SM
// dynamic shared memory per frame: staging doubles, the head of the region list, and in GM_BMS the frame's bitmap
__host__ __device__ inline size_t grow_smem_per_frame(int P, bool bitmap) {
  return 96 * sizeof(double) + REG_SMEM * sizeof(unsigned) + (bitmap ? (size_t)((P + 31) / 32 + 3) / 4 * 16 : 0);
}
template <int M>
__global__ void __launch_bounds__(32 * GROW_WARPS) k_lsd_grow(const __grid_constant__ LineParams L, uint4* pixAll,
                                                 unsigned* bmAll, const unsigned* __restrict__ seedsAll,
                                                 const int* __restrict__ nseeds, unsigned* regAll,
                                                 LsdRect* __restrict__ rectsAll, int* __restrict__ nrects,
                                                 int* __restrict__ status) {
  constexpr bool BM = M == GM_BMS || M == GM_BMG;
  extern __shared__ __align__(16) unsigned char grow_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * (blockDim.x >> 5) + warp;
  if (f >= L.batch) return;
  unsigned char* mySmem = grow_smem + (size_t)warp * grow_smem_per_frame(L.P, M == GM_BMS);
  GrowCtx C;
  C.stage = reinterpret_cast<double*>(mySmem);
  C.regS = reinterpret_cast<unsigned*>(mySmem + 96 * sizeof(double));
  C.bm = nullptr;
  if (BM) {
    const int bmWords = (L.P + 31) / 32;
    unsigned* gbm = bmAll + (size_t)f * ((bmWords + 3) / 4 * 4);
    if (M == GM_BMS) {
      uint4* dst = reinterpret_cast<uint4*>(C.regS + REG_SMEM);
      const uint4* src = reinterpret_cast<const uint4*>(gbm);
      for (int i = lane; i < (bmWords + 3) / 4; i += 32) dst[i] = __ldcs(src + i);
      C.bm = reinterpret_cast<unsigned*>(dst);
      __syncwarp();
    } else {
      C.bm = gbm;
    }
  }
  C.prefetch = (L.grow_variant & 2) != 0;
  C.prefetch2 = (L.grow_variant & 4) != 0;
  C.pix = pixAll + (size_t)f * L.P;
  C.regG = regAll + (size_t)f * L.P;
  C.sw = L.sw;
  C.sh = L.sh;
  C.P = L.P;
  C.lane = lane;
#ifdef PLSLAM_GROW_PROF
  long long prof[16];
  for (int k = 0; k < 16; ++k) prof[k] = 0;
  C.prof = prof;
  const long long gp_k0 = clock64();
  long long gp_t0 = 0;
#endif
  const unsigned* seeds = seedsAll + (size_t)f * L.P;
  const int ns = nseeds[f];
  LsdRect* rects = rectsAll + (size_t)f * L.rect_cap;
  int nrect = 0;
  int snext = lane < ns ? (int)seeds[lane] : -1;
  for (int base = 0; base < ns; base += 32) {
    int s = snext;
    snext = base + 32 + lane < ns ? (int)seeds[base + 32 + lane] : -1;  // prefetch the next chunk of seeds
    while (true) {
      const bool unused = s >= 0 && (BM ? !C.bm_test<M>(s) : !(*C.wptr(s) & USED_BIT));
      const unsigned m = __ballot_sync(0xffffffffu, unused);
      if (!m) break;
      const int j = __ffs(m) - 1;
      const int seed = __shfl_sync(0xffffffffu, s, j);
      if (lane <= j) s = -1;
      if (lane == 0) C.reg_set(0, ((unsigned)(seed / L.sw) << 16) | (unsigned)(seed % L.sw));
      __syncwarp();
      LsdRect rec;
      int n = 0;
      if (!lsd_seed_region<M>(C, L.prec, L.p, L.density_th, L.min_reg_size, L.grow_variant, &rec, &n)) continue;
      if (nrect < L.rect_cap) {
        if (lane == 0) rects[nrect] = rec;
      } else if (lane == 0) {
        atomicMax(status, PLSLAM_ERR_OVERFLOW);
      }
      ++nrect;
    }
  }
  if (lane == 0) nrects[f] = min(nrect, L.rect_cap);
#ifdef PLSLAM_GROW_PROF
  prof[4] = clock64() - gp_k0;
  if (lane == 0)
    for (int k = 0; k < 16; ++k) atomicAdd(&g_grow_prof[k], (unsigned long long)prof[k]);
#endif
}


// ------------------------------------------------------------------------------------------
// k_lsd_grow_sw: region growing with one CTA per frame: K - 1 worker warps and one retirement warp, NO round barriers.
// Regions are grown speculatively in parallel and RETIRE IN SEED ORDER through a window of WS seeds (slot = rank % WS),
// which keeps the result identical to the sequential loop of lsd.cpp:
//   * a worker takes the next GROUP of 32 consecutive seeds (an atomic counter), classifies them with one look at the
//     owner plane and grows the free ones one after the other (region_grow, region2rect, refine as in the sequential
//     loop), so seeds of one edge that sit next to each other in the list are absorbed as they would be sequentially;
//   * a region marks its pixels with (tag << 1), tag = seed rank + 1 (GrowCtx).  Tags below the head tag are final.  A
//     region that takes a pixel held or given back by a YOUNGER in-flight region robs it (atomicMin) and poisons that
//     region; a region that wants a pixel held or given back by an OLDER in-flight region poisons itself;
//   * a seed is passed over (DIRTY) when an older region still in flight holds it, or when it lies on the level line of
//     an older seed that is pending or being grown (same-edge heuristic: it will most likely be absorbed).  When a DIRTY
//     seed reaches the head of the window it must be held by a region that does not come after it, else it is grown
//     then — as the head, with nothing older in flight, i.e. exactly;
//   * a poisoned region gives its pixels back (squashed while running: from its list; after it finished: from the copy
//     of its list in the frame's pool) and is grown again once its blocker has retired, or at once at the head.
// The retirement warp handles 32 seeds per step (states are per seed) and never looks at the seed list except to
// validate DIRTY seeds; scanning is done by the workers.
// ------------------------------------------------------------------------------------------
enum { SW_EMPTY = 0, SW_PENDING = 1, SW_RUNNING = 2, SW_DONE = 3, SW_DIRTY = 4, SW_SQUASHED = 5, SW_QUEUED = 6 };
constexpr int SW_RQN = 128;  // re-issue queue entries (ranks)
enum { SWS_GROUPS = 0, SWS_REGIONS, SWS_SQUASH_RUN, SWS_SQUASH_DONE, SWS_INSERT, SWS_RECTS, SWS_VALSTEPS, SWS_SCHED_IDLE,
       SWS_WORK_IDLE, SWS_DIRTY, SWS_RERUNS, SWS_FRAMES,
       // cycle accounting (clock64): wall time of the frame, and where workers / the retirement warp spent it
       SWS_CYC_KERNEL, SWS_CYC_W_WAIT, SWS_CYC_W_GROW, SWS_CYC_W_SQUASH, SWS_CYC_W_SCAN, SWS_CYC_S_RETIRE, SWS_CYC_S_PLIST,
       SWS_CYC_S_IDLE, SWS_PIX_DONE, SWS_PIX_SQUASH, SWS_WINFULL, SWS_POOLFULL, SWS_HEUR_SKIP, SWS_P_HELD, SWS_P_TOMB, SWS_P_ROB,
       SWS_N };
__device__ unsigned long long g_aw_stats[32];
// "same edge" heuristic (it decides which seeds are passed over for now, never the result): a seed closer than
// g_sw_perp px to the level line through an older seed that is pending or being grown, within g_sw_along px along it, with
// a level-line angle within g_sw_ang degrees, will most likely be absorbed by that seed's region
// Measured (tools/sweep_sw.sh, 16 frames, K = 8 / 16): holding seeds back costs more (they are grown one by one at the
// head when the guess was wrong) than the squashed regions it avoids, so the heuristic is OFF by default (perp = 0);
// PLSLAM_SW_PERP / _ALONG / _ANG switch it on for experiments.
__device__ float g_sw_perp = 0.f, g_sw_along = 0.f, g_sw_ang = 0.f;

template <int K, int WS>
struct SwState {
  int st[WS];         // per seed of the window
  int poison[WS];
  int blocker[WS];    // tag of the older region that poisoned the seed's region
  int rect[WS];       // DONE: index of the region's rectangle in the frame's pool, -1 = none
  int lstOff[WS], lstCnt[WS];  // DONE: copy of the region's pixel list in the frame's list pool
  int plist[WS];      // ring of poisoned slots (slot + 1), written by whoever poisons
  unsigned plTail;
  // seeds that are pending or being grown, one row per worker warp: tag (0 = none), position, unit vector along the
  // level line, level-line angle
  unsigned pendTag[K - 1][32];
  float pendX[K - 1][32], pendY[K - 1][32], pendC[K - 1][32], pendS[K - 1][32], pendD[K - 1][32];
  int rq[SW_RQN];     // re-issue queue: rank + 1, 0 = empty
  unsigned rqHead, rqTail;
  unsigned headTag;
  int headRank, nextGroup;
  int rectTop, listTop;
  int done, abort;
  unsigned long long stats[SWS_N];
};
template <int K, int WS>
constexpr size_t sw_smem_bytes() {
  return ((sizeof(SwState<K, WS>) + 15) & ~size_t(15)) + (sizeof(double) * 96 + sizeof(unsigned) * REG_SMEM) * (size_t)(K - 1);
}
constexpr long long SW_WATCHDOG = 1ll << 21;  // polls (>= 40 ns each) without progress before the kernel gives up
// internal consistency checks: a failure raises status 100 + code and stops the frame's CTA
#define SW_ASSERT(cond, code)                                                   \
  do {                                                                          \
    if (!(cond)) {                                                              \
      atomicMax(status, 100 + (code));                                          \
      *reinterpret_cast<volatile int*>(&S.abort) = 1;                           \
    }                                                                           \
  } while (0)

template <int K, int WS>
__device__ void sw_retire_warp(SwState<K, WS>& S, const LineParams& L, const unsigned* own, const unsigned* __restrict__ seeds,
                               int ns, const LsdRect* rectPool, LsdRect* __restrict__ rects, int* nrectOut, int* status,
                               int lane) {
  const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;
  volatile int* vst = S.st;
  volatile int* vpoison = S.poison;
  volatile int* vblocker = S.blocker;
  volatile int* vrq = S.rq;
  volatile int* vplist = S.plist;
  int head = 0, nrect = 0;  // head = rank of the oldest seed that has not retired
  unsigned rqTail = 0, plHead = 0, plSeenTail = 0;
  int plSeenHead = -1;
  long long idle = 0;
  const long long tKernel = clock64();
  long long tPhase = tKernel;
#define SW_STAT(k, v) do { if (lane == 0) S.stats[k] += (v); } while (0)
  auto requeue = [&](int rank) {  // the seed `rank` is grown (again) as a job of its own
    if (lane == 0) {
      vst[rank % WS] = SW_QUEUED;
      __threadfence_block();
      long long spins = 0;
      while (vrq[rqTail % SW_RQN] != 0 && ++spins < SW_WATCHDOG) __nanosleep(20);
      SW_ASSERT(spins < SW_WATCHDOG, 1);
      vrq[rqTail % SW_RQN] = rank + 1;
      __threadfence_block();
      *reinterpret_cast<volatile unsigned*>(&S.rqTail) = rqTail + 1;
    }
    ++rqTail;
    SW_STAT(SWS_RERUNS, 1);
    __syncwarp();
  };
  while (head < ns) {
    bool progress = false;
    if (*reinterpret_cast<volatile int*>(&S.abort)) break;
    // ---- retire the longest prefix of the 32 oldest seeds that is finished and not poisoned ----
    {
      const int rank = head + lane;
      const bool valid = rank < ns;
      const int s = rank % WS;
      const int st = valid ? vst[s] : SW_EMPTY;
      __threadfence_block();
      const bool pz = valid && vpoison[s] != 0;
      const bool fin = valid && (st == SW_DONE || st == SW_DIRTY) && !pz;
      const unsigned finM = __ballot_sync(FULL, fin);
      int prefix = __ffs(~finM) - 1;  // (finM == FULL gives 32 - 1 + 1: ffs(0) = 0 -> -1; handled below)
      if (finM == FULL) prefix = 32;
      // DIRTY seeds of the prefix: each must be held by a region that does not come after it
      const bool chk = fin && st == SW_DIRTY && lane < prefix;
      if (__any_sync(FULL, chk)) {
        bool fail = false;
        if (chk) {
          const unsigned v = __ldcg(own + seeds[rank]);
          fail = (v & 1u) || (v >> 1) > (unsigned)rank + 1u;
        }
        SW_STAT(SWS_VALSTEPS, 1);
        const unsigned failM = __ballot_sync(FULL, fail);
        if (failM) {  // the sequential loop would have grown this seed: it is grown now, as the head
          const int f = __ffs(failM) - 1;
          prefix = f;
          if (lane == 0) {
            S.lstCnt[(head + f) % WS] = 0;
            vblocker[(head + f) % WS] = 0;
          }
          SW_STAT(SWS_INSERT, 1);
          requeue(head + f);
          progress = true;
        }
      }
      if (prefix > 0) {
        // rectangles of the retiring regions, in seed order
        const int ri = (lane < prefix && st == SW_DONE) ? S.rect[s] : -1;
        const unsigned rectM = __ballot_sync(FULL, ri >= 0);
        constexpr int RD = (int)(sizeof(LsdRect) / sizeof(double));
        int pos = nrect;
        for (unsigned r = rectM; r; r &= r - 1u, ++pos) {
          const int src = __shfl_sync(FULL, ri, __ffs(r) - 1);
          if (pos < L.rect_cap) {
            if (lane < RD) reinterpret_cast<double*>(rects + pos)[lane] = __ldcg(reinterpret_cast<const double*>(rectPool + src) + lane);
          } else if (lane == 0) {
            atomicMax(status, PLSLAM_ERR_OVERFLOW);
          }
        }
        SW_STAT(SWS_RECTS, __popc(rectM));
        nrect = pos;
        if (lane < prefix) {
          vpoison[s] = 0;
          vst[s] = SW_EMPTY;
        }
        head += prefix;
        __syncwarp();
        __threadfence_block();
        if (lane == 0) {
          *reinterpret_cast<volatile unsigned*>(&S.headTag) = (unsigned)head + 1u;
          *reinterpret_cast<volatile int*>(&S.headRank) = head;
        }
        progress = true;
      }
      // the seed now at the head: poisoned or squashed -> grown again at once (nothing older is in flight any more)
      if (head < ns) {
        const int hs = head % WS;
        const int hst = vst[hs];
        const bool hpz = vpoison[hs] != 0;
        if (hst == SW_SQUASHED || ((hst == SW_DONE || hst == SW_DIRTY) && hpz)) {
          if (hst != SW_SQUASHED) SW_STAT(SWS_SQUASH_DONE, 1);
          requeue(head);
          progress = true;
        }
      }
    }
    { const long long t = clock64(); SW_STAT(SWS_CYC_S_RETIRE, t - tPhase); tPhase = t; }
    // ---- poisoned regions behind the head are grown again once the region that blocked them has retired ----
    {
      // entries that must wait for their blocker go back to the ring; the ring is looked at again when a new entry has
      // arrived or the head has moved since, not in a busy loop
      const unsigned tl = *reinterpret_cast<volatile unsigned*>(&S.plTail);
      if (tl != plHead && (tl != plSeenTail || head != plSeenHead)) {
        unsigned repushed = 0;
        while (plHead != tl) {
          const unsigned todo = min(tl - plHead, 32u);
          int e = 0;
          if ((unsigned)lane < todo) {
            long long spins = 0;
            while ((e = vplist[(plHead + lane) % WS]) == 0 && ++spins < SW_WATCHDOG) {}
            vplist[(plHead + lane) % WS] = 0;
          }
          plHead += todo;
          const int s = e - 1;
          int kind = 0;  // 1 requeue, 2 keep waiting
          int rank = 0;
          if (e > 0) {
            const int st = vst[s];
            const bool pz = vpoison[s] != 0;
            rank = head + ((s - head % WS) + WS) % WS;
            if (st == SW_SQUASHED || ((st == SW_DONE || st == SW_DIRTY) && pz)) {
              kind = ((unsigned)vblocker[s] < (unsigned)head + 1u) ? 1 : 2;
              if (rank == head) kind = 0;  // the head is handled above
            } else if ((st == SW_RUNNING || st == SW_PENDING) && pz) {
              kind = 2;  // its worker has not noticed yet
            }
          }
          unsigned m = __ballot_sync(FULL, kind == 1);
          while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1u;
            const int rj = __shfl_sync(FULL, rank, j), sj = __shfl_sync(FULL, s, j);
            const int stj = vst[sj];  // read again: two ring entries may name the same slot
            if (!(stj == SW_SQUASHED || ((stj == SW_DONE || stj == SW_DIRTY) && vpoison[sj] != 0))) continue;
            if (stj != SW_SQUASHED) SW_STAT(SWS_SQUASH_DONE, 1);
            requeue(rj);
            progress = true;
          }
          if (kind == 2) {
            const unsigned p = atomicAdd(&S.plTail, 1u);
            vplist[p % WS] = e;
          }
          repushed += __popc(__ballot_sync(FULL, kind == 2));
        }
        plSeenTail = tl + repushed;  // a push by a worker meanwhile makes the tail differ and the ring is looked at again
        plSeenHead = head;
      }
    }
    { const long long t = clock64(); SW_STAT(SWS_CYC_S_PLIST, t - tPhase); tPhase = t; }
    if (!progress) {
      SW_STAT(SWS_SCHED_IDLE, 1);
      __nanosleep(40);
      { const long long t = clock64(); SW_STAT(SWS_CYC_S_IDLE, t - tPhase); tPhase = t; }
      if (++idle > SW_WATCHDOG) {
        if (lane == 0) {
          atomicMax(status, PLSLAM_ERR_INTERNAL);
          *reinterpret_cast<volatile int*>(&S.abort) = 1;
        }
        break;
      }
    } else {
      idle = 0;
    }
  }
  SW_STAT(SWS_CYC_KERNEL, clock64() - tKernel);
  if (lane == 0) {
    *nrectOut = min(nrect, L.rect_cap);
    __threadfence_block();
    *reinterpret_cast<volatile int*>(&S.done) = 1;
  }
#undef SW_STAT
}

// One worker warp: take a job (re-issued seeds first, else the next group of 32 seeds), grow its free seeds in order,
// publish every seed's state, repeat.
template <int K, int WS>
__device__ void sw_worker(SwState<K, WS>& S, GrowCtx& C, const LineParams& L, const unsigned* __restrict__ seeds, int ns,
                          unsigned* listPool, int listPoolCap, LsdRect* rectPool, int* status) {
  const unsigned FULL = 0xffffffffu;
  const int lane = C.lane;
  volatile int* vst = S.st;
  volatile int* vrq = S.rq;
  volatile int* vdone = &S.done;
  volatile int* vabort = &S.abort;
  volatile int* vheadRank = &S.headRank;
  const int me = (int)(threadIdx.x >> 5) - 1;  // my row of the table of pending seeds
  const int nGroups = (ns + SW_G - 1) / SW_G;
  int fresh = -1;  // a group taken from the counter that still waits for room in the window
  bool freshLeft = true;
  while (true) {
    // ---- acquire ----
    const long long tWait = clock64();
    int a = -1, cnt = 0;  // the job: seeds [a, a + cnt)
    bool single = false;
    long long spins = 0;
    while (true) {
      int gotA = -1, gotSingle = 0, stop = 0;
      if (lane == 0) {
        const unsigned h = *reinterpret_cast<volatile unsigned*>(&S.rqHead), tl = *reinterpret_cast<volatile unsigned*>(&S.rqTail);
        if (h != tl) {
          if (atomicCAS(&S.rqHead, h, h + 1u) == h) {
            const int v = vrq[h % SW_RQN];
            vrq[h % SW_RQN] = 0;
            gotA = v - 1;
            gotSingle = 1;
          }
        } else {
          if (fresh < 0 && freshLeft) {
            fresh = atomicAdd(&S.nextGroup, 1);
            if (fresh >= nGroups) { fresh = -1; freshLeft = false; }
          }
          if (fresh >= 0 && (fresh + 1) * SW_G <= *vheadRank + WS) {
            gotA = fresh * SW_G;
            fresh = -1;
          } else if (fresh >= 0) {
            S.stats[SWS_WINFULL] += 1;  // racy between workers: a statistic only
          }
        }
        stop = *vdone | *vabort;
      }
      gotA = __shfl_sync(FULL, gotA, 0);
      if (gotA >= 0) {
        a = gotA;
        single = __shfl_sync(FULL, gotSingle, 0) != 0;
        cnt = single ? 1 : min(SW_G, ns - a);
        break;
      }
      if (__shfl_sync(FULL, stop, 0)) return;
      __nanosleep(40);
      if (++spins > SW_WATCHDOG) {
        if (lane == 0) {
          atomicMax(status, PLSLAM_ERR_INTERNAL);
          *vabort = 1;
        }
        return;
      }
    }
    const long long tJob = clock64();
    if (lane == 0) {
      atomicAdd(&S.stats[SWS_WORK_IDLE], (unsigned long long)spins);
      atomicAdd(&S.stats[SWS_CYC_W_WAIT], (unsigned long long)(tJob - tWait));
    }
    if (a < 0 || a >= ns) {
      SW_ASSERT(false, 2);
      return;
    }
    __threadfence_block();
    const bool mineLane = lane < cnt;
    const int rank = a + lane, slot = rank % WS;
    const unsigned tagL = (unsigned)rank + 1u;
    int sd = mineLane ? (int)seeds[rank] : -1;
    if (single) {
      // a region that finished and was poisoned afterwards still holds its pixels: give them back (pool copy of its list)
      const int lc = S.lstCnt[a % WS], lo = S.lstOff[a % WS];
      const unsigned mark = ((unsigned)a + 1u) << 1;
      for (int i = lane; i < lc; i += 32) {
        const unsigned pxy = listPool[lo + i] & 0x7fffffffu;
        atomicCAS(C.own + (int)(pxy >> 16) * C.sw + (int)(pxy & 0xffff), mark, 0xffffffffu);
      }
      __threadfence();
      if (lane == 0) {
        S.lstCnt[a % WS] = 0;
        S.rect[a % WS] = -1;
        C.blocker[a % WS] = 0;
        *reinterpret_cast<volatile int*>(C.poison + a % WS) = 0;
        __threadfence_block();
        vst[a % WS] = SW_PENDING;
      }
    } else if (mineLane) {
      S.lstCnt[slot] = 0;
      S.rect[slot] = -1;
      C.blocker[slot] = 0;
      vst[slot] = SW_PENDING;  // (the slot was EMPTY with its poison flag cleared: the window had room for this group)
    }
    __syncwarp();
    // seed positions and level lines (same-edge heuristic)
    float sxf = 0.f, syf = 0.f, sdeg = 0.f, scs = 0.f, ssn = 0.f;
    if (sd >= 0) {
      const uint4 r = C.pix[sd];
      const int yy = sd / C.sw;
      sxf = (float)(sd - yy * C.sw);
      syf = (float)yy;
      sdeg = __uint_as_float(r.x);
      scs = __uint_as_float(r.y);
      ssn = __uint_as_float(r.z);
      S.pendX[me][lane] = sxf;
      S.pendY[me][lane] = syf;
      S.pendC[me][lane] = scs;
      S.pendS[me][lane] = ssn;
      S.pendD[me][lane] = sdeg;
    }
    bool published = false;  // pendTag row written (after the first classification: only seeds that look free)
    int nreg = 0;
    long long cycGrow = 0, cycSquash = 0;
    // Pass 0 walks the job's seeds in order; seeds the same-edge heuristic holds back are looked at once more in pass 1,
    // when the older seeds they waited for have most likely been grown (what is still held back then stays DIRTY).
    unsigned deferM = 0;
    int sdd = -1;
    for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) {
      if (!deferM) break;
      sd = ((deferM >> lane) & 1u) ? sdd : -1;
      published = false;
    }
    while (true) {
      const unsigned h0 = *C.headTag;
      const unsigned v = sd >= 0 ? __ldcg(C.own + sd) : 0u;
      const unsigned t = v >> 1;
      // 0 = not mine / dropped, 1 = grow, 2 = used for good (held by a retired region), 3 = dirty (an older region in
      // flight holds it or gave it back: its fate decides)
      int cls = 0;
      if (sd >= 0) {
        if (v == 0xffffffffu) cls = 1;
        else if (t < h0) cls = (v & 1u) ? 1 : 2;
        else if (t < tagL) cls = 3;
        else cls = 1;  // mine from an earlier run (tombstone), or a younger region: the take robs it
      }
      if (cls == 2) {
        vst[slot] = SW_DONE;
        sd = -1;
      }
      if (!published) {
        *reinterpret_cast<volatile unsigned*>(&S.pendTag[me][lane]) = cls == 1 ? tagL : 0u;
        published = true;
      }
      const unsigned gm = __ballot_sync(FULL, cls == 1);
      const int j = gm ? __ffs(gm) - 1 : 32;
      if (cls == 3 && lane < j) {  // passed over
        vst[slot] = SW_DIRTY;
        sd = -1;
        atomicAdd(&S.stats[SWS_DIRTY], 1ull);
      }
      if (!gm) break;
      const int seed = __shfl_sync(FULL, sd, j);
      const unsigned myTag = (unsigned)(a + j) + 1u;
      const int mySlot = (a + j) % WS;
      if (lane == j) {
        sdd = sd;
        sd = -1;
      }
      {
        // same-edge heuristic against the older seeds other workers have pending or are growing
        const float jx = __shfl_sync(FULL, sxf, j), jy = __shfl_sync(FULL, syf, j), jd = __shfl_sync(FULL, sdeg, j);
        bool hit = false;
        for (int w = 0; w < K - 1; ++w) {  // (warp-uniform trip count: the vote below needs every lane)
          if (w == me) continue;
          const unsigned at = *reinterpret_cast<volatile unsigned*>(&S.pendTag[w][lane]);
          if (at != 0u && at < myTag) {
            const float dx = jx - S.pendX[w][lane], dy = jy - S.pendY[w][lane];
            const float perp = fabsf(dx * S.pendS[w][lane] - dy * S.pendC[w][lane]);
            const float along = fabsf(dx * S.pendC[w][lane] + dy * S.pendS[w][lane]);
            float dd = fabsf(jd - S.pendD[w][lane]);
            if (dd > 180.f) dd = 360.f - dd;
            hit = hit || (perp < g_sw_perp && along < g_sw_along && dd < g_sw_ang);
          }
        }
        __syncwarp();
        if (__any_sync(FULL, hit)) {
          if (lane == 0) {
            *reinterpret_cast<volatile unsigned*>(&S.pendTag[me][j]) = 0u;
            if (pass == 1) vst[mySlot] = SW_DIRTY;
            atomicAdd(&S.stats[SWS_HEUR_SKIP], 1ull);
          }
          if (pass == 0) deferM |= 1u << j;
          continue;
        }
      }
      const long long tReg = clock64();
      C.myVal = myTag << 1;
      C.slot = mySlot;
      if (lane == 0) {
        vst[mySlot] = SW_RUNNING;
        C.reg_set(0, ((unsigned)(seed / C.sw) << 16) | (unsigned)(seed % C.sw));
      }
      __syncwarp();
      double reg_angle;
      int n = lsd_region_grow_spec<GM_MW>(C, L.prec, &reg_angle);
      ++nreg;
      bool keep = false;
      LsdRect rec;
      if (!C.mw_poisoned() && n >= L.min_reg_size) {
        lsd_region2rect(C, n, reg_angle, L.prec, L.p, &rec);
        keep = lsd_refine<GM_MW>(C, &n, reg_angle, L.prec, L.p, &rec, L.density_th, 1);
      }
      if (lane == 0) *reinterpret_cast<volatile unsigned*>(&S.pendTag[me][j]) = 0u;
      bool squashed = C.mw_poisoned();
      int roff = -1, loff = 0;
      if (!squashed) {
        // publish: the rectangle and the list (for a release after DONE) go to the frame's pools
        if (lane == 0) {
          if (keep) roff = atomicAdd(&S.rectTop, 1);
          if (n) loff = atomicAdd(&S.listTop, n);
        }
        roff = __shfl_sync(FULL, roff, 0);
        loff = __shfl_sync(FULL, loff, 0);
        if (roff >= L.rect_cap) {  // rectangle pool exhausted (runs that were squashed leak their entries)
          if (lane == 0) atomicMax(status, PLSLAM_ERR_OVERFLOW);
          roff = -1;
        }
        if (roff >= 0 && lane == 0) rectPool[roff] = rec;
        if (loff + n <= listPoolCap) {
          for (int i = lane; i < n; i += 32) listPool[loff + i] = C.reg_get(i);
        } else {
          // list pool exhausted: the region could not be released after DONE, so it only publishes as the head
          if (lane == 0) S.stats[SWS_POOLFULL] += 1;
          long long sp = 0;
          while (*vheadRank != a + j && !C.mw_poisoned() && !(*vabort) && ++sp < SW_WATCHDOG) __nanosleep(100);
          squashed = C.mw_poisoned() || *vheadRank != a + j;
          loff = 0;
          if (!squashed) n = 0;  // (published with an empty list: the head cannot be poisoned any more)
        }
      }
      if (squashed) {
        for (int i = lane; i < n; i += 32) {
          const unsigned pxy = C.reg_get(i) & 0x7fffffffu;
          atomicCAS(C.own + (int)(pxy >> 16) * C.sw + (int)(pxy & 0xffff), C.myVal, 0xffffffffu);
        }
      }
      __threadfence();  // owner-plane updates (releases are fire-and-forget) and pool copies are performed before the state is published
      if (lane == 0) {
        if (squashed) {
          vst[mySlot] = SW_SQUASHED;  // (its slot is in the poison ring: the retirement warp re-issues it)
          atomicAdd(&S.stats[SWS_SQUASH_RUN], 1ull);
          atomicAdd(&S.stats[SWS_PIX_SQUASH], (unsigned long long)n);
        } else {
          S.rect[mySlot] = roff;
          S.lstOff[mySlot] = loff;
          S.lstCnt[mySlot] = n;
          __threadfence_block();
          vst[mySlot] = SW_DONE;
          atomicAdd(&S.stats[SWS_PIX_DONE], (unsigned long long)n);
        }
      }
      (squashed ? cycSquash : cycGrow) += clock64() - tReg;
      __syncwarp();
    }
    }
    if (lane == 0) {
      atomicAdd(&S.stats[SWS_GROUPS], single ? 0ull : 1ull);
      atomicAdd(&S.stats[SWS_REGIONS], (unsigned long long)nreg);
      atomicAdd(&S.stats[SWS_CYC_W_GROW], (unsigned long long)cycGrow);
      atomicAdd(&S.stats[SWS_CYC_W_SQUASH], (unsigned long long)cycSquash);
      atomicAdd(&S.stats[SWS_CYC_W_SCAN], (unsigned long long)(clock64() - tJob - cycGrow - cycSquash));
    }
    __syncwarp();
  }
}

template <int K, int WS>
__global__ void __launch_bounds__(32 * K) k_lsd_grow_sw(const __grid_constant__ LineParams L, uint4* pixAll, unsigned* ownAll,
                                                       const unsigned* __restrict__ seedsAll,
                                                       const int* __restrict__ nseeds, unsigned* listAll, unsigned* listPoolAll,
                                                       LsdRect* rectPoolAll, LsdRect* __restrict__ rectsAll,
                                                       int* __restrict__ nrects, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char sw_smem[];
  SwState<K, WS>& S = *reinterpret_cast<SwState<K, WS>*>(sw_smem);
  unsigned char* wbase = sw_smem + ((sizeof(SwState<K, WS>) + 15) & ~size_t(15));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x;
  for (int i = threadIdx.x; i < WS; i += 32 * K) {
    S.st[i] = SW_EMPTY;
    S.poison[i] = 0;
    S.blocker[i] = 0;
    S.rect[i] = -1;
    S.lstOff[i] = S.lstCnt[i] = 0;
    S.plist[i] = 0;
  }
  for (int i = threadIdx.x; i < (K - 1) * 32; i += 32 * K) S.pendTag[i / 32][i % 32] = 0u;
  if (threadIdx.x < SW_RQN) S.rq[threadIdx.x] = 0;
  if (threadIdx.x < SWS_N) S.stats[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    S.plTail = 0;
    S.rqHead = S.rqTail = 0;
    S.headTag = 1u;
    S.headRank = 0;
    S.nextGroup = 0;
    S.rectTop = 0;
    S.listTop = 0;
    S.done = 0;
    S.abort = 0;
  }
  __syncthreads();
  const unsigned* seeds = seedsAll + (size_t)f * L.P;
  const int ns = nseeds[f];
  if (warp == 0) {
    sw_retire_warp<K, WS>(S, L, ownAll + (size_t)f * L.P, seeds, ns, rectPoolAll + (size_t)f * L.rect_cap,
                          rectsAll + (size_t)f * L.rect_cap, nrects + f, status, lane);
  } else {
    GrowCtx C;
#ifdef PLSLAM_GROW_PROF
    long long prof[16];
    for (int k = 0; k < 16; ++k) prof[k] = 0;
    C.prof = prof;
#endif
    unsigned char* mine = wbase + (size_t)(warp - 1) * (sizeof(double) * 96 + sizeof(unsigned) * REG_SMEM);
    C.stage = reinterpret_cast<double*>(mine);
    C.regS = reinterpret_cast<unsigned*>(mine + sizeof(double) * 96);
    C.prefetch = false;
    C.pix = pixAll + (size_t)f * L.P;
    C.own = ownAll + (size_t)f * L.P;
    // per worker: a list area of 2 P entries (region list + refine() scratch); per frame: list pool of 2 P entries and
    // rectangle pool of rect_cap entries
    C.regG = listAll + ((size_t)f * (K - 1) + (warp - 1)) * 2 * (size_t)L.P;
    C.cap = 2 * L.P;
    C.sw = L.sw;
    C.sh = L.sh;
    C.P = L.P;
    C.lane = lane;
    C.poison = S.poison;
    C.blocker = S.blocker;
    C.plist = S.plist;
    C.plTail = &S.plTail;
    C.pstat = &S.stats[SWS_P_HELD];
    C.headTag = &S.headTag;
    C.nslots = WS;
    C.slot = 0;
    C.myVal = 0;
    sw_worker<K, WS>(S, C, L, seeds, ns, listPoolAll + (size_t)f * 2 * (size_t)L.P, 2 * L.P, rectPoolAll + (size_t)f * L.rect_cap,
                     status);
  }
  __syncthreads();
  if (threadIdx.x < SWS_N) {
    const unsigned long long v = threadIdx.x == SWS_FRAMES ? 1ull : S.stats[threadIdx.x];
    if (v) atomicAdd(&g_aw_stats[threadIdx.x], v);
  }
}

// ------------------------------------------------------------------------------------------
// k_lsd_nfa: rect_improve() / rect_nfa() / nfa(), one warp per rectangle.  The row scan is the one
// compiled into cv2 4.13 (see oracle/lsd_oracle.cc rect_nfa_rows); lanes split the rows or the
// columns of the scan, counts are integer so the reduction order is irrelevant.  nfa() uses the CUDA
// double-precision log/exp/pow/sinh: its value only feeds comparisons.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool double_equal_dev(double a, double b) {
  if (a == b) return true;
  const double abs_diff = fabs(a - b), aa = fabs(a), bb = fabs(b);
  double abs_max = aa > bb ? aa : bb;
  if (abs_max < 2.2250738585072014e-308) abs_max = 2.2250738585072014e-308;
  return (abs_diff / abs_max) <= (100.0 * 2.2204460492503131e-16);
}
// log_gamma / nfa of lsd.cpp.  Integer powers are formed by repeated multiplication (pow() differs from
// it by ulps only, and the NFA value feeds comparisons, never an output).
__device__ __noinline__ double log_gamma_dev(double x) {
  if (x > 15.0) {
    const double x2 = x * x;
    return 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * (x2 * x2 * x2)));
  }
  const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
  double a = (x + 0.5) * log(x + 5.5) - (x + 5.5);
  double b = 0, xn = 1.0;
  for (int n = 0; n < 7; ++n) {
    a -= log(x + (double)n);
    b += q[n] * xn;
    xn *= x;
  }
  return a + log(b);
}
// nfa() only ever asks log_gamma for integer arguments (n+1, k+1, n-k+1 with n = pixels of a rectangle), so the
// values are tabulated once per process by k_lsd_lgamma_table with the very same device function.
constexpr int LGAMMA_TABLE = 1 << 15;
__device__ double g_lgamma[LGAMMA_TABLE];
__global__ void k_lsd_lgamma_table() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < LGAMMA_TABLE) g_lgamma[i] = i > 0 ? log_gamma_dev((double)i) : 0.0;
}
__device__ __forceinline__ double log_gamma_int(int x) { return x < LGAMMA_TABLE ? g_lgamma[x] : log_gamma_dev((double)x); }

__device__ __noinline__ double nfa_dev(int n, int k, double p, double LOG_NT) {
  if (n == 0 || k == 0) return -LOG_NT;
  if (n == k) return -LOG_NT - (double)n * log10(p);
  const double p_term = p / (1 - p);
  const double log1term = log_gamma_int(n + 1) - log_gamma_int(k + 1) - log_gamma_int(n - k + 1) +
                          (double)k * log(p) + (double)(n - k) * log(1.0 - p);
  double term = exp(log1term);
  if (double_equal_dev(term, 0)) {
    if (k > n * p) return -log1term / 2.30258509299404568402 - LOG_NT;
    return -LOG_NT;
  }
  double bin_tail = term;
  const double tolerance = 0.1;
  for (int i = k + 1; i <= n; ++i) {
    const double bin_term = (double)(n - i + 1) / (double)i;
    const double mult_term = bin_term * p_term;
    term *= mult_term;
    bin_tail += term;
    if (bin_term < 1) {
      const double err = term * ((1 - pow(mult_term, (double)(n - i + 1))) / (1 - mult_term) - 1);
      if (err < tolerance * fabs(-log10(bin_tail) - LOG_NT) * bin_tail) break;
    }
  }
  return -log10(bin_tail) - LOG_NT;
}

// angular distance used by isAligned(); +inf for undefined pixels
__device__ __forceinline__ double lsd_ntheta(double theta, float deg) {
  if (deg == NOTDEF_F) return 1e300;
  const double a = __dmul_rn((double)deg, PL_DEG_TO_RADS);
  double n_theta = __dsub_rn(theta, a);
  if (n_theta < 0) n_theta = -n_theta;
  if (n_theta > M_3_2_PI_D) {
    n_theta = __dsub_rn(n_theta, M_2__PI_D);
    if (n_theta < 0) n_theta = -n_theta;
  }
  return n_theta;
}

// The part of a rectangle that rect_improve() changes and rect_nfa() reads, kept in registers (an LsdRect passed by
// reference lives on the thread's stack: the round-1 form of this kernel moved 7x more local than global memory).
struct NfaRect {
  double x1, y1, x2, y2, width, p, prec;
};

// Row scan of rect_nfa(): counts the pixels of the rectangle (total) and, for up to 5 angular
// tolerances at once, the aligned ones.  Warp-cooperative; results are warp-uniform.
template <int NPREC>
__device__ __forceinline__ void rect_count_dev(const NfaRect& r, double theta, double rdx, double rdy, const double (&precs)[5],
                                               const float* __restrict__ pix, int sw, int sh, int lane, int& total_out,
                                               int (&alg_out)[5]) {
  const double hw = __dmul_rn(0.5, r.width);
  const double dyhw = __dmul_rn(rdy, hw), dxhw = __dmul_rn(rdx, hw);
  double ux0 = __dsub_rn(r.x1, dyhw), ux1 = __dsub_rn(r.x2, dyhw), ux2 = __dadd_rn(r.x2, dyhw), ux3 = __dadd_rn(r.x1, dyhw);
  double uy0 = __dadd_rn(r.y1, dxhw), uy1 = __dadd_rn(r.y2, dxhw), uy2 = __dsub_rn(r.y2, dxhw), uy3 = __dsub_rn(r.y1, dxhw);
  // the corner that comes first in (y, x) order becomes corner 0; the rotation is done with selects so that the corners
  // stay in registers (an array indexed by the offset would go to local memory)
  int off = 0;
  {
    double by = uy0, bx = ux0;
    if (uy1 < by || (uy1 == by && ux1 < bx)) { off = 1; by = uy1; bx = ux1; }
    if (uy2 < by || (uy2 == by && ux2 < bx)) { off = 2; by = uy2; bx = ux2; }
    if (uy3 < by || (uy3 == by && ux3 < bx)) { off = 3; }
  }
  if (off & 1) {
    double t = ux0; ux0 = ux1; ux1 = ux2; ux2 = ux3; ux3 = t;
    t = uy0; uy0 = uy1; uy1 = uy2; uy2 = uy3; uy3 = t;
  }
  if (off & 2) {
    double t = ux0; ux0 = ux2; ux2 = t; t = ux1; ux1 = ux3; ux3 = t;
    t = uy0; uy0 = uy2; uy2 = t; t = uy1; uy1 = uy3; uy3 = t;
  }
  const int iy0 = (int)ceil(uy0), iy1 = (int)ceil(uy1), iy2 = (int)ceil(uy2), iy3 = (int)ceil(uy3);
  const double s01 = (iy1 == iy0) ? 0.0 : __ddiv_rn(__dsub_rn(ux1, ux0), __dsub_rn(uy1, uy0));
  const double s12 = (iy2 == iy1) ? 0.0 : __ddiv_rn(__dsub_rn(ux2, ux1), __dsub_rn(uy2, uy1));
  const double s03 = (iy3 == iy0) ? 0.0 : __ddiv_rn(__dsub_rn(ux3, ux0), __dsub_rn(uy3, uy0));
  const double s32 = (iy3 == iy2) ? 0.0 : __ddiv_rn(__dsub_rn(ux2, ux3), __dsub_rn(uy2, uy3));
  int total = 0, alg[5] = {0, 0, 0, 0, 0};
  const int nrows = iy2 - iy0 + 1;
  const bool byRows = nrows >= 12;
  for (int yy = byRows ? iy0 + lane : iy0; yy <= iy2; yy += byRows ? 32 : 1) {
    if (yy < 0 || yy >= sh) continue;
    const double yd = (double)yy;
    const double xa = (iy1 < yy) ? __dadd_rn(__dmul_rn(__dsub_rn(yd, uy1), s12), ux1)
                                 : __dadd_rn(__dmul_rn(__dsub_rn(yd, uy0), s01), ux0);
    const double xb = (iy3 <= yy) ? __dadd_rn(__dmul_rn(__dsub_rn(yd, uy3), s32), ux3)
                                  : __dadd_rn(__dmul_rn(__dsub_rn(yd, uy0), s03), ux0);
    int xs = (int)ceil(xa);
    int xe = (int)xb;
    if (xs < 0) xs = 0;
    if (xe > sw - 1) xe = sw - 1;
    const float* row = pix + (size_t)yy * sw;
    for (int x = byRows ? xs : xs + lane; x <= xe; x += byRows ? 1 : 32) {
      ++total;
      const double nt = lsd_ntheta(theta, __ldg(row + x));
#pragma unroll
      for (int j = 0; j < NPREC; ++j)
        if (nt <= precs[j]) ++alg[j];
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    total += __shfl_xor_sync(0xffffffffu, total, d);
#pragma unroll
    for (int j = 0; j < NPREC; ++j) alg[j] += __shfl_xor_sync(0xffffffffu, alg[j], d);
  }
  total_out = total;
#pragma unroll
  for (int j = 0; j < NPREC; ++j) alg_out[j] = alg[j];
}

// rect_improve(): within a stage the five candidate rectangles do not depend on which of them is
// accepted, so their pixel counts are gathered first and the five nfa() evaluations (the expensive
// part: log-gamma, a binomial tail) run on five lanes at once; the accept chain is then replayed in order.
// Stage 0 is the rectangle as it arrives (one "variant", always accepted), so that the scan, nfa() and the replay have
// ONE call site each and everything they touch stays in registers.
__device__ __forceinline__ void lsd_nfa_rect(const LineParams& L, const float* __restrict__ pixAll,
                                             const LsdRect* __restrict__ rectsAll, LsdSegment* __restrict__ rectOut,
                                             uint8_t* __restrict__ rectValid, int f, int ri, int lane) {
  const float* pix = pixAll + (size_t)f * L.P;
  const LsdRect* rp = rectsAll + (size_t)f * L.rect_cap + ri;
  NfaRect rec;
  rec.x1 = rp->x1; rec.y1 = rp->y1; rec.x2 = rp->x2; rec.y2 = rp->y2; rec.width = rp->width; rec.p = rp->p; rec.prec = rp->prec;
  const double theta = rp->theta, rdx = rp->dx, rdy = rp->dy;
  const double LOG_EPS = L.log_eps, LOG_NT = L.log_nt;
  const int sw = L.sw, sh = L.sh;
  const double delta = 0.5, delta_2 = delta / 2.0;
  const double shx = __dmul_rn(-rdy, delta_2), shy = __dmul_rn(rdx, delta_2);  // the half-step of stages 3 and 4
  // one step of a stage (the caller has checked the stage's width test)
  auto step = [&](NfaRect& r, int stage) {
    if (stage == 1 || stage == 5) {
      r.p = __ddiv_rn(r.p, 2.0);
      r.prec = __dmul_rn(r.p, PL_PI);
    } else {
      if (stage == 3) {
        r.x1 = __dadd_rn(r.x1, shx); r.y1 = __dadd_rn(r.y1, shy);
        r.x2 = __dadd_rn(r.x2, shx); r.y2 = __dadd_rn(r.y2, shy);
      } else if (stage == 4) {
        r.x1 = __dsub_rn(r.x1, shx); r.y1 = __dsub_rn(r.y1, shy);
        r.x2 = __dsub_rn(r.x2, shx); r.y2 = __dsub_rn(r.y2, shy);
      }
      r.width = __dsub_rn(r.width, delta);
    }
  };
  double log_nfa = 0.0;
  for (int stage = 0; stage <= 5; ++stage) {
    if (stage > 0 && log_nfa > LOG_EPS) break;
    const NfaRect rec0 = rec;  // the stage's variants derive from the rectangle it starts with
    // variant n = the stage's step applied n+1 times (each step only if its width test holds, as in rect_improve);
    // lane n keeps (myTotal, myAlg, myP, myOk) of variant n
    int myTotal = 0, myAlg = 0;
    double myP = rec0.p;
    bool myOk = false;
    unsigned okMask = 0;
    int total = 0, alg[5] = {0, 0, 0, 0, 0};
    double precs[5] = {0, 0, 0, 0, 0};
    if (stage == 1 || stage == 5) {
      if (stage == 1 || __dsub_rn(rec0.width, delta) >= 0.5) {  // the width does not change: one test for the five
        okMask = 0x1fu;
        NfaRect r = rec0;
#pragma unroll
        for (int n = 0; n < 5; ++n) {
          step(r, 1);
          precs[n] = r.prec;
          if (lane == n) { myP = r.p; myOk = true; }
        }
        rect_count_dev<5>(rec0, theta, rdx, rdy, precs, pix, sw, sh, lane, total, alg);
        myTotal = total;
#pragma unroll
        for (int n = 0; n < 5; ++n)
          if (lane == n) myAlg = alg[n];
      }
    } else {
      NfaRect r = rec0;
      const int nv = stage == 0 ? 1 : 5;
#pragma unroll 1
      for (int n = 0; n < nv; ++n) {
        if (stage > 0) {
          if (!(__dsub_rn(r.width, delta) >= 0.5)) break;
          step(r, stage);
        }
        okMask |= 1u << n;
        precs[0] = r.prec;
        rect_count_dev<1>(r, theta, rdx, rdy, precs, pix, sw, sh, lane, total, alg);
        if (lane == n) { myTotal = total; myAlg = alg[0]; myP = r.p; myOk = true; }
      }
    }
    double myNfa = 0.0;
    if (myOk) myNfa = nfa_dev(myTotal, myAlg, myP, LOG_NT);
    NfaRect r = rec0;
#pragma unroll 1
    for (int n = 0; n < 5; ++n) {
      if (!((okMask >> n) & 1u)) break;
      if (stage > 0) step(r, stage);
      const double v = __shfl_sync(0xffffffffu, myNfa, n);
      if (stage == 0 || v > log_nfa) {
        log_nfa = v;
        rec = r;
      }
    }
  }
  if (lane == 0) {
    const size_t o = (size_t)f * L.rect_cap + ri;
    rectValid[o] = log_nfa > LOG_EPS ? 1 : 0;
    LsdSegment s;
    // "+0.5" offset then "/ SCALE" (flsd())
    s.x1 = (float)__ddiv_rn(__dadd_rn(rec.x1, 0.5), L.scale);
    s.y1 = (float)__ddiv_rn(__dadd_rn(rec.y1, 0.5), L.scale);
    s.x2 = (float)__ddiv_rn(__dadd_rn(rec.x2, 0.5), L.scale);
    s.y2 = (float)__ddiv_rn(__dadd_rn(rec.y2, 0.5), L.scale);
    s.width = __ddiv_rn(rec.width, L.scale);
    s.prec = rec.p;
    s.nfa = log_nfa;
    rectOut[o] = s;
  }
}

// Persistent launch with a work queue: rectangle counts are only known on the device (~50 to several hundred per frame) and
// a rectangle that enters rect_improve costs ~25x one that is meaningful at once, so a static assignment leaves most warps
// idle while a few finish their heavy rectangles.  k_lsd_nfa_prefix lays the rectangles of the batch out as one list
// (exclusive prefix of the per-frame counts) and clears the queue head; every warp of k_lsd_nfa then takes the next
// rectangle with an atomic increment until the list is empty.  Rectangles are taken in frame order, which keeps a frame's
// angle plane in L2 while its rectangles are evaluated.
__global__ void __launch_bounds__(32) k_lsd_nfa_prefix(int batch, const int* __restrict__ nrects, int* __restrict__ prefix,
                                                       int* __restrict__ head) {
  const int lane = threadIdx.x;
  int run = 0;
  for (int base = 0; base < batch; base += 32) {
    const int v = base + lane < batch ? nrects[base + lane] : 0;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (base + lane < batch) prefix[base + lane] = run + inc - v;
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) {
    prefix[batch] = run;
    *head = 0;
  }
}

__global__ void __launch_bounds__(256, 3) k_lsd_nfa(const __grid_constant__ LineParams L, const float* __restrict__ pixAll,
                                                 const LsdRect* __restrict__ rectsAll, const int* __restrict__ prefix,
                                                 int* __restrict__ head, LsdSegment* __restrict__ rectOut,
                                                 uint8_t* __restrict__ rectValid) {
  const int lane = threadIdx.x & 31;
  const int total = __ldg(prefix + L.batch);
  while (true) {
    int item = 0;
    if (lane == 0) item = atomicAdd(head, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) break;
    int lo = 0, hi = L.batch;  // the frame of the item: largest f with prefix[f] <= item
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(prefix + mid) <= item) lo = mid; else hi = mid;
    }
    lsd_nfa_rect(L, pixAll, rectsAll, rectOut, rectValid, lo, item - __ldg(prefix + lo), lane);
  }
}

// ------------------------------------------------------------------------------------------
// k_lsd_finish: per frame — ordered compaction of the accepted rectangles, KeyLine fields
// (LSDDetector::detectImpl), keep the `max_lines` strongest by response (auxiliar.h:67-72; ties in
// detection order), LBD (BinaryDescriptor::computeLBD + binaryConversion) and the line equations
// sp x ep / |(l0, l1)| of ExtractLineSegment.
// ------------------------------------------------------------------------------------------
__constant__ int c_lbd_comb[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6},
                                      {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8},
                                      {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

__device__ __forceinline__ plslam_keyline_t make_keyline(const LsdSegment& s, int W, int H, int class_id) {
  float e0 = s.x1, e1 = s.y1, e2 = s.x2, e3 = s.y2;
  if (e0 < 0) e0 = 0;
  if (e0 >= W) e0 = (float)W - 1.0f;
  if (e2 < 0) e2 = 0;
  if (e2 >= W) e2 = (float)W - 1.0f;
  if (e1 < 0) e1 = 0;
  if (e1 >= H) e1 = (float)H - 1.0f;
  if (e3 < 0) e3 = 0;
  if (e3 >= H) e3 = (float)H - 1.0f;
  plslam_keyline_t kl;
  kl.startPointX = e0; kl.startPointY = e1; kl.endPointX = e2; kl.endPointY = e3;
  kl.sPointInOctaveX = e0; kl.sPointInOctaveY = e1; kl.ePointInOctaveX = e2; kl.ePointInOctaveY = e3;
  const double dx = (double)__fsub_rn(e0, e2), dy = (double)__fsub_rn(e1, e3);
  kl.lineLength = (float)sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
  const int x0 = __float2int_rn(e0), y0 = __float2int_rn(e1), x1 = __float2int_rn(e2), y1 = __float2int_rn(e3);
  kl.numOfPixels = max(abs(x1 - x0), abs(y1 - y0)) + 1;
  kl.angle = pl_atan2f_dev(__fsub_rn(e3, e1), __fsub_rn(e2, e0));
  kl.class_id = class_id;
  kl.octave = 0;
  kl.size = __fmul_rn(__fsub_rn(e2, e0), __fsub_rn(e3, e1));
  kl.response = __fdiv_rn(kl.lineLength, (float)max(W, H));
  kl.pt_x = __fdiv_rn(__fadd_rn(e2, e0), 2.f);
  kl.pt_y = __fdiv_rn(__fadd_rn(e3, e1), 2.f);
  return kl;
}

__device__ __forceinline__ void sobel_at_dev(const uint8_t* img, int W, int H, int pitch, int x, int y, int& dx, int& dy) {
  const int xm = reflect101_dev(x - 1, W), xp = reflect101_dev(x + 1, W);
  const int ym = reflect101_dev(y - 1, H), yp = reflect101_dev(y + 1, H);
  const uint8_t* r0 = img + (size_t)ym * pitch;
  const uint8_t* r1 = img + (size_t)y * pitch;
  const uint8_t* r2 = img + (size_t)yp * pitch;
  dx = ((int)r0[xp] - (int)r0[xm]) + 2 * ((int)r1[xp] - (int)r1[xm]) + ((int)r2[xp] - (int)r2[xm]);
  dy = ((int)r2[xm] - (int)r0[xm]) + 2 * ((int)r2[x] - (int)r0[x]) + ((int)r2[xp] - (int)r0[xp]);
}

__global__ void __launch_bounds__(256) k_lsd_finish(const __grid_constant__ LineParams L, const uint8_t* __restrict__ img,
                                                    int pitch, size_t frame_stride, const int* __restrict__ nrects,
                                                    const LsdSegment* __restrict__ rectOut,
                                                    const uint8_t* __restrict__ rectValid, LsdSegment* __restrict__ segsAll,
                                                    int* __restrict__ nsegs, float* __restrict__ respAll,
                                                    float* __restrict__ rowsumAll,
                                                    plslam_keyline_t* __restrict__ keylines, uint8_t* __restrict__ desc,
                                                    double* __restrict__ funcs, int capacity, int* __restrict__ counts,
                                                    int* __restrict__ status) {
  __shared__ int warpTmp[33];
  __shared__ int sh_nkeep, sh_run;
  __shared__ unsigned long long sh_best[8];
  __shared__ int keepIdxS[256];  // segment index of kept line r when a strongest-N selection is active (N <= 256)
  const int f = blockIdx.x, t = threadIdx.x, T = blockDim.x;
  const int nr = nrects[f];
  const LsdSegment* rin = rectOut + (size_t)f * L.rect_cap;
  const uint8_t* valid = rectValid + (size_t)f * L.rect_cap;
  LsdSegment* segs = segsAll + (size_t)f * L.rect_cap;
  float* resp = respAll + (size_t)f * L.rect_cap;
  // ordered compaction of the accepted rectangles, one 256-wide chunk at a time
  if (t == 0) sh_run = 0;
  __syncthreads();
  for (int base = 0; base < nr; base += T) {
    const int i = base + t;
    const int v = (i < nr) ? valid[i] : 0;
    // block-exclusive scan of one flag per thread
    const int lane = t & 31, w = t >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, v);
    const int inWarp = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warpTmp[w] = __popc(bal);
    __syncthreads();
    int off = sh_run;
    for (int k = 0; k < w; ++k) off += warpTmp[k];
    if (v) segs[off + inWarp] = rin[i];
    __syncthreads();
    if (t == 0) {
      int tot = 0;
      for (int k = 0; k < (T + 31) / 32; ++k) tot += warpTmp[k];
      sh_run += tot;
    }
    __syncthreads();
  }
  const int nseg = sh_run;
  if (t == 0) nsegs[f] = nseg;
  const bool select = L.max_lines > 0 && nseg > L.max_lines;
  int nkeep = nseg;
  if (select) {
    // keep the max_lines strongest by response (ties: detection order): repeated block-wide arg-max
    nkeep = L.max_lines;
    for (int i = t; i < nseg; i += T) resp[i] = make_keyline(segs[i], L.W, L.H, i).response;
    __syncthreads();
    for (int r = 0; r < nkeep; ++r) {
      unsigned long long best = 0ull;
      for (int i = t; i < nseg; i += T) {
        const float v = resp[i];
        if (v >= 0.f) {
          const unsigned long long key = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)(0x7fffffff - i);
          best = key > best ? key : best;
        }
      }
#pragma unroll
      for (int d = 16; d; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
        best = o > best ? o : best;
      }
      if ((t & 31) == 0) sh_best[t >> 5] = best;
      __syncthreads();
      if (t == 0) {
        unsigned long long b2 = 0ull;
        for (int k = 0; k < (T + 31) / 32; ++k) b2 = sh_best[k] > b2 ? sh_best[k] : b2;
        const int idx = 0x7fffffff - (int)(b2 & 0xffffffffull);
        keepIdxS[r] = idx;
        resp[idx] = -1.f;
      }
      __syncthreads();
    }
  }
  if (nkeep > L.out_cap) {  // keep-all mode on a frame with more lines than the fixed output capacity: reported, not hidden
    nkeep = L.out_cap;
    if (t == 0) atomicMax(status, PLSLAM_ERR_OVERFLOW);
  }
  if (t == 0) { sh_nkeep = nkeep; counts[f] = nkeep; }
  __syncthreads();
#define KEEP_IDX(r) (select ? keepIdxS[r] : (r))
  plslam_keyline_t* KL = keylines + (size_t)f * capacity;
  for (int r = t; r < nkeep; r += T) {
    const plslam_keyline_t kl = make_keyline(segs[KEEP_IDX(r)], L.W, L.H, r);
    KL[r] = kl;
    // lineF = sp x ep / sqrt(l0^2 + l1^2) in double
    const double x1 = kl.startPointX, y1 = kl.startPointY, x2 = kl.endPointX, y2 = kl.endPointY;
    const double l0 = __dsub_rn(y1, y2), l1 = __dsub_rn(x2, x1), l2 = __dsub_rn(__dmul_rn(x1, y2), __dmul_rn(y1, x2));
    const double nrm = sqrt(__dadd_rn(__dmul_rn(l0, l0), __dmul_rn(l1, l1)));
    double* F = funcs + ((size_t)f * capacity + r) * 3;
    F[0] = __ddiv_rn(l0, nrm);
    F[1] = __ddiv_rn(l1, nrm);
    F[2] = __ddiv_rn(l2, nrm);
  }
}

// ------------------------------------------------------------------------------------------
// k_lbd: BinaryDescriptor::computeLBD + binaryConversion, one CTA per (frame, kept line).
// Thread = one of the 63 rows of the line support region (sequential float sums along the line, as the
// reference); then thread = one of the 72 (band, statistic) accumulators, each summing its <= 21 row
// contributions in row order; thread 0 normalises, clamps and binarises.
// ------------------------------------------------------------------------------------------
// Sobel taps of one position straight from the image (safety net of k_lbd when a chunk's box does not fit its tile)
__device__ __noinline__ void lbd_taps_global(const uint8_t* I, int pitch, int x, int y, int xm, int xp, int ym, int yp, int* dx,
                                             int* dy) {
  const uint8_t* r0 = I + (size_t)ym * pitch;
  const uint8_t* r1 = I + (size_t)y * pitch;
  const uint8_t* r2 = I + (size_t)yp * pitch;
  const int a = r0[xm], b = r0[x], c = r0[xp], d = r1[xm], e = r1[xp], g = r2[xm], h = r2[x], i = r2[xp];
  *dx = (c - a) + 2 * (e - d) + (i - g);
  *dy = (g - a) + 2 * (h - b) + (i - c);
}

constexpr int LBD_THREADS = 64;
__global__ void __launch_bounds__(LBD_THREADS) k_lbd(const __grid_constant__ LineParams L, const uint8_t* __restrict__ img, int pitch,
                                            size_t frame_stride, const plslam_keyline_t* __restrict__ keylines,
                                            const int* __restrict__ counts, uint8_t* __restrict__ desc, int capacity) {
  __shared__ float rowsum[LBD_ROWS][4];
  __shared__ float band[LBD_BANDS][8];
  const int f = blockIdx.y, r = blockIdx.x, t = threadIdx.x;
  if (r >= counts[f]) return;
  const uint8_t* I = img + (size_t)f * frame_stride;
  const plslam_keyline_t kl = keylines[(size_t)f * capacity + r];
  // Thread hID < 63 owns row hID of the line support region.  The walk along a row is a chain of float additions (the
  // positions and the four sums), so it stays sequential per row; the image is what the rows share.  The line is cut into
  // chunks of LBD_CH steps: every row first advances its positions through the chunk, rows 0 and 62 publish the bounding
  // box (positions are monotone in the row index, so the two outer rows bound all of them), the CTA copies that box of
  // the image (+1 pixel for the Sobel taps, at most 76 x 76 bytes for 32 steps x 63 rows at any angle) into shared memory
  // with coalesced row loads, and the rows then take their Sobel taps from the tile.  BORDER_REFLECT_101 is folded into
  // the tap indices (x - 1 -> 1 at x = 0, x + 1 -> W - 2 at x = W - 1), which stay inside the box.
  constexpr int LBD_CH = 32, LBD_TILE = 80, LBD_TROWS = 76;  // tile pitch 80: the box is widened to whole 4-byte words
  __shared__ __align__(16) uint8_t tile[LBD_TILE * LBD_TROWS];
  // the chunk's positions, [step][row] (row-major over steps keeps a row's thread on its own bank column); in shared memory rather
  // than registers so that the two loops over the chunk stay rolled: the unrolled form was 6.7 k instructions (107 KB of code)
  __shared__ unsigned posS[LBD_CH][LBD_ROWS + 1];
  const bool wordLoads = ((reinterpret_cast<uintptr_t>(I) | (uintptr_t)pitch) & 3u) == 0;
  __shared__ int boxS[2][4];
  const bool rowThread = t < LBD_ROWS;
  const int hID = t;
  const int lengthOfLSP = (int)(short)kl.numOfPixels;
  const int W = L.W, H = L.H;
  float dL0, dL1;
  {
    double sn_, cs_;
    pl_sincos_dev((double)kl.angle, &sn_, &cs_);
    dL0 = (float)cs_;
    dL1 = (float)sn_;
  }
  const float dO0 = -dL1, dO1 = dL0;
  float sCorX = 0.f, sCorY = 0.f;
  if (rowThread) {
    const short halfWidth = (short)((lengthOfLSP - 1) / 2), halfHeight = (LBD_ROWS - 1) / 2;
    const float midX = (float)__dmul_rn(0.5, (double)__fadd_rn(kl.sPointInOctaveX, kl.ePointInOctaveX));
    const float midY = (float)__dmul_rn(0.5, (double)__fadd_rn(kl.sPointInOctaveY, kl.ePointInOctaveY));
    // sCor0 after hID row steps (each step: sCorX0 -= dL[1]; sCorY0 += dL[0])
    sCorX = __fadd_rn(__fadd_rn(__fmul_rn(-dL0, (float)halfWidth), __fmul_rn(dL1, (float)halfHeight)), midX);
    sCorY = __fadd_rn(__fsub_rn(__fmul_rn(-dL1, (float)halfWidth), __fmul_rn(dL0, (float)halfHeight)), midY);
    for (int k = 0; k < hID; ++k) {
      sCorX = __fsub_rn(sCorX, dL1);
      sCorY = __fadd_rn(sCorY, dL0);
    }
  }
  float pgdL = 0, ngdL = 0, pgdO = 0, ngdO = 0;
  const int xlo = W > 1 ? 1 : 0, xhi = W > 1 ? W - 2 : 0, ylo = H > 1 ? 1 : 0, yhi = H > 1 ? H - 2 : 0;
  for (int w0 = 0; w0 < lengthOfLSP; w0 += LBD_CH) {
    const int nst = min(LBD_CH, lengthOfLSP - w0);
    int mnx = 0x7fff, mxx = 0, mny = 0x7fff, mxy = 0;
    if (rowThread) {
#pragma unroll 1
      for (int u = 0; u < nst; ++u) {
        short tc = (short)roundf(sCorX);
        const int x = (tc < 0) ? 0 : (tc > W - 1) ? W - 1 : tc;
        tc = (short)roundf(sCorY);
        const int y = (tc < 0) ? 0 : (tc > H - 1) ? H - 1 : tc;
        posS[u][hID] = ((unsigned)y << 16) | (unsigned)x;
        sCorX = __fadd_rn(sCorX, dL0);
        sCorY = __fadd_rn(sCorY, dL1);
        mnx = min(mnx, x); mxx = max(mxx, x);
        mny = min(mny, y); mxy = max(mxy, y);
      }
      if (hID == 0 || hID == LBD_ROWS - 1) {
        int* bs = boxS[hID ? 1 : 0];
        bs[0] = mnx; bs[1] = mxx; bs[2] = mny; bs[3] = mxy;
      }
    }
    __syncthreads();
    int bx0 = max(0, min(boxS[0][0], boxS[1][0]) - 1);
    const int bx1 = min(W - 1, max(boxS[0][1], boxS[1][1]) + 1);
    const int by0 = max(0, min(boxS[0][2], boxS[1][2]) - 1), by1 = min(H - 1, max(boxS[0][3], boxS[1][3]) + 1);
    if (wordLoads) bx0 &= ~3;
    const int tw = bx1 - bx0 + 1, th = by1 - by0 + 1;
    const bool tiled = tw <= LBD_TILE && th <= LBD_TROWS;  // always, by the bound above; the global path is the safety net
    if (tiled) {
      if (wordLoads) {
        // a row of the box is at most 20 words: two rows per warp step when it is at most 16
        const int twords = (tw + 3) >> 2, lane = t & 31, warp = t >> 5;
        const int rpw = twords <= 16 ? 2 : 1, col = rpw == 2 ? (lane & 15) : lane, sub = rpw == 2 ? (lane >> 4) : 0;
        for (int ty = warp * rpw + sub; ty < th; ty += (LBD_THREADS / 32) * rpw)
          if (col < twords)
            reinterpret_cast<unsigned*>(tile + ty * LBD_TILE)[col] =
                __ldg(reinterpret_cast<const unsigned*>(I + (size_t)(by0 + ty) * pitch + bx0) + col);
      } else {
        for (int ty = t >> 5; ty < th; ty += LBD_THREADS / 32) {
          const uint8_t* src = I + (size_t)(by0 + ty) * pitch + bx0;
          for (int tx = t & 31; tx < tw; tx += 32) tile[ty * LBD_TILE + tx] = __ldg(src + tx);
        }
      }
    }
    __syncthreads();
    if (rowThread) {
#pragma unroll 2
      for (int u = 0; u < nst; ++u) {
        {
          const unsigned pu = posS[u][hID];
          const int x = (int)(pu & 0xffffu), y = (int)(pu >> 16);
          const int xm = x > 0 ? x - 1 : xlo, xp = x < W - 1 ? x + 1 : xhi;
          const int ym = y > 0 ? y - 1 : ylo, yp = y < H - 1 ? y + 1 : yhi;
          int dx, dy;
          if (tiled) {
            const uint8_t* r0 = tile + (ym - by0) * LBD_TILE - bx0;
            const uint8_t* r1 = tile + (y - by0) * LBD_TILE - bx0;
            const uint8_t* r2 = tile + (yp - by0) * LBD_TILE - bx0;
            const int a = r0[xm], b = r0[x], c = r0[xp], d = r1[xm], e = r1[xp], g = r2[xm], h = r2[x], i = r2[xp];
            dx = (c - a) + 2 * (e - d) + (i - g);
            dy = (g - a) + 2 * (h - b) + (i - c);
          } else {
            lbd_taps_global(I, pitch, x, y, xm, xp, ym, yp, &dx, &dy);
          }
          const float gDL = __fadd_rn(__fmul_rn((float)dx, dL0), __fmul_rn((float)dy, dL1));
          const float gDO = __fadd_rn(__fmul_rn((float)dx, dO0), __fmul_rn((float)dy, dO1));
          if (gDL > 0) pgdL = __fadd_rn(pgdL, gDL); else ngdL = __fsub_rn(ngdL, gDL);
          if (gDO > 0) pgdO = __fadd_rn(pgdO, gDO); else ngdO = __fsub_rn(ngdO, gDO);
        }
      }
    }
  }
  if (rowThread) {
    const float cg = L.gaussG[hID];
    rowsum[hID][0] = __fmul_rn(cg, pgdL);
    rowsum[hID][1] = __fmul_rn(cg, ngdL);
    rowsum[hID][2] = __fmul_rn(cg, pgdO);
    rowsum[hID][3] = __fmul_rn(cg, ngdO);
  }
  __syncthreads();
  for (int tt = t; tt < LBD_BANDS * 8; tt += LBD_THREADS) {
    // accumulator (band b, statistic q): q = 0 pgdL, 1 ngdL, 2 pgdL^2, 3 ngdL^2, 4 pgdO, 5 ngdO, 6 pgdO^2, 7 ngdO^2
    const int b = tt >> 3, q = tt & 7;
    const int src = (q & 1) + ((q & 4) ? 2 : 0);  // which of the 4 row sums
    const bool sq = (q & 2) != 0;
    float acc = 0.f;
    const int h0 = max(0, (b - 1) * LBD_W), h1 = min(LBD_ROWS, (b + 2) * LBD_W);
    for (int hID = h0; hID < h1; ++hID) {
      const int b0 = hID / LBD_W, m = hID % LBD_W;
      // row hID adds to its own band with gaussL[m+7], to the band above (b0-1) with gaussL[m+14], below (b0+1) with gaussL[m]
      const float coef = (b == b0) ? L.gaussL[m + LBD_W] : (b == b0 - 1 ? L.gaussL[m + 2 * LBD_W] : L.gaussL[m]);
      const float v = rowsum[hID][src];
      acc = sq ? __fadd_rn(acc, __fmul_rn(__fmul_rn(coef, coef), __fmul_rn(v, v))) : __fadd_rn(acc, __fmul_rn(coef, v));
    }
    band[b][q] = acc;
  }
  __syncthreads();
  if (t == 0) {
    float des[LBD_BANDS * 8];
    const float invN2 = (float)(1.0 / (LBD_W * 2.0)), invN3 = (float)(1.0 / (LBD_W * 3.0));
    for (int b = 0; b < LBD_BANDS; ++b) {
      const float invN = (b == 0 || b == LBD_BANDS - 1) ? invN2 : invN3;
      float tmp = __fmul_rn(band[b][0], invN);
      des[b * 8] = tmp;
      des[b * 8 + 4] = sqrtf(__fsub_rn(__fmul_rn(band[b][2], invN), __fmul_rn(tmp, tmp)));
      tmp = __fmul_rn(band[b][1], invN);
      des[b * 8 + 1] = tmp;
      des[b * 8 + 5] = sqrtf(__fsub_rn(__fmul_rn(band[b][3], invN), __fmul_rn(tmp, tmp)));
      tmp = __fmul_rn(band[b][4], invN);
      des[b * 8 + 2] = tmp;
      des[b * 8 + 6] = sqrtf(__fsub_rn(__fmul_rn(band[b][6], invN), __fmul_rn(tmp, tmp)));
      tmp = __fmul_rn(band[b][5], invN);
      des[b * 8 + 3] = tmp;
      des[b * 8 + 7] = sqrtf(__fsub_rn(__fmul_rn(band[b][7], invN), __fmul_rn(tmp, tmp)));
    }
    float tempM = 0, tempS = 0;
    for (int b = 0; b < LBD_BANDS; ++b) {
      for (int q = 0; q < 4; ++q) tempM = __fadd_rn(tempM, __fmul_rn(des[b * 8 + q], des[b * 8 + q]));
      for (int q = 4; q < 8; ++q) tempS = __fadd_rn(tempS, __fmul_rn(des[b * 8 + q], des[b * 8 + q]));
    }
    tempM = __fdiv_rn(1.f, sqrtf(tempM));
    tempS = __fdiv_rn(1.f, sqrtf(tempS));
    for (int b = 0; b < LBD_BANDS; ++b) {
      for (int q = 0; q < 4; ++q) des[b * 8 + q] = __fmul_rn(des[b * 8 + q], tempM);
      for (int q = 4; q < 8; ++q) des[b * 8 + q] = __fmul_rn(des[b * 8 + q], tempS);
    }
    for (int i = 0; i < LBD_BANDS * 8; ++i)
      if ((double)des[i] > 0.4) des[i] = (float)0.4;
    float tmp = 0;
    for (int i = 0; i < LBD_BANDS * 8; ++i) tmp = __fadd_rn(tmp, __fmul_rn(des[i], des[i]));
    tmp = __fdiv_rn(1.f, sqrtf(tmp));
    for (int i = 0; i < LBD_BANDS * 8; ++i) des[i] = __fmul_rn(des[i], tmp);
    uint8_t* D = desc + ((size_t)f * capacity + r) * 32;
    for (int comb = 0; comb < 32; ++comb) {
      const int a = c_lbd_comb[comb][0] * 8, b = c_lbd_comb[comb][1] * 8;
      unsigned res = 0;
      for (int i = 0; i < 8; ++i)
        if (des[a + i] > des[b + i]) res += 1u << i;
      D[comb] = (uint8_t)res;
    }
  }
}

}  // namespace
int debug_aw_stats(unsigned long long* out32) {
  if (cudaDeviceSynchronize() != cudaSuccess) return PLSLAM_ERR_CUDA;
  if (cudaMemcpyFromSymbol(out32, g_aw_stats, sizeof(g_aw_stats)) != cudaSuccess) return PLSLAM_ERR_CUDA;
  unsigned long long z[32] = {0};
  if (cudaMemcpyToSymbol(g_aw_stats, z, sizeof(z)) != cudaSuccess) return PLSLAM_ERR_CUDA;
  return PLSLAM_OK;
}
#ifdef PLSLAM_GROW_PROF
int debug_grow_prof(unsigned long long* out16) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, g_grow_prof, sizeof(g_grow_prof));
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(g_grow_prof, z, sizeof(z));
  return 0;
}
#endif

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
LineExtractor::LineExtractor() {}

LineExtractor::~LineExtractor() {
  DevBuf* all[] = {&listpool, &rectstage, &recttmp, &owner, &degp, &g2p, &ubm, &nfaq, &scaled, &pix, &coef, &rowhist, &binstart, &maxg2, &seeds, &nseeds, &regbuf, &rects, &nrects,
                   &rectout, &segs, &nsegs, &resp, &rowsum, &status, &stageIn, &stageKl, &stageDesc, &stageFuncs, &stageCnt};
  for (DevBuf* b : all) b->release();
  if (ownStream) cudaStreamDestroy(ownStream);
  if (pinnedStatus) cudaFreeHost(pinnedStatus);
}

// OpenCV's 8-bit fixed-point Gaussian table (error diffusion from the borders, centre = 256 - 2*side)
static void gauss_table_u8(double sigma, int ksize, int* out) {
  std::vector<double> k(ksize);
  double sum = 0;
  const double scale2X = -0.5 / (sigma * sigma);
  for (int i = 0; i < ksize; ++i) {
    const double x = i - (ksize - 1) * 0.5;
    k[i] = std::exp(scale2X * x * x);
    sum += k[i];
  }
  double err = 0;
  int side = 0;
  for (int i = 0; i < ksize / 2; ++i) {
    const double v = k[i] / sum * 256.0 + err;
    const int r = (int)std::lrint(v);
    err = v - r;
    out[i] = out[ksize - 1 - i] = r;
    side += r;
  }
  out[ksize / 2] = 256 - 2 * side;
}

int LineExtractor::configure(int W, int H, int batch) {
  if (W == cfgW && H == cfgH && batch <= cfgB) return PLSLAM_OK;
  PL_CHECK_ARG(W >= 16 && H >= 16 && W <= 16000 && H <= 16000);
  if (device < 0) {
    PL_CUDA(cudaGetDevice(&device));
    PL_CUDA(cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, device));
    PL_CUDA(cudaStreamCreateWithFlags(&ownStream, cudaStreamNonBlocking));
    PL_CUDA(cudaMallocHost(&pinnedStatus, 64));
  }
  std::memset(&P, 0, sizeof(P));
  // createLineSegmentDetector(LSD_REFINE_ADV) defaults: scale 0.8, sigma_scale 0.6, quant 2, ang_th 22.5,
  // log_eps 0, density_th 0.7, n_bins 1024
  const double SCALE = 0.8, SIGMA_SCALE = 0.6, QUANT = 2.0, ANG_TH = 22.5;
  P.W = W;
  P.H = H;
  P.scale = SCALE;
  P.sw = (int)std::lrint(W * SCALE);
  P.sh = (int)std::lrint(H * SCALE);
  P.spitch = (int)align_up(P.sw, 32);
  P.P = P.sw * P.sh;
  const double sigma = SIGMA_SCALE / SCALE;
  const unsigned h = (unsigned)std::ceil(sigma * std::sqrt(2 * 3.0 * std::log(10.0)));
  P.ksize = 1 + 2 * (int)h;
  PL_CHECK_ARG(P.ksize == 7);
  gauss_table_u8(sigma, P.ksize, P.blurk);
  P.prec = PL_PI * ANG_TH / 180;
  P.p = ANG_TH / 180;
  P.rho = QUANT / std::sin(P.prec);
  P.log_nt = 5 * (std::log10((double)P.sw) + std::log10((double)P.sh)) / 2 + std::log10(11.0);
  P.min_reg_size = (int)(size_t)(-P.log_nt / std::log10(P.p));
  {  // smallest gx^2 + gy^2 with modgrad = sqrt(g2 * 0.25) > rho (same double operations as ll_angle(); monotone in g2)
    int g = 0;
    while (!(std::sqrt((double)g * 0.25) > P.rho)) ++g;
    P.g2_min = g;
  }
  P.density_th = 0.7;
  P.log_eps = 0.0;
  P.max_lines = max_lines;
  {  // experiment switch (default 3): bit 0 = in-batch speculation, bit 1 = L1 prefetch
    const char* gv = std::getenv("PLSLAM_GROW_VARIANT");
    P.grow_variant = gv ? std::atoi(gv) : 3;  // bit 0: in-batch speculation, bit 1: L1 prefetch of new region points' neighbourhoods
  }
  // accepted rectangles per frame: ~1 per 300 scaled pixels on the synthetic frames; capacity 1 per 32, overflow is reported
  rect_cap = std::min(std::max(4096, P.P / 32), 1 << 17);
  P.rect_cap = rect_cap;
  PL_CHECK_ARG(max_lines <= 256);
  P.out_cap = out_capacity();
  {  // LBD weights (BinaryDescriptor constructor)
    double u = (LBD_W * 3 - 1) / 2;
    double sg = (LBD_W * 2 + 1) / 2;
    double inv = -1 / (2 * sg * sg);
    for (int i = 0; i < LBD_W * 3; ++i) { const double d = i - u; P.gaussL[i] = (float)std::exp(d * d * inv); }
    u = (LBD_BANDS * LBD_W - 1) / 2;
    sg = u;
    inv = -1 / (2 * sg * sg);
    for (int i = 0; i < LBD_ROWS; ++i) { const double d = i - u; P.gaussG[i] = (float)std::exp(d * d * inv); }
  }
  // INTER_LINEAR_EXACT tables (step exactly 1/SCALE)
  std::vector<int> tab;
  auto emit = [&](int ssize, int dsize) {
    std::vector<int> ofs(dsize), c1(dsize);
    const double sc = 1.0 / SCALE;
    for (int d = 0; d < dsize; ++d) {
      const double fv = sc * (d + 0.5) - 0.5;
      const int i = (int)std::floor(fv);
      if (i >= 0 && ssize > 1) {
        if (i < ssize - 1) { ofs[d] = i; c1[d] = (int)std::lrint((fv - i) * 256.0); }
        else { ofs[d] = ssize - 1; c1[d] = 0; }
      } else { ofs[d] = 0; c1[d] = 0; }
    }
    tab.insert(tab.end(), ofs.begin(), ofs.end());
    tab.insert(tab.end(), c1.begin(), c1.end());
  };
  emit(W, P.sw);
  emit(H, P.sh);
  int rc;
  const size_t B = std::max(batch, cfgB);
  if ((rc = coef.ensure(tab.size() * sizeof(int)))) return rc;
  PL_CUDA(cudaMemcpy(coef.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
  if ((rc = scaled.ensure(B * P.spitch * P.sh))) return rc;
  if ((rc = pix.ensure(B * P.P * sizeof(uint4)))) return rc;
  if ((rc = degp.ensure(B * P.P * sizeof(float)))) return rc;
  if ((rc = g2p.ensure(B * P.P * sizeof(unsigned)))) return rc;
  if ((rc = rowhist.ensure(B * (size_t)div_up(P.sh, SORT_ROWS) * LSD_BINS * sizeof(unsigned)))) return rc;
  if ((rc = binstart.ensure(B * LSD_BINS * sizeof(unsigned)))) return rc;
  if ((rc = maxg2.ensure(B * sizeof(int)))) return rc;
  if ((rc = seeds.ensure(B * P.P * sizeof(unsigned)))) return rc;
  if ((rc = nseeds.ensure(B * sizeof(int)))) return rc;
  if ((rc = regbuf.ensure(B * P.P * sizeof(unsigned)))) return rc;
  if ((rc = rects.ensure(B * P.rect_cap * sizeof(LsdRect)))) return rc;
  if ((rc = nrects.ensure(B * sizeof(int)))) return rc;
  if ((rc = nfaq.ensure((B + 2) * sizeof(int)))) return rc;  // k_lsd_nfa's work list: prefix[B + 1] and the queue head
  if ((rc = rectout.ensure(B * P.rect_cap * (sizeof(LsdSegment) + 1)))) return rc;
  if ((rc = segs.ensure(B * P.rect_cap * sizeof(LsdSegment)))) return rc;
  if ((rc = nsegs.ensure(B * sizeof(int)))) return rc;
  if ((rc = resp.ensure(B * P.rect_cap * sizeof(float)))) return rc;
  if ((rc = rowsum.ensure(B * P.out_cap * LBD_ROWS * 4 * sizeof(float)))) return rc;
  if ((rc = status.ensure(sizeof(int)))) return rc;
  if (!statusArmed) {  // the status word is sticky (kernels only raise it): cleared when it is created, never on a reconfigure
    PL_CUDA(cudaMemset(status.p, 0, sizeof(int)));
    statusArmed = true;
  }
  PL_CHECK_ARG(P.sw < 65536 && P.sh < 32768);
  {  // the log-gamma table is a __device__ global shared by every extractor of the process: built once per device
    static PerDeviceOnce tab;
    if (tab.first()) {
      PL_CARVEOUT(k_lsd_lgamma_table);
      k_lsd_lgamma_table<<<div_up(LGAMMA_TABLE, 256), 256>>>();
      PL_CUDA(cudaDeviceSynchronize());
    }
  }
  cfgW = W;
  cfgH = H;
  cfgB = (int)B;
  return PLSLAM_OK;
}

int LineExtractor::extract_device(const uint8_t* d_images, int batch, int W, int H, int pitch, size_t frame_stride,
                                  plslam_keyline_t* d_keylines, uint8_t* d_desc, double* d_funcs, int capacity,
                                  int32_t* d_counts, cudaStream_t st) {
  PL_CHECK_ARG(d_images && d_keylines && d_desc && d_funcs && d_counts);
  PL_CHECK_ARG(batch >= 1 && batch <= 65535 && pitch >= W);
  int rc = configure(W, H, batch);
  if (rc) return rc;
  if (capacity < P.out_cap) {
    set_error("capacity %d < line output capacity %d", capacity, P.out_cap);
    return PLSLAM_ERR_CAPACITY;
  }
  last_batch = batch;
  PL_CUDA(cudaMemsetAsync(maxg2.p, 0, (size_t)batch * sizeof(int), st));
  // `status` is sticky: kernels only raise it, check_status() reads and re-arms it (several batches may be in flight)
  PL_STAGE_BEGIN(timer, "lsd_scale", st);
  PL_CARVEOUT(k_lsd_scale);
  k_lsd_scale<<<dim3(div_up(P.sw, ST_W), div_up(P.sh, ST_H), batch), 256, 0, st>>>(P, d_images, pitch, frame_stride,
                                                                                     coef.as<int>(), scaled.as<uint8_t>());
  PL_STAGE_END(timer, st);
  PL_STAGE_BEGIN(timer, "lsd_grad", st);
  PL_CARVEOUT(k_lsd_grad);
  // Where the one-warp-per-frame kernel keeps the `used` map (experiment switch PLSLAM_GROW_USED: 0 = in the pixel records
  // (default), 1 = bitmap in shared memory, 2 = bitmap in global memory).  Measured on B200, 640x480 (tools/r02_bm_ab.sh,
  // gpurun_out/r02_bm_ab.log): the bitmaps cut the kernel's loads to the records of available pixels only, yet one batch of
  // 256 frames takes 47.4 ms (shared) / 57.2 ms (global) against 49.5 ms, and with 16 batches in flight the shared bitmap
  // (24 KB per frame: 4 frames per SM instead of ~20) falls from 17.8 k to 11.7 k frames/s, the global one to 17.0 k: the
  // kernel is bound by its dependent instruction chain, not by the records it loads.
  static const int usedEnv = [] { const char* e = std::getenv("PLSLAM_GROW_USED"); return e ? std::atoi(e) : -1; }();
  static const int gwEnv = [] { const char* e = std::getenv("PLSLAM_GROW_GW"); return e ? std::atoi(e) : 0; }();
  const size_t bmStride = (size_t)((P.P + 31) / 32 + 3) / 4 * 4;  // words per frame
  int usedMode = usedEnv >= 0 && usedEnv <= GM_BMG ? usedEnv : GM_REC;
  if (usedMode == GM_BMS && grow_smem_per_frame(P.P, true) > 200 * 1024) usedMode = GM_BMG;
  if (usedMode != GM_REC) {
    int rcb;
    if ((rcb = ubm.ensure((size_t)std::max(batch, cfgB) * bmStride * sizeof(unsigned)))) return rcb;
    if (P.sw % 32) PL_CUDA(cudaMemsetAsync(ubm.p, 0, (size_t)batch * bmStride * sizeof(unsigned), st));
  }
  k_lsd_grad<<<dim3(div_up(P.sw, GRAD_TW), div_up(P.sh, GRAD_TH), batch), 256, 0, st>>>(P, scaled.as<uint8_t>(), pix.as<uint4>(), degp.as<float>(), g2p.as<unsigned>(),
                                                                             maxg2.as<int>(), usedMode != GM_REC ? ubm.as<unsigned>() : nullptr);
  PL_STAGE_END(timer, st);
  PL_STAGE_BEGIN(timer, "lsd_rowhist", st);
  PL_CARVEOUT(k_lsd_rowhist);
  k_lsd_rowhist<<<dim3(div_up(P.sh, SORT_ROWS), batch), 256, 0, st>>>(P, g2p.as<unsigned>(), maxg2.as<int>(), rowhist.as<unsigned>());
  PL_STAGE_END(timer, st);
  PL_STAGE_BEGIN(timer, "lsd_colscan", st);
  PL_CARVEOUT(k_lsd_colscan);
  k_lsd_colscan<<<batch, LSD_BINS, 0, st>>>(P, div_up(P.sh, SORT_ROWS), rowhist.as<unsigned>(), binstart.as<unsigned>(), nseeds.as<int>());
  PL_STAGE_END(timer, st);
  PL_STAGE_BEGIN(timer, "lsd_scatter", st);
  PL_CARVEOUT(k_lsd_scatter);
  {
    static PerDeviceOnce attr;
    if (attr.first()) PL_CUDA(cudaFuncSetAttribute(k_lsd_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  PL_CHECK_ARG((size_t)SORT_ROWS * P.sw * sizeof(unsigned) <= 160 * 1024);
  k_lsd_scatter<<<dim3(div_up(P.sh, SORT_ROWS), batch), 256, (size_t)SORT_ROWS * P.sw * sizeof(unsigned), st>>>(P, g2p.as<unsigned>(), maxg2.as<int>(), rowhist.as<unsigned>(),
                                                              binstart.as<unsigned>(), seeds.as<unsigned>());
  PL_STAGE_END(timer, st);
  if (mark_event && mark_where == 1) PL_CUDA(cudaEventRecord(mark_event, st));
  PL_STAGE_BEGIN(timer, "lsd_grow", st);
  P.batch = batch;
  // PLSLAM_GROW_MODE: 0 = one warp per frame (k_lsd_grow), 2 = speculative multi-warp growing with in-order retirement
  // (k_lsd_grow_sw); unset = automatic.  PLSLAM_SW_K = warps per frame of mode 2 (8, 16 or 32).
  static const int growMode = [] { const char* e = std::getenv("PLSLAM_GROW_MODE"); return e ? std::atoi(e) : -1; }();  // -1 auto
  static const int swKenv = [] { const char* e = std::getenv("PLSLAM_SW_K"); return e ? std::atoi(e) : 0; }();
  int swK = swKenv ? swKenv : 8;  // (16 and 32 workers do not pay: the retirement order, not the worker count, limits)
  swK = swK >= 32 ? 32 : swK >= 16 ? 16 : 8;
  // per worker warp a list area of 2 P entries, per frame a list pool of 2 P entries: mode 2 needs them within 24 GB
  const size_t swBytes = (size_t)std::max(batch, cfgB) * (2 * (size_t)(swK - 1) + 2) * P.P * sizeof(unsigned);
  // Measured on B200 (tools/prof_aw.py, one batch on the GPU, 640x480): 29 -> 11 ms for 1 frame, 45 -> 26 ms for 64,
  // 50 -> 34 ms for 148, 55 -> 49 ms for 256; 186 -> 63 ms for 8 frames of 1280x720.  With many batches in flight the
  // one-warp-per-frame kernel fills the machine with less work per region, so mode 2 is chosen while the GPU holds at most
  // two frames per SM.
  const bool sw = growMode >= 0 ? growMode != 0
                                : ((long long)batch * batches_in_flight <= 2 * numSMs && swBytes <= ((size_t)24 << 30));
  if (sw) {
    int rc2;
    static bool swTune = false;
    if (!swTune) {  // experiment knobs of the same-edge heuristic (never affect results)
      swTune = true;
      if (const char* e = std::getenv("PLSLAM_SW_PERP")) { float v = (float)std::atof(e); cudaMemcpyToSymbol(g_sw_perp, &v, sizeof(v)); }
      if (const char* e = std::getenv("PLSLAM_SW_ALONG")) { float v = (float)std::atof(e); cudaMemcpyToSymbol(g_sw_along, &v, sizeof(v)); }
      if (const char* e = std::getenv("PLSLAM_SW_ANG")) { float v = (float)std::atof(e); cudaMemcpyToSymbol(g_sw_ang, &v, sizeof(v)); }
    }
    if ((rc2 = owner.ensure((size_t)cfgB * P.P * sizeof(unsigned)))) return rc2;
    if ((rc2 = regbuf.ensure((size_t)cfgB * (swK - 1) * 2 * P.P * sizeof(unsigned)))) return rc2;
    if ((rc2 = listpool.ensure((size_t)cfgB * 2 * P.P * sizeof(unsigned)))) return rc2;
    if ((rc2 = recttmp.ensure((size_t)cfgB * P.rect_cap * sizeof(LsdRect)))) return rc2;
    PL_CUDA(cudaMemsetAsync(owner.p, 0xff, (size_t)batch * P.P * sizeof(unsigned), st));
#define PL_SW_LAUNCH(KK, WW)                                                                                                  \
  do {                                                                                                                        \
    static PerDeviceOnce attr;                                                                                                \
    if (attr.first()) {                                                                                                       \
      PL_CUDA(cudaFuncSetAttribute(k_lsd_grow_sw<KK, WW>, cudaFuncAttributeMaxDynamicSharedMemorySize,                        \
                                   (int)sw_smem_bytes<KK, WW>()));                                                            \
    }                                                                                                                         \
    k_lsd_grow_sw<KK, WW><<<batch, 32 * KK, sw_smem_bytes<KK, WW>(), st>>>(                                                   \
        P, pix.as<uint4>(), owner.as<unsigned>(), seeds.as<unsigned>(), nseeds.as<int>(), regbuf.as<unsigned>(),              \
        listpool.as<unsigned>(), recttmp.as<LsdRect>(), rects.as<LsdRect>(), nrects.as<int>(), status.as<int>());             \
  } while (0)
    if (swK == 32) PL_SW_LAUNCH(32, 1024);
    else if (swK == 16) PL_SW_LAUNCH(16, 1024);
    else PL_SW_LAUNCH(8, 2048);
#undef PL_SW_LAUNCH
  } else {
    // frames per CTA: 4 unless the bitmaps in shared memory would push a CTA beyond a quarter of the SM
    const size_t perFrame = grow_smem_per_frame(P.P, usedMode == GM_BMS);
    int gw = gwEnv ? std::min(std::max(gwEnv, 1), GROW_WARPS) : GROW_WARPS;
    while (gw > 1 && perFrame * gw > 60 * 1024) gw >>= 1;
    // experiment switch: extra dynamic shared memory per CTA, to cap how many region-growing CTAs an SM holds (its warps take
    // 72 registers each: 7 CTAs of 4 frames fill the register file and leave no room for the other kernels of the pipeline)
    static const size_t growPad = [] { const char* e = std::getenv("PLSLAM_GROW_PAD"); return e ? (size_t)std::atoi(e) : (size_t)0; }();
#define PL_GROW_LAUNCH(MODE)                                                                                                  \
  do {                                                                                                                        \
    static PerDeviceOnce attr;                                                                                                \
    if (attr.first())                                                                                                         \
      PL_CUDA(cudaFuncSetAttribute(k_lsd_grow<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));               \
    PL_CARVEOUT(k_lsd_grow<MODE>);                                                                                            \
    k_lsd_grow<MODE><<<div_up(batch, gw), 32 * gw, perFrame * gw + growPad, st>>>(P, pix.as<uint4>(), ubm.as<unsigned>(),               \
                                                                        seeds.as<unsigned>(), nseeds.as<int>(),               \
                                                                        regbuf.as<unsigned>(), rects.as<LsdRect>(),           \
                                                                        nrects.as<int>(), status.as<int>());                  \
  } while (0)
    if (usedMode == GM_BMS) PL_GROW_LAUNCH(GM_BMS);
    else if (usedMode == GM_BMG) PL_GROW_LAUNCH(GM_BMG);
    else PL_GROW_LAUNCH(GM_REC);
#undef PL_GROW_LAUNCH
  }
  PL_STAGE_END(timer, st);
  if (mark_event && mark_where == 2) PL_CUDA(cudaEventRecord(mark_event, st));
  static const bool grow_only = std::getenv("PLSLAM_DEBUG_STOP_AFTER_GROW") != nullptr;  // profiling aid (tools/)
  if (grow_only) return PLSLAM_OK;
  LsdSegment* rout = rectout.as<LsdSegment>();
  uint8_t* rvalid = reinterpret_cast<uint8_t*>(rout + (size_t)cfgB * P.rect_cap);
  PL_STAGE_BEGIN(timer, "lsd_nfa", st);
  PL_CARVEOUT(k_lsd_nfa);
  k_lsd_nfa_prefix<<<1, 32, 0, st>>>(batch, nrects.as<int>(), nfaq.as<int>(), nfaq.as<int>() + batch + 1);
  k_lsd_nfa<<<numSMs * 3, 256, 0, st>>>(P, degp.as<float>(), rects.as<LsdRect>(), nfaq.as<int>(), nfaq.as<int>() + batch + 1,
                                        rout, rvalid);
  PL_STAGE_END(timer, st);
  PL_STAGE_BEGIN(timer, "lsd_finish", st);
  PL_CARVEOUT(k_lsd_finish);
  k_lsd_finish<<<batch, 256, 0, st>>>(P, d_images, pitch, frame_stride, nrects.as<int>(), rout, rvalid,
                                            segs.as<LsdSegment>(), nsegs.as<int>(), resp.as<float>(), rowsum.as<float>(), d_keylines, d_desc,
                                            d_funcs, capacity, d_counts, status.as<int>());
  PL_STAGE_END(timer, st);
  PL_STAGE_BEGIN(timer, "lbd", st);
  PL_CARVEOUT(k_lbd);
  k_lbd<<<dim3(P.out_cap, batch), LBD_THREADS, 0, st>>>(P, d_images, pitch, frame_stride, d_keylines, d_counts, d_desc, capacity);
  PL_STAGE_END(timer, st);
  PL_CUDA(cudaGetLastError());
  return PLSLAM_OK;
}

int LineExtractor::check_status(cudaStream_t st) {
  PL_CUDA(cudaMemcpyAsync(pinnedStatus, status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  PL_CUDA(cudaStreamSynchronize(st));
  const int s = *reinterpret_cast<int*>(pinnedStatus);
  if (s != PLSLAM_OK) {
    set_error(s >= PLSLAM_ERR_INTERNAL ? "device status %d (region-growing scheduler: 5 = watchdog, 100 + n = consistency check n)" : "device status %d (rectangle list overflow: raise rect_cap)", s);
    cudaMemsetAsync(status.p, 0, sizeof(int), st);  // re-arm
  }
  return s;
}

int LineExtractor::extract_host(const uint8_t* images, int batch, int W, int H, int pitch, size_t frame_stride,
                                plslam_keyline_t* keylines, uint8_t* desc, double* funcs, int capacity, int32_t* counts) {
  PL_CHECK_ARG(images && keylines && desc && funcs && counts && batch >= 1 && pitch >= W);
  int rc = configure(W, H, batch);
  if (rc) return rc;
  const int cap = P.out_cap;
  const size_t dpitch = align_up(W, 32), dstride = dpitch * H;
  if ((rc = stageIn.ensure(dstride * batch))) return rc;
  if ((rc = stageKl.ensure((size_t)batch * cap * sizeof(plslam_keyline_t)))) return rc;
  if ((rc = stageDesc.ensure((size_t)batch * cap * 32))) return rc;
  if ((rc = stageFuncs.ensure((size_t)batch * cap * 3 * sizeof(double)))) return rc;
  if ((rc = stageCnt.ensure((size_t)batch * sizeof(int)))) return rc;
  cudaStream_t st = ownStream;
  for (int f = 0; f < batch; ++f)
    PL_CUDA(cudaMemcpy2DAsync(stageIn.as<uint8_t>() + f * dstride, dpitch, images + f * frame_stride, pitch, W, H,
                              cudaMemcpyHostToDevice, st));
  rc = extract_device(stageIn.as<uint8_t>(), batch, W, H, (int)dpitch, dstride, stageKl.as<plslam_keyline_t>(),
                      stageDesc.as<uint8_t>(), stageFuncs.as<double>(), cap, stageCnt.as<int>(), st);
  if (rc) return rc;
  PL_CUDA(cudaMemcpyAsync(counts, stageCnt.p, (size_t)batch * sizeof(int), cudaMemcpyDeviceToHost, st));
  if ((rc = check_status(st))) return rc;
  for (int f = 0; f < batch; ++f) {
    const int n = counts[f];
    if (n > capacity || n > cap) {
      set_error("frame %d has %d lines, capacity %d", f, n, std::min(capacity, cap));
      return PLSLAM_ERR_CAPACITY;
    }
    if (!n) continue;
    PL_CUDA(cudaMemcpyAsync(keylines + (size_t)f * capacity, stageKl.as<plslam_keyline_t>() + (size_t)f * cap,
                            (size_t)n * sizeof(plslam_keyline_t), cudaMemcpyDeviceToHost, st));
    PL_CUDA(cudaMemcpyAsync(desc + (size_t)f * capacity * 32, stageDesc.as<uint8_t>() + (size_t)f * cap * 32, (size_t)n * 32,
                            cudaMemcpyDeviceToHost, st));
    PL_CUDA(cudaMemcpyAsync(funcs + (size_t)f * capacity * 3, stageFuncs.as<double>() + (size_t)f * cap * 3,
                            (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  PL_CUDA(cudaStreamSynchronize(st));
  return PLSLAM_OK;
}

// BinaryDescriptor::compute(image, keylines, descriptors) of OpenCV-contrib line_descriptor, as ExtractLineSegment calls
// it (reference include/ExtractLineSegment.h:38): the LBD bytes of GIVEN key lines (host arrays in, host bytes out).
int LineExtractor::compute_lbd_host(const uint8_t* image, int W, int H, int pitch, const plslam_keyline_t* keylines, int n,
                                    uint8_t* desc) {
  PL_CHECK_ARG(image && pitch >= W && n >= 0);
  if (n == 0) return PLSLAM_OK;
  PL_CHECK_ARG(keylines && desc && n <= 65535);
  int rc = configure(W, H, 1);
  if (rc) return rc;
  const size_t dpitch = align_up(W, 32);
  DevBuf dImg, dKl, dDesc, dCnt;
  if ((rc = dImg.ensure(dpitch * H)) || (rc = dKl.ensure((size_t)n * sizeof(plslam_keyline_t))) || (rc = dDesc.ensure((size_t)n * 32)) ||
      (rc = dCnt.ensure(sizeof(int)))) {
    dImg.release(); dKl.release(); dDesc.release(); dCnt.release();
    return rc;
  }
  cudaStream_t st = ownStream;
  cudaError_t e = cudaMemcpy2DAsync(dImg.p, dpitch, image, pitch, W, H, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dKl.p, keylines, (size_t)n * sizeof(plslam_keyline_t), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dCnt.p, &n, sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    PL_CARVEOUT(k_lbd);
    k_lbd<<<dim3(n, 1), LBD_THREADS, 0, st>>>(P, dImg.as<uint8_t>(), (int)dpitch, dpitch * H, dKl.as<plslam_keyline_t>(), dCnt.as<int>(),
                                    dDesc.as<uint8_t>(), n);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(desc, dDesc.p, (size_t)n * 32, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  dImg.release(); dKl.release(); dDesc.release(); dCnt.release();
  if (e != cudaSuccess) {
    set_error("compute_lbd: %s", cudaGetErrorString(e));
    return PLSLAM_ERR_CUDA;
  }
  return PLSLAM_OK;
}

int LineExtractor::scaled_size(int* w, int* h) const {
  PL_CHECK_ARG(cfgW > 0);
  *w = P.sw;
  *h = P.sh;
  return PLSLAM_OK;
}

int LineExtractor::copy_scaled(int frame, uint8_t* out, size_t bytes) {
  PL_CHECK_ARG(cfgW > 0 && frame >= 0 && frame < last_batch && out && bytes >= (size_t)P.P);
  PL_CUDA(cudaDeviceSynchronize());
  PL_CUDA(cudaMemcpy2D(out, P.sw, scaled.as<uint8_t>() + (size_t)frame * P.spitch * P.sh, P.spitch, P.sw, P.sh,
                       cudaMemcpyDeviceToHost));
  return PLSLAM_OK;
}

int LineExtractor::copy_angles(int frame, float* deg_out, int32_t* g2_out, size_t n) {
  PL_CHECK_ARG(cfgW > 0 && frame >= 0 && frame < last_batch && n >= (size_t)P.P);
  PL_CUDA(cudaDeviceSynchronize());
  std::vector<uint4> h(P.P);
  PL_CUDA(cudaMemcpy(h.data(), pix.as<uint4>() + (size_t)frame * P.P, (size_t)P.P * sizeof(uint4), cudaMemcpyDeviceToHost));
  for (int i = 0; i < P.P; ++i) {
    if (deg_out) std::memcpy(&deg_out[i], &h[i].x, 4);
    if (g2_out) g2_out[i] = (int32_t)(h[i].w & 0x7fffffffu);  // bit 31 = LSD `used` flag
  }
  return PLSLAM_OK;
}

int LineExtractor::copy_segments(int frame, LsdSegment* out, int capacity, int* n_out) {
  PL_CHECK_ARG(cfgW > 0 && frame >= 0 && frame < last_batch && n_out);
  PL_CUDA(cudaDeviceSynchronize());
  int n = 0;
  PL_CUDA(cudaMemcpy(&n, nsegs.as<int>() + frame, sizeof(int), cudaMemcpyDeviceToHost));
  *n_out = n;
  if (n > capacity) return PLSLAM_ERR_CAPACITY;
  if (n) PL_CUDA(cudaMemcpy(out, segs.as<LsdSegment>() + (size_t)frame * P.rect_cap, (size_t)n * sizeof(LsdSegment), cudaMemcpyDeviceToHost));
  return PLSLAM_OK;
}

}  // namespace plslam
extern "C" int plslam_debug_grow_stats(unsigned long long* out32) {
  if (!out32) return PLSLAM_ERR_INVALID;
  return plslam::debug_aw_stats(out32);
}
#ifdef PLSLAM_GROW_PROF
extern "C" int plslam_debug_grow_prof(unsigned long long* out16) { return plslam::debug_grow_prof(out16); }
#endif
