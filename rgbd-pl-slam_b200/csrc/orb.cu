// orb.cu — B200-native ORB extractor: batched pyramid, per-cell FAST-9 + NMS, quad-tree keypoint
// distribution, intensity-centroid orientation, 7x7 integer blur and 256-bit rBRIEF.
//
// Replaces ORB_SLAM2::ORBextractor (reference include/ORBextractor.h:45-111; implementation only
// as machine code in lib/libORB_SLAM2.so, addresses cited per kernel).  All kernels are integer /
// byte work bounded by HBM traffic; none uses tensor cores.  One launch per stage covers the
// whole batch (grid.y = frame), so a 256-frame batch is ~16 launches.
#include "orb.cuh"

#include <cuda.h>
#include "fast_score.cuh"
#include "pl_math.cuh"

#include <algorithm>
#include <cmath>

namespace plslam {

namespace {

// bit_pattern_31_ (256 x 4 coordinates in [-13, 12]), see tools/extract_pattern.py.  Kept in global
// memory (not __constant__): every lane reads its own 32 bytes, which the constant cache would serialise.
__device__ __align__(16) int8_t g_pattern[1024];
const int h_pattern[1024] = {
#include "orb_pattern.inc"
};

__device__ __forceinline__ const uint8_t* level_ptr(const OrbParams& P, const OrbImages& I, int f, int l, int& pitch) {
  if (l == 0) {
    pitch = I.pitch0;
    return I.img0 + (size_t)f * I.stride0;
  }
  pitch = P.lv[l].pitch;
  return I.pyr + (size_t)f * P.pyrFrameStride + P.lv[l].off;
}

// ------------------------------------------------------------------------------------------
// K1  pyramid level l from level l-1: cv::resize INTER_LINEAR, 11-bit fixed point
// (ComputePyramid @0x70430, resize call @0x70b07; arithmetic SURVEY B.1).
// One thread = 4 horizontally adjacent output pixels (one uchar4 store).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int resize_px(const uint8_t* S0, const uint8_t* S1, int sx, int xab, int b0, int b1) {
  const int a0 = (short)(xab & 0xffff), a1 = xab >> 16;
  int t0 = S0[sx] * a0, t1 = S1[sx] * a0;
  if (a1) {
    t0 += S0[sx + 1] * a1;
    t1 += S1[sx + 1] * a1;
  }
  return (((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2;
}

__global__ void __launch_bounds__(256) k_resize(const __grid_constant__ OrbParams P, OrbImages I, int l,
                                                const int* __restrict__ coef) {
  const OrbLevel& D = P.lv[l];
  const int f = blockIdx.z;
  const int x4 = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x4 >= D.w || y >= D.h) return;
  int sp;
  const uint8_t* S = level_ptr(P, I, f, l - 1, sp);
  const int sh = P.lv[l - 1].h;
  // tables: int2 {source offset, c0 | c1 << 16} per destination column, then per destination row (padded to 4 entries)
  const int2* xtab = reinterpret_cast<const int2*>(coef + D.coefOff);
  const int2* ytab = xtab + ((D.w + 3) & ~3);
  const int2 ye = __ldg(ytab + y);
  const int sy = ye.x, b0 = (short)(ye.y & 0xffff), b1 = ye.y >> 16;
  const uint8_t* S0 = S + (size_t)sy * sp;
  const uint8_t* S1 = S + (size_t)min(sy + 1, sh - 1) * sp;
  const int4 e01 = __ldg(reinterpret_cast<const int4*>(xtab + x4)), e23 = __ldg(reinterpret_cast<const int4*>(xtab + x4) + 1);
  const uint32_t out = (uint32_t)(resize_px(S0, S1, e01.x, e01.y, b0, b1) & 0xff) |
                       ((uint32_t)(resize_px(S0, S1, e01.z, e01.w, b0, b1) & 0xff) << 8) |
                       ((uint32_t)(resize_px(S0, S1, e23.x, e23.y, b0, b1) & 0xff) << 16) |
                       ((uint32_t)(resize_px(S0, S1, e23.z, e23.w, b0, b1) & 0xff) << 24);
  uint8_t* Dp = I.pyr + (size_t)f * P.pyrFrameStride + D.off + (size_t)y * D.pitch + x4;
  *reinterpret_cast<uint32_t*>(Dp) = out;
}

// ------------------------------------------------------------------------------------------
// K2  per-cell FAST-9 + 3x3 NMS + empty-cell retry (ComputeKeyPointsOctTree FAST stage
// @0x75fa0-0x76890; cv::FAST call sites @0x763d4 / @0x76753; arithmetic SURVEY B.3).
// One warp = one 30-px cell (with its 3-px dead border).  The cell tile is staged in shared
// memory; pixels that pass the 4-point quick test at minThFAST are warp-compacted into a queue
// so the full 16-arc score is evaluated with all lanes busy.
//
// Arc score s = max over the 16 arcs of 9 contiguous circle pixels of min|v-p| (better polarity).
// corner(th) <=> s > th, response = s-1.  NMS(th) = strict 3x3 local maxima of the raw s map
// with s > th (neighbours below th count as 0 in OpenCV, which is < s either way), so the local
// maxima are found once and the threshold only selects among them.
//
// Candidate record (64 bit): [63:56] response+1 (=s), [55:28] list-order key, [27:14] Y, [13:0] X
// (X, Y border-local).  key = ((cellRow*nCols + cellCol)*(hCell+6) + yLocal)*(wCell+6) + xLocal reproduces the
// reference list order (cell-major, then FAST's row-major) without ordered writes.
// ------------------------------------------------------------------------------------------
constexpr int FAST_LIST_CAP = 512;                  // survivor list entries per cell (warp)
constexpr int FAST_QUEUE_BYTES = 2 * FAST_LIST_CAP;  // u16 positions
__global__ void __launch_bounds__(256) k_fast(const __grid_constant__ OrbParams P, OrbImages I,
                                              unsigned long long* __restrict__ cand, int* __restrict__ candCount,
                                              int* __restrict__ status) {
  extern __shared__ uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cell = blockIdx.x * 8 + warp;
  const int f = blockIdx.y;
  if (cell >= P.totalCells) return;
  int l = 0;
  while (l + 1 < P.nlevels && cell >= P.lv[l + 1].cellBase) ++l;
  const OrbLevel& L = P.lv[l];
  const int ci = (cell - L.cellBase) / L.nCols, cj = (cell - L.cellBase) % L.nCols;
  const int iniY = ORB_MINB + ci * L.hCell, iniX = ORB_MINB + cj * L.wCell;
  if (iniY >= L.maxBorderY - 3 || iniX >= L.maxBorderX - 6) return;
  const int pw = min(iniX + L.wCell + 6, L.maxBorderX) - iniX;
  const int ph = min(iniY + L.hCell + 6, L.maxBorderY) - iniY;
  if (pw < 7 || ph < 7) return;

  const int pp = P.patchPitch;  // multiple of 4
  const int tileBytes = pp * P.patchRows;
  uint8_t* patch = smem + (size_t)warp * (2 * tileBytes + FAST_QUEUE_BYTES);
  uint8_t* score = patch + tileBytes;
  unsigned short* queue = reinterpret_cast<unsigned short*>(score + tileBytes);  // FAST_QUEUE_BYTES / 2 entries
  uint32_t* PW = reinterpret_cast<uint32_t*>(patch);
  uint32_t* SW = reinterpret_cast<uint32_t*>(score);
  const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;

  int sp;
  const uint8_t* S = level_ptr(P, I, f, l, sp);
  S += (size_t)iniY * sp + iniX;
  // Stage the cell with aligned 32-bit loads (lane = (row, word) of a group of rows): pixel x of the cell then lives
  // in byte column shift + x of the shared tile.  Byte loads when the level's pitch is not a multiple of 4.
  for (int i = lane; i < tileBytes / 4; i += 32) SW[i] = 0;
  const int nwd = pp >> 2;                           // words per tile row
  const int rpi = nwd <= 32 ? 32 / nwd : 1;          // tile rows handled per warp iteration
  const int r = nwd <= 32 ? lane / nwd : 0;          // this lane's row within the group ...
  const int wl = nwd <= 32 ? lane - r * nwd : lane;  // ... and word within the row
  const bool rowLane = r < rpi;
  int shift = (int)(reinterpret_cast<uintptr_t>(S) & 3);
  if ((sp & 3) == 0 && shift + pw <= pp) {
    const uint32_t* Sw = reinterpret_cast<const uint32_t*>(S - shift);
    const int nw = (shift + pw + 3) >> 2;
    const int spw = sp >> 2;
    for (int wb = 0; wb < nw; wb += 32) {
      const int w = wb + wl;
      const bool on = rowLane && w < nw;
      const uint32_t* src = Sw + (size_t)r * spw + w;  // this lane's column of words, rows r, r + rpi, ...
      uint32_t* dst = PW + r * nwd + w;
      // four rows per step: the loads are issued together, so a cell costs ~4 memory round trips instead of ~13
      for (int y = r; y < ph; y += 4 * rpi) {
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (on && y + k * rpi < ph) ? __ldg(src + (size_t)k * rpi * spw) : 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (on && y + k * rpi < ph) dst[k * rpi * nwd] = v[k];
        src += (size_t)4 * rpi * spw;
        dst += 4 * rpi * nwd;
      }
    }
  } else {
    shift = 0;
    for (int y = 0; y < ph; ++y)
      for (int x = lane; x < pw; x += 32) patch[y * pp + x] = S[(size_t)y * sp + x];
  }
  __syncwarp();

  // All passes below work on 4 pixels per lane with the byte-SIMD instructions.  Detection area: rows [3, yEnd),
  // byte columns [cLo, cHi).
  const int yEnd = ph - 3, cLo = shift + 3, cHi = shift + pw - 3;
  auto valid_mask_of = [&](int w) -> unsigned {  // bytes of word w inside [cLo, cHi)
    const int lo = max(cLo - 4 * w, 0), hi = min(cHi - 4 * w, 4);
    return hi > lo ? ((FULL >> (32 - 8 * hi)) & (FULL << (8 * lo))) : 0u;
  };
  // the usual case (a tile row fits one warp iteration): the lane's word, hence its mask, never changes
  const unsigned vmLane = (rowLane && wl < nwd) ? valid_mask_of(wl) : 0u;
  auto valid_mask = [&](int w) -> unsigned { return nwd <= 32 ? vmLane : valid_mask_of(w); };
  // exclusive prefix over the lanes of a per-lane count in [0, 4]; *total = warp sum
  auto lane_prefix = [&](int cnt, int* total) -> int {
    const unsigned b0 = __ballot_sync(FULL, cnt & 1), b1 = __ballot_sync(FULL, cnt & 2), b2 = __ballot_sync(FULL, cnt & 4);
    *total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
    return __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
  };
  const int th0 = P.minTh;
  const unsigned th4 = (unsigned)th0 * 0x01010101u;

  // pass 1: 4-point quick test at minThFAST (N/S and E/W compass pixels); the survivors are appended to a list (in
  // FAST's row-major order) and scored 32 at a time with the 16-arc test.  The list is kept for the later passes, which
  // then only visit the ~25 % of pixels that can be corners; if it would overflow it degrades to a rolling queue
  // and the later passes scan the whole cell.
  int ns = 0, sc = 0;  // entries appended / scored
  bool overflow = false;
  for (int yb = 3; yb < yEnd; yb += rpi)
    for (int wb = 0; wb < nwd; wb += 32) {
      const int y = yb + r, w = wb + wl;
      unsigned pass = 0;
      if (rowLane && y < yEnd && w < nwd) {
        const unsigned vm = valid_mask(w);
        if (vm) {
          const uint32_t* R = PW + y * nwd;
          const unsigned C = R[w], N = R[w - 3 * nwd], So = R[w + 3 * nwd];
          const unsigned E = __funnelshift_r(C, R[w + 1], 24), We = __funnelshift_r(R[w - 1], C, 8);
          const unsigned a = __vcmpgtu4(__vabsdiffu4(C, N), th4) | __vcmpgtu4(__vabsdiffu4(C, So), th4);
          const unsigned b = __vcmpgtu4(__vabsdiffu4(C, E), th4) | __vcmpgtu4(__vabsdiffu4(C, We), th4);
          pass = a & b & vm & 0x01010101u;
        }
      }
      if (ns + 128 > FAST_LIST_CAP) {  // no room for a full iteration: score what is pending, restart as a queue
        for (; sc < ns; sc += 32)
          if (sc + lane < ns) {
            const int q = queue[sc + lane];
            score[q] = (uint8_t)fast_arc_score(patch + q, pp);
          }
        __syncwarp();
        overflow = true;
        ns = sc = 0;
      }
      int tot;
      int at = ns + lane_prefix(__popc(pass), &tot);
      const int base = y * pp + 4 * w;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if ((pass >> (8 * k)) & 1u) queue[at++] = (unsigned short)(base + k);
      ns += tot;
      __syncwarp();
      while (ns - sc >= 32) {
        const int q = queue[sc + lane];
        score[q] = (uint8_t)fast_arc_score(patch + q, pp);
        sc += 32;
      }
      __syncwarp();
    }
  if (sc + lane < ns) {
    const int q = queue[sc + lane];
    score[q] = (uint8_t)fast_arc_score(patch + q, pp);
  }
  __syncwarp();

  if (!overflow) {
    // pass 2 over the survivors: strict 3x3 local maxima of the raw score map (scores outside the detection area are
    // 0); the maxima (score, else 0) go to the pixel tile, which is no longer read as pixels
    int nHi = 0, nLo = 0;
    for (int b0 = 0; b0 < ns; b0 += 32) {
      int keep = 0;
      if (b0 + lane < ns) {
        const int q = queue[b0 + lane];
        const uint8_t* sp8 = score + q;
        const int c = sp8[0];
        if (c > th0) {
          const int m = max(max(max(sp8[-pp - 1], sp8[-pp]), max(sp8[-pp + 1], sp8[-1])),
                            max(max(sp8[1], sp8[pp - 1]), max(sp8[pp], sp8[pp + 1])));
          if (c > m) keep = c;
        }
        patch[q] = (uint8_t)keep;
      }
      nHi += __popc(__ballot_sync(FULL, keep > P.iniTh));
      nLo += __popc(__ballot_sync(FULL, keep > 0));
    }
    __syncwarp();
    const int thSel = nHi > 0 ? P.iniTh : th0;
    const int nOut = nHi > 0 ? nHi : nLo;
    if (nOut == 0) return;
    int slot = 0;
    if (lane == 0) slot = atomicAdd(candCount + f * ORB_MAXL + l, nOut);
    slot = __shfl_sync(FULL, slot, 0);
    if (slot + nOut > L.candCap) {
      if (lane == 0) atomicMax(status, PLSLAM_ERR_OVERFLOW);
      return;
    }
    unsigned long long* out = cand + (size_t)f * P.candFrameStride + L.candOff + slot;
    // pass 3 over the survivors: emit the maxima above the selected threshold (list order = FAST's row-major order)
    int wr = 0;
    for (int b0 = 0; b0 < ns; b0 += 32) {
      int q = 0, keep = 0;
      if (b0 + lane < ns) {
        q = queue[b0 + lane];
        keep = patch[q];
      }
      const bool emit = keep > thSel;
      const unsigned m = __ballot_sync(FULL, emit);
      if (emit) {
        const int y = q / pp, x = q - y * pp - shift;
        const unsigned key = (((unsigned)(ci * L.nCols + cj) * (unsigned)(L.hCell + 6) + (unsigned)y) * (unsigned)(L.wCell + 6) + (unsigned)x);
        const unsigned X = x + cj * L.wCell, Y = y + ci * L.hCell;
        out[wr + __popc(m & lt)] =
            ((unsigned long long)keep << 56) | ((unsigned long long)key << 28) | ((unsigned long long)Y << 14) | X;
      }
      wr += __popc(m);
    }
    return;
  }

  // survivor list overflowed (a very noisy cell): full-cell passes
  // pass 2: strict 3x3 local maxima of the raw score map (scores outside the detection area are 0); the maxima
  // (score, else 0) overwrite the pixel tile, which is no longer read as pixels
  const unsigned ini4 = (unsigned)P.iniTh * 0x01010101u;
  int nHi = 0, nLo = 0;
  for (int yb = 3; yb < yEnd; yb += rpi)
    for (int wb = 0; wb < nwd; wb += 32) {
      const int y = yb + r, w = wb + wl;
      if (rowLane && y < yEnd && w < nwd) {
        const unsigned vm = valid_mask(w);
        unsigned keep4 = 0;
        const uint32_t* R = SW + y * nwd;
        const unsigned C = R[w];
        if (C & vm) {
          const uint32_t* U = R - nwd;
          const uint32_t* D = R + nwd;
          const unsigned u0 = U[w], d0 = D[w];
          const unsigned m =
              __vmaxu4(__vmaxu4(__vmaxu4(__funnelshift_r(U[w - 1], u0, 24), u0),
                                __vmaxu4(__funnelshift_r(u0, U[w + 1], 8), __funnelshift_r(R[w - 1], C, 24))),
                       __vmaxu4(__vmaxu4(__funnelshift_r(C, R[w + 1], 8), __funnelshift_r(D[w - 1], d0, 24)),
                                __vmaxu4(d0, __funnelshift_r(d0, D[w + 1], 8))));
          keep4 = C & __vcmpgtu4(C, m) & __vcmpgtu4(C, th4) & vm;
        }
        PW[y * nwd + w] = keep4;
        nHi += __popc(__vcmpgtu4(keep4, ini4) & 0x01010101u);
        nLo += __popc(__vcmpgtu4(keep4, 0u) & 0x01010101u);
      }
    }
  nHi = __reduce_add_sync(FULL, nHi);
  nLo = __reduce_add_sync(FULL, nLo);
  __syncwarp();
  const int thSel = nHi > 0 ? P.iniTh : th0;
  const int nOut = nHi > 0 ? nHi : nLo;
  if (nOut == 0) return;
  int slot = 0;
  if (lane == 0) slot = atomicAdd(candCount + f * ORB_MAXL + l, nOut);
  slot = __shfl_sync(0xffffffffu, slot, 0);
  if (slot + nOut > L.candCap) {
    if (lane == 0) atomicMax(status, PLSLAM_ERR_OVERFLOW);
    return;
  }
  unsigned long long* out = cand + (size_t)f * P.candFrameStride + L.candOff + slot;
  // pass 3: emit the maxima above the selected threshold in FAST's row-major order
  const unsigned sel4 = (unsigned)thSel * 0x01010101u;
  int wr = 0;
  for (int yb = 3; yb < yEnd; yb += rpi)
    for (int wb = 0; wb < nwd; wb += 32) {
      const int y = yb + r, w = wb + wl;
      unsigned kw = 0, e = 0;
      if (rowLane && y < yEnd && w < nwd) {
        kw = PW[y * nwd + w];
        e = __vcmpgtu4(kw, sel4) & 0x01010101u;
      }
      int tot;
      int at = wr + lane_prefix(__popc(e), &tot);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if ((e >> (8 * k)) & 1u) {
          const int x = 4 * w + k - shift;
          const unsigned keep = (kw >> (8 * k)) & 0xffu;
          const unsigned key = (((unsigned)(ci * L.nCols + cj) * (unsigned)(L.hCell + 6) + (unsigned)y) * (unsigned)(L.wCell + 6) + (unsigned)x);
          const unsigned X = x + cj * L.wCell, Y = y + ci * L.hCell;
          out[at++] = ((unsigned long long)keep << 56) | ((unsigned long long)key << 28) | ((unsigned long long)Y << 14) | X;
        }
      wr += tot;
    }
}

// ------------------------------------------------------------------------------------------
// K3  DistributeOctTree (@0x73c60) + DivideNode (@0x70c60): one CTA per (frame, level).
// The std::list of nodes is kept as an array in list order (index == position); every pass
// rebuilds it: children of the divided nodes are "push_front"ed in division order (so they end
// reversed at the head) and untouched nodes follow in their old order.  A full pass divides all
// expandable nodes in list order; the finishing phase divides them in (size, creation seq)
// descending order and stops at the first prefix that reaches N nodes — all divisions are
// independent, only the cut depends on the running count, so it is a prefix sum.
// Tie-break pinned to creation sequence (the reference compares heap addresses, @0x74d74).
// ------------------------------------------------------------------------------------------
struct QtNode {
  short x0, y0, x1, y1;
};

__global__ void __launch_bounds__(256) k_quadtree(const __grid_constant__ OrbParams P,
                                                  const unsigned long long* __restrict__ cand,
                                                  const int* __restrict__ candCount, unsigned short* __restrict__ knodeAll,
                                                  uint2* __restrict__ lvlKp, int* __restrict__ lvlCnt) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int l = blockIdx.x, f = blockIdx.y, t = threadIdx.x, T = blockDim.x;
  const OrbLevel& L = P.lv[l];
  const int NC = P.nodeCap;
  // shared layout
  QtNode* box[2];
  int* cnt[2];
  box[0] = reinterpret_cast<QtNode*>(smem);
  box[1] = box[0] + NC;
  cnt[0] = reinterpret_cast<int*>(box[1] + NC);
  cnt[1] = cnt[0] + NC;
  int* seq[2];
  seq[0] = cnt[1] + NC;
  seq[1] = seq[0] + NC;
  int* childCnt = seq[1] + NC;      // [NC][4]; later reused as child new position
  int* newPos = childCnt + 4 * NC;  // new index of a non-divided node / scan scratch
  int* ord = newPos + NC;           // division order (node indices)
  int* pushOff = ord + NC;          // per order slot: exclusive count of pushed children
  int* expOff = pushOff + NC;       // per order slot: exclusive count of pushed children with cnt>1
  int* divRank = expOff + NC;       // per node: rank in division order or -1
  unsigned long long* best = reinterpret_cast<unsigned long long*>(divRank + NC + (NC & 1));
  __shared__ int warpTmp[33];
  __shared__ int sh_n, sh_D, sh_mode, sh_done, sh_E;

  const int nk = min(candCount[f * ORB_MAXL + l], L.candCap);
  const unsigned long long* K = cand + (size_t)f * P.candFrameStride + L.candOff;
  unsigned short* knode = knodeAll + (size_t)f * P.candFrameStride + L.candOff;
  const int N = L.quota;
  int* outCnt = lvlCnt + f * ORB_MAXL + l;
  if (nk == 0) {
    if (t == 0) *outCnt = 0;
    return;
  }

  // roots
  const int nIni = L.nIni;
  const int H = L.maxBorderY - ORB_MINB;
  for (int i = t; i < NC; i += T) { cnt[0][i] = 0; cnt[1][i] = 0; }
  __syncthreads();
  for (int k = t; k < nk; k += T) {
    const int X = (int)(K[k] & 0x3fff);
    int r = (int)__fdiv_rn((float)X, L.hX);
    r = min(r, nIni - 1);
    knode[k] = (unsigned short)r;
    atomicAdd(&cnt[0][r], 1);
  }
  __syncthreads();
  if (t == 0) {
    // erase empty roots (list order preserved); nIni is 1..few
    int n = 0;
    for (int i = 0; i < nIni; ++i) {
      const int c = cnt[0][i];
      newPos[i] = c ? n : -1;
      if (c) {
        QtNode b;
        b.x0 = (short)(int)(L.hX * (float)i);
        b.x1 = (short)(int)(L.hX * (float)(i + 1));
        b.y0 = 0;
        b.y1 = (short)H;
        box[1][n] = b;
        cnt[1][n] = c;
        ++n;
      }
    }
    sh_n = n;
    sh_mode = 0;
    sh_done = 0;
  }
  __syncthreads();
  if (nIni > 1) {
    for (int k = t; k < nk; k += T) knode[k] = (unsigned short)newPos[knode[k]];
  }
  int cur = 1;  // buffer holding the current list
  __syncthreads();

  while (true) {
    const int n = sh_n, mode = sh_mode;
    QtNode* B = box[cur];
    int* C = cnt[cur];
    int* SQ = seq[cur];
    // A. division order
    if (mode == 0) {
      for (int i = t; i < n; i += T) newPos[i] = C[i] > 1 ? 1 : 0;
      __syncthreads();
      const int E = block_scan_excl(newPos, n, warpTmp);
      for (int i = t; i < n; i += T) {
        if (C[i] > 1) { ord[newPos[i]] = i; divRank[i] = newPos[i]; } else divRank[i] = -1;
      }
      if (t == 0) sh_E = E;
    } else {
      // rank by (size desc, seq desc) among expandable nodes
      for (int i = t; i < n; i += T) {
        int r = -1;
        const int ci = C[i];
        if (ci > 1) {
          const int si = SQ[i];
          r = 0;
          for (int j = 0; j < n; ++j) {
            const int cj = C[j];
            if (cj > 1 && (cj > ci || (cj == ci && SQ[j] > si))) ++r;
          }
          ord[r] = i;
        }
        divRank[i] = r;
      }
      if (t == 0) {
        int E = 0;
        for (int i = 0; i < n; ++i) E += C[i] > 1;
        sh_E = E;
      }
    }
    for (int i = t; i < 4 * n; i += T) childCnt[i] = 0;
    __syncthreads();
    const int E = sh_E;
    // B. quadrant of every key of an expandable node
    for (int k = t; k < nk; k += T) {
      const int nd = knode[k];
      if (C[nd] > 1) {
        const unsigned long long r = K[k];
        const int X = (int)(r & 0x3fff), Y = (int)((r >> 14) & 0x3fff);
        const QtNode b = B[nd];
        const int mx = b.x0 + (b.x1 - b.x0 + 1) / 2, my = b.y0 + (b.y1 - b.y0 + 1) / 2;
        const int q = (X < mx ? 0 : 1) + (Y < my ? 0 : 2);
        atomicAdd(&childCnt[4 * nd + q], 1);
      }
    }
    __syncthreads();
    // C. pushes per division slot, cut for the finishing phase
    for (int e = t; e < E; e += T) {
      const int nd = ord[e];
      int ne = 0, nx = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = childCnt[4 * nd + q];
        ne += c > 0;
        nx += c > 1;
      }
      pushOff[e] = ne;
      expOff[e] = nx;
    }
    __syncthreads();
    block_scan_excl(pushOff, E, warpTmp);
    if (t == 0) {
      int D = E;
      if (mode == 1) {
        // smallest prefix with n + sum(pushed - 1) >= N
        for (int e = 0; e < E; ++e) {
          const int nd = ord[e];
          int ne = 0;
          for (int q = 0; q < 4; ++q) ne += childCnt[4 * nd + q] > 0;
          const int after = n + (pushOff[e] + ne) - (e + 1);
          if (after >= N) { D = e + 1; break; }
        }
      }
      sh_D = D;
    }
    __syncthreads();
    const int D = sh_D;
    block_scan_excl(expOff, D, warpTmp);
    // total pushes / expandable children among the first D slots
    int Ctot, Etot;
    {
      const int nd = ord[D - 1 < 0 ? 0 : D - 1];
      int ne = 0, nx = 0;
      for (int q = 0; q < 4; ++q) {
        const int c = childCnt[4 * nd + q];
        ne += c > 0;
        nx += c > 1;
      }
      Ctot = D > 0 ? pushOff[D - 1] + ne : 0;
      Etot = D > 0 ? expOff[D - 1] + nx : 0;
    }
    // D. positions of untouched nodes
    for (int i = t; i < n; i += T) newPos[i] = (divRank[i] < 0 || divRank[i] >= D) ? 1 : 0;
    __syncthreads();
    const int keepTot = block_scan_excl(newPos, n, warpTmp);
    QtNode* B2 = box[cur ^ 1];
    int* C2 = cnt[cur ^ 1];
    int* SQ2 = seq[cur ^ 1];
    for (int i = t; i < n; i += T) {
      const int r = divRank[i];
      if (r < 0 || r >= D) {
        const int np = Ctot + newPos[i];
        newPos[i] = np;
        B2[np] = B[i];
        C2[np] = C[i];
        SQ2[np] = SQ[i];
      } else {
        const QtNode b = B[i];
        const int mx = b.x0 + (b.x1 - b.x0 + 1) / 2, my = b.y0 + (b.y1 - b.y0 + 1) / 2;
        int p = pushOff[r], x = expOff[r];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = childCnt[4 * i + q];
          if (c > 0) {
            const int np = Ctot - 1 - p;
            QtNode cb;
            cb.x0 = (q & 1) ? (short)mx : b.x0;
            cb.x1 = (q & 1) ? b.x1 : (short)mx;
            cb.y0 = (q & 2) ? (short)my : b.y0;
            cb.y1 = (q & 2) ? b.y1 : (short)my;
            B2[np] = cb;
            C2[np] = c;
            SQ2[np] = c > 1 ? x : -1;
            childCnt[4 * i + q] = np;  // reuse as new position
            ++p;
            x += c > 1;
          }
        }
      }
    }
    __syncthreads();
    // E. relabel keys
    for (int k = t; k < nk; k += T) {
      const int nd = knode[k];
      const int r = divRank[nd];
      if (r < 0 || r >= D) {
        knode[k] = (unsigned short)newPos[nd];
      } else {
        const unsigned long long rec = K[k];
        const int X = (int)(rec & 0x3fff), Y = (int)((rec >> 14) & 0x3fff);
        const QtNode b = B[nd];
        const int mx = b.x0 + (b.x1 - b.x0 + 1) / 2, my = b.y0 + (b.y1 - b.y0 + 1) / 2;
        const int q = (X < mx ? 0 : 1) + (Y < my ? 0 : 2);
        knode[k] = (unsigned short)childCnt[4 * nd + q];
      }
    }
    __syncthreads();
    if (t == 0) {
      const int n2 = Ctot + keepTot;
      sh_n = n2;
      if (n2 >= N || n2 == n) sh_done = 1;
      else if (mode == 0 && n2 + 3 * Etot > N) sh_mode = 1;
    }
    cur ^= 1;
    __syncthreads();
    if (sh_done) break;
  }

  // best key per node: max response, first in list order on ties (cmova @0x75a77-0x75a7f)
  const int n = sh_n;
  for (int i = t; i < n; i += T) best[i] = 0ull;
  __syncthreads();
  for (int k = t; k < nk; k += T) {
    const unsigned long long r = K[k];
    const unsigned long long v = ((r >> 56) << 28) | (0x0fffffffull - ((r >> 28) & 0x0fffffffull));
    atomicMax(&best[knode[k]], v);
  }
  __syncthreads();
  uint2* out = lvlKp + (size_t)f * P.maxKp + L.kpOff;
  for (int i = t; i < n; i += T) {
    const unsigned long long v = best[i];
    const unsigned key = 0x0fffffffu - (unsigned)(v & 0x0fffffffull);
    const int s = (int)(v >> 28);
    const unsigned wSpan = (unsigned)(L.wCell + 6), hSpan = (unsigned)(L.hCell + 6);
    const int xl = (int)(key % wSpan), yl = (int)((key / wSpan) % hSpan);
    const unsigned cellIdx = key / (wSpan * hSpan);
    const int cj = (int)(cellIdx % (unsigned)L.nCols), ci = (int)(cellIdx / (unsigned)L.nCols);
    const int X = xl + cj * L.wCell + ORB_MINB, Y = yl + ci * L.hCell + ORB_MINB;
    out[i] = make_uint2((unsigned)X | ((unsigned)Y << 16), (unsigned)(s - 1));
  }
  if (t == 0) *outCnt = n;
}

// ------------------------------------------------------------------------------------------
// K4  7x7 Gaussian blur, sigma 2, BORDER_REFLECT_101, on the borderless level (GaussianBlur call
// @0x77487; arithmetic SURVEY B.2): dst = (sum_y sum_x k[y]k[x]p + 32768) >> 16.
// Tile = 128x16 outputs per CTA.  The (128+16)x22 source box is fetched by one TMA bulk-tensor copy
// (cp.async.bulk.tensor.3d, tensor = [frame][row][byte] of the level) into shared memory and signalled
// through an mbarrier; TMA zero-fills outside the level, so only CTAs on the level's rim patch the
// 3-px reflected halo by hand.  Levels whose base/pitch are not 16-byte aligned (a caller's level-0
// image) fall back to plain loads inside the same kernel.  Horizontal pass: 4 outputs per item with
// funnel-shifted byte windows and dp4a; vertical pass: one thread per column half with a register window.
// ------------------------------------------------------------------------------------------
// TMA needs the innermost start coordinate 16-byte aligned, so the box starts 16 px left of the tile
// (BOX_XOFF = 13 columns before the 3-px halo) and is 160 bytes wide.
constexpr int BT_W = 128, BT_H = 16, BOX_W = 160, BOX_H = BT_H + 6, BOX_XOFF = 13;

struct BlurMaps {
  CUtensorMap m[ORB_MAXL];
  unsigned tmaMask;  // bit l set: level l is read through its tensor map
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// K1, TMA form: the source patch of a 128x8 output tile (<= 176 x 12 bytes at level ratios up to 1.23) arrives by one
// cp.async.bulk.tensor.3d from the [frame][row][byte] view of the source level; the 256 threads then take their taps from shared
// memory.  Used for every level whose source view is TMA-legal (16-byte aligned base, pitch and frame stride) and whose ratio
// fits the box; k_resize (plain loads) is the fallback.  The box starts at the tile's first source column rounded down to 16.
constexpr int RS_TW = 128, RS_TH = 8, RS_BW = 176, RS_BH = 12;
struct ResizeMaps {
  CUtensorMap m[ORB_MAXL];  // m[l]: level l as the SOURCE of level l + 1, box RS_BW x RS_BH
  unsigned tmaMask;         // bit l set: level l (destination) is built by k_resize_tma
};

struct AllMaps {
  BlurMaps blur;
  ResizeMaps rs;
};

__global__ void __launch_bounds__(256) k_resize_tma(const __grid_constant__ OrbParams P, const __grid_constant__ ResizeMaps Mp,
                                                    OrbImages I, int l, const int* __restrict__ coef) {
  __shared__ __align__(128) uint8_t patch[RS_BH][RS_BW];
  __shared__ __align__(8) unsigned long long mbar;
  const OrbLevel& D = P.lv[l];
  const int f = blockIdx.z;
  const int x0 = blockIdx.x * RS_TW, y0 = blockIdx.y * RS_TH;
  const int2* xtab = reinterpret_cast<const int2*>(coef + D.coefOff);
  const int2* ytab = xtab + ((D.w + 3) & ~3);
  const int bx = __ldg(&xtab[x0].x) & ~15, by = __ldg(&ytab[y0].x);
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const unsigned bar = smem_u32(&mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(RS_BW * RS_BH) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(&patch[0][0])), "l"(reinterpret_cast<unsigned long long>(&Mp.m[l - 1])), "r"(bx), "r"(by), "r"(f), "r"(bar)
        : "memory");
    unsigned done = 0;  // one thread polls, the others park at the CTA barrier (as in k_blur)
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar)
          : "memory");
    }
  }
  __syncthreads();
  const int x4 = x0 + threadIdx.x * 4, y = y0 + threadIdx.y;
  if (x4 >= D.w || y >= D.h) return;
  const int sh = P.lv[l - 1].h;
  const int2 ye = __ldg(ytab + y);
  const int sy = ye.x, b0 = (short)(ye.y & 0xffff), b1 = ye.y >> 16;
  const uint8_t* S0 = &patch[sy - by][0] - bx;
  const uint8_t* S1 = &patch[min(sy + 1, sh - 1) - by][0] - bx;
  const int4 e01 = __ldg(reinterpret_cast<const int4*>(xtab + x4)), e23 = __ldg(reinterpret_cast<const int4*>(xtab + x4) + 1);
  const uint32_t out = (uint32_t)(resize_px(S0, S1, e01.x, e01.y, b0, b1) & 0xff) |
                       ((uint32_t)(resize_px(S0, S1, e01.z, e01.w, b0, b1) & 0xff) << 8) |
                       ((uint32_t)(resize_px(S0, S1, e23.x, e23.y, b0, b1) & 0xff) << 16) |
                       ((uint32_t)(resize_px(S0, S1, e23.z, e23.w, b0, b1) & 0xff) << 24);
  uint8_t* Dp = I.pyr + (size_t)f * P.pyrFrameStride + D.off + (size_t)y * D.pitch + x4;
  *reinterpret_cast<uint32_t*>(Dp) = out;
}

// The tensor maps travel as a __grid_constant__ kernel parameter (12 x 128 B): no device copy of them exists, so a caller
// that cycles through any number of input buffers never waits for a descriptor upload.
__global__ void __launch_bounds__(256) k_blur(const __grid_constant__ OrbParams P, const __grid_constant__ BlurMaps Mp,
                                              OrbImages I, const int* __restrict__ lvlCnt,
                                              const unsigned* __restrict__ tileTab) {
  __shared__ __align__(128) uint8_t raw[BOX_H][BOX_W];
  __shared__ __align__(16) unsigned short hs[BOX_H][BT_W];
  __shared__ __align__(8) unsigned long long mbar;
  const int f = blockIdx.y;
  const unsigned tt = __ldg(tileTab + blockIdx.x);  // level << 28 | tile row << 14 | tile column (host table)
  const int l = (int)(tt >> 28);
  if (lvlCnt[f * ORB_MAXL + l] == 0) return;  // the reference blurs only levels with keypoints
  const OrbLevel& L = P.lv[l];
  const int tx = (int)(tt & 0x3fffu) * BT_W, ty = (int)((tt >> 14) & 0x3fffu) * BT_H;
  const int w = L.w, h = L.h;
  int sp;
  const uint8_t* S = level_ptr(P, I, f, l, sp);
  const bool useTma = (Mp.tmaMask >> l) & 1u;
  const bool rim = tx < 3 || ty < 3 || tx + BT_W + 3 > w || ty + BT_H + 3 > h;
  if (useTma) {
    if (threadIdx.x == 0) {
      const unsigned bar = smem_u32(&mbar);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BOX_W * BOX_H) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
          ::"r"(smem_u32(&raw[0][0])), "l"(reinterpret_cast<unsigned long long>(&Mp.m[l])), "r"(tx - 16), "r"(ty - 3), "r"(f),
            "r"(bar)
          : "memory");
    }
    // one thread polls the mbarrier, the other 255 park at the CTA barrier: a polling loop in every warp burns issue
    // slots that the kernels of the other batches in flight could use
    if (threadIdx.x == 0) {
      const unsigned bar = smem_u32(&mbar);
      unsigned done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar)
            : "memory");
      }
    }
    __syncthreads();
    if (rim) {
      // reflect-101 halo outside the level (TMA wrote zeros there).  Only the cells that can reach a valid output are
      // patched: up to 3 rows above / below the level and up to 3 columns left / right of it; warp = row, lane = column.
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      auto refl = [](int g, int n) {
        g = g < 0 ? -g : (g >= n ? 2 * (n - 1) - g : g);
        return max(0, min(g, n - 1));
      };
      const int xxEnd = min(BT_W + 6, w + 3 - (tx - 3));  // columns beyond w + 2 feed no valid output
      for (int yy = warp; yy < BOX_H; yy += 8) {
        const int gy = ty + yy - 3;
        if (gy >= h + 3) break;
        const uint8_t* row = S + (size_t)refl(gy, h) * sp;
        if (gy < 0 || gy >= h) {  // a whole row outside the level
          for (int xx = lane; xx < xxEnd; xx += 32) raw[yy][BOX_XOFF + xx] = row[refl(tx + xx - 3, w)];
        } else {                  // inside rows: only the columns left of 0 and right of w - 1
          if (tx < 3 && lane < 3 - tx) raw[yy][BOX_XOFF + lane] = row[refl(tx + lane - 3, w)];
          const int x0 = w - (tx - 3);  // first box column at or beyond w
          if (x0 < BT_W + 6 && lane < 3 && x0 + lane < BT_W + 6 && x0 + lane >= 0)
            raw[yy][BOX_XOFF + x0 + lane] = row[refl(w + lane, w)];
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < BOX_H * (BT_W + 6); i += 256) {
      const int yy = i / (BT_W + 6), xx = i - yy * (BT_W + 6);
      int gx = tx + xx - 3, gy = ty + yy - 3;
      gx = gx < 0 ? -gx : (gx >= w ? 2 * (w - 1) - gx : gx);
      gy = gy < 0 ? -gy : (gy >= h ? 2 * (h - 1) - gy : gy);
      gx = max(0, min(gx, w - 1));
      gy = max(0, min(gy, h - 1));
      raw[yy][BOX_XOFF + xx] = S[(size_t)gy * sp + gx];
    }
  }
  __syncthreads();
  // horizontal pass: item = (row, 4 consecutive outputs); byte windows by funnel shift, taps by dp4a
  const unsigned k0123 = (unsigned)P.blurk[0] | ((unsigned)P.blurk[1] << 8) | ((unsigned)P.blurk[2] << 16) | ((unsigned)P.blurk[3] << 24);
  const unsigned k456 = (unsigned)P.blurk[4] | ((unsigned)P.blurk[5] << 8) | ((unsigned)P.blurk[6] << 16);
  for (int i = threadIdx.x; i < BOX_H * (BT_W / 4); i += 256) {
    const int r = i / (BT_W / 4), x4 = (i - r * (BT_W / 4)) * 4;
    // output x4+j reads raw bytes [BOX_XOFF + x4 + j, +7): BOX_XOFF = 12 + 1, so word 3 + x4/4 shifted by 1 + j bytes
    const unsigned* wp = reinterpret_cast<const unsigned*>(&raw[r][x4 + BOX_XOFF - 1]);
    const unsigned w0 = wp[0], w1 = wp[1], w2 = wp[2];
    unsigned o[4];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const unsigned lo = __funnelshift_r(w0, w1, 8 * (j + 1));
      const unsigned hi = __funnelshift_r(w1, w2, 8 * (j + 1));
      o[j] = __dp4a(hi, k456, __dp4a(lo, k0123, 0u));
    }
    o[3] = __dp4a(w2, k456, __dp4a(w1, k0123, 0u));
    uint2 pk;
    pk.x = o[0] | (o[1] << 16);
    pk.y = o[2] | (o[3] << 16);
    *reinterpret_cast<uint2*>(&hs[r][x4]) = pk;
  }
  __syncthreads();
  // vertical pass: thread = (column, 8-row half), 14-value register window
  {
    const int c = threadIdx.x & (BT_W - 1), half = threadIdx.x >> 7;
    const int gx = tx + c;
    if (gx < w) {
      int win[14];
#pragma unroll
      for (int i = 0; i < 14; ++i) win[i] = hs[half * 8 + i][c];
      const int k0 = P.blurk[0], k1 = P.blurk[1], k2 = P.blurk[2], k3 = P.blurk[3], k4 = P.blurk[4], k5 = P.blurk[5],
                k6 = P.blurk[6];
      uint8_t* Dst = I.blurred + (size_t)f * P.pyrFrameStride + L.off + gx;
#pragma unroll
      for (int y = 0; y < 8; ++y) {
        const int gy = ty + half * 8 + y;
        if (gy < h) {
          const int a = 32768 + k0 * win[y] + k1 * win[y + 1] + k2 * win[y + 2] + k3 * win[y + 3] + k4 * win[y + 4] +
                        k5 * win[y + 5] + k6 * win[y + 6];
          Dst[(size_t)gy * L.pitch] = (uint8_t)(a >> 16);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// K5  IC_Angle (@0x6fb10) + computeOrbDescriptor (@0x777a8-0x77c72) + keypoint rescale
// (@0x77cb1-0x77d10): one warp per keypoint.  Orientation: lane = column u in [-15,15], loop
// over rows (coalesced 31-byte row reads), integer moments reduced by shuffles, then
// cv::fastAtan2's float polynomial evaluated without FMA contraction.  Descriptor: lane = output
// byte; its 16 pattern points are rotated with the pinned sin/cos (double Cody-Waite + fdlibm
// kernels, every op individually rounded) and the FMA form of the shipped binary (@0x77888-0x778a7).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_orient_desc(const __grid_constant__ OrbParams P, OrbImages I,
                                                     const uint2* __restrict__ lvlKp, const int* __restrict__ lvlCnt,
                                                     plslam_keypoint_t* __restrict__ kps, uint8_t* __restrict__ desc,
                                                     int capacity, int* __restrict__ counts) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.y;
  const int g = blockIdx.x * 8 + warp;
  // level of keypoint g (level-major output order)
  int l = 0, base = 0, total = 0;
  {
    int acc = 0;
    bool found = false;
    for (int i = 0; i < P.nlevels; ++i) {
      const int c = lvlCnt[f * ORB_MAXL + i];
      if (!found && g < acc + c) { l = i; base = acc; found = true; }
      acc += c;
    }
    total = acc;
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[f] = total;
    if (!found) return;
  }
  const OrbLevel& L = P.lv[l];
  const uint2 rec = lvlKp[(size_t)f * P.maxKp + L.kpOff + (g - base)];
  const int X = rec.x & 0xffff, Y = rec.x >> 16;

  // --- orientation on the unblurred level ---
  int sp;
  const uint8_t* S = level_ptr(P, I, f, l, sp);
  const uint8_t* center = S + (size_t)Y * sp + X;
  int m10 = 0, m01 = 0;
  {
    const int u = lane - ORB_HALF_PATCH;  // lanes 0..30
    const int au = abs(u);
    if (lane < 31) {
#pragma unroll 1
      for (int v = -ORB_HALF_PATCH; v <= ORB_HALF_PATCH; ++v) {
        if (au <= P.umax[abs(v)]) {
          const int p = center[v * sp + u];
          m10 += u * p;
          m01 += v * p;
        }
      }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      m10 += __shfl_xor_sync(0xffffffffu, m10, d);
      m01 += __shfl_xor_sync(0xffffffffu, m01, d);
    }
  }
  const float angle = fast_atan2_dev((float)m01, (float)m10);

  // --- descriptor on the blurred level ---
  float sn, cs;
  pl_sincosf_dev(__fmul_rn(angle, 0.017453292f), &sn, &cs);
  const uint8_t* Bc = I.blurred + (size_t)f * P.pyrFrameStride + L.off + (size_t)Y * L.pitch + X;
  const int bp = L.pitch;
  int val = 0;
  // lane's 16 pattern points = 32 signed bytes
  uint32_t pw[8];
  {
    const uint4 a = reinterpret_cast<const uint4*>(g_pattern)[lane * 2];
    const uint4 b = reinterpret_cast<const uint4*>(g_pattern)[lane * 2 + 1];
    pw[0] = a.x; pw[1] = a.y; pw[2] = a.z; pw[3] = a.w;
    pw[4] = b.x; pw[5] = b.y; pw[6] = b.z; pw[7] = b.w;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int tv[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t h = pw[k] >> (16 * j);
      const float px = (float)(int)(int8_t)(h & 0xff), py = (float)(int)(int8_t)((h >> 8) & 0xff);
      const int r = cv_round(__fmaf_rn(px, sn, __fmul_rn(py, cs)));
      const int c = cv_round(__fmaf_rn(px, cs, -__fmul_rn(py, sn)));
      tv[j] = Bc[r * bp + c];
    }
    val |= (tv[0] < tv[1]) << k;
  }
  if (g < capacity) {
    desc[((size_t)f * capacity + g) * 32 + lane] = (uint8_t)val;
    if (lane < 7) {
      float fx = (float)X, fy = (float)Y;
      if (l != 0) { fx = __fmul_rn(fx, L.scale); fy = __fmul_rn(fy, L.scale); }
      uint32_t w;
      switch (lane) {
        case 0: w = __float_as_uint(fx); break;
        case 1: w = __float_as_uint(fy); break;
        case 2: w = __float_as_uint(L.patchSize); break;
        case 3: w = __float_as_uint(angle); break;
        case 4: w = __float_as_uint((float)rec.y); break;
        case 5: w = (uint32_t)l; break;
        default: w = 0xffffffffu; break;
      }
      reinterpret_cast<uint32_t*>(kps + (size_t)f * capacity + g)[lane] = w;
    }
  }
}

inline int cvRoundf_host(float v) { return (int)lrintf(v); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

// [frame][row][byte] view of one level of `batch` frames; false if the layout is not TMA-legal
bool encode_level_map(CUtensorMap* m, const void* base, int w, int h, size_t pitch, size_t frameStride, int batch,
                      int boxW = BOX_W, int boxH = BOX_H) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch & 15) || (frameStride & 15)) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)batch};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frameStride};
  const cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------

// ORBextractor::ORBextractor (ORBextractor.h:51-52, @0x73050)
OrbExtractor::OrbExtractor(int nf, float sf, int nl, int ini, int mn)
    : nfeatures(nf), nlevels(nl), iniThFAST(ini), minThFAST(mn), scaleFactor((double)sf) {
  mvScaleFactor.resize(nl);
  mvLevelSigma2.resize(nl);
  mvInvScaleFactor.resize(nl);
  mvInvLevelSigma2.resize(nl);
  mvScaleFactor[0] = 1.f;
  mvLevelSigma2[0] = 1.f;
  for (int i = 1; i < nl; ++i) {
    mvScaleFactor[i] = (float)((double)mvScaleFactor[i - 1] * scaleFactor);
    mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
  }
  for (int i = 0; i < nl; ++i) {
    mvInvScaleFactor[i] = 1.f / mvScaleFactor[i];
    mvInvLevelSigma2[i] = 1.f / mvLevelSigma2[i];
  }
  mnFeaturesPerLevel.resize(nl);
  const float factor = (float)(1.0 / scaleFactor);
  float nDesired = (float)nf * (1.f - factor) / (1.f - (float)std::pow((double)factor, (double)nl));
  int sum = 0;
  for (int l = 0; l < nl - 1; ++l) {
    mnFeaturesPerLevel[l] = cvRoundf_host(nDesired);
    sum += mnFeaturesPerLevel[l];
    nDesired *= factor;
  }
  mnFeaturesPerLevel[nl - 1] = std::max(nf - sum, 0);
  umax.assign(ORB_HALF_PATCH + 1, 0);
  const int vmax = (int)std::floor(ORB_HALF_PATCH * std::sqrt(2.f) / 2 + 1);
  const int vmin = (int)std::ceil(ORB_HALF_PATCH * std::sqrt(2.f) / 2);
  const double hp2 = ORB_HALF_PATCH * ORB_HALF_PATCH;
  for (int v = 0; v <= vmax; ++v) umax[v] = (int)lrint(std::sqrt(hp2 - v * v));
  for (int v = ORB_HALF_PATCH, v0 = 0; v >= vmin; --v) {
    while (umax[v0] == umax[v0 + 1]) ++v0;
    umax[v] = v0;
    ++v0;
  }
}

OrbExtractor::~OrbExtractor() {
  DevBuf* all[] = {&pyr, &blurred, &coef, &cand, &candCount, &knode, &lvlKp, &lvlCnt, &status, &tileTab,
                   &stageIn, &stageKps, &stageDesc, &stageCnt};
  for (DevBuf* b : all) b->release();
  if (ownStream) cudaStreamDestroy(ownStream);
  if (pinnedStatus) cudaFreeHost(pinnedStatus);
}

int OrbExtractor::set_blur_kernel(const int32_t k[7]) {
  int s = 0;
  for (int i = 0; i < 7; ++i) s += k[i];
  PL_CHECK_ARG(s == 256);
  for (int i = 0; i < 7; ++i) { blurk[i] = k[i]; P.blurk[i] = k[i]; }
  return PLSLAM_OK;
}

int OrbExtractor::max_keypoints() const {
  int s = 0;
  for (int q : mnFeaturesPerLevel) s += q + 3;
  return s;
}

int OrbExtractor::level_size(int level, int* w, int* h) const {
  PL_CHECK_ARG(level >= 0 && level < nlevels && cfgW > 0);
  *w = P.lv[level].w;
  *h = P.lv[level].h;
  return PLSLAM_OK;
}

// Level geometry, FAST cell grid, quad-tree roots, resize tables, workspace (per image size / batch).
int OrbExtractor::configure(int W, int H, int batch) {
  if (W == cfgW && H == cfgH && batch <= cfgB) return PLSLAM_OK;
  PL_CHECK_ARG(nlevels >= 1 && nlevels <= ORB_MAXL);
  PL_CHECK_ARG(W <= 16000 && H <= 16000);
  if (device < 0) {
    PL_CUDA(cudaGetDevice(&device));
    int8_t pat[1024];
    for (int i = 0; i < 1024; ++i) pat[i] = (int8_t)h_pattern[i];
    PL_CUDA(cudaMemcpyToSymbol(g_pattern, pat, sizeof(pat)));
    PL_CUDA(cudaStreamCreateWithFlags(&ownStream, cudaStreamNonBlocking));
    PL_CUDA(cudaMallocHost(&pinnedStatus, 64));
  }
  std::memset(&P, 0, sizeof(P));
  P.nlevels = nlevels;
  P.iniTh = iniThFAST;
  P.minTh = minThFAST;
  for (int i = 0; i < 16; ++i) P.umax[i] = umax[i];
  for (int i = 0; i < 7; ++i) P.blurk[i] = blurk[i];
  size_t off = 0, candOff = 0, coefOff = 0;
  int cellBase = 0, tileBase = 0, kpOff = 0, maxQuota = 0, maxPw = 7, maxPh = 7;
  std::vector<int> coefHost;
  for (int l = 0; l < nlevels; ++l) {
    OrbLevel& L = P.lv[l];
    // ComputePyramid sizes (@0x7051e-0x705ba)
    L.w = cvRoundf_host((float)W * mvInvScaleFactor[l]);
    L.h = cvRoundf_host((float)H * mvInvScaleFactor[l]);
    PL_CHECK_ARG(L.w >= 2 * ORB_EDGE + 1 && L.h >= 2 * ORB_EDGE + 1);
    L.pitch = (int)align_up(L.w, 32);
    L.off = off;
    off += align_up((size_t)L.pitch * L.h, 256);
    // FAST grid (@0x760c6-0x76196)
    L.maxBorderX = L.w - ORB_EDGE + 3;
    L.maxBorderY = L.h - ORB_EDGE + 3;
    const float width = (float)(L.maxBorderX - ORB_MINB), height = (float)(L.maxBorderY - ORB_MINB);
    L.nCols = (int)(width / 30.f);
    L.nRows = (int)(height / 30.f);
    if (L.nCols > 0 && L.nRows > 0) {
      L.wCell = (int)std::ceil(width / L.nCols);
      L.hCell = (int)std::ceil(height / L.nRows);
    } else {
      L.nCols = L.nRows = 0;
      L.wCell = L.hCell = 1;
    }
    // a level with a single column (row) of cells has cells of up to 59 px; the list-order key is mixed-radix over
    // (cell, y, x) and must fit its 28 bits, the staged tile its shared memory
    PL_CHECK_ARG(L.wCell + 6 <= 128 && L.hCell + 6 <= 128);
    PL_CHECK_ARG((unsigned long long)std::max(L.nCols * L.nRows, 1) * (L.wCell + 6) * (L.hCell + 6) < (1ull << 28));
    maxPw = std::max(maxPw, L.wCell + 6);
    maxPh = std::max(maxPh, L.hCell + 6);
    L.cellBase = cellBase;
    cellBase += L.nCols * L.nRows;
    // DistributeOctTree roots (@0x73cca-0x73d49)
    L.quota = mnFeaturesPerLevel[l];
    maxQuota = std::max(maxQuota, L.quota);
    L.nIni = (int)std::round((float)(L.maxBorderX - ORB_MINB) / (L.maxBorderY - ORB_MINB));
    PL_CHECK_ARG(L.nIni >= 1);  // the reference divides by zero for portrait images narrower than h/2
    L.hX = (float)(L.maxBorderX - ORB_MINB) / L.nIni;
    // strict 3x3 local maxima cannot be 8-adjacent: at most one per 2x2 block
    L.candCap = ((L.w + 1) / 2) * ((L.h + 1) / 2);
    L.candOff = candOff;
    candOff += align_up((size_t)L.candCap, 32);
    L.kpOff = kpOff;
    kpOff += L.quota + 3;
    L.tilesX = div_up(L.w, BT_W);
    L.tileBase = tileBase;
    tileBase += L.tilesX * div_up(L.h, BT_H);
    L.scale = mvScaleFactor[l];
    L.patchSize = (float)(int)(31.f * mvScaleFactor[l]);
    if (l > 0) {
      // resize tables: xofs, xa, yofs, ya
      L.coefOff = coefOff;
      const OrbLevel& S = P.lv[l - 1];
      // layout: per destination column an int2 {source offset, c0 | c1 << 16}, then the same per destination row; each table
      // padded to a multiple of 4 entries so that four consecutive entries are two aligned 16-byte loads
      auto emit = [&](int ssize, int dsize) {
        const int dpad = (dsize + 3) & ~3;
        std::vector<int> tab(2 * (size_t)dpad, 0);
        const double scale = 1.0 / ((double)dsize / ssize);
        for (int d = 0; d < dsize; ++d) {
          float fx = (float)((d + 0.5) * scale - 0.5);
          int s = (int)std::floor(fx);
          fx -= s;
          if (s < 0) { s = 0; fx = 0.f; }
          if (s >= ssize - 1) { s = ssize - 1; fx = 0.f; }
          const int c0 = cvRoundf_host((1.f - fx) * 2048.f), c1 = cvRoundf_host(fx * 2048.f);
          tab[2 * d] = s;
          tab[2 * d + 1] = (c0 & 0xffff) | (c1 << 16);
        }
        for (int d = dsize; d < dpad; ++d) {  // padding repeats the last entry (read by the lanes past the row end, never stored)
          tab[2 * d] = tab[2 * (dsize - 1)];
          tab[2 * d + 1] = tab[2 * (dsize - 1) + 1];
        }
        coefHost.insert(coefHost.end(), tab.begin(), tab.end());
      };
      emit(S.w, L.w);
      emit(S.h, L.h);
      coefOff = coefHost.size();
    }
  }
  P.totalCells = cellBase;
  P.totalTiles = tileBase;
  {  // blur tile table: one word per CTA instead of a level search and two integer divisions per thread
    std::vector<unsigned> tab;
    tab.reserve(tileBase);
    for (int l = 0; l < nlevels; ++l) {
      const int ty_n = div_up(P.lv[l].h, BT_H);
      for (int ty = 0; ty < ty_n; ++ty)
        for (int tx = 0; tx < P.lv[l].tilesX; ++tx) tab.push_back(((unsigned)l << 28) | ((unsigned)ty << 14) | (unsigned)tx);
    }
    int rct = tileTab.ensure(tab.size() * sizeof(unsigned));
    if (rct) return rct;
    PL_CUDA(cudaMemcpy(tileTab.p, tab.data(), tab.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
  }
  P.maxKp = kpOff;
  P.nodeCap = std::max(maxQuota + 8, 16);
  P.patchPitch = (int)align_up(maxPw + 3, 4);  // + 3: room for the alignment shift of the staged tile
  P.patchRows = maxPh;
  P.pyrFrameStride = off;
  P.candFrameStride = candOff;
  PL_CHECK_ARG(P.nodeCap < 65535);

  int rc;
  const int B = std::max(batch, cfgB);
  if ((rc = pyr.ensure(off * B))) return rc;
  if ((rc = blurred.ensure(off * B))) return rc;
  if ((rc = coef.ensure(std::max<size_t>(coefHost.size(), 1) * sizeof(int)))) return rc;
  if (!coefHost.empty())
    PL_CUDA(cudaMemcpy(coef.p, coefHost.data(), coefHost.size() * sizeof(int), cudaMemcpyHostToDevice));
  if ((rc = cand.ensure(candOff * B * sizeof(unsigned long long)))) return rc;
  if ((rc = knode.ensure(candOff * B * sizeof(unsigned short)))) return rc;
  if ((rc = candCount.ensure((size_t)B * ORB_MAXL * sizeof(int)))) return rc;
  if ((rc = lvlCnt.ensure((size_t)B * ORB_MAXL * sizeof(int)))) return rc;
  if ((rc = lvlKp.ensure((size_t)B * P.maxKp * sizeof(uint2)))) return rc;
  if ((rc = status.ensure(sizeof(int)))) return rc;
  PL_CUDA(cudaMemset(status.p, 0, sizeof(int)));
  cfgW = W;
  cfgH = H;
  cfgB = B;
  return PLSLAM_OK;
}

static size_t quadtree_smem(int NC) {
  // box[2], cnt[2], seq[2], childCnt[4], newPos, ord, pushOff, expOff, divRank, pad, best(u64)
  return (size_t)NC * (2 * sizeof(QtNode) + 2 * 4 + 2 * 4 + 16 + 5 * 4) + 8 + (size_t)NC * 8;
}

int OrbExtractor::extract_device(const uint8_t* d_images, int batch, int W, int H, int pitch, size_t frame_stride,
                                 plslam_keypoint_t* d_kps, uint8_t* d_desc, int capacity, int32_t* d_counts,
                                 cudaStream_t st) {
  PL_CHECK_ARG(d_images && d_kps && d_desc && d_counts);
  PL_CHECK_ARG(batch >= 1 && batch <= 65535 && W > 0 && H > 0 && pitch >= W);
  PL_CHECK_ARG(frame_stride >= (size_t)pitch * (H - 1) + W);
  int rc = configure(W, H, batch);
  if (rc) return rc;
  if (capacity < P.maxKp) {
    set_error("capacity %d < plslam_orb_max_keypoints() = %d", capacity, P.maxKp);
    return PLSLAM_ERR_CAPACITY;
  }
  OrbImages I;
  I.img0 = d_images;
  I.pitch0 = pitch;
  I.stride0 = frame_stride;
  I.pyr = pyr.as<uint8_t>();
  I.blurred = blurred.as<uint8_t>();
  last_img0 = d_images;
  last_pitch0 = pitch;
  last_stride0 = frame_stride;
  last_batch = batch;

  PL_CUDA(cudaMemsetAsync(candCount.p, 0, (size_t)batch * ORB_MAXL * sizeof(int), st));
  // `status` is sticky: kernels only raise it, check_status() reads and re-arms it (several batches may be in flight)
  // The TMA descriptors only depend on the buffers and the frame geometry: they are re-encoded (host only, ~1 us per
  // level) when those change and passed to the kernel by value; four encoded sets are kept (input buffers the caller
  // alternates between).
  const uintptr_t key[6] = {(uintptr_t)d_images, (uintptr_t)pitch, (uintptr_t)frame_stride, (uintptr_t)batch,
                            (uintptr_t)pyr.p, (uintptr_t)(W * 65536 + H)};
  static_assert(sizeof(AllMaps) <= sizeof(mapsCache[0]), "mapsCache entry too small");
  int slotIdx = -1;
  for (int e = 0; e < 4; ++e)
    if (std::memcmp(key, mapsKey[e], sizeof(key)) == 0) slotIdx = e;
  if (slotIdx < 0) {
    slotIdx = mapsNext;
    mapsNext = (mapsNext + 1) & 3;
    AllMaps A;
    std::memset(&A, 0, sizeof(A));
    BlurMaps& M = A.blur;
    for (int l = 0; l < nlevels; ++l) {
      const void* base = l ? (const void*)(pyr.as<uint8_t>() + P.lv[l].off) : (const void*)d_images;
      const size_t lp = l ? (size_t)P.lv[l].pitch : (size_t)pitch, ls = l ? (size_t)P.pyrFrameStride : frame_stride;
      if (encode_level_map(&M.m[l], base, P.lv[l].w, P.lv[l].h, lp, ls, batch)) M.tmaMask |= 1u << l;
      // the same view with the box of k_resize_tma, when level l + 1 can be built from it: the source patch of a tile is
      // at most ceil(127 r) + 2 + 15 columns (box origin rounded down to 16) and ceil(7 r) + 2 rows, r = size ratio
      if (l + 1 < nlevels && (long long)(RS_TW - 1) * P.lv[l].w <= (long long)(RS_BW - 19) * P.lv[l + 1].w &&
          (long long)2 * (RS_TH - 1) * P.lv[l].h <= (long long)(2 * RS_BH - 5) * P.lv[l + 1].h &&
          encode_level_map(&A.rs.m[l], base, P.lv[l].w, P.lv[l].h, lp, ls, batch, RS_BW, RS_BH))
        A.rs.tmaMask |= 1u << (l + 1);
    }
    std::memcpy(mapsCache[slotIdx], &A, sizeof(A));
    std::memcpy(mapsKey[slotIdx], key, sizeof(key));
  }
  const AllMaps& allMaps = *reinterpret_cast<const AllMaps*>(mapsCache[slotIdx]);
  PL_STAGE_BEGIN(timer, "orb_pyramid(7 launches)", st);
  for (int l = 1; l < nlevels; ++l) {
    dim3 grid(div_up(P.lv[l].w, 128), div_up(P.lv[l].h, 8), batch);
    // PLSLAM_RESIZE_TMA=0 forces the plain-load form (measured: 55.6 us per launch against 62.5 us for the TMA form, 63.7 us
    // for the round-1 kernel; 7 launches per 256 frames, i.e. 0.05 ms of an 8.6 ms step: the TMA form stays the default)
    static const bool resizeTma = [] { const char* e = std::getenv("PLSLAM_RESIZE_TMA"); return !e || std::atoi(e) != 0; }();
    if (resizeTma && ((allMaps.rs.tmaMask >> l) & 1u)) {
      PL_CARVEOUT(k_resize_tma);
      k_resize_tma<<<grid, dim3(32, 8), 0, st>>>(P, allMaps.rs, I, l, coef.as<int>());
    } else {
      PL_CARVEOUT(k_resize);
      k_resize<<<grid, dim3(32, 8), 0, st>>>(P, I, l, coef.as<int>());
    }
  }
  PL_STAGE_END(timer, st);
  {
    const size_t smem = 8 * (2 * (size_t)P.patchPitch * P.patchRows + FAST_QUEUE_BYTES);
    static PerDeviceOnce attr;
    if (attr.first()) {
      PL_CUDA(cudaFuncSetAttribute(k_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      PL_CUDA(cudaFuncSetAttribute(k_quadtree, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    PL_STAGE_BEGIN(timer, "orb_fast", st);
    PL_CARVEOUT(k_fast);
    k_fast<<<dim3(div_up(P.totalCells, 8), batch), 256, smem, st>>>(P, I, cand.as<unsigned long long>(),
                                                                    candCount.as<int>(), status.as<int>());
    PL_STAGE_END(timer, st);
  }
  {
    const size_t smem = quadtree_smem(P.nodeCap);
    if (smem > 200 * 1024) {
      set_error("nfeatures too large for the shared-memory quad-tree (%zu B)", smem);
      return PLSLAM_ERR_INVALID;
    }
    PL_STAGE_BEGIN(timer, "orb_quadtree", st);
    PL_CARVEOUT(k_quadtree);
    k_quadtree<<<dim3(nlevels, batch), 256, smem, st>>>(P, cand.as<unsigned long long>(), candCount.as<int>(),
                                                        knode.as<unsigned short>(), lvlKp.as<uint2>(), lvlCnt.as<int>());
    PL_STAGE_END(timer, st);
  }
  PL_STAGE_BEGIN(timer, "orb_blur", st);
  {
    bool k8 = true;
    for (int i = 0; i < 7; ++i) k8 = k8 && blurk[i] >= 0 && blurk[i] <= 255;
    PL_CHECK_ARG(k8);
    const BlurMaps& dMaps = allMaps.blur;
    PL_CARVEOUT(k_blur);
    k_blur<<<dim3(P.totalTiles, batch), 256, 0, st>>>(P, dMaps, I, lvlCnt.as<int>(), tileTab.as<unsigned>());
  }
  PL_STAGE_END(timer, st);
  PL_STAGE_BEGIN(timer, "orb_orient_desc", st);
  PL_CARVEOUT(k_orient_desc);
  k_orient_desc<<<dim3(div_up(P.maxKp, 8), batch), 256, 0, st>>>(P, I, lvlKp.as<uint2>(), lvlCnt.as<int>(), d_kps,
                                                                 d_desc, capacity, d_counts);
  PL_STAGE_END(timer, st);
  PL_CUDA(cudaGetLastError());
  return PLSLAM_OK;
}

int OrbExtractor::check_status(cudaStream_t st) {
  PL_CUDA(cudaMemcpyAsync(pinnedStatus, status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  PL_CUDA(cudaStreamSynchronize(st));
  const int s = *reinterpret_cast<int*>(pinnedStatus);
  if (s != PLSLAM_OK) {
    set_error("device status %d (internal candidate buffer overflow)", s);
    cudaMemsetAsync(status.p, 0, sizeof(int), st);  // re-arm
  }
  return s;
}

int OrbExtractor::extract_host(const uint8_t* images, int batch, int W, int H, int pitch, size_t frame_stride,
                               plslam_keypoint_t* kps, uint8_t* desc, int capacity, int32_t* counts) {
  PL_CHECK_ARG(images && kps && desc && counts && batch >= 1 && W > 0 && H > 0 && pitch >= W);
  int rc = configure(W, H, batch);
  if (rc) return rc;
  const int cap = P.maxKp;
  const size_t dpitch = align_up(W, 32), dstride = dpitch * H;
  if ((rc = stageIn.ensure(dstride * batch))) return rc;
  if ((rc = stageKps.ensure((size_t)batch * cap * sizeof(plslam_keypoint_t)))) return rc;
  if ((rc = stageDesc.ensure((size_t)batch * cap * 32))) return rc;
  if ((rc = stageCnt.ensure((size_t)batch * sizeof(int)))) return rc;
  cudaStream_t st = ownStream;
  if (frame_stride == (size_t)pitch * H) {
    PL_CUDA(cudaMemcpy2DAsync(stageIn.p, dpitch, images, pitch, W, (size_t)H * batch, cudaMemcpyHostToDevice, st));
  } else {
    for (int f = 0; f < batch; ++f)
      PL_CUDA(cudaMemcpy2DAsync(stageIn.as<uint8_t>() + f * dstride, dpitch, images + f * frame_stride, pitch, W, H,
                                cudaMemcpyHostToDevice, st));
  }
  rc = extract_device(stageIn.as<uint8_t>(), batch, W, H, (int)dpitch, dstride, stageKps.as<plslam_keypoint_t>(),
                      stageDesc.as<uint8_t>(), cap, stageCnt.as<int>(), st);
  if (rc) return rc;
  PL_CUDA(cudaMemcpyAsync(counts, stageCnt.p, (size_t)batch * sizeof(int), cudaMemcpyDeviceToHost, st));
  if ((rc = check_status(st))) return rc;
  for (int f = 0; f < batch; ++f) {
    const int n = counts[f];
    if (n > capacity) {
      set_error("frame %d has %d keypoints, capacity %d", f, n, capacity);
      return PLSLAM_ERR_CAPACITY;
    }
  }
  if (capacity == cap) {
    PL_CUDA(cudaMemcpyAsync(kps, stageKps.p, (size_t)batch * cap * sizeof(plslam_keypoint_t), cudaMemcpyDeviceToHost, st));
    PL_CUDA(cudaMemcpyAsync(desc, stageDesc.p, (size_t)batch * cap * 32, cudaMemcpyDeviceToHost, st));
  } else {
    for (int f = 0; f < batch; ++f) {
      const int n = counts[f];
      if (!n) continue;
      PL_CUDA(cudaMemcpyAsync(kps + (size_t)f * capacity, stageKps.as<plslam_keypoint_t>() + (size_t)f * cap,
                              (size_t)n * sizeof(plslam_keypoint_t), cudaMemcpyDeviceToHost, st));
      PL_CUDA(cudaMemcpyAsync(desc + (size_t)f * capacity * 32, stageDesc.as<uint8_t>() + (size_t)f * cap * 32,
                              (size_t)n * 32, cudaMemcpyDeviceToHost, st));
    }
  }
  PL_CUDA(cudaStreamSynchronize(st));
  return PLSLAM_OK;
}

int OrbExtractor::copy_level(int frame, int level, int which, uint8_t* out, size_t out_bytes) {
  PL_CHECK_ARG(cfgW > 0 && frame >= 0 && frame < last_batch && level >= 0 && level < nlevels && out);
  const OrbLevel& L = P.lv[level];
  PL_CHECK_ARG(out_bytes >= (size_t)L.w * L.h);
  const uint8_t* src;
  size_t sp;
  if (which == 0 && level == 0) {
    src = last_img0 + (size_t)frame * last_stride0;
    sp = last_pitch0;
  } else {
    src = (which ? blurred.as<uint8_t>() : pyr.as<uint8_t>()) + (size_t)frame * P.pyrFrameStride + L.off;
    sp = L.pitch;
  }
  PL_CUDA(cudaDeviceSynchronize());
  PL_CUDA(cudaMemcpy2D(out, L.w, src, sp, L.w, L.h, cudaMemcpyDeviceToHost));
  return PLSLAM_OK;
}

int OrbExtractor::copy_candidates(int frame, int level, int32_t* xyr, int capacity, int* n_out) {
  PL_CHECK_ARG(cfgW > 0 && frame >= 0 && frame < last_batch && level >= 0 && level < nlevels && n_out);
  PL_CUDA(cudaDeviceSynchronize());
  int n = 0;
  PL_CUDA(cudaMemcpy(&n, candCount.as<int>() + frame * ORB_MAXL + level, sizeof(int), cudaMemcpyDeviceToHost));
  *n_out = n;
  if (n > capacity) return PLSLAM_ERR_CAPACITY;
  if (n == 0) return PLSLAM_OK;
  PL_CHECK_ARG(xyr);
  std::vector<unsigned long long> rec(n);
  PL_CUDA(cudaMemcpy(rec.data(), cand.as<unsigned long long>() + (size_t)frame * P.candFrameStride + P.lv[level].candOff,
                     (size_t)n * 8, cudaMemcpyDeviceToHost));
  std::sort(rec.begin(), rec.end(), [](unsigned long long a, unsigned long long b) {
    return ((a >> 28) & 0x0fffffffull) < ((b >> 28) & 0x0fffffffull);
  });
  for (int i = 0; i < n; ++i) {
    xyr[3 * i] = (int)(rec[i] & 0x3fff);
    xyr[3 * i + 1] = (int)((rec[i] >> 14) & 0x3fff);
    xyr[3 * i + 2] = (int)(rec[i] >> 56) - 1;
  }
  return PLSLAM_OK;
}

}  // namespace plslam
