// common.cuh — shared helpers for the sm_100a kernels and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/plslam_b200.h"

namespace plslam {

// thread-local error text returned by plslam_last_error()
void set_error(const char* fmt, ...);

#define PL_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::plslam::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PLSLAM_ERR_CUDA;                                                                 \
    }                                                                                         \
  } while (0)

#define PL_CHECK_ARG(cond)                                                    \
  do {                                                                        \
    if (!(cond)) {                                                            \
      ::plslam::set_error("invalid argument: %s (%s:%d)", #cond, __FILE__, __LINE__); \
      return PLSLAM_ERR_INVALID;                                              \
    }                                                                         \
  } while (0)

// Every kernel of the library asks for the same L1/shared-memory split.  Kernels whose carve-outs differ cannot be
// resident on one SM at the same time, and the front-end relies on short streaming kernels of one batch running
// beside the long region-growing kernels of other batches.  PLSLAM_CARVEOUT (percent of shared memory, -1 = leave
// the driver's per-kernel choice) overrides the default.
int carveout_pct();
// cudaFuncSetAttribute applies to the CURRENT device only: a flag per call site AND device (thread-safe), so that a second
// GPU used from the same process gets its opt-in to large dynamic shared memory and its carve-out preference too.
struct PerDeviceOnce {
  std::atomic<unsigned long long> done[2];
  PerDeviceOnce() { done[0] = 0; done[1] = 0; }
  bool first() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) return true;
    d &= 127;
    const unsigned long long bit = 1ull << (d & 63);
    return !(done[d >> 6].fetch_or(bit) & bit);
  }
};
#define PL_CARVEOUT(kernel)                                                                                   \
  do {                                                                                                        \
    static ::plslam::PerDeviceOnce _pl_once;                                                                  \
    if (_pl_once.first() && ::plslam::carveout_pct() >= 0)                                                    \
      cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, ::plslam::carveout_pct()); \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int div_up(int a, int b) { return (a + b - 1) / b; }

// Simple owning device buffer that only grows.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return PLSLAM_OK;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    PL_CUDA(cudaMalloc(&p, need));
    bytes = need;
    return PLSLAM_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// Optional per-stage device timing: CUDA events recorded on the launching stream around each kernel.
struct StageTimer {
  struct Rec {
    const char* name;
    cudaEvent_t a, b;
  };
  bool enabled = false;
  std::vector<Rec> recs;
  size_t used = 0;
  void reset() { used = 0; }
  void begin(const char* name, cudaStream_t st) {
    if (!enabled) return;
    if (used == recs.size()) {
      Rec r{name, nullptr, nullptr};
      cudaEventCreate(&r.a);
      cudaEventCreate(&r.b);
      recs.push_back(r);
    }
    recs[used].name = name;
    cudaEventRecord(recs[used].a, st);
  }
  void end(cudaStream_t st) {
    if (!enabled) return;
    cudaEventRecord(recs[used].b, st);
    ++used;
  }
  int collect(const char** names, float* ms, int cap) {
    int n = 0;
    for (size_t i = 0; i < used && n < cap; ++i, ++n) {
      cudaEventSynchronize(recs[i].b);
      float t = 0;
      cudaEventElapsedTime(&t, recs[i].a, recs[i].b);
      names[n] = recs[i].name;
      ms[n] = t;
    }
    return n;
  }
  ~StageTimer() {
    for (auto& r : recs) {
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
  }
};
// NVTX range per stage (tracing) when PLSLAM_NVTX=1: the ranges bracket the ENQUEUE of a stage's kernels on
// the host thread (header-only nvtx3; a no-op unless a tool such as nsys is attached).
inline bool nvtx_on() {
  static const bool on = [] { const char* e = std::getenv("PLSLAM_NVTX"); return e && e[0] == '1'; }();
  return on;
}
#define PL_STAGE_BEGIN(tm, name, st) do { if (::plslam::nvtx_on()) nvtxRangePushA(name); if (tm) (tm)->begin(name, st); } while (0)
#define PL_STAGE_END(tm, st) do { if (tm) (tm)->end(st); if (::plslam::nvtx_on()) nvtxRangePop(); } while (0)

#ifdef __CUDACC__
// cvRound on float: round-half-even (x86 vcvtss2si in the reference binary)
__device__ __forceinline__ int cv_round(float v) { return __float2int_rn(v); }

// block-wide exclusive scan of a shared-memory int array, in place; returns the total.
// All threads of the block must call it; n may be any size. `warp_tmp` needs 33 ints of smem.
__device__ inline int block_scan_excl(int* a, int n, int* warp_tmp) {
  const int T = blockDim.x, t = threadIdx.x, lane = t & 31, w = t >> 5, nw = (T + 31) >> 5;
  const int per = (n + T - 1) / T;
  const int b = min(t * per, n), e = min(b + per, n);
  int sum = 0;
  for (int i = b; i < e; ++i) sum += a[i];
  int incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) warp_tmp[w] = incl;
  __syncthreads();
  if (w == 0) {
    int v = lane < nw ? warp_tmp[lane] : 0, iv = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, iv, d);
      if (lane >= d) iv += u;
    }
    if (lane < nw) warp_tmp[lane] = iv - v;
    if (lane == 31) warp_tmp[32] = iv;
  }
  __syncthreads();
  int run = warp_tmp[w] + incl - sum;
  for (int i = b; i < e; ++i) {
    int v = a[i];
    a[i] = run;
    run += v;
  }
  int total = warp_tmp[32];
  __syncthreads();
  return total;
}
#endif

}  // namespace plslam
