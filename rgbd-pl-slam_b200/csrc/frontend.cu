// frontend.cu — batched front-end: ORB + LSD/LBD extraction (+ frame-pair kNN matching) for a batch of
// frames, i.e. the hot path Frame::Frame runs per RGB-D frame (reference include/Frame.h:60,67,70).
// ORB and line stages are independent, so they run on two internal streams forked from the caller's
// stream: the latency-bound line kernels (one warp per frame) overlap the throughput-bound ORB kernels.
#include <string>
#include <new>

#include <chrono>
#include <thread>
#include <cstdlib>

#include "common.cuh"
#include "lines.cuh"
#include "orb.cuh"

namespace plslam {
namespace {

// PLSLAM_TRACE_HOST=1: report enqueue calls that block the submitting thread for more than a millisecond
inline bool trace_host() {
  static const bool on = std::getenv("PLSLAM_TRACE_HOST") != nullptr;
  return on;
}

__global__ void k_make_pair_jobs(const uint8_t* desc, const int32_t* counts, int capacity, int npairs, int32_t* out,
                                 plslam_knn_job_t* jobs) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  plslam_knn_job_t j;
  j.query = desc + (size_t)(2 * p) * capacity * 32;
  j.train = desc + (size_t)(2 * p + 1) * capacity * 32;
  j.out = out + (size_t)p * capacity * 4;
  j.nq = min(counts[2 * p], capacity);
  j.nt = min(counts[2 * p + 1], capacity);
  jobs[p] = j;
}

}  // namespace

// One pipeline slot = one complete workspace (ORB + line extractors, streams, staging).  Independent
// batches submitted on different slots overlap on the GPU: k_lsd_grow keeps one warp per frame busy for
// tens of milliseconds at ~0.13 IPC, so several batches in flight are needed to fill the issue slots.
// experiment switch PLSLAM_STREAM_PRIO = three digits "OHC": 1 = greatest stream priority for the ORB branch stream (O), the
// host-path / line branch stream (H), the copy streams (C); default "000"
static int stream_prio_digit(int i) {
  static const std::string v = [] { const char* e = std::getenv("PLSLAM_STREAM_PRIO"); return std::string(e ? e : "000"); }();
  return (int)v.size() > i && v[i] == '1';
}

struct Slot {
  OrbExtractor orb;
  LineExtractor lines;
  StageTimer tOrb, tLines;
  cudaStream_t sOrb = nullptr, sHost = nullptr;  // the line branch runs on the caller's stream, the ORB branch beside it
  cudaEvent_t evFork = nullptr, evOrb = nullptr, evLines = nullptr, evDone = nullptr;
  // host path, optional (PLSLAM_UPLOAD_STREAM=1): two input staging buffers and a shared upload stream, so the upload of a
  // slot's next batch runs while its current batch computes
  cudaStream_t sUp = nullptr;  // owned by the Frontend
  cudaEvent_t evUp[2] = {nullptr, nullptr}, evFree[2] = {nullptr, nullptr};
  bool bufUsed[2] = {false, false};
  int bufNext = 0;

  DevBuf jobsOrb, jobsLines;
  DevBuf dIn[2], dKps, dDesc, dKpCnt, dKl, dLdesc, dFuncs, dLCnt, dOrbM, dLineM;
  int* pinnedStatus = nullptr;
  bool used = false, hostPending = false, h2dOnce = false;
  Slot(int nf, float sf, int nl, int ini, int mn, int max_lines) : orb(nf, sf, nl, ini, mn) { lines.set_max_lines(max_lines); }
  ~Slot() {
    DevBuf* all[] = {&jobsOrb, &jobsLines, &dIn[0], &dIn[1], &dKps, &dDesc, &dKpCnt, &dKl, &dLdesc, &dFuncs, &dLCnt, &dOrbM, &dLineM};
    for (DevBuf* b : all) b->release();
    for (OutSet& o : wOut) {
      DevBuf* w[] = {&o.kps, &o.desc, &o.kpCnt, &o.kl, &o.ldesc, &o.funcs, &o.lCnt, &o.orbM, &o.lineM};
      for (DevBuf* b : w) b->release();
    }
    if (sOrb) cudaStreamDestroy(sOrb);
    if (sHost) cudaStreamDestroy(sHost);
    cudaEvent_t evs[] = {evFork, evOrb, evLines, evDone, evUp[0], evUp[1], evFree[0], evFree[1], evRead[0], evRead[1], evDl[0], evDl[1]};
    for (cudaEvent_t e : evs)
      if (e) cudaEventDestroy(e);
    if (evUpDone) cudaEventDestroy(evUpDone);
    if (evDownDone) cudaEventDestroy(evDownDone);
    if (pinnedStatus) cudaFreeHost(pinnedStatus);
  }
  int init() {
    if (sOrb) return PLSLAM_OK;
    {
      int least = 0, greatest = 0;
      PL_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      PL_CUDA(cudaStreamCreateWithPriority(&sOrb, cudaStreamNonBlocking, stream_prio_digit(0) ? greatest : least));
      PL_CUDA(cudaStreamCreateWithPriority(&sHost, cudaStreamNonBlocking, stream_prio_digit(1) ? greatest : least));
    }
    PL_CUDA(cudaEventCreateWithFlags(&evFork, cudaEventDisableTiming));
    PL_CUDA(cudaEventCreateWithFlags(&evOrb, cudaEventDisableTiming));
    PL_CUDA(cudaEventCreateWithFlags(&evLines, cudaEventDisableTiming));
    PL_CUDA(cudaEventCreateWithFlags(&evDone, cudaEventDisableTiming));
    for (int b = 0; b < 2; ++b) {
      PL_CUDA(cudaEventCreateWithFlags(&evUp[b], cudaEventDisableTiming));
      PL_CUDA(cudaEventCreateWithFlags(&evFree[b], cudaEventDisableTiming));
    }
    PL_CUDA(cudaMallocHost((void**)&pinnedStatus, 64));
    return PLSLAM_OK;
  }
  int process_device(const uint8_t* d_images, int batch, int W, int H, int pitch, size_t stride,
                     const plslam_frontend_io_t& io, int match_pairs, cudaStream_t st, bool timing) {
    int rc = init();
    if (rc) return rc;
    PL_CHECK_ARG(io.keypoints && io.descriptors && io.kp_counts && io.keylines && io.line_descriptors &&
                 io.line_functions && io.line_counts);
    const int kpCap = orb.max_keypoints(), lnCap = lines.out_capacity();
    const int npairs = batch / 2;
    if (match_pairs) {
      PL_CHECK_ARG(io.orb_matches && io.line_matches && npairs >= 1);
      if ((rc = jobsOrb.ensure((size_t)npairs * sizeof(plslam_knn_job_t)))) return rc;
      if ((rc = jobsLines.ensure((size_t)npairs * sizeof(plslam_knn_job_t)))) return rc;
    }
    tOrb.enabled = tLines.enabled = timing;
    tOrb.reset();
    tLines.reset();
    orb.timer = timing ? &tOrb : nullptr;
    lines.timer = timing ? &tLines : nullptr;
    // the slot's previous batch (possibly submitted from another stream) must have drained: its workspace is reused
    if (used) PL_CUDA(cudaStreamWaitEvent(st, evDone, 0));
    // experiment switch PLSLAM_ONE_STREAM: 1 = both branches in the caller's stream, lines first; 2 = ORB first (one hardware
    // connection per slot instead of two: no aliasing of 2 x depth streams on the 32 connections, no concurrency inside a slot)
    static const int oneStream = [] { const char* e = std::getenv("PLSLAM_ONE_STREAM"); return e ? std::atoi(e) : 0; }();
    cudaStream_t sOrb = oneStream ? st : this->sOrb;
    // experiment switch PLSLAM_ORB_AFTER: the ORB branch of a slot starts with the batch (0), once the line branch has reached
    // its region-growing kernel (1: the kernels in front of it do not compete with ORB work), or after that kernel (2)
    // Measured (profiles/r02_sched_sweeps.log): 29.4 k frames/s with 0, 29.7 k with 1 (three runs each), 26.5 k with 2: default 1.
    static const int orbAfter = [] { const char* e = std::getenv("PLSLAM_ORB_AFTER"); return e ? std::atoi(e) : 1; }();
    lines.mark_event = (!oneStream && orbAfter) ? evFork : nullptr;
    lines.mark_where = orbAfter;
    if (!oneStream && !orbAfter) {
      PL_CUDA(cudaEventRecord(evFork, st));
      PL_CUDA(cudaStreamWaitEvent(sOrb, evFork, 0));
    }
    cudaStream_t sLines = st;  // every stream is a hardware connection: two per slot keep deep pipelines from aliasing queues
    // line branch first: its long sequential kernel should start as early as possible
    const auto tA = std::chrono::steady_clock::now();
    if (oneStream == 2) {
      rc = orb.extract_device(d_images, batch, W, H, pitch, stride, io.keypoints, io.descriptors, kpCap, io.kp_counts, sOrb);
      if (rc) return rc;
    }
    rc = lines.extract_device(d_images, batch, W, H, pitch, stride, io.keylines, io.line_descriptors, io.line_functions,
                              lnCap, io.line_counts, sLines);
    if (rc) return rc;
    const auto tB = std::chrono::steady_clock::now();
    if (!oneStream && orbAfter) PL_CUDA(cudaStreamWaitEvent(sOrb, evFork, 0));
    if (oneStream != 2) {
      rc = orb.extract_device(d_images, batch, W, H, pitch, stride, io.keypoints, io.descriptors, kpCap, io.kp_counts, sOrb);
      if (rc) return rc;
    }
    const auto tC = std::chrono::steady_clock::now();
    if (trace_host()) {
      const double a = std::chrono::duration<double, std::milli>(tB - tA).count(), b = std::chrono::duration<double, std::milli>(tC - tB).count();
      if (a > 1.0 || b > 1.0) fprintf(stderr, "[plslam trace] enqueue lines %.2f ms, orb %.2f ms\n", a, b);
    }
    if (match_pairs) {
      PL_STAGE_BEGIN(orb.timer, "match_orb_knn2", sOrb);
      rc = plslam_match_knn2_pairs_device(io.descriptors, io.kp_counts, kpCap, npairs, io.orb_matches,
                                          jobsOrb.as<plslam_knn_job_t>(), sOrb);
      PL_STAGE_END(orb.timer, sOrb);
      if (rc) return rc;
      PL_STAGE_BEGIN(lines.timer, "match_lbd_knn2", sLines);
      rc = plslam_match_knn2_pairs_device(io.line_descriptors, io.line_counts, lnCap, npairs, io.line_matches,
                                          jobsLines.as<plslam_knn_job_t>(), sLines);
      PL_STAGE_END(lines.timer, sLines);
      if (rc) return rc;
    }
    if (!oneStream) {
      PL_CUDA(cudaEventRecord(evOrb, sOrb));
      PL_CUDA(cudaStreamWaitEvent(st, evOrb, 0));
    }
    PL_CUDA(cudaEventRecord(evDone, st));
    used = true;
    return PLSLAM_OK;
  }
  int check_status(cudaStream_t st) {
    int rc = init();
    if (rc) return rc;
    if (!used) return PLSLAM_OK;
    PL_CUDA(cudaMemcpyAsync(pinnedStatus, orb.device_status(), sizeof(int), cudaMemcpyDeviceToHost, st));
    PL_CUDA(cudaMemcpyAsync(pinnedStatus + 1, lines.device_status(), sizeof(int), cudaMemcpyDeviceToHost, st));
    PL_CUDA(cudaStreamSynchronize(st));
    if (pinnedStatus[0] != PLSLAM_OK || pinnedStatus[1] != PLSLAM_OK) {
      set_error("device status orb=%d lines=%d (internal fixed-capacity buffer overflow)", pinnedStatus[0], pinnedStatus[1]);
      cudaMemsetAsync(const_cast<int*>(orb.device_status()), 0, sizeof(int), st);  // re-arm the sticky flags
      cudaMemsetAsync(const_cast<int*>(lines.device_status()), 0, sizeof(int), st);
      return PLSLAM_ERR_OVERFLOW;
    }
    return PLSLAM_OK;
  }
  // enqueue H2D + kernels + D2H on the slot's host stream; returns without waiting
  int submit_host(const uint8_t* images, int batch, int W, int H, int pitch, size_t stride, const plslam_frontend_io_t& io,
                  int match_pairs, bool timing) {
    int rc = init();
    if (rc) return rc;
    PL_CHECK_ARG(images && batch >= 1 && pitch >= W);
    const int kpCap = orb.max_keypoints(), lnCap = lines.out_capacity();
    const int npairs = batch / 2;
    const size_t dpitch = align_up(W, 32), dstride = dpitch * H;
    const int b = sUp ? bufNext : 0;
    if (sUp) bufNext ^= 1;
    if ((rc = dIn[b].ensure(dstride * batch)) || (rc = dKps.ensure((size_t)batch * kpCap * sizeof(plslam_keypoint_t))) ||
        (rc = dDesc.ensure((size_t)batch * kpCap * 32)) || (rc = dKpCnt.ensure((size_t)batch * 4)) ||
        (rc = dKl.ensure((size_t)batch * lnCap * sizeof(plslam_keyline_t))) || (rc = dLdesc.ensure((size_t)batch * lnCap * 32)) ||
        (rc = dFuncs.ensure((size_t)batch * lnCap * 24)) || (rc = dLCnt.ensure((size_t)batch * 4)))
      return rc;
    if (match_pairs) {
      if ((rc = dOrbM.ensure((size_t)std::max(npairs, 1) * kpCap * 16)) || (rc = dLineM.ensure((size_t)std::max(npairs, 1) * lnCap * 16)))
        return rc;
    }
    cudaStream_t st = sHost;
    static const int dbgSkip = [] { const char* e = std::getenv("PLSLAM_DEBUG_SKIP_COPIES"); return e ? std::atoi(e) : 0; }();  // profiling aid: 1 = H2D only on first use, 2 = no D2H
    cudaStream_t up = sUp ? sUp : st;
    if (sUp && bufUsed[b]) PL_CUDA(cudaStreamWaitEvent(up, evFree[b], 0));  // the batch that last read this buffer
    uint8_t* dst = dIn[b].as<uint8_t>();
    if ((dbgSkip & 4) && h2dOnce && bufUsed[b]) {  // re-upload one frame only: is the cost the bytes or the dependency?
      PL_CUDA(cudaMemcpyAsync(dst, images, dstride, cudaMemcpyHostToDevice, up));
    } else if ((dbgSkip & 1) && h2dOnce && bufUsed[b]) {  // the staging buffer still holds the same frames
    } else if (stride == (size_t)pitch * H && (size_t)pitch == dpitch) {
      // densely packed frames on both sides: one linear copy (a 2-D copy of 122 880 rows is descriptor bound)
      PL_CUDA(cudaMemcpyAsync(dst, images, dstride * batch, cudaMemcpyHostToDevice, up));
    } else if (stride == (size_t)pitch * H) {
      PL_CUDA(cudaMemcpy2DAsync(dst, dpitch, images, pitch, W, (size_t)H * batch, cudaMemcpyHostToDevice, up));
    } else {
      for (int f = 0; f < batch; ++f)
        PL_CUDA(cudaMemcpy2DAsync(dst + f * dstride, dpitch, images + f * stride, pitch, W, H, cudaMemcpyHostToDevice, up));
    }
    if (sUp) {
      PL_CUDA(cudaEventRecord(evUp[b], up));
      PL_CUDA(cudaStreamWaitEvent(st, evUp[b], 0));
    }
    bufUsed[b] = true;
    h2dOnce = true;
    plslam_frontend_io_t d{};
    d.keypoints = dKps.as<plslam_keypoint_t>();
    d.descriptors = dDesc.as<uint8_t>();
    d.kp_counts = dKpCnt.as<int32_t>();
    d.keylines = dKl.as<plslam_keyline_t>();
    d.line_descriptors = dLdesc.as<uint8_t>();
    d.line_functions = dFuncs.as<double>();
    d.line_counts = dLCnt.as<int32_t>();
    d.orb_matches = match_pairs ? dOrbM.as<int32_t>() : nullptr;
    d.line_matches = match_pairs ? dLineM.as<int32_t>() : nullptr;
    rc = process_device(dst, batch, W, H, (int)dpitch, dstride, d, match_pairs, st, timing);
    if (rc) return rc;
    // the kernels that read dIn[b] are all enqueued (process_device joins its two branches back into st): the next upload
    // into this buffer waits for this event
    if (sUp) PL_CUDA(cudaEventRecord(evFree[b], st));

    if (!(dbgSkip & 2)) {
    PL_CUDA(cudaMemcpyAsync(io.keypoints, d.keypoints, (size_t)batch * kpCap * sizeof(plslam_keypoint_t), cudaMemcpyDeviceToHost, st));
    PL_CUDA(cudaMemcpyAsync(io.descriptors, d.descriptors, (size_t)batch * kpCap * 32, cudaMemcpyDeviceToHost, st));
    PL_CUDA(cudaMemcpyAsync(io.kp_counts, d.kp_counts, (size_t)batch * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA(cudaMemcpyAsync(io.keylines, d.keylines, (size_t)batch * lnCap * sizeof(plslam_keyline_t), cudaMemcpyDeviceToHost, st));
    PL_CUDA(cudaMemcpyAsync(io.line_descriptors, d.line_descriptors, (size_t)batch * lnCap * 32, cudaMemcpyDeviceToHost, st));
    PL_CUDA(cudaMemcpyAsync(io.line_functions, d.line_functions, (size_t)batch * lnCap * 24, cudaMemcpyDeviceToHost, st));
    PL_CUDA(cudaMemcpyAsync(io.line_counts, d.line_counts, (size_t)batch * 4, cudaMemcpyDeviceToHost, st));
    if (match_pairs) {
      PL_CUDA(cudaMemcpyAsync(io.orb_matches, d.orb_matches, (size_t)npairs * kpCap * 16, cudaMemcpyDeviceToHost, st));
      PL_CUDA(cudaMemcpyAsync(io.line_matches, d.line_matches, (size_t)npairs * lnCap * 16, cudaMemcpyDeviceToHost, st));
    }
    }
    PL_CUDA(cudaEventRecord(evDone, st));
    hostPending = true;
    return PLSLAM_OK;
  }
  // ---- host-scheduled path (plslam_frontend_acquire_slot / _submit_host_slot / _wait_host) ----
  // A copy with dependent kernels behind it in the same stream costs about twice its duration of whole-pipeline time
  // (DESIGN.md section 6).  Here copies run on a stream of their own, kernels on theirs, and the hand-over between the
  // two is done by the host polling events, so no compute stream ever waits on a copy.
  enum HostState { H_IDLE = 0, H_UPLOAD, H_COMPUTE, H_DOWNLOAD };
  HostState hstate = H_IDLE;
  cudaEvent_t evUpDone = nullptr, evDownDone = nullptr;
  struct Task {
    int batch = 0, W = 0, H = 0, match_pairs = 0;
    size_t dpitch = 0, dstride = 0;
    plslam_frontend_io_t io{};
    bool timing = false;
  } task;
  int stage_upload(cudaStream_t sCopy, const uint8_t* images, int batch, int W, int H, int pitch, size_t stride,
                   const plslam_frontend_io_t& io, int match_pairs, bool timing) {
    int rc = init();
    if (rc) return rc;
    PL_CHECK_ARG(images && batch >= 1 && pitch >= W && hstate == H_IDLE);
    if (!evUpDone) {
      PL_CUDA(cudaEventCreateWithFlags(&evUpDone, cudaEventDisableTiming));
      PL_CUDA(cudaEventCreateWithFlags(&evDownDone, cudaEventDisableTiming));
    }
    const int kpCap = orb.max_keypoints(), lnCap = lines.out_capacity();
    const int npairs = batch / 2;
    const size_t dpitch = align_up(W, 32), dstride = dpitch * H;
    if ((rc = dIn[0].ensure(dstride * batch)) || (rc = dKps.ensure((size_t)batch * kpCap * sizeof(plslam_keypoint_t))) ||
        (rc = dDesc.ensure((size_t)batch * kpCap * 32)) || (rc = dKpCnt.ensure((size_t)batch * 4)) ||
        (rc = dKl.ensure((size_t)batch * lnCap * sizeof(plslam_keyline_t))) || (rc = dLdesc.ensure((size_t)batch * lnCap * 32)) ||
        (rc = dFuncs.ensure((size_t)batch * lnCap * 24)) || (rc = dLCnt.ensure((size_t)batch * 4)))
      return rc;
    if (match_pairs) {
      if ((rc = dOrbM.ensure((size_t)std::max(npairs, 1) * kpCap * 16)) || (rc = dLineM.ensure((size_t)std::max(npairs, 1) * lnCap * 16)))
        return rc;
    }
    uint8_t* dst = dIn[0].as<uint8_t>();
    if (stride == (size_t)pitch * H && (size_t)pitch == dpitch) {
      PL_CUDA(cudaMemcpyAsync(dst, images, dstride * batch, cudaMemcpyHostToDevice, sCopy));
    } else if (stride == (size_t)pitch * H) {
      PL_CUDA(cudaMemcpy2DAsync(dst, dpitch, images, pitch, W, (size_t)H * batch, cudaMemcpyHostToDevice, sCopy));
    } else {
      for (int f = 0; f < batch; ++f)
        PL_CUDA(cudaMemcpy2DAsync(dst + f * dstride, dpitch, images + f * stride, pitch, W, H, cudaMemcpyHostToDevice, sCopy));
    }
    PL_CUDA(cudaEventRecord(evUpDone, sCopy));
    task.batch = batch; task.W = W; task.H = H; task.match_pairs = match_pairs;
    task.dpitch = dpitch; task.dstride = dstride; task.io = io; task.timing = timing;
    hstate = H_UPLOAD;
    hostPending = true;
    return PLSLAM_OK;
  }
  plslam_frontend_io_t device_io() {
    plslam_frontend_io_t d{};
    d.keypoints = dKps.as<plslam_keypoint_t>();
    d.descriptors = dDesc.as<uint8_t>();
    d.kp_counts = dKpCnt.as<int32_t>();
    d.keylines = dKl.as<plslam_keyline_t>();
    d.line_descriptors = dLdesc.as<uint8_t>();
    d.line_functions = dFuncs.as<double>();
    d.line_counts = dLCnt.as<int32_t>();
    d.orb_matches = task.match_pairs ? dOrbM.as<int32_t>() : nullptr;
    d.line_matches = task.match_pairs ? dLineM.as<int32_t>() : nullptr;
    return d;
  }
  // advance this slot's batch by whatever stage has completed; returns a status code
  int pump(cudaStream_t sCopy) {  // sCopy here: the download stream
    if (hstate == H_UPLOAD && cudaEventQuery(evUpDone) == cudaSuccess) {
      const plslam_frontend_io_t d = device_io();
      int rc = process_device(dIn[0].as<uint8_t>(), task.batch, task.W, task.H, (int)task.dpitch, task.dstride, d, task.match_pairs,
                              sHost, task.timing);
      if (rc) return rc;
      hstate = H_COMPUTE;
    }
    if (hstate == H_COMPUTE && cudaEventQuery(evDone) == cudaSuccess) {
      const plslam_frontend_io_t d = device_io();
      const plslam_frontend_io_t& io = task.io;
      const int batch = task.batch, kpCap = orb.max_keypoints(), lnCap = lines.out_capacity(), npairs = batch / 2;
      PL_CUDA(cudaMemcpyAsync(io.keypoints, d.keypoints, (size_t)batch * kpCap * sizeof(plslam_keypoint_t), cudaMemcpyDeviceToHost, sCopy));
      PL_CUDA(cudaMemcpyAsync(io.descriptors, d.descriptors, (size_t)batch * kpCap * 32, cudaMemcpyDeviceToHost, sCopy));
      PL_CUDA(cudaMemcpyAsync(io.kp_counts, d.kp_counts, (size_t)batch * 4, cudaMemcpyDeviceToHost, sCopy));
      PL_CUDA(cudaMemcpyAsync(io.keylines, d.keylines, (size_t)batch * lnCap * sizeof(plslam_keyline_t), cudaMemcpyDeviceToHost, sCopy));
      PL_CUDA(cudaMemcpyAsync(io.line_descriptors, d.line_descriptors, (size_t)batch * lnCap * 32, cudaMemcpyDeviceToHost, sCopy));
      PL_CUDA(cudaMemcpyAsync(io.line_functions, d.line_functions, (size_t)batch * lnCap * 24, cudaMemcpyDeviceToHost, sCopy));
      PL_CUDA(cudaMemcpyAsync(io.line_counts, d.line_counts, (size_t)batch * 4, cudaMemcpyDeviceToHost, sCopy));
      if (task.match_pairs) {
        PL_CUDA(cudaMemcpyAsync(io.orb_matches, d.orb_matches, (size_t)npairs * kpCap * 16, cudaMemcpyDeviceToHost, sCopy));
        PL_CUDA(cudaMemcpyAsync(io.line_matches, d.line_matches, (size_t)npairs * lnCap * 16, cudaMemcpyDeviceToHost, sCopy));
      }
      PL_CUDA(cudaEventRecord(evDownDone, sCopy));
      hstate = H_DOWNLOAD;
    }
    if (hstate == H_DOWNLOAD && cudaEventQuery(evDownDone) == cudaSuccess) hstate = H_IDLE;
    return PLSLAM_OK;
  }
  // ---- wave path (plslam_frontend_submit_host_wave) ----
  // The pipeline is fastest when its slots run in phase (every kernel type gets the whole machine in turn); an upload in
  // front of every step staggers the slots by one copy each and costs ~2 copy durations per step.
  // A wave = up to `depth` batches submitted together: their uploads run back to back on the upload stream into the input
  // buffer set the previous wave is NOT reading, every slot starts after the wave's last upload, results leave on the
  // download stream from the result set the previous wave is NOT using.
  struct OutSet {
    DevBuf kps, desc, kpCnt, kl, ldesc, funcs, lCnt, orbM, lineM;
  };
  OutSet wOut[2];
  cudaEvent_t evRead[2] = {nullptr, nullptr}, evDl[2] = {nullptr, nullptr};
  bool wUsed[2] = {false, false};
  unsigned wUse = 0;  // wave submissions this slot has taken: its parity selects the buffer set
  int wave_upload(int b, cudaStream_t sUpload, const uint8_t* images, int batch, int W, int H, int pitch, size_t stride,
                  int match_pairs) {
    int rc = init();
    if (rc) return rc;
    PL_CHECK_ARG(images && batch >= 1 && pitch >= W);
    if (!evRead[0])
      for (int i = 0; i < 2; ++i) {
        PL_CUDA(cudaEventCreateWithFlags(&evRead[i], cudaEventDisableTiming));
        PL_CUDA(cudaEventCreateWithFlags(&evDl[i], cudaEventDisableTiming));
      }
    const int kpCap = orb.max_keypoints(), lnCap = lines.out_capacity(), npairs = std::max(batch / 2, 1);
    const size_t dpitch = align_up(W, 32), dstride = dpitch * H;
    // both buffer sets at once: an allocation while the previous wave computes would stall the submitting thread
    for (int i = 0; i < 2; ++i) {
      OutSet& o = wOut[i];
      if ((rc = dIn[i].ensure(dstride * batch)) || (rc = o.kps.ensure((size_t)batch * kpCap * sizeof(plslam_keypoint_t))) ||
          (rc = o.desc.ensure((size_t)batch * kpCap * 32)) || (rc = o.kpCnt.ensure((size_t)batch * 4)) ||
          (rc = o.kl.ensure((size_t)batch * lnCap * sizeof(plslam_keyline_t))) || (rc = o.ldesc.ensure((size_t)batch * lnCap * 32)) ||
          (rc = o.funcs.ensure((size_t)batch * lnCap * 24)) || (rc = o.lCnt.ensure((size_t)batch * 4)))
        return rc;
      if (match_pairs && ((rc = o.orbM.ensure((size_t)npairs * kpCap * 16)) || (rc = o.lineM.ensure((size_t)npairs * lnCap * 16)))) return rc;
    }
    if (wUsed[b]) PL_CUDA(cudaStreamWaitEvent(sUpload, evRead[b], 0));  // the buffer's previous reader (two waves ago)
    uint8_t* dst = dIn[b].as<uint8_t>();
    if (stride == (size_t)pitch * H && (size_t)pitch == dpitch) {
      PL_CUDA(cudaMemcpyAsync(dst, images, dstride * batch, cudaMemcpyHostToDevice, sUpload));
    } else if (stride == (size_t)pitch * H) {
      PL_CUDA(cudaMemcpy2DAsync(dst, dpitch, images, pitch, W, (size_t)H * batch, cudaMemcpyHostToDevice, sUpload));
    } else {
      for (int f = 0; f < batch; ++f)
        PL_CUDA(cudaMemcpy2DAsync(dst + f * dstride, dpitch, images + f * stride, pitch, W, H, cudaMemcpyHostToDevice, sUpload));
    }
    return PLSLAM_OK;
  }
  int wave_compute(int b, cudaEvent_t evWaveUp, cudaStream_t sDownload, int batch, int W, int H, const plslam_frontend_io_t& io,
                   int match_pairs, bool timing) {
    const int kpCap = orb.max_keypoints(), lnCap = lines.out_capacity(), npairs = batch / 2;
    const size_t dpitch = align_up(W, 32), dstride = dpitch * H;
    OutSet& o = wOut[b];
    plslam_frontend_io_t d{};
    d.keypoints = o.kps.as<plslam_keypoint_t>();
    d.descriptors = o.desc.as<uint8_t>();
    d.kp_counts = o.kpCnt.as<int32_t>();
    d.keylines = o.kl.as<plslam_keyline_t>();
    d.line_descriptors = o.ldesc.as<uint8_t>();
    d.line_functions = o.funcs.as<double>();
    d.line_counts = o.lCnt.as<int32_t>();
    d.orb_matches = match_pairs ? o.orbM.as<int32_t>() : nullptr;
    d.line_matches = match_pairs ? o.lineM.as<int32_t>() : nullptr;
    PL_CUDA(cudaStreamWaitEvent(sHost, evWaveUp, 0));
    if (wUsed[b]) PL_CUDA(cudaStreamWaitEvent(sHost, evDl[b], 0));  // this result set has left for the host
    int rc = process_device(dIn[b].as<uint8_t>(), batch, W, H, (int)dpitch, dstride, d, match_pairs, sHost, timing);
    if (rc) return rc;
    PL_CUDA(cudaEventRecord(evRead[b], sHost));
    PL_CUDA(cudaStreamWaitEvent(sDownload, evRead[b], 0));
    PL_CUDA(cudaMemcpyAsync(io.keypoints, d.keypoints, (size_t)batch * kpCap * sizeof(plslam_keypoint_t), cudaMemcpyDeviceToHost, sDownload));
    PL_CUDA(cudaMemcpyAsync(io.descriptors, d.descriptors, (size_t)batch * kpCap * 32, cudaMemcpyDeviceToHost, sDownload));
    PL_CUDA(cudaMemcpyAsync(io.kp_counts, d.kp_counts, (size_t)batch * 4, cudaMemcpyDeviceToHost, sDownload));
    PL_CUDA(cudaMemcpyAsync(io.keylines, d.keylines, (size_t)batch * lnCap * sizeof(plslam_keyline_t), cudaMemcpyDeviceToHost, sDownload));
    PL_CUDA(cudaMemcpyAsync(io.line_descriptors, d.line_descriptors, (size_t)batch * lnCap * 32, cudaMemcpyDeviceToHost, sDownload));
    PL_CUDA(cudaMemcpyAsync(io.line_functions, d.line_functions, (size_t)batch * lnCap * 24, cudaMemcpyDeviceToHost, sDownload));
    PL_CUDA(cudaMemcpyAsync(io.line_counts, d.line_counts, (size_t)batch * 4, cudaMemcpyDeviceToHost, sDownload));
    if (match_pairs) {
      PL_CUDA(cudaMemcpyAsync(io.orb_matches, d.orb_matches, (size_t)npairs * kpCap * 16, cudaMemcpyDeviceToHost, sDownload));
      PL_CUDA(cudaMemcpyAsync(io.line_matches, d.line_matches, (size_t)npairs * lnCap * 16, cudaMemcpyDeviceToHost, sDownload));
    }
    PL_CUDA(cudaEventRecord(evDl[b], sDownload));
    wUsed[b] = true;
    hostPending = true;
    return PLSLAM_OK;
  }
  int wait_host() {
    if (!hostPending) return PLSLAM_OK;
    hostPending = false;
    return check_status(sHost);
  }
};

struct Frontend {
  std::vector<Slot*> slots;
  int next = 0, lastSlot = 0;
  bool timing = false;
  cudaStream_t sCopy = nullptr, sDown = nullptr;  // upload / download streams of the host-scheduled path
  int ensure_copy_stream() {
    int least = 0, greatest = 0;
    PL_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    if (!sCopy) PL_CUDA(cudaStreamCreateWithPriority(&sCopy, cudaStreamNonBlocking, stream_prio_digit(2) ? greatest : least));
    if (!sDown) PL_CUDA(cudaStreamCreateWithPriority(&sDown, cudaStreamNonBlocking, stream_prio_digit(2) ? greatest : least));
    if (evWave.empty()) {
      evWave.assign(slots.size(), nullptr);  // one wave takes at least one slot: never more waves in flight than slots
      for (cudaEvent_t& e : evWave) PL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    return PLSLAM_OK;
  }
  std::vector<cudaEvent_t> evWave;
  unsigned waveIdx = 0;
  int waveCursor = 0;  // first slot of the next wave: successive waves rotate over the slots
  int pump_all() {
    int rc = PLSLAM_OK;
    for (Slot* s : slots) {
      const int r = s->pump(sDown);
      if (r && !rc) rc = r;
    }
    return rc;
  }
  // optional upload stream of the host path, shared by the slots (PLSLAM_UPLOAD_STREAM=1; measured: no gain, off by default)
  cudaStream_t sUp = nullptr;
  int ensure_upload_stream() {
    static const bool enabled = [] { const char* e = std::getenv("PLSLAM_UPLOAD_STREAM"); return e && std::atoi(e) != 0; }();
    if (enabled && !sUp && slots.size() > 1) {
      PL_CUDA(cudaStreamCreateWithFlags(&sUp, cudaStreamNonBlocking));
      for (Slot* s : slots) s->sUp = sUp;
    }
    return PLSLAM_OK;
  }

  Frontend(int nf, float sf, int nl, int ini, int mn, int max_lines, int depth) {
    for (int i = 0; i < depth; ++i) {
      slots.push_back(new Slot(nf, sf, nl, ini, mn, max_lines));
      slots.back()->lines.batches_in_flight = depth;
    }
  }
  ~Frontend() {
    for (Slot* s : slots) delete s;
    if (sUp) cudaStreamDestroy(sUp);
    if (sCopy) cudaStreamDestroy(sCopy);
    if (sDown) cudaStreamDestroy(sDown);
    for (cudaEvent_t e : evWave)
      if (e) cudaEventDestroy(e);
  }
  Slot& take() {
    lastSlot = next;
    next = (next + 1) % (int)slots.size();
    return *slots[lastSlot];
  }
};

}  // namespace plslam

using namespace plslam;

struct plslam_frontend {
  Frontend impl;
  plslam_frontend(int nf, float sf, int nl, int ini, int mn, int ml, int depth) : impl(nf, sf, nl, ini, mn, ml, depth) {}
};

extern "C" {

int plslam_match_knn2_pairs_device(const uint8_t* d_desc, const int32_t* d_counts, int capacity, int npairs,
                                   int32_t* d_out, plslam_knn_job_t* d_jobs_scratch, void* stream) {
  PL_CHECK_ARG(d_desc && d_counts && d_out && d_jobs_scratch && capacity >= 1 && npairs >= 1);
  PL_CARVEOUT(k_make_pair_jobs);
  k_make_pair_jobs<<<div_up(npairs, 128), 128, 0, (cudaStream_t)stream>>>(d_desc, d_counts, capacity, npairs, d_out,
                                                                          d_jobs_scratch);
  PL_CUDA(cudaGetLastError());
  return plslam_match_knn2_batch_device(d_jobs_scratch, npairs, capacity, stream);
}

int plslam_frontend_create_pipelined(plslam_frontend_t** out, int nfeatures, float scaleFactor, int nlevels, int iniThFAST,
                                     int minThFAST, int max_lines, int depth) {
  PL_CHECK_ARG(out != nullptr);
  *out = nullptr;
  PL_CHECK_ARG(nfeatures > 0 && nlevels >= 1 && nlevels <= ORB_MAXL && scaleFactor > 1.0f && max_lines >= 0);
  PL_CHECK_ARG(iniThFAST >= minThFAST && minThFAST >= 1 && iniThFAST < 255 && depth >= 1 && depth <= 64);
  *out = new (std::nothrow) plslam_frontend(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_lines, depth);
  if (!*out) {
    set_error("out of host memory");
    return PLSLAM_ERR_INVALID;
  }
  return PLSLAM_OK;
}
int plslam_frontend_create(plslam_frontend_t** out, int nfeatures, float scaleFactor, int nlevels, int iniThFAST,
                           int minThFAST, int max_lines) {
  return plslam_frontend_create_pipelined(out, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_lines, 1);
}
void plslam_frontend_destroy(plslam_frontend_t* h) { delete h; }
int plslam_frontend_depth(const plslam_frontend_t* h) { return h ? (int)h->impl.slots.size() : 0; }
int plslam_frontend_capacities(const plslam_frontend_t* h, int* kp_capacity, int* line_capacity) {
  PL_CHECK_ARG(h);
  if (kp_capacity) *kp_capacity = h->impl.slots[0]->orb.max_keypoints();
  if (line_capacity) *line_capacity = h->impl.slots[0]->lines.out_capacity();
  return PLSLAM_OK;
}
int plslam_frontend_process_device(plslam_frontend_t* h, const uint8_t* d_images, int batch, int width, int height,
                                   int pitch, size_t frame_stride, const plslam_frontend_io_t* io, int match_pairs,
                                   void* stream) {
  PL_CHECK_ARG(h && io && d_images);
  return h->impl.take().process_device(d_images, batch, width, height, pitch, frame_stride, *io, match_pairs,
                                       (cudaStream_t)stream, h->impl.timing);
}
int plslam_frontend_submit_host(plslam_frontend_t* h, const uint8_t* images, int batch, int width, int height, int pitch,
                                size_t frame_stride, const plslam_frontend_io_t* io, int match_pairs) {
  PL_CHECK_ARG(h && io);
  // No host wait here: the slot's streams order this batch behind its previous one (staging buffers, workspaces), and
  // the overflow flags are sticky until plslam_frontend_wait_host() reads them.  Blocking on the slot's previous batch
  // would stall the one submitting thread while other slots sit idle.
  int rcu = h->impl.ensure_upload_stream();
  if (rcu) return rcu;
  Slot& s = h->impl.take();
  return s.submit_host(images, batch, width, height, pitch, frame_stride, *io, match_pairs, h->impl.timing);
}
// Completion-ordered submission: acquire returns a slot whose previous batch has finished (polling the slots' completion
// events), submit_host_slot enqueues on exactly that slot.  Every copy and kernel enqueued this way is ready to run, so the
// copy engines' in-order queues never hold a batch behind another slot's unfinished one.
int plslam_frontend_acquire_slot(plslam_frontend_t* h) {
  if (!h) return -1;
  Frontend& F = h->impl;
  if (F.ensure_copy_stream()) return -1;
  const int n = (int)F.slots.size();
  for (;;) {
    if (F.pump_all()) return -1;
    for (int k = 0; k < n; ++k) {
      const int i = (F.next + k) % n;
      Slot* s = F.slots[i];
      if (s->hstate == Slot::H_IDLE && (!s->used || !s->evDone || cudaEventQuery(s->evDone) == cudaSuccess)) {
        F.next = (i + 1) % n;
        return i;
      }
    }
    std::this_thread::sleep_for(std::chrono::microseconds(20));
  }
}
int plslam_frontend_submit_host_slot(plslam_frontend_t* h, int slot, const uint8_t* images, int batch, int width, int height,
                                     int pitch, size_t frame_stride, const plslam_frontend_io_t* io, int match_pairs) {
  PL_CHECK_ARG(h && io && slot >= 0 && slot < (int)h->impl.slots.size());
  int rc = h->impl.ensure_copy_stream();
  if (rc) return rc;
  h->impl.lastSlot = slot;
  rc = h->impl.slots[slot]->stage_upload(h->impl.sCopy, images, batch, width, height, pitch, frame_stride, *io, match_pairs,
                                         h->impl.timing);
  if (rc) return rc;
  return h->impl.pump_all();
}
// Wave submission: n_batches (<= depth) batches enter the pipeline together and in phase (Slot::wave_*).  Returns once
// everything is enqueued; the uploads of this wave overlap the kernels of the previous one.
int plslam_frontend_submit_host_wave(plslam_frontend_t* h, const uint8_t* const* images, int n_batches, int batch, int width,
                                     int height, int pitch, size_t frame_stride, const plslam_frontend_io_t* ios, int match_pairs) {
  PL_CHECK_ARG(h && images && ios && n_batches >= 1 && n_batches <= (int)h->impl.slots.size());
  Frontend& F = h->impl;
  int rc = F.ensure_copy_stream();
  if (rc) return rc;
  for (int i = 0; i < n_batches; ++i) {
    const plslam_frontend_io_t& io = ios[i];
    PL_CHECK_ARG(io.keypoints && io.descriptors && io.kp_counts && io.keylines && io.line_descriptors && io.line_functions &&
                 io.line_counts && (!match_pairs || (io.orb_matches && io.line_matches)));
  }
  const int ns = (int)F.slots.size();
  cudaEvent_t evUp = F.evWave[F.waveIdx % (unsigned)ns];
  for (int i = 0; i < n_batches; ++i) {
    Slot* s = F.slots[(F.waveCursor + i) % ns];
    if ((rc = s->wave_upload((int)(s->wUse & 1u), F.sCopy, images[i], batch, width, height, pitch, frame_stride, match_pairs))) return rc;
  }
  PL_CUDA(cudaEventRecord(evUp, F.sCopy));
  for (int i = 0; i < n_batches; ++i) {
    Slot* s = F.slots[(F.waveCursor + i) % ns];
    if ((rc = s->wave_compute((int)(s->wUse & 1u), evUp, F.sDown, batch, width, height, ios[i], match_pairs, F.timing))) return rc;
    ++s->wUse;
  }
  ++F.waveIdx;
  F.lastSlot = (F.waveCursor + n_batches - 1) % ns;
  F.waveCursor = (F.waveCursor + n_batches) % ns;
  return PLSLAM_OK;
}
int plslam_frontend_wait_host(plslam_frontend_t* h) {
  PL_CHECK_ARG(h);
  int rc = PLSLAM_OK;
  if (h->impl.sCopy) {  // host-scheduled batches: drive them to completion
    for (;;) {
      rc = h->impl.pump_all();
      if (rc) return rc;
      bool busy = false;
      for (Slot* s : h->impl.slots) busy = busy || s->hstate != Slot::H_IDLE;
      if (!busy) break;
      std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
  }
  if (h->impl.sDown) PL_CUDA(cudaStreamSynchronize(h->impl.sDown));  // results of the wave path
  for (Slot* s : h->impl.slots) {
    const int r = s->wait_host();
    if (r && !rc) rc = r;
  }
  return rc;
}
int plslam_frontend_process_host(plslam_frontend_t* h, const uint8_t* images, int batch, int width, int height,
                                 int pitch, size_t frame_stride, const plslam_frontend_io_t* io, int match_pairs) {
  int rc = plslam_frontend_submit_host(h, images, batch, width, height, pitch, frame_stride, io, match_pairs);
  if (rc) return rc;
  return h->impl.slots[h->impl.lastSlot]->wait_host();
}
int plslam_frontend_check_status(plslam_frontend_t* h, void* stream) {
  PL_CHECK_ARG(h);
  int rc = PLSLAM_OK;
  for (Slot* s : h->impl.slots) {
    const int r = s->check_status((cudaStream_t)stream);
    if (r && !rc) rc = r;
  }
  return rc;
}
int plslam_frontend_enable_timing(plslam_frontend_t* h, int enable) {
  PL_CHECK_ARG(h);
  h->impl.timing = enable != 0;
  return PLSLAM_OK;
}
int plslam_frontend_stage_times(plslam_frontend_t* h, const char** names, float* ms, int capacity) {
  if (!h || !names || !ms) return 0;
  Slot& s = *h->impl.slots[h->impl.lastSlot];
  int n = s.tLines.collect(names, ms, capacity);
  n += s.tOrb.collect(names + n, ms + n, capacity - n);
  return n;
}
int plslam_frontend_launches_per_call(const plslam_frontend_t* h, int match_pairs) {
  if (!h) return 0;
  // ORB: (nlevels-1) resize + fast + quadtree + blur + orient_desc; lines: 10 kernels; matching: 2 x (jobs + knn2)
  return (h->impl.slots[0]->orb.nlevels - 1) + 4 + 10 + (match_pairs ? 4 : 0);
}

}  // extern "C"
