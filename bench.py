#!/usr/bin/env python3
"""bench.py — frames/s of the ORB + LSD/LBD extract + match hot path on N B200s (BASELINE.json metric).

A "step" is one pass of the front-end over one batch of synthetic 640x480 frames (config C2 of
BASELINE.md: batch = 256 frames = 128 frame pairs per GPU): ORB extraction (8-level pyramid, FAST,
quad-tree, orientation, rBRIEF), LSD + LBD line extraction and brute-force Hamming kNN(k=2) matching of
both descriptor sets inside every frame pair.  Frames shard over ranks (weak scaling, no data-path
collective).

  python bench.py [--gpus N] [--steps K] [--warmup W]           our CUDA path (rank 0 prints one JSON line)
  python bench.py --impl reference ...                         the reference algorithm on the host cores
                                                               (oracle restatement: the reference ships no
                                                               buildable source for this path)
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "rgbd-pl-slam_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

# one hardware queue per stream of the pipelined front-end (read when the CUDA context is created; see c_api.cu)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# keep stdout to the one JSON line: NCCL's version banner / debug output goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"  # the VERSION level prints its banner with a bare printf to stdout

import numpy as np

METRIC = "frames/sec ORB+LSD extract+match 640x480 RGB-D"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step (even: frame pairs)")
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--nfeatures", type=int, default=1000)
    ap.add_argument("--max-lines", type=int, default=40)
    ap.add_argument("--depth", type=int, default=32,
                    help="batches (steps) in flight per GPU: pipeline slots of the front-end (measured on B200, 640x480, batch 256: "
                         "16.9 k frames/s at 8, 23.1 k at 16, 23.2 k at 32 before the last kernel changes - profiles/r02h_depth_sweep.log; 28.7 k at 32 "
                         "at the end of round 2; flat from 16 on: 16 batches are what the register file holds of k_lsd_grow; ~3 GB of workspace per slot)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="frames in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="BASELINE.json configs.  c2: 640x480 extract + kNN matching (headline); c3: 1280x720, nFeatures 2000, "
                         "128 frames (64 pairs) per step; c4: c2 plus ComputeBoW + SearchByBoW on every pair; c5: 3840x2160, "
                         "nFeatures 8000, 16 frames per step")
    ap.add_argument("--wave", type=int, default=0,
                    help="batches per plslam_frontend_submit_host_wave call of the e2e leg (0 = depth / 2: successive waves rotate "
                         "over the slots, so only the first wave's upload is exposed; measured at 64 steps, 32 slots: 26.2 k "
                         "frames/s with 32 per call, 27.1 k with 16, 22.7 k with 8, 25.0 k with 4 - profiles/r02_wave_sweep.log)")
    ap.add_argument("--wave-ramp", default="auto",
                    help="comma-separated sizes of the first waves of the e2e leg (then --wave); auto = depth/5 + the rest when the "
                         "run passes over the slots only once (steps <= slots), none otherwise")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-frame latency block")
    ap.add_argument("--voc-levels", type=int, default=6, help="depth L of the synthetic k=10 vocabulary (ORBvoc.txt: 6)")
    return ap.parse_args()


WORKLOAD_PRESETS = {  # (width, height, nfeatures, batch, depth) of BASELINE.json's configs 3 and 5
    # batches in flight measured on B200 (profiles/r02_c35_sweep.log): C3 2.05 k frames/s at 4, 3.6 k at 12, 3.96 k at 24 (110 GB of
    # workspace); C5 36.7 frames/s at 2, 79.8 at 6, 113 at 12, 90 at 20 (beyond ~300 frames in flight region growing switches to
    # its one-warp-per-frame form, which needs far more frames than fit)
    "c3": (1280, 720, 2000, 128, 24),
    "c5": (3840, 2160, 8000, 16, 12),
}


def apply_workload(a):
    if a.workload in WORKLOAD_PRESETS:
        w, h, nf, b, d = WORKLOAD_PRESETS[a.workload]
        defaults = parse_defaults()
        if a.width == defaults.width and a.height == defaults.height:
            a.width, a.height = w, h
        if a.nfeatures == defaults.nfeatures:
            a.nfeatures = nf
        if a.batch == defaults.batch:
            a.batch = b
        if a.depth == defaults.depth:
            a.depth = d
    return a


def parse_defaults():
    return argparse.Namespace(width=640, height=480, nfeatures=1000, batch=256, depth=32)


def workload_name(a):
    base = "%dx%d synthetic RGB-D, batch=%d frames (%d pairs)/GPU, ORB nFeatures=%d 8 levels + LSD/LBD top-%d, kNN2 ORB+LBD per pair" % (
        a.width, a.height, a.batch, a.batch // 2, a.nfeatures, a.max_lines)
    if getattr(a, "workload", "c2") == "c4":
        return "C4: " + base + " + ComputeBoW (k=10 L=%d synthetic vocabulary) + SearchByBoW per pair" % a.voc_levels
    return getattr(a, "workload", "c2").upper() + ": " + base


def broadcast_vocabulary(a, pl, dist, rank, world):
    """Rank 0 builds the vocabulary and exports its flat device image; one NCCL broadcast puts it on every GPU
    (the only collective besides the timing barrier; nothing NCCL on the per-frame path)."""
    import torch
    from plslam_b200.synth import synth_vocabulary_arrays
    info = {}
    voc = None
    nbytes = torch.zeros(1, dtype=torch.int64, device="cuda")
    blob = None
    if rank == 0:
        voc = pl.ORBVocabulary.from_arrays(10, a.voc_levels, *synth_vocabulary_arrays(10, a.voc_levels, seed=0))
        blob = voc.export_blob()
        nbytes[0] = blob.numel()
    if world > 1:
        dist.broadcast(nbytes, 0)
        if rank != 0:
            blob = torch.empty(int(nbytes.item()), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(blob, 0)
        e1.record()
        torch.cuda.synchronize()
        info = {"voc_broadcast_bytes": int(nbytes.item()), "voc_broadcast_ms": e0.elapsed_time(e1)}
        if rank != 0:
            voc = pl.ORBVocabulary.from_blob(blob)
    info["voc_nodes"] = voc.n_nodes
    return voc, info


def make_frames(a, rank):
    from plslam_b200.synth import synth_pair
    frames = np.empty((a.batch, a.height, a.width), np.uint8)
    for p in range(a.batch // 2):
        x, y = synth_pair(rank * 100000 + p, a.width, a.height)
        frames[2 * p], frames[2 * p + 1] = x, y
    return frames


# ------------------------------------------------------------------------------------------------
# algorithmic bytes per frame and stage (every stage reads its input once and
# writes its output once)
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(a, kp_avg, cand_avg):
    W, H = a.width, a.height
    inv = [1.0]
    for _ in range(7):
        inv.append(inv[-1] / 1.2)
    areas = [int(round(W * s)) * int(round(H * s)) for s in inv]
    S = sum(areas)
    P = int(round(W * 0.8)) * int(round(H * 0.8))
    K = kp_avg
    stages = {
        "orb_pyramid(7 launches)": sum(areas[l - 1] + areas[l] for l in range(1, 8)),
        "orb_fast": S + 8 * cand_avg,
        "orb_quadtree": 8 * cand_avg + 8 * K,
        "orb_blur": 2 * S,
        "orb_orient_desc": K * (709 + 4) + K * (512 + 32) + K * 28,
        "match_orb_knn2": (2 * K * 32 + K * 16) // 2,  # per frame (a pair reads 2K rows, writes K results)
        "lsd_scale": 2 * W * H + W * H + P,
        "lsd_grad": P + 16 * P,
        "lsd_rowhist": 8 * P,
        "lsd_colscan": 8 * P,
        "lsd_scatter": 8 * P,
        "lsd_grow": 9 * P,
        "lsd_nfa": 4 * P,
        "lsd_finish": 40 * 650 + a.max_lines * (68 + 24),
        "lbd": W * H + a.max_lines * (60 * 63 * 4 + 32),
        "match_lbd_knn2": (2 * a.max_lines * 32 + a.max_lines * 16) // 2,
    }
    return stages


def survey_bytes_per_frame(a, K):
    """SURVEY.md section 8d's algorithmic bytes of one frame (ORB + LSD + LBD; 21.46 MB at 640x480, K = 1000)."""
    W, H = a.width, a.height
    inv = [1.0]
    for _ in range(7):
        inv.append(inv[-1] / 1.2)
    A = [int(round(W * s)) * int(round(H * s)) for s in inv]
    S = sum(A)
    P = int(np.ceil(0.8 * W)) * int(np.ceil(0.8 * H))
    orb = A[0] + sum(A[l - 1] + A[l] for l in range(1, 8)) + S + 2 * S + K * (709 + 4) + K * (512 + 32) + K * 28
    lsd = 2 * W * H + (W * H + P) + (P + 16 * P + 12 * P) + 24 * P + 9 * P
    lbd = W * H + 4 * W * H + 40 * (60 * 63 * 4 + 32)
    return orb + lsd + lbd


def single_frame_latency(a, frames, n=200):
    """The drop-in call of the reference's Frame constructor: one ORBextractor::operator() + one
    LineSegment::ExtractLineSegment per camera frame, through the C-ABI host entry points (pageable host image in, host
    results out, synchronous), next to the same two calls of the CPU port on one thread."""
    import plslam_b200 as pl
    orb = pl.ORBextractor(a.nfeatures, 1.2, 8, 20, 7)
    ls = pl.LineSegment(a.max_lines)
    for i in range(5):
        orb(frames[i % len(frames)])
        ls.ExtractLineSegment(frames[i % len(frames)])
    t_orb, t_ls = [], []
    for i in range(n):
        img = frames[i % len(frames)]
        t0 = time.perf_counter()
        orb(img)
        t1 = time.perf_counter()
        ls.ExtractLineSegment(img)
        t2 = time.perf_counter()
        t_orb.append((t1 - t0) * 1e3)
        t_ls.append((t2 - t1) * 1e3)
    tot = np.array(t_orb) + np.array(t_ls)
    out = {"frames": n, "api": "plslam_orb_extract + plslam_lines_extract (ORBextractor::operator(), LineSegment::ExtractLineSegment), "
                              "host image in, host results out, one frame per call",
           "p50_ms": float(np.percentile(tot, 50)), "p99_ms": float(np.percentile(tot, 99)),
           "orb_p50_ms": float(np.percentile(t_orb, 50)), "lines_p50_ms": float(np.percentile(t_ls, 50))}
    try:
        from oracle import bindings as ob
        ob.build()
        o = ob.OrbOracle(a.nfeatures, 1.2, 8, 20, 7)
        m = min(6, len(frames))
        t0 = time.perf_counter()
        for i in range(m):
            o.extract(frames[i])
        t1 = time.perf_counter()
        for i in range(m):
            ob.extract_lines(frames[i], a.max_lines)
        t2 = time.perf_counter()
        out["cpu_port_1thread_ms"] = {"orb": (t1 - t0) / m * 1e3, "lines": (t2 - t1) / m * 1e3, "frames": m}
    except Exception as e:  # the oracle is the checker: its absence must not lose the bench line
        out["cpu_port_1thread_ms"] = "unavailable: %s" % e
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML during the timed region."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference algorithm) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_pairs_per_second(a, frames, threads):
    """ORB + LSD/LBD extraction of both frames of each pair + kNN2 matching, frame-pair tasks over `threads` threads."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import bindings as ob
    ob.build()
    tl = threading.local()

    def task(p):
        if not hasattr(tl, "orb"):
            tl.orb = ob.OrbOracle(a.nfeatures, 1.2, 8, 20, 7)
        r = []
        for f in (2 * p, 2 * p + 1):
            k, d = tl.orb.extract(frames[f])
            kl, ld, lf, _ = ob.extract_lines(frames[f], a.max_lines)
            r.append((d, ld))
        ob.knn2(r[0][0], r[1][0])
        ob.knn2(r[0][1], r[1][1])
        return len(r[0][0])

    npairs = len(frames) // 2
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(task, range(npairs)))
    return time.perf_counter() - t0


def opencv_primitives_fps(a, frames, threads):
    """Independent reference point: the OpenCV calls the reference's path is made of, as shipped in cv2
    (SIMD-optimised): 8-level resize pyramid, whole-level FAST at both thresholds' cheaper one, 7x7 blur per level,
    LSD_REFINE_ADV, BFMatcher(HAMMING).knnMatch(k=2) of 1000 x 1000 rows per pair.  No quad-tree, orientation, rBRIEF or
    LBD (cv2 has no ORB-SLAM extractor and no line_descriptor module here), so it is a LOWER bound on the CPU cost."""
    try:
        import cv2
    except Exception:
        return None
    from concurrent.futures import ThreadPoolExecutor
    cv2.setNumThreads(1)
    rng = np.random.default_rng(0)
    d1 = rng.integers(0, 256, (a.nfeatures, 32)).astype(np.uint8)
    d2 = rng.integers(0, 256, (a.nfeatures, 32)).astype(np.uint8)
    tl = threading.local()

    def task(p):
        if not hasattr(tl, "lsd"):
            tl.lsd = cv2.createLineSegmentDetector(cv2.LSD_REFINE_ADV)
            tl.fast = cv2.FastFeatureDetector_create(20, True)
            tl.bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        for f in (2 * p, 2 * p + 1):
            cur = frames[f]
            for l in range(8):
                if l:
                    cur = cv2.resize(cur, (int(round(a.width / 1.2 ** l)), int(round(a.height / 1.2 ** l))), interpolation=cv2.INTER_LINEAR)
                tl.fast.detect(cur)
                cv2.GaussianBlur(cur, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
            tl.lsd.detect(frames[f])
        tl.bf.knnMatch(d1, d2, k=2)

    npairs = len(frames) // 2
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(task, range(min(npairs, threads))))  # untimed warm-up: detector creation, first-touch
        t0 = time.perf_counter()
        list(ex.map(task, range(npairs)))
        dt = time.perf_counter() - t0
    return len(frames) / dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    area = (a.width * a.height) / (640.0 * 480.0)  # bounded sample: about the same CPU work whatever the frame size
    sample = a.cpu_sample or max(2, int(max(2 * threads, 16) / area))
    sample -= sample % 2
    a_s = argparse.Namespace(**vars(a))
    a_s.batch = sample
    frames = make_frames(a_s, 0)
    for _ in range(a.warmup):
        cpu_pairs_per_second(a, frames, threads)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cpu_pairs_per_second(a, frames, threads)
    dt = time.perf_counter() - t0
    fps = sample * a.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(a), "frames_per_step_sample": sample,
                       "note": "reference CPU algorithm = oracle restatement (the reference ships no source for this path; "
                               "its prebuilt .so cannot load here: SURVEY.md 8c), all host threads, one frame pair per task"},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d frames (%d pairs) per step x %d steps" % (sample, sample // 2, a.steps)},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def wave_ramp_sizes(spec, steps, depth):
    """Sizes of the first waves of the e2e leg: `spec` = "auto" (depth/5 + the rest when the run passes over the slots only once,
    none otherwise) or a comma-separated list."""
    if spec == "auto":
        first = max(1, depth // 5)
        return [first, depth - first] if steps <= depth and depth >= 4 else []
    return [int(v) for v in spec.split(",") if v]


def wave_plan(n, depth, wave, ramp):
    """Batches per plslam_frontend_submit_host_wave call for n steps: the ramp sizes first, then `wave`; never more than the
    pipeline holds, never more than what is left."""
    ramp = list(ramp)
    out = []
    k = 0
    while k < n:
        m = max(1, min(ramp.pop(0) if ramp else wave, n - k, depth))
        out.append(m)
        k += m
    return out


def bind_host_threads(local, world):
    """One rank per GPU on one node: keep the rank's submitting thread (and the pinned buffers it first touches) on the cores next
    to its GPU.  NVML names the GPU's CPU affinity; when every GPU reports the same set (round 1: 0-31, NUMA 0 for all eight)
    the set is split evenly over the local ranks so that they do not migrate over each other.  Returns what was done."""
    info = {}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = sorted(c for c in range(ncpu) if (words[c // 64] >> (c % 64)) & 1)
        allowed = sorted(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed] or allowed
        nloc = max(1, min(world, int(os.environ.get("LOCAL_WORLD_SIZE", world))))
        if nloc > 1 and len(cpus) >= 2 * nloc:
            k = len(cpus) // nloc
            cpus = cpus[(local % nloc) * k:(local % nloc + 1) * k]
        os.sched_setaffinity(0, cpus)
        info = {"cpus": "%d-%d (%d)" % (cpus[0], cpus[-1], len(cpus))}
        try:
            info["numa_node"] = int(open("/sys/bus/pci/devices/%s/numa_node" % pynvml.nvmlDeviceGetPciInfo(h).busId.lower()[4:]).read())
        except Exception:
            pass
    except Exception as e:
        info = {"unbound": str(e)[:80]}
    return info


def run_ours(a):
    import torch
    import plslam_b200 as pl
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    host_binding = bind_host_threads(local, world) if world > 1 else {}
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    voc, voc_info = None, {}
    if a.workload == "c4" or world > 1:
        try:
            voc, voc_info = broadcast_vocabulary(a, pl, dist, rank, world)
        except Exception as e:  # the vocabulary only matters for C4; never lose a C2 scaling run over it
            if a.workload == "c4":
                raise
            voc_info = {"voc_broadcast": "failed: %s" % e}

    frames = make_frames(a, rank)
    depth = max(1, min(a.depth, a.steps))
    fe = pl.Frontend(a.nfeatures, 1.2, 8, 20, 7, a.max_lines, depth=depth)
    d_images = torch.from_numpy(frames).cuda()
    outs = [fe.alloc(a.batch, device="cuda") for _ in range(depth)]
    out = outs[0]
    # PLSLAM_STREAM_PRIO (experiment switch of the library, digits ORB / host+lines / copies): the device path runs the line
    # branch on the caller's stream, so that stream carries the priority
    streams = [torch.cuda.Stream(priority=-1 if os.environ.get("PLSLAM_STREAM_PRIO", "000")[1:2] == "1" else 0) for _ in range(depth)]
    h_images = torch.from_numpy(frames).pin_memory()
    h_outs = [fe.alloc(a.batch, pinned=True) for _ in range(depth)]

    c4 = a.workload == "c4"
    fvs = bows = None
    if c4:
        fvs = [voc.featvec_batch_device(o["descriptors"], o["kp_counts"], 4) for o in outs]
        bows = [pl.bow_pairs_device(o["keypoints"], o["descriptors"], o["kp_counts"], fv) for o, fv in zip(outs, fvs)]

    def device_steps(n):
        """n steps, step k on stream/slot k % depth (up to `depth` batches in flight)."""
        main = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(main)
        for k in range(n):
            i = k % depth
            fe.process_device(d_images, outs[i], True, stream=streams[i])
            if c4:  # Frame::ComputeBoW + ORBmatcher::SearchByBoW on the pairs of the batch, same stream, no host round trip
                voc.featvec_batch_device(outs[i]["descriptors"], outs[i]["kp_counts"], 4, out=fvs[i], stream=streams[i])
                pl.bow_pairs_device(outs[i]["keypoints"], outs[i]["descriptors"], outs[i]["kp_counts"], fvs[i], out=bows[i],
                                    stream=streams[i])
        for s in streams:
            main.wait_stream(s)

    # ---- device-resident throughput (`value`) ----
    device_steps(max(a.warmup, 3, depth))
    torch.cuda.synchronize()
    fe.check_status()
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0.record()
    device_steps(a.steps)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.result()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    fe.check_status()
    value = world * a.batch * a.steps / (dev_ms * 1e-3)

    # ---- end to end through the C-ABI with host buffers (`e2e`) ----
    # Two public forms of the host path are timed, the better one is reported (both named in the line):
    #   stream : plslam_frontend_submit_host per step (upload, kernels, download of a step chained in the slot's stream)
    #   wave   : plslam_frontend_submit_host_wave, `depth` steps per call (uploads one wave ahead on an upload stream, the
    #            slots of a wave start together, downloads on their own stream)
    wave = a.wave if a.wave > 0 else max(1, depth // 2)
    # Sizes of the first waves of a run (then `wave`).  A run that passes over the slots only once (steps <= slots: the driver's
    # --steps 20) has no steady state whose phase could be kept, so it starts with a small wave (the GPU starts after 4 uploads
    # instead of 10) and sends the rest as one: its staggered start also fills the tail of the first batches' region growing.
    # Measured at 20 steps (profiles/r02_sched_sweeps.log): 21.4-21.6 k frames/s end to end with waves of 10 + 10, 24.4-26.8 k with
    # 3-8 + rest; at 64 steps a ramp costs the phase (28.2 k -> 23.8-25.6 k), so longer runs keep equal waves.
    wave_ramp = wave_ramp_sizes(a.wave_ramp, a.steps, depth)
    wave_sub = [0]  # batches submitted through the wave form so far: batch t goes to slot t % depth, buffer set (t // depth) & 1

    def host_steps(n, mode):
        if mode == "wave":
            k = 0
            for m in wave_plan(n, depth, wave, wave_ramp):
                t0 = wave_sub[0]
                fe.submit_host_wave([h_images] * m, [h_outs2[((t0 + i) // depth) & 1][(t0 + i) % depth] for i in range(m)], True)
                wave_sub[0] += m
                k += m
        else:
            for k in range(n):
                fe.submit_host(h_images, h_outs[k % depth], True)
        fe.wait_host()

    h_outs2 = [h_outs, [fe.alloc(a.batch, pinned=True) for _ in range(depth)]]
    e2e_modes = {}
    for mode in ("stream", "wave"):
        host_steps(2 * depth if mode == "wave" else max(2, depth), mode)  # two waves: both buffer sets touched
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        host_steps(a.steps, mode)
        e2e_modes[mode] = max_over_ranks(time.perf_counter() - t0)
        barrier()
    # results of the last wave-form step must equal the device-resident ones (same frames)
    ho, do = h_outs2[0][0], {k: v.cpu() for k, v in out.items() if k in ("kp_counts", "line_counts", "descriptors", "line_descriptors")}
    assert torch.equal(ho["kp_counts"], do["kp_counts"]) and torch.equal(ho["line_counts"], do["line_counts"]), "host path counts differ"
    for f in (0, a.batch - 1):
        nk, nl = int(do["kp_counts"][f]), int(do["line_counts"][f])
        assert torch.equal(ho["descriptors"][f, :nk], do["descriptors"][f, :nk]), "host path ORB descriptors differ"
        assert torch.equal(ho["line_descriptors"][f, :nl], do["line_descriptors"][f, :nl]), "host path LBD descriptors differ"
    e2e_mode = min(e2e_modes, key=e2e_modes.get)
    e2e_s = e2e_modes[e2e_mode]
    e2e = world * a.batch * a.steps / e2e_s
    h2d = int(h_images.numel())
    d2h = int(sum(v.numel() * v.element_size() for v in h_outs[0].values()))

    line = None
    if rank == 0:
        kp_avg = float(out["kp_counts"].float().mean())
        # ---- per-stage device times (CUDA events on the launching streams) for the roofline ----
        fe.enable_timing(True)
        acc = {}
        reps = 3
        for _ in range(reps):
            fe.process_device(d_images, out, True)
            torch.cuda.synchronize()  # one batch in flight: stage times are uncontended
            for n, ms in fe.stage_times():
                acc[n] = acc.get(n, 0.0) + ms / reps
        fe.enable_timing(False)
        stages = algorithmic_bytes(a, kp_avg, 5.0 * kp_avg)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
        dom = max(acc, key=acc.get)
        dom_bytes = stages.get(dom, 0) * a.batch
        achieved = dom_bytes / (acc[dom] * 1e-3) / 1e9
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
            traffic = prof.get(dom)
        except Exception:
            pass
        step_ms = dev_ms / a.steps
        lat_sum = sum(acc.values())
        # With `depth` batches in flight the stage times above are LATENCIES of one batch (they overlap with the kernels of
        # the other batches); what a stage costs the pipeline at saturation is modelled as its share of the summed
        # latencies times the measured step time.
        sat_ms = {k: step_ms * v / lat_sum for k, v in acc.items()} if lat_sum > 0 else {}
        survey_total = survey_bytes_per_frame(a, int(round(kp_avg))) * a.batch
        # all-pairs Hamming is not an HBM problem: popc32 per second against the SM's integer rate
        kq = int(round(kp_avg))
        popc = None
        if acc.get("match_orb_knn2"):
            n_popc = (a.batch // 2) * kq * kq * 8
            sm_mhz = (clocks.get("sm_mhz") or 1965.0)
            popc_peak = 148 * 16 * sm_mhz * 1e6
            popc = {"kernel": "match_orb_knn2", "popc32_per_launch": n_popc, "kernel_ms": acc["match_orb_knn2"],
                    "achieved_popc32_per_s": n_popc / (acc["match_orb_knn2"] * 1e-3), "peak_popc32_per_s": popc_peak,
                    "frac": n_popc / (acc["match_orb_knn2"] * 1e-3) / popc_peak,
                    "peak_source": "nominal: 16 popc per clock per SM (CUDA arithmetic-throughput table) x 148 SMs x measured SM clock"}
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": acc[dom],
                    "kernel_ms_at_saturation_model": sat_ms.get(dom),
                    "frac_at_saturation_model": (dom_bytes / (sat_ms[dom] * 1e-3) / 1e9 / peak) if sat_ms.get(dom) else None,
                    "note": "lsd_grow (region growing) is latency bound (greedy growth is sequential inside a frame: one warp per "
                            "frame when many batches are in flight, 8 warps per frame with in-order retirement for small batches), "
                            "not HBM bound; kernel_ms is the latency of ONE batch with %d batches in flight elsewhere idle, the "
                            "*_at_saturation_model figures spread the measured step time over the stages by latency share; see "
                            "DESIGN.md" % depth,
                    "pipeline": {"algorithmic_bytes_per_step": survey_total, "bytes_per_frame": survey_total // a.batch,
                                 "source": "SURVEY.md 8d (21.46 MB per 640x480 frame)",
                                 "achieved": survey_total / (step_ms * 1e-3) / 1e9,
                                 "frac": survey_total / (step_ms * 1e-3) / 1e9 / peak},
                    "popc": popc,
                    "stages_ms": {k: round(v, 4) for k, v in sorted(acc.items(), key=lambda kv: -kv[1])},
                    "stages_ms_at_saturation_model": {k: round(v, 4) for k, v in sorted(sat_ms.items(), key=lambda kv: -kv[1])},
                    # algorithmic GB/s of every stage (same definition as `achieved`), one batch in flight, the ORB and
                    # line branches running side by side on two streams
                    "stages_gbs": {k: round(stages.get(k, 0) * a.batch / (v * 1e-3) / 1e9, 1)
                                   for k, v in sorted(acc.items(), key=lambda kv: -kv[1]) if v > 0}}
        latency = None
        if world == 1 and not a.no_latency:
            try:
                latency = single_frame_latency(a, frames)
            except Exception as e:
                latency = {"error": str(e)}
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            threads = os.cpu_count() or 1
            area = (a.width * a.height) / (640.0 * 480.0)
            sample = a.cpu_sample or min(a.batch, max(2, int(max(4 * threads, 32) / area)))
            sample -= sample % 2
            dt = cpu_pairs_per_second(a, frames[:sample], threads)
            cpu = {"value": sample / dt, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "first %d frames (%d pairs) of the same batch, one frame pair per task, %d threads" % (sample, sample // 2, threads)}
            ocv = opencv_primitives_fps(a, frames[:sample], threads)
            if ocv:
                cpu["opencv_primitives"] = {"value": ocv, "unit": UNIT,
                                            "note": "cv2 4.13 resize+FAST+blur pyramid, LSD_REFINE_ADV, BFMatcher knn2 on the same sample and "
                                                    "threads: a lower bound on the CPU cost (no quad-tree / orientation / rBRIEF / LBD)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic",
                "config": {"workload": workload_name(a), "frames_per_gpu_per_step": a.batch,
                           "cache": "inputs + intermediates (~12 MB per 640x480 frame, scaled with the frame area) exceed the 126 MB L2; no flush needed",
                           "steps_in_flight": depth, **voc_info,
                           "parallelism": "frames sharded over %d rank(s), no data-path collective" % world},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        # what one rank moved over PCIe while it was timed (its share of the wall clock of the e2e run): far below
                        # the ~57 GB/s a B200 measured here sustains, i.e. the bus is not what the e2e figure waits for
                        "per_rank_copy_gbs": {"h2d": h2d * a.steps / e2e_s / 1e9, "d2h": d2h * a.steps / e2e_s / 1e9},
                        "host_binding": host_binding,
                        **({"note": "single pass over the slots (steps <= slots): the e2e leg starts with a small first wave (--wave-ramp), "
                                    "so its batches run staggered - the first batches' latency-bound region growing overlaps the "
                                    "throughput-bound kernels of the rest - while the device-resident loop starts all batches in phase and "
                                    "ends with the tail of every batch's slowest frames; e2e can therefore exceed `value` in such a run "
                                    "(DESIGN.md section 6, profiles/r02_sched_sweeps.log)"} if wave_ramp else {}),
                        "by_api": {m: world * a.batch * a.steps / t for m, t in e2e_modes.items()},
                        "api": ("plslam_frontend_submit_host_wave (%s steps per call, successive calls rotate over the slots) + plslam_frontend_wait_host" % ("+".join(map(str, wave_ramp)) + " then %d" % wave if wave_ramp else str(wave)) if e2e_mode == "wave"
                                else "plslam_frontend_submit_host x K + plslam_frontend_wait_host") +
                               " (pinned host buffers; H2D, kernels and D2H of every step inside the timed region, up to `steps_in_flight` steps overlapped)" +
                               ("; the C4 extras (ComputeBoW + SearchByBoW) are device-path only and not part of this e2e figure" if c4 else "")},
                "gpu_launches": world * a.steps * (fe.launches_per_call(True) + (5 if c4 else 0)),
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "latency": latency,
                "counts": {"keypoints_per_frame": kp_avg, "lines_per_frame": float(out["line_counts"].float().mean())}}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def main():
    a = apply_workload(parse())
    if a.batch % 2:
        a.batch += 1
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
