"""Region-growing modes side by side: each mode runs in its own process (the switches are read once per process),
every accepted LSD segment and every output of the line extractor is compared bit for bit with mode 0 (one warp per
frame, the sequential scan), the line branch is timed with CUDA events, and the scheduler counters of k_lsd_grow_aw are
printed per frame.

  python tools/prof_aw.py                       # default matrix
  python tools/prof_aw.py child <mode> <K> <B> <W> <H> <gen> <out.npz>     (internal)
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rgbd-pl-slam_b200"))
STAT_NAMES = ["groups", "regions", "squash_run", "squash_done", "insert", "rects", "valsteps", "sched_idle", "work_idle",
              "dirty", "reruns", "frames", "cyc_kernel", "cyc_w_wait", "cyc_w_grow", "cyc_w_squash", "cyc_w_scan",
              "cyc_s_retire", "cyc_s_plist", "cyc_s_idle", "pix_done", "pix_squash", "winfull", "poolfull", "heur_skip",
              "p_held", "p_tomb", "p_rob"]


def frames(gen, B, W, H):
    import numpy as np
    from plslam_b200.synth import synth_frame
    if gen == "curvy":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import cv2

        def curvy(seed):
            rng = np.random.default_rng(seed)
            img = synth_frame(seed, W, H).copy()
            for _ in range(25):
                c = (int(rng.integers(0, W)), int(rng.integers(0, H)))
                ax = (int(rng.integers(8, 120)), int(rng.integers(8, 90)))
                cv2.ellipse(img, c, ax, float(rng.uniform(0, 180)), 0, 360, int(rng.integers(0, 256)),
                            int(rng.integers(1, 4)) if rng.random() < 0.7 else -1)
            f = cv2.GaussianBlur(img.astype(np.float32), (0, 0), 1.0) + rng.normal(0.0, float(rng.uniform(2, 9)), size=img.shape).astype(np.float32)
            return np.clip(np.rint(f), 0, 255).astype(np.uint8)
        base = [curvy(7000 + i) for i in range(min(B, 16))]
    else:
        base = [synth_frame(i, W, H) for i in range(min(B, 16))]
    reps = (B + len(base) - 1) // len(base)
    return np.stack((base * reps)[:B])


def child(mode, K, B, W, H, gen, out):
    import ctypes as C
    import numpy as np
    import torch
    import plslam_b200 as pl
    imgs = frames(gen, B, W, H)
    d = torch.from_numpy(imgs).cuda()
    ls = pl.LineSegment(max_lines=0 if W * H <= 1280 * 720 else 40)  # (keep-all output capacity is 8192 lines)
    res = ls.extract_batch_device(d)
    torch.cuda.synchronize()
    ls.check_status()
    L = pl.lib()
    st = (C.c_ulonglong * 32)()
    L.plslam_debug_grow_stats.argtypes = [C.POINTER(C.c_ulonglong)]
    L.plslam_debug_grow_stats(st)
    stats = list(st)
    segs = [ls.segments(f) for f in range(min(B, 16))]
    kl = pl.keylines_from_tensor(res[0])
    cnt = res[3].cpu().numpy()
    desc = res[1].cpu().numpy()
    desc = np.concatenate([desc[f, :cnt[f]] for f in range(B)]) if cnt.sum() else np.zeros(0)
    # timing: line branch, several launches
    for _ in range(2):
        ls.extract_batch_device(d, res)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    R = 3
    e0.record()
    for _ in range(R):
        ls.extract_batch_device(d, res)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / R
    ls.check_status()
    np.savez(out, ms=ms, stats=np.array(stats, dtype=np.uint64), cnt=cnt, desc=desc,
             kl=np.concatenate([kl[f, :cnt[f]] for f in range(B)]) if cnt.sum() else np.zeros(0),
             **{"seg%d" % i: s for i, s in enumerate(segs)})


def run(mode, K, B, W, H, gen, tag):
    out = "/tmp/prof_aw_%s.npz" % tag
    env = dict(os.environ, PLSLAM_GROW_MODE=str(mode), PLSLAM_SW_K=str(K), PLSLAM_DEBUG_STOP_AFTER_GROW="")
    env.pop("PLSLAM_DEBUG_STOP_AFTER_GROW")
    t0 = time.time()
    p = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(mode), str(K), str(B), str(W), str(H), gen, out],
                       env=env, capture_output=True, text=True, timeout=150)
    if p.returncode != 0:
        print("  mode %d K %d: FAILED rc=%d (%.0f s)\n%s" % (mode, K, p.returncode, time.time() - t0, (p.stderr or "")[-1500:]), flush=True)
        return None
    import numpy as np
    return dict(np.load(out))


def same(a, b):
    import numpy as np
    if a is None or b is None:
        return False
    for k in a:
        if k in ("ms", "stats"):
            continue
        if a[k].shape != b[k].shape or a[k].tobytes() != b[k].tobytes():
            return False
    return True


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), sys.argv[7], sys.argv[8])
        sys.exit(0)
    cases = [(1, 640, 480, "synth"), (16, 640, 480, "synth"), (16, 640, 480, "curvy"), (256, 640, 480, "synth"), (8, 1280, 720, "curvy")]
    if len(sys.argv) > 1:
        cases = [(int(a.split(",")[0]), int(a.split(",")[1]), int(a.split(",")[2]), a.split(",")[3]) for a in sys.argv[1:]]
    bad = 0
    for (B, W, H, gen) in cases:
        print("== batch %d, %dx%d, %s" % (B, W, H, gen), flush=True)
        ref = run(0, 0, B, W, H, gen, "ref")
        if ref is None:
            bad += 1
            continue
        print("  mode 0 (warp per frame):        %8.3f ms/batch  lines/frame %.1f" % (float(ref["ms"]), ref["cnt"].mean()), flush=True)
        ks = [int(k) for k in os.environ.get("PROF_AW_K", "8,16,32").split(",")]
        for (mode, K) in [(2, k) for k in ks]:
            if mode == 2 and K == 32 and B > 148:
                continue
            r = run(mode, K, B, W, H, gen, "m%dk%d" % (mode, K))
            if r is None:
                bad += 1
                continue
            ok = same(ref, r)
            bad += 0 if ok else 1
            line = "  mode %d K %2d: %8.3f ms/batch  %s" % (mode, K, float(r["ms"]), "IDENTICAL" if ok else "*** MISMATCH ***")
            if mode == 2:
                s = r["stats"].astype(float)
                fr = max(s[11], 1.0)
                line += "  | per frame: " + " ".join("%s=%.0f" % (STAT_NAMES[i], s[i] / fr) for i in range(11))
                line += "\n      kcycles/frame: " + " ".join("%s=%.0f" % (STAT_NAMES[i][4:], s[i] / fr / 1e3) for i in range(12, 20))
                line += " | " + " ".join("%s=%.0f" % (STAT_NAMES[i], s[i] / fr) for i in range(20, 28))
            print(line, flush=True)
    print("prof_aw: %d problem(s)" % bad)
    sys.exit(1 if bad else 0)
