"""Extended parity sweep (beyond the test-suite): many seeds and odd frame sizes through the batched front-end, every
frame compared with the CPU oracle (keypoints, ORB descriptors, KeyLines, LBD bytes, line equations, all LSD segments)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'rgbd-pl-slam_b200'))
from concurrent.futures import ThreadPoolExecutor
import threading
import numpy as np, torch
import plslam_b200 as pl
from plslam_b200.synth import synth_frame
from oracle import bindings as ob
ob.build()
tl = threading.local()
def oracle_one(img):
    if not hasattr(tl, "orb"):
        tl.orb = ob.OrbOracle(1000)
    k, d = tl.orb.extract(img)
    kl, ld, lf, _ = ob.extract_lines(img, 40)
    seg, _ = ob.lsd_detect(img, compat=0)
    return k, d, kl, ld, lf, seg
bad = 0; total = 0
t0 = time.time()
def curvy(seed, W, H):
    """synthetic frame with ellipses and heavier noise: curved edges send many regions through refine() /
    reduce_region_radius(), textured noise produces many tiny regions"""
    import cv2
    rng = np.random.default_rng(seed)
    img = synth_frame(seed, W, H).copy()
    for _ in range(25):
        c = (int(rng.integers(0, W)), int(rng.integers(0, H)))
        ax = (int(rng.integers(8, 120)), int(rng.integers(8, 90)))
        cv2.ellipse(img, c, ax, float(rng.uniform(0, 180)), 0, 360, int(rng.integers(0, 256)), int(rng.integers(1, 4)) if rng.random() < 0.7 else -1)
    f = cv2.GaussianBlur(img.astype(np.float32), (0, 0), 1.0) + rng.normal(0.0, float(rng.uniform(2, 9)), size=img.shape).astype(np.float32)
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)
SCALE = int(os.environ.get("SWEEP_SCALE", "1"))
for (W, H, B, seed0, gen) in ((640, 480, 192 * SCALE, 1000, synth_frame), (640, 480, 128 * SCALE, 7000, curvy), (640, 480, 24, 5000, synth_frame),
                              (517, 389, 48, 2000, curvy), (801, 601, 32, 3000, synth_frame), (1280, 720, 16, 4000, curvy)):
    imgs = np.stack([gen(seed0 + i, W, H) for i in range(B)])
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        ref = list(ex.map(oracle_one, imgs))
    fe = pl.Frontend(depth=1)
    out = fe.alloc(B, device="cuda")
    fe.process_device(torch.from_numpy(imgs).cuda(), out, False)
    torch.cuda.synchronize(); fe.check_status()
    kps = pl.kps_from_tensor(out["keypoints"]); kls = pl.keylines_from_tensor(out["keylines"])
    desc = out["descriptors"].cpu().numpy(); ldesc = out["line_descriptors"].cpu().numpy(); funcs = out["line_functions"].cpu().numpy()
    kc = out["kp_counts"].cpu().numpy(); lc = out["line_counts"].cpu().numpy()
    ls = pl.LineSegment(max_lines=0)
    for f in range(B):
        k, d, kl, ld, lf, seg = ref[f]
        ok = kc[f] == len(k) and lc[f] == len(kl)
        ok = ok and all(np.array_equal(kps[f, :kc[f]][n], k[n]) for n in k.dtype.names) and np.array_equal(desc[f, :kc[f]], d)
        ok = ok and all(np.array_equal(kls[f, :lc[f]][n], kl[n]) for n in kl.dtype.names)
        ok = ok and np.array_equal(ldesc[f, :lc[f]], ld) and np.array_equal(funcs[f, :lc[f]], lf)
        if f % 8 == 0:  # all accepted segments (not only the 40 strongest) through the single-frame entry point
            ls.ExtractLineSegment(imgs[f])
            g = ls.segments(0)
            ok = ok and len(g) == len(seg) and np.array_equal(g[:, :6], seg[:, :6])
        total += 1
        if not ok:
            bad += 1
            print("MISMATCH %dx%d seed %d" % (W, H, seed0 + f), flush=True)
    print("%dx%d: %d frames checked (%.0f s)" % (W, H, B, time.time() - t0), flush=True)
print("parity sweep: %d frames, %d mismatches" % (total, bad))
sys.exit(1 if bad else 0)
