"""Cycle breakdown of k_lsd_grow from the instrumented build (`make -C rgbd-pl-slam_b200 PROF=1`)."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'rgbd-pl-slam_b200'))
import numpy as np, torch
import plslam_b200 as pl
pl.LIB_PATH = os.path.join(ROOT, 'rgbd-pl-slam_b200', 'libplslam_b200_prof.so')
from plslam_b200.synth import synth_frame
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
base = np.stack([synth_frame(i) for i in range(16)])
imgs = torch.from_numpy(np.concatenate([base] * (B // 16))).cuda()
ls = pl.LineSegment()
out = ls.extract_batch_device(imgs)
torch.cuda.synchronize(); ls.check_status()
L = pl.lib()
buf = (C.c_ulonglong * 16)()
L.plslam_debug_grow_prof(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ls.extract_batch_device(imgs, out); e1.record(); torch.cuda.synchronize()
L.plslam_debug_grow_prof(buf)
v = [x / B for x in buf]
names = ["load", "loop", "rect", "refine(incl regrow)", "kernel", "single-pixel first growths", "frontier points over batches", "batches with a candidate", "regions", "batches", "rounds", "mis-speculations", "points", "", "", "first growths below min_reg_size"]
print("lines pipeline %.2f ms/batch; per frame:" % e0.elapsed_time(e1))
for n, x in zip(names, v):
    if n:
        print("  %-22s %12.0f" % (n, x))
print("  seed scan+other (cycles) %10.0f" % (v[4] - v[0] - v[1] - v[2] - v[3] + 0))
if buf[5]:
    print("  CTA-per-frame mode: rounds %.0f, slots started %.2f and committed %.2f per round; rounds cut by poison %.0f, by a skipped free seed %.0f (per frame)"
          % (v[5], buf[6] / buf[5], buf[7] / buf[5], v[13], v[14]))
