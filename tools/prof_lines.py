"""Quick device-time profile of the line extractor on a synthetic batch."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'rgbd-pl-slam_b200'))
import numpy as np, torch
import plslam_b200 as pl
from plslam_b200.synth import synth_frame
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (640, 480)
base = np.stack([synth_frame(i, W, H) for i in range(16)])
imgs = torch.from_numpy(np.concatenate([base] * (B // 16))).cuda()
ls = pl.LineSegment()
out = ls.extract_batch_device(imgs)
torch.cuda.synchronize(); ls.check_status()
for it in range(2):
    ls.extract_batch_device(imgs, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 3
e0.record()
for it in range(K):
    ls.extract_batch_device(imgs, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("LINES batch=%d %dx%d: %.3f ms/batch, %.0f frames/s, counts[0:4]=%s" % (B, W, H, ms, B / ms * 1e3, out[3][:4].tolist()))
