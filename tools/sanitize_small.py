"""Small end-to-end run for compute-sanitizer (memcheck): odd-sized frames through every kernel of the path."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'rgbd-pl-slam_b200'))
import numpy as np, torch
import plslam_b200 as pl
from plslam_b200.synth import synth_frame, synth_pair, synth_vocabulary_arrays, synth_depth
SIZES = ((333, 250, 2), (640, 480, 4), (401, 303, 2))
if os.environ.get("SAN_SMALL"):
    SIZES = ((333, 250, 2), (270, 200, 2))
for (W, H, B) in SIZES:
    imgs = np.stack([synth_frame(i, W, H) for i in range(B)])
    fe = pl.Frontend(depth=2)
    out = fe.alloc(B, device="cuda")
    fe.process_device(torch.from_numpy(imgs).cuda(), out, True)
    torch.cuda.synchronize(); fe.check_status()
    print(W, H, out["kp_counts"].tolist(), out["line_counts"].tolist())
    if (W, H) == (640, 480):
        voc = pl.ORBVocabulary.from_arrays(10, 3, *synth_vocabulary_arrays(10, 3, 0))
        fv = voc.featvec_batch_device(out["descriptors"], out["kp_counts"], 1)
        res = pl.bow_pairs_device(out["keypoints"], out["descriptors"], out["kp_counts"], fv)
        depth = torch.from_numpy(np.stack([synth_depth(i).astype(np.float32) / 5000 for i in range(B)])).cuda()
        post = pl.frame_post_device(pl.TUM1_CALIB, pl.frame_image_bounds(pl.TUM1_CALIB, W, H), out["keypoints"], out["kp_counts"], depth)
        torch.cuda.synchronize()
        print("bow matches", res["nmatches"].tolist(), "grid items", post["grid_start"][:, -1].tolist())
print("done")
