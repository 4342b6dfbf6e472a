"""Lines-only throughput against the number of batches in flight (one LineSegment instance + stream each)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'rgbd-pl-slam_b200'))
import numpy as np, torch
import plslam_b200 as pl
from plslam_b200.synth import synth_frame
B = 256
base = np.stack([synth_frame(i) for i in range(16)])
imgs = torch.from_numpy(np.concatenate([base] * (B // 16))).cuda()
for K in [int(x) for x in sys.argv[1:]] or [1, 2, 4, 8, 16]:
    ls = [pl.LineSegment() for _ in range(K)]
    st = [torch.cuda.Stream() for _ in range(K)]
    outs = [l.extract_batch_device(imgs) for l in ls]
    torch.cuda.synchronize()
    def run(reps):
        for r in range(reps):
            for k in range(K):
                with torch.cuda.stream(st[k]):
                    ls[k].extract_batch_device(imgs, outs[k], stream=st[k])
    run(1); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    e0.record()
    for s in st: s.wait_stream(torch.cuda.current_stream())
    run(reps)
    for s in st: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("in flight %2d: %.1f ms per batch-of-256 (amortised), %.0f frames/s" % (K, ms / (reps * K), reps * K * B / ms * 1e3), flush=True)
    del ls, outs
