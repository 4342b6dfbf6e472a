"""Stage times of one 256-frame batch (depth 1) and device throughput at depth 16, for quick A/B of kernel changes.
usage: prof_stages.py [depth16=1|0]   (PLSLAM_LIB selects another build of the library)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'rgbd-pl-slam_b200'))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch
import plslam_b200 as pl
from plslam_b200.synth import synth_pair
B = 256
frames = np.stack([f for s in range(B // 2) for f in synth_pair(s)])
d_images = torch.from_numpy(frames).cuda()
fe = pl.Frontend(depth=1)
out = fe.alloc(B, device="cuda")
for _ in range(2):
    fe.process_device(d_images, out, True)
torch.cuda.synchronize(); fe.check_status()
fe.enable_timing(True)
acc = {}
for _ in range(3):
    fe.process_device(d_images, out, True)
    torch.cuda.synchronize()
    for k, v in fe.stage_times():
        acc[k] = acc.get(k, 0.0) + v / 3
fe.enable_timing(False)
print("stages_ms", {k: round(v, 3) for k, v in sorted(acc.items(), key=lambda kv: -kv[1])})
if len(sys.argv) < 2 or sys.argv[1] != "0":
    depth = 16
    fe = pl.Frontend(depth=depth)
    outs = [fe.alloc(B, device="cuda") for _ in range(depth)]
    streams = [torch.cuda.Stream() for _ in range(depth)]
    def run(n):
        main = torch.cuda.current_stream()
        for s in streams: s.wait_stream(main)
        for k in range(n):
            fe.process_device(d_images, outs[k % depth], True, stream=streams[k % depth])
        for s in streams: main.wait_stream(s)
    run(depth); torch.cuda.synchronize(); fe.check_status()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 48
    e0.record(); run(n); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("depth16: %.2f ms/step, %.0f frames/s" % (ms / n, n * B / ms * 1e3))
