"""Uploads spread over the run like the host path issues them (one per step, each released when the step `depth` earlier
has finished), with and without anything waiting for them.  Separates "a copy concurrent with steady-state kernels costs
pipeline time" from "a dependency on a copy costs pipeline time"."""
import sys, os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'rgbd-pl-slam_b200'))
import numpy as np, torch, argparse
import plslam_b200 as pl
import bench
a = argparse.Namespace(batch=256, width=640, height=480)
frames = bench.make_frames(a, 0)
depth, steps = 15, 60
fe = pl.Frontend(depth=depth)
d_images = torch.from_numpy(frames).cuda()
h_images = torch.from_numpy(frames).pin_memory()
scratch = [torch.empty_like(d_images) for _ in range(2)]
outs = [fe.alloc(256, device="cuda") for _ in range(depth)]
streams = [torch.cuda.Stream() for _ in range(depth)]
up = torch.cuda.Stream()
def run(n, mode):
    """mode: none | h2d (throttled, nobody waits) | h2d_dep (throttled, step k waits for upload k) | d2d (throttled device copy)
    | h2d_small (throttled, 1/16 of the bytes)"""
    main = torch.cuda.current_stream()
    for s in streams: s.wait_stream(main)
    up.wait_stream(main)
    done = [None] * n
    for k in range(n):
        st = streams[k % depth]
        if mode != "none":
            if k >= depth: up.wait_event(done[k - depth])
            with torch.cuda.stream(up):
                if mode == "d2d": scratch[k % 2].copy_(d_images, non_blocking=True)
                elif mode == "h2d_small": scratch[k % 2][:16].copy_(h_images[:16], non_blocking=True)
                else: scratch[k % 2].copy_(h_images, non_blocking=True)
            if mode == "h2d_dep":
                ev = torch.cuda.Event(); ev.record(up); st.wait_event(ev)
        fe.process_device(d_images, outs[k % depth], True, stream=st)
        done[k] = torch.cuda.Event(); done[k].record(st)
    for s in streams: main.wait_stream(s)
    main.wait_stream(up)
for mode in ("none", "h2d", "h2d_dep", "d2d", "h2d_small", "none"):
    run(depth, mode); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(steps, mode); e1.record(); torch.cuda.synchronize()
    print("%-10s %.2f ms/step" % (mode, e0.elapsed_time(e1) / steps), flush=True)
