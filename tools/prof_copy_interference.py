"""Does host->device DMA traffic slow the kernels of the front-end?  Device-resident pipeline with and without a background
stream that uploads 78.6 MB once per step."""
import sys, os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'rgbd-pl-slam_b200'))
import numpy as np, torch, argparse
import plslam_b200 as pl
import bench
a = argparse.Namespace(batch=256, width=640, height=480)
frames = bench.make_frames(a, 0)
depth, steps = 16, 48
fe = pl.Frontend(depth=depth)
d_images = torch.from_numpy(frames).cuda()
h_images = torch.from_numpy(frames).pin_memory()
scratch = [torch.empty_like(d_images) for _ in range(2)]
outs = [fe.alloc(256, device="cuda") for _ in range(depth)]
streams = [torch.cuda.Stream() for _ in range(depth)]
up = torch.cuda.Stream()
def run(n, copies):
    main = torch.cuda.current_stream()
    for s in streams: s.wait_stream(main)
    up.wait_stream(main)
    for k in range(n):
        if copies:
            with torch.cuda.stream(up):
                scratch[k % 2].copy_(h_images, non_blocking=True)
        fe.process_device(d_images, outs[k % depth], True, stream=streams[k % depth])
    for s in streams: main.wait_stream(s)
    main.wait_stream(up)
def run_into_input(n, shift=0):
    """an upload ordered before every step like the host path; shift = 0: into the buffer that step reads, shift = 8: into
    the buffer another slot reads 8 steps later"""
    main = torch.cuda.current_stream()
    for s in streams: s.wait_stream(main)
    for k in range(n):
        with torch.cuda.stream(streams[k % depth]):
            d_in[(k + shift) % depth].copy_(h_images, non_blocking=True)
        fe.process_device(d_in[k % depth], outs[k % depth], True, stream=streams[k % depth])
    for s in streams: main.wait_stream(s)
d_in = [d_images.clone() for _ in range(depth)]
run_into_input(depth); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run_into_input(steps); e1.record(); torch.cuda.synchronize()
print("upload into the input buffer of each step (torch streams, device API): %.2f ms/step" % (e0.elapsed_time(e1) / steps), flush=True)
run_into_input(depth, 8); torch.cuda.synchronize()
e0.record(); run_into_input(steps, 8); e1.record(); torch.cuda.synchronize()
print("same uploads in the chain, but into a buffer read 8 steps later: %.2f ms/step" % (e0.elapsed_time(e1) / steps), flush=True)
for copies in (0, 1):
    run(depth, copies); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(steps, copies); e1.record(); torch.cuda.synchronize()
    print("background uploads %d: %.2f ms/step" % (copies, e0.elapsed_time(e1) / steps), flush=True)
