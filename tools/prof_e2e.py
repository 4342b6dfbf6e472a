"""Where does the end-to-end (host buffers) path lose against the device-resident path?  Host time per submit call, and
throughput with the copies removed one at a time (via tiny frames is not possible: use timing only)."""
import sys, os, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'rgbd-pl-slam_b200'))
import numpy as np, torch, argparse
import plslam_b200 as pl
import bench
a = argparse.Namespace(batch=256, width=640, height=480)
frames = bench.make_frames(a, 0)
depth, steps = int(os.environ.get('E2E_DEPTH', '16')), 64
fe = pl.Frontend(depth=depth)
h_images = torch.from_numpy(frames).pin_memory()
h_outs = [fe.alloc(256, pinned=True) for _ in range(depth)]
for k in range(2 * depth): fe.submit_host(h_images, h_outs[k % depth], True)
fe.wait_host()
lat = []
t0 = time.perf_counter()
for k in range(steps):
    t1 = time.perf_counter()
    fe.submit_host(h_images, h_outs[k % depth], True)
    lat.append(time.perf_counter() - t1)
t_submit = time.perf_counter() - t0
fe.wait_host()
t_all = time.perf_counter() - t0
lat = np.array(lat) * 1e3
print("e2e: %.2f ms/step (%.0f frames/s); all submits returned after %.1f ms; submit call ms: median %.3f, p90 %.3f, max %.3f, first-16 mean %.3f, rest mean %.3f"
      % (t_all / steps * 1e3, steps * 256 / t_all, t_submit * 1e3, np.median(lat), np.percentile(lat, 90), lat.max(), lat[:16].mean(), lat[16:].mean()))

# completion-ordered submission
for k in range(depth): fe.submit_host_slot(fe.acquire_slot(), h_images, h_outs[k], True)
fe.wait_host()
t0 = time.perf_counter()
for k in range(steps):
    s = fe.acquire_slot()
    fe.submit_host_slot(s, h_images, h_outs[s], True)
fe.wait_host()
t_all = time.perf_counter() - t0
print("e2e, completion-ordered slots: %.2f ms/step (%.0f frames/s)" % (t_all / steps * 1e3, steps * 256 / t_all))
