"""Is the end-to-end loss a phase effect?  The device-resident loop enqueues all `depth` steps at once, so the slots run the same
kernel at the same time; a per-step upload in front of each step staggers them.  (1) a 1.4 ms spin kernel instead of the upload;
(2) wave scheduling: the uploads of a whole wave of `depth` steps run in the background (second set of input buffers) while the
previous wave computes, every step of the wave waits for the wave's LAST upload, results leave on a download stream."""
import sys, os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'rgbd-pl-slam_b200'))
import numpy as np, torch, argparse, time
import plslam_b200 as pl
import bench
a = argparse.Namespace(batch=256, width=640, height=480)
frames = bench.make_frames(a, 0)
depth = int(os.environ.get("E2E_DEPTH", "15"))
waves = 4
steps = depth * waves
fe = pl.Frontend(depth=depth)
d_images = torch.from_numpy(frames).cuda()
h_images = torch.from_numpy(frames).pin_memory()
streams = [torch.cuda.Stream() for _ in range(depth)]
up, down = torch.cuda.Stream(), torch.cuda.Stream()
d_in = [[torch.empty_like(d_images) for _ in range(depth)] for _ in range(2)]
d_out = [[fe.alloc(256, device="cuda") for _ in range(depth)] for _ in range(2)]
h_out = [fe.alloc(256, pinned=True) for _ in range(depth)]
spin_cycles = int(1.4e-3 * 1.9e9)

def run_plain(n, spin):
    main = torch.cuda.current_stream()
    for s in streams: s.wait_stream(main)
    for k in range(n):
        st = streams[k % depth]
        if spin:
            with torch.cuda.stream(st): torch.cuda._sleep(spin_cycles)
        fe.process_device(d_images, d_out[0][k % depth], True, stream=st)
    for s in streams: main.wait_stream(s)

def run_waves(nw, download=True):
    main = torch.cuda.current_stream()
    for s in streams: s.wait_stream(main)
    up.wait_stream(main); down.wait_stream(main)
    done = {}
    dl = {}
    for w in range(nw):
        b = w % 2
        with torch.cuda.stream(up):
            for s in range(depth):
                if w >= 2: up.wait_event(done[(w - 2, s)])   # the buffer's previous reader
                d_in[b][s].copy_(h_images, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(up)
        for s in range(depth):
            st = streams[s]
            st.wait_event(ev)
            if w >= 2 and download: st.wait_event(dl[(w - 2, s)])   # result buffer drained
            fe.process_device(d_in[b][s], d_out[b][s], True, stream=st)
            done[(w, s)] = torch.cuda.Event(); done[(w, s)].record(st)
            if download:
                down.wait_event(done[(w, s)])
                with torch.cuda.stream(down):
                    for key, v in d_out[b][s].items():
                        if key in h_out[s]: h_out[s][key].copy_(v, non_blocking=True)
                    dl[(w, s)] = torch.cuda.Event(); dl[(w, s)].record(down)
    for s in streams: main.wait_stream(s)
    main.wait_stream(up); main.wait_stream(down)

def timed(fn, *args):
    fn(*args); torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(*args); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3

print("depth %d, %d steps" % (depth, steps))
print("device-resident, all enqueued at once : %.2f ms/step" % (timed(run_plain, steps, False) / steps), flush=True)
print("1.4 ms spin kernel before every step  : %.2f ms/step" % (timed(run_plain, steps, True) / steps), flush=True)
print("waves, uploads only                   : %.2f ms/step" % (timed(run_waves, waves, False) / steps), flush=True)
print("waves, uploads + downloads            : %.2f ms/step" % (timed(run_waves, waves, True) / steps), flush=True)
print("device-resident again                 : %.2f ms/step" % (timed(run_plain, steps, False) / steps), flush=True)
