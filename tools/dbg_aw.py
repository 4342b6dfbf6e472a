"""Per-frame comparison of the region-growing modes: which frames differ from the sequential scan, and how.
  python tools/dbg_aw.py [K] [B] [gen] [reps]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rgbd-pl-slam_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def child(B, gen, out, reps):
    import numpy as np
    import torch
    import plslam_b200 as pl
    from prof_aw import frames
    imgs = frames(gen, B, 640, 480)
    d = torch.from_numpy(imgs).cuda()
    ls = pl.LineSegment(max_lines=0)
    res = {}
    for r in range(reps):
        ls.extract_batch_device(d)
        torch.cuda.synchronize()
        ls.check_status()
        for f in range(B):
            res["r%d_f%d" % (r, f)] = ls.segments(f)
    np.savez(out, **res)


if __name__ == "__main__":
    if sys.argv[1] == "child":
        child(int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5]))
        sys.exit(0)
    import numpy as np
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    gen = sys.argv[3] if len(sys.argv) > 3 else "synth"
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    outs = {}
    for mode in (0, 2):
        out = "/tmp/dbg_aw_%d.npz" % mode
        env = dict(os.environ, PLSLAM_GROW_MODE=str(mode), PLSLAM_SW_K=str(K))
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(B), gen, out, str(reps if mode else 1)], env=env,
                           capture_output=True, text=True, timeout=200)
        if p.returncode:
            print("mode %d failed: %s" % (mode, p.stderr[-2000:]))
            sys.exit(1)
        outs[mode] = dict(np.load(out))
    nbad = 0
    for r in range(reps):
        for f in range(B):
            a, b = outs[0]["r0_f%d" % f], outs[2]["r%d_f%d" % (r, f)]
            if a.shape == b.shape and a.tobytes() == b.tobytes():
                continue
            nbad += 1
            n = min(len(a), len(b))
            first = next((i for i in range(n) if a[i].tobytes() != b[i].tobytes()), n)
            sa = set(x.tobytes() for x in a)
            sb = set(x.tobytes() for x in b)
            print("rep %d frame %d: %d vs %d segments, first difference at %d, only-ref %d only-aw %d" %
                  (r, f, len(a), len(b), first, len(sa - sb), len(sb - sa)))
            if first < n:
                print("   ref:", a[first][:5], "\n   aw: ", b[first][:5])
    print("dbg_aw K=%d B=%d %s: %d differing (frame, rep) of %d" % (K, B, gen, nbad, reps * B))
