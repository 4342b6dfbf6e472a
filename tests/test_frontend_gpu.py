"""The batched, pipelined front-end (plslam_frontend_*) against the oracle: every frame of a bench-sized batch,
both the device-resident and the asynchronous host path, with several batches in flight (pipeline slots must not
interfere), including the frame-pair kNN matches."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_batch(oracle, frames, nfeatures=1000, max_lines=40):
    import threading
    tl = threading.local()

    def one(f):
        if not hasattr(tl, "orb"):
            tl.orb = oracle.OrbOracle(nfeatures)
        k, d = tl.orb.extract(frames[f])
        kl, ld, lf, _ = oracle.extract_lines(frames[f], max_lines)
        return k, d, kl, ld, lf

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        return list(ex.map(one, range(len(frames))))


def _compare(oracle, pl, out, ref, B, tag):
    kps = pl.kps_from_tensor(out["keypoints"]); kls = pl.keylines_from_tensor(out["keylines"])
    desc = out["descriptors"].cpu().numpy(); ldesc = out["line_descriptors"].cpu().numpy()
    funcs = out["line_functions"].cpu().numpy()
    kc = out["kp_counts"].cpu().numpy(); lc = out["line_counts"].cpu().numpy()
    om = out["orb_matches"].cpu().numpy(); lm = out["line_matches"].cpu().numpy()
    for f in range(B):
        k, d, kl, ld, lf = ref[f]
        assert kc[f] == len(k) and lc[f] == len(kl), "%s frame %d counts" % (tag, f)
        for fld in k.dtype.names:
            assert np.array_equal(kps[f, :kc[f]][fld], k[fld]), "%s frame %d keypoint %s" % (tag, f, fld)
        assert np.array_equal(desc[f, :kc[f]], d), "%s frame %d ORB descriptors" % (tag, f)
        for fld in kl.dtype.names:
            assert np.array_equal(kls[f, :lc[f]][fld], kl[fld]), "%s frame %d keyline %s" % (tag, f, fld)
        assert np.array_equal(ldesc[f, :lc[f]], ld) and np.array_equal(funcs[f, :lc[f]], lf), "%s frame %d LBD" % (tag, f)
    for p in range(B // 2):
        a, b = ref[2 * p], ref[2 * p + 1]
        assert np.array_equal(om[p, :len(a[1])], oracle.knn2(a[1], b[1])), "%s pair %d ORB matches" % (tag, p)
        assert np.array_equal(lm[p, :len(a[3])], oracle.knn2(a[3], b[3])), "%s pair %d line matches" % (tag, p)


def test_full_bench_batch_all_paths(oracle):
    import torch
    import bench
    import argparse
    import plslam_b200 as pl
    B = 64
    a = argparse.Namespace(batch=B, width=640, height=480)
    frames = bench.make_frames(a, 0)
    ref = _oracle_batch(oracle, frames)
    depth = 3
    fe = pl.Frontend(depth=depth)
    d_images = torch.from_numpy(frames).cuda()
    # device path: 5 submissions over 3 slots / 3 streams; all must give the same, correct result
    outs = [fe.alloc(B, device="cuda") for _ in range(5)]
    streams = [torch.cuda.Stream() for _ in range(depth)]
    torch.cuda.synchronize()
    for k in range(5):
        fe.process_device(d_images, outs[k], True, stream=streams[k % depth])
    torch.cuda.synchronize()
    fe.check_status()
    for k in (0, 3, 4):
        _compare(oracle, pl, outs[k], ref, B, "device submission %d" % k)
    # asynchronous host path, 4 submissions in flight over 3 slots
    h_images = torch.from_numpy(frames).pin_memory()
    h_outs = [fe.alloc(B, pinned=True) for _ in range(4)]
    for k in range(4):
        fe.submit_host(h_images, h_outs[k], True)
    fe.wait_host()
    for k in (0, 3):
        _compare(oracle, pl, h_outs[k], ref, B, "host submission %d" % k)
    # synchronous host path from pageable memory
    out = fe.alloc(B)
    fe.process_host(frames, out, True)
    _compare(oracle, pl, out, ref, B, "pageable host")


def test_4k_frame_nfeatures_8000(oracle):
    """Config C5 shape: 3840x2160, nFeatures = 8000 (single frame: the oracle needs a few seconds)."""
    import plslam_b200 as pl
    from plslam_b200.synth import synth_frame
    img = synth_frame(1, 3840, 2160)
    ex = pl.ORBextractor(nfeatures=8000)
    k, d = ex(img)
    ok, od = oracle.OrbOracle(nfeatures=8000).extract(img)
    assert len(k) == len(ok) >= 8000
    for fld in ok.dtype.names:
        assert np.array_equal(k[fld], ok[fld]), fld
    assert np.array_equal(d, od)
    ls = pl.LineSegment()
    kl, ld, lf = ls.ExtractLineSegment(img)
    okl, old, olf, _ = oracle.extract_lines(img, 40)
    for fld in okl.dtype.names:
        assert np.array_equal(kl[fld], okl[fld]), fld
    assert np.array_equal(ld, old) and np.array_equal(lf, olf)
