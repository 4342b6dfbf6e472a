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
    # host-scheduled slots (copies on their own streams, hand-over by event polling), 5 submissions over 3 slots
    s_outs = [fe.alloc(B, pinned=True) for _ in range(5)]
    for k in range(5):
        fe.submit_host_slot(fe.acquire_slot(), h_images, s_outs[k], True)
    fe.wait_host()
    for k in (0, 2, 4):
        _compare(oracle, pl, s_outs[k], ref, B, "host-scheduled submission %d" % k)
    # wave submission: 3 waves (3, 3 and 2 batches) over the 3 slots, uploads of a wave overlapping the wave before it
    w_outs = [fe.alloc(B, pinned=True) for _ in range(8)]
    for v in w_outs:
        v["kp_counts"].fill_(-7)
    fe.submit_host_wave([h_images] * 3, w_outs[0:3], True)
    fe.submit_host_wave([h_images] * 3, w_outs[3:6], True)
    fe.submit_host_wave([h_images] * 2, w_outs[6:8], True)
    fe.wait_host()
    for k in (0, 2, 4, 5, 7):
        _compare(oracle, pl, w_outs[k], ref, B, "wave submission %d" % k)
    # waves smaller than the pipeline rotate over the slots (sizes 2, 1, 2, 3 over 3 slots: every slot alternates between
    # its two buffer sets); alternate batches carry the frames in reversed order, so a stale staging buffer or a result set
    # overwritten before it left for the host would show
    h_rev = torch.from_numpy(np.ascontiguousarray(frames[::-1])).pin_memory()
    ref_rev = ref[::-1]
    r_outs = [fe.alloc(B, pinned=True) for _ in range(8)]
    ins = [h_images if k % 2 == 0 else h_rev for k in range(8)]
    k = 0
    for m in (2, 1, 2, 3):
        fe.submit_host_wave(ins[k:k + m], r_outs[k:k + m], True)
        k += m
    fe.wait_host()
    for k in (0, 1, 2, 4, 5, 7):
        _compare(oracle, pl, r_outs[k], ref if k % 2 == 0 else ref_rev, B, "rotating wave submission %d" % k)
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


def test_full_size_batch_properties():
    """BASELINE config C2 at its full size (256 frames = 128 pairs, 16 batches in flight), checked through properties that do
    not need the CPU oracle on every frame: bitwise determinism across pipeline slots and submissions, independence of a
    frame's result from the batch it travels in, level-major keypoint order inside every frame, and the kNN results
    re-derived with independent torch arithmetic (Hamming distance of the reported neighbours, ordering d1 <= d2, and
    minimality of d1 over the whole train set) for every query of every pair."""
    import torch
    import bench
    import argparse
    import plslam_b200 as pl
    B, depth = 256, 16
    a = argparse.Namespace(batch=B, width=640, height=480)
    frames = bench.make_frames(a, 0)
    d_images = torch.from_numpy(frames).cuda()
    fe = pl.Frontend(depth=depth)
    outs = [fe.alloc(B, device="cuda") for _ in range(depth)]
    streams = [torch.cuda.Stream() for _ in range(depth)]
    torch.cuda.synchronize()
    for rep in range(2):  # 32 submissions: every slot is used twice
        for k in range(depth):
            fe.process_device(d_images, outs[k], True, stream=streams[k])
    torch.cuda.synchronize()
    fe.check_status()
    ref = outs[0]
    kc, lc = ref["kp_counts"].long(), ref["line_counts"].long()
    assert int(kc.min()) > 900 and int(lc.min()) == 40
    cap = ref["descriptors"].shape[1]
    valid = (torch.arange(cap, device="cuda")[None, :] < kc[:, None])
    lvalid = (torch.arange(ref["line_descriptors"].shape[1], device="cuda")[None, :] < lc[:, None])
    for k in (1, 7, depth - 1):  # determinism across slots / submissions (valid rows only: the rest is never written)
        o = outs[k]
        assert torch.equal(o["kp_counts"], ref["kp_counts"]) and torch.equal(o["line_counts"], ref["line_counts"])
        assert torch.equal(o["descriptors"][valid], ref["descriptors"][valid])
        assert torch.equal(o["keypoints"][valid], ref["keypoints"][valid])
        assert torch.equal(o["line_descriptors"][lvalid], ref["line_descriptors"][lvalid])
        assert torch.equal(o["keylines"][lvalid], ref["keylines"][lvalid])
        assert torch.equal(o["line_functions"][lvalid], ref["line_functions"][lvalid])
    # batch independence: frames 10..13 alone give the same rows
    fe2 = pl.Frontend()
    small = fe2.alloc(4, device="cuda")
    fe2.process_device(d_images[10:14].contiguous(), small, True)
    torch.cuda.synchronize()
    for j in range(4):
        n = int(kc[10 + j])
        assert int(small["kp_counts"][j]) == n
        assert torch.equal(small["descriptors"][j, :n], ref["descriptors"][10 + j, :n])
        assert torch.equal(small["keypoints"][j, :n], ref["keypoints"][10 + j, :n])
        assert torch.equal(small["line_descriptors"][j, :40], ref["line_descriptors"][10 + j, :40])
    # keypoints are level-major (octave is int32 word 5 of cv::KeyPoint) in every frame
    octv = ref["keypoints"][:, :, 5].long()
    octv = torch.where(valid, octv, torch.full_like(octv, 99))
    assert bool((octv[:, 1:] >= octv[:, :-1]).all())
    # kNN (k = 2) of every pair re-derived with torch
    desc = ref["descriptors"]
    m = ref["orb_matches"].long()
    pop = torch.tensor([bin(i).count("1") for i in range(256)], device="cuda", dtype=torch.int16)
    for p in range(0, B // 2, 8):
        q, t = desc[2 * p, :kc[2 * p]], desc[2 * p + 1, :kc[2 * p + 1]]
        d = pop[(q[:, None, :] ^ t[None, :, :]).long()].sum(-1)  # nq x nt Hamming distances
        mm = m[p, :len(q)]
        rows = torch.arange(len(q), device="cuda")
        assert torch.equal(d[rows, mm[:, 0]].long(), mm[:, 1]) and torch.equal(d[rows, mm[:, 2]].long(), mm[:, 3])
        s, _ = torch.sort(d.long(), dim=1)
        assert torch.equal(s[:, 0], mm[:, 1]) and torch.equal(s[:, 1], mm[:, 3])
        assert torch.equal(d.argmin(1), mm[:, 0])  # ties -> lowest index


def test_featureless_frames_in_a_batch(oracle):
    """Empty and ragged results inside one batch: constant frames (no gradient at all: the seed sort, the region growing and
    the rectangle queue see zero work), a frame of faint noise below every threshold, a single step edge (one line, no corner)
    and two ordinary frames, through the extractors and the pipelined front-end with the pair matcher."""
    import torch
    import plslam_b200 as pl
    from plslam_b200.synth import synth_frame
    H, W = 480, 640
    rng = np.random.default_rng(5)
    frames = np.stack([
        synth_frame(3, W, H),
        np.zeros((H, W), np.uint8),
        np.full((H, W), 255, np.uint8),
        (128 + rng.integers(0, 3, (H, W))).astype(np.uint8),
        np.concatenate([np.full((H, W // 2), 40, np.uint8), np.full((H, W - W // 2), 200, np.uint8)], axis=1),
        synth_frame(4, W, H),
    ])
    B = len(frames)
    ref = _oracle_batch(oracle, frames)
    assert [len(r[0]) for r in ref][1:5] == [0, 0, 0, 0] and [len(r[2]) for r in ref][1:5] == [0, 0, 0, 1]
    # extractors on their own
    kl, desc, funcs, counts = pl.LineSegment().extract_batch_host(frames)
    assert list(counts) == [len(r[2]) for r in ref]
    # front-end: device path and host wave path, frame pairs (0,1) (2,3) (4,5) include empty-vs-empty and empty-vs-full matches
    fe = pl.Frontend(depth=2)
    out = fe.alloc(B, device="cuda")
    fe.process_device(torch.from_numpy(frames).cuda(), out, True)
    torch.cuda.synchronize()
    fe.check_status()
    _compare(oracle, pl, out, ref, B, "featureless frames, device path")
    h_out = fe.alloc(B, pinned=True)
    fe.submit_host_wave([torch.from_numpy(frames).pin_memory()], [h_out], True)
    fe.wait_host()
    _compare(oracle, pl, h_out, ref, B, "featureless frames, wave path")


def test_odd_batch_other_capacities(oracle):
    """A batch of 5 frames (two pairs, the last frame unpaired) with nFeatures = 500 and the 15 strongest lines, 517x389 frames
    (no dimension a multiple of a tile), three submissions over two slots."""
    import torch
    import plslam_b200 as pl
    from plslam_b200.synth import synth_frame
    H, W, B = 389, 517, 5
    frames = np.stack([synth_frame(500 + i, W, H) for i in range(B)])
    ref = _oracle_batch(oracle, frames, nfeatures=500, max_lines=15)
    fe = pl.Frontend(nfeatures=500, max_lines=15, depth=2)
    d_images = torch.from_numpy(frames).cuda()
    outs = [fe.alloc(B, device="cuda") for _ in range(3)]
    for o in outs:
        fe.process_device(d_images, o, True)
    torch.cuda.synchronize()
    fe.check_status()
    for k in (0, 2):
        _compare(oracle, pl, outs[k], ref, B, "odd batch, submission %d" % k)
    h_out = fe.alloc(B, pinned=True)
    fe.submit_host(torch.from_numpy(frames).pin_memory(), h_out, True)
    fe.wait_host()
    _compare(oracle, pl, h_out, ref, B, "odd batch, host path")
