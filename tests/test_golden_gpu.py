"""The CUDA path (through the C-ABI) directly against the committed golden vectors of tests/golden/ — data produced by
cv2 4.13 and by the reference's own DBoW2 sources, not by this repository's oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_bow_transform_equals_reference_dbow2():
    import plslam_b200 as pl
    g = np.load(os.path.join(G, "dbow2_ref.npz"))
    V = pl.ORBVocabulary(os.path.join(G, "voc_k6_L3.txt"))
    for lu in (0, 1, 2):
        t = V.transform(g["desc"], lu)
        for k, v in t.items():
            assert np.array_equal(v, g["lu%d_%s" % (lu, k)]), (lu, k)
    out = pl.knn2_host(g["desc"][:200], g["desc"][200:])
    # nearest-neighbour distances are FORB::distance values: check the diagonal pairs through an all-pairs run
    d = g["forb_distance"]
    assert np.all(out[:, 1] <= d)


def test_lsd_segments_against_cv2():
    import plslam_b200 as pl
    g = np.load(os.path.join(G, "cv2_lsd.npz"))
    ls = pl.LineSegment(max_lines=0)
    for i in range(2):
        ls.ExtractLineSegment(g["img%d" % i])
        seg = ls.segments(0)
        ref = g["lines%d" % i]
        s = {tuple(np.float32(r[:4])) for r in seg}
        same = sum(tuple(r) in s for r in ref)
        # pinned sin/cos instead of libm's: at most a couple of segments per frame may differ from cv2
        assert abs(len(seg) - len(ref)) <= 2 and same >= len(ref) - 2, (len(seg), len(ref), same)


def test_undistort_against_cv2():
    import torch
    import plslam_b200 as pl
    p = np.load(os.path.join(G, "cv2_primitives.npz"))
    c = p["undist_calib"]
    cal = dict(zip(("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3"), [float(v) for v in c]), bf=40.0)
    xy = p["undist_in"]
    inside = (xy[:, 0] < 640) & (xy[:, 1] < 480)  # the depth lookup needs in-image keypoints
    kps = np.zeros(int(inside.sum()), pl.KP_DTYPE)
    kps["x"], kps["y"] = xy[inside, 0], xy[inside, 1]
    out = pl.frame_post_host(cal, pl.frame_image_bounds(cal, 640, 480), kps, np.ones((480, 640), np.float32))
    assert np.array_equal(out["un_xy"], p["undist_out"][inside])
    b = pl.frame_image_bounds(cal, 640, 480)
    m = p["undist_out"][:4]  # the four image corners
    assert np.array_equal(b, np.array([min(m[0, 0], m[2, 0]), max(m[1, 0], m[3, 0]), min(m[0, 1], m[1, 1]), max(m[2, 1], m[3, 1])], np.float32))


def test_orb_extractor_equals_the_reference_library():
    """The CUDA ORBextractor against outputs of the reference's OWN code: ORB_SLAM2::ORBextractor::operator() executed from
    lib/libORB_SLAM2.so in the build container (tests/golden/reference_code.py; fixture reference_library.npz, ex*).
    Standard frame sizes only (640x480 nFeatures 1000, 1280x720 nFeatures 2000)."""
    import hashlib
    import plslam_b200 as pl
    from plslam_b200.synth import synth_frame
    g = np.load(os.path.join(G, "reference_library.npz"))
    done = 0
    for k in range(int(g["ex_n"])):
        seed, W, H, nf = (int(v) for v in g["ex%d_args" % k])
        if (W, H) not in ((640, 480), (1280, 720)):
            continue
        img = synth_frame(seed, W, H)
        assert hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() == str(g["ex%d_img_sha256" % k])
        kps, desc = pl.ORBextractor(nfeatures=nf)(img)
        ref = g["ex%d_kps" % k]
        assert len(kps) == len(ref)
        for f in ("x", "y", "size", "response", "octave"):
            assert np.array_equal(kps[f], ref[f]), (k, f)
        assert np.abs(kps["angle"] - ref["angle"]).max() <= 1e-4 and np.array_equal(kps["angle"], ref["angle"]), k
        assert np.array_equal(desc, g["ex%d_desc" % k]), k
        done += 1
    assert done == 3


# ------------------------------------------------------------------------------------------------
# The matcher KERNELS directly against outputs of the reference's own machine code (reference_library.npz,
# projection_boundary.npz): the inputs are rebuilt from the seeds with the CUDA extractor (it equals the oracle bit for
# bit, tests/test_orb_gpu.py), no oracle call is involved.
# ------------------------------------------------------------------------------------------------
def _features(seed, cache={}):
    import plslam_b200 as pl
    from plslam_b200.synth import synth_pair
    if seed not in cache:
        a, b = synth_pair(seed)
        ex = pl.ORBextractor()
        cache[seed] = (ex(a), ex(b))
    return cache[seed]


def _scale_factors():
    import plslam_b200 as pl
    return np.asarray(pl.ORBextractor().GetScaleFactors(), np.float32)


def test_k_projection_equals_the_reference_matcher():
    """k_projection (ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono), @0x80d00) on the pj* fixtures."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import plslam_b200 as pl
    from matchdata import projection_case
    g = np.load(os.path.join(G, "reference_library.npz"))
    sf = _scale_factors()
    total = 0
    for k in range(int(g["pj_n"])):
        seed, motion, th, mono = g["pj%d_args" % k]
        (ka, da), (kb, db) = _features(int(seed))
        last, cur, cam, _, tc, tl = projection_case(ka, da, kb, db, sf, seed=int(seed), motion=float(motion))
        m, n = pl.search_by_projection_host(last, cur, cam, sf, tc, tl, float(th), bool(mono), True)
        assert n == int(g["pj%d_n" % k]) and np.array_equal(m, g["pj%d_match" % k]), k
        total += n
    assert total > 3000


def test_k_projection_on_window_edge_cases_equals_the_reference_matcher():
    """k_projection on projection_boundary.npz: current key points planted within one ulp of the search window's edge, where
    an evaluation of the projection that is not the binary's (float division, fused multiply-adds, float gemm sums) decides
    differently; expected results come from the reference's own code (tests/golden/make_projection_boundary.py)."""
    import plslam_b200 as pl
    g = np.load(os.path.join(G, "projection_boundary.npz"))
    planted = 0
    for k in range(int(g["n"])):
        last = {n[len("c%d_last_" % k):]: g[n] for n in g.files if n.startswith("c%d_last_" % k)}
        cur = {n[len("c%d_cur_" % k):]: g[n] for n in g.files if n.startswith("c%d_cur_" % k)}
        a = g["c%d_args" % k]
        m, n = pl.search_by_projection_host(last, cur, g["c%d_cam" % k], g["c%d_sf" % k], g["c%d_tc" % k], g["c%d_tl" % k],
                                            float(a[2]), bool(a[3]), True)
        assert n == int(g["c%d_n" % k]) and np.array_equal(m, g["c%d_match" % k]), k
        planted += int(a[4])
    assert planted > 2000


def test_k_local_points_equals_the_reference_matcher():
    """k_local_points (ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), @0x79f10) on the lp* fixtures."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import plslam_b200 as pl
    from matchdata import local_points_case
    g = np.load(os.path.join(G, "reference_library.npz"))
    sf = _scale_factors()
    for k in range(int(g["lp_n"])):
        seed, th, nnr, jit = g["lp%d_args" % k]
        (ka, da), (kb, db) = _features(int(seed))
        mp, fr, cam4 = local_points_case(ka, da, kb, db, seed=10 * int(seed) + int(th), jitter=float(jit))
        m, n = pl.search_local_points_host(mp, fr, cam4, sf, float(th), float(nnr))
        assert n == int(g["lp%d_n" % k]) > 300 and np.array_equal(m, g["lp%d_match" % k]), k


def test_k_bow_equals_the_reference_matcher():
    """k_bow (ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), @0x80150) on the bw* fixtures."""
    import sys
    import torch
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import plslam_b200 as pl
    from matchdata import fake_feature_vector
    g = np.load(os.path.join(G, "reference_library.npz"))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    jobs, keep = [], []
    for k in range(int(g["bw_n"])):
        seed, nnr, ori, nbits = g["bw%d_args" % k]
        (ka, da), (kb, db) = _features(int(seed))
        kf = dict(desc=da, angle=np.ascontiguousarray(ka["angle"]), valid=g["bw%d_valid" % k])
        kf["nodes"], kf["start"], kf["idx"] = fake_feature_vector(da, int(nbits), seed=7)
        f = dict(desc=db, angle=np.ascontiguousarray(kb["angle"]))
        f["nodes"], f["start"], f["idx"] = fake_feature_vector(db, int(nbits), seed=7)
        d, e = {n: dev(v) for n, v in kf.items()}, {n: dev(v) for n, v in f.items()}
        m = torch.empty(len(db), dtype=torch.int32, device="cuda")
        n = torch.zeros(1, dtype=torch.int32, device="cuda")
        keep.append((d, e, m, n))
        jobs.append(pl.BowJob(d["desc"].data_ptr(), d["angle"].data_ptr(), d["valid"].data_ptr(), d["nodes"].data_ptr(),
                              d["start"].data_ptr(), d["idx"].data_ptr(), e["desc"].data_ptr(), e["angle"].data_ptr(),
                              e["nodes"].data_ptr(), e["start"].data_ptr(), e["idx"].data_ptr(), m.data_ptr(), n.data_ptr(),
                              len(da), len(db), len(kf["nodes"]), len(f["nodes"]), float(nnr), int(ori)))
    pl.bow_batch_device(jobs, max(max(j.n1, j.n2) for j in jobs), "cuda")
    torch.cuda.synchronize()
    for k, (_, _, m, n) in enumerate(keep):
        assert int(n) == int(g["bw%d_n" % k]) > 200 and np.array_equal(m.cpu().numpy(), g["bw%d_match" % k]), k


def test_k_bow_keyframes_equals_the_reference_matcher():
    """k_bow in its key frame / key frame form (ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12), @0x82cc0) on the
    bk* fixtures of reference_library2.npz, through plslam_match_bow_kfkf_host."""
    import plslam_b200 as pl
    from test_golden_cpu import _kfkf_cases
    g = np.load(os.path.join(G, "reference_library2.npz"))
    ex = pl.ORBextractor()
    done = 0
    for k, kf1, kf2, nnr, ori in _kfkf_cases(g, ex):
        m, n = pl.search_by_bow_kfkf_host(kf1, kf2, nnr, ori)
        assert n == int(g["bk%d_n" % k]) > 150 and np.array_equal(m, g["bk%d_match" % k]), k
        done += 1
    assert done == 8


def test_k_projection_keyframe_mode_equals_the_reference_matcher():
    """k_projection in its key-frame mode (ORBmatcher::SearchByProjection(Frame&, KeyFrame*, set<MapPoint*>&, th, ORBdist),
    @0x7e8c0) on the rk* fixtures of reference_library2.npz; the device PredictScale (restated glibc logf) against the
    oracle's on a sweep of distances."""
    import plslam_b200 as pl
    from oracle import bindings as ob
    from test_golden_cpu import _reloc_cases
    g = np.load(os.path.join(G, "reference_library2.npz"))
    ex = pl.ORBextractor()
    total = 0
    for k, kf, cur, cam, sf, lsf, tcw, th, od, ori in _reloc_cases(g, ex, _scale_factors()):
        m, n = pl.search_by_projection_kf_host(kf, cur, cam, sf, lsf, tcw, th, od, ori)
        assert n == int(g["rk%d_n" % k]) and np.array_equal(m, g["rk%d_match" % k]), k
        total += n
    assert total > 1500
    lsf = float(np.log(np.float32(1.2)))
    rng = np.random.default_rng(9)
    for d in list(rng.uniform(0.05, 30.0, 200)) + [10.0 / 1.2 ** e for e in range(-2, 10)]:
        assert pl.predict_scale(10.0, float(d), lsf, 8) == ob.predict_scale(10.0, float(d), lsf, 8), d


def test_k_kf_projection_equals_the_reference_matcher():
    """k_kf_projection (ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th), @0x880f0) on the lc*
    fixtures of reference_library2.npz (similarity scales 0.6 / 1 / 1.7)."""
    import plslam_b200 as pl
    from test_golden_cpu import _loop_cases
    g = np.load(os.path.join(G, "reference_library2.npz"))
    ex = pl.ORBextractor()
    total = 0
    for k, kf, mp, scw, mi, th in _loop_cases(g, ex, _scale_factors()):
        m, n = pl.search_by_projection_sim3_host(kf, mp, scw, mi, th)
        assert n == int(g["lc%d_n" % k]) and np.array_equal(m, g["lc%d_match" % k]), k
        total += n
    assert total > 1500


def test_k_fuse_search_equals_the_reference_matcher():
    """k_fuse_search (the matching core of both ORBmatcher::Fuse overloads, @0x7a500 / @0x7bb20) on the fu* and fs* fixtures: the
    reference's sequence of AddObservation / AddMapPoint / Replace calls, vpReplacePoint and nFused follow from the kernel's
    per-point matches plus the replay of the bookkeeping."""
    import plslam_b200 as pl
    from oracle import bindings as ob
    from test_golden_cpu import _fuse_cases, _fuse_sim3_cases
    g = np.load(os.path.join(G, "reference_library2.npz"))
    ex = pl.ORBextractor()
    total = 0
    for k, kf, mp, kp_, th in _fuse_cases(g, ex, _scale_factors()):
        nf, log = ob.fuse_replay(pl.fuse_search_host(kf, mp, th), mp, kp_)
        assert nf == int(g["fu%d_n" % k]) and np.array_equal(np.array(log, np.int32).reshape(-1, 3), g["fu%d_log" % k]), k
        total += nf
    for k, kf, mp, kp_, th in _fuse_sim3_cases(g, ex, _scale_factors()):
        nf, log, rep = ob.fuse_replay_sim3(pl.fuse_search_host(kf, mp, th, scw=kf["scw"]), mp, kp_)
        assert nf == int(g["fs%d_n" % k]) and np.array_equal(np.array(log, np.int32).reshape(-1, 3), g["fs%d_log" % k]), k
        assert np.array_equal(rep, g["fs%d_replace" % k]), k
        total += nf
    assert total > 1500


def test_search_by_sim3_equals_the_reference_matcher():
    """ORBmatcher::SearchBySim3 (@0x838b0) on the s3* fixtures: both directions on the device (k_fuse_search, sim3 mode), the
    derived transforms from plslam_sim3_transforms, the agreement pass on the host."""
    import plslam_b200 as pl
    from oracle import bindings as ob
    from test_golden_cpu import _sim3_cases
    g = np.load(os.path.join(G, "reference_library2.npz"))
    ex = pl.ORBextractor()
    total = 0
    for k, kf1, kf2, mp1, mp2, s12, R12, t12, mi, th in _sim3_cases(g, ex, _scale_factors()):
        m, n = pl.search_by_sim3_host(kf1, kf2, mp1, mp2, s12, R12, t12, th, mi)
        assert n == int(g["s3%d_n" % k]) and np.array_equal(m, g["s3%d_match" % k]), k
        total += n
    assert total > 500


def test_k_is_in_frustum_equals_the_reference():
    """k_is_in_frustum (Frame::isInFrustum, @0xf5190) on the fz* fixtures: 3 x 3000 map points, every field the reference's
    function leaves in a MapPoint."""
    import plslam_b200 as pl
    from test_golden_cpu import _frustum_cases
    g = np.load(os.path.join(G, "reference_library2.npz"))
    seen = 0
    for k, c, lim in _frustum_cases(g):
        r = pl.is_in_frustum_host(c["xyz"], c["normal"], c["dist_range"], c["cam8"], c["tcw"], c["ow"], c["mbf"], c["log_sf"], c["n_levels"], lim)
        for name in ("in_view", "proj", "level", "viewcos"):
            assert np.array_equal(r[name], g["fz%d_%s" % (k, name)]), (k, name)
        seen += int(r["in_view"].sum())
    assert seen > 2000
    assert len(pl.is_in_frustum_host(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 2)), c["cam8"], c["tcw"], c["ow"], 40.0, c["log_sf"], 8, 0.5)["in_view"]) == 0


def test_k_triangulation_equals_the_reference_matcher():
    """k_triangulation (ORBmatcher::SearchForTriangulation, @0x86b30, epipole included) on the tr* fixtures."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import plslam_b200 as pl
    from matchdata import triangulation_case
    from plslam_b200.synth import synth_frame
    g = np.load(os.path.join(G, "reference_library.npz"))
    cases, total = {}, 0
    for k in range(int(g["tr_n"])):
        a = g["tr%d_args" % k]
        seed, only, ori, nbits = int(a[0]), bool(a[1]), bool(a[2]), int(a[4])
        key = (seed, float(a[3]), nbits, tuple(a[5:8]))
        if key not in cases:
            kps, desc = pl.ORBextractor()(synth_frame(seed))
            cases[key] = triangulation_case(kps, desc, seed=seed, stereo_fraction=float(a[3]), nbits=nbits, t21=tuple(a[5:8]))
        kf1, kf2, F12, pose, cam, sf, sg = cases[key]
        ex, ey = pl.epipole(*pose, *cam)
        m, n, _ = pl.search_for_triangulation_host(kf1, kf2, F12, ex, ey, sf, sg, only, ori)
        assert n == int(g["tr%d_n" % k]) and np.array_equal(m, g["tr%d_match" % k]), k
        total += n
    assert total > 1500


def test_k_search_init_equals_the_reference_matcher():
    """k_search_init (ORBmatcher::SearchForInitialization, @0x7db00) on the si* fixtures: vnMatches12, the return value and the
    updated vbPrevMatched."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import plslam_b200 as pl
    from matchdata import frame_grid
    g = np.load(os.path.join(G, "reference_library.npz"))
    feats = {}
    for k in range(int(g["si_n"])):
        seed, win, nnr, ori = g["si%d_args" % k]
        seed = int(seed)
        if seed not in feats:
            (ka, da), (kb, db) = _features(seed)
            mk = lambda kk, d: dict(xy=np.stack([kk["x"], kk["y"]], 1).astype(np.float32), octave=kk["octave"].astype(np.int32),
                                    angle=kk["angle"].astype(np.float32), desc=d)
            f1, f2 = mk(ka, da), mk(kb, db)
            gs, gi, (mnx, mxx, mny, mxy, gwi, ghi) = frame_grid(f2["xy"], 640, 480)
            f2["grid_start"], f2["grid_items"] = gs, gi
            feats[seed] = (f1, f2, np.array([mnx, mny, gwi, ghi], np.float32))
        f1, f2, cam4 = feats[seed]
        m, n, prev = pl.search_for_initialization_host(f1, f2, cam4, f1["xy"].copy(), int(win), float(nnr), bool(ori))
        assert n == int(g["si%d_n" % k]) > 50 and np.array_equal(m, g["si%d_match" % k]), k
        assert np.array_equal(prev, g["si%d_prev" % k]), k
