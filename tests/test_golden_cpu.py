"""The oracle against the COMMITTED golden vectors of tests/golden/ (made by tests/golden/make_golden.py from cv2 4.13,
from the reference's own DBoW2 sources and from constants of the reference binary).  Nothing here imports cv2 or reads
/root/reference: this is the pin that travels to the GPU box."""
import hashlib
import json
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def prim():
    return np.load(os.path.join(G, "cv2_primitives.npz"))


def test_reference_binary_constants(oracle):
    import plslam_b200 as pl
    c = json.load(open(os.path.join(G, "reference_binary.json")))
    p = oracle.orb_pattern().astype("<i4")
    assert hashlib.sha256(p.tobytes()).hexdigest() == c["bit_pattern_31_sha256"]
    assert p[:16].tolist() == c["bit_pattern_31_first16"]
    assert (pl.TH_LOW, pl.TH_HIGH, pl.HISTO_LENGTH) == (c["TH_LOW"], c["TH_HIGH"], c["HISTO_LENGTH"]) == (50, 100, 30)


def test_resize_blur_fast_atan2(oracle, prim):
    assert np.array_equal(oracle.resize_linear(prim["img"], 133, 107), prim["resize_img_133x107"])
    assert np.array_equal(oracle.resize_linear(prim["noise"], 69, 51), prim["resize_noise_69x51"])
    assert np.array_equal(oracle.blur7(prim["img"]), prim["blur_img"])
    assert np.array_equal(oracle.blur7(prim["noise"]), prim["blur_noise"])
    cell = np.ascontiguousarray(prim["img"][16:54, 100:137])
    for th in (20, 7):
        assert np.array_equal(oracle.fast9(prim["img"], th), prim["fast%d_img" % th])
        assert np.array_equal(oracle.fast9(cell, th), prim["fast%d_cell" % th])
    assert len(prim["fast20_img"]) > 20 and len(prim["fast7_img"]) > len(prim["fast20_img"])
    mine = np.array([oracle.fast_atan2(y, x) for y, x in prim["atan2_yx"]], np.float32)
    assert np.array_equal(mine, prim["atan2_deg"])


def test_undistort_points(oracle, prim):
    c = prim["undist_calib"]
    cal = dict(zip(("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3"), [float(v) for v in c]), bf=40.0)
    assert np.array_equal(oracle.undistort_points(cal, prim["undist_in"]), prim["undist_out"])


def test_lsd_front_half(oracle, prim):
    k = oracle.gauss_table_u8(0.75, 7)
    assert np.array_equal(oracle.gauss_blur_u8(prim["noise"], k), prim["lsd_blur_noise"])
    b = oracle.gauss_blur_u8(prim["img"], k)
    ref = prim["lsd_scaled_img"]
    assert np.array_equal(oracle.resize_linear_exact(b, ref.shape[1], ref.shape[0], 0.8), ref)


def test_lsd_segments_identical_to_cv2(oracle):
    g = np.load(os.path.join(G, "cv2_lsd.npz"))
    for i in range(2):
        mine, _ = oracle.lsd_detect(g["img%d" % i], compat=1)  # libm trig mode = what cv2 computes
        assert len(mine) == len(g["lines%d" % i]) > 20
        assert np.array_equal(mine[:, :4].astype(np.float32), g["lines%d" % i])
        assert np.array_equal(mine[:, 4], g["width%d" % i])
        assert np.array_equal(mine[:, 5], g["prec%d" % i])
        assert np.allclose(mine[:, 6], g["nfa%d" % i], rtol=0, atol=1e-9)
        # the pinned-trig mode (the one the CUDA path reproduces) agrees on (nearly) every segment
        pinned, _ = oracle.lsd_detect(g["img%d" % i], compat=0)
        s = {tuple(np.float32(r[:4])) for r in pinned}
        assert sum(tuple(r) in s for r in g["lines%d" % i]) >= len(g["lines%d" % i]) - 2


def test_bow_transform_equals_reference_dbow2(oracle):
    g = np.load(os.path.join(G, "dbow2_ref.npz"))
    voc = oracle.VocOracle(os.path.join(G, "voc_k6_L3.txt"))
    desc = g["desc"]
    for lu in (0, 1, 2):
        t = voc.transform(desc, lu)
        for k, v in t.items():
            assert np.array_equal(v, g["lu%d_%s" % (lu, k)]), (lu, k)
    d = np.array([oracle.descriptor_distance(a, b) for a, b in zip(desc[:200], desc[200:])], np.int32)
    assert np.array_equal(d, g["forb_distance"])


def test_product_host_pieces_against_golden():
    """Host-side scalar pieces of the product (no GPU needed): DescriptorDistance, vocabulary loader + BoW assembly use the
    same golden data through the C-ABI where no device is required."""
    import plslam_b200 as pl
    g = np.load(os.path.join(G, "dbow2_ref.npz"))
    desc = g["desc"]
    d = np.array([pl.DescriptorDistance(a, b) for a, b in zip(desc[:200], desc[200:])], np.int32)
    assert np.array_equal(d, g["forb_distance"])


def test_oracle_equals_the_reference_machine_code(oracle):
    """tests/golden/reference_code.npz holds outputs of lib/libORB_SLAM2.so's own code for four leaf functions of the
    matcher path (tests/golden/reference_code.py executes them); the epipolar cases are pairs of ADJACENT float32 inputs
    on either side of the reference's decision boundary, so the rounding sequence (the binary's FMA pattern) is pinned."""
    g = np.load(os.path.join(G, "reference_code.npz"))
    for v, want in zip(g["radius_in"], g["radius_out"]):
        assert oracle.radius_by_viewing_cos(v) == want, float(v)
    assert oracle.radius_by_viewing_cos(np.float32(0.998)) == 2.5  # (double)0.998f > 0.998
    for a, b, want in zip(g["dd_a"], g["dd_b"], g["dd_out"]):
        assert oracle.descriptor_distance(a, b) == want
    for h, want in zip(g["tm_in"], g["tm_out"]):
        assert oracle.three_maxima(h) == tuple(int(x) for x in want), h
    got = np.array([oracle.check_dist_epipolar_line(k1, k2, F, g["ep_sigma2"][o])
                    for F, k1, k2, o in zip(g["ep_F"], g["ep_kp1"], g["ep_kp2"], g["ep_oct"])], np.uint8)
    assert np.array_equal(got, g["ep_out"]), "%d of %d differ" % (int((got != g["ep_out"]).sum()), len(got))
    assert 300 < int(g["ep_out"].sum()) < 1200


@pytest.mark.skipif(not os.path.exists("/root/reference/lib/libORB_SLAM2.so"), reason="build container only")
def test_reference_machine_code_still_gives_the_golden_answers():
    import hashlib
    import sys
    sys.path.insert(0, G)
    from reference_code import RefCode, SO
    g = np.load(os.path.join(G, "reference_code.npz"))
    assert hashlib.sha256(open(SO, "rb").read()).hexdigest() == str(g["so_sha256"])
    r = RefCode()
    assert [r.descriptor_distance(a, b) for a, b in zip(g["dd_a"][:20], g["dd_b"][:20])] == list(g["dd_out"][:20])
    assert [r.radius_by_viewing_cos(float(v)) for v in g["radius_in"][:9]] == list(g["radius_out"][:9])


def test_oracle_equals_the_loaded_reference_library(oracle):
    """tests/golden/reference_library.npz: outputs of lib/libORB_SLAM2.so itself (dlopen'ed over stub dependencies by
    tests/golden/reference_code.py) — ORBextractor's constructor tables and ORBextractor::DistributeOctTree, the latter under a
    monotonic allocator (its pair<int, ExtractorNode*> sort breaks ties by pointer value; allocation order is the instance
    the oracle restates)."""
    g = np.load(os.path.join(G, "reference_library.npz"))
    for k, pr in enumerate(g["ctor_params"]):
        o = oracle.OrbOracle(int(pr[0]), float(pr[1]), int(pr[2]), int(pr[3]), int(pr[4]))
        t = o.tables()
        for name in ("quota", "scale", "inv_scale", "sigma2", "inv_sigma2", "umax"):
            assert np.array_equal(t[name], g["ctor%d_%s" % (k, name)]), (k, name)
        assert np.array_equal(np.asarray(oracle.orb_pattern(), np.int32).reshape(-1, 2)[:512], g["ctor%d_pattern" % k])
    o = oracle.OrbOracle()
    for k in range(int(g["qt_n"])):
        x0, x1, y0, y1, N = (int(v) for v in g["qt%d_args" % k])
        got = o.distribute(g["qt%d_in" % k], x0, x1, y0, y1, N)
        assert np.array_equal(got, g["qt%d_out" % k]), "quad-tree case %d" % k


def test_oracle_keypoints_equal_the_reference_compute_keypoints_oct_tree(oracle):
    """ORBextractor::ComputeKeyPointsOctTree + computeOrientation executed from lib/libORB_SLAM2.so (fixture
    reference_library.npz, ck*): cell grid, threshold retry, quad-tree, border offsets, patch size, octave and IC angle of
    every keypoint, in the reference's order.  The oracle's final keypoints are those scaled by mvScaleFactor[level]
    (operator() does `keypoint->pt *= scale` for level > 0, @0x77d10)."""
    g = np.load(os.path.join(G, "reference_library.npz"))
    for k in range(int(g["ck_n"])):
        o = oracle.OrbOracle(int(g["ck%d_nf" % k]))
        kps, _ = o.extract(g["ck%d_img" % k])
        scale = o.tables()["scale"]
        ref, counts = g["ck%d_kps" % k], g["ck%d_counts" % k]
        assert len(kps) == len(ref) == int(counts.sum()) > 100
        assert np.array_equal(kps["octave"], ref["octave"])
        assert np.array_equal(np.bincount(kps["octave"], minlength=8), counts)
        s = scale[ref["octave"]]
        for name, want in (("x", np.where(ref["octave"] > 0, ref["x"] * s, ref["x"])), ("y", np.where(ref["octave"] > 0, ref["y"] * s, ref["y"])),
                           ("angle", ref["angle"]), ("response", ref["response"]), ("size", ref["size"])):
            assert np.array_equal(kps[name].view(np.uint32), want.astype(np.float32).view(np.uint32)), (k, name)


def _reference_extractions():
    from plslam_b200.synth import synth_frame
    g = np.load(os.path.join(G, "reference_library.npz"))
    for k in range(int(g["ex_n"])):
        seed, W, H, nf = (int(v) for v in g["ex%d_args" % k])
        img = g["ex%d_img" % k] if "ex%d_img" % k in g.files else synth_frame(seed, W, H)
        assert hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() == str(g["ex%d_img_sha256" % k])
        yield k, nf, img, g["ex%d_kps" % k], g["ex%d_desc" % k]


def test_oracle_equals_the_reference_orb_extractor_end_to_end(oracle):
    """ORB_SLAM2::ORBextractor::operator() executed from lib/libORB_SLAM2.so — the reference's own code from pyramid to
    descriptors, OpenCV entry points replaced by ABI-exact shims over the cv2-pinned primitives (tests/golden/
    reference_code.py) — on six frames (256x200 ... 1280x720, nFeatures 300 ... 2000): every keypoint field and every
    descriptor byte of the oracle is identical, in the same order."""
    n = 0
    for k, nf, img, kps, desc in _reference_extractions():
        ok, od = oracle.OrbOracle(nf).extract(img)
        assert len(ok) == len(kps) > 200, k
        for f in kps.dtype.names:
            assert np.array_equal(ok[f].view(np.uint32), kps[f].view(np.uint32)), (k, f)
        assert np.array_equal(od, desc), k
        n += 1
    assert n == 6


def test_grid_and_window_search_equal_the_reference_frame_code(oracle):
    """Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea executed from lib/libORB_SLAM2.so on a faked Frame
    (fixture reference_library.npz, fg*): the oracle's grid (frame_oracle.cc) and its window search (the function both
    SearchByProjection restatements use) reproduce the reference's cells and candidate lists, order included."""
    g = np.load(os.path.join(G, "reference_library.npz"))
    kps, b = g["fg_kps"], g["fg_bounds"]
    xy = np.stack([kps["x"], kps["y"]], 1).astype(np.float32)
    calib = dict(fx=500.0, fy=500.0, cx=320.0, cy=240.0, k1=0.0, k2=0.0, p1=0.0, p2=0.0, k3=0.0, bf=40.0)  # k1 == 0: no undistortion
    fp = oracle.frame_post(calib, np.array([b[0], b[1], b[2], b[3]], np.float32), xy, np.zeros((480, 640), np.float32))
    assert np.array_equal(fp["grid_start"], g["fg_start"]) and np.array_equal(fp["grid_items"], g["fg_items"])
    assert 0 < len(g["fg_items"]) < len(kps)  # some keypoints lie outside the bounds and are in no cell
    cam4 = np.array([b[0], b[2], np.float32(64.0) / (b[1] - b[0]), np.float32(48.0) / (b[3] - b[2])], np.float32)
    off = np.concatenate([[0], np.cumsum(g["fg_res_len"])])
    octave = kps["octave"].astype(np.int32)
    for k, (x, y, r, lo, hi) in enumerate(g["fg_queries"]):
        got = oracle.get_features_in_area(np.float32(x), np.float32(y), np.float32(r), int(lo), int(hi), cam4, g["fg_start"],
                                          g["fg_items"], xy, octave)
        assert np.array_equal(got, g["fg_res"][off[k]:off[k + 1]]), (k, x, y, r, lo, hi)
    assert int(g["fg_res_len"].sum()) > 3000


def test_local_map_search_equals_the_reference_matcher(oracle):
    """ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (@0x79f10) executed from lib/libORB_SLAM2.so on
    faked Frame / MapPoint objects (fixture reference_library.npz, lp*): the oracle assigns the same map point to every
    keypoint and returns the same count."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from matchdata import local_points_case
    from plslam_b200.synth import synth_pair
    g = np.load(os.path.join(G, "reference_library.npz"))
    o = oracle.OrbOracle()
    sf = o.tables()["scale"]
    feats = {}
    for k in range(int(g["lp_n"])):
        seed, th, nnr, jit = g["lp%d_args" % k]
        seed = int(seed)
        if seed not in feats:
            a, b = synth_pair(seed)
            feats[seed] = (o.extract(a), o.extract(b))
        (ka, da), (kb, db) = feats[seed]
        mp, fr, cam4 = local_points_case(ka, da, kb, db, seed=10 * seed + int(th), jitter=float(jit))
        m, n = oracle.search_local_points(mp, fr, cam4, sf, float(th), float(nnr))
        assert n == int(g["lp%d_n" % k]) > 300 and np.array_equal(m, g["lp%d_match" % k]), k


def test_search_by_bow_equals_the_reference_matcher(oracle):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (@0x80150) executed from lib/libORB_SLAM2.so on faked
    KeyFrame / Frame objects whose feature vectors are real std::map's (fixture reference_library.npz, bw*)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from matchdata import fake_feature_vector
    from plslam_b200.synth import synth_pair
    g = np.load(os.path.join(G, "reference_library.npz"))
    o = oracle.OrbOracle()
    feats = {}
    for k in range(int(g["bw_n"])):
        seed, nnr, ori, nbits = g["bw%d_args" % k]
        seed, nbits = int(seed), int(nbits)
        if seed not in feats:
            a, b = synth_pair(seed)
            feats[seed] = (o.extract(a), o.extract(b))
        (ka, da), (kb, db) = feats[seed]
        kf = dict(desc=da, angle=np.ascontiguousarray(ka["angle"]), valid=g["bw%d_valid" % k])
        kf["nodes"], kf["start"], kf["idx"] = fake_feature_vector(da, nbits, seed=7)
        f = dict(desc=db, angle=np.ascontiguousarray(kb["angle"]))
        f["nodes"], f["start"], f["idx"] = fake_feature_vector(db, nbits, seed=7)
        m, n = oracle.search_by_bow(kf, f, float(nnr), bool(ori))
        assert n == int(g["bw%d_n" % k]) > 200 and np.array_equal(m, g["bw%d_match" % k]), k


def _kfkf_cases(g, extract):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from matchdata import fake_feature_vector
    from plslam_b200.synth import synth_pair
    feats = {}
    for k in range(int(g["bk_n"])):
        seed, nnr, ori, nbits = g["bk%d_args" % k]
        seed, nbits = int(seed), int(nbits)
        if seed not in feats:
            a, b = synth_pair(seed)
            feats[seed] = (extract(a), extract(b))
        (ka, da), (kb, db) = feats[seed]
        kf1 = dict(desc=da, angle=np.ascontiguousarray(ka["angle"]), valid=g["bk%d_valid1" % k])
        kf1["nodes"], kf1["start"], kf1["idx"] = fake_feature_vector(da, nbits, seed=7)
        kf2 = dict(desc=db, angle=np.ascontiguousarray(kb["angle"]), valid=g["bk%d_valid2" % k])
        kf2["nodes"], kf2["start"], kf2["idx"] = fake_feature_vector(db, nbits, seed=7)
        yield k, kf1, kf2, float(nnr), bool(ori)


def test_search_by_bow_keyframes_equals_the_reference_matcher(oracle):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) (@0x82cc0, loop closing) executed from
    lib/libORB_SLAM2.so on two faked KeyFrames (fixture reference_library2.npz, bk*)."""
    g = np.load(os.path.join(G, "reference_library2.npz"))
    o = oracle.OrbOracle()
    for k, kf1, kf2, nnr, ori in _kfkf_cases(g, o.extract):
        m, n = oracle.search_by_bow_kfkf(kf1, kf2, nnr, ori)
        assert n == int(g["bk%d_n" % k]) > 150 and np.array_equal(m, g["bk%d_match" % k]), k


def _reloc_cases(g, extract, scale_factors):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from matchdata import relocalisation_case
    from plslam_b200.synth import synth_pair
    feats = {}
    for k in range(int(g["rk_n"])):
        seed, motion, th, od, ori = g["rk%d_args" % k]
        seed = int(seed)
        if seed not in feats:
            a, b = synth_pair(seed)
            feats[seed] = (extract(a), extract(b))
        (ka, da), (kb, db) = feats[seed]
        kf, cur, cam, sf, lsf, tcw = relocalisation_case(ka, da, kb, db, scale_factors, seed=seed, motion=float(motion))
        yield k, kf, cur, cam, sf, lsf, tcw, float(th), int(od), bool(ori)


def test_search_by_projection_keyframe_equals_the_reference_matcher(oracle):
    """ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist) (@0x7e8c0, relocalisation) executed
    from lib/libORB_SLAM2.so on a faked Frame / KeyFrame / MapPoints with a real std::set; MapPoint::PredictScale (logf of the C
    library), isBad, GetWorldPos and Frame::GetFeaturesInArea are the library's own (fixture reference_library2.npz, rk*)."""
    g = np.load(os.path.join(G, "reference_library2.npz"))
    o = oracle.OrbOracle()
    total = 0
    for k, kf, cur, cam, sf, lsf, tcw, th, od, ori in _reloc_cases(g, o.extract, o.tables()["scale"]):
        m, n = oracle.search_by_projection_kf(kf, cur, cam, sf, lsf, tcw, th, od, ori)
        assert n == int(g["rk%d_n" % k]) and np.array_equal(m, g["rk%d_match" % k]), k
        total += n
    assert total > 1500


def _loop_cases(g, extract, scale_factors):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from matchdata import loop_projection_case
    from plslam_b200.synth import synth_pair
    feats = {}
    for k in range(int(g["lc_n"])):
        seed, motion, scale, th = g["lc%d_args" % k]
        seed = int(seed)
        if seed not in feats:
            a, b = synth_pair(seed)
            feats[seed] = (extract(a), extract(b))
        (ka, da), (kb, db) = feats[seed]
        kf, mp, scw, mi = loop_projection_case(ka, da, kb, db, scale_factors, seed=seed, motion=float(motion), scale=float(scale))
        yield k, kf, mp, scw, mi, int(th)


def test_search_by_projection_sim3_equals_the_reference_matcher(oracle):
    """ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th) (@0x880f0, loop closing) executed from
    lib/libORB_SLAM2.so on a faked KeyFrame (real nested-vector grid) and faked MapPoints, similarity scales 0.6 / 1 / 1.7;
    KeyFrame::GetFeaturesInArea, IsInImage and MapPoint::PredictScale(dist, KeyFrame*) are the library's own (fixture lc*)."""
    g = np.load(os.path.join(G, "reference_library2.npz"))
    o = oracle.OrbOracle()
    total = 0
    for k, kf, mp, scw, mi, th in _loop_cases(g, o.extract, o.tables()["scale"]):
        m, n = oracle.search_by_projection_sim3(kf, mp, scw, mi, th)
        assert n == int(g["lc%d_n" % k]) and np.array_equal(m, g["lc%d_match" % k]), k
        total += n
    assert total > 1500


def _fuse_cases(g, extract, scale_factors):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from matchdata import fuse_case
    from plslam_b200.synth import synth_pair
    feats = {}
    for k in range(int(g["fu_n"])):
        seed, motion, th = g["fu%d_args" % k]
        seed = int(seed)
        if seed not in feats:
            a, b = synth_pair(seed)
            feats[seed] = (extract(a), extract(b))
        (ka, da), (kb, db) = feats[seed]
        kf, mp, kp_ = fuse_case(ka, da, kb, db, scale_factors, seed=seed, motion=float(motion))
        yield k, kf, mp, kp_, float(th)


def test_fuse_equals_the_reference_matcher(oracle):
    """ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (@0x7a500) executed from lib/libORB_SLAM2.so on a faked KeyFrame
    and faked MapPoints, its four map-graph callees replaced by logging stand-ins: the sequence of AddObservation / AddMapPoint /
    Replace calls (which point, which key-frame feature, which direction) and nFused must follow from the oracle's matching core
    plus the replay of the bookkeeping (fixture fu*)."""
    g = np.load(os.path.join(G, "reference_library2.npz"))
    o = oracle.OrbOracle()
    total = 0
    for k, kf, mp, kp_, th in _fuse_cases(g, o.extract, o.tables()["scale"]):
        best = oracle.fuse_search(kf, mp, th)
        nf, log = oracle.fuse_replay(best, mp, kp_)
        assert nf == int(g["fu%d_n" % k]), k
        assert np.array_equal(np.array(log, np.int32).reshape(-1, 3), g["fu%d_log" % k]), k
        total += nf
    assert total > 500


def _fuse_sim3_cases(g, extract, scale_factors):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from matchdata import fuse_case
    from plslam_b200.synth import synth_pair
    feats = {}
    for k in range(int(g["fs_n"])):
        seed, motion, scale, th = g["fs%d_args" % k]
        seed = int(seed)
        if seed not in feats:
            a, b = synth_pair(seed)
            feats[seed] = (extract(a), extract(b))
        (ka, da), (kb, db) = feats[seed]
        kf, mp, kp_ = fuse_case(ka, da, kb, db, scale_factors, seed=seed, motion=float(motion), scale=float(scale), sim3=True)
        yield k, kf, mp, kp_, float(th)


def test_fuse_sim3_equals_the_reference_matcher(oracle):
    """ORBmatcher::Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint) (@0x7bb20, LoopClosing::SearchAndFuse) executed from
    lib/libORB_SLAM2.so (same stand-ins as above; pKF->GetMapPoints() is the library's): call log, vpReplacePoint and nFused must
    follow from the oracle's matching core plus the replay (fixture fs*, similarity scales 0.7 / 1 / 1.6)."""
    g = np.load(os.path.join(G, "reference_library2.npz"))
    o = oracle.OrbOracle()
    total = 0
    for k, kf, mp, kp_, th in _fuse_sim3_cases(g, o.extract, o.tables()["scale"]):
        best = oracle.fuse_search_sim3(kf, mp, kf["scw"], th)
        nf, log, rep = oracle.fuse_replay_sim3(best, mp, kp_)
        assert nf == int(g["fs%d_n" % k]), k
        assert np.array_equal(np.array(log, np.int32).reshape(-1, 3), g["fs%d_log" % k]), k
        assert np.array_equal(rep, g["fs%d_replace" % k]), k
        total += nf
    assert total > 1000


def _sim3_cases(g, extract, scale_factors):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from matchdata import sim3_case
    from plslam_b200.synth import synth_pair
    feats = {}
    for k in range(int(g["s3_n"])):
        seed, s12, th = g["s3%d_args" % k]
        seed = int(seed)
        if seed not in feats:
            a, b = synth_pair(seed)
            feats[seed] = (extract(a), extract(b))
        (ka, da), (kb, db) = feats[seed]
        yield (k,) + sim3_case(ka, da, kb, db, scale_factors, seed=seed, s12=float(s12)) + (float(th),)


def test_search_by_sim3_equals_the_reference_matcher(oracle):
    """ORBmatcher::SearchBySim3 (@0x838b0) executed from lib/libORB_SLAM2.so on two faked KeyFrames with their own faked map
    points (three more matrix-expression shims: s * Mat, s * Mat.t(), -Mat): the mutually consistent matches it writes into
    vpMatches12 and nFound (fixture s3*, scales 0.8 / 1 / 1.3)."""
    g = np.load(os.path.join(G, "reference_library2.npz"))
    o = oracle.OrbOracle()
    total = 0
    for k, kf1, kf2, mp1, mp2, s12, R12, t12, mi, th in _sim3_cases(g, o.extract, o.tables()["scale"]):
        m, n = oracle.search_by_sim3(kf1, kf2, mp1, mp2, s12, R12, t12, th, mi)
        assert n == int(g["s3%d_n" % k]) and np.array_equal(m, g["s3%d_match" % k]), k
        total += n
    assert total > 500


def _frustum_cases(g):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from matchdata import frustum_case
    for k in range(int(g["fz_n"])):
        seed, motion, lim = g["fz%d_args" % k]
        c = frustum_case(3000, seed=int(seed), motion=float(motion))
        assert np.array_equal(c["ow"], g["fz%d_ow" % k])
        yield k, c, float(lim)


def test_is_in_frustum_equals_the_reference(oracle):
    """Frame::isInFrustum (@0xf5190) executed from lib/libORB_SLAM2.so on a faked Frame and 3 x 3000 faked MapPoints: the
    in-view flag, mTrackProjX / Y / XR, mnTrackScaleLevel and mTrackViewCos it leaves in each point (fixture fz*)."""
    g = np.load(os.path.join(G, "reference_library2.npz"))
    seen = 0
    for k, c, lim in _frustum_cases(g):
        r = oracle.is_in_frustum(c["xyz"], c["normal"], c["dist_range"], c["cam8"], c["tcw"], c["ow"], c["mbf"], c["log_sf"], c["n_levels"], lim)
        for name in ("in_view", "proj", "level", "viewcos"):
            assert np.array_equal(r[name], g["fz%d_%s" % (k, name)]), (k, name)
        seen += int(r["in_view"].sum())
    assert seen > 2000


def test_logf_and_predict_scale(oracle):
    """The restated glibc logf equals the C library's on a sweep of bit patterns (PredictScale calls logf, @0x8fc7b), and
    PredictScale clamps to [0, nLevels - 1]."""
    import ctypes as C
    L = oracle.lib()
    L.oracle_logf_mismatches.argtypes, L.oracle_logf_mismatches.restype = [C.c_uint32] * 3, C.c_int
    assert L.oracle_logf_mismatches(0x00800000, (0x7f800000 - 0x00800000) // 997, 997) == 0
    assert L.oracle_logf_mismatches(0x3f000000, 1 << 22, 3) == 0      # [0.5, ...) densely
    assert L.oracle_logf_mismatches(1, 5000, 1601) == 0               # subnormals
    lsf = float(np.log(np.float32(1.2)))
    assert oracle.predict_scale(10.0, 10.0, lsf, 8) == 0
    assert oracle.predict_scale(10.0, 20.0, lsf, 8) == 0              # ratio < 1: negative level clamps to 0
    assert oracle.predict_scale(10.0, 10.0 / 1.2 ** 3.5, lsf, 8) == 4
    assert oracle.predict_scale(10.0, 0.01, lsf, 8) == 7


def test_search_by_projection_equals_the_reference_matcher(oracle):
    """ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (@0x80d00, TrackWithMotionModel's matcher) executed from
    lib/libORB_SLAM2.so on faked Frame / MapPoint objects, its cv::Mat expressions evaluated with cv::gemm's arithmetic
    (fixture reference_library.npz, pj*): forward, backward and sideways motion, stereo and mono."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from matchdata import projection_case
    from plslam_b200.synth import synth_pair
    g = np.load(os.path.join(G, "reference_library.npz"))
    o = oracle.OrbOracle()
    sf = o.tables()["scale"]
    feats, total = {}, 0
    for k in range(int(g["pj_n"])):
        seed, motion, th, mono = g["pj%d_args" % k]
        seed = int(seed)
        if seed not in feats:
            a, b = synth_pair(seed)
            feats[seed] = (o.extract(a), o.extract(b))
        (ka, da), (kb, db) = feats[seed]
        last, cur, cam, _, tc, tl = projection_case(ka, da, kb, db, sf, seed=seed, motion=float(motion))
        m, n = oracle.search_by_projection(last, cur, cam, sf, tc, tl, float(th), bool(mono), True)
        assert n == int(g["pj%d_n" % k]) and np.array_equal(m, g["pj%d_match" % k]), k
        total += n
    assert total > 3000


def test_search_by_projection_window_edge_cases_equal_the_reference_matcher(oracle):
    """The same function on projection_boundary.npz: current key points planted within one ulp of the search window's edge
    (in x, and in the stereo coordinate), where an evaluation of the projection other than the binary's — float division
    @0x81c92, fused multiply-adds @0x81cba / @0x81cd9 / @0x81eb5, cv::gemm's float sums — decides differently (the double
    evaluation of round 1 fails every one of these cases).  Expected results: the reference's own code
    (tests/golden/make_projection_boundary.py)."""
    g = np.load(os.path.join(G, "projection_boundary.npz"))
    planted = 0
    for k in range(int(g["n"])):
        last = {n[len("c%d_last_" % k):]: g[n] for n in g.files if n.startswith("c%d_last_" % k)}
        cur = {n[len("c%d_cur_" % k):]: g[n] for n in g.files if n.startswith("c%d_cur_" % k)}
        a = g["c%d_args" % k]
        m, n = oracle.search_by_projection(last, cur, g["c%d_cam" % k], g["c%d_sf" % k], g["c%d_tc" % k], g["c%d_tl" % k],
                                           float(a[2]), bool(a[3]), True)
        assert n == int(g["c%d_n" % k]) and np.array_equal(m, g["c%d_match" % k]), k
        planted += int(a[4])
    assert planted > 2000


def test_search_for_triangulation_equals_the_reference_matcher(oracle):
    """ORBmatcher::SearchForTriangulation (@0x86b30) executed from lib/libORB_SLAM2.so on faked KeyFrame objects — the reference
    computes its own epipole through KeyFrame's pose getters — against the oracle fed with oracle.epipole (fixture
    reference_library.npz, tr*).  The CUDA kernel is compared with the same oracle in tests/test_triangulation_gpu.py."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from matchdata import triangulation_case
    from plslam_b200.synth import synth_frame
    g = np.load(os.path.join(G, "reference_library.npz"))
    cases, total = {}, 0
    for k in range(int(g["tr_n"])):
        a = g["tr%d_args" % k]
        seed, only, ori, nbits = int(a[0]), bool(a[1]), bool(a[2]), int(a[4])
        key = (seed, float(a[3]), nbits, tuple(a[5:8]))
        if key not in cases:
            kps, desc = oracle.OrbOracle().extract(synth_frame(seed))
            cases[key] = triangulation_case(kps, desc, seed=seed, stereo_fraction=float(a[3]), nbits=nbits, t21=tuple(a[5:8]))
        kf1, kf2, F12, pose, cam, sf, sg = cases[key]
        ex, ey = oracle.epipole(*pose, *cam)
        m, n = oracle.search_for_triangulation(kf1, kf2, F12, ex, ey, sf, sg, only, ori)
        assert n == int(g["tr%d_n" % k]) and np.array_equal(m, g["tr%d_match" % k]), k
        total += n
    assert total > 1500


def test_rgbd_stereo_equals_the_reference_frame_code(oracle):
    """Frame::ComputeStereoFromRGBD (@0xf6860) executed from lib/libORB_SLAM2.so on a faked Frame (fixture st*): depth is read
    at the truncated DISTORTED keypoint, mvuRight = undistorted x - mbf / d where d > 0, -1 elsewhere."""
    g = np.load(os.path.join(G, "reference_library.npz"))
    xy, un, depth = g["st_xy"], g["st_un"], g["st_depth"].astype(np.float32)
    # frame_post undistorts itself; with k1 == 0 it copies, so feed the undistorted points through a second call for the grid
    # only and check the stereo part on its own inputs: depth lookup uses xy, the subtraction uses un
    calib = dict(fx=500.0, fy=500.0, cx=320.0, cy=240.0, k1=0.0, k2=0.0, p1=0.0, p2=0.0, k3=0.0, bf=40.0)
    fp = oracle.frame_post(calib, np.array([0, 640, 0, 480], np.float32), xy, depth)
    assert np.array_equal(fp["depth"].view(np.uint32), g["st_z"].view(np.uint32))
    has = g["st_z"] > 0
    assert 1500 < int(has.sum()) < len(xy)
    want = np.where(has, un[:, 0] - np.float32(40.0) / np.where(has, g["st_z"], 1).astype(np.float32), np.float32(-1)).astype(np.float32)
    assert np.array_equal(want.view(np.uint32), g["st_uright"].view(np.uint32))   # the reference's arithmetic, restated in numpy
    # and the oracle's own subtraction, on points whose undistorted position equals the distorted one
    assert np.array_equal(fp["uright"][has].view(np.uint32), (xy[has, 0] - np.float32(40.0) / g["st_z"][has]).astype(np.float32).view(np.uint32))


def test_search_for_initialization_equals_the_reference_matcher(oracle):
    """ORBmatcher::SearchForInitialization (@0x7db00) executed from lib/libORB_SLAM2.so on faked Frames (fixture si*): matches,
    count and the updated vbPrevMatched."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from matchdata import frame_grid
    from plslam_b200.synth import synth_pair
    g = np.load(os.path.join(G, "reference_library.npz"))
    o = oracle.OrbOracle()
    feats = {}
    for k in range(int(g["si_n"])):
        seed, win, nnr, ori = g["si%d_args" % k]
        seed = int(seed)
        if seed not in feats:
            a, b = synth_pair(seed)
            (ka, da), (kb, db) = o.extract(a), o.extract(b)
            mk = lambda kk, d: dict(xy=np.stack([kk["x"], kk["y"]], 1).astype(np.float32), octave=kk["octave"].astype(np.int32),
                                    angle=kk["angle"].astype(np.float32), desc=d)
            f1, f2 = mk(ka, da), mk(kb, db)
            gs, gi, (mnx, mxx, mny, mxy, gwi, ghi) = frame_grid(f2["xy"], 640, 480)
            f2["grid_start"], f2["grid_items"] = gs, gi
            feats[seed] = (f1, f2, np.array([mnx, mny, gwi, ghi], np.float32))
        f1, f2, cam4 = feats[seed]
        m, n, prev = oracle.search_for_initialization(f1, f2, cam4, f1["xy"].copy(), int(win), float(nnr), bool(ori))
        assert n == int(g["si%d_n" % k]) > 50 and np.array_equal(m, g["si%d_match" % k]), k
        assert np.array_equal(prev, g["si%d_prev" % k]), k


def test_undistortion_and_bounds_equal_the_reference_frame_code(oracle):
    """Frame::UndistortKeyPoints (@0xf8630) and Frame::ComputeImageBounds (@0xf6010) executed from lib/libORB_SLAM2.so on a faked
    Frame (fixture un*; cv::undistortPoints served by the cv2-pinned restatement): the k1 == 0 shortcut, the N x 2 / 2-channel
    round trip, the corner order and min/max choices of the bounds."""
    g = np.load(os.path.join(G, "reference_library.npz"))
    for k in range(int(g["un_n"])):
        c = dict(zip(("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3"), (float(v) for v in g["un%d_calib" % k])), bf=40.0)
        assert np.array_equal(oracle.undistort_points(c, g["un%d_xy" % k]).view(np.uint32), g["un%d_out" % k].view(np.uint32)), k
        assert np.array_equal(oracle.image_bounds(c, 640, 480), g["un%d_bounds" % k]), k
    assert np.array_equal(g["un1_out"], g["un1_xy"])  # the undistorted calibration copies
