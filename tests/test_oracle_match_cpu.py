"""CPU tests of the matcher oracle against independent numpy restatements (known-answer material: FORB::distance,
TH_LOW/TH_HIGH/HISTO_LENGTH) and of the synthetic-data helpers."""
import numpy as np

from matchdata import fake_feature_vector, frame_grid, projection_case


def test_knn2_against_numpy(oracle):
    rng = np.random.default_rng(1)
    q = rng.integers(0, 256, (64, 32)).astype(np.uint8)
    t = rng.integers(0, 8, (50, 32)).astype(np.uint8)  # many ties
    d = np.unpackbits(q[:, None, :] ^ t[None, :, :], axis=2).sum(2)
    got = oracle.knn2(q, t)
    for i in range(len(q)):
        order = np.argsort(d[i], kind="stable")
        assert got[i].tolist() == [order[0], d[i][order[0]], order[1], d[i][order[1]]]
    assert oracle.knn2(q[:3], t[:1])[:, 2:].tolist() == [[-1, -1]] * 3
    assert oracle.knn2(q[:2], t[:0]).tolist() == [[-1, -1, -1, -1]] * 2


def test_bow_and_projection_oracles_are_consistent(oracle):
    from plslam_b200.synth import synth_pair
    orc = oracle.OrbOracle()
    a, b = synth_pair(5)
    (ka, da), (kb, db) = orc.extract(a), orc.extract(b)
    kf = dict(desc=da, angle=np.ascontiguousarray(ka["angle"]), valid=np.ones(len(da), np.uint8))
    kf["nodes"], kf["start"], kf["idx"] = fake_feature_vector(da, seed=3)
    f = dict(desc=db, angle=np.ascontiguousarray(kb["angle"]))
    f["nodes"], f["start"], f["idx"] = fake_feature_vector(db, seed=3)
    m, n = oracle.search_by_bow(kf, f, 0.7, True)
    assert n == int((m >= 0).sum()) and n > 50
    # every accepted match shares a node and satisfies TH_LOW
    node_kf = np.empty(len(da), np.int32); node_f = np.empty(len(db), np.int32)
    for k in range(len(kf["nodes"])):
        node_kf[kf["idx"][kf["start"][k]:kf["start"][k + 1]]] = kf["nodes"][k]
    for k in range(len(f["nodes"])):
        node_f[f["idx"][f["start"][k]:f["start"][k + 1]]] = f["nodes"][k]
    for i2 in np.nonzero(m >= 0)[0]:
        assert node_f[i2] == node_kf[m[i2]]
        assert oracle.descriptor_distance(da[m[i2]], db[i2]) <= 50
    # without the orientation filter there can only be more matches
    assert oracle.search_by_bow(kf, f, 0.7, False)[1] >= n
    last, cur, cam, sf, tc, tl = projection_case(ka, da, kb, db, orc.tables()["scale"], seed=1)
    mp, npj = oracle.search_by_projection(last, cur, cam, sf, tc, tl, 15.0, False, False)
    assert npj > 100
    for i2 in np.nonzero(mp >= 0)[0]:
        assert last["valid"][mp[i2]] and not cur["taken"][i2]
        assert oracle.descriptor_distance(da[mp[i2]], db[i2]) <= 100


def test_frame_grid_is_a_partition():
    rng = np.random.default_rng(2)
    xy = np.stack([rng.random(500) * 639, rng.random(500) * 479], 1).astype(np.float32)
    start, items, _ = frame_grid(xy, 640, 480)
    assert start[-1] == len(items) <= 500 and len(start) == 64 * 48 + 1
    assert len(set(items.tolist())) == len(items)


def test_search_for_triangulation_oracle_properties(oracle):
    """Structure of ORBmatcher::SearchForTriangulation's result on synthetic two-view geometry (the arithmetic of its leaves is
    pinned against the reference's machine code in test_golden_cpu.py)."""
    import plslam_b200 as pl
    from plslam_b200.synth import synth_frame
    from matchdata import triangulation_case
    kps, desc = oracle.OrbOracle().extract(synth_frame(21))
    kf1, kf2, F12, (R2w, t2w, Cw), (fx, fy, cx, cy), sf, sg = triangulation_case(kps, desc, seed=1)
    ex, ey = oracle.epipole(R2w, t2w, Cw, fx, fy, cx, cy)
    assert (ex, ey) == pl.epipole(R2w, t2w, Cw, fx, fy, cx, cy)          # host scalar code of the product, no GPU needed
    assert abs(ex - (fx * t2w[0] / t2w[2] + cx)) < 1e-2
    node_of = lambda kf: {int(i): int(kf["nodes"][k]) for k in range(len(kf["nodes"])) for i in kf["idx"][kf["start"][k]:kf["start"][k + 1]]}
    nd1, nd2 = node_of(kf1), node_of(kf2)
    for only_stereo in (False, True):
        m, n = oracle.search_for_triangulation(kf1, kf2, F12, ex, ey, sf, sg, only_stereo=only_stereo, check_ori=True)
        m0, n0 = oracle.search_for_triangulation(kf1, kf2, F12, ex, ey, sf, sg, only_stereo=only_stereo, check_ori=False)
        idx = np.flatnonzero(m >= 0)
        assert n == len(idx) and n0 == int((m0 >= 0).sum()) and 30 < n <= n0
        assert len(set(m[idx].tolist())) == len(idx)                       # a KF2 feature is matched once (vbMatched2)
        assert set(idx.tolist()) <= set(np.flatnonzero(m0 >= 0).tolist())   # the rotation filter only removes
        for i in idx:
            j = int(m[i])
            assert not kf1["has_mp"][i] and not kf2["has_mp"][j] and nd1[int(i)] == nd2[j]
            assert oracle.descriptor_distance(kf1["desc"][i], kf2["desc"][j]) <= 50
            assert oracle.check_dist_epipolar_line(kf1["xy"][i], kf2["xy"][j], F12, sg[kf2["octave"][j]])
            if only_stereo:
                assert kf1["uright"][i] >= 0 and kf2["uright"][j] >= 0
            elif kf1["uright"][i] < 0 and kf2["uright"][j] < 0:
                d = kf2["xy"][j] - np.array([ex, ey], np.float32)
                assert float(d @ d) >= 100 * sf[kf2["octave"][j]] * (1 - 1e-5)
