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
