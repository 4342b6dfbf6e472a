"""GPU parity of the Hamming matching kernels (through the C-ABI) against the CPU oracle: indices, Hamming
scores and match counts must be identical."""
import numpy as np
import pytest

from matchdata import fake_feature_vector, projection_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def feats(oracle):
    from plslam_b200.synth import synth_pair
    orc = oracle.OrbOracle()
    out = []
    for seed in (0, 1, 2):
        a, b = synth_pair(seed)
        out.append((orc.extract(a), orc.extract(b)))
    return out, orc.tables()["scale"]


def test_knn2_matches_oracle_and_descriptor_distance(oracle, feats):
    import torch
    import plslam_b200 as pl
    rng = np.random.default_rng(0)
    (ka, da), (kb, db) = feats[0][0]
    assert np.array_equal(oracle.knn2(da, db), pl.knn2_host(da, db))
    for _ in range(50):
        i, j = rng.integers(len(da)), rng.integers(len(db))
        assert pl.DescriptorDistance(da[i], db[j]) == oracle.descriptor_distance(da[i], db[j]) == \
            int(np.unpackbits(da[i] ^ db[j]).sum())
    # ragged / tiny / duplicate-heavy cases, batched on device
    cases = [(rng.integers(0, 256, (nq, 32)).astype(np.uint8), rng.integers(0, 4, (nt, 32)).astype(np.uint8))
             for nq, nt in ((1, 1), (5, 2), (300, 1500), (1025, 1030), (40, 40), (3, 0))]
    pairs = []
    for q, t in cases:
        tq, tt = torch.from_numpy(q).cuda(), torch.from_numpy(t.reshape(-1, 32)).cuda()
        pairs.append((tq, tt, torch.empty((len(q), 4), dtype=torch.int32, device="cuda")))
    pl.knn2_batch_device(pairs)
    torch.cuda.synchronize()
    for (q, t), (_, _, o) in zip(cases, pairs):
        assert np.array_equal(oracle.knn2(q, t), o.cpu().numpy())


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_search_by_bow_matches_oracle(oracle, feats):
    import torch
    import plslam_b200 as pl
    jobs, keep, expect = [], [], []
    for s, ((ka, da), (kb, db)) in enumerate(feats[0]):
        rng = np.random.default_rng(s)
        for nnratio, ori in ((0.7, True), (0.9, False)):
            kf = dict(desc=da, angle=np.ascontiguousarray(ka["angle"]), valid=(rng.random(len(da)) < 0.85).astype(np.uint8))
            kf["nodes"], kf["start"], kf["idx"] = fake_feature_vector(da, seed=7)
            f = dict(desc=db, angle=np.ascontiguousarray(kb["angle"]))
            f["nodes"], f["start"], f["idx"] = fake_feature_vector(db, seed=7)
            expect.append(oracle.search_by_bow(kf, f, nnratio, ori))
            d = {k: _dev(v) for k, v in kf.items()}
            e = {k: _dev(v) for k, v in f.items()}
            m = torch.empty(len(db), dtype=torch.int32, device="cuda"); n = torch.zeros(1, dtype=torch.int32, device="cuda")
            keep.append((d, e, m, n))
            jobs.append(pl.BowJob(d["desc"].data_ptr(), d["angle"].data_ptr(), d["valid"].data_ptr(), d["nodes"].data_ptr(),
                                  d["start"].data_ptr(), d["idx"].data_ptr(), e["desc"].data_ptr(), e["angle"].data_ptr(),
                                  e["nodes"].data_ptr(), e["start"].data_ptr(), e["idx"].data_ptr(), m.data_ptr(), n.data_ptr(),
                                  len(da), len(db), len(kf["nodes"]), len(f["nodes"]), nnratio, int(ori)))
    pl.bow_batch_device(jobs, max(max(j.n1, j.n2) for j in jobs), "cuda")
    torch.cuda.synchronize()
    total = 0
    for (em, en), (_, _, m, n) in zip(expect, keep):
        assert int(n) == en
        assert np.array_equal(em, m.cpu().numpy())
        total += en
    assert total > 100


def test_search_by_projection_matches_oracle(oracle, feats):
    import torch
    import plslam_b200 as pl
    fs, scale = feats
    jobs, keep, expect = [], [], []
    for s, ((ka, da), (kb, db)) in enumerate(fs):
        for th, mono, ori, motion in ((15.0, False, True, 0.02), (30.0, False, True, 0.3), (7.0, True, False, 0.0), (15.0, False, True, -0.3)):
            last, cur, cam, sf, tc, tl = projection_case(ka, da, kb, db, scale, seed=s, motion=motion)
            expect.append(oracle.search_by_projection(last, cur, cam, sf, tc, tl, th, mono, ori))
            L = {k: _dev(v) for k, v in last.items()}
            Cc = {k: _dev(v) for k, v in cur.items()}
            sfd = _dev(sf)
            m = torch.empty(len(db), dtype=torch.int32, device="cuda"); n = torch.zeros(1, dtype=torch.int32, device="cuda")
            keep.append((L, Cc, sfd, m, n))
            j = pl.ProjJob(L["valid"].data_ptr(), L["xyz"].data_ptr(), L["desc"].data_ptr(), L["octave"].data_ptr(),
                           L["angle"].data_ptr(), L["obs"].data_ptr(), Cc["xy"].data_ptr(), Cc["octave"].data_ptr(),
                           Cc["angle"].data_ptr(), Cc["desc"].data_ptr(), Cc["uright"].data_ptr(), Cc["taken"].data_ptr(),
                           Cc["grid_start"].data_ptr(), Cc["grid_items"].data_ptr(), sfd.data_ptr(), m.data_ptr(), n.data_ptr())
            j.cam[:] = cam.tolist(); j.tcw_cur[:] = tc.ravel().tolist(); j.tcw_last[:] = tl.ravel().tolist()
            j.th = th; j.n1 = len(da); j.n2 = len(db); j.mono = int(mono); j.check_orientation = int(ori)
            jobs.append(j)
    pl.projection_batch_device(jobs, max(j.n1 for j in jobs), max(j.n2 for j in jobs), "cuda")
    torch.cuda.synchronize()
    total = 0
    for (em, en), (_, _, _, m, n) in zip(expect, keep):
        assert int(n) == en
        assert np.array_equal(em, m.cpu().numpy())
        total += en
    assert total > 200


def test_search_local_points_matches_oracle(oracle, feats):
    """ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) (Tracking::SearchLocalPoints): batched device form and the
    host form against the sequential oracle, th = 1 (no factor) and th = 3 (after relocalisation), two nnratio values."""
    import torch
    import plslam_b200 as pl
    from matchdata import local_points_case
    fs, scale = feats
    jobs, keep, expect = [], [], []
    for s, ((ka, da), (kb, db)) in enumerate(fs):
        for th, nnr, jitter in ((1.0, 0.8, 2.0), (3.0, 0.8, 6.0), (1.0, 0.6, 1.0), (5.0, 0.9, 10.0)):
            mp, fr, cam4 = local_points_case(ka, da, kb, db, seed=10 * s + int(th), jitter=jitter)
            expect.append(oracle.search_local_points(mp, fr, cam4, scale, th, nnr))
            Mp = {k: _dev(v) for k, v in mp.items()}
            Fr = {k: _dev(v) for k, v in fr.items()}
            sfd = _dev(scale)
            m = torch.empty(len(db), dtype=torch.int32, device="cuda"); n = torch.zeros(1, dtype=torch.int32, device="cuda")
            keep.append((Mp, Fr, sfd, m, n))
            j = pl.LocalJob(Mp["valid"].data_ptr(), Mp["proj"].data_ptr(), Mp["level"].data_ptr(), Mp["viewcos"].data_ptr(),
                            Mp["desc"].data_ptr(), Mp["obs"].data_ptr(), Fr["xy"].data_ptr(), Fr["octave"].data_ptr(),
                            Fr["desc"].data_ptr(), Fr["uright"].data_ptr(), Fr["taken"].data_ptr(), Fr["grid_start"].data_ptr(),
                            Fr["grid_items"].data_ptr(), sfd.data_ptr(), m.data_ptr(), n.data_ptr())
            j.cam[:] = cam4.tolist(); j.th = th; j.nnratio = nnr; j.m = len(da); j.n = len(db)
            jobs.append(j)
            if s == 0:
                hm, hn = pl.search_local_points_host(mp, fr, cam4, scale, th, nnr)
                assert hn == expect[-1][1] and np.array_equal(hm, expect[-1][0])
    pl.local_points_batch_device(jobs, max(j.n for j in jobs), "cuda")
    torch.cuda.synchronize()
    total = 0
    for (em, en), (_, _, _, m, n) in zip(expect, keep):
        assert int(n) == en
        assert np.array_equal(em, m.cpu().numpy())
        total += en
    assert total > 300
