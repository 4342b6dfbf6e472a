"""GPU parity of the vocabulary descent (plslam_voc_*) against the oracle (which tests/test_bow_cpu.py pins against the
reference's own DBoW2 build), the blob export/import used for the start-up broadcast, and SearchByBoW fed with the
resulting real FeatureVectors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def voc(oracle, tmp_path_factory):
    path = str(tmp_path_factory.mktemp("voc") / "voc.txt")
    oracle.write_vocabulary_text(path, 10, 4, seed=5)
    return path


def test_transform_matches_oracle_and_reference(oracle, voc):
    import torch
    import plslam_b200 as pl
    from plslam_b200.synth import synth_frame
    V = pl.ORBVocabulary(voc)
    O = oracle.VocOracle(voc)
    assert (V.k, V.L, V.n_nodes, V.n_words) == (10, 4, O.n_nodes, O.n_words)
    rng = np.random.default_rng(0)
    k, d = oracle.OrbOracle().extract(synth_frame(3))
    for desc, levelsup in ((d, 2), (rng.integers(0, 256, (5000, 32)).astype(np.uint8), 4), (d[:1], 0), (d[:7], 9)):
        w, wt, nd = V.transform_features(desc, levelsup)
        ow, owt, ond = O.transform_features(desc, levelsup)
        assert np.array_equal(w, ow) and np.array_equal(wt, owt) and np.array_equal(nd, ond)
        a, b = V.transform(desc, levelsup), O.transform(desc, levelsup)
        for key in a:
            assert np.array_equal(a[key], b[key]), key
        if oracle.dbow2_ref() is not None:
            c = oracle.VocReference(voc).transform(desc, levelsup)
            for key in a:
                assert np.array_equal(a[key], c[key]), "vs reference DBoW2: " + key
    # blob round trip (what rank 0 broadcasts at start-up)
    V2 = pl.ORBVocabulary.from_blob(V.export_blob())
    w2, wt2, nd2 = V2.transform_features(d, 2)
    ow, owt, ond = O.transform_features(d, 2)
    assert np.array_equal(w2, ow) and np.array_equal(nd2, ond)
    # device-resident entry point
    dw, dwt, dnd = V.transform_features_device(torch.from_numpy(d).cuda(), 2)
    assert np.array_equal(dw.cpu().numpy(), ow) and np.array_equal(dwt.cpu().numpy(), owt)


def test_trailing_empty_line_phantom_node(oracle, tmp_path):
    """The shipped ORBvoc.txt ends with a newline; the product loader must create the same phantom node the reference's does."""
    import plslam_b200 as pl
    path = str(tmp_path / "voc.txt")
    oracle.write_vocabulary_text(path, 5, 3, seed=9)
    with open(path, "a") as f:
        f.write("\n")
    V, O = pl.ORBVocabulary(path), oracle.VocOracle(path)
    assert (V.n_nodes, V.n_words) == (O.n_nodes, O.n_words)
    rng = np.random.default_rng(4)
    desc = rng.integers(0, 256, (3000, 32)).astype(np.uint8)
    desc[:200] &= rng.integers(0, 256, (200, 32)).astype(np.uint8) & rng.integers(0, 256, (200, 32)).astype(np.uint8)
    a, b = V.transform(desc, 1), O.transform(desc, 1)
    for key in a:
        assert np.array_equal(a[key], b[key]), key


def test_search_by_bow_with_real_feature_vectors(oracle, voc):
    import torch
    import plslam_b200 as pl
    from plslam_b200.synth import synth_pair
    V = pl.ORBVocabulary(voc)
    orc = oracle.OrbOracle()
    a, b = synth_pair(4)
    (ka, da), (kb, db) = orc.extract(a), orc.extract(b)
    fa, fb = V.transform(da, 2), V.transform(db, 2)
    kf = dict(desc=da, angle=np.ascontiguousarray(ka["angle"]), valid=np.ones(len(da), np.uint8), nodes=fa["fv_nodes"].astype(np.int32),
              start=fa["fv_start"], idx=fa["fv_idx"].astype(np.int32))
    f = dict(desc=db, angle=np.ascontiguousarray(kb["angle"]), nodes=fb["fv_nodes"].astype(np.int32), start=fb["fv_start"],
             idx=fb["fv_idx"].astype(np.int32))
    em, en = oracle.search_by_bow(kf, f, 0.7, True)
    dv = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    D = {k: dv(v) for k, v in kf.items()}
    E = {k: dv(v) for k, v in f.items()}
    m = torch.empty(len(db), dtype=torch.int32, device="cuda"); n = torch.zeros(1, dtype=torch.int32, device="cuda")
    job = pl.BowJob(D["desc"].data_ptr(), D["angle"].data_ptr(), D["valid"].data_ptr(), D["nodes"].data_ptr(), D["start"].data_ptr(),
                    D["idx"].data_ptr(), E["desc"].data_ptr(), E["angle"].data_ptr(), E["nodes"].data_ptr(), E["start"].data_ptr(),
                    E["idx"].data_ptr(), m.data_ptr(), n.data_ptr(), len(da), len(db), len(kf["nodes"]), len(f["nodes"]), 0.7, 1)
    pl.bow_batch_device([job], max(len(da), len(db)), "cuda")
    torch.cuda.synchronize()
    assert int(n) == en and np.array_equal(m.cpu().numpy(), em) and en > 30


def test_c4_device_chain_extract_bow_search(oracle, voc):
    """Config C4 without host round trips: batched extraction -> ComputeBoW (descent + FeatureVector CSR on the device) ->
    SearchByBoW on the frame pairs of the batch, against the oracle chain (ORB oracle -> vocabulary oracle -> SearchByBoW
    oracle) pair by pair."""
    import torch
    import plslam_b200 as pl
    from plslam_b200.synth import synth_pair
    V = pl.ORBVocabulary(voc)
    O = oracle.VocOracle(voc)
    orc = oracle.OrbOracle()
    pairs = [synth_pair(20 + s) for s in range(3)]
    imgs = np.stack([im for pr in pairs for im in pr])
    imgs[5] = 128  # a pair whose second frame has no features
    ex = pl.ORBextractor()
    d_kps, d_desc, d_cnt = ex.extract_batch_device(torch.from_numpy(imgs).cuda())
    fv = V.featvec_batch_device(d_desc, d_cnt, levelsup=2)
    bv = V.bowvec_batch_device(fv, d_cnt)  # Frame::mBowVec of every frame, accumulated and normalised on the device
    res = pl.bow_pairs_device(d_kps, d_desc, d_cnt, fv, nnratio=0.7, check_ori=True)
    torch.cuda.synchronize()
    cnt = d_cnt.cpu().numpy()
    feats = []
    for f in range(6):
        k, d = orc.extract(imgs[f])
        n = int(cnt[f])
        assert n == len(k)
        t = O.transform(d, 2)
        nn = int(fv["fv_count"][f])
        assert nn == len(t["fv_nodes"])
        assert np.array_equal(fv["fv_nodes"][f, :nn].cpu().numpy(), t["fv_nodes"].astype(np.int32))
        assert np.array_equal(fv["fv_start"][f, :nn + 1].cpu().numpy(), t["fv_start"])
        m = int(t["fv_start"][-1])
        assert np.array_equal(fv["fv_idx"][f, :m].cpu().numpy(), t["fv_idx"].astype(np.int32))
        nb = int(bv["bow_count"][f])
        assert nb == len(t["bow_ids"])
        assert np.array_equal(bv["bow_ids"][f, :nb].cpu().numpy(), t["bow_ids"].astype(np.int32))
        assert np.array_equal(bv["bow_vals"][f, :nb].cpu().numpy(), t["bow_vals"])  # bit-identical doubles
        feats.append((k, d, t))
    total = 0
    for p in range(3):
        (ka, da, ta), (kb, db, tb) = feats[2 * p], feats[2 * p + 1]
        kf = dict(desc=da, angle=np.ascontiguousarray(ka["angle"]), valid=np.ones(len(da), np.uint8),
                  nodes=ta["fv_nodes"].astype(np.int32), start=ta["fv_start"], idx=ta["fv_idx"].astype(np.int32))
        f = dict(desc=db, angle=np.ascontiguousarray(kb["angle"]), nodes=tb["fv_nodes"].astype(np.int32), start=tb["fv_start"],
                 idx=tb["fv_idx"].astype(np.int32))
        if len(db) == 0:
            assert int(res["nmatches"][p]) == 0
            continue
        em, en = oracle.search_by_bow(kf, f, 0.7, True)
        assert int(res["nmatches"][p]) == en
        assert np.array_equal(res["match"][p, :len(db)].cpu().numpy(), em)
        total += en
    assert total > 60
