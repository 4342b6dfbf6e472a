"""Pins the bag-of-words oracle and the Hamming distance against the reference's OWN DBoW2 sources, compiled
unmodified into oracle/_ref/libdbow2_ref.so (FORB.cpp:82-102, TemplatedVocabulary.h:1151-1284,1362-1448)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def ref(oracle):
    if oracle.dbow2_ref() is None:
        pytest.skip("oracle/_ref/libdbow2_ref.so not built (needs /root/reference)")
    return oracle


def test_descriptor_distance_equals_reference_forb_distance(ref):
    import plslam_b200 as pl
    rng = np.random.default_rng(0)
    for _ in range(300):
        a, b = rng.integers(0, 256, 32).astype(np.uint8), rng.integers(0, 256, 32).astype(np.uint8)
        d = ref.VocReference.forb_distance(a, b)
        assert d == ref.descriptor_distance(a, b) == pl.DescriptorDistance(a, b)


@pytest.mark.parametrize("k,L,early", [(10, 3, 0.0), (4, 5, 0.0), (6, 4, 0.2)])
def test_transform_equals_reference_dbow2(ref, tmp_path, k, L, early):
    path = str(tmp_path / "voc.txt")
    ref.write_vocabulary_text(path, k, L, seed=k * 10 + L, leaf_fraction_early=early)
    mine, theirs = ref.VocOracle(path), ref.VocReference(path)
    assert mine.n_words == theirs.n_words
    rng = np.random.default_rng(1)
    # with leaves above level L the reference leaves `nid` unassigned (an uninitialised local, TemplatedVocabulary.h:1178)
    # whenever the descent stops before level L - levelsup; only levels every descent reaches are compared there
    cases = ((1, 1), (500, 4), (1000, 2), (37, 0), (200, 9)) if early == 0.0 else ((300, L - 1), (300, L), (50, 9))
    for n, levelsup in cases:
        desc = rng.integers(0, 256, (n, 32)).astype(np.uint8)
        a, b = mine.transform(desc, levelsup), theirs.transform(desc, levelsup)
        for key in a:
            assert np.array_equal(a[key], b[key]), (key, n, levelsup)
        assert abs(a["bow_vals"].sum() - 1.0) < 1e-12


def test_transform_on_real_orb_features(ref, tmp_path):
    from plslam_b200.synth import synth_frame
    path = str(tmp_path / "voc.txt")
    ref.write_vocabulary_text(path, 10, 4, seed=3)
    mine, theirs = ref.VocOracle(path), ref.VocReference(path)
    k, d = ref.OrbOracle().extract(synth_frame(2))
    a, b = mine.transform(d, 2), theirs.transform(d, 2)
    for key in a:
        assert np.array_equal(a[key], b[key]), key
    assert len(a["fv_idx"]) == len(d)


def test_trailing_empty_line_creates_the_same_phantom_node(ref, tmp_path):
    """ORBvoc.txt ends with '\\n'; the reference's loader then appends one more node whose parent / leaf flag are the
    previous line's (unassigned locals) and whose weight is 0 (TemplatedVocabulary.h:1401-1443)."""
    path = str(tmp_path / "voc.txt")
    ref.write_vocabulary_text(path, 5, 3, seed=9)
    with open(path, "a") as f:
        f.write("\n")
    mine, theirs = ref.VocOracle(path), ref.VocReference(path)
    assert mine.n_words == theirs.n_words
    rng = np.random.default_rng(4)
    desc = rng.integers(0, 256, (3000, 32)).astype(np.uint8)
    desc[:200] &= rng.integers(0, 256, (200, 32)).astype(np.uint8) & rng.integers(0, 256, (200, 32)).astype(np.uint8)  # near-zero rows
    a, b = mine.transform(desc, 1), theirs.transform(desc, 1)
    for key in a:
        assert np.array_equal(a[key], b[key]), key


REAL_VOC = "/root/reference/Vocabulary/ORBvoc.txt.tar.gz"


@pytest.mark.skipif(not __import__("os").path.exists(REAL_VOC), reason="the reference's ORBvoc archive is not on this machine")
def test_real_orbvoc_equals_reference_dbow2(ref, tmp_path):
    """The shipped vocabulary (k=10, L=6, 1,082,073 nodes + the phantom), levelsup=4 as Frame::ComputeBoW uses (Frame.cc:354-362)."""
    import tarfile
    with tarfile.open(REAL_VOC) as t:
        t.extract("ORBvoc.txt", tmp_path)
    path = str(tmp_path / "ORBvoc.txt")
    mine, theirs = ref.VocOracle(path), ref.VocReference(path)
    assert mine.n_words == theirs.n_words == 971815
    from plslam_b200.synth import synth_frame
    o = ref.OrbOracle()
    rng = np.random.default_rng(0)
    sets = [o.extract(synth_frame(s))[1] for s in range(3)] + [rng.integers(0, 256, (5000, 32)).astype(np.uint8)]
    for d in sets:
        a, b = mine.transform(d, 4), theirs.transform(d, 4)
        for key in a:
            assert np.array_equal(a[key], b[key]), key
