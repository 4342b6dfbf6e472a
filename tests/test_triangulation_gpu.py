"""ORBmatcher::SearchForTriangulation on the GPU (k_triangulation through plslam_match_triangulation_host) against the
oracle, whose leaf arithmetic is pinned against the reference's machine code (tests/test_golden_cpu.py).  Bit-exact: match
indices and counts."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _case(oracle, seed, **kw):
    from plslam_b200.synth import synth_frame
    from matchdata import triangulation_case
    kps, desc = oracle.OrbOracle().extract(synth_frame(seed))
    return triangulation_case(kps, desc, seed=seed, **kw)


@pytest.mark.parametrize("seed,kw", [(21, {}), (22, dict(stereo_fraction=0.0, t21=(0.005, 0.002, 0.15))),
                                     (23, dict(stereo_fraction=1.0)), (24, dict(nbits=1)), (25, dict(nbits=6, stereo_fraction=0.3))])
def test_search_for_triangulation_equals_oracle(oracle, seed, kw):
    import plslam_b200 as pl
    kf1, kf2, F12, (R2w, t2w, Cw), (fx, fy, cx, cy), sf, sg = _case(oracle, seed, **kw)
    ex, ey = pl.epipole(R2w, t2w, Cw, fx, fy, cx, cy)
    total = 0
    for only_stereo in (False, True):
        for check_ori in (True, False):
            want, nwant = oracle.search_for_triangulation(kf1, kf2, F12, ex, ey, sf, sg, only_stereo, check_ori)
            got, ngot, pairs = pl.search_for_triangulation_host(kf1, kf2, F12, ex, ey, sf, sg, only_stereo, check_ori)
            assert ngot == nwant and np.array_equal(got, want), (seed, only_stereo, check_ori)
            assert pairs == [(int(i), int(v)) for i, v in enumerate(want) if v >= 0]
            total += nwant
    if kw.get("stereo_fraction", 0.5) > 0 or True:
        assert total > 0


def test_ties_and_empty_inputs(oracle):
    """An equal distance later in the node list replaces the earlier candidate (dist > bestDist skips, @0x87a0d); key frames
    without features or without common nodes give no pairs."""
    import plslam_b200 as pl
    kf1, kf2, F12, (R2w, t2w, Cw), (fx, fy, cx, cy), sf, sg = _case(oracle, 26, nbits=1)
    ex, ey = pl.epipole(R2w, t2w, Cw, fx, fy, cx, cy)
    # duplicate every KF2 feature: each duplicate ties with its original
    dup = {k: (np.concatenate([v, v]) if k not in ("nodes", "start", "idx") else v) for k, v in kf2.items()}
    n2 = len(kf2["desc"])
    from matchdata import fake_feature_vector
    dup["nodes"], dup["start"], dup["idx"] = fake_feature_vector(dup["desc"], 1, seed=26 + 100)
    want, nwant = oracle.search_for_triangulation(kf1, dup, F12, ex, ey, sf, sg, False, True)
    got, ngot, _ = pl.search_for_triangulation_host(kf1, dup, F12, ex, ey, sf, sg, False, True)
    assert ngot == nwant > 0 and np.array_equal(got, want)
    assert (want[want >= 0] >= n2).any()  # some winners are the later duplicates
    empty = {k: v[:0] for k, v in kf2.items()}
    empty["start"] = np.zeros(1, np.int32)
    got, ngot, pairs = pl.search_for_triangulation_host(kf1, empty, F12, ex, ey, sf, sg, False, True)
    assert ngot == 0 and pairs == [] and (got == -1).all()
    other = dict(kf2); other["nodes"] = kf2["nodes"] + 1000
    got, ngot, pairs = pl.search_for_triangulation_host(kf1, other, F12, ex, ey, sf, sg, False, True)
    assert ngot == 0 and pairs == []
