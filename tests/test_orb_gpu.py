"""GPU parity: the CUDA ORB extractor (through the C-ABI) against the CPU oracle, stage by stage
and end to end.  Bit-exact for pyramid / blur pixels, FAST candidates (position, response, list
order), keypoint order/positions/octave/response and descriptor bytes; angles must match bit for
bit as well (same float sequence), the north-star tolerance of 1e-4 is asserted as the bound."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ext(**kw):
    import plslam_b200 as pl
    return pl.ORBextractor(**kw)


def _compare_frame(o_kps, o_desc, g_kps, g_desc, tag=""):
    assert len(o_kps) == len(g_kps), "%s keypoint count %d vs %d" % (tag, len(o_kps), len(g_kps))
    for fld in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(o_kps[fld], g_kps[fld]), "%s field %s differs" % (tag, fld)
    assert np.max(np.abs(o_kps["angle"] - g_kps["angle"]), initial=0) <= 1e-4, tag + " angle tolerance"
    assert np.array_equal(o_kps["angle"], g_kps["angle"]), tag + " angle bits"
    assert np.array_equal(o_desc, g_desc), tag + " descriptor bytes"


def test_stages_and_end_to_end_640x480(oracle):
    from plslam_b200.synth import synth_frame
    img = synth_frame(0)
    ex = _ext()
    orc = oracle.OrbOracle()
    o_kps, o_desc = orc.extract(img)
    g_kps, g_desc = ex(img)
    t_o, t_g = orc.tables(), ex.tables()
    for k in t_o:
        assert np.array_equal(t_o[k], t_g[k]), k
    for l in range(8):
        assert np.array_equal(orc.level(l), ex.level(0, l)), "pyramid level %d" % l
        assert np.array_equal(orc.candidates(l), ex.candidates(0, l)), "FAST candidates level %d" % l
        ob = orc.level(l, blurred=True)
        if ob is not None:
            assert np.array_equal(ob, ex.level(0, l, blurred=True)), "blurred level %d" % l
    _compare_frame(o_kps, o_desc, g_kps, g_desc)
    assert len(g_kps) >= 1000


# (303, 401): level 7 is 112 x 85, one ROW of FAST cells 59 px high (cells wider than 58 px were rejected in round 1);
# (200, 270): single rows / columns of cells on several levels
@pytest.mark.parametrize("shape", [(480, 640), (720, 1280), (250, 333), (303, 401), (200, 270)])
def test_batch_parity(oracle, shape):
    from plslam_b200.synth import synth_frame
    H, W = shape
    B = 6
    imgs = np.stack([synth_frame(100 + i, W, H) for i in range(B)])
    nf = 2000 if W > 1000 else 1000
    ex = _ext(nfeatures=nf)
    orc = oracle.OrbOracle(nfeatures=nf)
    kps, desc, counts = ex.extract_batch_host(imgs)
    for f in range(B):
        o_kps, o_desc = orc.extract(imgs[f])
        n = counts[f]
        _compare_frame(o_kps, o_desc, kps[f, :n], desc[f, :n], "frame %d" % f)


def test_noise_and_flat_frames(oracle):
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 256, (480, 640)).astype(np.uint8)
    flat = np.full((480, 640), 77, np.uint8)
    ex = _ext()
    orc = oracle.OrbOracle()
    for name, img in (("noise", noise), ("flat", flat)):
        o_kps, o_desc = orc.extract(img)
        g_kps, g_desc = ex(img)
        _compare_frame(o_kps, o_desc, g_kps, g_desc, name)
    assert len(ex(flat)[0]) == 0


def test_device_resident_batch(oracle):
    import torch
    import plslam_b200 as pl
    from plslam_b200.synth import synth_frame
    imgs = np.stack([synth_frame(300 + i) for i in range(4)])
    ex = _ext()
    d = torch.from_numpy(imgs).cuda()
    kps, desc, counts = ex.extract_batch_device(d)
    ex.check_status()
    torch.cuda.synchronize()
    k = pl.kps_from_tensor(kps)
    orc = oracle.OrbOracle()
    for f in range(4):
        o_kps, o_desc = orc.extract(imgs[f])
        n = int(counts[f])
        _compare_frame(o_kps, o_desc, k[f, :n], desc[f, :n].cpu().numpy(), "frame %d" % f)


def test_empty_image_is_silent():
    ex = _ext()
    k, d = ex(np.empty((0, 0), np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)


def test_device_batch_unaligned_view(oracle):
    """Device-resident frames that are a misaligned VIEW of a larger buffer (odd base address, pitch and frame stride that are
    no multiples of 16): level 0 is then not TMA-legal, so level 1 comes from the plain-load k_resize while the other levels
    use k_resize_tma, k_blur reads level 0 with plain loads, and the word-staging paths of k_fast, k_lsd_scale and k_lbd see a
    base that is not word aligned."""
    import torch
    import plslam_b200 as pl
    from plslam_b200.synth import synth_frame
    H, W, B = 250, 333, 3
    imgs = np.stack([synth_frame(400 + i, W, H) for i in range(B)])
    big = torch.zeros((B, H + 3, W + 9), dtype=torch.uint8, device="cuda")
    view = big[:, 2:2 + H, 5:5 + W]
    view.copy_(torch.from_numpy(imgs).cuda())
    assert view.data_ptr() % 4 != 0 and view.stride(1) % 16 != 0
    ex = _ext()
    kps, desc, counts = ex.extract_batch_device(view)
    ex.check_status()
    ls = pl.LineSegment()
    kl, ldesc, funcs, lcounts = ls.extract_batch_device(view)
    ls.check_status()
    torch.cuda.synchronize()
    k = pl.kps_from_tensor(kps)
    klines = pl.keylines_from_tensor(kl)
    orc = oracle.OrbOracle()
    for f in range(B):
        o_kps, o_desc = orc.extract(imgs[f])
        n = int(counts[f])
        _compare_frame(o_kps, o_desc, k[f, :n], desc[f, :n].cpu().numpy(), "frame %d" % f)
        okl, odesc, ofun, _ = oracle.extract_lines(imgs[f], 40)
        m = int(lcounts[f])
        assert m == len(okl), "frame %d line count" % f
        for fld in okl.dtype.names:
            assert np.array_equal(klines[f, :m][fld], okl[fld]), "frame %d keyline %s" % (f, fld)
        assert np.array_equal(ldesc[f, :m].cpu().numpy(), odesc) and np.array_equal(funcs[f, :m].cpu().numpy(), ofun), "frame %d LBD" % f


@pytest.mark.parametrize("cfg", [(300, 1.5, 4, 30, 10), (1500, 1.1, 8, 15, 5), (500, 1.2, 10, 20, 7), (50, 2.0, 3, 20, 7)])
def test_other_extractor_settings(oracle, cfg):
    """ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) away from the ORB-SLAM2 defaults: level ratios the
    TMA pyramid box does not hold (1.5, 2.0: plain-load k_resize), 10 levels, small quotas in the quad-tree."""
    from plslam_b200.synth import synth_frame
    nf, sf, nl, ini, mn = cfg
    ex = _ext(nfeatures=nf, scaleFactor=sf, nlevels=nl, iniThFAST=ini, minThFAST=mn)
    orc = oracle.OrbOracle(nf, sf, nl, ini, mn)
    for seed in (7, 8):
        img = synth_frame(seed)
        o_kps, o_desc = orc.extract(img)
        g_kps, g_desc = ex(img)
        assert len(o_kps) > 0
        _compare_frame(o_kps, o_desc, g_kps, g_desc, "cfg %s seed %d" % (cfg, seed))
