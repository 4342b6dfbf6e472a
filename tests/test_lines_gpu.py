"""GPU parity: the CUDA LSD + LBD line extractor (through the C-ABI) against the CPU oracle in its
PINNED mode.  Stage by stage (scaled image, level-line field, accepted segments) and end to end
(KeyLines, LBD descriptor bytes, line equations).  Integer/byte outputs and segment end points are
required to be bit-identical; the north-star tolerance for end points (1e-4 px) is asserted too."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check_frame(oracle, ls, img, f, tag):
    sc, ang, mg = oracle.lsd_stage(img)
    assert np.array_equal(sc, ls.scaled(f)), tag + " scaled image"
    deg, g2 = ls.level_lines(f)
    o_deg = np.where(ang == -1024.0, -1024.0, ang * (180.0 / np.pi))
    assert np.array_equal(deg == -1024.0, ang == -1024.0), tag + " defined mask"
    # the record stores degrees; radians = deg * (pi/180) in double must reproduce the oracle's angle bit for bit
    rad = deg.astype(np.float64) * (np.pi / 180)
    assert np.array_equal(rad[ang != -1024.0], ang[ang != -1024.0]), tag + " level-line angles"
    assert np.array_equal(np.sqrt(g2[:-1, :-1] / 4.0), mg[:-1, :-1]), tag + " gradient norm"
    o_seg, _ = oracle.lsd_detect(img, compat=0)
    g_seg = ls.segments(f)
    assert len(o_seg) == len(g_seg), "%s segment count %d vs %d" % (tag, len(o_seg), len(g_seg))
    assert np.max(np.abs(o_seg[:, :4] - g_seg[:, :4]), initial=0) <= 1e-4, tag + " end point tolerance"
    assert np.array_equal(o_seg[:, :6], g_seg[:, :6]), tag + " segments (x1,y1,x2,y2,width,prec)"
    assert np.allclose(o_seg[:, 6], g_seg[:, 6], rtol=1e-9, atol=1e-9), tag + " log-NFA"


def _check_lines(o, g, tag):
    okl, odesc, ofun, _ = o
    gkl, gdesc, gfun = g
    assert len(okl) == len(gkl), tag + " line count"
    for fld in okl.dtype.names:
        assert np.array_equal(okl[fld], gkl[fld]), "%s keyline field %s" % (tag, fld)
    assert np.array_equal(odesc, gdesc), tag + " LBD bytes"
    assert np.array_equal(ofun, gfun), tag + " line functions"


def test_lsd_stages_and_lines_640x480(oracle):
    import plslam_b200 as pl
    from plslam_b200.synth import synth_frame
    ls = pl.LineSegment()
    for seed in (0, 1):
        img = synth_frame(seed)
        g = ls.ExtractLineSegment(img)
        _check_frame(oracle, ls, img, 0, "seed %d" % seed)
        _check_lines(oracle.extract_lines(img, 40), g, "seed %d" % seed)
        assert len(g[0]) == 40


@pytest.mark.parametrize("shape", [(480, 640), (720, 1280), (250, 333)])
def test_lines_batch_parity(oracle, shape):
    import plslam_b200 as pl
    from plslam_b200.synth import synth_frame
    H, W = shape
    B = 5
    imgs = np.stack([synth_frame(200 + i, W, H) for i in range(B)])
    ls = pl.LineSegment()
    kl, desc, funcs, counts = ls.extract_batch_host(imgs)
    for f in range(B):
        n = counts[f]
        _check_frame(oracle, ls, imgs[f], f, "frame %d" % f)
        _check_lines(oracle.extract_lines(imgs[f], 40), (kl[f, :n], desc[f, :n], funcs[f, :n]), "frame %d" % f)


def test_all_lines_mode_and_degenerate_frames(oracle):
    import plslam_b200 as pl
    from plslam_b200.synth import synth_frame
    ls = pl.LineSegment(max_lines=0)
    img = synth_frame(7)
    g = ls.ExtractLineSegment(img)
    _check_lines(oracle.extract_lines(img, 0), g, "all lines")
    assert len(g[0]) > 100
    flat = np.full((480, 640), 90, np.uint8)
    assert len(ls.ExtractLineSegment(flat)[0]) == 0
    noise = np.random.default_rng(3).integers(0, 256, (240, 320)).astype(np.uint8)
    _check_lines(oracle.extract_lines(noise, 0), ls.ExtractLineSegment(noise), "noise")


def test_lines_device_resident(oracle):
    import torch
    import plslam_b200 as pl
    from plslam_b200.synth import synth_frame
    imgs = np.stack([synth_frame(400 + i) for i in range(3)])
    ls = pl.LineSegment()
    kl, desc, funcs, counts = ls.extract_batch_device(torch.from_numpy(imgs).cuda())
    ls.check_status()
    k = pl.keylines_from_tensor(kl)
    for f in range(3):
        n = int(counts[f])
        _check_lines(oracle.extract_lines(imgs[f], 40), (k[f, :n], desc[f, :n].cpu().numpy(), funcs[f, :n].cpu().numpy()),
                     "frame %d" % f)
