"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/plslam_b200.h
declares, argument validation and the no-CPU-fallback behaviour work without a GPU, struct layouts match the
reference's cv::KeyPoint / KeyLine, and the host-side scalar pieces agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "plslam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(plslam_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import plslam_b200 as pl
    lib = pl.lib()
    names = _declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layouts():
    import plslam_b200 as pl
    assert pl.KP_DTYPE.itemsize == 28 and pl.KEYLINE_DTYPE.itemsize == 68
    assert C.sizeof(pl.KnnJob) == 32 and C.sizeof(pl.BowJob) == 144 and C.sizeof(pl.ProjJob) == 328 and C.sizeof(pl.TriJob) == 240
    assert C.sizeof(pl.FuseJob) == 304 and C.sizeof(pl.KfProjJob) == 240 and C.sizeof(pl.FrustumJob) == 168 and C.sizeof(pl.FrontendIO) == 72 and C.sizeof(pl.LocalJob) == 160 and C.sizeof(pl.FrameCalib) == 40
    assert [n for n in pl.KP_DTYPE.names] == ["x", "y", "size", "angle", "response", "octave", "class_id"]


def test_constructor_tables_without_gpu(oracle):
    import plslam_b200 as pl
    for nf in (1000, 2000, 8000):
        ex = pl.ORBextractor(nfeatures=nf)
        a, b = ex.tables(), oracle.OrbOracle(nfeatures=nf).tables()
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        assert ex.max_keypoints == int(a["quota"].sum()) + 3 * 8
    assert pl.ORBextractor().GetLevels() == 8


def test_argument_validation_and_empty_image():
    import plslam_b200 as pl
    with pytest.raises(pl.PlslamError):
        pl.ORBextractor(nfeatures=0)
    with pytest.raises(pl.PlslamError):
        pl.ORBextractor(nlevels=99)
    with pytest.raises(pl.PlslamError):
        pl.ORBextractor().set_blur_kernel([1, 2, 3, 4, 5, 6, 7])  # must sum to 256
    k, d = pl.ORBextractor()(np.empty((0, 0), np.uint8))  # empty image: silent return (reference @0x76dda)
    assert len(k) == 0 and d.shape == (0, 32)
    kl, ld, fn = pl.LineSegment().ExtractLineSegment(np.empty((0, 0), np.uint8))
    assert len(kl) == 0


def test_no_cpu_fallback_without_a_device():
    import torch
    import plslam_b200 as pl
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert pl.lib().plslam_device_count() == 0
    img = np.zeros((480, 640), np.uint8)
    with pytest.raises(pl.PlslamError) as e:
        pl.ORBextractor()(img)
    assert e.value.code == pl.ERR_CUDA
    with pytest.raises(pl.PlslamError):
        pl.LineSegment().ExtractLineSegment(img)
    with pytest.raises(pl.PlslamError):
        pl.knn2_host(np.zeros((2, 32), np.uint8), np.zeros((2, 32), np.uint8))


def test_scalar_descriptor_distance_matches_oracle(oracle):
    import plslam_b200 as pl
    rng = np.random.default_rng(0)
    for _ in range(200):
        a, b = rng.integers(0, 256, 32).astype(np.uint8), rng.integers(0, 256, 32).astype(np.uint8)
        assert pl.DescriptorDistance(a, b) == oracle.descriptor_distance(a, b) == int(np.unpackbits(a ^ b).sum())
    z = np.zeros(32, np.uint8)
    assert pl.DescriptorDistance(z, z) == 0 and pl.DescriptorDistance(z, ~z) == 256


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "rgbd-pl-slam_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, os.path.join(dp, f)
