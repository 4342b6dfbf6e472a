"""GPU parity of the Frame post-extraction kernel (plslam_frame_post_*) against the oracle, on the extractor's own
device-resident output, and chained into SearchByProjection (the grid it builds is the one the matcher consumes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _depth(seed, holes=0.15):
    from plslam_b200.synth import synth_depth
    d = synth_depth(seed).astype(np.float32) * np.float32(1.0 / 5000.0)  # convertTo(CV_32F, 1/DepthMapFactor), TUM1.yaml:35
    rng = np.random.default_rng(seed)
    d[rng.random(d.shape) < holes] = 0.0
    return d


@pytest.mark.parametrize("distorted", [True, False])
def test_frame_post_batch_on_extractor_output(oracle, distorted):
    import torch
    import plslam_b200 as pl
    from plslam_b200.synth import synth_frame
    cal = pl.TUM1_CALIB if distorted else dict(pl.TUM1_CALIB, k1=0.0)
    B = 6
    imgs = np.stack([synth_frame(300 + i) for i in range(B)])
    imgs[5] = 128  # a frame without keypoints
    depth = np.stack([_depth(i) for i in range(B)])
    ex = pl.ORBextractor()
    d_kps, d_desc, d_cnt = ex.extract_batch_device(torch.from_numpy(imgs).cuda())
    bounds = pl.frame_image_bounds(cal, 640, 480)
    assert np.array_equal(bounds, oracle.image_bounds(cal, 640, 480))
    out = pl.frame_post_device(cal, bounds, d_kps, d_cnt, torch.from_numpy(depth).cuda())
    torch.cuda.synchronize()
    kps = pl.kps_from_tensor(d_kps)
    cnt = d_cnt.cpu().numpy()
    assert cnt[5] == 0 and cnt[:5].min() > 900
    for f in range(B):
        n = int(cnt[f])
        xy = np.stack([kps[f, :n]["x"], kps[f, :n]["y"]], 1)
        o = oracle.frame_post(cal, bounds, xy, depth[f])
        assert np.array_equal(out["un_xy"][f, :n].cpu().numpy(), o["un_xy"]), "frame %d undistorted points" % f
        assert np.array_equal(out["uright"][f, :n].cpu().numpy(), o["uright"])
        assert np.array_equal(out["depth"][f, :n].cpu().numpy(), o["depth"])
        gs = out["grid_start"][f].cpu().numpy()
        assert np.array_equal(gs, o["grid_start"])
        assert np.array_equal(out["grid_items"][f, :gs[-1]].cpu().numpy(), o["grid_items"])
    # shared depth map (frame stride 0) and the single-frame host form
    n = int(cnt[0])
    h = pl.frame_post_host(cal, bounds, kps[0, :n], depth[0])
    o = oracle.frame_post(cal, bounds, np.stack([kps[0, :n]["x"], kps[0, :n]["y"]], 1), depth[0])
    for k in o:
        assert np.array_equal(h[k], o[k]), k
    h0 = pl.frame_post_host(cal, bounds, kps[0, :0], depth[0])
    assert h0["grid_start"][-1] == 0 and len(h0["un_xy"]) == 0


def test_frame_post_random_keypoints_full_image(oracle):
    """Keypoints anywhere in the image (cells 64 / 48 reached by rounding are dropped), more than one pass of the CTA."""
    import torch
    import plslam_b200 as pl
    rng = np.random.default_rng(9)
    cal = pl.TUM1_CALIB
    bounds = pl.frame_image_bounds(cal, 640, 480)
    B, cap = 3, 8200
    kps = np.zeros((B, cap), pl.KP_DTYPE)
    cnt = np.array([8200, 1, 3000], np.int32)
    for f in range(B):
        kps[f]["x"] = rng.uniform(0, 639.9, cap).astype(np.float32)
        kps[f]["y"] = rng.uniform(0, 479.9, cap).astype(np.float32)
    depth = _depth(2)
    d_kps = torch.from_numpy(kps.view(np.int32).reshape(B, cap, 7)).cuda()
    out = pl.frame_post_device(cal, bounds, d_kps, torch.from_numpy(cnt).cuda(), torch.from_numpy(depth[None]).cuda())
    torch.cuda.synchronize()
    dropped = 0
    for f in range(B):
        n = int(cnt[f])
        o = oracle.frame_post(cal, bounds, np.stack([kps[f, :n]["x"], kps[f, :n]["y"]], 1), depth)
        assert np.array_equal(out["un_xy"][f, :n].cpu().numpy(), o["un_xy"])
        assert np.array_equal(out["uright"][f, :n].cpu().numpy(), o["uright"])
        assert np.array_equal(out["depth"][f, :n].cpu().numpy(), o["depth"])
        gs = out["grid_start"][f].cpu().numpy()
        assert np.array_equal(gs, o["grid_start"])
        assert np.array_equal(out["grid_items"][f, :gs[-1]].cpu().numpy(), o["grid_items"])
        dropped += n - gs[-1]
    assert dropped > 0


def test_grid_feeds_search_by_projection(oracle):
    """The device-built grid + undistorted points + uRight drive SearchByProjection to the same matches as the oracle
    fed with the oracle's own Frame post-processing."""
    import torch
    import plslam_b200 as pl
    from plslam_b200.synth import synth_pair
    import matchdata as md
    cal = dict(pl.TUM1_CALIB, k1=0.0)  # projection_case() models an undistorted pinhole camera
    a, b = synth_pair(11)
    orc = oracle.OrbOracle()
    ka, da = orc.extract(a)
    kb, dbb = orc.extract(b)
    last, cur, cam, sf, tcw_cur, tcw_last = md.projection_case(ka, da, kb, dbb, orc.tables()["scale"], seed=5)
    depth = _depth(4, holes=0.3)
    bounds = pl.frame_image_bounds(cal, 640, 480)
    post = pl.frame_post_host(cal, bounds, kb, depth)
    opost = oracle.frame_post(cal, bounds, np.stack([kb["x"], kb["y"]], 1), depth)
    cur_g = dict(cur, xy=post["un_xy"], uright=post["uright"], grid_start=post["grid_start"],
                 grid_items=post["grid_items"] if len(post["grid_items"]) else np.zeros(1, np.int32))
    cur_o = dict(cur, xy=opost["un_xy"], uright=opost["uright"], grid_start=opost["grid_start"],
                 grid_items=opost["grid_items"] if len(opost["grid_items"]) else np.zeros(1, np.int32))
    cam = cam.copy()
    cam[4] = cal["bf"]; cam[5] = np.float32(cal["bf"]) / cam[0]
    m_o, n_o = oracle.search_by_projection(last, cur_o, cam, sf, tcw_cur, tcw_last, 15.0)
    m_g, n_g = pl.search_by_projection_host(last, cur_g, cam, sf, tcw_cur, tcw_last, 15.0)
    assert n_o == n_g and n_o > 50
    assert np.array_equal(m_o, m_g)
