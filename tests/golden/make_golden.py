#!/usr/bin/env python3
"""Generates the committed golden vectors under tests/golden/ (run in the BUILD container, where cv2 4.13, the
reference tree /root/reference and oracle/_ref exist; the GPU box only reads the .npz/.json files).

Sources of truth, none of them this repository's own oracle:
  cv2_primitives.npz   cv2 4.13: resize INTER_LINEAR, GaussianBlur 7x7 sigma 2, FAST-9/16 + NMS, fastAtan2,
                       undistortPoints (the OpenCV calls of ORBextractor / Frame: SURVEY.md Appendix B, section 8f-2)
  cv2_lsd.npz          cv2 4.13 createLineSegmentDetector(LSD_REFINE_ADV) on two small synthetic frames
  dbow2_ref.npz        the reference's OWN Thirdparty/DBoW2 sources (compiled into oracle/_ref/libdbow2_ref.so):
                       FORB::distance and ORBVocabulary::transform on a small vocabulary
  reference_binary.json  constants read from /root/reference/lib/libORB_SLAM2.so (rBRIEF pattern, matcher thresholds)
  tum_io.json          the reference's own association lists (Examples/RGB-D/associations/*.txt) parsed with str.split /
                       float(): entry count, digest, first and last entry of each; the trajectory line of the identity pose
                       from cv2.gemm + Python's "%.9f"  (`python tests/golden/make_golden.py tum` writes only this file)

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "rgbd-pl-slam_b200")):
    sys.path.insert(0, p)


def tum_io():
    import cv2
    d = "/root/reference/Examples/RGB-D/associations"
    out = {"associations": {}}
    for name in sorted(os.listdir(d)):
        ts, rgb, dep = [], [], []
        for line in open(os.path.join(d, name), "rb").read().decode("latin-1").split("\n"):
            if line == "":
                continue
            tok = line.split()
            assert len(tok) == 4, (name, line)  # well-formed lists only: the split rule then equals the reference's extraction
            ts.append(float(tok[0])); rgb.append(tok[1]); dep.append(tok[3])
        h = hashlib.sha256()
        for t, a, b in zip(ts, rgb, dep):
            h.update(("%s|%s|%s\n" % (float(t).hex(), a, b)).encode())
        out["associations"][name] = {"n": len(ts), "sha256": h.hexdigest(), "first": [ts[0].hex(), rgb[0], dep[0]],
                                     "last": [ts[-1].hex(), rgb[-1], dep[-1]]}
    twc = cv2.gemm(np.eye(3, dtype=np.float32), np.zeros((3, 1), np.float32), -1.0, None, 0.0).ravel()
    out["identity_line"] = "%.6f %.9f %.9f %.9f %.9f %.9f %.9f %.9f\n" % ((0.0,) + tuple(float(x) for x in twc) + (0.0, 0.0, 0.0, 1.0))
    json.dump(out, open(os.path.join(HERE, "tum_io.json"), "w"), indent=1)
    print("tum_io.json:", {k: v["n"] for k, v in out["associations"].items()}, repr(out["identity_line"]))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "tum":
        return tum_io()
    import cv2
    from plslam_b200.synth import synth_frame
    assert cv2.__version__.startswith("4.13"), cv2.__version__
    rng = np.random.default_rng(2026)

    # ---- OpenCV primitives ----
    img = np.ascontiguousarray(synth_frame(0)[100:228, 200:360])  # 128 x 160
    noise = rng.integers(0, 256, (61, 83)).astype(np.uint8)
    out = {"img": img, "noise": noise}
    out["resize_img_133x107"] = cv2.resize(img, (133, 107), interpolation=cv2.INTER_LINEAR)
    out["resize_noise_69x51"] = cv2.resize(noise, (69, 51), interpolation=cv2.INTER_LINEAR)
    out["blur_img"] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    out["blur_noise"] = cv2.GaussianBlur(noise, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    for th in (20, 7):
        det = cv2.FastFeatureDetector_create(th, True)
        for name, sub in (("img", img), ("cell", np.ascontiguousarray(img[16:54, 100:137]))):
            out["fast%d_%s" % (th, name)] = np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in det.detect(sub)],
                                                     np.int32).reshape(-1, 3)
    yx = rng.integers(-60000, 60000, (3000, 2)).astype(np.float32)
    yx[:4] = [[0, 0], [0, 5], [5, 0], [-3, -3]]
    out["atan2_yx"] = yx
    out["atan2_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    cal = dict(fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, k1=0.262383, k2=-0.953104, p1=-0.005358,
               p2=0.002628, k3=1.163314)  # /root/reference/Examples/RGB-D/TUM1.yaml
    K = np.array([[cal["fx"], 0, cal["cx"]], [0, cal["fy"], cal["cy"]], [0, 0, 1]], np.float32)
    D = np.array([cal[k] for k in ("k1", "k2", "p1", "p2", "k3")], np.float32).reshape(5, 1)
    pts = np.stack([rng.uniform(0, 640, 2000), rng.uniform(0, 480, 2000)], 1).astype(np.float32)
    pts[:4] = [[0, 0], [640, 0], [0, 480], [640, 480]]
    out["undist_in"] = pts
    out["undist_out"] = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
    out["undist_calib"] = np.array([cal[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3")], np.float32)
    # the LSD front half: sigma 0.75 blur + x0.8 INTER_LINEAR_EXACT
    out["lsd_blur_noise"] = cv2.GaussianBlur(noise, (7, 7), 0.75)
    out["lsd_scaled_img"] = cv2.resize(cv2.GaussianBlur(img, (7, 7), 0.75), None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR_EXACT)
    np.savez_compressed(os.path.join(HERE, "cv2_primitives.npz"), **out)

    # ---- LSD ----
    det = cv2.createLineSegmentDetector(cv2.LSD_REFINE_ADV)
    lsd = {}
    for i, (seed, W, H) in enumerate(((3, 320, 240), (4, 250, 333))):
        f = synth_frame(seed, W, H)
        lines, width, prec, nfa = det.detect(f)
        lsd["img%d" % i] = f
        lsd["lines%d" % i] = lines.reshape(-1, 4)
        lsd["width%d" % i] = width.ravel()
        lsd["prec%d" % i] = prec.ravel()
        lsd["nfa%d" % i] = nfa.ravel()
    np.savez_compressed(os.path.join(HERE, "cv2_lsd.npz"), **lsd)

    # ---- the reference's DBoW2 ----
    from oracle import bindings as ob
    ob.build()
    assert ob.dbow2_ref() is not None, "oracle/_ref/libdbow2_ref.so missing: run `make -C oracle` where /root/reference exists"
    voc_path = os.path.join(HERE, "voc_k6_L3.txt")
    ob.write_vocabulary_text(voc_path, 6, 3, seed=21)
    desc = rng.integers(0, 256, (400, 32)).astype(np.uint8)
    _, od = ob.OrbOracle().extract(synth_frame(5))
    desc[:200] = od[:200]
    ref = ob.VocReference(voc_path)
    bow = {"desc": desc}
    for lu in (0, 1, 2):
        t = ref.transform(desc, lu)
        for k, v in t.items():
            bow["lu%d_%s" % (lu, k)] = v
    a, b = desc[:200], desc[200:]
    bow["forb_distance"] = np.array([ob.VocReference.forb_distance(x, y) for x, y in zip(a, b)], np.int32)
    np.savez_compressed(os.path.join(HERE, "dbow2_ref.npz"), **bow)

    # ---- constants of the reference binary ----
    so = open("/root/reference/lib/libORB_SLAM2.so", "rb").read()
    pat = struct.unpack("<1024i", so[0x141c40:0x142c40])
    histo, th_low, th_high = struct.unpack("<3i", so[0x1269e0:0x1269ec])
    json.dump({"bit_pattern_31_sha256": hashlib.sha256(struct.pack("<1024i", *pat)).hexdigest(), "bit_pattern_31_first16": pat[:16],
               "HISTO_LENGTH": histo, "TH_LOW": th_low, "TH_HIGH": th_high,
               "source": "lib/libORB_SLAM2.so .data offset 0x141c40 (VA 0x341c40), .rodata 0x1269e0"},
              open(os.path.join(HERE, "reference_binary.json"), "w"), indent=1)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
    if len(sys.argv) == 1:
        tum_io()
