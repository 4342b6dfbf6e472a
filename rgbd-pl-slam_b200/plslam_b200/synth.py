"""Synthetic RGB-D frames for the benchmark and the parity tests.

synth_frame(seed, W, H): grey canvas, filled/outlined rectangles, line segments, a light
Gaussian blur and N(0,3) noise.  Deterministic for a given (seed, W, H) and cv2 build.
"""
import numpy as np


def synth_frame(seed, W=640, H=480):
    import cv2
    rng = np.random.default_rng(seed)
    s = (W * H) / (640.0 * 480.0)
    img = np.full((H, W), 128, np.uint8)
    for _ in range(int(round(40 * s))):
        x0 = int(rng.integers(0, W - 40)); y0 = int(rng.integers(0, H - 40))
        w = int(rng.integers(20, 200)); h = int(rng.integers(20, 160))
        g = int(rng.integers(0, 256))
        filled = rng.random() < 0.6
        cv2.rectangle(img, (x0, y0), (min(x0 + w, W - 1), min(y0 + h, H - 1)), g, -1 if filled else 2)
    for _ in range(int(round(30 * s))):
        p0 = (int(rng.integers(0, W)), int(rng.integers(0, H)))
        p1 = (int(rng.integers(0, W)), int(rng.integers(0, H)))
        g = int(rng.integers(0, 256)); t = int(rng.integers(1, 4))
        cv2.line(img, p0, p1, g, t)
    f = cv2.GaussianBlur(img.astype(np.float32), (0, 0), 0.8)
    f += rng.normal(0.0, 3.0, size=f.shape).astype(np.float32)
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)


def synth_depth(seed, W=640, H=480):
    rng = np.random.default_rng(seed + 7_000_000)
    z = (1.5 + 0.5 * np.arange(W, dtype=np.float64)[None, :] / W) * 5000.0
    z = np.repeat(z, H, 0) + rng.normal(0.0, 10.0, size=(H, W))
    return np.clip(np.rint(z), 0, 65535).astype(np.uint16)


def synth_pair(seed, W=640, H=480):
    """Second frame = first warped by a 3-px translation + 1 degree rotation, fresh noise."""
    import cv2
    a = synth_frame(seed, W, H)
    M = cv2.getRotationMatrix2D((W / 2.0, H / 2.0), 1.0, 1.0)
    M[0, 2] += 3.0
    M[1, 2] += 3.0
    b = cv2.warpAffine(a, M, (W, H), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    rng = np.random.default_rng(seed + 1_000_000)
    b = np.clip(np.rint(b.astype(np.float32) + rng.normal(0.0, 1.0, size=b.shape)), 0, 255).astype(np.uint8)
    return a, b


def synth_batch(seed0, B, W=640, H=480):
    return np.stack([synth_frame(seed0 + i, W, H) for i in range(B)])
