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


def synth_vocabulary_arrays(k=10, L=6, seed=0):
    """A complete k-ary vocabulary tree of depth L in node (breadth-first) order, the arrays plslam_voc_create takes:
    parent [n], is_leaf [n], descriptors [n][32], weights [n].  ORBvoc.txt has k=10, L=6 (1 111 111 nodes when complete).
    Children are noisy copies of their parent's descriptor so that descents discriminate like a trained tree."""
    rng = np.random.default_rng(seed)
    n = (k ** (L + 1) - 1) // (k - 1)
    parent = np.zeros(n, np.int32)
    parent[1:] = (np.arange(1, n, dtype=np.int64) - 1) // k
    first_leaf = (k ** L - 1) // (k - 1)
    is_leaf = np.zeros(n, np.uint8)
    is_leaf[first_leaf:] = 1
    desc = np.zeros((n, 32), np.uint8)
    desc[1:k + 1] = rng.integers(0, 256, (k, 32), dtype=np.uint8)
    lo = 1
    for level in range(1, L):
        cnt = k ** level
        hi = lo + cnt
        ch_lo = hi
        noise = rng.integers(0, 256, (cnt * k, 32), dtype=np.uint8) & rng.integers(0, 256, (cnt * k, 32), dtype=np.uint8) \
            & rng.integers(0, 256, (cnt * k, 32), dtype=np.uint8)
        desc[ch_lo:ch_lo + cnt * k] = np.repeat(desc[lo:hi], k, axis=0) ^ noise
        lo = hi
    weights = np.zeros(n, np.float64)
    weights[first_leaf:] = np.round(rng.random(n - first_leaf) * 10, 5) + 0.01
    return parent, is_leaf, desc, weights
