"""Seeded synthetic inputs for the benchmark and parity tests. numpy only, no cv2."""
import numpy as np


def stereo_frame(seed, w=1242, h=375, disparity=8):
    """One synthetic stereo pair: 1500 random grey rectangles on 128, 3x3 Gaussian (sigma 0.8), N(0, 2^2) noise.

    Right image = left image shifted by `disparity` px (np.roll), as SURVEY.md §8d specifies.
    """
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 128.0, np.float64)
    n = 1500
    xs = rng.integers(0, w, n)
    ys = rng.integers(0, h, n)
    ws = rng.integers(6, 90, n)
    hs = rng.integers(6, 60, n)
    gs = rng.integers(0, 256, n)
    for x, y, rw, rh, g in zip(xs, ys, ws, hs, gs):
        img[y:y + rh, x:x + rw] = g
    k = np.exp(-0.5 * (np.array([-1.0, 0.0, 1.0]) / 0.8) ** 2)
    k /= k.sum()
    p = np.pad(img, 1, mode="reflect")
    img = k[0] * p[1:-1, :-2] + k[1] * p[1:-1, 1:-1] + k[2] * p[1:-1, 2:]
    p = np.pad(img, 1, mode="reflect")
    img = k[0] * p[:-2, 1:-1] + k[1] * p[1:-1, 1:-1] + k[2] * p[2:, 1:-1]
    img = img + rng.normal(0.0, 2.0, img.shape)
    left = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    right = np.roll(left, -disparity, axis=1)
    return np.ascontiguousarray(left), np.ascontiguousarray(right)


def frame_seed(idx):
    return 1234 + idx
