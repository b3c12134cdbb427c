"""Seeded synthetic inputs of the shapes BASELINE.json names.

Generated on the host with numpy's PCG64 so the CPU oracle and the CUDA path see identical bytes.
"""
from __future__ import annotations

import numpy as np


def corner_image(height: int, width: int, seed: int = 0) -> np.ndarray:
    """uint8 gray image: level-128 background, H*W/4096 axis-aligned rectangles, N(0,1) noise.

    About 2 % of the pixels end up above the reference's default Harris threshold (10000) and
    above 1 % of the lambda-minus maximum, which is what the corner benchmarks assume."""
    rng = np.random.default_rng(seed)
    img = np.full((height, width), 128.0)
    n_rect = max(1, (height * width) // 4096)
    ys = rng.integers(0, height, n_rect)
    xs = rng.integers(0, width, n_rect)
    hs = rng.integers(8, 64, n_rect)
    ws = rng.integers(8, 64, n_rect)
    lv = rng.integers(16, 240, n_rect)
    for y, x, h, w, v in zip(ys, xs, hs, ws, lv):
        img[y:y + h, x:x + w] = v
    img += rng.standard_normal((height, width))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def sift_image(height: int, width: int, seed: int = 0, max_blobs: int = 4096) -> np.ndarray:
    """float64 gray image in [0, 1]: 0.5 + Gaussian blobs (sigma U(1.5, 8), amplitude U(-.5, .5))
    + N(0, 0.02) noise, clipped."""
    rng = np.random.default_rng(seed)
    img = np.full((height, width), 0.5)
    n_blob = int(min(max_blobs, max(4, (height * width) // 1024)))
    cy = rng.uniform(0, height, n_blob)
    cx = rng.uniform(0, width, n_blob)
    sg = rng.uniform(1.5, 8.0, n_blob)
    am = rng.uniform(-0.5, 0.5, n_blob)
    for y, x, s, a in zip(cy, cx, sg, am):
        r = int(np.ceil(4 * s))
        y0, y1 = max(0, int(y) - r), min(height, int(y) + r + 1)
        x0, x1 = max(0, int(x) - r), min(width, int(x) + r + 1)
        if y1 <= y0 or x1 <= x0:
            continue
        yy = np.arange(y0, y1)[:, None] - y
        xx = np.arange(x0, x1)[None, :] - x
        img[y0:y1, x0:x1] += a * np.exp(-(yy * yy + xx * xx) / (2 * s * s))
    img += rng.normal(0.0, 0.02, (height, width))
    return np.clip(img, 0.0, 1.0)


def template_from(main: np.ndarray, size: int, seed: int = 0, rot90: int = 0) -> np.ndarray:
    """A size x size crop of `main` (optionally rotated by multiples of 90 degrees) so that true
    matches exist between the pair."""
    rng = np.random.default_rng(seed + 7919)
    y = int(rng.integers(0, main.shape[0] - size + 1))
    x = int(rng.integers(0, main.shape[1] - size + 1))
    return np.ascontiguousarray(np.rot90(main[y:y + size, x:x + size], rot90))
