"""Synthetic CT-like inputs and mask families (numpy only).

The reference ships no sample data, so every workload in
BASELINE.json is driven by these seeded generators:

* ``ct_slice``      – 16-bit RAW slice: background ~1000, elliptical body ~2000,
                      three "organ" discs 2600..3000, Gaussian noise sigma=30.
* ``ct_volume``     – a stack of slices, seeds = slice indices.
* ``stress_mask``   – cfg5 mask families fed straight to mask2polygon.

They are deliberately independent of both the oracle and the CUDA path.
"""
from __future__ import annotations

import numpy as np


def ct_slice(seed: int, w: int = 512, h: int = 512) -> np.ndarray:
    """One headerless little-endian u16 slice, row-major ``h x w`` (the layout
    ``MMapFile`` maps at /root/reference/src/preprocess.cpp:86-87)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.full((h, w), 1000.0)
    cx = w / 2 + rng.normal(0, 10)
    cy = h / 2 + rng.normal(0, 10)
    ax, ay = 0.42 * w, 0.33 * h
    body = ((xx - cx) / ax) ** 2 + ((yy - cy) / ay) ** 2 <= 1.0
    img[body] = 2000.0
    for _ in range(3):
        r = rng.uniform(0.05, 0.16) * w
        ox = cx + rng.uniform(-0.22, 0.22) * w
        oy = cy + rng.uniform(-0.15, 0.15) * h
        val = rng.uniform(2600, 3000)
        disc = (xx - ox) ** 2 + (yy - oy) ** 2 <= r * r
        img[disc & body] = val
    img += rng.normal(0, 30, size=(h, w))
    return np.clip(np.rint(img), 0, 65535).astype(np.uint16)


def ct_volume(n: int, w: int = 512, h: int = 512, first_seed: int = 0) -> np.ndarray:
    return np.stack([ct_slice(first_seed + i, w, h) for i in range(n)])


def _box_blur(a: np.ndarray, r: int, passes: int = 3) -> np.ndarray:
    """Cheap separable blur (numpy only) approximating a Gaussian."""
    out = a.astype(np.float64)
    k = 2 * r + 1
    for _ in range(passes):
        for axis in (0, 1):
            c = np.cumsum(np.pad(out, [(r + 1, r) if ax == axis else (0, 0) for ax in (0, 1)],
                                 mode="wrap"), axis=axis)
            if axis == 0:
                out = (c[k:, :] - c[:-k, :]) / k
            else:
                out = (c[:, k:] - c[:, :-k]) / k
    return out


def stress_mask(kind: str, h: int = 2048, w: int = 2048, seed: int = 3) -> np.ndarray:
    """cfg5 mask families (values 0/255, u8)."""
    rng = np.random.default_rng(seed)
    if kind == "blobs":
        f = _box_blur(rng.random((h, w)), 3)
        m = f > np.median(f)
    elif kind == "rings":
        yy, xx = np.mgrid[0:h, 0:w]
        m = np.zeros((h, w), bool)
        step = max(h, w) // 8
        for cy in range(step // 2, h, step):
            for cx in range(step // 2, w, step):
                d = np.maximum(np.abs(yy - cy), np.abs(xx - cx))
                rr = np.hypot(yy - cy, xx - cx)
                sel = d < step // 2 - 2
                m |= sel & ((rr.astype(np.int64) // 6) % 2 == 0)
    elif kind == "checker":
        yy, xx = np.mgrid[0:h, 0:w]
        m = ((yy + xx) % 2) == 0
    elif kind == "diag":
        yy, xx = np.mgrid[0:h, 0:w]
        m = ((yy - xx) % 4 == 0) | ((yy + xx) % 6 == 0)
    elif kind == "noise":
        m = rng.random((h, w)) < 0.5
    elif kind == "sparse":
        m = rng.random((h, w)) < 0.02
    elif kind == "zeros":
        m = np.zeros((h, w), bool)
    elif kind == "ones":
        m = np.ones((h, w), bool)
    else:
        raise ValueError(kind)
    return (m.astype(np.uint8) * 255)


STRESS_KINDS = ("blobs", "rings", "checker", "diag", "noise", "sparse", "zeros", "ones")
