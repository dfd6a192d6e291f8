"""One-off fuzz of the contour-ordering kernels against cv2: random sizes (including widths that are not multiples of
32 and single rows / columns), densities and structures, every variant, single slices and same-size batches.

    python tools/fuzz_contours.py [n_cases] [seed]
"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402


def ref(m):
    cs, _ = cv2.findContours((m > 127).astype(np.uint8) * 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    return [c.reshape(-1, 2) for c in cs]


def same(a, b):
    return len(a) == len(b) and all(x.shape == y.shape and (x == y).all() for x, y in zip(a, b))


def make(rng, h, w):
    k = rng.integers(0, 6)
    if k == 0:
        m = rng.random((h, w)) < rng.random()
    elif k == 1:
        f = rng.random((h, w))
        for _ in range(int(rng.integers(1, 4))):
            f = (f + np.roll(f, 1, 0) + np.roll(f, 1, 1) + np.roll(f, -1, 0) + np.roll(f, -1, 1)) / 5
        m = f > np.median(f)
    elif k == 2:
        m = np.ones((h, w), bool)
        m[rng.random((h, w)) < 0.1] = False
    elif k == 3:
        yy, xx = np.mgrid[0:h, 0:w]
        m = ((xx // max(1, int(rng.integers(1, 5)))) + (yy // max(1, int(rng.integers(1, 5))))) % 2 == 0
    elif k == 4:
        yy, xx = np.mgrid[0:h, 0:w]
        r = np.hypot(xx - w / 2, yy - h / 2)
        m = (r.astype(int) // max(1, int(rng.integers(1, 6)))) % 2 == 0
    else:
        m = np.zeros((h, w), bool)
        for _ in range(int(rng.integers(1, 12))):
            y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
            m[y0:y0 + int(rng.integers(1, 40)), x0:x0 + int(rng.integers(1, 40))] = True
    return m.astype(np.uint8) * 255


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    eng = ms.Engine(None)
    bad = 0
    for variant in ("rank", "smem", "window", "crack"):
        os.environ["MEDSEG_TRACE"] = variant
        n_contours = 0
        for i in range(n):
            h, w = int(rng.integers(1, 260)), int(rng.integers(1, 260))
            if i % 50 == 0:
                h, w = int(rng.integers(300, 700)), int(rng.integers(300, 700))
            if i % 7 == 0:      # a small batch of same-size slices
                batch = np.stack([make(rng, h, w) for _ in range(3)])
                polys = eng.mask2polygon(batch)
                got = [polys.slice(j) for j in range(3)]
                want = [ref(b) for b in batch]
            else:
                m = make(rng, h, w)
                got, want = [eng.mask2polygon(m).slice(0)], [ref(m)]
            for g, wv in zip(got, want):
                n_contours += len(wv)
                if not same(g, wv):
                    bad += 1
                    if bad < 5:
                        print("MISMATCH", variant, i, h, w, len(g), len(wv))
        print(variant, "cases", n, "contours", n_contours, "bad so far", bad)
    eng.cleanup()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
