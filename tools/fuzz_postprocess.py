"""One-off fuzz of K5 (hole fill + 3x3 open + area filter) against the cv2 oracle: random sizes, class maps with blobs,
holes, speckle and border-touching regions, every foreground value.

    python tools/fuzz_postprocess.py [n_cases] [seed]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from oracle import pipeline as op  # noqa: E402  (checker only)


def make(rng, h, w):
    m = np.zeros((h, w), np.uint8)
    for _ in range(int(rng.integers(1, 8))):
        cy, cx = rng.integers(0, h), rng.integers(0, w)
        ry, rx = rng.integers(2, max(3, h // 2)), rng.integers(2, max(3, w // 2))
        yy, xx = np.ogrid[0:h, 0:w]
        m[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1] = rng.integers(1, 4)
    sp = rng.random((h, w))
    m[sp < 0.02] = 0
    m[sp > 0.985] = rng.integers(1, 4)
    return m


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    eng = ms.Engine(None)
    bad = 0
    for i in range(n):
        h, w = int(rng.integers(1, 200)), int(rng.integers(1, 200))
        if i % 25 == 0:
            h, w = int(rng.integers(200, 600)), int(rng.integers(200, 600))
        fg = int(rng.integers(1, 4))
        batch = np.stack([make(rng, h, w) for _ in range(int(rng.integers(1, 4)))])
        got = eng.postprocess(batch, fg_value=fg)
        for j in range(len(batch)):
            want = op.postprocess_mask(batch[j], fg=fg)
            if not (got[j] == want).all():
                bad += 1
                if bad < 5:
                    print("MISMATCH", i, j, h, w, fg, int((got[j] != want).sum()))
    print("cases", n, "bad", bad)
    eng.cleanup()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
