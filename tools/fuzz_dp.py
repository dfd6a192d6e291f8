"""One-off fuzz of the opt-in Douglas-Peucker kernel (dp_simplify.cuh) against cv2.approxPolyDP on the contours cv2.findContours
returns: random sizes (incl. widths that are not multiples of 32), structures and epsilons (tie-prone multiples of 0.25 and
random ones), single slices and same-size batches, non-identity coordinate mappings.

    python tools/fuzz_dp.py [n_cases] [seed]
"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from tools.fuzz_contours import make, ref, same  # noqa: E402


def want(m, eps, sx, sy):
    out = []
    for c in ref(m):
        a = cv2.approxPolyDP(c.reshape(-1, 1, 2), eps, True).reshape(-1, 2)
        out.append(np.stack([(a[:, 0] * sx).astype(np.int32), (a[:, 1] * sy).astype(np.int32)], axis=1))   # (int)(x * scale)
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    eng = ms.Engine(None)
    bad = n_contours = n_in = n_out = 0
    for i in range(n):
        h, w = int(rng.integers(4, 260)), int(rng.integers(4, 260))
        if i % 40 == 0:
            h, w = int(rng.integers(300, 900)), int(rng.integers(300, 900))
        eps = float(rng.choice([0.25, 0.5, 0.75, 1.0, 1.25, 1.5, 2.0, 2.5, 3.0, 5.0])) if i % 2 else float(rng.uniform(0.01, 8.0))
        ow, oh = (w, h) if i % 3 else (int(rng.integers(1, 3 * w)), int(rng.integers(1, 3 * h)))
        eng.set_dp_epsilon(eps)
        nb = 3 if i % 7 == 0 else 1
        batch = np.stack([make(rng, h, w) for _ in range(nb)])
        polys = eng.mask2polygon(batch, orig_w=ow, orig_h=oh)
        for j in range(nb):
            wv = want(batch[j], eps, ow / w, oh / h)
            g = polys.slice(j)
            n_contours += len(wv)
            n_in += sum(len(c) for c in ref(batch[j]))
            n_out += sum(len(c) for c in wv)
            if not same(g, wv):
                bad += 1
                if bad < 5:
                    print("MISMATCH", i, h, w, eps, len(g), len(wv))
    print("cases", n, "contours", n_contours, "vertices in", n_in, "kept", n_out, "mismatching slices", bad)
    eng.cleanup()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
