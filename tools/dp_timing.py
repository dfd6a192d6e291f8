"""Cost of the opt-in Douglas-Peucker step: K6 (mask2polygon through the device-pointer C ABI, incl. the polygon D2H) on a batch
of CT-like masks with dp_epsilon = 0 and > 0, and on a 2048 x 2048 stress mask.

    python tools/dp_timing.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402
from oracle import pipeline as op  # noqa: E402
from tools.stage_roofline import timed  # noqa: E402


def main():
    eng = ms.Engine(None)
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream
    vol = synth.ct_volume(32)
    masks = np.stack([(op.preprocess_raw(v) > 70).astype(np.uint8) * 255 for v in vol])
    cases = [("32 CT-like 512x512 masks", torch.from_numpy(masks).cuda(), 512, 512, 32)]
    for kind in ("blobs", "noise"):
        m = synth.stress_mask(kind)
        cases.append((f"2048x2048 '{kind}'", torch.from_numpy(m[None]).cuda(), 2048, 2048, 1))
    for name, d, h, w, b in cases:
        row = []
        for eps in (0.0, 1.0, 2.0, 4.0):
            eng.set_dp_epsilon(eps)
            p = eng.mask2polygon_dev(d.data_ptr(), h, w, b, 127, st)
            t = timed(lambda: eng.mask2polygon_dev(d.data_ptr(), h, w, b, 127, st), iters=20)
            row.append(f"eps {eps:g}: {t * 1e3:7.1f} us, {p.n_contours} contours, {p.n_points} vertices")
        print(name)
        for r in row:
            print("   ", r)
    eng.cleanup()


if __name__ == "__main__":
    main()
