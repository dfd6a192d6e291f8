"""A/B of the three contour-ordering kernels of K6 (MEDSEG_TRACE = smem | window | crack) through the device-pointer
C ABI, CUDA-event timed, with cv2.findContours on the host beside them.

    python tools/trace_variants.py [out.json]
"""
import json
import os
import sys
import time

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402
from tools.stage_roofline import timed  # noqa: E402


def main():
    eng = ms.Engine(None)
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream
    rows = []
    cases = [(k, synth.stress_mask(k)[None]) for k in ("blobs", "rings", "checker", "diag", "noise", "sparse")]
    cases.append(("blobs x4", np.stack([synth.stress_mask("blobs", seed=s) for s in range(4)])))
    cases.append(("blobs 512x512 x32", np.stack([synth.stress_mask("blobs", 512, 512, seed=s) for s in range(32)])))
    for name, m in cases:
        b, h, w = m.shape
        d = torch.from_numpy(m).cuda()
        t0 = time.perf_counter()
        for i in range(b):
            cv2.findContours(m[i], cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        row = {"case": name, "shape": [b, h, w], "cv2_ms": (time.perf_counter() - t0) * 1e3}
        for variant in ("rank", "smem", "window", "crack"):
            if variant in ("smem", "rank") and h > 1024:
                continue
            os.environ["MEDSEG_TRACE"] = variant
            polys = eng.mask2polygon_dev(d.data_ptr(), h, w, b, 127, st)
            row[variant + "_ms"] = timed(lambda: eng.mask2polygon_dev(d.data_ptr(), h, w, b, 127, st), iters=5, warm=2)
            row["contours"], row["points"] = polys.n_contours, polys.n_points
        rows.append(row)
        print(row)
    if len(sys.argv) > 1:
        json.dump(rows, open(sys.argv[1], "w"), indent=1)
    eng.cleanup()


if __name__ == "__main__":
    main()
