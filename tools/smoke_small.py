"""Small smoke workload (seconds): every stage kernel, all four contour-ordering variants,
one UNet forward at batch 1, the batched file path.

    python tools/smoke_small.py
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    eng = ms.Engine(None)
    masks = [(rng.random((37, 71)) < 0.5).astype(np.uint8) * 255, synth.stress_mask("blobs", 512, 512, seed=1),
             synth.stress_mask("rings", 512, 512, seed=1), np.zeros((64, 64), np.uint8), np.full((33, 65), 255, np.uint8)]
    for variant in ("rank", "smem", "window", "crack"):
        os.environ["MEDSEG_TRACE"] = variant
        for m in masks:
            p = eng.mask2polygon(m)
            print(variant, m.shape, p.n_contours, p.n_points)
    os.environ.pop("MEDSEG_TRACE")
    cls = rng.integers(0, 3, (2, 512, 512)).astype(np.uint8)
    cls[:, 100:400, 120:380] = 2
    print("postprocess", eng.postprocess(cls).sum())
    print("preprocess", eng.preprocess(synth.ct_slice(1, w=600, h=400)).sum())
    eng.cleanup()
    td = tempfile.mkdtemp()
    blob = ms.make_weight_blob(os.path.join(td, "w.msegw"), 3, 7)
    e = ms.Engine({"weights": blob, "max_batch": 2})
    polys, _, _ = e.process_batch(synth.ct_volume(2))
    print("pipeline", polys.n_contours, polys.n_points)
    for i in range(3):
        synth.ct_slice(i).tofile(os.path.join(td, f"s{i}.raw"))
    print("directory", e.process_directory(td, 512, 512, os.path.join(td, "out")))
    e.cleanup()


if __name__ == "__main__":
    main()
