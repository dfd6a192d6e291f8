"""One launch sequence of the stage kernels on a batch of CT-like class masks (for `ncu -k regex:slice_kernel`)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402
from oracle import pipeline as op  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
e = ms.Engine(None)
vol = synth.ct_volume(batch)
masks = []
for i in range(batch):
    n = op.preprocess_raw(vol[i]).astype(np.int32)
    masks.append(np.where(n > 170, 1, np.where(n > 70, 2, 0)).astype(np.uint8))
m = np.stack(masks)
for _ in range(2):
    clean = e.postprocess(m)
    polys = e.mask2polygon(clean, threshold=1)
print("contours", polys.n_contours, "points", polys.n_points)
e.cleanup()
