"""One-off fuzz of the whole per-batch path against the oracle's integer stages: random raw sizes and batch sizes; the
normalised slices must be bit-exact, and postprocess / contours / coordinate mapping must be bit-exact on the kernel's own
UNet masks (the UNet itself is compared with the fp32 oracle in tests/test_gpu_unet.py).

    python tools/fuzz_pipeline.py [n_batches] [seed]
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402
from oracle import pipeline as op  # noqa: E402  (checker only)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    td = tempfile.mkdtemp()
    blob = ms.make_weight_blob(os.path.join(td, "w.msegw"), 3, 11)
    eng = ms.Engine({"weights": blob, "max_batch": 6})
    bad = 0
    contours = 0
    for it in range(n):
        w, h, b = int(rng.integers(64, 1100)), int(rng.integers(64, 1100)), int(rng.integers(1, 7))
        vol = np.stack([synth.ct_slice(int(rng.integers(0, 10000)), w=w, h=h) for _ in range(b)])
        if it % 5 == 0:
            vol[0] = rng.integers(0, 65535, size=(h, w), dtype=np.uint16)          # pure noise
        if it % 7 == 0:
            vol[-1] = 1234                                                          # constant slice: mn == mx
        if it % 2:                                                                  # alternate the sync and the async (graph) path
            polys, norm, mask = eng.process_batch(vol, want_norm=True, want_mask=True)
        else:
            _, norm, mask = eng.process_batch(vol, want_norm=True, want_mask=True)
            eng.submit_batch(0, vol)
            polys = eng.wait_batch(0)
        raw = eng.process(norm)
        for i in range(b):
            ok = (norm[i] == op.preprocess_raw(vol[i])).all() and (mask[i] == op.postprocess_mask(raw[i])).all()
            ref = op.map_contour_points(op.extract_contours(op.mask_to_image(mask[i])), w / 512, h / 512)
            got = polys.slice(i)
            ok = ok and len(ref) == len(got) and all(a.shape == r.shape and (a == r).all() for a, r in zip(got, ref))
            contours += len(ref)
            if not ok:
                bad += 1
                print("MISMATCH", it, i, w, h, b)
    print("batches", n, "contours", contours, "bad", bad)
    eng.cleanup()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
