"""Runs selected UNet layers a few times (for ncu captures / quick timing).
    python tools/run_layers.py <batch> <iters> [layer,layer,...]"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402


def main():
    batch, iters = int(sys.argv[1]), int(sys.argv[2])
    want = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    td = tempfile.mkdtemp()
    blob = ms.make_weight_blob(os.path.join(td, "w.msegw"), 3, 1234)
    eng = ms.Engine({"weights": blob, "max_batch": batch})
    vol = synth.ct_volume(min(batch, 4))
    vol = np.concatenate([vol] * ((batch + len(vol) - 1) // len(vol)))[:batch]
    eng.process(eng.preprocess(vol))
    for li, name in enumerate(eng.layer_names()):
        if want and name not in want:
            continue
        t, fl = eng.time_layer(li, batch, iters)
        print(f"{name:12s} {t:8.3f} ms  {fl / t / 1e9:8.1f} TFLOP/s")
    eng.cleanup()


if __name__ == "__main__":
    main()
