"""Two UNet forward passes at the bench batch (for an ncu capture of exactly one launch per layer: skip the first
pass's 22 conv launches, capture the second pass's 22).

    python tools/forward_once.py [batch] [n_classes]
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    n_classes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    td = tempfile.mkdtemp()
    blob = ms.make_weight_blob(os.path.join(td, "w.msegw"), n_classes, 1234)
    cfg = {"weights": blob, "max_batch": batch}
    if n_classes == 1:
        cfg["head"] = "binary"
    eng = ms.Engine(cfg)
    vol = synth.ct_volume(min(batch, 8))
    vol = np.concatenate([vol] * ((batch + len(vol) - 1) // len(vol)))[:batch]
    norm = eng.preprocess(vol)
    for _ in range(2):
        eng.process(norm)
    print("layers:", ",".join(eng.layer_names()))
    eng.cleanup()


if __name__ == "__main__":
    main()
