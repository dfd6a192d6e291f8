"""Batch-1 latency of the whole path (p50 / p90 over 50 runs) and the per-layer UNet times at batch 1."""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms
from medseg_b200 import synth

def main():
    td = tempfile.mkdtemp()
    blob = ms.make_weight_blob(os.path.join(td, "w.msegw"), 3, 1234)
    for mode in ("1", "0"):
        os.environ["MEDSEG_LATENCY_MODE"] = mode
        e = ms.Engine({"weights": blob, "max_batch": 4})
        src = synth.ct_volume(1)
        for _ in range(5): e.process_batch(src)
        lat = []
        for _ in range(50):
            t = time.perf_counter(); e.process_batch(src); lat.append((time.perf_counter() - t) * 1e3)
        tot = 0.0; rows = []
        for i, n in enumerate(e.layer_names()):
            t, fl = e.time_layer(i, 1, 20); tot += t; rows.append(f"{n}:{t*1e3:.0f}us")
        print(f"latency_mode={mode}: p50 {np.median(lat):.3f} ms p90 {np.quantile(lat, 0.9):.3f} ms | UNet batch-1 {tot:.3f} ms | " + " ".join(rows))
        e.cleanup()

if __name__ == "__main__":
    main()
