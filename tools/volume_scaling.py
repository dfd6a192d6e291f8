"""cfg3 in ONE process: a 256-slice volume through ms_process_volume_host on 1, 2, 4, 8 GPUs of the box (whatever is there).

    python tools/volume_scaling.py [out.json]
"""
import json
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402


def main():
    td = tempfile.mkdtemp()
    blob = ms.make_weight_blob(os.path.join(td, "w.msegw"), 1, 1234)
    vol = torch.from_numpy(synth.ct_volume(256)).pin_memory().numpy()
    rows, ref = [], None
    n = 1
    while n <= torch.cuda.device_count():
        e = ms.Engine({"weights": blob, "max_batch": 32, "head": "binary", "devices": list(range(n))})
        for _ in range(3):
            p, _, _ = e.process_volume(vol)
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            p, _, _ = e.process_volume(vol)
        dt = (time.perf_counter() - t0) / reps
        if ref is None:
            ref = p
        same = bool((p.xy == ref.xy).all() and (p.contour_start == ref.contour_start).all() and (p.slice_start == ref.slice_start).all())
        rows.append({"gpus": n, "slices_per_s": 256 / dt, "ms_per_volume": dt * 1e3, "identical_to_1gpu": same})
        print(rows[-1], flush=True)
        e.cleanup()
        n *= 2
    if len(sys.argv) > 1:
        json.dump({"what": "one 256-slice 512x512 volume, one call of ms_process_volume_host, one process", "rows": rows}, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
