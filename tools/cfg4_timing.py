"""cfg4 (1024 x 1024, 4-label argmax head, per-class contours) stage timings at batch 4 through the device-pointer C ABI:
K1, UNet forward, and per class K5 (postprocess with FOREGROUND_VALUE = k) + K6 (contours of mask == k, incl. the polygon D2H).

    python tools/cfg4_timing.py [out.json]
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402
from tools.stage_roofline import timed  # noqa: E402


def main():
    B, S = 4, 1024
    td = tempfile.mkdtemp()
    blob = ms.make_weight_blob(os.path.join(td, "u4.msegw"), n_classes=4, seed=77)
    eng = ms.Engine({"weights": blob, "max_batch": B, "net_h": S, "net_w": S, "n_classes": 4})
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream
    vol = np.stack([synth.ct_slice(i, w=S, h=S) for i in range(B)])
    d_src = torch.from_numpy(vol.view(np.int16)).cuda()
    d_norm = torch.empty((B, S, S), dtype=torch.uint8, device="cuda")
    d_raw = torch.empty_like(d_norm)
    d_clean = torch.empty_like(d_norm)
    out = {"batch": B, "size": S, "gflop_per_slice": eng.info.flops_per_slice / 1e9}
    out["K1_ms"] = timed(lambda: eng.preprocess_dev(d_src.data_ptr(), S, S, B, d_norm.data_ptr(), 0, st))
    out["unet_ms"] = timed(lambda: eng.unet_forward_dev(d_norm.data_ptr(), B, d_raw.data_ptr(), 0, st), iters=10)
    out["unet_tflops"] = eng.info.flops_per_slice * B / out["unet_ms"] / 1e9
    per_class = {}
    for k in (1, 2, 3):
        t5 = timed(lambda: eng.postprocess_dev(d_raw.data_ptr(), d_clean.data_ptr(), S, S, B, k, st))
        p = eng.mask2polygon_dev(d_clean.data_ptr(), S, S, B, k - 1, st)
        t6 = timed(lambda: eng.mask2polygon_dev(d_clean.data_ptr(), S, S, B, k - 1, st))
        per_class[k] = {"K5_ms": t5, "K6_ms": t6, "contours": int(p.n_contours), "points": int(p.n_points)}
    out["per_class"] = per_class
    total = out["K1_ms"] + out["unet_ms"] + sum(v["K5_ms"] + v["K6_ms"] for v in per_class.values())
    out["slices_per_s_device"] = B / total * 1e3
    # end to end from host memory: one C-ABI call (K1 + UNet once, K5 / K6 per label on the device) vs the stage-by-stage composition
    import time
    pin = torch.from_numpy(vol.view(np.int16)).pin_memory().numpy().view(np.uint16)
    for name, fn in (("one_call_ms", lambda: eng.process_batch_multiclass(pin, (1, 2, 3))), ("stage_calls_ms", lambda: eng.process_multiclass(pin, (1, 2, 3)))):
        for _ in range(3):
            fn()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        out[name] = (time.perf_counter() - t0) / 10 * 1e3
    out["slices_per_s_one_call"] = B / out["one_call_ms"] * 1e3
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
    eng.cleanup()


if __name__ == "__main__":
    main()
