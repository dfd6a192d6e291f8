"""Batch-1 latency of the whole path (wall clock, p50 / p90 over 100 calls of ms_process_batch_host) and where it goes:
per-stage device times at batch 1 (CUDA events around the device-pointer entry points) and the per-layer UNet times.

    python tools/latency.py [out.json]
"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402
from tools.stage_roofline import timed  # noqa: E402


def main():
    td = tempfile.mkdtemp()
    blob = ms.make_weight_blob(os.path.join(td, "w.msegw"), 3, 1234)
    e = ms.Engine({"weights": blob, "max_batch": 4})
    src = synth.ct_volume(1)
    for _ in range(5):
        e.process_batch(src)
    lat = []
    for _ in range(100):
        t = time.perf_counter()
        e.process_batch(src)
        lat.append((time.perf_counter() - t) * 1e3)
    out = {"p50_ms": float(np.median(lat)), "p90_ms": float(np.quantile(lat, 0.9)), "min_ms": float(np.min(lat))}
    # the same slice through ms_submit_batch_host + ms_wait_batch: the kernel chain is one CUDA graph launch
    pin = torch.from_numpy(src).pin_memory()
    for graph in ("1", "0"):
        os.environ["MEDSEG_GRAPH"] = graph
        eg = ms.Engine({"weights": blob, "max_batch": 4})
        for _ in range(5):
            eg.submit_batch(0, pin.numpy())
            eg.wait_batch(0)
        lat = []
        for _ in range(100):
            t = time.perf_counter()
            eg.submit_batch(0, pin.numpy())
            eg.wait_batch(0)
            lat.append((time.perf_counter() - t) * 1e3)
        out["p50_submit_wait_graph" + graph + "_ms"] = float(np.median(lat))
        eg.cleanup()
    os.environ.pop("MEDSEG_GRAPH")
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream
    d_src = torch.from_numpy(src).cuda()
    d_u8 = torch.empty((1, 512, 512), dtype=torch.uint8, device="cuda")
    d_raw = torch.empty_like(d_u8)
    d_mask = torch.empty_like(d_u8)
    out["K1_preprocess_ms"] = timed(lambda: e.preprocess_dev(d_src.data_ptr(), 512, 512, 1, d_u8.data_ptr(), 0, st), iters=50)
    out["unet_forward_ms"] = timed(lambda: e.unet_forward_dev(d_u8.data_ptr(), 1, d_raw.data_ptr(), 0, st), iters=50)
    out["K5_postprocess_ms"] = timed(lambda: e.postprocess_dev(d_raw.data_ptr(), d_mask.data_ptr(), 512, 512, 1, 2, st), iters=50)
    vis = (d_mask == 2).to(torch.uint8) * 255
    out["K6_mask2polygon_incl_sync_and_d2h_ms"] = timed(lambda: e.mask2polygon_dev(vis.data_ptr(), 512, 512, 1, 127, st), iters=50)
    out["whole_path_device_ms"] = timed(lambda: e.process_batch_dev(d_src.data_ptr(), 512, 512, 1, st), iters=50)
    layers = {}
    for i, n in enumerate(e.layer_names()):
        t, _ = e.time_layer(i, 1, 20)
        layers[n] = round(t * 1e3, 1)
    out["unet_layers_us"] = layers
    out["unet_layers_sum_ms"] = sum(layers.values()) / 1e3
    print(json.dumps(out))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
    e.cleanup()


if __name__ == "__main__":
    main()
