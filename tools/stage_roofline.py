"""HBM roofline of the non-tensor stages (K1 preprocess, K5 postprocess, K6 mask2polygon) through the
device-pointer C ABI, CUDA-event timed: achieved = ALGORITHMIC bytes / time.

    python tools/stage_roofline.py [out.json]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402
from oracle import pipeline as op  # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    peak = 6552.6
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    eng = ms.Engine(None)
    ts = torch.cuda.Stream()          # non-default stream shared by the library's launches and the CUDA events
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream
    rows = []
    for batch in (32, 256):
        vol = synth.ct_volume(8)
        vol = np.concatenate([vol] * (batch // 8))
        d_src = torch.from_numpy(vol).cuda()
        d_u8 = torch.empty((batch, 512, 512), dtype=torch.uint8, device="cuda")
        d_bf = torch.empty((batch, 512, 512), dtype=torch.bfloat16, device="cuda")
        t = timed(lambda: eng.preprocess_dev(d_src.data_ptr(), 512, 512, batch, d_u8.data_ptr(), d_bf.data_ptr(), st))
        by = batch * (2 * 512 * 512 + 2 * 512 * 512)
        rows.append({"stage": "K1 preprocess (u16 -> u8 + bf16)", "batch": batch, "ms": t, "alg_bytes": by, "GBps": by / t / 1e6})
        t = timed(lambda: eng.preprocess_dev(d_src.data_ptr(), 512, 512, batch, d_u8.data_ptr(), 0, st))
        rows.append({"stage": "K1 preprocess (u16 -> u8, as in the pipeline)", "batch": batch, "ms": t, "alg_bytes": by, "GBps": by / t / 1e6})
        # a realistic class mask: body ellipse = 2, organs = 1, speckle
        masks = []
        for i in range(8):
            n = op.preprocess_raw(vol[i]).astype(np.int32)
            m = np.where(n > 170, 1, np.where(n > 70, 2, 0)).astype(np.uint8)
            masks.append(m)
        m = np.concatenate([np.stack(masks)] * (batch // 8))
        d_m = torch.from_numpy(m).cuda()
        d_o = torch.empty_like(d_m)
        t = timed(lambda: eng.postprocess_dev(d_m.data_ptr(), d_o.data_ptr(), 512, 512, batch, 2, st))
        by = batch * 2 * 512 * 512
        rows.append({"stage": "K5 postprocess (hole fill + open + area filter)", "batch": batch, "ms": t, "alg_bytes": by, "GBps": by / t / 1e6})
        vis = (d_o == 2).to(torch.uint8) * 255
        polys = eng.mask2polygon_dev(vis.data_ptr(), 512, 512, batch, 127, st)
        t = timed(lambda: eng.mask2polygon_dev(vis.data_ptr(), 512, 512, batch, 127, st), iters=10)
        by = batch * 512 * 512 + 8 * polys.n_points + 4 * polys.n_contours
        rows.append({"stage": "K6 mask2polygon (incl. D2H of polygons)", "batch": batch, "ms": t, "alg_bytes": int(by), "GBps": by / t / 1e6,
                     "contours": polys.n_contours, "points": polys.n_points})
    for kind in ("blobs", "noise"):
        m = synth.stress_mask(kind)
        d = torch.from_numpy(m).cuda()
        polys = eng.mask2polygon_dev(d.data_ptr(), 2048, 2048, 1, 127, st)
        t = timed(lambda: eng.mask2polygon_dev(d.data_ptr(), 2048, 2048, 1, 127, st), iters=5)
        by = 2048 * 2048 + 8 * polys.n_points + 4 * polys.n_contours
        rows.append({"stage": f"K6 mask2polygon cfg5 2048x2048 '{kind}' (incl. D2H)", "batch": 1, "ms": t, "alg_bytes": int(by), "GBps": by / t / 1e6,
                     "contours": polys.n_contours, "points": polys.n_points})
    for r in rows:
        r["frac_of_measured_hbm_peak"] = r["GBps"] / peak
        print(f"{r['stage']:58s} batch {r['batch']:4d}  {r['ms']:8.3f} ms  {r['GBps']:9.1f} GB/s  {100 * r['frac_of_measured_hbm_peak']:5.1f}% of {peak:.0f}")
    if len(sys.argv) > 1:
        json.dump({"hbm_peak_GBps": peak, "rows": rows}, open(sys.argv[1], "w"), indent=1)
    eng.cleanup()


if __name__ == "__main__":
    main()
