"""A/B timing of UNet layer kernels under identical thermal / power state: all configurations live in one
process, hold the same realistic activations (a full batch forward runs first) and are timed in an
interleaved order, `reps` times; reports median and min per layer and the SM clock seen.

    python tools/ab_layers.py <batch> <reps> <cfg,cfg,...> [layer,layer,...]
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402

CONFIGS = {
    "pertap": {"MEDSEG_HALO": "0"},
    "halo1": {"MEDSEG_HALO": "1", "MEDSEG_CTA2": "0"},
    "halo2": {"MEDSEG_HALO": "1", "MEDSEG_CTA2": "2"},
    "auto": {},
    "nostream2": {"MEDSEG_STREAM2": "0"},
    "nodeep2": {"MEDSEG_DEEP2": "0"},
    "stream128": {"MEDSEG_RES_BIG": "0"},
}
ENV_KEYS = ("MEDSEG_NAIVE_CONV", "MEDSEG_HALO", "MEDSEG_DESC_MODE", "MEDSEG_HALO_PITCH", "MEDSEG_CTA2", "MEDSEG_STREAM2", "MEDSEG_DEEP2", "MEDSEG_RES_BIG")


def sm_clock():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
    except Exception:
        return None, None


def main():
    batch, reps = int(sys.argv[1]), int(sys.argv[2])
    cfgs = sys.argv[3].split(",")
    want = sys.argv[4].split(",") if len(sys.argv) > 4 else None
    td = tempfile.mkdtemp()
    blob = ms.make_weight_blob(os.path.join(td, "w.msegw"), 3, 1234)
    vol = synth.ct_volume(min(batch, 8))
    vol = np.concatenate([vol] * ((batch + len(vol) - 1) // len(vol)))[:batch]
    engines = {}
    for c in cfgs:
        for k in ENV_KEYS:
            os.environ.pop(k, None)
        os.environ.update(CONFIGS[c])
        e = ms.Engine({"weights": blob, "max_batch": batch})
        e.process(e.preprocess(vol))
        engines[c] = e
    names = engines[cfgs[0]].layer_names()
    layers = [i for i, n in enumerate(names) if not want or n in want]
    times = {c: {i: [] for i in layers} for c in cfgs}
    clocks = []
    for r in range(reps):
        for i in layers:
            for c in cfgs:
                t, fl = engines[c].time_layer(i, batch, 5)
                times[c][i].append(t)
        clocks.append(sm_clock())
    print("clock/power samples:", clocks)
    fl = {i: engines[cfgs[0]].time_layer(i, batch, 1)[1] for i in layers}
    hdr = f"{'layer':12s}" + "".join(f"{c + ' med':>12s}{c + ' min':>12s}{'TF(med)':>9s}" for c in cfgs)
    print(hdr)
    tot = {c: 0.0 for c in cfgs}
    for i in layers:
        row = f"{names[i]:12s}"
        for c in cfgs:
            med, mn = float(np.median(times[c][i])), float(np.min(times[c][i]))
            tot[c] += med
            row += f"{med:12.3f}{mn:12.3f}{fl[i] / med / 1e9:9.0f}"
        print(row)
    print(f"{'total':12s}" + "".join(f"{tot[c]:12.3f}{'':12s}{'':9s}" for c in cfgs))


if __name__ == "__main__":
    main()
