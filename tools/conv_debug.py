"""GPU bring-up aid: runs the UNet forward with the CUDA-core reference kernels (MEDSEG_NAIVE_CONV=1)
and with one or more tcgen05 configurations on the same slice(s), reports the first activation
buffer that differs and the per-layer timings of each configuration.

    python tools/conv_debug.py [n_slices] [time_batch]
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402
from oracle import pipeline as op  # noqa: E402

BUFS = ["e1a", "cat1", "p1", "e2a", "cat2", "p2", "e3a", "cat3", "p3", "e4a", "cat4", "p4", "ba", "bb", "d4a", "d4b", "d3a",
        "d3b", "d2a", "d2b", "d1a"]
CONFIGS = {
    "pertap": {"MEDSEG_HALO": "0"},
    "halo1": {"MEDSEG_HALO": "1", "MEDSEG_CTA2": "0"},
    "halo2": {"MEDSEG_HALO": "1", "MEDSEG_CTA2": "2"},
    "auto": {},
    "nostream2": {"MEDSEG_STREAM2": "0"},
    "nodeep2": {"MEDSEG_DEEP2": "0"},
    "stream128": {"MEDSEG_RES_BIG": "0"},
}


def make_engine(blob, nb, env):
    for k in ("MEDSEG_NAIVE_CONV", "MEDSEG_HALO", "MEDSEG_DESC_MODE", "MEDSEG_HALO_PITCH", "MEDSEG_CTA2", "MEDSEG_STREAM2", "MEDSEG_DEEP2", "MEDSEG_RES_BIG"):
        os.environ.pop(k, None)
    os.environ.update(env)
    return ms.Engine({"weights": blob, "max_batch": nb})


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    tb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    names = sys.argv[3].split(",") if len(sys.argv) > 3 else list(CONFIGS)
    td = tempfile.mkdtemp()
    blob = ms.make_weight_blob(os.path.join(td, "w.msegw"), 3, 1234)
    norm = np.stack([op.preprocess_raw(synth.ct_slice(i)) for i in range(nb)])
    en = make_engine(blob, nb, {"MEDSEG_NAIVE_CONV": "1"})
    mn, ln = en.process(norm, want_logits=True)
    ref = {name: en.read_activation(name, nb) for name in BUFS}
    en.cleanup()
    for cname in names:
        try:
            et = make_engine(blob, max(nb, tb), CONFIGS[cname])
            mt, lt = et.process(norm, want_logits=True)
        except ms.MedsegError as e:
            print(f"[{cname}] FAILED: {e}")
            continue
        bad = 0
        for name in BUFS:
            a, b = ref[name], et.read_activation(name, nb)
            d = np.abs(a - b)
            scale = max(np.abs(a).max(), 1e-6)
            if not (d.max() / scale < 2e-2):
                bad += 1
                if bad <= 3:
                    idx = np.argsort(d)[-4:]
                    print(f"[{cname}] {name:5s} DIFFERS |ref|max={scale:.4f} maxdiff={d.max():.5f} meandiff={d.mean():.6f} nan={int(np.isnan(b).sum())}"
                          f" worst idx {idx.tolist()} ref {a[idx].tolist()} got {b[idx].tolist()}")
        dl = np.abs(ln - lt)
        print(f"[{cname}] buffers differing: {bad}/{len(BUFS)}  logits maxdiff={dl.max():.5f}  mask agreement={(mn == mt).mean():.6f}  "
              f"RESULT {'OK' if bad == 0 and dl.max() < 2e-2 else 'MISMATCH'}")
        if tb:
            tot = 0.0
            rows = []
            for li, lname in enumerate(et.layer_names()):
                t, fl = et.time_layer(li, tb, 10)
                tot += t
                rows.append(f"{lname}:{t:.3f}ms/{fl / t / 1e9:.0f}TF")
            print(f"[{cname}] batch {tb}: UNet {tot:.3f} ms  " + " ".join(rows))
        et.cleanup()


if __name__ == "__main__":
    main()
