"""GPU bring-up aid: runs the UNet forward twice -- CUDA-core reference kernels (MEDSEG_NAIVE_CONV=1)
and the tcgen05 path -- on the same slice and reports the first activation buffer that differs."""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402
from oracle import pipeline as op  # noqa: E402

BUFS = ["e1a", "cat1", "p1", "e2a", "cat2", "p2", "e3a", "cat3", "p3", "e4a", "cat4", "p4", "ba", "bb", "d4a", "d4b", "d3a",
        "d3b", "d2a", "d2b", "d1a"]


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    td = tempfile.mkdtemp()
    blob = ms.make_weight_blob(os.path.join(td, "w.msegw"), 3, 1234)
    norm = np.stack([op.preprocess_raw(synth.ct_slice(i)) for i in range(nb)])
    os.environ["MEDSEG_NAIVE_CONV"] = "1"
    en = ms.Engine({"weights": blob, "max_batch": nb})
    os.environ["MEDSEG_NAIVE_CONV"] = "0"
    et = ms.Engine({"weights": blob, "max_batch": nb})
    t = time.time(); mn, ln = en.process(norm, want_logits=True); tn = time.time() - t
    t = time.time(); mt, lt = et.process(norm, want_logits=True); tt = time.time() - t
    print(f"naive {tn*1e3:.1f} ms, tcgen05 {tt*1e3:.1f} ms (first call, includes upload)")
    bad = 0
    for name in BUFS:
        a, b = en.read_activation(name, nb), et.read_activation(name, nb)
        d = np.abs(a - b)
        scale = max(np.abs(a).max(), 1e-6)
        flag = "" if d.max() / scale < 2e-2 else "   <-- DIFFERS"
        bad += bool(flag)
        print(f"{name:5s} n={a.size:10d} |ref|max={scale:9.4f} maxdiff={d.max():9.5f} meandiff={d.mean():9.6f} nan={int(np.isnan(b).sum())}{flag}")
        if flag and bad == 1:
            idx = np.argsort(d)[-5:]
            print("   worst idx", idx, "ref", a[idx], "got", b[idx])
    dl = np.abs(ln - lt)
    print(f"logits maxdiff={dl.max():.5f} mask agreement={(mn == mt).mean():.6f}")
    print("RESULT", "OK" if bad == 0 and dl.max() < 2e-2 else "MISMATCH")


if __name__ == "__main__":
    main()
