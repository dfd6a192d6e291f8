"""Where the one-CTA-per-slice kernel (csrc/slice_fused.cuh) spends its time: clock64 stamps of slice 0 at the phase
boundaries (MEDSEG_FUSED_DBG=1), for a CT-like slice through the whole path.

    MEDSEG_FUSED_DBG=1 python tools/fused_phases.py [batch]
"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

os.environ["MEDSEG_FUSED_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402

NAMES = ["mask->bits", "holes: scan+label", "fill", "open", "label opened", "keep + write mask", "bg label", "external starts",
         "frame bits + border scan", "crack init", "cut", "pointer jumping", "lengths", "flags", "prefix", "emit"]


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    td = tempfile.mkdtemp()
    blob = ms.make_weight_blob(os.path.join(td, "w.msegw"), 3, 1234)
    e = ms.Engine({"weights": blob, "max_batch": batch})
    vol = synth.ct_volume(batch)
    for _ in range(3):
        e.process_batch(vol)
    buf = (C.c_longlong * 32)()
    assert ms.lib().ms_debug_fused_phases(C.addressof(buf)) == 0
    t = np.array(buf[:17], dtype=np.int64)
    total = t[16] - t[0]
    for i, n in enumerate(NAMES):
        print(f"{n:28s} {int(t[i + 1] - t[i]):9d} cycles  {100.0 * (t[i + 1] - t[i]) / total:5.1f} %")
    print(f"{'total':28s} {int(total):9d} cycles  (~{total / 1.9e3:.1f} us at 1.9 GHz)")
    d = np.array(buf[20:25], dtype=np.int64)
    q = np.array(buf[20:32], dtype=np.int64)
    print("holes label: scan+init %d | hook sweep %d | compress#1 %d (%d rounds) | leftover sweep %d | compress#2 %d (%d rounds) | head x %d | tails+flush %d cycles; runs %d"
          % (q[0] - t[1], q[6] - q[0], q[1] - q[6], q[5] // 100, q[7] - q[1], q[2] - q[7], q[5] % 100, q[8] - q[2], q[3] - q[8], q[4]))
    e.cleanup()


if __name__ == "__main__":
    main()
