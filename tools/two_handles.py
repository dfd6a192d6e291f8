"""Experiment: does running the post stages (K5/K6) of one batch under the UNet of the next help?  Two handles (two
streams, separate workspaces) alternate whole batches; compare with one handle's double-buffered submit/wait.

    python tools/two_handles.py [steps]
"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402


def pinned(a):
    t = torch.from_numpy(a).pin_memory()
    return t, t.numpy()


def run_one(eng, vols, steps):
    eng.submit_batch(0, vols[0])
    for i in range(steps):
        if i + 1 < steps:
            eng.submit_batch((i + 1) % 2, vols[(i + 1) % len(vols)])
        eng.wait_batch(i % 2)


def run_two(engs, vols, steps):
    # batch i goes to handle i % 2, slot (i // 2) % 2; up to three batches are in flight
    def sub(i):
        engs[i % 2].submit_batch((i // 2) % 2, vols[i % len(vols)])
    depth = 3
    for i in range(min(depth, steps)):
        sub(i)
    for i in range(steps):
        engs[i % 2].wait_batch((i // 2) % 2)
        if i + depth < steps:
            sub(i + depth)


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    B = 32
    td = tempfile.mkdtemp()
    blob = ms.make_weight_blob(os.path.join(td, "w.msegw"), 1, 1234)
    cfg = {"weights": blob, "max_batch": B, "head": "binary"}
    keep, vols = [], []
    for k in range(4):
        t, a = pinned(synth.ct_volume(B, first_seed=k * B))
        keep.append(t)
        vols.append(a)
    for mode in ("smem", "crack"):
        os.environ["MEDSEG_TRACE"] = mode
        engs = [ms.Engine(cfg), ms.Engine(cfg)]
        for e in engs:
            run_one(e, vols, 4)
        torch.cuda.synchronize()
        for name, fn in (("one handle ", lambda: run_one(engs[0], vols, steps)), ("two handles", lambda: run_two(engs, vols, steps)),
                         ("one handle ", lambda: run_one(engs[0], vols, steps)), ("two handles", lambda: run_two(engs, vols, steps))):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"MEDSEG_TRACE={mode}  {name}: {B * steps / dt:8.1f} slices/s  ({dt / steps * 1e3:.3f} ms/batch)")
        for e in engs:
            e.cleanup()


if __name__ == "__main__":
    main()
