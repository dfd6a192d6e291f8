"""N3: files/s of the batched directory walk (ms_process_directory) against the per-file loop the reference's main.cpp
runs (ms_process_raw_file per file), both writing all five artefacts per slice to a tmpfs-or-disk directory.

    python tools/dir_throughput.py [n_files] [out.json]
"""
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms  # noqa: E402
from medseg_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    root = tempfile.mkdtemp(prefix="medseg_dir_")
    src = os.path.join(root, "in")
    os.makedirs(src)
    for i in range(n):
        synth.ct_slice(i).tofile(os.path.join(src, f"slice_{i:04d}.raw"))
    blob = ms.make_weight_blob(os.path.join(root, "w.msegw"), n_classes=3, seed=1)
    eng = ms.Engine({"weights": blob, "max_batch": 32})
    rows = {}
    eng.process_directory(src, 512, 512, os.path.join(root, "warm"))          # allocations, page cache
    for writers in ("1", "4", "16", "default"):
        if writers == "default":
            os.environ.pop("MEDSEG_WRITERS", None)
        else:
            os.environ["MEDSEG_WRITERS"] = writers
        out = os.path.join(root, "out_" + writers)
        t0 = time.perf_counter()
        found, good, bad = eng.process_directory(src, 512, 512, out)
        dt = time.perf_counter() - t0
        assert (found, good, bad) == (n, n, 0)
        rows["directory_writers_" + writers] = {"files_per_s": n / dt, "s": dt}
        shutil.rmtree(out)
    m = min(n, 64)
    out = os.path.join(root, "out_single")
    t0 = time.perf_counter()
    for i in range(m):
        eng.process_raw_file(os.path.join(src, f"slice_{i:04d}.raw"), 512, 512, out)
    dt = time.perf_counter() - t0
    rows["per_file_loop"] = {"files_per_s": m / dt, "s": dt, "files": m}
    rows["n_files"] = n
    rows["host_cores"] = os.cpu_count()
    rows["artefact_bytes_per_file"] = sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out) if f.startswith("slice_0000"))
    # every GPU of the box behind one handle ("devices"): the list is split into contiguous blocks, one pipeline per GPU
    import torch
    eng.cleanup()
    g = torch.cuda.device_count()
    if g > 1:
        os.environ.pop("MEDSEG_WRITERS", None)
        eng = ms.Engine({"weights": blob, "max_batch": 32, "devices": list(range(g))})
        eng.process_directory(src, 512, 512, os.path.join(root, "warm_multi"))
        out_m = os.path.join(root, "out_multi")
        t0 = time.perf_counter()
        found, good, bad = eng.process_directory(src, 512, 512, out_m)
        dt = time.perf_counter() - t0
        assert (found, good, bad) == (n, n, 0)
        rows[f"directory_{g}_gpus_one_process"] = {"files_per_s": n / dt, "s": dt}
    print(json.dumps(rows))
    if len(sys.argv) > 2:
        json.dump(rows, open(sys.argv[2], "w"), indent=1)
    eng.cleanup()
    shutil.rmtree(root)


if __name__ == "__main__":
    main()
