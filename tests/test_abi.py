"""CPU suite, part 2: the C-ABI library loads, exports every declared symbol, fails loudly without a
GPU, and its host-only code (JSON formatter, weight blob, facade link) is correct."""
import ctypes
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT, HAVE_GPU
import cases


def test_library_exports_every_declared_symbol(ms):
    with open(os.path.join(ROOT, "include", "medseg_b200.h")) as f:
        hdr = f.read()
    declared = set(re.findall(r"MS_API\s+[\w\s\*]+?\b(ms_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    assert declared == set(ms.ABI), declared ^ set(ms.ABI)
    lib = ms.lib()
    for name in declared:
        assert getattr(lib, name) is not None


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(ms):
    with pytest.raises(ms.MedsegError) as ei:
        ms.Engine(None)
    assert ei.value.code == ms.MS_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_bad_arguments_do_not_crash(ms):
    lib = ms.lib()
    assert lib.ms_init(None, None, None) == ms.MS_ERR_ARG
    assert lib.ms_get_info(None, None) == ms.MS_ERR_ARG
    assert lib.ms_launch_count(None) == 0
    assert lib.ms_preprocess_host(None, None, 1, 1, 1, None) == ms.MS_ERR_ARG
    assert lib.ms_last_error(None) is not None
    lib.ms_destroy(None)
    h = ctypes.c_void_p()
    rc = lib.ms_init(b"/nonexistent/weights.msegw", None, ctypes.byref(h))
    assert rc == ms.MS_ERR_IO and not h.value
    assert b"not found" in lib.ms_last_error(None)


def test_json_formatter_byte_exact(ms):
    for name, (base, w, h, contours) in cases.json_cases().items():
        with open(os.path.join(ROOT, "tests", "golden", f"labelme_{name}.json"), "rb") as f:
            want = f.read().decode()
        assert ms.polygons_to_json([np.array(c, np.int32) for c in contours], base, w, h) == want, name


def test_json_formatter_random_vs_oracle(ms):
    from oracle import pipeline as op
    rng = np.random.default_rng(5)
    for _ in range(20):
        cs = [rng.integers(0, 4096, (int(rng.integers(1, 40)), 2)).astype(np.int32) for _ in range(int(rng.integers(0, 5)))]
        assert ms.polygons_to_json(cs, "x_y-1", 777, 333) == op.generate_json(cs, "x_y-1", 777, 333)


def test_weight_blob_roundtrip(ms, tmp_path):
    from medseg_b200 import weights as W
    assert W.n_params(3) == 31_036_611
    w = W.make_weights(7, 3)
    p = str(tmp_path / "w.msegw")
    W.save_blob(p, w, 3)
    arch, w2 = W.load_blob(p)
    assert arch["n_classes"] == 3 and list(w2) == list(w)
    for k in w:
        assert (w[k] == w2[k]).all()
    w3 = W.make_weights(7, 3)
    assert all((w[k] == w3[k]).all() for k in w)   # seeded, reproducible


def test_synth_is_deterministic(ms):
    from medseg_b200 import synth
    a, b = synth.ct_slice(3), synth.ct_slice(3)
    assert a.dtype == np.uint16 and a.shape == (512, 512) and (a == b).all()
    assert 800 < a.min() < 1100 and 2500 < a.max() < 3300
    for k in synth.STRESS_KINDS:
        m = synth.stress_mask(k, 64, 64)
        assert m.dtype == np.uint8 and set(np.unique(m)) <= {0, 255}


@pytest.mark.skipif(not os.path.exists("/root/reference/src/main.cpp"), reason="reference tree not present on this box")
def test_reference_main_links_unchanged(ms, tmp_path):
    """Drop-in check: the reference's own src/main.cpp compiles against include/ and links to the library."""
    exe = str(tmp_path / "seg_main")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-include", "algorithm", "-I", os.path.join(ROOT, "include"),
                        "/root/reference/src/main.cpp", "-o", exe, ms.LIB_PATH, "-Wl,-rpath," + os.path.dirname(ms.LIB_PATH)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([exe], input="help\nprocess x 1 1\nexit\n", capture_output=True, text=True, timeout=60)
    assert out.returncode == 0
    assert "Error: Engine not initialized" in out.stderr
    assert "Exiting..." in out.stdout


def test_json_batch_equals_per_slice_documents(ms):
    """ms_polygons_to_json_batch (threaded, N2): per slice exactly the single-document formatter's bytes; slices without
    contours contribute nothing (the reference writes no file for them, src/mask2polygon.cpp:183-186)."""
    rng = np.random.default_rng(11)
    per_slice = [[rng.integers(-5, 5000, (int(rng.integers(1, 300)), 2)).astype(np.int32) for _ in range(int(k))]
                 for k in (2, 0, 1, 5, 0, 3, 1, 1, 0, 4)]
    flat = [c for s in per_slice for c in s]
    cs = np.zeros(len(flat) + 1, np.int32)
    cs[1:] = np.cumsum([len(c) for c in flat])
    ss = np.zeros(len(per_slice) + 1, np.int32)
    ss[1:] = np.cumsum([len(s) for s in per_slice])
    polys = ms.Polygons(np.concatenate(flat), cs, ss)
    names = [f"vol.{i:03d}" for i in range(len(per_slice))]
    for threads in (1, 3, 0):
        buf, offs = ms.polygons_to_json_batch(polys, names, 1024, 768, n_threads=threads)
        for i, s in enumerate(per_slice):
            got = bytes(buf[offs[i]:offs[i + 1]]).decode()
            assert got == (ms.polygons_to_json(s, names[i], 1024, 768) if s else ""), i
    # too small a buffer: the needed size comes back and nothing is written
    small = np.zeros(16, np.uint8)
    buf, offs = ms.polygons_to_json_batch(polys, names, 1024, 768, buf=small)
    assert buf.size == offs[-1] > 16 and (small == 0).all()


def test_png_writer_and_simd_checksums(tmp_path):
    """csrc/png_min.hpp on the host: PCLMULQDQ CRC-32 and SSSE3 Adler-32 == the table / scalar forms (random lengths and
    alignments, known answers), and the row-streaming encoder round-trips through the decoder with valid chunk CRCs --
    with the SIMD paths and with them switched off.  cv2 (libpng verifies CRC and Adler) reads its files."""
    import cv2
    exe = str(tmp_path / "png_check")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "hostsim", "png_check.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    for env in ({}, {"MEDSEG_SCALAR_CHECKSUMS": "1"}):
        out = subprocess.run([exe], capture_output=True, text=True, env={**os.environ, **env}, timeout=120)
        assert out.returncode == 0 and out.stdout.startswith("ok "), out.stdout
