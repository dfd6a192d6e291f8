"""CPU suite, part 1: the oracle itself, pinned against the committed golden vectors, and the
OpenCV-free restatements (plain-C oracle, host build of the device trace code) pinned against cv2."""
import ctypes
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, contours_equal
from oracle import pipeline as op
import cases

GOLD = os.path.join(ROOT, "tests", "golden")


def _run_contours(fn, mask, thr=127):
    H, W = mask.shape
    cap, capc = 4 * H * W + 16, H * W + 1
    xy = np.zeros(cap * 2, np.int32)
    cs = np.zeros(capc + 1, np.int32)
    npts = ctypes.c_int64(0)
    mask = np.ascontiguousarray(mask)
    nc = fn(mask.ctypes.data_as(ctypes.c_void_p), H, W, thr, xy.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(cap),
            cs.ctypes.data_as(ctypes.c_void_p), capc, ctypes.byref(npts))
    assert nc >= 0
    return [xy[2 * cs[i]:2 * cs[i + 1]].reshape(-1, 2) for i in range(nc)]


def golden_contours():
    z = np.load(os.path.join(GOLD, "contours_small.npz"))
    for i in range(int(z["n"])):
        h, w = z[f"shape_{i}"]
        m = np.unpackbits(z[f"bits_{i}"])[:h * w].reshape(h, w).astype(np.uint8) * 255
        lens = z[f"len_{i}"]
        xy = z[f"xy_{i}"].astype(np.int32)
        offs = np.concatenate([[0], np.cumsum(lens)])
        yield m, [xy[offs[k]:offs[k + 1]] for k in range(len(lens))]


def test_oracle_contours_match_golden():
    n = 0
    for m, want in golden_contours():
        assert contours_equal(op.extract_contours(m), want)
        n += 1
    assert n == 31


def test_golden_inputs_reproducible():
    masks = cases.contour_case_masks()
    for (m, _), m2 in zip(golden_contours(), masks):
        assert (m == m2).all()


def test_known_answers():
    # rectangle -> TL, BL, BR, TR ; 5-px line -> 2 endpoints ; plus sign -> 12 vertices
    m = np.zeros((8, 8), np.uint8); m[2:6, 2:6] = 255
    assert op.extract_contours(m)[0].tolist() == [[2, 2], [2, 5], [5, 5], [5, 2]]
    m = np.zeros((7, 7), np.uint8); m[3, 1:6] = 255
    assert op.extract_contours(m)[0].tolist() == [[1, 3], [5, 3]]
    m = np.zeros((7, 7), np.uint8); m[1:6, 3] = 255; m[3, 1:6] = 255
    assert len(op.extract_contours(m)[0]) == 12
    m = np.zeros((3, 3), np.uint8); m[1, 1] = 255
    assert op.extract_contours(m)[0].tolist() == [[1, 1]]
    m = np.zeros((3, 3), np.uint8); m[1, 1] = 127
    assert op.extract_contours(m) == []          # threshold(127): 127 -> 0, 128 -> 255
    m[1, 1] = 128
    assert len(op.extract_contours(m)) == 1


def test_c_oracle_contours(oracle_c):
    for m, want in golden_contours():
        assert contours_equal(_run_contours(oracle_c.orc_find_contours, m), want)


def _random_masks(n, seed):
    rng = np.random.default_rng(seed)
    for it in range(n):
        H, W = int(rng.integers(1, 48)), int(rng.integers(1, 48))
        k = it % 5
        if k == 0:
            m = rng.random((H, W)) < rng.random()
        elif k == 1:
            m = np.zeros((H, W), bool)
            for q in range(0, min(H, W) // 2, 2):
                m[q:H - q, q:W - q] = (q // 2) % 2 == 0
            m ^= rng.random((H, W)) < 0.05
        elif k == 2:
            m = np.ones((H, W), bool)
            m[rng.random((H, W)) < 0.1] = 0
        elif k == 3:
            m = rng.random((H, W)) < 0.08
        else:
            f = rng.random((H, W))
            f = (f + np.roll(f, 1, 0) + np.roll(f, 1, 1) + np.roll(f, -1, 0) + np.roll(f, -1, 1)) / 5
            m = f > 0.5
        yield m.astype(np.uint8) * 255


def test_c_oracle_and_hostsim_trace_vs_cv2(oracle_c):
    """Pins the closed form and the exact device trace code against cv2."""
    so = os.path.join(ROOT, "tests", "hostsim", "libtrace_sim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so,
                    os.path.join(ROOT, "tests", "hostsim", "trace_sim.cpp")], check=True)
    sim = ctypes.CDLL(so)
    total = 0
    for m in _random_masks(1500, 11):
        want = op.extract_contours(m)
        assert contours_equal(_run_contours(oracle_c.orc_find_contours, m), want)
        assert contours_equal(_run_contours(sim.sim_find_contours, m), want)
        total += len(want)
    assert total > 5000


def test_oracle_postprocess_golden(oracle_c):
    z = np.load(os.path.join(GOLD, "postprocess_small.npz"))
    for i in range(int(z["n"])):
        h, w = z[f"shape_{i}"]
        m = z[f"in_{i}"]
        want = np.unpackbits(z[f"out_{i}"])[:h * w].reshape(h, w).astype(np.uint8) * 2
        assert (op.postprocess_mask(m) == want).all()
        out = np.zeros_like(m)
        mc = np.ascontiguousarray(m)
        oracle_c.orc_postprocess(mc.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), int(h), int(w), 2,
                                 ctypes.c_float(0.06))
        assert (out == want).all()


def test_min_area_thresholds():
    # int(w*h*0.06f) = 15,728 / 62,914 / 251,658
    assert [op._min_area(s, s) for s in (512, 1024, 2048)] == [15728, 62914, 251658]


def test_oracle_preprocess_golden(oracle_c):
    with open(os.path.join(GOLD, "preprocess_hashes.json")) as f:
        want = json.load(f)
    for name, src in cases.preprocess_cases().items():
        got = op.preprocess_raw(src)
        assert hashlib.sha256(got.tobytes()).hexdigest() == want[name], name
        out = np.zeros((512, 512), np.uint8)
        s = np.ascontiguousarray(src)
        oracle_c.orc_preprocess(s.ctypes.data_as(ctypes.c_void_p), s.shape[1], s.shape[0], 512, 512, out.ctypes.data_as(ctypes.c_void_p))
        assert (out == got).all(), name


def test_argmax_first_max_and_nan(oracle_c):
    lg = np.zeros((3, 2, 2), np.float32)
    lg[:, 0, 0] = [1, 1, 0]          # tie -> lowest index
    lg[:, 0, 1] = [0, 2, 2]          # tie -> 1
    lg[:, 1, 0] = [np.nan, np.nan, np.nan]   # NaN never wins -> 0
    lg[:, 1, 1] = [-1, -2, 5]
    assert op.argmax_first3(lg).tolist() == [[0, 1], [0, 2]]
    out = np.zeros(4, np.uint8)
    l2 = np.ascontiguousarray(lg.reshape(3, 4))
    oracle_c.orc_argmax(l2.ctypes.data_as(ctypes.c_void_p), 3, ctypes.c_size_t(4), out.ctypes.data_as(ctypes.c_void_p))
    assert out.tolist() == [0, 1, 0, 2]


def test_oracle_json_matches_nlohmann_golden():
    for name, (base, w, h, contours) in cases.json_cases().items():
        with open(os.path.join(GOLD, f"labelme_{name}.json"), "rb") as f:
            want = f.read().decode()
        got = op.generate_json([np.array(c, np.int32).reshape(-1, 2) for c in contours], base, w, h)
        assert got == want, name
    with open(os.path.join(GOLD, "sidecar_slice_007.json")) as f:
        assert op.sidecar_json_text("slice_007.raw", 600, 400) == f.read()


def test_map_points_truncates():
    c = [np.array([[3, 5], [511, 511]], np.int32)]
    m = op.map_contour_points(c, 600 / 512, 400 / 512)
    assert m[0].tolist() == [[3, 3], [598, 399]]


def test_as_shipped_file_path_equals_in_memory(tmp_path):
    """The oracle's as-shipped form (every PNG / JSON round trip of src/process.cpp:188-262) gives the in-memory
    restatement's results: PNG is lossless, the sidecar carries the sizes."""
    import cv2
    from medseg_b200 import synth, weights as W
    from oracle.unet_torch import load_unet
    net = load_unet(W.make_weights(1234, 3), 3)
    src = synth.ct_slice(3, w=600, h=400)
    src.tofile(tmp_path / "a.raw")
    r = op.process_single_image_files(str(tmp_path / "a.raw"), 600, 400, str(tmp_path / "out"), net)
    m = op.process_slice(src, net)
    assert [os.path.basename(p) for p in r["written"]] == ["a_normalized.png", "a_original_sizes.json", "a_mask.png", "a_contour_overlay.png", "a.json"]
    assert (cv2.imread(str(tmp_path / "out" / "a_normalized.png"), cv2.IMREAD_UNCHANGED) == m["norm"]).all()
    assert (cv2.imread(str(tmp_path / "out" / "a_mask.png"), cv2.IMREAD_UNCHANGED) == m["vis"]).all()
    assert open(tmp_path / "out" / "a.json").read() == op.generate_json(m["mapped"], "a", 600, 400)
    assert open(tmp_path / "out" / "a_original_sizes.json").read() == op.sidecar_json_text("a.raw", 600, 400)


# ----------------------------------------------------------------------------- Douglas-Peucker (opt-in extra)
DP_EPS = (0.5, 1.0, 1.5, 2.0, 3.7, 10.0)


def dp_fuzz_contours(n_masks, seed):
    """Contours for the D-P pins: speckle / nested / smooth masks (spikes, revisited pixels, long smooth borders) and the
    CT-like body outline."""
    import cv2
    rng = np.random.default_rng(seed)
    for t in range(n_masks):
        h, w = (int(v) for v in rng.integers(8, 160, 2))
        k = t % 3
        if k == 0:
            m = (rng.random((h, w)) < rng.uniform(0.3, 0.7)).astype(np.uint8) * 255
        elif k == 1:
            m = np.zeros((h, w), np.uint8)
            for _ in range(int(rng.integers(1, 6))):
                cv2.ellipse(m, (int(rng.integers(0, w)), int(rng.integers(0, h))), (int(rng.integers(2, w)), int(rng.integers(2, h))),
                            float(rng.uniform(0, 180)), 0, 360, 255, -1)
        else:
            m = cv2.GaussianBlur((rng.random((h, w)) * 255).astype(np.uint8), (0, 0), float(rng.uniform(1, 4)))
            m = (m > 127).astype(np.uint8) * 255
        for c in op.extract_contours(m):
            yield c, float(rng.uniform(0.01, 6.0))


def test_approx_poly_dp_restatement_vs_cv2():
    """The oracle's closed-curve Douglas-Peucker against cv2.approxPolyDP (4.13) itself: same vertices, same start, same
    order -- on tie-prone epsilons (multiples of 0.5) and random ones."""
    import cv2
    n = 0
    for c, eps_r in dp_fuzz_contours(240, 5):
        for eps in DP_EPS + (eps_r,):
            want = cv2.approxPolyDP(c.reshape(-1, 1, 2), eps, True).reshape(-1, 2)
            got = op.approx_poly_dp(c, eps)
            assert got.shape == want.shape and (got == want).all(), (eps, c.tolist())
            n += 1
    assert n > 5000
    # degenerate inputs: one point, two points, all points within eps of each other
    for pts in ([[3, 4]], [[3, 4], [9, 4]], [[0, 0], [1, 0], [1, 1], [0, 1]]):
        c = np.array(pts, np.int32)
        for eps in (0.5, 1.0, 5.0):
            want = cv2.approxPolyDP(c.reshape(-1, 1, 2), eps, True).reshape(-1, 2)
            assert np.array_equal(op.approx_poly_dp(c, eps), want), (pts, eps)
