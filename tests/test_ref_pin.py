"""CPU suite: the oracle PINNED TO THE REFERENCE'S OWN OBJECT CODE.

tests/golden/ref_pin.npz / ref_pin_text.json hold outputs of /root/reference/src/{preprocess,postprocess,mask2polygon}.cpp
compiled unmodified (oracle/ref_build/Makefile -> oracle/_ref/libref_pipeline.so) and were produced by the committed
tests/golden/make_ref_golden.py.  Part 1 checks oracle/pipeline.py (cv2 4.13) and the plain-C oracle against those
fixtures everywhere.  Part 2 (where the library is present: this container, and the GPU box as a prebuilt file) runs
the reference object code live on fresh random inputs, and checks the OpenCV stub's primitives against cv2."""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import ROOT, contours_equal
from oracle import pipeline as op
from oracle import ref
import cases

GOLD = os.path.join(ROOT, "tests", "golden")
live = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_pipeline.so not built (needs /root/reference)")


@pytest.fixture(scope="module")
def pin():
    with open(os.path.join(GOLD, "ref_pin_text.json")) as f:
        return np.load(os.path.join(GOLD, "ref_pin.npz")), json.load(f)


def _split(xy, lens):
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(int)
    return [xy[offs[k]:offs[k + 1]] for k in range(len(lens))]


# ------------------------------------------------------------------ part 1: oracle vs committed reference outputs
def test_oracle_preprocess_vs_reference_fixture(pin, oracle_c):
    z, text = pin
    for name, src in cases.preprocess_cases().items():
        got = op.preprocess_raw(src)
        assert hashlib.sha256(got.tobytes()).hexdigest() == text[f"pre_sha_{name}"], name
        assert op.sidecar_json_text("slice_007.raw", src.shape[1], src.shape[0]) == text[f"pre_sidecar_{name}"], name
        out = np.empty((512, 512), np.uint8)
        s = np.ascontiguousarray(src)
        oracle_c.orc_preprocess(s.ctypes.data_as(ctypes.c_void_p), src.shape[1], src.shape[0], 512, 512, out.ctypes.data_as(ctypes.c_void_p))
        assert hashlib.sha256(out.tobytes()).hexdigest() == text[f"pre_sha_{name}"], name
    assert (op.preprocess_raw(cases.preprocess_cases()["rand_333x517"]) == z["pre_full_rand_333x517"]).all()


def test_oracle_postprocess_vs_reference_fixture(pin):
    z, _ = pin
    pm = cases.postprocess_case_masks() + cases.ref_mask_cases()
    assert int(z["post_n"]) == len(pm)
    for i, m in enumerate(pm):
        want = np.unpackbits(z[f"post_{i}"])[:m.size].reshape(m.shape)
        assert ((op.postprocess_mask(m) == 2) == want).all(), i
    # the 512 x 512 cases exercise both sides of every rule
    kept = [int(np.unpackbits(z[f"post_{i}"]).sum()) for i in range(len(pm) - 4, len(pm))]
    assert all(k > 15728 for k in kept)


def test_oracle_contours_and_mapping_vs_reference_fixture(pin):
    z, _ = pin
    masks = cases.contour_case_masks()
    assert int(z["cnt_n"]) == len(masks)
    for i, m in enumerate(masks):
        want = _split(z[f"cnt_xy_{i}"], z[f"cnt_len_{i}"])
        got = op.extract_contours(m)
        assert contours_equal(got, want), i
        for k, (sx, sy) in enumerate(cases.MAP_SCALES):
            assert contours_equal(op.map_contour_points(got, sx, sy), _split(z[f"map_xy_{i}_{k}"], z[f"cnt_len_{i}"])), (i, k)


def test_oracle_json_vs_reference_fixture(pin):
    _, text = pin
    for name, (base, w, h, contours) in cases.json_cases().items():
        assert op.generate_json([np.array(c, np.int32) for c in contours], base, w, h) == text[f"labelme_{name}"], name
        # the same document the standalone generator wrote with the same header in round 1
        assert text[f"labelme_{name}"] == open(os.path.join(GOLD, f"labelme_{name}.json")).read()


def test_oracle_process_single_mask_vs_reference_fixture(pin):
    """P7e: the file protocol (sidecar lookup, contour extraction, overlay in network space, mapped JSON)."""
    z, text = pin
    m = op.postprocess_mask(cases.ref_mask_cases()[0])
    vis = op.mask_to_image(m)
    norm = (np.add.outer(np.arange(512), np.arange(512)) % 251).astype(np.uint8)
    contours = op.extract_contours(vis)
    assert op.generate_json(op.map_contour_points(contours, 600 / 512, 400 / 512), "s", 600, 400) == text["psm_json"]
    ov = op.create_overlay_image(contours, norm)
    assert hashlib.sha256(ov.tobytes()).hexdigest() == text["psm_overlay_sha"]
    red = (ov[:, :, 2] == 255) & (ov[:, :, 1] == 0) & (ov[:, :, 0] == 0)
    assert (np.packbits(red) == z["psm_overlay_red"]).all()


# ------------------------------------------------------------------ part 2: live reference object code
@live
def test_reference_object_code_reproduces_fixtures(pin):
    z, text = pin
    src = cases.preprocess_cases()["rand_333x517"]
    px, side = ref.preprocess_raw(src, "slice_007.raw")
    assert (px == z["pre_full_rand_333x517"]).all() and side == text["pre_sidecar_rand_333x517"]
    m = cases.ref_mask_cases()[1]
    n = len(cases.postprocess_case_masks())
    assert ((ref.postprocess_mask(m) == 2) == np.unpackbits(z[f"post_{n + 1}"])[:m.size].reshape(m.shape)).all()


@live
def test_reference_live_vs_oracle_random():
    rng = np.random.default_rng(2027)
    for it in range(6):                                             # P1c: odd sizes, narrow ranges, .5 ties
        w, h = (int(v) for v in rng.integers(40, 900, 2))
        lo = int(rng.integers(0, 60000))
        src = rng.integers(lo, lo + int(rng.integers(1, 5000)), (h, w)).astype(np.uint16)
        px, side = ref.preprocess_raw(src, f"r{it}.raw")
        assert (px == op.preprocess_raw(src)).all(), (w, h)
        assert side == op.sidecar_json_text(f"r{it}.raw", w, h)
    for it in range(25):                                            # P5 + P7a-c on random class masks
        h, w = (int(v) for v in rng.integers(8, 160, 2))
        f = rng.random((h, w))
        for _ in range(int(rng.integers(0, 4))):
            f = (f + np.roll(f, 1, 0) + np.roll(f, -1, 0) + np.roll(f, 1, 1) + np.roll(f, -1, 1)) / 5
        m = np.where(f > np.quantile(f, rng.uniform(0.2, 0.7)), 2, rng.integers(0, 2, (h, w))).astype(np.uint8)
        got = ref.postprocess_mask(m)
        assert (got == op.postprocess_mask(m)).all(), it
        for img in (op.mask_to_image(got), op.mask_to_image(m), (rng.random((h, w)) * 255).astype(np.uint8)):
            a, b = ref.extract_contours(img), op.extract_contours(img)
            assert contours_equal(a, b), it
            sx, sy = float(rng.uniform(0.3, 9)), float(rng.uniform(0.3, 9))
            assert contours_equal(ref.map_contour_points(a, sx, sy), op.map_contour_points(b, sx, sy))
            if b:
                mp = op.map_contour_points(b, sx, sy)
                assert ref.generate_json(mp, "vol.3_x", w * 3, h * 3) == op.generate_json(mp, "vol.3_x", w * 3, h * 3)


@live
def test_reference_process_single_mask_files(tmp_path):
    """process_single_mask end to end through files; also its early exits: no contours -> no JSON (src/mask2polygon.cpp:183-186),
    size mismatch and missing sidecar entry -> swallowed, nothing written (:172-179, :219-221)."""
    import cv2
    vis = np.zeros((512, 512), np.uint8)
    vis[100:300, 150:400] = 255
    vis[120:140, 170:190] = 0
    norm = (np.add.outer(np.arange(512), np.arange(512)) % 199).astype(np.uint8)
    ref.write_png_gray(str(tmp_path / "a_mask.png"), vis)
    ref.write_png_gray(str(tmp_path / "a_normalized.png"), norm)
    assert (cv2.imread(str(tmp_path / "a_mask.png"), cv2.IMREAD_UNCHANGED) == vis).all()       # the stub writes real PNGs
    (tmp_path / "a_sizes.json").write_text(op.sidecar_json_text("a.raw", 640, 480))
    ref.process_single_mask(str(tmp_path / "a_mask.png"), str(tmp_path), str(tmp_path / "a_sizes.json"), str(tmp_path / "a_normalized.png"), "a")
    contours = op.extract_contours(vis)
    assert open(tmp_path / "a.json").read() == op.generate_json(op.map_contour_points(contours, 640 / 512, 480 / 512), "a", 640, 480)
    assert (cv2.imread(str(tmp_path / "a_contour_overlay.png")) == op.create_overlay_image(contours, norm)).all()
    ref.write_png_gray(str(tmp_path / "e_mask.png"), np.zeros((512, 512), np.uint8))
    (tmp_path / "e_sizes.json").write_text(op.sidecar_json_text("e.raw", 640, 480))
    ref.process_single_mask(str(tmp_path / "e_mask.png"), str(tmp_path), str(tmp_path / "e_sizes.json"), str(tmp_path / "a_normalized.png"), "e")
    assert not os.path.exists(tmp_path / "e.json") and not os.path.exists(tmp_path / "e_contour_overlay.png")
    ref.write_png_gray(str(tmp_path / "m_mask.png"), np.full((64, 64), 255, np.uint8))
    (tmp_path / "m_sizes.json").write_text(op.sidecar_json_text("m.raw", 640, 480))
    ref.process_single_mask(str(tmp_path / "m_mask.png"), str(tmp_path), str(tmp_path / "m_sizes.json"), "", "m")
    assert not os.path.exists(tmp_path / "m.json")
    ref.process_single_mask(str(tmp_path / "a_mask.png"), str(tmp_path), str(tmp_path / "a_sizes.json"), "", "zzz")
    assert not os.path.exists(tmp_path / "zzz.json")


@live
def test_opencv_stub_primitives_vs_cv2(tmp_path):
    """The stub's OpenCV primitives are not the reference's: each is checked against cv2 4.13 through the reference
    functions that call it (threshold + findContours, CCL stats + compare + setTo + morphology, drawContours, PNG codec)."""
    import cv2
    rng = np.random.default_rng(5)
    for v in (0, 1, 126, 127, 128, 255):                           # cv::threshold(127): 127 -> 0, 128 -> 255
        img = np.full((3, 3), v, np.uint8)
        assert len(ref.extract_contours(img)) == len(op.extract_contours(img)) == (1 if v > 127 else 0)
    for it in range(12):                                           # general lines too (the stub rounds a DDA; cv2 LINE_8)
        img = np.zeros((64, 64), np.uint8)
        cv2.circle(img, (32, 32), int(rng.integers(3, 28)), 255, -1)
        ref.write_png_gray(str(tmp_path / "c_mask.png"), img)
        ref.write_png_gray(str(tmp_path / "c_n.png"), img // 2)
        (tmp_path / "c.json").write_text(op.sidecar_json_text("c.raw", 64, 64, 64, 64))
        ref.process_single_mask(str(tmp_path / "c_mask.png"), str(tmp_path), str(tmp_path / "c.json"), str(tmp_path / "c_n.png"), "c")
        got = ref.read_png_bgr(str(tmp_path / "c_contour_overlay.png"))
        assert (got == op.create_overlay_image(op.extract_contours(img), img // 2)).all()
        assert (cv2.imread(str(tmp_path / "c_contour_overlay.png")) == got).all()
