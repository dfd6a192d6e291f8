"""GPU parity against the REFERENCE'S OWN OBJECT CODE: the CUDA stages (through the C ABI) vs tests/golden/ref_pin.*
(outputs of /root/reference/src/{preprocess,postprocess,mask2polygon}.cpp compiled unmodified, see
tests/golden/make_ref_golden.py), and -- where oracle/_ref/libref_pipeline.so travelled to this box -- vs that code live."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import ROOT, contours_equal
from oracle import ref
import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def pin():
    with open(os.path.join(GOLD, "ref_pin_text.json")) as f:
        return np.load(os.path.join(GOLD, "ref_pin.npz")), json.load(f)


def _split(xy, lens):
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(int)
    return [xy[offs[k]:offs[k + 1]] for k in range(len(lens))]


def test_k1_preprocess_vs_reference_object_code(stage_engine, pin):
    z, text = pin
    for name, src in cases.preprocess_cases().items():
        got = stage_engine.preprocess(src)[0]
        assert hashlib.sha256(got.tobytes()).hexdigest() == text[f"pre_sha_{name}"], name
    assert (stage_engine.preprocess(cases.preprocess_cases()["rand_333x517"])[0] == z["pre_full_rand_333x517"]).all()


def test_k5_postprocess_vs_reference_object_code(stage_engine, pin):
    z, _ = pin
    pm = cases.postprocess_case_masks() + cases.ref_mask_cases()
    for i, m in enumerate(pm):
        want = np.unpackbits(z[f"post_{i}"])[:m.size].reshape(m.shape)
        got = stage_engine.postprocess(m)
        assert ((got == 2) == want).all() and set(np.unique(got)) <= {0, 2}, i
    big = np.stack(cases.ref_mask_cases())                       # batched: slices do not interact
    got = stage_engine.postprocess(big)
    n0 = len(cases.postprocess_case_masks())
    for k in range(len(big)):
        assert ((got[k] == 2) == np.unpackbits(z[f"post_{n0 + k}"]).reshape(512, 512)).all(), k


def test_k6_contours_mapping_json_vs_reference_object_code(stage_engine, ms, pin):
    z, text = pin
    for i, m in enumerate(cases.contour_case_masks()):
        lens = z[f"cnt_len_{i}"]
        h, w = m.shape
        assert contours_equal(stage_engine.mask2polygon(m).slice(0), _split(z[f"cnt_xy_{i}"], lens)), i           # scale 1
        got = stage_engine.mask2polygon(m, orig_w=4 * w, orig_h=3 * h).slice(0)                                     # MAP_SCALES[2]
        assert contours_equal(got, _split(z[f"map_xy_{i}_2"], lens)), i
    for name, (base, w, h, contours) in cases.json_cases().items():
        assert ms.polygons_to_json([np.array(c, np.int32) for c in contours], base, w, h) == text[f"labelme_{name}"], name
    # P7e's document for a post-processed 512 x 512 mask of a 600 x 400 slice, stage by stage on the GPU
    clean = stage_engine.postprocess(cases.ref_mask_cases()[0])
    polys = stage_engine.mask2polygon(clean, threshold=1, orig_w=600, orig_h=400)
    assert ms.polygons_to_json(polys.slice(0), "s", 600, 400) == text["psm_json"]


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_pipeline.so did not travel")
def test_stages_vs_live_reference_object_code(stage_engine, ms):
    rng = np.random.default_rng(99)
    for it in range(4):
        w, h = (int(v) for v in rng.integers(100, 1300, 2))
        src = rng.integers(0, int(rng.integers(2, 65536)), (h, w)).astype(np.uint16)
        assert (stage_engine.preprocess(src)[0] == ref.preprocess_raw(src)[0]).all(), (w, h)
    for it in range(10):
        h, w = (int(v) for v in rng.integers(16, 300, 2))
        f = rng.random((h, w))
        for _ in range(int(rng.integers(0, 4))):
            f = (f + np.roll(f, 1, 0) + np.roll(f, -1, 0) + np.roll(f, 1, 1) + np.roll(f, -1, 1)) / 5
        m = np.where(f > np.quantile(f, rng.uniform(0.2, 0.7)), 2, rng.integers(0, 2, (h, w))).astype(np.uint8)
        want = ref.postprocess_mask(m)
        assert (stage_engine.postprocess(m) == want).all(), it
        vis = np.where(m == 2, 255, np.where(m == 1, 128, 0)).astype(np.uint8)
        ow, oh = int(rng.integers(8, 4000)), int(rng.integers(8, 4000))
        contours = ref.extract_contours(vis)
        mapped = ref.map_contour_points(contours, ow / w, oh / h)
        got = stage_engine.mask2polygon(vis, orig_w=ow, orig_h=oh).slice(0)
        assert contours_equal(got, mapped), it
        if mapped:
            assert ms.polygons_to_json(got, "r.%d" % it, ow, oh) == ref.generate_json(mapped, "r.%d" % it, ow, oh)
