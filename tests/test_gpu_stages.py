"""GPU parity tests (run with -m gpu on a B200): the integer / byte stages through the C ABI against the
oracle.  Bar: bit-exact."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import ROOT, contours_equal
from oracle import pipeline as op
import cases
from test_oracle import golden_contours, _random_masks

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


# ----------------------------------------------------------------------------- preprocess
def test_preprocess_golden_hashes(stage_engine):
    with open(os.path.join(GOLD, "preprocess_hashes.json")) as f:
        want = json.load(f)
    for name, src in cases.preprocess_cases().items():
        got = stage_engine.preprocess(src)[0]
        assert hashlib.sha256(got.tobytes()).hexdigest() == want[name], name


def test_preprocess_batch_vs_oracle(stage_engine, ms):
    from medseg_b200 import synth
    vol = synth.ct_volume(5)
    got = stage_engine.preprocess(vol)
    for i in range(5):
        assert (got[i] == op.preprocess_raw(vol[i])).all()
    # per-slice min/max must not leak across the batch
    vol2 = vol.copy()
    vol2[2] = vol2[2] // 2
    got2 = stage_engine.preprocess(vol2)
    assert (got2[1] == got[1]).all() and (got2[2] == op.preprocess_raw(vol2[2])).all()


def test_preprocess_odd_sizes_and_ties(stage_engine):
    rng = np.random.default_rng(1)
    for (w, h) in [(1, 1), (2, 3), (511, 513), (2048, 2048), (1537, 700)]:
        src = rng.integers(0, 65536, (h, w), dtype=np.uint16)
        assert (stage_engine.preprocess(src)[0] == op.preprocess_raw(src)).all(), (w, h)
    # exact .5 ties after normalisation: values 0..510 step 1 with range 510 -> v*255/510 = v/2
    src = (np.arange(512 * 512) % 511).astype(np.uint16).reshape(512, 512)
    assert (stage_engine.preprocess(src)[0] == op.preprocess_raw(src)).all()


# ----------------------------------------------------------------------------- postprocess
def test_postprocess_golden(stage_engine):
    z = np.load(os.path.join(GOLD, "postprocess_small.npz"))
    for i in range(int(z["n"])):
        h, w = z[f"shape_{i}"]
        want = np.unpackbits(z[f"out_{i}"])[:h * w].reshape(h, w).astype(np.uint8) * 2
        got = stage_engine.postprocess(z[f"in_{i}"])
        assert (got == want).all(), i


def _class_mask(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    m = np.zeros((h, w), np.uint8)
    for _ in range(int(rng.integers(1, 4))):
        cy, cx = rng.uniform(0.2, 0.8) * h, rng.uniform(0.2, 0.8) * w
        ry, rx = rng.uniform(0.1, 0.45) * h, rng.uniform(0.1, 0.45) * w
        m[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1] = 2
    for _ in range(int(rng.integers(0, 6))):
        cy, cx, r = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(2, 0.2 * min(h, w))
        m[(yy - cy) ** 2 + (xx - cx) ** 2 < r * r] = rng.integers(0, 2)
    sp = rng.random((h, w))
    m[sp < 0.02] = 0
    m[sp > 0.985] = 2
    m[(sp > 0.5) & (sp < 0.505)] = 1
    return m


def test_postprocess_random_batched(stage_engine):
    rng = np.random.default_rng(9)
    for (h, w, b) in [(512, 512, 6), (256, 320, 3), (97, 33, 4), (1024, 1024, 1)]:
        batch = np.stack([_class_mask(rng, h, w) for _ in range(b)])
        got = stage_engine.postprocess(batch)
        for i in range(b):
            assert (got[i] == op.postprocess_mask(batch[i])).all(), (h, w, i)


def test_postprocess_edge_cases(stage_engine):
    full = np.full((64, 64), 2, np.uint8)
    assert (stage_engine.postprocess(full) == 2).all()                 # full image survives the open
    assert (stage_engine.postprocess(np.zeros((64, 64), np.uint8)) == 0).all()
    m = np.zeros((128, 128), np.uint8); m[:2, :] = 2; m[60:62, :] = 2   # border strip vs interior strip (both < 6 %)
    assert (stage_engine.postprocess(m) == op.postprocess_mask(m)).all()
    m = np.full((100, 100), 2, np.uint8); m[40:60, 40:60] = 1; m[10:12, 10:12] = 0   # class-1 hole is filled too
    got = stage_engine.postprocess(m)
    assert (got == op.postprocess_mask(m)).all() and got[50, 50] == 2
    # other foreground values (cfg4 per-class use)
    m3 = np.where(m == 2, 3, m).astype(np.uint8)
    assert (stage_engine.postprocess(m3, fg_value=3) == op.postprocess_mask(m3, fg=3)).all()
    # idempotence
    g2 = stage_engine.postprocess(got)
    assert (g2 == got).all()


# ----------------------------------------------------------------------------- mask2polygon
def test_mask2polygon_golden(stage_engine):
    for m, want in golden_contours():
        got = stage_engine.mask2polygon(m).slice(0)
        assert contours_equal(got, want)


def test_mask2polygon_random_small(stage_engine):
    n = 0
    for m in _random_masks(400, 23):
        want = op.extract_contours(m)
        assert contours_equal(stage_engine.mask2polygon(m).slice(0), want)
        n += len(want)
    assert n > 1000


def test_mask2polygon_batched_ragged(stage_engine):
    rng = np.random.default_rng(3)
    batch = np.stack([(rng.random((96, 160)) < p).astype(np.uint8) * 255 for p in (0.0, 0.1, 0.5, 0.9, 1.0, 0.02)])
    polys = stage_engine.mask2polygon(batch)
    for i in range(len(batch)):
        assert contours_equal(polys.slice(i), op.extract_contours(batch[i])), i
    assert polys.slice(0) == []


def test_mask2polygon_threshold_and_mapping(stage_engine):
    m = np.zeros((32, 32), np.uint8)
    m[4:20, 4:20] = 128
    m[25:30, 25:30] = 127
    got = stage_engine.mask2polygon(m, orig_w=600, orig_h=400).slice(0)
    want = op.map_contour_points(op.extract_contours(m), 600 / 32, 400 / 32)
    assert len(got) == 1 and contours_equal(got, want)


@pytest.mark.parametrize("kind", ["blobs", "rings", "checker", "diag", "noise", "sparse", "zeros", "ones"])
def test_mask2polygon_stress_2048(stage_engine, ms, kind):
    """cfg5: 2048 x 2048 masks with thousands of nested / touching components."""
    from medseg_b200 import synth
    m = synth.stress_mask(kind)
    want = op.extract_contours(m)
    got = stage_engine.mask2polygon(m).slice(0)
    assert len(got) == len(want)
    assert contours_equal(got, want)


@pytest.mark.parametrize("variant", ["rank", "smem", "window", "crack"])
def test_mask2polygon_trace_variants(stage_engine, ms, monkeypatch, variant):
    """The three contour-ordering kernels (whole slice in shared memory, 256 x 256 window walk, crack list ranking)
    are interchangeable: same golden / random / ragged cases, bit-exact, whichever one MEDSEG_TRACE forces."""
    from medseg_b200 import synth
    monkeypatch.setenv("MEDSEG_TRACE", variant)
    for m, want in golden_contours():
        assert contours_equal(stage_engine.mask2polygon(m).slice(0), want)
    for m in _random_masks(150, 77):
        assert contours_equal(stage_engine.mask2polygon(m).slice(0), op.extract_contours(m))
    rng = np.random.default_rng(9)
    batch = np.stack([(rng.random((70, 131)) < p).astype(np.uint8) * 255 for p in (0.0, 0.3, 1.0, 0.6, 0.05)])
    polys = stage_engine.mask2polygon(batch, orig_w=262, orig_h=35)
    for i in range(len(batch)):
        want = op.map_contour_points(op.extract_contours(batch[i]), 262 / 131, 35 / 70)
        assert contours_equal(polys.slice(i), want), i
    for kind in ("blobs", "rings", "checker", "diag", "noise"):
        m = synth.stress_mask(kind, 512, 512, seed=4)
        assert contours_equal(stage_engine.mask2polygon(m).slice(0), op.extract_contours(m)), kind


def test_mask2polygon_rotation_property(stage_engine, ms):
    """Size-independent property at full size: the multiset of traced border pixels is invariant
    under a 180-degree rotation of the mask (coordinates mirrored)."""
    from medseg_b200 import synth
    m = synth.stress_mask("blobs", 1024, 1024, seed=5)
    a = stage_engine.mask2polygon(m)
    b = stage_engine.mask2polygon(np.ascontiguousarray(m[::-1, ::-1]))
    assert a.n_contours == b.n_contours
    pa = {tuple(p) for p in a.xy}
    pb = {(1023 - x, 1023 - y) for x, y in b.xy}
    # CHAIN_APPROX_SIMPLE keeps direction-change vertices, which are rotation invariant as a set
    assert pa == pb
