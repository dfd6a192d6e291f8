"""GPU tests of the OPT-IN Douglas-Peucker step ("dp_epsilon", BASELINE.json north_star (3)); off by default, never on the
reference's path (src/mask2polygon.cpp:34 keeps CHAIN_APPROX_SIMPLE).  Bar: vertex-for-vertex what
cv2.approxPolyDP(contour, eps, closed=True) returns for the contours cv2.findContours returns on the same mask, then the
reference's (int)(x * scale) mapping (src/mask2polygon.cpp:41-63)."""
import cv2
import numpy as np
import pytest

from conftest import contours_equal
from oracle import pipeline as op
from test_oracle import _random_masks

pytestmark = pytest.mark.gpu


def _want(mask, eps, sx=1.0, sy=1.0, thr=127):
    cs = op.extract_contours(((mask > thr).astype(np.uint8)) * 255)
    return op.map_contour_points([cv2.approxPolyDP(c.reshape(-1, 1, 2), eps, True).reshape(-1, 2) for c in cs], sx, sy)


@pytest.fixture()
def dp_engine(stage_engine):
    yield stage_engine
    stage_engine.set_dp_epsilon(0.0)


def test_dp_off_is_the_default_and_changes_nothing(dp_engine, ms):
    from medseg_b200 import synth
    assert dp_engine.dp_epsilon() == 0.0
    m = synth.stress_mask("blobs", 512, 512, seed=2)
    a = dp_engine.mask2polygon(m)
    dp_engine.set_dp_epsilon(1.5)
    b = dp_engine.mask2polygon(m)
    dp_engine.set_dp_epsilon(0.0)
    c = dp_engine.mask2polygon(m)
    assert contours_equal(a.slice(0), op.extract_contours(m)) and contours_equal(c.slice(0), a.slice(0))
    assert b.n_contours == a.n_contours and b.n_points < a.n_points


@pytest.mark.parametrize("eps", [0.5, 1.0, 1.5, 2.0, 3.7])
def test_dp_random_small_masks_vs_cv2(dp_engine, eps):
    dp_engine.set_dp_epsilon(eps)
    n = 0
    for m in _random_masks(200, 31):
        want = _want(m, eps)
        assert contours_equal(dp_engine.mask2polygon(m).slice(0), want), eps
        n += len(want)
    assert n > 500


def test_dp_batched_ragged_with_mapping(dp_engine):
    """A ragged batch (empty, dense, full slices), smooth shapes, non-identity mapping, several epsilons."""
    rng = np.random.default_rng(3)
    batch = []
    for p in (0.0, 0.3, 1.0, 0.6, 0.05, 0.5):
        f = cv2.GaussianBlur(rng.random((160, 224)).astype(np.float32), (0, 0), 5.0)
        batch.append(((f > np.quantile(f, 1 - p)) if 0 < p < 1 else np.full(f.shape, p >= 1)).astype(np.uint8) * 255)
    batch = np.stack(batch)
    for eps in (0.75, 1.0, 2.5, 6.0, 40.0, 1000.0):
        dp_engine.set_dp_epsilon(eps)
        polys = dp_engine.mask2polygon(batch, orig_w=448, orig_h=80)
        for i in range(len(batch)):
            assert contours_equal(polys.slice(i), _want(batch[i], eps, 448 / 224, 80 / 160)), (eps, i)
    assert polys.slice(0) == []


@pytest.mark.parametrize("kind", ["blobs", "rings", "noise", "diag"])
def test_dp_long_contours_2048(dp_engine, ms, kind):
    """cfg5-sized masks: contours far longer than the shared-memory staging (global-scratch path), thousands of short ones."""
    from medseg_b200 import synth
    m = synth.stress_mask(kind)
    for eps in (1.0, 2.0):
        dp_engine.set_dp_epsilon(eps)
        got = dp_engine.mask2polygon(m).slice(0)
        want = _want(m, eps)
        assert len(got) == len(want) and contours_equal(got, want), (kind, eps)
    if kind in ("blobs", "diag"):
        assert max(len(c) for c in op.extract_contours(m)) > 6144      # the long-contour path really ran


def test_dp_whole_path_config_key_sync_and_graph(ms, blob3):
    """"dp_epsilon" in the JSON config: the synchronous call, the asynchronous (CUDA graph) call and the multi-class call all
    return the simplified polygons of their own final mask; switching it off at run time restores the reference output."""
    from medseg_b200 import synth
    eps = 2.0
    vol = synth.ct_volume(4, 640, 480, first_seed=21)
    eng = ms.Engine({"weights": blob3, "max_batch": 4, "dp_epsilon": eps})
    try:
        assert eng.dp_epsilon() == eps
        polys, _, mask = eng.process_batch(vol, want_mask=True)
        sx, sy = 640 / 512, 480 / 512
        for i in range(4):
            assert contours_equal(polys.slice(i), _want(op.mask_to_image(mask[i]), eps, sx, sy)), i
        for it in range(4):                      # the second submit per slot replays the captured graph
            eng.submit_batch(it % 2, vol)
            got = eng.wait_batch(it % 2)
            assert (got.contour_start == polys.contour_start).all() and (got.xy == polys.xy).all(), it
        eng.set_dp_epsilon(0.0)
        plain, _, mask2 = eng.process_batch(vol, want_mask=True)
        assert (mask2 == mask).all()
        for i in range(4):
            assert contours_equal(plain.slice(i), op.map_contour_points(op.extract_contours(op.mask_to_image(mask[i])), sx, sy)), i
        for it in range(3):
            eng.submit_batch(it % 2, vol)
            got = eng.wait_batch(it % 2)
            assert (got.xy == plain.xy).all(), it
        assert plain.n_points > polys.n_points
    finally:
        eng.cleanup()
