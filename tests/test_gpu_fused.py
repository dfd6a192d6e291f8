"""GPU parity tests for the one-CTA-per-slice kernel (csrc/slice_fused.cuh: K5 + K6 in shared memory) -- the default
path for slices that fit on chip.  Every case is checked against the integer oracle (cv2) AND against the multi-kernel
path (MEDSEG_FUSED=0), including the kernel's internal fallbacks: more runs than the shared-memory label tables hold
(global tables), more border pixels / contours than the crack table holds (two-pass border walk), and polygon sets
larger than the device capacities (grow + re-run)."""
import numpy as np
import pytest

from conftest import contours_equal
from oracle import pipeline as op
from test_gpu_stages import _class_mask

pytestmark = pytest.mark.gpu


def _both(monkeypatch, fn):
    monkeypatch.setenv("MEDSEG_FUSED", "1")
    a = fn()
    monkeypatch.setenv("MEDSEG_FUSED", "0")
    b = fn()
    monkeypatch.delenv("MEDSEG_FUSED")
    return a, b


def test_fused_postprocess_matches_oracle_and_multikernel(stage_engine, monkeypatch):
    rng = np.random.default_rng(41)
    for (h, w, n) in [(512, 512, 8), (512, 768, 2), (384, 640, 3), (130, 70, 5), (1, 1, 2), (3, 40, 2), (37, 1, 2)]:
        batch = np.stack([_class_mask(rng, h, w) if min(h, w) >= 16 else rng.integers(0, 3, (h, w)).astype(np.uint8) for _ in range(n)])
        a, b = _both(monkeypatch, lambda: stage_engine.postprocess(batch))
        assert (a == b).all(), (h, w)
        for i in range(n):
            assert (a[i] == op.postprocess_mask(batch[i])).all(), (h, w, i)


def test_fused_noisy_masks_use_global_run_tables(stage_engine, monkeypatch):
    """~65 k runs per slice: far beyond the ~10 k the shared-memory tables hold at 512 x 512."""
    rng = np.random.default_rng(42)
    batch = rng.integers(0, 3, (3, 512, 512)).astype(np.uint8)
    batch[1, 100:400, 100:400] = 2                       # a large component that survives, with speckle holes
    batch[1][rng.random((512, 512)) < 0.02] = 0
    a, b = _both(monkeypatch, lambda: stage_engine.postprocess(batch))
    assert (a == b).all()
    for i in range(3):
        assert (a[i] == op.postprocess_mask(batch[i])).all(), i
    assert (a[1] == 2).sum() > 15728


@pytest.mark.parametrize("kind", ["blobs", "rings", "checker", "diag", "noise", "sparse", "zeros", "ones"])
def test_fused_contours_stress_512(stage_engine, ms, monkeypatch, kind):
    """Thousands of nested / touching components per 512 x 512 slice: crack-table overflow -> two-pass walk, > 16 k
    contours, capacity growth.  Batched so slices reserve their records concurrently."""
    from medseg_b200 import synth
    batch = np.stack([synth.stress_mask(kind, 512, 512, seed=s) for s in (1, 2, 3)])
    a, b = _both(monkeypatch, lambda: stage_engine.mask2polygon(batch, orig_w=777, orig_h=1300))
    assert (a.slice_start == b.slice_start).all() and (a.contour_start == b.contour_start).all() and (a.xy == b.xy).all()
    for i in range(3):
        want = op.map_contour_points(op.extract_contours(batch[i]), 777 / 512, 1300 / 512)
        assert contours_equal(a.slice(i), want), (kind, i)


def test_fused_ragged_batch_and_threshold(stage_engine, monkeypatch):
    rng = np.random.default_rng(43)
    batch = np.stack([(rng.random((200, 333)) < p).astype(np.uint8) * int(v) for p, v in
                      ((0.0, 255), (0.1, 255), (0.5, 128), (0.9, 127), (1.0, 255), (0.02, 200), (0.6, 129))])
    a, b = _both(monkeypatch, lambda: stage_engine.mask2polygon(batch))
    assert (a.xy == b.xy).all() and (a.contour_start == b.contour_start).all()
    for i in range(len(batch)):
        assert contours_equal(a.slice(i), op.extract_contours(batch[i])), i
    assert a.slice(0) == [] and a.slice(3) == []            # threshold(127): 127 is background


def test_fused_whole_path_matches_multikernel(unet_engine, ms, monkeypatch):
    """The pipeline (K1 -> UNet -> fused K5 + K6) returns what the multi-kernel chain returns: masks, polygons, and the
    asynchronous / CUDA-graph entry point on both."""
    from medseg_b200 import synth
    vol = synth.ct_volume(4, first_seed=900)

    def run():
        polys, norm, mask = unet_engine.process_batch(vol, want_norm=True, want_mask=True)
        for it in range(3):
            unet_engine.submit_batch(it % 2, vol)
            got = unet_engine.wait_batch(it % 2)
            assert (got.xy == polys.xy).all() and (got.contour_start == polys.contour_start).all()
        return polys, mask

    (pa, ma), (pb, mb) = _both(monkeypatch, run)
    assert (ma == mb).all()
    assert (pa.slice_start == pb.slice_start).all() and (pa.contour_start == pb.contour_start).all() and (pa.xy == pb.xy).all()
    for i in range(4):
        assert contours_equal(pa.slice(i), op.extract_contours(op.mask_to_image(ma[i]))), i


def test_fused_nested_components_external_only(stage_engine):
    """RETR_EXTERNAL through the fused background labelling: islands inside holes are dropped, a component touching the
    left edge is external, and after postprocess a >= 6 % island inside a >= 6 % ring's hole still is not."""
    m = np.zeros((512, 512), np.uint8)
    m[20:500, 20:500] = 2
    m[60:460, 60:460] = 0            # hole: 160 k px, not filled (>= 15,728)
    m[150:370, 150:370] = 2          # island: 48 k px, survives the area filter
    m[:, 0] = 2                      # a 1-px column on the left edge: erased by the open
    clean = stage_engine.postprocess(m)
    assert (clean == op.postprocess_mask(m)).all() and clean[200, 200] == 2
    polys = stage_engine.mask2polygon(clean, threshold=1)
    want = op.extract_contours(op.mask_to_image(clean))
    assert len(want) == 1 and contours_equal(polys.slice(0), want)
