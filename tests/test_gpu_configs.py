"""GPU parity tests for the remaining BASELINE.json configurations: cfg3 (a slice-sharded volume) and
cfg4 (1024 x 1024, four labels, per-class contours)."""
import numpy as np
import pytest

from conftest import contours_equal
from oracle import pipeline as op

pytestmark = pytest.mark.gpu


def test_cfg3_volume_sharded_equals_unsharded(unet_engine, ms):
    """cfg3: a synthetic volume split into contiguous per-rank blocks gives, block by block, exactly the polygons
    of the unsharded run (slices are independent; ranks never communicate)."""
    from medseg_b200 import synth
    from medseg_b200.sharding import shard_range
    vol = synth.ct_volume(8, first_seed=100)
    whole = []
    for i in range(0, 8, 4):                       # max_batch of the fixture engine is 4
        whole += unet_engine.process_batch(vol[i:i + 4])[0].per_slice()
    for world in (2, 4):
        got = []
        for rank in range(world):
            lo, hi = shard_range(8, world, rank)
            got += unet_engine.process_batch(vol[lo:hi])[0].per_slice()
        assert len(got) == 8
        for a, b in zip(got, whole):
            assert contours_equal(a, b)
    assert all(len(c) >= 1 for c in whole)


def test_cfg4_1024_four_labels_per_class_contours(ms, tmp_path):
    from medseg_b200 import synth, weights as W
    from oracle.unet_torch import load_unet, unet_logits
    blob = ms.make_weight_blob(str(tmp_path / "u4.msegw"), n_classes=4, seed=77)
    eng = ms.Engine({"weights": blob, "max_batch": 1, "net_h": 1024, "net_w": 1024, "n_classes": 4})
    assert eng.info.flops_per_slice == 1_539_343_122_432
    src = synth.ct_slice(7, w=1024, h=1024)
    norm = eng.preprocess(src)
    assert (norm[0] == op.preprocess_raw(src, 1024, 1024)).all()
    mask, logits = eng.process(norm, want_logits=True)
    arch, w = W.load_blob(blob)
    want = unet_logits(load_unet(w, 4), norm)
    err = np.abs(logits - want)
    print("1024^2 logits: max err %.4g p99.9 %.4g" % (err.max(), np.quantile(err, 0.999)))
    assert np.quantile(err, 0.999) < 2e-2
    assert (mask[0] == op.argmax_first3(logits[0], 4)).all()
    assert (mask[0] == op.argmax_first3(want[0], 4)).mean() >= 0.995
    assert len(np.unique(mask)) >= 3                               # several labels are present
    raw, per_class = eng.process_multiclass(src, classes=(1, 2, 3))
    assert (raw == mask).all()
    n_found = 0
    for k, (clean, polys) in per_class.items():
        ref_clean = op.postprocess_mask(raw[0], fg=k)
        assert (clean[0] == ref_clean).all(), k
        ref = op.extract_contours(np.where(ref_clean == k, 255, 0).astype(np.uint8))
        assert contours_equal(polys.slice(0), ref), k
        n_found += len(ref)
    assert n_found >= 1
    # the same through ONE C-ABI call (ms_process_batch_multiclass_host: K1 + UNet once, K5 / K6 per label on the device)
    raw2, per_class2 = eng.process_batch_multiclass(src, classes=(1, 2, 3), want_masks=True)
    assert (raw2 == raw).all()
    for k in (1, 2, 3):
        assert (per_class2[k][0] == per_class[k][0]).all(), k
        assert contours_equal(per_class2[k][1].slice(0), per_class[k][1].slice(0)), k
    eng.cleanup()


def test_multiclass_call_512_fused_path_vs_oracle(unet_engine, torch_unet3, ms):
    """The multi-class call at 512 x 512 (3-class argmax head, batch 3, non-identity mapping): every label runs the fused
    K5 + K6 kernel with FOREGROUND_VALUE = k; per label the clean mask equals the oracle's postprocess of the kernel's own
    argmax mask and the polygons equal cv2's on that clean mask, mapped with (int)(x * scale)."""
    from medseg_b200 import synth
    vol = synth.ct_volume(3, 640, 480, first_seed=60)
    raw, per_class = unet_engine.process_batch_multiclass(vol, classes=(2, 1), want_masks=True)
    polys_only = unet_engine.process_batch_multiclass(vol, classes=(2, 1))
    sx, sy = 640 / 512, 480 / 512
    n = 0
    for k in (1, 2):
        clean, polys = per_class[k]
        for i in range(3):
            want_clean = op.postprocess_mask(raw[i], fg=k)
            assert (clean[i] == want_clean).all(), (k, i)
            want = op.map_contour_points(op.extract_contours(np.where(want_clean == k, 255, 0).astype(np.uint8)), sx, sy)
            assert contours_equal(polys.slice(i), want), (k, i)
            assert contours_equal(polys_only[k].slice(i), want), (k, i)
            n += len(want)
    assert n >= 3
    # label 2 alone is the reference's own configuration: identical to the single-label whole-path call
    ref_polys, _, ref_mask = unet_engine.process_batch(vol, want_mask=True)
    assert (per_class[2][0] == ref_mask).all() and (per_class[2][1].xy == ref_polys.xy).all()


@pytest.mark.parametrize("net_h,net_w", [(128, 256), (256, 512), (384, 256), (128, 1536)])
def test_other_net_sizes_whole_path(ms, tmp_path, net_h, net_w):
    """Configurable network size (multiples of 128 x 256): the smallest net has ONE 8 x 16 tile at
    the bottleneck, non-square nets exercise the tile decomposition; logits vs the fp32 oracle, then the integer stages
    bit-exact on the kernel's own mask, with a resample from a different raw size."""
    from medseg_b200 import synth, weights as W
    from oracle.unet_torch import load_unet, unet_logits
    blob = ms.make_weight_blob(str(tmp_path / "u.msegw"), n_classes=3, seed=5)
    eng = ms.Engine({"weights": blob, "max_batch": 3, "net_h": net_h, "net_w": net_w})
    vol = np.stack([synth.ct_slice(300 + i, w=640, h=400) for i in range(3)])
    norm = eng.preprocess(vol)
    for i in range(3):
        assert (norm[i] == op.preprocess_raw(vol[i], net_w, net_h)).all()
    mask, logits = eng.process(norm, want_logits=True)
    arch, w = W.load_blob(blob)
    want = unet_logits(load_unet(w, 3), norm)
    err = np.abs(logits - want)
    print(net_h, net_w, "logits: max err %.4g p99.9 %.4g" % (err.max(), np.quantile(err, 0.999)))
    assert np.quantile(err, 0.999) < 2e-2
    for i in range(3):
        assert (mask[i] == op.argmax_first3(logits[i])).all()
    polys, norm2, clean = eng.process_batch(vol, want_norm=True, want_mask=True)
    assert (norm2 == norm).all()
    for i in range(3):
        assert (clean[i] == op.postprocess_mask(mask[i])).all()
        ref = op.map_contour_points(op.extract_contours(op.mask_to_image(clean[i])), 640 / net_w, 400 / net_h)
        assert contours_equal(polys.slice(i), ref), i
    eng.cleanup()


def test_two_devices_in_one_process(ms, blob3):
    """One handle per GPU inside ONE process: same results on both devices, calls interleaved."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from medseg_b200 import synth
    engs = [ms.Engine({"weights": blob3, "max_batch": 4, "device": d}) for d in (0, 1)]
    vols = [synth.ct_volume(4, first_seed=500 + 4 * i) for i in range(3)]
    res = [[], []]
    for v in vols:
        for d in (1, 0):
            res[d].append(engs[d].process_batch(v)[0])
    for a, b in zip(res[0], res[1]):
        assert a.n_contours == b.n_contours and (a.xy == b.xy).all() and (a.contour_start == b.contour_start).all()
    for d in (0, 1):                       # the asynchronous (graph) path too
        for i in range(4):
            engs[d].submit_batch(i % 2, vols[i % 3])
            got = engs[d].wait_batch(i % 2)
            assert (got.xy == res[d][i % 3].xy).all()
    for e in engs:
        e.cleanup()


def test_cfg3_volume_call_streams_sub_batches(unet_engine, ms):
    """cfg3 through ONE call (ms_process_volume_host): a 10-slice volume on a max_batch-4 handle is streamed as 4 + 4 + 2
    through the double-buffered pair and comes back, in slice order, as exactly the per-batch results; the side outputs
    (normalised slices, clean masks) switch the worker to the synchronous per-batch call and give the same polygons."""
    from medseg_b200 import synth
    vol = synth.ct_volume(10, first_seed=700)
    want, wn, wm = [], [], []
    for i in range(0, 10, 4):
        p, n, m = unet_engine.process_batch(vol[i:i + 4], want_norm=True, want_mask=True)
        want += p.per_slice()
        wn.append(n)
        wm.append(m)
    for it in range(2):                              # second pass replays the slots' CUDA graphs
        got, _, _ = unet_engine.process_volume(vol)
        assert len(got.slice_start) == 11 and got.n_contours == sum(len(c) for c in want)
        for a, b in zip(got.per_slice(), want):
            assert contours_equal(a, b)
    got, norm, mask = unet_engine.process_volume(vol, want_norm=True, want_mask=True)
    assert (norm == np.concatenate(wn)).all() and (mask == np.concatenate(wm)).all()
    for a, b in zip(got.per_slice(), want):
        assert contours_equal(a, b)
    assert unet_engine.device_count() == 1


def test_cfg3_volume_sharded_over_gpus_in_one_process(ms, blob3, tmp_path):
    """cfg3 as written: one volume, one call, contiguous slice blocks per GPU ("devices"), one host thread per GPU, results
    concatenated in slice order == the single-GPU result; the file-list path fans out the same way and writes the same
    bytes."""
    import os
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from medseg_b200 import synth
    g = min(4, torch.cuda.device_count())
    one = ms.Engine({"weights": blob3, "max_batch": 4, "device": 0})
    many = ms.Engine({"weights": blob3, "max_batch": 4, "devices": list(range(g))})
    assert many.device_count() == g
    vol = synth.ct_volume(4 * g + 3, first_seed=800)
    a, _, _ = one.process_volume(vol)
    for it in range(2):
        b, _, _ = many.process_volume(vol)
        assert (a.slice_start == b.slice_start).all() and (a.contour_start == b.contour_start).all() and (a.xy == b.xy).all()
    src = tmp_path / "in"
    src.mkdir()
    for i in range(len(vol)):
        vol[i].tofile(src / f"s{i:03d}.raw")
    assert one.process_directory(str(src), 512, 512, str(tmp_path / "o1")) == (len(vol), len(vol), 0)
    assert many.process_directory(str(src), 512, 512, str(tmp_path / "oN")) == (len(vol), len(vol), 0)
    for f in sorted(os.listdir(tmp_path / "o1")):
        assert open(tmp_path / "o1" / f, "rb").read() == open(tmp_path / "oN" / f, "rb").read(), f
    assert many.launch_count() > 0
    one.cleanup()
    many.cleanup()
