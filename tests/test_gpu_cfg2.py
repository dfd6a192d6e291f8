"""GPU parity tests AT THE BENCHMARKED CONFIGURATION (BASELINE.json configs[1]: batch 32 of 512 x 512, binary head,
bf16, weight seed 1234 -- exactly the engine bench.py builds).

At batch 32 the per-layer kernel selection differs from the small-batch fixtures (unet.cu: run_layer): every level-3/4
layer and the bottleneck run `conv_halo2_kernel<256, EPI_STORE, 0>` (K = 2304 / 4608 / 9216), which the max_batch-4
engine only reaches at level 2.  These tests compare exactly those launches with the fp32 oracle, and every slice of
the 32 with the integer oracle.  The reference's inference call is /root/reference/src/process.cpp:123-175."""
import numpy as np
import pytest

from conftest import contours_equal
from oracle import pipeline as op

pytestmark = pytest.mark.gpu

B = 32
LOGIT_TOL = 2e-2      # BASELINE.json north_star: "UNet logits within 2e-2 abs (bf16)"
MASK_AGREE = 0.999    # "end-to-end masks at >= 99.9 % pixel agreement"
PROBE = (0, 9, 22, 31)   # slices compared with the fp32 torch oracle (spread over the batch: first / last CTA waves)


@pytest.fixture(scope="module")
def cfg2(ms, tmp_path_factory):
    from medseg_b200 import synth, weights as W
    from oracle.unet_torch import load_unet
    d = tmp_path_factory.mktemp("cfg2")
    blob = ms.make_weight_blob(str(d / "unet.msegw"), n_classes=1, seed=1234)          # bench.py: run_ours
    eng = ms.Engine({"weights": blob, "max_batch": B, "device": 0, "net_h": 512, "net_w": 512, "head": "binary"})
    vol = synth.ct_volume(B, 512, 512, first_seed=0)                                   # bench.py rank 0, batch 0
    arch, w = W.load_blob(blob)
    net = load_unet(w, 1)
    yield eng, vol, net
    eng.cleanup()


def test_cfg2_kernel_selection_is_the_benchmarked_one(cfg2):
    eng, _, _ = cfg2
    k = dict(zip(eng.layer_names(), eng.layer_kernels()))
    # the ten launches of the dominant instantiation (VERDICT r1 "Roofline, recomputed")
    for name in ("enc3a", "enc3b", "enc4a", "enc4b", "bott_a", "bott_b", "dec4a", "dec4b", "dec3a", "dec3b"):
        assert "256" in k[name] and "halo2" in k[name], (name, k[name])
    assert "EPI_HEAD" in k["dec1b_head"]


def test_cfg2_logits_and_deep_activations_vs_fp32_oracle(cfg2):
    """(a) logits of four slices of the 32 vs the fp32 oracle, (b) the bottleneck output `bb` (bott_b: N tile 256,
    K = 9216), the level-4 skip (enc4b, K = 4608), the ConvT output up4 inside `cat4`, and dec4b / dec3b (the decoder's
    <256> launches) at batch 32 vs the torch taps."""
    import torch
    eng, vol, net = cfg2
    norm = eng.preprocess(vol)
    for i in PROBE:
        assert (norm[i] == op.preprocess_raw(vol[i])).all()
    mask, logits = eng.process(norm, want_logits=True)
    assert logits.shape == (B, 1, 512, 512)
    acts = {n: eng.read_activation(n, B) for n in ("bb", "cat4", "d4b", "d3b")}
    for i in PROBE:
        taps = {}
        with torch.no_grad():
            x = torch.from_numpy(norm[i:i + 1].astype(np.float32) / np.float32(255.0))[:, None]
            want = net(x, taps).numpy()[0]
            up4 = net.up1.up(taps["x5"])
            d4 = net.up1.conv(torch.cat([taps["x4"], up4], dim=1))
            up3 = net.up2.up(d4)
            d3 = net.up2.conv(torch.cat([taps["x3"], up3], dim=1))
        err = np.abs(logits[i] - want)
        print("slice %d logits: max %.4g p99.9 %.4g" % (i, err.max(), np.quantile(err, 0.999)))
        assert np.quantile(err, 0.999) < LOGIT_TOL, i
        assert err.max() < 4 * LOGIT_TOL, i
        assert (mask[i] == op.binary_head(logits[i])).all()                       # fused head == head of its own logits
        assert (mask[i] == op.binary_head(want)).mean() >= 0.995, i
        for name, ref, c0, c1 in (("bb", taps["x5"], 0, 1024), ("cat4", taps["x4"], 0, 512), ("cat4", up4, 512, 1024),
                                  ("d4b", d4, 0, 512), ("d3b", d3, 0, 256)):
            ref = ref.numpy()[0]
            C = acts[name].size // (B * ref.shape[1] * ref.shape[2])
            got = acts[name].reshape(B, C, ref.shape[1], ref.shape[2])[i, c0:c1]
            rel = np.abs(got - ref).max() / np.abs(ref).max()
            print("  %s[%d:%d] rel max err %.4g" % (name, c0, c1, rel))
            assert rel < 3e-2, (name, i)


def test_cfg2_every_slice_integer_stages_bit_exact(cfg2):
    """(c) all 32 slices: the kernels' own raw masks through postprocess + mask2polygon, bit-exact against the integer
    oracle; the fused whole-path call returns exactly the stage-by-stage result; end-to-end masks >= 99.9 % for the
    probe slices against the fp32 oracle run from the RAW slice."""
    eng, vol, net = cfg2
    norm = eng.preprocess(vol)
    raw = eng.process(norm)
    clean = eng.postprocess(raw)
    polys, norm2, mask2 = eng.process_batch(vol, want_norm=True, want_mask=True)
    assert (norm2 == norm).all() and (mask2 == clean).all()
    n_contours = 0
    for i in range(B):
        assert (clean[i] == op.postprocess_mask(raw[i])).all(), i
        want = op.map_contour_points(op.extract_contours(op.mask_to_image(clean[i])), 1.0, 1.0)
        assert contours_equal(polys.slice(i), want), i
        n_contours += len(want)
    assert n_contours >= B                      # every CT-like slice yields a body contour
    for i in PROBE:
        ref = op.process_slice(vol[i], net, head="binary", n_classes=1)
        assert (norm[i] == ref["norm"]).all()
        agree = (clean[i] == ref["mask"]).mean()
        print("slice %d final mask agreement %.6f" % (i, agree))
        assert agree >= MASK_AGREE, i
    # the asynchronous (CUDA graph) entry point bench.py's e2e uses returns the same polygons
    for it in range(3):
        eng.submit_batch(it % 2, vol)
        got = eng.wait_batch(it % 2)
        assert (got.slice_start == polys.slice_start).all() and (got.contour_start == polys.contour_start).all() and (got.xy == polys.xy).all()


def test_cfg2_forward_is_deterministic_under_repetition(cfg2):
    """The pipelines hand accumulators, shared-memory stages and staging slabs between warps with relaxed arrives and
    asynchronous-proxy fences; a missing ordering shows up as run-to-run differences long before it shows up as a wrong
    answer.  40 repetitions of the bench batch (kernels back to back, persistent CTAs, PDL between layers): logits, deep
    activations, masks and polygons must be bit-identical every time."""
    import hashlib
    eng, vol, _ = cfg2
    norm = eng.preprocess(vol)
    seen = set()
    for it in range(40):
        mask, logits = eng.process(norm, want_logits=True)
        acts = b"".join(eng.read_activation(n, B).tobytes() for n in ("p1", "bb")) if it in (0, 39) else b""
        seen.add((hashlib.sha256(mask.tobytes()).hexdigest(), hashlib.sha256(logits.tobytes()).hexdigest(),
                  hashlib.sha256(acts).hexdigest() if acts else ""))
    assert len({s[:2] for s in seen}) == 1, "forward pass is not deterministic"
    assert len({s[2] for s in seen if s[2]}) == 1, "intermediate activations are not deterministic"
    polys0, _, _ = eng.process_batch(vol)
    for it in range(6):
        eng.submit_batch(it % 2, vol)
        got = eng.wait_batch(it % 2)
        assert (got.xy == polys0.xy).all() and (got.contour_start == polys0.contour_start).all()
