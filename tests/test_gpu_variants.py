"""GPU tests of kernel VARIANTS that compute the same thing two ways: each pair must agree bit for bit with each other and
with the oracle.  (The selection rules in unet.cu / preprocess.cu pick one form per layer / geometry from measurements;
these tests force the other form through the environment knobs so neither rots.)"""
import os

import numpy as np
import pytest

from oracle import pipeline as op

pytestmark = pytest.mark.gpu


class _Env:
    def __init__(self, **kv):
        self.kv, self.old = kv, {}

    def __enter__(self):
        for k, v in self.kv.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_k1_cluster_kernel_matches_two_kernel_form_and_reference_stage(stage_engine, ms):
    """K1 identity geometry: the 8-CTA cluster kernel (registers + DSMEM min/max) vs minmax + normalise vs the oracle
    (src/preprocess.cpp:65-118).  Cases: CT-like batch, constant slice (min == max, :92), full-range noise, one hot pixel
    in the LAST vector of the slice (owned by the last CTA of the cluster), a batch of 33 (odd number of clusters)."""
    from medseg_b200 import synth
    rng = np.random.default_rng(7)
    vol = synth.ct_volume(33, 512, 512, first_seed=3)
    vol[1] = 1234                                             # constant slice
    vol[2] = rng.integers(0, 65536, (512, 512), dtype=np.uint16)
    vol[3] = 100
    vol[3, 511, 511] = 60000                                  # the max lives in the last CTA's last vector
    vol[4] = 65535
    vol[4, 0, 0] = 0                                          # the min lives in the first CTA's first vector
    with _Env(MEDSEG_K1_CLUSTER="1"):
        a = stage_engine.preprocess(vol)
    with _Env(MEDSEG_K1_CLUSTER="0"):
        b = stage_engine.preprocess(vol)
    assert (a == b).all()
    for i in range(vol.shape[0]):
        assert (a[i] == op.preprocess_raw(vol[i])).all(), i


def test_convt_pair_kernel_matches_per_tap_kernel(ms, blob3, torch_unet3):
    """The four up-sampling layers through the cta_group::2 GEMM (MEDSEG_CONVT_PAIR=all) and through the per-tap kernel
    (=0): identical bf16 outputs (same K order, same epilogue), and both within tolerance of the fp32 oracle's ConvT."""
    import torch
    from medseg_b200 import synth
    vol = synth.ct_volume(4, 512, 512, first_seed=11)
    out = {}
    for mode in ("0", "all"):
        with _Env(MEDSEG_CONVT_PAIR=mode):
            eng = ms.Engine({"weights": blob3, "max_batch": 4})
        kern = dict(zip(eng.layer_names(), eng.layer_kernels()))
        for up in ("up4", "up3", "up2", "up1"):
            assert ("convt_pair" in kern[up]) == (mode == "all"), (mode, kern[up])
        norm = eng.preprocess(vol)
        mask, logits = eng.process(norm, want_logits=True)
        out[mode] = {"logits": logits, "mask": mask, **{n: eng.read_activation(n, 4) for n in ("cat4", "cat3", "cat2", "cat1")}}
        eng.cleanup()
    for k in out["0"]:
        assert np.array_equal(out["0"][k], out["all"][k]), k
    taps = {}
    with torch.no_grad():
        x = torch.from_numpy(norm[:1].astype(np.float32) / np.float32(255.0))[:, None]
        want = torch_unet3(x, taps).numpy()[0]
    err = np.abs(out["all"]["logits"][0] - want)
    assert np.quantile(err, 0.999) < 2e-2 and err.max() < 8e-2


@pytest.mark.parametrize("n_classes", [3, 1])
def test_rowpair_kernel_matches_halo_kernel_and_fp32_oracle(ms, tmp_path_factory, n_classes):
    """The Cout = 64 layers (enc1b with its fused 2x2 max-pool, dec1b + head) through the row-pair kernel (two output rows per
    GEMM row, N = 128 accumulators) and through the halo kernel (MEDSEG_ROWPAIR=0): both within the bf16 tolerance of the
    fp32 oracle's skip map x1, its pooled map and the logits; masks of the two kernels agree to >= 99.99 %."""
    import torch
    import torch.nn.functional as F
    from medseg_b200 import synth, weights as W
    from oracle.unet_torch import load_unet
    blob = ms.make_weight_blob(str(tmp_path_factory.mktemp("rp") / "u.msegw"), n_classes=n_classes, seed=77)
    arch, w = W.load_blob(blob)
    net = load_unet(w, n_classes)
    B = 3
    vol = synth.ct_volume(B, 512, 512, first_seed=40)
    out = {}
    # halo kernels | pair row-pair kernel where the weights stay resident | + dec1a (streamed weights) | single-CTA row-pair kernel
    modes = {"0": dict(MEDSEG_ROWPAIR="0"), "1": dict(MEDSEG_ROWPAIR="1"), "2": dict(MEDSEG_ROWPAIR="2"),
             "s": dict(MEDSEG_ROWPAIR="2", MEDSEG_ROWPAIR2="0")}
    for mode, env in modes.items():
        with _Env(**env):
            cfg = {"weights": blob, "max_batch": B}
            if n_classes == 1:
                cfg["head"] = "binary"
            eng = ms.Engine(cfg)
        kern = dict(zip(eng.layer_names(), eng.layer_kernels()))
        want_kernel = "conv_rowpair_kernel" if mode == "s" else "conv_rowpair2_kernel"
        for name in ("enc1b", "dec1b_head"):
            assert (want_kernel in kern[name]) == (mode != "0"), (mode, kern[name])
        assert (want_kernel in kern["dec1a"]) == (mode in ("2", "s")), (mode, kern["dec1a"])
        norm = eng.preprocess(vol)
        mask, logits = eng.process(norm, want_logits=True)
        out[mode] = {"mask": mask, "logits": logits, "cat1": eng.read_activation("cat1", B).reshape(B, 128, 512, 512)[:, :64],
                     "p1": eng.read_activation("p1", B).reshape(B, 64, 256, 256), "d1a": eng.read_activation("d1a", B)}
        eng.cleanup()
    for mode in ("1", "2", "s"):
        assert (out["0"]["mask"] == out[mode]["mask"]).mean() >= 0.9999
        assert np.abs(out["0"]["d1a"] - out[mode]["d1a"]).max() <= 1e-2 * np.abs(out["0"]["d1a"]).max(), mode
    for i in (0, B - 1):
        taps = {}
        with torch.no_grad():
            x = torch.from_numpy(norm[i:i + 1].astype(np.float32) / np.float32(255.0))[:, None]
            want = net(x, taps).numpy()[0]
            x1 = taps["x1"]
            p1 = F.max_pool2d(x1, 2).numpy()[0]
            x1 = x1.numpy()[0]
        for mode in modes:
            o = out[mode]
            assert np.abs(o["cat1"][i] - x1).max() / np.abs(x1).max() < 2e-2, mode
            assert np.abs(o["p1"][i] - p1).max() / np.abs(p1).max() < 2e-2, mode
            err = np.abs(o["logits"][i] - want)
            assert np.quantile(err, 0.999) < 2e-2 and err.max() < 8e-2, (mode, err.max())
        # the pooled map is exactly the max of the stored skip map (both come from the same bf16 values)
        for mode in ("2", "s"):
            c1 = torch.from_numpy(out[mode]["cat1"][i:i + 1])
            assert np.array_equal(F.max_pool2d(c1, 2).numpy()[0], out[mode]["p1"][i])
