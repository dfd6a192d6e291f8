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
