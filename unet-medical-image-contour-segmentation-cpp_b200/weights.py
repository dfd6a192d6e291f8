"""Weight blob for the canonical UNet.

The reference loads an opaque TensorRT engine (/root/reference/src/initialize.cpp:48-60)
whose ``*.pt -> *.onnx -> *.trt`` export chain is git-ignored, so there is no checkpoint.
Both the CUDA path and the oracle therefore load the same seeded random-init blob written
here.  Tensor names follow the Pytorch-UNet ``state_dict`` the repo name points at:

    inc / down1..4 : double_conv = [conv3x3(no bias), BN, ReLU] x 2
    up1..4         : ConvTranspose2d(k=2, s=2, bias) + double_conv on cat([skip, up])
    outc           : conv1x1 (bias)

Blob layout (little endian):
    8 B   magic  b"MSEGW001"
    4 B   u32    header length L
    L B   JSON   {"arch": {...}, "tensors": [{"name", "shape", "offset"}...]}  (fp32 only)
    pad to 64 B, then raw fp32 data; offsets are relative to the start of the data section.

BatchNorm tensors are stored *unfolded* (weight, bias, running_mean, running_var); the CUDA
engine folds them into the conv at load time, the torch oracle runs real BatchNorm2d(eval).
"""
from __future__ import annotations

import json
import re
import struct
from collections import OrderedDict

import numpy as np

MAGIC = b"MSEGW001"
BN_EPS = 1e-5
WIDTHS = (64, 128, 256, 512, 1024)


def _double_conv_names(prefix: str, cin: int, cout: int):
    yield f"{prefix}.0.weight", (cout, cin, 3, 3)
    for t in ("weight", "bias", "running_mean", "running_var"):
        yield f"{prefix}.1.{t}", (cout,)
    yield f"{prefix}.3.weight", (cout, cout, 3, 3)
    for t in ("weight", "bias", "running_mean", "running_var"):
        yield f"{prefix}.4.{t}", (cout,)


def tensor_specs(n_classes: int = 3, in_ch: int = 1):
    """Ordered (name, shape) list of the canonical UNet."""
    specs = list(_double_conv_names("inc.double_conv", in_ch, WIDTHS[0]))
    for i in range(4):
        specs += list(_double_conv_names(f"down{i+1}.maxpool_conv.1.double_conv", WIDTHS[i], WIDTHS[i + 1]))
    for i in range(4):
        cin = WIDTHS[4 - i]
        specs.append((f"up{i+1}.up.weight", (cin, cin // 2, 2, 2)))
        specs.append((f"up{i+1}.up.bias", (cin // 2,)))
        specs += list(_double_conv_names(f"up{i+1}.conv.double_conv", cin, cin // 2))
    specs.append(("outc.conv.weight", (n_classes, WIDTHS[0], 1, 1)))
    specs.append(("outc.conv.bias", (n_classes,)))
    return specs


def n_params(n_classes: int = 3) -> int:
    """Trainable parameter count (running stats are buffers): 31,036,611 for n_classes=3."""
    return sum(int(np.prod(s)) for n, s in tensor_specs(n_classes) if "running_" not in n)


def make_weights(seed: int = 1234, n_classes: int = 3, head_gain: float = 0.05, plant: bool = True) -> "OrderedDict[str, np.ndarray]":
    """Seeded, engineered random init (SURVEY.md §7 hard part H1).

    A default-initialised net gives a constant argmax and a purely random one gives
    pixel-level speckle whose argmax flips under bf16 rounding.  This recipe keeps activations
    O(1) through all 23 layers (He-normal convs, mildly randomised BN statistics) and, with
    ``plant=True``, plants one *signal path* (see `plant_signal_path`) so that CT-like phantoms
    yield large, decisive regions of several classes with ragged, noise-driven boundaries.
    """
    rng = np.random.default_rng(seed)
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape in tensor_specs(n_classes):
        if name.endswith("running_mean"):
            t = rng.normal(0.0, 0.1, shape)
        elif name.endswith("running_var"):
            t = rng.uniform(0.8, 1.25, shape)
        elif re.search(r"double_conv\.[14]\.", name):  # BN gamma / beta
            t = rng.uniform(0.9, 1.1, shape) if name.endswith("weight") else rng.normal(0.0, 0.1, shape)
        elif name.startswith("outc"):
            if name.endswith("weight"):
                t = rng.normal(0.0, head_gain / np.sqrt(shape[1]), shape)
            else:
                t = rng.normal(0.0, 0.1, shape)
        elif ".up." in name:
            if name.endswith("weight"):  # ConvT: each output pixel sees Cin inputs
                t = rng.normal(0.0, np.sqrt(2.0 / shape[0]), shape)
            else:
                t = rng.normal(0.0, 0.05, shape)
        else:  # conv3x3, He-normal on fan_in
            fan_in = shape[1] * 9
            t = rng.normal(0.0, np.sqrt(2.0 / fan_in), shape)
        out[name] = np.ascontiguousarray(t, dtype=np.float32)
    if plant:
        plant_signal_path(out, n_classes)
    return out


# thresholds (in normalised intensity, 0..1) at which the winning class changes, and the class
# that wins above each threshold; below the first threshold class 0 wins.
_PLANT = {
    1: ((0.27,), (0,), 1.5),                       # binary head: logit > 0 above 0.27
    3: ((0.27, 0.68), (2, 1), 1.5),                # bg -> 0, body -> 2 (FOREGROUND_VALUE), organs -> 1
    4: ((0.27, 0.68, 0.86), (2, 1, 3), 1.5),       # cfg4: four labels
}


def plant_signal_path(w: "OrderedDict[str, np.ndarray]", n_classes: int) -> None:
    """Route a smoothed copy of the input through channel 0 of the level-1 skip connection
    (inc -> cat -> up4.conv) and let the 1x1 head threshold it, on top of the random features.

    All planted values are exactly representable in bf16.  Channel 0 never reads the other
    channels; every other channel still reads channel 0 like any feature, and the full
    network is still evaluated -- the deep path only perturbs the logits (std ~0.07).
    """
    k16 = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], np.float32) / 16.0

    def ident_bn(prefix):
        w[prefix + ".weight"][0] = 1.0
        w[prefix + ".bias"][0] = 0.0
        w[prefix + ".running_mean"][0] = 0.0
        w[prefix + ".running_var"][0] = 1.0

    a = w["inc.double_conv.0.weight"]; a[0] = 0; a[0, 0] = k16
    ident_bn("inc.double_conv.1")
    a = w["inc.double_conv.3.weight"]; a[0] = 0; a[0, 0] = k16
    ident_bn("inc.double_conv.4")
    a = w["up4.conv.double_conv.0.weight"]; a[0] = 0; a[0, 0, 1, 1] = 1.0   # cat channel 0 = skip channel 0
    ident_bn("up4.conv.double_conv.1")
    a = w["up4.conv.double_conv.3.weight"]; a[0] = 0; a[0, 0, 1, 1] = 1.0
    ident_bn("up4.conv.double_conv.4")
    thr, cls, s0 = _PLANT.get(n_classes, _PLANT[3])
    hw, hb = w["outc.conv.weight"], w["outc.conv.bias"]
    hw[:, 0] = 0.0
    slope, icpt = 0.0, 0.0
    steps = (s0, 2.0, 2.0, 2.0)
    for t, c, ds in zip(thr, cls, steps):
        slope, icpt = slope + ds, icpt - ds * t            # piecewise-linear upper envelope
        if c < n_classes:
            hw[c, 0, 0, 0] = slope
            hb[c] = icpt
    if n_classes > 1:
        hb[0] = 0.0


def save_blob(path: str, weights: "OrderedDict[str, np.ndarray]", n_classes: int) -> None:
    tensors, off = [], 0
    for name, t in weights.items():
        tensors.append({"name": name, "shape": list(t.shape), "offset": off})
        off += (t.size * 4 + 63) // 64 * 64
    header = json.dumps({"arch": {"kind": "unet", "in_ch": 1, "widths": list(WIDTHS), "n_classes": n_classes,
                                  "bn_eps": BN_EPS}, "tensors": tensors}).encode()
    pre = 8 + 4 + len(header)
    pad = (-pre) % 64
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(header)))
        f.write(header)
        f.write(b"\0" * pad)
        for name, t in weights.items():
            b = t.astype("<f4").tobytes()
            f.write(b)
            f.write(b"\0" * ((-len(b)) % 64))


def load_blob(path: str):
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != MAGIC:
        raise ValueError("not a MSEGW001 weight blob: " + path)
    (hl,) = struct.unpack("<I", raw[8:12])
    header = json.loads(raw[12:12 + hl])
    base = (12 + hl + 63) // 64 * 64
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for t in header["tensors"]:
        n = int(np.prod(t["shape"]))
        a = np.frombuffer(raw, dtype="<f4", count=n, offset=base + t["offset"]).reshape(t["shape"])
        out[t["name"]] = a
    return header["arch"], out
