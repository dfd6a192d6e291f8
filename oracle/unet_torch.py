"""ORACLE (test infrastructure only) -- fp32 torch-CPU UNet.

Stand-in for the opaque TensorRT engine the reference launches at
/root/reference/src/process.cpp:94,147.  The only contract the reference pins is the I/O
(`"input"` fp32 [1,1,512,512] -> `"output"` fp32 [1,C,512,512], src/process.cpp:70,81-85);
the architecture is the canonical UNet fixed in SURVEY.md §8(a) row P3 (31,036,611 params).
Parity: unpinned by the reference (no engine, no weights in the repo).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class DoubleConv(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
            nn.Conv2d(cout, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.double_conv(x)


class Down(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(cin, cout))

    def forward(self, x):
        return self.maxpool_conv(x)


class Up(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.up = nn.ConvTranspose2d(cin, cin // 2, kernel_size=2, stride=2)
        self.conv = DoubleConv(cin, cout)

    def forward(self, x1, x2):
        x1 = self.up(x1)
        return self.conv(torch.cat([x2, x1], dim=1))  # skip first, then upsampled


class OutConv(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        return self.conv(x)


class UNet(nn.Module):
    def __init__(self, n_classes=3, in_ch=1):
        super().__init__()
        self.inc = DoubleConv(in_ch, 64)
        self.down1, self.down2, self.down3, self.down4 = Down(64, 128), Down(128, 256), Down(256, 512), Down(512, 1024)
        self.up1, self.up2, self.up3, self.up4 = Up(1024, 512), Up(512, 256), Up(256, 128), Up(128, 64)
        self.outc = OutConv(64, n_classes)

    def forward(self, x, taps=None):
        x1 = self.inc(x)
        x2 = self.down1(x1)
        x3 = self.down2(x2)
        x4 = self.down3(x3)
        x5 = self.down4(x4)
        y = self.up1(x5, x4)
        y = self.up2(y, x3)
        y = self.up3(y, x2)
        y = self.up4(y, x1)
        if taps is not None:
            taps.update(x1=x1, x2=x2, x3=x3, x4=x4, x5=x5, d1=y)
        return self.outc(y)


def load_unet(weights: "dict[str, np.ndarray]", n_classes: int) -> UNet:
    net = UNet(n_classes)
    sd = net.state_dict()
    for k in sd:
        if k.endswith("num_batches_tracked"):
            continue
        sd[k] = torch.from_numpy(np.array(weights[k], dtype=np.float32))
    net.load_state_dict(sd)
    return net.eval()


@torch.no_grad()
def unet_logits(net: UNet, x_u8: np.ndarray) -> np.ndarray:
    """x_u8: [B,H,W] uint8 -> logits fp32 [B,C,H,W].  The /255.0f is
    `preprocess_image` (/root/reference/src/process.cpp:36-39)."""
    x = torch.from_numpy(x_u8.astype(np.float32)) / np.float32(255.0)
    return net(x[:, None]).numpy()
