"""Pure-Python model of the crack formulation behind the parallel contour-ordering kernels
(csrc/mask2polygon.cu, namespace crack), checked against cv2.findContours.

A crack is a directed pixel edge with foreground on its left: (x, y, s), s = 0 W side walked south, 1 S side walked
east, 2 E side walked north, 3 N side walked west.  succ / pred are the rules the kernels evaluate from the 3x3
neighbourhood.  The test establishes the two facts the kernels rely on:
  * following succ from the W side of a component's raster-first pixel and dropping consecutive repeats of the owner
    pixel yields cv2's CHAIN_APPROX_NONE sequence (hence list ranking can replace border following);
  * CHAIN_APPROX_SIMPLE can be decided per crack from the 3x3 neighbourhood alone: the first crack of each pixel visit
    keeps the pixel iff the direction to the next visited pixel differs from the direction from the previous one.
Reference behaviour: cv::findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE), /root/reference/src/mask2polygon.cpp:34.
"""
import numpy as np
import cv2

DX = [1, 1, 0, -1, -1, -1, 0, 1]     # 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE, y grows downwards
DY = [0, -1, -1, -1, 0, 1, 1, 1]


class CrackImage:
    def __init__(self, fg):
        self.fg = fg
        self.H, self.W = fg.shape

    def F(self, x, y):
        return 0 <= x < self.W and 0 <= y < self.H and bool(self.fg[y, x])

    def code(self, x, y):
        return [self.F(x + DX[d], y + DY[d]) for d in range(8)]

    def succ(self, x, y, s):
        c = self.code(x, y)
        dg, sg = (5 + 2 * s) & 7, (6 + 2 * s) & 7
        if c[dg]:
            return x + DX[dg], y + DY[dg], (s + 3) & 3
        if c[sg]:
            return x + DX[sg], y + DY[sg], s
        return x, y, (s + 1) & 3

    def pred_dir(self, x, y, s):
        c = self.code(x, y)
        for d in ((3 + 2 * s) & 7, (2 + 2 * s) & 7):
            if c[d]:
                return d
        return -1

    def pred(self, x, y, s):
        c = self.code(x, y)
        dg, sg = (3 + 2 * s) & 7, (2 + 2 * s) & 7
        if c[dg]:
            return x + DX[dg], y + DY[dg], (s + 1) & 3
        if c[sg]:
            return x + DX[sg], y + DY[sg], s
        return x, y, (s + 3) & 3


def border_cycle(im, start):
    cyc, c = [], (start[0], start[1], 0)
    while True:
        cyc.append(c)
        n = im.succ(*c)
        assert im.pred(*n) == c          # succ is a permutation
        c = n
        if c == (start[0], start[1], 0):
            return cyc


def visits(cyc):
    out = []
    for x, y, _ in cyc:
        if not out or out[-1] != (x, y):
            out.append((x, y))
    if len(out) > 1 and out[-1] == out[0]:
        out.pop()
    return out


def simple_per_crack(im, start):
    """What flags_kernel computes, crack by crack, in arbitrary order."""
    cyc = border_cycle(im, start)
    L = len(cyc)
    back, s = 0, 0
    while back < 3 and im.pred_dir(start[0], start[1], s) < 0:
        s = (s + 3) & 3
        back += 1
    kept = {}
    order = np.random.default_rng(L).permutation(L)      # any order: the decision is local
    for pos in order:
        x, y, s = cyc[pos]
        code = im.code(x, y)
        pd = im.pred_dir(x, y, s)
        keep = False
        if pd >= 0:
            d_prev = (pd + 4) & 7
            d_out = -1
            for t in range(4):
                ss = (s + t) & 3
                for d in ((5 + 2 * ss) & 7, (6 + 2 * ss) & 7):
                    if d_out < 0 and code[d]:
                        d_out = d
            keep = d_out != d_prev
        elif not any(code) and s == 0:
            keep = True
        if keep:
            kept[(pos + back) % L] = (x, y)
    return [kept[k] for k in sorted(kept)]


def _masks(n, seed):
    rng = np.random.default_rng(seed)
    for it in range(n):
        H, W = int(rng.integers(1, 28)), int(rng.integers(1, 28))
        k = it % 4
        if k == 0:
            m = rng.random((H, W)) < rng.random()
        elif k == 1:
            m = rng.random((H, W)) < 0.5
        elif k == 2:
            m = np.ones((H, W), bool)
            m[rng.random((H, W)) < 0.15] = False
        else:
            f = rng.random((H, W))
            m = (f + np.roll(f, 1, 0) + np.roll(f, 1, 1)) / 3 > 0.5
        yield m


def test_crack_cycle_is_cv2_visit_sequence():
    n = 0
    for m in _masks(400, 1):
        im = CrackImage(m)
        cs, _ = cv2.findContours(m.astype(np.uint8) * 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
        for c in cs:
            pts = [tuple(p) for p in c.reshape(-1, 2).tolist()]
            assert visits(border_cycle(im, pts[0])) == pts
            n += 1
    assert n > 1000


def test_simple_rule_is_local_per_crack():
    n = 0
    for m in _masks(400, 2):
        im = CrackImage(m)
        cs, _ = cv2.findContours(m.astype(np.uint8) * 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        for c in cs:
            pts = [tuple(p) for p in c.reshape(-1, 2).tolist()]
            assert simple_per_crack(im, pts[0]) == pts
            n += 1
    assert n > 1000
