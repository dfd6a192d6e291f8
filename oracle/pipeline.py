"""ORACLE (test infrastructure only) -- CPU restatement of the reference's per-slice path.

Every function cites the reference lines it restates (paths relative to /root/reference).
The integer/byte stages call the *same OpenCV functions* the reference calls, through
cv2 (opencv-python-headless 4.13.0; the reference's own OpenCV version is unpinned -- it has
no build file).  The reference has no tests, fixtures or golden vectors; this
module is pinned by
  * oracle/_ref/libref_pipeline.so -- the reference's OWN preprocess.cpp / postprocess.cpp /
    mask2polygon.cpp compiled unmodified against an OpenCV stub (oracle/ref_build, binding
    oracle/ref.py; tests/test_ref_pin.py, tests/golden/ref_*),
  * tests/golden/*  generated from this module by tests/golden/make_golden.py, and
  * oracle/c/medseg_oracle.c, an OpenCV-free restatement checked against this module.
What stays pinned to cv2 4.13 only (no reference object code can exist for it) is the arithmetic the
reference delegates to OpenCV itself: findContours, connectedComponentsWithStats, morphologyEx.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product never does.
"""
from __future__ import annotations

import numpy as np
import cv2

NET = 512
FOREGROUND_VALUE = 2          # src/postprocess.cpp:5
MORPH_KERNEL_SIZE = 3         # src/postprocess.cpp:8
MIN_AREA_RATIO = np.float32(0.06)  # src/postprocess.cpp:9


# --------------------------------------------------------------------------- preprocess
def compute_minmax(src: np.ndarray):
    """src/preprocess.cpp:65-74."""
    return int(src.min()), int(src.max())


def preprocess_raw(src_u16: np.ndarray, out_w: int = NET, out_h: int = NET) -> np.ndarray:
    """src/preprocess.cpp:81-118: top-left aligned bilinear in double, min/max normalise,
    `(uchar)(... + 0.5)` truncation.  `src_u16` is h x w row-major.  The reference
    hard-codes out=512x512 (:81); other sizes are the cfg4 extension."""
    h, w = src_u16.shape
    step_x = float(w) / out_w                                   # :82
    step_y = float(h) / out_h                                   # :83
    mn, mx = compute_minmax(src_u16)                            # :91
    if mn == mx:
        mx = mn + 1                                             # :92
    scale8 = 255.0 / float(mx - mn)                             # :93
    fx = np.arange(out_w, dtype=np.float64) * step_x            # :100
    fy = np.arange(out_h, dtype=np.float64) * step_y
    ix = fx.astype(np.int64)                                    # :101
    iy = fy.astype(np.int64)                                    # :102
    ix1 = np.minimum(ix + 1, w - 1)                             # :103
    iy1 = np.minimum(iy + 1, h - 1)                             # :104
    dx = (fx - ix)[None, :]                                     # :105
    dy = (fy - iy)[:, None]
    s = src_u16.astype(np.float64)
    v00 = s[iy][:, ix]
    v01 = s[iy][:, ix1]
    v10 = s[iy1][:, ix]
    v11 = s[iy1][:, ix1]
    # :112-115, left-to-right evaluation, no fused multiply-add
    v = ((1 - dx) * (1 - dy)) * v00
    v = v + (dx * (1 - dy)) * v01
    v = v + ((1 - dx) * dy) * v10
    v = v + (dx * dy) * v11
    q = (v - mn) * scale8 + 0.5                                 # :116
    return q.astype(np.uint8)                                   # truncation toward zero; 0.5 <= q < 256


def size_sidecar(filename: str, w: int, h: int, out_w: int = NET, out_h: int = NET) -> dict:
    """src/preprocess.cpp:126-132."""
    return {filename: {"original_width": w, "original_height": h, "scaled_width": out_w, "scaled_height": out_h}}


def preprocess_image(gray_u8: np.ndarray) -> np.ndarray:
    """src/process.cpp:36-39: float(u8) / 255.0f."""
    return gray_u8.astype(np.float32) / np.float32(255.0)


# --------------------------------------------------------------------------- argmax
def argmax_first3(logits: np.ndarray, n: int = 3) -> np.ndarray:
    """src/process.cpp:158-170: strict `>` from -FLT_MAX over channels 0..n-1 (the reference
    hard-codes n=3); ties -> lowest index, NaN -> 0.  logits [C,H,W] -> u8 [H,W]."""
    maxp = np.full(logits.shape[1:], -np.finfo(np.float32).max, np.float32)
    idx = np.zeros(logits.shape[1:], np.uint8)
    for c in range(n):
        upd = logits[c] > maxp
        maxp = np.where(upd, logits[c], maxp)
        idx[upd] = c
    return idx


def binary_head(logits: np.ndarray, fg: int = FOREGROUND_VALUE) -> np.ndarray:
    """cfg2 extension (not in the reference): one logit, mask = fg where logit > 0."""
    return np.where(logits[0] > 0, fg, 0).astype(np.uint8)


# --------------------------------------------------------------------------- postprocess
def _min_area(w: int, h: int, ratio=MIN_AREA_RATIO) -> int:
    # src/postprocess.cpp:30,66: static_cast<int>(w * h * 0.06f) -- int*int then float32 multiply
    return int(np.float32(w * h) * np.float32(ratio))


def fill_holes_inside_foreground(mask: np.ndarray, fg: int = FOREGROUND_VALUE, ratio=MIN_AREA_RATIO) -> None:
    """src/postprocess.cpp:13-44 (in place)."""
    binm = np.where(mask == fg, 255, 0).astype(np.uint8)                       # :18
    inv = cv2.bitwise_not(binm)                                                # :22
    nc, labels, stats, _ = cv2.connectedComponentsWithStats(inv, connectivity=8)  # :26
    h, w = mask.shape
    min_area = _min_area(w, h, ratio)                                          # :30
    for i in range(1, nc):
        left, top, ww, hh, area = (int(v) for v in stats[i])
        right, bottom = left + ww - 1, top + hh - 1
        if left > 0 and top > 0 and right < w - 1 and bottom < h - 1 and area < min_area:  # :40
            mask[labels == i] = fg                                             # :41


def postprocess_mask(src: np.ndarray, fg: int = FOREGROUND_VALUE, ratio=MIN_AREA_RATIO,
                     ksize: int = MORPH_KERNEL_SIZE) -> np.ndarray:
    """src/postprocess.cpp:47-79."""
    mask = src.copy()
    fill_holes_inside_foreground(mask, fg, ratio)                              # :54
    binm = np.where(mask == fg, 255, 0).astype(np.uint8)                       # :57
    kernel = cv2.getStructuringElement(cv2.MORPH_RECT, (ksize, ksize))         # :58-59
    binm = cv2.morphologyEx(binm, cv2.MORPH_OPEN, kernel)                      # :60
    nc, labels, stats, _ = cv2.connectedComponentsWithStats(binm, connectivity=8)  # :64
    min_area = _min_area(mask.shape[1], mask.shape[0], ratio)                  # :66
    keep = np.zeros(mask.shape, np.uint8)
    for i in range(1, nc):
        if int(stats[i, cv2.CC_STAT_AREA]) >= min_area:                        # :70
            keep[labels == i] = 255
    out = np.zeros_like(mask)                                                  # :75
    out[keep > 0] = fg                                                         # :76
    return out


def mask_to_image(mask: np.ndarray) -> np.ndarray:
    """src/process.cpp:178-185: LUT 1->128, 2->255, else 0."""
    lut = np.zeros(256, np.uint8)
    lut[1], lut[2] = 128, 255
    return lut[mask]



# --------------------------------------------------------------------------- Douglas-Peucker (opt-in, not in the reference)
def approx_poly_dp(contour, eps: float) -> np.ndarray:
    """Closed-curve Douglas-Peucker as cv2.approxPolyDP(contour, eps, True) of OpenCV 4.13 computes it (the reference never
    simplifies: src/mask2polygon.cpp:34 stops at CHAIN_APPROX_SIMPLE; BASELINE.json north_star (3) names D-P as an extra, so it
    is opt-in, "dp_epsilon" = 0 = off).  Restated from the library's behaviour and pinned by fuzzing against cv2 itself
    (tests/test_oracle.py::test_approx_poly_dp_restatement_vs_cv2):
      1. three sweeps "farthest point from the current start" fix the two initial split points (the output starts at the second);
      2. a slice (s0, s1) is split at the interior point farthest FROM THE SEGMENT (not the line: points that project outside
         the segment count with their distance to the nearer end), first maximum in walk order, while that distance > eps;
      3. the start points of the leaf slices, in walk order, go through one clean-up pass that drops a vertex when it lies
         within eps / sqrt(2) of the chord of its neighbours, the chord is not axis-parallel and the path does not turn back.
    Distances are compared exactly (integers on the common scale |s1 - s0|^2); the eps test is `double(num) <= eps^2 * double(den)`.
    `contour`: (n, 2) integer points.  Returns (m, 2) int32."""
    pts = [(int(x), int(y)) for x, y in np.asarray(contour).reshape(-1, 2)]
    count = len(pts)
    if count == 0:
        return np.zeros((0, 2), np.int32)
    eps2 = float(eps) * float(eps)
    nxt = lambda p: p + 1 if p + 1 < count else 0
    out, stack = [], []
    pos, right_start, le_eps, start = 0, 0, False, pts[0]
    for _ in range(3):
        max_d = 0
        pos = (pos + right_start) % count
        start, pos = pts[pos], nxt(pos)
        for j in range(1, count):
            pt, pos = pts[pos], nxt(pos)
            d = (pt[0] - start[0]) ** 2 + (pt[1] - start[1]) ** 2
            if d > max_d:
                max_d, right_start = d, j
        le_eps = float(max_d) <= eps2
    if le_eps:
        out.append(start)
    else:
        s0 = pos % count
        s1 = (right_start + s0) % count
        stack += [(s1, s0), (s0, s1)]
    while stack:
        s0, s1 = stack.pop()
        a, b = pts[s0], pts[s1]
        pos = nxt(s0)
        if pos == s1:
            out.append(a)
            continue
        dx, dy = b[0] - a[0], b[1] - a[1]
        L = dx * dx + dy * dy
        best, split = 0, s0
        while pos != s1:
            pt = pts[pos]
            px, py = pt[0] - a[0], pt[1] - a[1]
            dot = px * dx + py * dy
            if L == 0 or dot < 0:
                v = (px * px + py * py) * max(L, 1)
            elif dot > L:
                v = ((pt[0] - b[0]) ** 2 + (pt[1] - b[1]) ** 2) * L
            else:
                v = (py * dx - px * dy) ** 2
            if v > best:
                best, split = v, pos
            pos = nxt(pos)
        if float(best) <= eps2 * float(max(L, 1)):
            out.append(a)
        else:
            stack += [(split, s1), (s0, split)]
    # clean-up pass, in place with wrap-around exactly as the library does it
    count = new_count = len(out)
    dst = list(out)
    rd = lambda p: (dst[p], p + 1 if p + 1 < count else 0)
    start, pos = rd(count - 1)
    wpos = pos
    pt, pos = rd(pos)
    i = 0
    while i < count and new_count > 2:
        end, pos = rd(pos)
        dx, dy = float(end[0] - start[0]), float(end[1] - start[1])
        dist = abs((pt[0] - start[0]) * dy - (pt[1] - start[1]) * dx)
        sip = (pt[0] - start[0]) * (end[0] - pt[0]) + (pt[1] - start[1]) * (end[1] - pt[1])
        if dist * dist <= 0.5 * eps2 * (dx * dx + dy * dy) and dx != 0 and dy != 0 and sip >= 0:
            new_count -= 1
            dst[wpos] = start = end
            wpos = wpos + 1 if wpos + 1 < count else 0
            pt, pos = rd(pos)
            i += 2
            continue
        dst[wpos] = start = pt
        wpos = wpos + 1 if wpos + 1 < count else 0
        pt = end
        i += 1
    return np.array(dst[:new_count], dtype=np.int32).reshape(-1, 2)


def simplify_contours(contours, eps: float):
    """The opt-in step between extract_contours and map_contour_points: every contour through approx_poly_dp (eps <= 0: off)."""
    if eps <= 0:
        return list(contours)
    return [approx_poly_dp(c, eps) for c in contours]

# --------------------------------------------------------------------------- mask2polygon
def extract_contours(mask_img: np.ndarray):
    """src/mask2polygon.cpp:29-36.  Returns a list of int32 [n,2] (x,y) arrays."""
    _, binary = cv2.threshold(mask_img, 127, 255, cv2.THRESH_BINARY)           # :31
    contours, _ = cv2.findContours(binary, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)  # :34
    return [c.reshape(-1, 2).astype(np.int32) for c in contours]


def map_contour_points(contours, scale_x: float, scale_y: float):
    """src/mask2polygon.cpp:41-63: (int)(pt * scale) in double, truncation toward zero."""
    out = []
    for c in contours:
        m = np.empty_like(c)
        m[:, 0] = (c[:, 0].astype(np.float64) * scale_x).astype(np.int32)      # :54
        m[:, 1] = (c[:, 1].astype(np.float64) * scale_y).astype(np.int32)      # :55
        out.append(m)
    return out


def _dump(v, indent: int, cur: int) -> str:
    """nlohmann::json::dump(4) restated for the value kinds generate_json produces
    (include/nlohmann/json.hpp 3.12.0 serializer::dump, pretty-print branch): objects with
    alphabetical keys (std::map), arrays one element per line, `{}` / `[]` for empties."""
    if v is None:
        return "null"
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if isinstance(v, str):
        return '"' + v.replace("\\", "\\\\").replace('"', '\\"') + '"'
    pad, pad_in = " " * cur, " " * (cur + indent)
    if isinstance(v, dict):
        if not v:
            return "{}"
        items = [f'{pad_in}"{k}": {_dump(v[k], indent, cur + indent)}' for k in sorted(v)]
        return "{\n" + ",\n".join(items) + "\n" + pad + "}"
    if isinstance(v, (list, tuple)):
        if not len(v):
            return "[]"
        items = [pad_in + _dump(e, indent, cur + indent) for e in v]
        return "[\n" + ",\n".join(items) + "\n" + pad + "]"
    raise TypeError(type(v))


def generate_json(contours, base_name: str, original_width: int, original_height: int) -> str:
    """src/mask2polygon.cpp:68-109 -- returns the exact file text (`f << std::setw(4) << j << std::endl`)."""
    shapes = []
    for c in contours:
        pts = [[int(x), int(y)] for x, y in c]
        shapes.append({"label": 1, "labelIndex": 0, "points": pts if pts else None, "shape_type": "polygon",
                       "description": "", "mask": None, "group_id": None, "flags": {}})
    j = {"version": "1.0.2.812", "imagePath": base_name + ".raw", "imageData": None, "flags": {},
         "shapes": shapes, "imageWidth": original_width, "imageHeight": original_height}
    return _dump(j, 4, 0) + "\n"


def sidecar_json_text(filename: str, w: int, h: int, out_w: int = NET, out_h: int = NET) -> str:
    """src/preprocess.cpp:133-134: `jf << j << std::endl` (compact dump, alphabetical keys)."""
    d = size_sidecar(filename, w, h, out_w, out_h)[filename]
    inner = ",".join(f'"{k}":{d[k]}' for k in sorted(d))
    return '{"' + filename + '":{' + inner + "}}\n"


def create_overlay_image(contours, gray_u8: np.ndarray) -> np.ndarray:
    """src/mask2polygon.cpp:114-129: red 1-px drawContours on the normalised image (BGR)."""
    img = cv2.cvtColor(gray_u8, cv2.COLOR_GRAY2BGR)
    cv2.drawContours(img, [c.reshape(-1, 1, 2) for c in contours], -1, (0, 0, 255), 1)
    return img


# --------------------------------------------------------------------------- whole slice
def process_slice(src_u16: np.ndarray, net, head: str = "argmax", n_classes: int = 3):
    """In-memory restatement of process_single_image (src/process.cpp:188-262) without the
    PNG/JSON disk round-trips (PNG is lossless, so values are identical).  `net` is an
    oracle.unet_torch.UNet.  Returns dict of every intermediate."""
    from .unet_torch import unet_logits
    h, w = src_u16.shape
    norm = preprocess_raw(src_u16)                                  # :211
    logits = unet_logits(net, norm[None])[0]                        # :224
    raw_mask = argmax_first3(logits, 3) if head == "argmax" else binary_head(logits)
    mask = postprocess_mask(raw_mask)                               # :231
    vis = mask_to_image(mask)                                       # :234
    contours = extract_contours(vis)                                # src/mask2polygon.cpp:182
    mapped = map_contour_points(contours, w / NET, h / NET)         # :199-203
    return dict(norm=norm, logits=logits, raw_mask=raw_mask, mask=mask, vis=vis, contours=contours, mapped=mapped)


def process_single_image_files(raw_path: str, width: int, height: int, output_dir: str, net, head: str = "argmax") -> dict:
    """The reference AS SHIPPED: process_single_image (src/process.cpp:188-262) with every PNG / JSON round trip it
    makes -- preprocess_raw writes `_normalized.png` + `_original_sizes.json` (src/preprocess.cpp:121-134), the PNG is read
    back for inference (:217), `_mask.png` is written (:236-239) and read back by process_single_mask
    (src/mask2polygon.cpp:166), which reads the PNG again for the overlay (:117) and writes `_contour_overlay.png` and
    `<stem>.json`.  Returns the paths written."""
    import json
    import os
    from .unet_torch import unet_logits
    os.makedirs(output_dir, exist_ok=True)
    stem = os.path.splitext(os.path.basename(raw_path))[0]
    png = os.path.join(output_dir, stem + "_normalized.png")
    sizes = os.path.join(output_dir, stem + "_original_sizes.json")
    mask_png = os.path.join(output_dir, stem + "_mask.png")
    src = np.fromfile(raw_path, dtype=np.uint16, count=width * height).reshape(height, width)     # MMapFile, :28-61
    cv2.imwrite(png, preprocess_raw(src))                                                       # :122
    with open(sizes, "w") as f:
        f.write(sidecar_json_text(os.path.basename(raw_path), width, height))                   # :126-134
    gray = cv2.imread(png, cv2.IMREAD_UNCHANGED)                                                # src/process.cpp:217
    logits = unet_logits(net, gray[None])[0]                                                    # :224
    raw_mask = argmax_first3(logits, 3) if head == "argmax" else binary_head(logits)
    cv2.imwrite(mask_png, mask_to_image(postprocess_mask(raw_mask)))                            # :231-239
    written = [png, sizes, mask_png]
    vis = cv2.imread(mask_png, cv2.IMREAD_GRAYSCALE)                                            # src/mask2polygon.cpp:166
    with open(sizes) as f:
        info = json.load(f)[os.path.basename(raw_path)]                                        # :146-160
    contours = extract_contours(vis)                                                            # :182
    if contours:                                                                                # :183-186
        overlay = os.path.join(output_dir, stem + "_contour_overlay.png")
        cv2.imwrite(overlay, create_overlay_image(contours, cv2.imread(png, cv2.IMREAD_GRAYSCALE)))   # :114-129, 189-193
        mapped = map_contour_points(contours, info["original_width"] / info["scaled_width"], info["original_height"] / info["scaled_height"])
        out_json = os.path.join(output_dir, stem + ".json")
        with open(out_json, "w") as f:
            f.write(generate_json(mapped, stem, info["original_width"], info["original_height"]))      # :206-207
        written += [overlay, out_json]
    return {"written": written, "contours": contours}
