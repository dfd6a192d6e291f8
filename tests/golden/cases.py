"""Seeded input generators shared by make_golden.py and the tests (numpy only)."""
import numpy as np


def contour_case_masks():
    rng = np.random.default_rng(20261018)
    masks = []

    def add(m):
        masks.append((np.asarray(m, bool).astype(np.uint8) * 255))

    add(np.zeros((5, 7)))                       # empty
    add(np.ones((6, 4)))                        # full
    add(np.ones((1, 1)))                        # single pixel image
    add(np.ones((1, 9)))                        # 1 x N
    add(np.ones((9, 1)))                        # N x 1
    m = np.zeros((7, 7)); m[3, 1:6] = 1; add(m)             # horizontal 5-px line
    m = np.zeros((7, 7)); m[1:6, 3] = 1; m[3, 1:6] = 1; add(m)  # plus sign (1-px strokes)
    m = np.zeros((9, 9)); m[np.arange(9), np.arange(9)] = 1; add(m)  # diagonal
    m = np.zeros((8, 8)); m[2:6, 2:6] = 1; add(m)           # rectangle
    m = np.zeros((12, 12)); m[1:11, 1:11] = 1; m[3:9, 3:9] = 0; m[5:7, 5:7] = 1; add(m)  # ring + island
    yy, xx = np.mgrid[0:16, 0:16]; add((yy + xx) % 2 == 0)   # checkerboard (corner touching)
    m = np.ones((10, 10)); m[0, :] = 0; add(m)              # touching 3 edges
    for _ in range(12):
        h, w = rng.integers(3, 48, 2)
        add(rng.random((h, w)) < rng.uniform(0.1, 0.9))
    for _ in range(6):
        h, w = rng.integers(20, 64, 2)
        f = rng.random((h, w))
        for _ in range(3):
            f = (f + np.roll(f, 1, 0) + np.roll(f, -1, 0) + np.roll(f, 1, 1) + np.roll(f, -1, 1)) / 5
        add(f > np.median(f))
    # nested rings, depth 4
    m = np.zeros((40, 40))
    for k in range(0, 18, 2):
        m[k:40 - k, k:40 - k] = (k // 2) % 2 == 0
    add(m)
    return masks


def postprocess_case_masks():
    rng = np.random.default_rng(777)
    out = []
    for (h, w) in [(64, 64), (96, 128), (128, 128), (100, 77)]:
        for _ in range(3):
            m = np.zeros((h, w), np.uint8)
            yy, xx = np.mgrid[0:h, 0:w]
            # big foreground blob with holes, class-1 islands, speckle
            cy, cx = h / 2 + rng.normal(0, 4), w / 2 + rng.normal(0, 4)
            m[((yy - cy) / (0.4 * h)) ** 2 + ((xx - cx) / (0.42 * w)) ** 2 < 1] = 2
            for _ in range(4):
                hy, hx, r = rng.integers(5, h - 5), rng.integers(5, w - 5), rng.integers(1, max(2, min(h, w) // 6))
                m[(yy - hy) ** 2 + (xx - hx) ** 2 < r * r] = rng.integers(0, 2)
            sp = rng.random((h, w))
            m[sp < 0.03] = 0
            m[sp > 0.97] = 2
            m[(sp > 0.5) & (sp < 0.51)] = 1
            out.append(m)
    out.append(np.full((32, 32), 2, np.uint8))
    out.append(np.zeros((32, 32), np.uint8))
    m = np.zeros((64, 64), np.uint8); m[:2, :] = 2; m[30:32, 10:50] = 2; out.append(m)   # 2-px strips: border vs interior
    return out


def preprocess_cases():
    rng = np.random.default_rng(4242)
    cases = {}
    for (w, h) in [(512, 512), (1024, 768), (333, 517), (640, 480), (64, 48)]:
        cases[f"rand_{w}x{h}"] = rng.integers(0, 65536, (h, w), dtype=np.uint16)
    cases["narrow_600x400"] = rng.integers(900, 3200, (400, 600)).astype(np.uint16)
    cases["const_128x128"] = np.full((128, 128), 1234, np.uint16)
    cases["max_100x100"] = np.full((100, 100), 65535, np.uint16)
    g = (np.add.outer(np.arange(700), np.arange(900)) % 4096).astype(np.uint16)
    cases["ramp_900x700"] = g
    return cases


def json_cases():
    return {
        "one": ("slice_000", 512, 512, [[(10, 20), (10, 300), (400, 300), (400, 20)]]),
        "two": ("vol-1.slice_2", 1024, 768, [[(0, 0)], [(5, 6), (7, 8), (9, 10)]]),
        "none": ("empty", 64, 64, []),
    }


# scale factors (orig / scaled) for map_contour_points: identity, up, down, non-representable ratios
MAP_SCALES = [(1.0, 1.0), (600 / 512, 400 / 512), (2048 / 512, 1536 / 512), (333 / 512, 517 / 512), (1000 / 48, 7 / 3)]


def ref_mask_cases():
    """512 x 512 class masks {0,1,2}: a body with holes of both background classes, border-touching holes, small blobs
    on either side of the 6 % area threshold (15,728 px), 1-2 px bridges that the 3 x 3 open cuts."""
    rng = np.random.default_rng(31337)
    out = []
    yy, xx = np.mgrid[0:512, 0:512]
    for k in range(4):
        m = np.zeros((512, 512), np.uint8)
        cy, cx = 256 + rng.normal(0, 12), 256 + rng.normal(0, 12)
        m[((yy - cy) / 170) ** 2 + ((xx - cx) / 215) ** 2 < 1] = 2
        for _ in range(10):                                     # holes of class 0 / 1 inside the body
            hy, hx, r = rng.integers(120, 390), rng.integers(100, 410), rng.integers(2, 70)
            m[(yy - hy) ** 2 + (xx - hx) ** 2 < r * r] = rng.integers(0, 2)
        for _ in range(6):                                      # blobs around the area threshold: r ~ 71 px <-> 15.7 k px
            by, bx, r = rng.integers(30, 480), rng.integers(30, 480), rng.integers(3, 80)
            m[(yy - by) ** 2 + (xx - bx) ** 2 < r * r] = 2
        m[200 + 10 * k:202 + 10 * k, :] = 2                      # a 2-px bar across the image (erased by the open, cut at the body)
        sp = rng.random((512, 512))
        m[sp < 0.01] = 0
        m[sp > 0.99] = 2
        m[(sp > 0.50) & (sp < 0.505)] = 1
        if k == 3:
            m[:, :3] = 2                                        # foreground hugging the left edge
            m[0:40, 100:140] = 0
        out.append(m)
    return out
