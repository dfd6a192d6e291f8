"""ORACLE (test infrastructure only) -- ctypes binding of oracle/_ref/libref_pipeline.so: the REFERENCE'S OWN object code
(/root/reference/src/{preprocess,postprocess,mask2polygon}.cpp compiled unmodified against the OpenCV stub, recipe
oracle/ref_build/Makefile).  `available()` is False where the library was not built (it is built in the container that
has /root/reference and travels to the GPU box as a prebuilt file).

Only tests/, tests/golden/make_ref_golden.py and bench.py's CPU legs may import this module; the product never does."""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_pipeline.so")
_lib = None


def available() -> bool:
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(SO)
        P, I, L = C.c_void_p, C.c_int, C.c_int64
        l.ref_preprocess_raw.restype, l.ref_preprocess_raw.argtypes = I, [C.c_char_p, C.c_char_p, C.c_char_p, I, I]
        l.ref_read_png_gray.restype, l.ref_read_png_gray.argtypes = I, [C.c_char_p, P, L, C.POINTER(I), C.POINTER(I)]
        l.ref_read_png_bgr.restype, l.ref_read_png_bgr.argtypes = I, [C.c_char_p, P, L, C.POINTER(I), C.POINTER(I)]
        l.ref_write_png_gray.restype, l.ref_write_png_gray.argtypes = I, [C.c_char_p, P, I, I]
        l.ref_postprocess_mask.restype, l.ref_postprocess_mask.argtypes = I, [P, P, I, I]
        l.ref_extract_contours.restype, l.ref_extract_contours.argtypes = I, [P, I, I, P, L, P, I, C.POINTER(L)]
        l.ref_map_contour_points.restype, l.ref_map_contour_points.argtypes = None, [P, P, I, C.c_double, C.c_double]
        l.ref_generate_json.restype, l.ref_generate_json.argtypes = I, [P, P, I, C.c_char_p, C.c_char_p, I, I]
        l.ref_process_single_mask.restype, l.ref_process_single_mask.argtypes = None, [C.c_char_p] * 5
        _lib = l
    return _lib


def read_png_gray(path: str) -> np.ndarray:
    w, h = C.c_int(0), C.c_int(0)
    buf = np.empty(1 << 24, np.uint8)
    if not lib().ref_read_png_gray(path.encode(), buf.ctypes.data, buf.size, C.byref(w), C.byref(h)):
        raise IOError("stub imread failed: " + path)
    return buf[:w.value * h.value].reshape(h.value, w.value).copy()


def read_png_bgr(path: str) -> np.ndarray:
    w, h = C.c_int(0), C.c_int(0)
    buf = np.empty(3 << 24, np.uint8)
    if not lib().ref_read_png_bgr(path.encode(), buf.ctypes.data, buf.size, C.byref(w), C.byref(h)):
        raise IOError("stub imread failed: " + path)
    return buf[:w.value * h.value * 3].reshape(h.value, w.value, 3).copy()


def write_png_gray(path: str, img: np.ndarray) -> None:
    img = np.ascontiguousarray(img, np.uint8)
    if not lib().ref_write_png_gray(path.encode(), img.ctypes.data, img.shape[1], img.shape[0]):
        raise IOError("stub imwrite failed: " + path)


def preprocess_raw(src_u16: np.ndarray, filename: str = "slice.raw"):
    """Preprocess::preprocess_raw on a temporary RAW file -> (u8 512 x 512 pixels, sidecar JSON text)."""
    h, w = src_u16.shape
    with tempfile.TemporaryDirectory() as td:
        raw = os.path.join(td, filename)
        np.ascontiguousarray(src_u16, np.uint16).tofile(raw)
        png, js = os.path.join(td, "o", "n.png"), os.path.join(td, "sizes.json")
        if not lib().ref_preprocess_raw(raw.encode(), png.encode(), js.encode(), w, h):
            raise RuntimeError("reference preprocess_raw returned false")
        with open(js) as f:
            return read_png_gray(png), f.read()


def postprocess_mask(mask: np.ndarray) -> np.ndarray:
    m = np.ascontiguousarray(mask, np.uint8)
    out = np.empty_like(m)
    if not lib().ref_postprocess_mask(m.ctypes.data, out.ctypes.data, m.shape[0], m.shape[1]):
        raise RuntimeError("reference postprocess_mask threw")
    return out


def _csr(contours):
    cs = np.zeros(len(contours) + 1, np.int32)
    for i, c in enumerate(contours):
        cs[i + 1] = cs[i] + len(c)
    xy = np.ascontiguousarray(np.concatenate([np.asarray(c, np.int32).reshape(-1, 2) for c in contours]) if len(contours)
                              else np.zeros((0, 2), np.int32))
    return xy, cs


def extract_contours(mask_img: np.ndarray):
    m = np.ascontiguousarray(mask_img, np.uint8)
    H, W = m.shape
    cap, capc = 4 * H * W + 16, H * W + 1
    xy, cs, n = np.zeros((cap, 2), np.int32), np.zeros(capc + 1, np.int32), C.c_int64(0)
    nc = lib().ref_extract_contours(m.ctypes.data, H, W, xy.ctypes.data, cap, cs.ctypes.data, capc, C.byref(n))
    return [xy[cs[i]:cs[i + 1]].copy() for i in range(nc)]


def map_contour_points(contours, scale_x: float, scale_y: float):
    xy, cs = _csr(contours)
    lib().ref_map_contour_points(xy.ctypes.data, cs.ctypes.data, len(contours), float(scale_x), float(scale_y))
    return [xy[cs[i]:cs[i + 1]].copy() for i in range(len(contours))]


def generate_json(contours, base_name: str, ow: int, oh: int) -> str:
    xy, cs = _csr(contours)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "o.json")
        if not lib().ref_generate_json(xy.ctypes.data, cs.ctypes.data, len(contours), p.encode(), base_name.encode(), ow, oh):
            raise RuntimeError("reference generate_json threw")
        with open(p) as f:
            return f.read()


def process_single_mask(mask_png: str, output_dir: str, sizes_json: str, original_png: str, base_name: str) -> None:
    lib().ref_process_single_mask(mask_png.encode(), output_dir.encode(), sizes_json.encode(), original_png.encode(), base_name.encode())
