"""medseg_b200 -- host-side mirror of the reference's stage API over libmedseg_b200.so.

The reference is compiled C++ (its host side lives in csrc/facade.cpp with the reference's own
namespaces); this Python module is the same interface for tests and bench.py, bound with ctypes to
the C ABI declared in include/medseg_b200.h.  Stage names follow the reference:

    initialize  -> MedicalSeg::initialize_engine      (/root/reference/src/initialize.cpp:26)
    preprocess  -> Preprocess::preprocess_raw         (src/preprocess.cpp:76)
    process     -> MedicalSeg::execute_inference      (src/process.cpp:123)   [UNet + head]
    postprocess -> postprocess_mask                   (src/postprocess.cpp:47)
    mask2polygon-> Mask2Polygon::extract_contours + map_contour_points (src/mask2polygon.cpp:29,41)
    cleanup     -> MedicalSeg::cleanup_resources      (src/cleanup.cpp:10)

There is no CPU fallback: if the shared library is missing, or no sm_100 GPU is present, every call
raises.  (The directory name contains '-', so import it through the `medseg_b200` shim at the
repository root.)
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmedseg_b200.so")
CSRC = os.path.join(_HERE, "csrc")

MS_OK, MS_ERR_ARG, MS_ERR_IO, MS_ERR_FORMAT, MS_ERR_CUDA, MS_ERR_CAPACITY, MS_ERR_STATE, MS_ERR_INTERNAL = 0, -1, -2, -3, -4, -5, -6, -7


class MedsegError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class ms_info(C.Structure):
    _fields_ = [("device", C.c_int32), ("sm_count", C.c_int32), ("net_h", C.c_int32), ("net_w", C.c_int32),
                ("n_classes", C.c_int32), ("max_batch", C.c_int32), ("foreground_value", C.c_int32),
                ("min_area_ratio", C.c_float), ("has_weights", C.c_int32), ("n_params", C.c_int64),
                ("flops_per_slice", C.c_int64)]


class ms_polygons(C.Structure):
    _fields_ = [("xy", C.c_void_p), ("cap_points", C.c_int64), ("contour_start", C.c_void_p), ("cap_contours", C.c_int64),
                ("slice_start", C.c_void_p), ("n_points", C.c_int64), ("n_contours", C.c_int64)]


# every symbol include/medseg_b200.h declares: name -> (restype, argtypes)
_P, _I, _L = C.c_void_p, C.c_int, C.c_int64
ABI = {
    "ms_init": (_I, [C.c_char_p, C.c_char_p, C.POINTER(_P)]),
    "ms_init_json": (_I, [C.c_char_p, C.c_char_p, C.POINTER(_P)]),
    "ms_destroy": (None, [_P]),
    "ms_last_error": (C.c_char_p, [_P]),
    "ms_get_info": (_I, [_P, C.POINTER(ms_info)]),
    "ms_alloc_pinned": (_P, [C.c_size_t]),
    "ms_free_pinned": (None, [_P]),
    "ms_preprocess_dev": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "ms_preprocess_host": (_I, [_P, _P, _I, _I, _I, _P]),
    "ms_unet_forward_dev": (_I, [_P, _P, _I, _P, _P, _P]),
    "ms_unet_forward_host": (_I, [_P, _P, _I, _P, _P]),
    "ms_postprocess_dev": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "ms_postprocess_host": (_I, [_P, _P, _P, _I, _I, _I, _I]),
    "ms_mask2polygon_host": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, C.POINTER(ms_polygons)]),
    "ms_mask2polygon_dev": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, C.POINTER(ms_polygons), _P]),
    "ms_process_batch_host": (_I, [_P, _P, _I, _I, _I, C.POINTER(ms_polygons), _P, _P]),
    "ms_process_batch_dev": (_I, [_P, _P, _I, _I, _I, C.POINTER(_L), C.POINTER(_L), _P]),
    "ms_last_counts": (_I, [_P, C.POINTER(_L), C.POINTER(_L)]),
    "ms_get_transfer_bytes": (_I, [_P, C.POINTER(_L), C.POINTER(_L)]),
    "ms_submit_batch_host": (_I, [_P, _I, _P, _I, _I, _I]),
    "ms_wait_batch": (_I, [_P, _I, C.POINTER(ms_polygons)]),
    "ms_process_volume_host": (_I, [_P, _P, _I, _I, _L, C.POINTER(ms_polygons), _P, _P]),
    "ms_device_count": (_I, [_P]),
    "ms_process_raw_file": (_I, [_P, C.c_char_p, _I, _I, C.c_char_p]),
    "ms_process_raw_files": (_I, [_P, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), _L, _I, _I, _P, C.POINTER(_L), C.POINTER(_L)]),
    "ms_process_directory": (_I, [_P, C.c_char_p, _I, _I, C.c_char_p, _I, _I, _I, C.POINTER(_L), C.POINTER(_L), C.POINTER(_L)]),
    "ms_polygons_to_json": (_L, [_P, _P, _I, C.c_char_p, _I, _I, _P, _L]),
    "ms_polygons_to_json_batch": (_L, [_P, _P, _P, _I, C.POINTER(C.c_char_p), _I, _I, _I, _P, _L, _P]),
    "ms_process_batch_multiclass_host": (_I, [_P, _P, _I, _I, _I, _P, _I, _P, _P, _P]),
    "ms_launch_count": (_L, [_P]),
    "ms_set_dp_epsilon": (_I, [_P, C.c_double]),
    "ms_dp_epsilon": (C.c_double, [_P]),
    "ms_time_layer": (_I, [_P, _I, _I, _I, C.POINTER(C.c_float), C.POINTER(C.c_double)]),
    "ms_layer_count": (_I, [_P]),
    "ms_profile_layers_begin": (_I, [_P, _I]),
    "ms_profile_layers_read": (_I, [_P, C.POINTER(C.c_float), _I, C.POINTER(_I)]),
    "ms_profile_stages_read": (_I, [_P, C.POINTER(C.c_float), C.POINTER(_I)]),
    "ms_layer_name": (C.c_char_p, [_P, _I]),
    "ms_layer_kernel": (C.c_char_p, [_P, _I]),
    "ms_debug_fused_phases": (_I, [_P]),
    "ms_debug_read_activation": (_L, [_P, C.c_char_p, _I, _P, _L]),
}

_lib_cache = None


def build(verbose: bool = False) -> str:
    """Compile libmedseg_b200.so in-tree (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libmedseg_b200.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded shared library with typed entry points.  Fails loudly when it is missing."""
    global _lib_cache
    if _lib_cache is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in ABI.items():
            fn = getattr(l, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib_cache = l
    return _lib_cache


def _ptr(a) -> int:
    if a is None:
        return 0
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)


def _u8(a, name) -> np.ndarray:
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint8:
        raise TypeError(f"{name} must be uint8")
    return a


class Polygons:
    """CSR polygon set as returned by the C ABI, plus list-of-arrays views."""

    def __init__(self, xy: np.ndarray, contour_start: np.ndarray, slice_start: np.ndarray):
        self.xy, self.contour_start, self.slice_start = xy, contour_start, slice_start

    @property
    def n_contours(self) -> int:
        return int(self.slice_start[-1])

    @property
    def n_points(self) -> int:
        return int(self.contour_start[self.n_contours])

    def slice(self, s: int) -> List[np.ndarray]:
        a, b = int(self.slice_start[s]), int(self.slice_start[s + 1])
        return [self.xy[self.contour_start[c]:self.contour_start[c + 1]] for c in range(a, b)]

    def per_slice(self) -> List[List[np.ndarray]]:
        return [self.slice(s) for s in range(len(self.slice_start) - 1)]


class Engine:
    """One handle = one GPU.  `source` is a config dict, a path to a config JSON or weight blob, or
    None for a stage-only handle (no UNet)."""

    def __init__(self, source=None, log_dir: Optional[str] = None):
        self._l = lib()
        self._h = _P()
        ld = log_dir.encode() if log_dir else None
        if isinstance(source, dict):
            rc = self._l.ms_init_json(json.dumps(source).encode(), ld, C.byref(self._h))
        else:
            rc = self._l.ms_init(source.encode() if source else None, ld, C.byref(self._h))
        if rc != MS_OK:
            self._h = _P()
            raise MedsegError(rc, self._l.ms_last_error(None).decode(errors="replace"))
        inf = ms_info()
        self._l.ms_get_info(self._h, C.byref(inf))
        self.info = inf
        self._cap_pts, self._cap_cnt = 1 << 16, 1 << 10

    # reference-style names
    initialize = classmethod(lambda cls, engine_path, log_dir=None: cls(engine_path, log_dir))

    def _check(self, rc: int):
        if rc != MS_OK:
            raise MedsegError(rc, self._l.ms_last_error(self._h).decode(errors="replace"))

    def cleanup(self):
        if self._h:
            self._l.ms_destroy(self._h)
            self._h = _P()

    close = cleanup

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.cleanup()

    def __del__(self):
        try:
            self.cleanup()
        except Exception:
            pass

    # ------------------------------------------------------------------ stages (host buffers)
    def preprocess(self, src_u16: np.ndarray) -> np.ndarray:
        src = np.ascontiguousarray(src_u16)
        if src.dtype != np.uint16:
            raise TypeError("src must be uint16")
        if src.ndim == 2:
            src = src[None]
        b, h, w = src.shape
        out = np.empty((b, self.info.net_h, self.info.net_w), np.uint8)
        self._check(self._l.ms_preprocess_host(self._h, _ptr(src), w, h, b, _ptr(out)))
        return out

    def process(self, norm_u8: np.ndarray, want_logits: bool = False):
        """UNet forward + head: u8 [B,H,W] -> class mask u8 [B,H,W] (and fp32 logits [B,C,H,W])."""
        x = _u8(norm_u8, "norm")
        if x.ndim == 2:
            x = x[None]
        b = x.shape[0]
        mask = np.empty_like(x)
        logits = np.empty((b, self.info.n_classes, x.shape[1], x.shape[2]), np.float32) if want_logits else None
        self._check(self._l.ms_unet_forward_host(self._h, _ptr(x), b, _ptr(mask), _ptr(logits)))
        return (mask, logits) if want_logits else mask

    def postprocess(self, mask: np.ndarray, fg_value: int = 0) -> np.ndarray:
        m = _u8(mask, "mask")
        squeeze = m.ndim == 2
        if squeeze:
            m = m[None]
        out = np.empty_like(m)
        self._check(self._l.ms_postprocess_host(self._h, _ptr(m), _ptr(out), m.shape[1], m.shape[2], m.shape[0], fg_value))
        return out[0] if squeeze else out

    def _poly_call(self, call, batch: int, retry: bool = True) -> Polygons:
        if not retry:   # one-shot calls (the result is consumed by the call): size the buffers generously up front
            self._cap_pts, self._cap_cnt = max(self._cap_pts, batch * 8192), max(self._cap_cnt, batch * 64)
        while True:
            xy = np.empty((self._cap_pts, 2), np.int32)
            cs = np.empty(self._cap_cnt + 1, np.int32)
            ss = np.empty(batch + 1, np.int32)
            pg = ms_polygons(_ptr(xy), self._cap_pts, _ptr(cs), self._cap_cnt, _ptr(ss), 0, 0)
            rc = call(C.byref(pg))
            if rc == MS_ERR_CAPACITY and (pg.n_points > self._cap_pts or pg.n_contours > self._cap_cnt):
                self._cap_pts = max(self._cap_pts, int(pg.n_points) + 16)
                self._cap_cnt = max(self._cap_cnt, int(pg.n_contours) + 16)
                continue
            self._check(rc)
            return Polygons(xy[:pg.n_points], cs[:pg.n_contours + 1], ss)

    def mask2polygon(self, mask: np.ndarray, threshold: int = 127, orig_w: Optional[int] = None,
                     orig_h: Optional[int] = None) -> Polygons:
        m = _u8(mask, "mask")
        if m.ndim == 2:
            m = m[None]
        b, h, w = m.shape
        ow, oh = orig_w or w, orig_h or h
        return self._poly_call(lambda pg: self._l.ms_mask2polygon_host(self._h, _ptr(m), h, w, b, threshold, ow, oh, pg), b)

    def process_batch(self, src_u16: np.ndarray, want_norm: bool = False, want_mask: bool = False):
        """RAW u16 slices [B,h,w] -> polygons in original coordinates (the whole per-slice path)."""
        src = src_u16 if isinstance(src_u16, np.ndarray) and src_u16.flags.c_contiguous else np.ascontiguousarray(src_u16)
        if src.dtype != np.uint16:
            raise TypeError("src must be uint16")
        if src.ndim == 2:
            src = src[None]
        b, h, w = src.shape
        norm = np.empty((b, self.info.net_h, self.info.net_w), np.uint8) if want_norm else None
        mask = np.empty((b, self.info.net_h, self.info.net_w), np.uint8) if want_mask else None
        polys = self._poly_call(
            lambda pg: self._l.ms_process_batch_host(self._h, _ptr(src), w, h, b, pg, _ptr(norm), _ptr(mask)), b)
        return polys, norm, mask

    def process_volume(self, vol_u16: np.ndarray, want_norm: bool = False, want_mask: bool = False):
        """cfg3: one volume [N,h,w] through one call, sharded by contiguous slice blocks over the handle's GPUs
        ("devices" in the config).  Returns (Polygons over all N slices, norm, mask)."""
        v = vol_u16
        if v.dtype != np.uint16 or not v.flags.c_contiguous or v.ndim != 3:
            raise TypeError("volume must be a C-contiguous uint16 [N,h,w] array")
        n, h, w = v.shape
        norm = np.empty((n, self.info.net_h, self.info.net_w), np.uint8) if want_norm else None
        mask = np.empty((n, self.info.net_h, self.info.net_w), np.uint8) if want_mask else None
        self._cap_pts, self._cap_cnt = max(self._cap_pts, n * 1024), max(self._cap_cnt, n * 8)
        polys = self._poly_call(lambda pg: self._l.ms_process_volume_host(self._h, _ptr(v), w, h, n, pg, _ptr(norm), _ptr(mask)), n)
        return polys, norm, mask

    def device_count(self) -> int:
        return int(self._l.ms_device_count(self._h))

    def submit_batch(self, slot: int, src_u16: np.ndarray) -> None:
        """Asynchronous half of the double-buffered pipeline: enqueue H2D + the whole path for `src_u16` [B,h,w]
        (keep the array alive until wait_batch(slot) returns)."""
        if src_u16.dtype != np.uint16 or not src_u16.flags.c_contiguous or src_u16.ndim != 3:
            raise TypeError("src must be a C-contiguous uint16 [B,h,w] array")
        b, h, w = src_u16.shape
        self._check(self._l.ms_submit_batch_host(self._h, slot, _ptr(src_u16), w, h, b))
        self._slot_batch = getattr(self, "_slot_batch", {})
        self._slot_batch[slot] = b

    def wait_batch(self, slot: int) -> Polygons:
        return self._poly_call(lambda pg: self._l.ms_wait_batch(self._h, slot, pg), self._slot_batch[slot], retry=False)

    def process_multiclass(self, src_u16: np.ndarray, classes: Sequence[int]):
        """cfg4 extension: per-class contours.  preprocess -> UNet argmax -> for each class k: postprocess with
        FOREGROUND_VALUE = k (src/postprocess.cpp:5 made a parameter) -> contours of (mask == k), mapped to
        the original size.  Returns (raw_mask [B,H,W], {k: (clean_mask, Polygons)})."""
        src = np.ascontiguousarray(src_u16)
        if src.ndim == 2:
            src = src[None]
        b, h, w = src.shape
        raw = self.process(self.preprocess(src))
        out = {}
        for k in classes:
            clean = self.postprocess(raw, fg_value=int(k))
            out[int(k)] = (clean, self.mask2polygon(clean, threshold=int(k) - 1, orig_w=w, orig_h=h))
        return raw, out

    def process_batch_multiclass(self, src_u16: np.ndarray, classes: Sequence[int], want_masks: bool = False):
        """cfg4 through ONE C-ABI call (ms_process_batch_multiclass_host): K1 + UNet once, then K5 (FOREGROUND_VALUE = k) + K6 per
        label on the device.  Returns {k: Polygons} or, with want_masks, (raw_mask, {k: (clean_mask, Polygons)})."""
        src = np.ascontiguousarray(src_u16)
        if src.ndim == 2:
            src = src[None]
        if src.dtype != np.uint16:
            raise TypeError("src must be uint16")
        b, h, w = src.shape
        cls = np.asarray(list(classes), np.int32)
        n = len(cls)
        nh, nw = self.info.net_h, self.info.net_w
        raw = np.empty((b, nh, nw), np.uint8) if want_masks else None
        clean = np.empty((n, b, nh, nw), np.uint8) if want_masks else None
        while True:
            xy = [np.empty((self._cap_pts, 2), np.int32) for _ in range(n)]
            cs = [np.empty(self._cap_cnt + 1, np.int32) for _ in range(n)]
            ss = [np.empty(b + 1, np.int32) for _ in range(n)]
            pgs = (ms_polygons * n)(*[ms_polygons(_ptr(xy[i]), self._cap_pts, _ptr(cs[i]), self._cap_cnt, _ptr(ss[i]), 0, 0) for i in range(n)])
            rc = self._l.ms_process_batch_multiclass_host(self._h, _ptr(src), w, h, b, _ptr(cls), n, C.cast(pgs, C.c_void_p), _ptr(raw), _ptr(clean))
            need_p, need_c = max(int(p.n_points) for p in pgs), max(int(p.n_contours) for p in pgs)
            if rc == MS_ERR_CAPACITY and (need_p > self._cap_pts or need_c > self._cap_cnt):
                self._cap_pts, self._cap_cnt = max(self._cap_pts, need_p + 16), max(self._cap_cnt, need_c + 16)
                continue
            self._check(rc)
            break
        polys = {int(k): Polygons(xy[i][:pgs[i].n_points], cs[i][:pgs[i].n_contours + 1], ss[i]) for i, k in enumerate(cls)}
        if want_masks:
            return raw, {int(k): (clean[i], polys[int(k)]) for i, k in enumerate(cls)}
        return polys

    def process_raw_file(self, raw_path: str, w: int, h: int, out_dir: str) -> None:
        self._check(self._l.ms_process_raw_file(self._h, raw_path.encode(), w, h, out_dir.encode()))

    def process_raw_files(self, raw_paths, w: int, h: int, out_dirs) -> np.ndarray:
        """Batched per-file loop (src/main.cpp:148-164).  `out_dirs` is one directory or one per file.
        Returns a bool array: file i produced its artefacts."""
        raw_paths = [os.fspath(p) for p in raw_paths]
        if isinstance(out_dirs, (str, os.PathLike)):
            out_dirs = [out_dirs] * len(raw_paths)
        out_dirs = [os.fspath(p) for p in out_dirs]
        if len(out_dirs) != len(raw_paths):
            raise ValueError("one output directory per file")
        n = len(raw_paths)
        a = (C.c_char_p * max(n, 1))(*[p.encode() for p in raw_paths])
        b = (C.c_char_p * max(n, 1))(*[p.encode() for p in out_dirs])
        ok = np.zeros(max(n, 1), np.uint8)
        n_ok, n_bad = _L(0), _L(0)
        self._check(self._l.ms_process_raw_files(self._h, a, b, n, w, h, ok.ctypes.data, C.byref(n_ok), C.byref(n_bad)))
        assert n_ok.value + n_bad.value == n
        return ok[:n].astype(bool)

    def process_directory(self, input_dir: str, w: int, h: int, out_dir: str, recursive: bool = False, shard_index: int = 0,
                          shard_count: int = 1):
        """Directory branch of src/main.cpp:134-168.  Returns (n_found, n_ok, n_failed); with shard_count > 1 the
        sorted file list is split into contiguous blocks (sharding.shard_range) and only this shard's is processed."""
        found, good, bad = _L(0), _L(0), _L(0)
        self._check(self._l.ms_process_directory(self._h, os.fspath(input_dir).encode(), w, h, os.fspath(out_dir).encode(),
                                                 1 if recursive else 0, shard_index, shard_count, C.byref(found), C.byref(good),
                                                 C.byref(bad)))
        return found.value, good.value, bad.value

    # ------------------------------------------------------------------ device-pointer entry points
    def preprocess_dev(self, d_src: int, w: int, h: int, batch: int, d_out_u8: int, d_out_bf16: int = 0, stream: int = 0):
        self._check(self._l.ms_preprocess_dev(self._h, d_src, w, h, batch, d_out_u8, d_out_bf16 or None, stream or None))

    def unet_forward_dev(self, d_in: int, batch: int, d_mask: int, d_logits: int = 0, stream: int = 0):
        self._check(self._l.ms_unet_forward_dev(self._h, d_in, batch, d_mask, d_logits or None, stream or None))

    def postprocess_dev(self, d_in: int, d_out: int, h: int, w: int, batch: int, fg_value: int = 0, stream: int = 0):
        self._check(self._l.ms_postprocess_dev(self._h, d_in, d_out, h, w, batch, fg_value, stream or None))

    def process_batch_dev(self, d_src: int, w: int, h: int, batch: int, stream: int = 0, wait: bool = True):
        """Whole path on device-resident input.  wait=False: nothing is read back and the host does not block
        (collect the totals later with last_counts())."""
        if not wait:
            self._check(self._l.ms_process_batch_dev(self._h, d_src, w, h, batch, None, None, stream or None))
            return None
        npts, ncnt = _L(0), _L(0)
        self._check(self._l.ms_process_batch_dev(self._h, d_src, w, h, batch, C.byref(npts), C.byref(ncnt), stream or None))
        return npts.value, ncnt.value

    def last_counts(self):
        """(n_points, n_contours) of the last process_batch_dev(wait=False); waits for it."""
        npts, ncnt = _L(0), _L(0)
        self._check(self._l.ms_last_counts(self._h, C.byref(npts), C.byref(ncnt)))
        return npts.value, ncnt.value

    def transfer_bytes(self):
        """(host->device, device->host) bytes this handle has copied since creation."""
        a, b = _L(0), _L(0)
        self._check(self._l.ms_get_transfer_bytes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def mask2polygon_dev(self, d_mask: int, h: int, w: int, batch: int, threshold: int = 127, stream: int = 0) -> Polygons:
        return self._poly_call(
            lambda pg: self._l.ms_mask2polygon_dev(self._h, d_mask, h, w, batch, threshold, w, h, pg, stream or None), batch)

    def set_dp_epsilon(self, eps: float) -> None:
        """Opt-in Douglas-Peucker (cv2.approxPolyDP, closed) on every contour before the coordinate mapping; 0 = off = the
        reference's CHAIN_APPROX_SIMPLE output."""
        self._check(self._l.ms_set_dp_epsilon(self._h, float(eps)))

    def dp_epsilon(self) -> float:
        return float(self._l.ms_dp_epsilon(self._h))

    # ------------------------------------------------------------------ instrumentation
    def launch_count(self) -> int:
        return int(self._l.ms_launch_count(self._h))

    def layer_names(self) -> List[str]:
        return [self._l.ms_layer_name(self._h, i).decode() for i in range(self._l.ms_layer_count(self._h))]

    def profile_layers_begin(self, max_forwards: int) -> None:
        """Bracket every layer launch of the next `max_forwards` eager forward passes with CUDA events."""
        self._check(self._l.ms_profile_layers_begin(self._h, max_forwards))

    def profile_layers_read(self):
        """-> (average ms per layer, number of passes recorded); switches the timing off."""
        n = self._l.ms_layer_count(self._h)
        buf = (C.c_float * n)()
        passes = _I(0)
        self._check(self._l.ms_profile_layers_read(self._h, buf, n, C.byref(passes)))
        return [float(v) for v in buf], passes.value

    def profile_stages_read(self):
        """-> ({"K1": ms, "unet": ms, "K5": ms, "K6": ms}, passes) recorded since profile_layers_begin."""
        buf = (C.c_float * 4)()
        passes = _I(0)
        self._check(self._l.ms_profile_stages_read(self._h, buf, C.byref(passes)))
        return dict(zip(("K1", "unet", "K5", "K6"), (float(v) for v in buf))), passes.value

    def layer_kernels(self) -> List[str]:
        """Kernel instantiation each UNet layer runs on (as ncu prints it)."""
        return [self._l.ms_layer_kernel(self._h, i).decode() for i in range(self._l.ms_layer_count(self._h))]

    def time_layer(self, layer: int, batch: int, iters: int = 20):
        ms, fl = C.c_float(0), C.c_double(0)
        self._check(self._l.ms_time_layer(self._h, layer, batch, iters, C.byref(ms), C.byref(fl)))
        return ms.value, fl.value

    def read_activation(self, name: str, batch: int) -> np.ndarray:
        n = self._l.ms_debug_read_activation(self._h, name.encode(), batch, None, 0)
        if n < 0:
            self._check(int(n))
        out = np.empty(n, np.float32)
        n2 = self._l.ms_debug_read_activation(self._h, name.encode(), batch, _ptr(out), n)
        if n2 < 0:
            self._check(int(n2))
        return out


def polygons_to_json(contours: Sequence[np.ndarray], base_name: str, orig_w: int, orig_h: int) -> str:
    """Byte-exact text of Mask2Polygon::generate_json (src/mask2polygon.cpp:68-109).  Pure host code."""
    l = lib()
    cs = np.zeros(len(contours) + 1, np.int32)
    for i, c in enumerate(contours):
        cs[i + 1] = cs[i] + len(c)
    xy = np.ascontiguousarray(np.concatenate([np.asarray(c, np.int32).reshape(-1, 2) for c in contours])
                              if len(contours) else np.zeros((0, 2), np.int32))
    n = l.ms_polygons_to_json(_ptr(xy), _ptr(cs), len(contours), base_name.encode(), orig_w, orig_h, None, 0)
    if n < 0:
        raise MedsegError(int(n), "ms_polygons_to_json failed")
    buf = C.create_string_buffer(n)
    l.ms_polygons_to_json(_ptr(xy), _ptr(cs), len(contours), base_name.encode(), orig_w, orig_h, C.addressof(buf), n)
    return buf.raw.decode()


def polygons_to_json_batch(polys: "Polygons", base_names: Sequence[str], orig_w: int, orig_h: int, n_threads: int = 0,
                           buf: Optional[np.ndarray] = None):
    """LabelMe documents of every slice of a polygon set, formatted by a pool of host threads (ms_polygons_to_json_batch).
    Returns (uint8 buffer, offsets[n_slices + 1]); slice s is buf[offsets[s]:offsets[s + 1]] (empty when it has no contour)."""
    l = lib()
    n = len(polys.slice_start) - 1
    names = (C.c_char_p * max(n, 1))(*[b.encode() for b in base_names])
    offs = np.zeros(n + 1, np.int64)
    xy = np.ascontiguousarray(polys.xy, np.int32)
    cs = np.ascontiguousarray(polys.contour_start, np.int32)
    ss = np.ascontiguousarray(polys.slice_start, np.int32)
    if buf is None:
        buf = np.empty(max(1 << 16, int(polys.n_points) * 100 + n * 1024), np.uint8)
    while True:
        total = l.ms_polygons_to_json_batch(_ptr(xy), _ptr(cs), _ptr(ss), n, names, orig_w, orig_h, n_threads, _ptr(buf), buf.size, _ptr(offs))
        if total < 0:
            raise MedsegError(int(total), "ms_polygons_to_json_batch failed")
        if total <= buf.size:
            return buf[:total], offs
        buf = np.empty(int(total) + 1024, np.uint8)


def make_weight_blob(path: str, n_classes: int = 3, seed: int = 1234) -> str:
    from . import weights as W
    W.save_blob(path, W.make_weights(seed, n_classes), n_classes)
    return path
