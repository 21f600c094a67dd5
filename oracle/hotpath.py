"""ctypes binding for oracle/liboracle.so (hotpath_oracle.c). TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    """Compile liboracle.so with the committed Makefile (building the checker is not using it)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("hotpath_oracle.c", "jpeg_oracle.c", "Makefile")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        u8p = ctypes.POINTER(ctypes.c_uint8)
        f32p = ctypes.POINTER(ctypes.c_float)
        i32p = ctypes.POINTER(ctypes.c_int32)
        L.orc_resize_axis_taps.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, i32p, i32p, f32p]
        L.orc_resize_axis_taps.restype = ctypes.c_int
        L.orc_resize_triangle.argtypes = [u8p, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.orc_resize_triangle.restype = ctypes.c_int
        L.orc_normalise_nchw.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p]
        L.orc_normalise_nchw.restype = None
        L.orc_bbox_area.argtypes = [f32p]
        L.orc_bbox_area.restype = ctypes.c_float
        L.orc_iou.argtypes = [f32p, f32p]
        L.orc_iou.restype = ctypes.c_float
        L.orc_postproc.argtypes = [f32p, f32p, ctypes.c_int, ctypes.c_float, ctypes.c_float, f32p, i32p, ctypes.c_int]
        L.orc_postproc.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def axis_taps(src_len: int, dst_len: int):
    """(left[D], ntaps[D], w[D,T]) of one resize axis (image 0.24.5 sample.rs)."""
    L = lib()
    t = L.orc_resize_axis_taps(src_len, dst_len, 0, None, None, None)
    left = np.zeros(dst_len, np.int32)
    nt = np.zeros(dst_len, np.int32)
    w = np.zeros((dst_len, t), np.float32)
    L.orc_resize_axis_taps(src_len, dst_len, t, _p(left, ctypes.c_int32), _p(nt, ctypes.c_int32), _p(w, ctypes.c_float))
    return left, nt, w


def resize_triangle(rgb: np.ndarray, nw: int, nh: int, round_intermediate: bool = False) -> np.ndarray:
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, c = rgb.shape
    assert c == 3
    out = np.empty((nh, nw, 3), np.uint8)
    rc = lib().orc_resize_triangle(_p(rgb, ctypes.c_uint8), w, h, _p(out, ctypes.c_uint8), nw, nh, int(round_intermediate))
    if rc != 0:
        raise MemoryError("orc_resize_triangle")
    return out


def normalise_nchw(hwc: np.ndarray, preset: int = 0) -> np.ndarray:
    hwc = np.ascontiguousarray(hwc, np.uint8)
    h, w, _ = hwc.shape
    out = np.empty((3, h, w), np.float32)
    lib().orc_normalise_nchw(_p(hwc, ctypes.c_uint8), w, h, preset, _p(out, ctypes.c_float))
    return out


def bbox_area(b) -> float:
    b = np.ascontiguousarray(b, np.float32)
    return float(lib().orc_bbox_area(_p(b, ctypes.c_float)))


def iou(a, b) -> float:
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return float(lib().orc_iou(_p(a, ctypes.c_float), _p(b, ctypes.c_float)))


def postproc(scores: np.ndarray, boxes: np.ndarray, min_conf: float, max_iou: float):
    """scores [K,2], boxes [K,4] -> (dets [n,5] = x0,y0,x1,y1,conf in selection order, prior idx [n])."""
    scores = np.ascontiguousarray(scores, np.float32)
    boxes = np.ascontiguousarray(boxes, np.float32)
    K = scores.shape[0]
    out = np.empty((max(K, 1), 5), np.float32)
    idx = np.empty(max(K, 1), np.int32)
    n = lib().orc_postproc(_p(scores, ctypes.c_float), _p(boxes, ctypes.c_float), K, min_conf, max_iou,
                           _p(out, ctypes.c_float), _p(idx, ctypes.c_int32), K)
    if n < 0:
        raise MemoryError("orc_postproc")
    return out[:n].copy(), idx[:n].copy()
