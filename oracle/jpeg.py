"""ctypes binding of oracle/jpeg_oracle.c (sample-domain half of libjpeg-turbo's default decoder). TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes

import numpy as np

from . import hotpath


def reconstruct(coefs: np.ndarray, w: int, h: int, hs, vs, quant: np.ndarray) -> np.ndarray:
    """coefs [nblocks,64] int16 (decode order, natural order inside a block), quant [ncomp,64] u16 -> RGB [h,w,3] u8."""
    L = hotpath.lib()
    L.orc_jpeg_reconstruct.restype = ctypes.c_int
    ncomp = len(hs)
    coefs = np.ascontiguousarray(coefs, np.int16)
    quant = np.ascontiguousarray(quant, np.uint16)
    hs_a, vs_a = (ctypes.c_uint32 * 3)(*hs, *([1] * (3 - ncomp))), (ctypes.c_uint32 * 3)(*vs, *([1] * (3 - ncomp)))
    out = np.empty((h, w, 3), np.uint8)
    rc = L.orc_jpeg_reconstruct(coefs.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(w), ctypes.c_uint32(h), ctypes.c_uint32(ncomp),
                                hs_a, vs_a, quant.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise MemoryError("orc_jpeg_reconstruct")
    return out
