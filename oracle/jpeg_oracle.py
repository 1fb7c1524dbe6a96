"""oracle/jpeg_oracle.py -- TEST INFRASTRUCTURE ONLY: ctypes wrapper of oracle/jpeg_ref.c.

The CPU restatement of `cv::imread` for baseline JPEG (reference: samples/main.cpp:24-40), pinned bit-exact
against cv2.imdecode in tests/test_oracle_jpeg.py.  Used only by tests/, smoke() and bench.py's CPU legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "build", "libjpeg_ref.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "build/libjpeg_ref.so"])
        _LIB = C.CDLL(path)
        _LIB.jpeg_ref_info.argtypes = [C.c_void_p, C.c_size_t] + [C.POINTER(C.c_int)] * 6
        _LIB.jpeg_ref_num_blocks.argtypes = [C.c_void_p, C.c_size_t]
        _LIB.jpeg_ref_num_blocks.restype = C.c_long
        _LIB.jpeg_ref_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    return _LIB


def info(data: bytes) -> dict:
    v = [C.c_int() for _ in range(6)]
    buf = np.frombuffer(data, np.uint8)
    st = lib().jpeg_ref_info(buf.ctypes.data, buf.size, *[C.byref(x) for x in v])
    if st:
        raise ValueError(f"jpeg oracle: error {st}")
    return dict(zip(("width", "height", "components", "h_samp", "v_samp", "restart_interval"), (x.value for x in v)))


def decode(data: bytes, want_coefficients: bool = False):
    """-> BGR uint8 [H, W, 3] (and the quantised coefficient blocks [n, 64] in scan order, natural order)."""
    buf = np.frombuffer(data, np.uint8)
    meta = info(data)
    out = np.empty((meta["height"], meta["width"], 3), np.uint8)
    coef = None
    if want_coefficients:
        coef = np.empty((lib().jpeg_ref_num_blocks(buf.ctypes.data, buf.size), 64), np.int16)
    st = lib().jpeg_ref_decode(buf.ctypes.data, buf.size, out.ctypes.data, coef.ctypes.data if coef is not None else None)
    if st:
        raise ValueError(f"jpeg oracle: error {st}")
    return (out, coef) if want_coefficients else out
