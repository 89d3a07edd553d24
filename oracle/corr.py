"""Correlation oracle: ctypes loader for corr_oracle.c plus a float64 numpy restatement.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import ctypes as C

import numpy as np

from . import build_c

_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_c.build())
        _lib.corr_oracle.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6
        _lib.corr_oracle.restype = C.c_int
    return _lib


def correlation_c(first, second, max_disp=20, stride2=2):
    """Reference summation order, fp32 (models/unflow_correlation.py:37-105)."""
    first = np.ascontiguousarray(first, dtype=np.float32)
    second = np.ascontiguousarray(second, dtype=np.float32)
    B, Cc, H, W = first.shape
    D = 2 * (max_disp // stride2) + 1
    out = np.empty((B, D * D, H, W), dtype=np.float32)
    rc = _load().corr_oracle(first.ctypes.data, second.ctypes.data, out.ctypes.data, B, Cc, H, W, max_disp, stride2)
    if rc != 0:
        raise MemoryError("corr_oracle failed")
    return out


def correlation_f64(first, second, max_disp=20, stride2=2):
    """Order-free float64 value of the same definition (for tolerance checks)."""
    first = np.asarray(first, dtype=np.float64)
    second = np.asarray(second, dtype=np.float64)
    B, Cc, H, W = first.shape
    R = max_disp // stride2
    D = 2 * R + 1
    pad = np.zeros((B, Cc, H + 2 * max_disp, W + 2 * max_disp))
    pad[:, :, max_disp:max_disp + H, max_disp:max_disp + W] = second
    out = np.empty((B, D * D, H, W))
    for iy in range(D):
        for ix in range(D):
            dy = (iy - R) * stride2 + max_disp
            dx = (ix - R) * stride2 + max_disp
            out[:, iy * D + ix] = (first * pad[:, :, dy:dy + H, dx:dx + W]).sum(1) / Cc
    return out
