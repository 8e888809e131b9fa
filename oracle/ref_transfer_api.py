"""ctypes binding of oracle/_ref/libpano_ref_transfer.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's Reinhard colour transfer (class transfer, transfer.cpp) compiled from /root/reference with its Win32
thread calls executed serially (oracle/ref_transfer_shim.h)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libpano_ref_transfer.so")
_lib = None


def available() -> bool:
    return os.path.exists(SO)


def color_transfer(src: np.ndarray, tem: np.ndarray) -> np.ndarray:
    """transfer(src, tem, out): src / tem planar uint8 [3][H][W] (sizes may differ) -> out like src"""
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO)
    s = np.ascontiguousarray(src, np.uint8)
    t = np.ascontiguousarray(tem, np.uint8)
    out = np.empty_like(s)
    rc = _lib.ref_color_transfer(s.ctypes.data_as(C.c_void_p), s.shape[2], s.shape[1], t.ctypes.data_as(C.c_void_p),
                                 t.shape[2], t.shape[1], out.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError(f"ref_color_transfer failed ({rc})")
    return out
