"""ctypes binding of oracle/_ref/libpano_ref_ex6.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The library is the reference's `src/ex6` variant (second caller of the hot path) compiled by
oracle/Makefile behind oracle/ref_ex6_harness.cpp, with the RANSAC seed pinned (ref_ex6_shim.h).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .ref_api import KEY_DTYPE, REF_DATA, _p  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
EX6_SO = os.path.join(HERE, "_ref", "libpano_ref_ex6.so")
DEFAULT_SEED = 666666

_lib = None


def available() -> bool:
    return os.path.exists(EX6_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(EX6_SO)
        _lib.ex6_stitch_mem.restype = C.c_void_p
        _lib.ex6_stitch_dir.restype = C.c_void_p
        _lib.ex6_set_seed(DEFAULT_SEED)
    return _lib


def set_seed(seed: int) -> None:
    """the value time(0) yields inside RANSAC (src/ex6/ImageProcess.cpp:403)"""
    lib().ex6_set_seed(C.c_uint(seed))


def dataset_dir(k: int) -> str:
    """dataset1 is byte-identical to the root variant's Input/ (staged once)."""
    return os.path.join(REF_DATA, "Input" if k == 1 else f"ex6_dataset{k}")


DATASET_SIZES = {1: 4, 2: 18, 3: 11}


def blend(a, b) -> np.ndarray:
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    _, h, w = a.shape
    out = np.empty_like(a)
    if lib().ex6_blend(_p(a), _p(b), w, h, _p(out)) != 0:
        raise RuntimeError("ex6_blend")
    return out


def cimg_blur2(planes: np.ndarray) -> np.ndarray:
    p = np.ascontiguousarray(planes, np.float32)
    c, h, w = p.shape
    out = np.empty_like(p)
    lib().ex6_cimg_blur2(_p(p), w, h, c, _p(out))
    return out


def ransac(src, dst) -> np.ndarray:
    src = np.ascontiguousarray(src, KEY_DTYPE)
    dst = np.ascontiguousarray(dst, KEY_DTYPE)
    H = np.empty(8, np.float64)
    lib().ex6_ransac(_p(src), _p(dst), len(src), _p(H))
    return H


def tail(img) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    _, h, w = img.shape
    out = np.empty_like(img)
    lib().ex6_tail(_p(img), w, h, _p(out))
    return out


def _pano_out(P):
    L = lib()
    w, h, tf, tm = C.c_int(), C.c_int(), C.c_double(), C.c_double()
    L.ex6_pano_info(P, C.byref(w), C.byref(h), C.byref(tf), C.byref(tm))
    out = np.empty((3, h.value, w.value), np.uint8)
    L.ex6_pano_copy(P, _p(out))
    buf = C.create_string_buffer(1 << 16)
    L.ex6_pano_log(P, buf, 1 << 16)
    nfeat = []
    i = 0
    while True:
        v = L.ex6_pano_nfeat(P, i)
        if v < 0:
            break
        nfeat.append(v)
        i += 1
    L.ex6_pano_free(P)
    return out, dict(t_features=tf.value, t_matching=tm.value, nfeat=nfeat, log=buf.value.decode())


def stitch_mem(imgs):
    """imgs: list of planar uint8 [3][H][W] arrays -> (panorama [3][H][W], info)."""
    imgs = [np.ascontiguousarray(i, np.uint8) for i in imgs]
    n = len(imgs)
    ptrs = (C.c_void_p * n)(*[i.ctypes.data for i in imgs])
    ws = (C.c_int * n)(*[i.shape[2] for i in imgs])
    hs = (C.c_int * n)(*[i.shape[1] for i in imgs])
    P = lib().ex6_stitch_mem(ptrs, ws, hs, n)
    if not P:
        raise RuntimeError("ex6: projected width > height (the reference exits, src/ex6/ImageProcess.cpp:35-38)")
    return _pano_out(C.c_void_p(P))
