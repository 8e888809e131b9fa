"""ctypes binding of oracle/_ref/libpano_ref.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The library is the reference itself (chensh236/ComputerVisionImageStich2, root variant) compiled by oracle/Makefile from
/root/reference, behind the harness in oracle/ref_harness.cpp.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libpano_ref.so")
REF_DATA = os.path.join(HERE, "_ref", "data")

KEY_DTYPE = np.dtype(
    [("o", "<i4"), ("ix", "<i4"), ("iy", "<i4"), ("is", "<i4"), ("x", "<f4"), ("y", "<f4"), ("s", "<f4"), ("sigma", "<f4")]
)  # VlSiftKeypoint, vl/sift.h:19-31

_lib = None


def available() -> bool:
    return os.path.exists(REF_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(REF_SO)
        _lib.ref_sift_dump_run.restype = C.c_void_p
        _lib.ref_stitch_mem.restype = C.c_void_p
        _lib.ref_stitch_dir.restype = C.c_void_p
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def fnv1a64(buf: np.ndarray) -> str:
    """FNV-1a 64 over the planar bytes (SURVEY.md 8c hash convention)."""
    h = 0xCBF29CE484222325
    # vectorised would need big-int tricks; a python loop over ~2 MB is fine for tests but chunk it via int.from_bytes
    data = np.ascontiguousarray(buf).tobytes()
    mask = 0xFFFFFFFFFFFFFFFF
    for b in data:
        h = ((h ^ b) * 0x100000001B3) & mask
    return f"{h:016x}"


def load_bmp(path: str) -> np.ndarray:
    w, h = C.c_int(), C.c_int()
    if lib().ref_load_bmp(path.encode(), C.byref(w), C.byref(h), None) != 0:
        raise RuntimeError(path)
    out = np.empty((3, h.value, w.value), np.uint8)
    lib().ref_load_bmp(path.encode(), C.byref(w), C.byref(h), _p(out))
    return out


def project(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    _, h, w = img.shape
    out = np.empty_like(img)
    lib().ref_project(_p(img), w, h, _p(out))
    return out


def gray(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    _, h, w = img.shape
    out = np.empty((h, w), np.uint8)
    lib().ref_gray(_p(img), w, h, _p(out))
    return out


def sift_features(gray_u8: np.ndarray, cap: int = 1 << 20):
    g = np.ascontiguousarray(gray_u8, np.uint8)
    h, w = g.shape
    cap = max(4096, min(cap, (w * h) // 64))   # one run of the reference in the common case (~1 feature per 400 pixels)
    while True:
        descr = np.empty((cap, 128), np.float32)
        keys = np.empty(cap, KEY_DTYPE)
        n = lib().ref_sift_features(_p(g), w, h, _p(descr), _p(keys), cap)
        if n <= cap:
            return descr[:n].copy(), keys[:n].copy()
        cap = n


def sift_dump(im_f32: np.ndarray, noctaves=4, nlevels=2, o_min=0):
    """Per-octave raw VLFeat state: list of dicts with gss, dog, grad, keys, nangles, angles, descr, descr_key."""
    im = np.ascontiguousarray(im_f32, np.float32)
    h, w = im.shape
    L = lib()
    D = C.c_void_p(L.ref_sift_dump_run(_p(im), w, h, noctaves, nlevels, o_min))
    out = []
    nl = nlevels + 3
    for o in range(L.ref_sift_dump_noctaves(D)):
        ow, oh, nk, nd, hg = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.ref_sift_dump_info(D, o, C.byref(ow), C.byref(oh), C.byref(nk), C.byref(nd), C.byref(hg))
        ow, oh, nk, nd = ow.value, oh.value, nk.value, nd.value
        d = dict(w=ow, h=oh)
        d["gss"] = np.empty((nl, oh, ow), np.float32)
        d["dog"] = np.empty((nl - 1, oh, ow), np.float32)
        d["keys"] = np.empty(nk, KEY_DTYPE)
        d["nangles"] = np.empty(nk, np.int32)
        d["angles"] = np.empty((nk, 4), np.float64)
        d["descr"] = np.empty((nd, 128), np.float32)
        d["descr_key"] = np.empty(nd, np.int32)
        d["descr_written"] = np.empty(nd, np.int32)
        for what, name in ((0, "gss"), (1, "dog"), (3, "keys"), (4, "nangles"), (5, "angles"), (6, "descr"),
                           (7, "descr_key"), (8, "descr_written")):
            if d[name].size:
                L.ref_sift_dump_copy(D, o, what, _p(d[name]))
        if hg.value:
            d["grad"] = np.empty((nlevels, oh, ow, 2), np.float32)
            L.ref_sift_dump_copy(D, o, 2, _p(d["grad"]))
        else:
            d["grad"] = None
        out.append(d)
    L.ref_sift_dump_free(D)
    return out


def match(descA, keysA, descB, keysB):
    descA = np.ascontiguousarray(descA, np.float32)
    descB = np.ascontiguousarray(descB, np.float32)
    keysA = np.ascontiguousarray(keysA, KEY_DTYPE)
    keysB = np.ascontiguousarray(keysB, KEY_DTYPE)
    cap = len(keysB)
    oa, ob = np.empty(cap, KEY_DTYPE), np.empty(cap, KEY_DTYPE)
    n = lib().ref_match(_p(descA), _p(keysA), len(keysA), _p(descB), _p(keysB), len(keysB), _p(oa), _p(ob), cap)
    return oa[:n].copy(), ob[:n].copy()


def kdforest_query(data, queries, k=2, metric=0, per_query=True):
    """VLFeat kd-forest as getImgPair drives it (ImageProcess.cpp:280-327): (idx [nq][k] int64, dist [nq][k] f64).
    metric: 0 = VlDistanceL1, 1 = VlDistanceL2."""
    data = np.ascontiguousarray(data, np.float32)
    queries = np.ascontiguousarray(queries, np.float32)
    n, dim = data.shape
    nq = len(queries)
    idx = np.empty((nq, k), np.int64)
    dist = np.empty((nq, k), np.float64)
    if lib().ref_kdforest_query(_p(data), n, dim, metric, _p(queries), nq, k, int(per_query), _p(idx), _p(dist)) != 0:
        raise RuntimeError("ref_kdforest_query")
    return idx, dist


def ransac(src, dst) -> np.ndarray:
    src = np.ascontiguousarray(src, KEY_DTYPE)
    dst = np.ascontiguousarray(dst, KEY_DTYPE)
    H = np.empty(8, np.float64)
    lib().ref_ransac(_p(src), _p(dst), len(src), _p(H))
    return H


def fit4(src, dst) -> np.ndarray:
    src = np.ascontiguousarray(src, KEY_DTYPE)
    dst = np.ascontiguousarray(dst, KEY_DTYPE)
    H = np.empty(8, np.float64)
    lib().ref_fit4(_p(src), _p(dst), _p(H))
    return H


def inliers(src, dst, H8) -> np.ndarray:
    src = np.ascontiguousarray(src, KEY_DTYPE)
    dst = np.ascontiguousarray(dst, KEY_DTYPE)
    H8 = np.ascontiguousarray(H8, np.float64)
    idx = np.empty(len(src), np.int32)
    n = lib().ref_inliers(_p(src), _p(dst), len(src), _p(H8), _p(idx))
    return idx[:n].copy()


def refit(src, dst, idx) -> np.ndarray:
    src = np.ascontiguousarray(src, KEY_DTYPE)
    dst = np.ascontiguousarray(dst, KEY_DTYPE)
    idx = np.ascontiguousarray(idx, np.int32)
    H = np.empty(8, np.float64)
    lib().ref_refit(_p(src), _p(dst), len(src), _p(idx), len(idx), _p(H))
    return H


def warp_bounds(w, h, H8) -> np.ndarray:
    H8 = np.ascontiguousarray(H8, np.float64)
    out = np.empty(4, np.float32)
    lib().ref_warp_bounds(w, h, _p(H8), _p(out))
    return out


def warp(src, H8, offx, offy, cw, ch) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    _, h, w = src.shape
    H8 = np.ascontiguousarray(H8, np.float64)
    out = np.empty((3, ch, cw), np.uint8)
    lib().ref_warp(_p(src), w, h, _p(H8), C.c_float(offx), C.c_float(offy), cw, ch, _p(out))
    return out


def shift(src, offx, offy, cw, ch) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    _, h, w = src.shape
    out = np.empty((3, ch, cw), np.uint8)
    lib().ref_shift(_p(src), w, h, int(offx), int(offy), cw, ch, _p(out))
    return out


def blend(a, b) -> np.ndarray:
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    _, h, w = a.shape
    out = np.empty_like(a)
    if lib().ref_blend(_p(a), _p(b), w, h, _p(out)) != 0:
        raise RuntimeError("ref_blend: unexpected output shape")
    return out


def cimg_blur2(planes: np.ndarray) -> np.ndarray:
    p = np.ascontiguousarray(planes, np.float32)
    c, h, w = p.shape
    out = np.empty_like(p)
    lib().ref_cimg_blur2(_p(p), w, h, c, _p(out))
    return out


def cimg_resize3(planes: np.ndarray, nw: int, nh: int) -> np.ndarray:
    p = np.ascontiguousarray(planes, np.float32)
    c, h, w = p.shape
    out = np.empty((c, nh, nw), np.float32)
    if lib().ref_cimg_resize3(_p(p), w, h, c, nw, nh, _p(out)) != 0:
        raise RuntimeError("ref_cimg_resize3")
    return out


def equalize(img) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    _, h, w = img.shape
    out = np.empty_like(img)
    lib().ref_equalize(_p(img), w, h, _p(out))
    return out


def equalize_mix(img) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    _, h, w = img.shape
    out = np.empty_like(img)
    lib().ref_equalize_mix(_p(img), w, h, _p(out))
    return out


def _pano_out(P):
    L = lib()
    w, h, tf, tm = C.c_int(), C.c_int(), C.c_double(), C.c_double()
    L.ref_pano_info(P, C.byref(w), C.byref(h), C.byref(tf), C.byref(tm))
    out = np.empty((3, h.value, w.value), np.uint8)
    L.ref_pano_copy(P, _p(out))
    buf = C.create_string_buffer(4096)
    L.ref_pano_log(P, buf, 4096)
    nfeat = []
    i = 0
    while True:
        v = L.ref_pano_nfeat(P, i)
        if v < 0:
            break
        nfeat.append(v)
        i += 1
    L.ref_pano_free(P)
    return out, dict(t_features=tf.value, t_matching=tm.value, nfeat=nfeat, log=buf.value.decode())


def stitch_mem(imgs):
    """imgs: list of planar uint8 [3][H][W] arrays -> (panorama [3][H][W], info)."""
    imgs = [np.ascontiguousarray(i, np.uint8) for i in imgs]
    n = len(imgs)
    ptrs = (C.c_void_p * n)(*[i.ctypes.data for i in imgs])
    ws = (C.c_int * n)(*[i.shape[2] for i in imgs])
    hs = (C.c_int * n)(*[i.shape[1] for i in imgs])
    P = C.c_void_p(lib().ref_stitch_mem(ptrs, ws, hs, n))
    return _pano_out(P)


def stitch_dir(d: str, n: int):
    if not d.endswith("/"):
        d += "/"
    P = C.c_void_p(lib().ref_stitch_dir(d.encode(), n))
    return _pano_out(P)
