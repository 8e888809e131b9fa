"""ctypes binding of tests/emul/libpano_emul.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

libpano_emul.so is a g++ build (-ffp-contract=off) of the per-item arithmetic bodies that the CUDA kernels execute
(csrc/*_device.cuh, __host__ __device__) driven by plain serial loops.  The CPU test tier uses it to check, without a
GPU, that the product's arithmetic reproduces the reference bit for bit; the GPU tier then checks the real kernels.
The product library never loads this file.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL_DIR = os.path.join(HERE, "emul")
EMUL_SO = os.path.join(EMUL_DIR, "libpano_emul.so")

KEY_DTYPE = np.dtype(
    [("o", "<i4"), ("ix", "<i4"), ("iy", "<i4"), ("is", "<i4"), ("x", "<f4"), ("y", "<f4"), ("s", "<f4"), ("sigma", "<f4")]
)
PAIR_DTYPE = np.dtype([("src", KEY_DTYPE), ("dst", KEY_DTYPE)])

_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", EMUL_DIR], check=True, stdout=subprocess.DEVNULL)
        _lib = C.CDLL(EMUL_SO)
        _lib.emul_sift_run.restype = C.c_void_p
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def project(img):
    img = np.ascontiguousarray(img, np.uint8)
    _, h, w = img.shape
    out = np.empty_like(img)
    lib().emul_project(_p(img), w, h, _p(out))
    return out


def gray(img):
    img = np.ascontiguousarray(img, np.uint8)
    _, h, w = img.shape
    out = np.empty((h, w), np.uint8)
    lib().emul_gray(_p(img), w, h, _p(out))
    return out


def sift_dump(im_f32, noctaves=4, nlevels=2):
    """Same structure as oracle.ref_api.sift_dump (no dog)."""
    im = np.ascontiguousarray(im_f32, np.float32)
    h, w = im.shape
    L = lib()
    D = C.c_void_p(L.emul_sift_run(_p(im), w, h, noctaves, nlevels))
    out = []
    nl = nlevels + 3
    for o in range(L.emul_sift_noctaves(D)):
        ow, oh, pitch, nk, nd = (C.c_int() for _ in range(5))
        L.emul_sift_info(D, o, C.byref(ow), C.byref(oh), C.byref(pitch), C.byref(nk), C.byref(nd))
        ow, oh, pitch, nk, nd = ow.value, oh.value, pitch.value, nk.value, nd.value
        d = dict(w=ow, h=oh)
        gss = np.empty((nl, oh, pitch), np.float32)
        grad = np.empty((nlevels, oh, pitch, 2), np.float32)
        d["keys"] = np.empty(nk, KEY_DTYPE)
        d["nangles"] = np.empty(nk, np.int32)
        d["angles"] = np.empty((nk, 4), np.float64)
        d["descr"] = np.empty((nd, 128), np.float32)
        d["descr_key"] = np.empty(nd, np.int32)
        d["descr_written"] = np.empty(nd, np.int32)
        if oh * pitch:
            L.emul_sift_copy(D, o, 0, _p(gss))
            L.emul_sift_copy(D, o, 2, _p(grad))
        for what, name in ((3, "keys"), (4, "nangles"), (5, "angles"), (6, "descr"), (7, "descr_key"), (8, "descr_written")):
            if d[name].size:
                L.emul_sift_copy(D, o, what, _p(d[name]))
        d["gss"] = np.ascontiguousarray(gss[:, :, :ow])
        d["grad"] = np.ascontiguousarray(grad[:, :, :ow, :])
        out.append(d)
    L.emul_sift_free(D)
    return out


def serial_mismatches():
    return lib().emul_serial_mismatches()


def feature_table(dump):
    """std::map<vector<float>, VlSiftKeypoint> insertion semantics over a sift_dump (ImageProcess.cpp:57, 80-86)."""
    rows = []
    for o in dump:
        wr = o["descr_written"].astype(bool)
        for j in np.nonzero(wr)[0]:
            rows.append((o["descr"][j], o["keys"][o["descr_key"][j]]))
    if not rows:
        return np.empty((0, 128), np.float32), np.empty(0, KEY_DTYPE)
    d = np.stack([r[0] for r in rows])
    k = np.array([r[1] for r in rows], KEY_DTYPE)
    order = np.lexsort(d.T[::-1], axis=0)  # stable, first column most significant
    d, k = d[order], k[order]
    keep = np.ones(len(d), bool)
    keep[1:] = np.any(d[1:] != d[:-1], axis=1)
    d, k = d[keep], k[keep].copy()
    k["ix"] = k["x"].astype(np.int32)
    k["iy"] = k["y"].astype(np.int32)
    return np.ascontiguousarray(d), k


def match_idx(dA, dB):
    dA = np.ascontiguousarray(dA, np.float32)
    dB = np.ascontiguousarray(dB, np.float32)
    idx = np.empty(len(dB), np.int32)
    lib().emul_match(_p(dA), len(dA), _p(dB), len(dB), _p(idx))
    return idx


def match_idx_prefilter(dA, dB, cand_cap=0):
    """The pre-filter pipeline (quantise, SAD statistics, decision, candidates, exact re-rank) on the CPU.
    Returns (idx, stats) with stats = survivors, overflow queries, largest candidate list, unbounded rows."""
    dA = np.ascontiguousarray(dA, np.float32)
    dB = np.ascontiguousarray(dB, np.float32)
    idx = np.empty(len(dB), np.int32)
    st = np.zeros(4, np.int64)
    lib().emul_match_prefilter(_p(dA), len(dA), _p(dB), len(dB), _p(idx), _p(st), int(cand_cap))
    return idx, {"survivors": int(st[0]), "overflow": int(st[1]), "max_candidates": int(st[2]), "unbounded_rows": int(st[3])}


def match_pair_grouped(dX, dY, sample_rows=0, cand_cap=0):
    """The grouped pass (seed, grouped bound, exact SAD of the pairs it cannot skip, decision, candidates, exact re-rank)
    for both directions of one image pair on the CPU.  Returns (idx of getImgPair(X, Y), idx of getImgPair(Y, X), stats)."""
    dX = np.ascontiguousarray(dX, np.float32)
    dY = np.ascontiguousarray(dY, np.float32)
    ixy = np.empty(len(dY), np.int32)
    iyx = np.empty(len(dX), np.int32)
    st = np.zeros(4, np.int64)
    lib().emul_match_grouped(_p(dX), len(dX), _p(dY), len(dY), _p(ixy), _p(iyx), _p(st), int(sample_rows), int(cand_cap))
    return ixy, iyx, {"survivors": int(st[0]), "exact": int(st[1]), "pairs": int(st[2]), "unqualified": int(st[3]) & 1,
                      "accepts": int(st[3]) >> 1}


def _pairs(src, dst):
    p = np.empty(len(src), PAIR_DTYPE)
    p["src"] = src
    p["dst"] = dst
    return p


def ransac(src, dst, seed=None):
    p = _pairs(src, dst)
    H = np.empty(8, np.float64)
    if seed is None:
        rc = lib().emul_ransac(_p(p), len(p), _p(H))
    else:
        rc = lib().emul_ransac_seeded(_p(p), len(p), C.c_uint(seed), _p(H))
    if rc != 0:
        raise RuntimeError("emul_ransac failed")
    return H


def draw_samples(npairs, seed, use_libc=False):
    idx = np.empty((72, 4), np.int32)
    n = lib().emul_draw_samples(npairs, C.c_uint(seed), _p(idx), int(use_libc))
    assert n == 72
    return idx


def fit4(src, dst):
    p = _pairs(src, dst)
    H = np.empty(8, np.float64)
    lib().emul_fit4(_p(p), _p(H))
    return H


def refit(src, dst, idx):
    p = _pairs(src, dst)
    idx = np.ascontiguousarray(idx, np.int32)
    H = np.empty(8, np.float64)
    if lib().emul_refit(_p(p), _p(idx), len(idx), _p(H)) != 0:
        raise RuntimeError("emul_refit failed")
    return H


def plan_canvas(dw, dh, H8, rw, rh, ex6=False):
    H8 = np.ascontiguousarray(H8, np.float64)
    mm = np.empty(4, np.float32)
    wh = np.empty(2, np.int32)
    (lib().emul_plan_canvas_ex6 if ex6 else lib().emul_plan_canvas)(dw, dh, _p(H8), rw, rh, _p(mm), _p(wh))
    return mm, wh


def warp(src, H8, offx, offy, cw, ch):
    src = np.ascontiguousarray(src, np.uint8)
    H8 = np.ascontiguousarray(H8, np.float64)
    _, h, w = src.shape
    out = np.empty((3, ch, cw), np.uint8)
    lib().emul_warp(_p(src), w, h, _p(H8), C.c_float(offx), C.c_float(offy), cw, ch, _p(out))
    return out


def cimg_blur2(p, deriche=False):
    p = np.ascontiguousarray(p, np.float32)
    c, h, w = p.shape
    out = np.empty_like(p)
    (lib().emul_cimg_blur2_deriche if deriche else lib().emul_cimg_blur2)(_p(p), w, h, c, _p(out))
    return out


def cimg_resize3(p, nw, nh):
    p = np.ascontiguousarray(p, np.float32)
    c, h, w = p.shape
    out = np.empty((c, nh, nw), np.float32)
    lib().emul_cimg_resize3(_p(p), w, h, c, nw, nh, _p(out))
    return out


def blend(a, b, ex6=False):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    _, h, w = a.shape
    out = np.empty_like(a)
    rc = (lib().emul_blend_ex6 if ex6 else lib().emul_blend)(_p(a), _p(b), w, h, _p(out))
    if rc != 0:
        raise RuntimeError("emul_blend: empty middle row")
    return out


def equalize_mix(img, ex6=False):
    img = np.ascontiguousarray(img, np.uint8)
    _, h, w = img.shape
    out = np.empty_like(img)
    (lib().emul_equalize_mix_ex6 if ex6 else lib().emul_equalize_mix)(_p(img), w, h, _p(out))
    return out


def color_transfer(src, tem):
    s = np.ascontiguousarray(src, np.uint8)
    t = np.ascontiguousarray(tem, np.uint8)
    out = np.empty_like(s)
    lib().emul_color_transfer(_p(s), s.shape[2], s.shape[1], _p(t), t.shape[2], t.shape[1], _p(out))
    return out
