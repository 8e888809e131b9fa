"""computervisionimagestich2_b200 -- B200-native (sm_100a) panorama-stitching hot path.

The product is the C-ABI shared library ``libpano_b200.so`` built from ``csrc/`` (hand-written CUDA kernels + a C++ host
orchestrator mirroring the reference's ``ImageProcess`` class).  This Python package is only a thin ctypes binding over
that ABI, used by the tests and by ``bench.py``; it contains no numerical code and has no CPU fallback: every call
fails loudly when the library is missing or no CUDA device is present.

Reference interface mirrored (chensh236/ComputerVisionImageStich2): ``ImageProcess(dir, n)`` -> ``Stitcher``,
``Projection::imageProjection`` -> ``project``, ``siftAlgorithm`` -> ``sift_features``, ``getImgPair`` -> ``match``,
``RANSAC`` -> ``ransac``, ``warpingImageByHomography`` / ``movingImageByOffset`` -> ``warp_shift``,
``blendTwoImages`` -> ``blend``, the equalisation tail of ``matching`` -> ``equalize_mix``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpano_b200.so")

KEY_DTYPE = np.dtype(
    [("o", "<i4"), ("ix", "<i4"), ("iy", "<i4"), ("is", "<i4"), ("x", "<f4"), ("y", "<f4"), ("s", "<f4"), ("sigma", "<f4")]
)  # VlSiftKeypoint, vl/sift.h:19-31
PAIR_DTYPE = np.dtype([("src", KEY_DTYPE), ("dst", KEY_DTYPE)])  # ImgPair, ImageProcess.h:43-47


class Times(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("project", "sift", "table", "match", "ransac", "warp", "blend", "tail", "total")] + [
        ("match_pairs_evaluated", C.c_int64), ("sift_pixels", C.c_int64), ("n_match_calls", C.c_int), ("n_blends", C.c_int)]


class PanoError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load libpano_b200.so; raises if it was not built (python -m computervisionimagestich2_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PanoError(f"{LIB_PATH} is missing: build it with `python -m computervisionimagestich2_b200.build` "
                            "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.pano_b200_last_error.restype = C.c_char_p
        L.pano_b200_last_error.argtypes = [C.c_void_p]
        L.pano_b200_free.argtypes = [C.c_void_p]
        L.pano_b200_destroy.argtypes = [C.c_void_p]
        L.pano_b200_alloc_pinned.restype = C.c_void_p
        L.pano_b200_alloc_pinned.argtypes = [C.c_size_t]
        L.pano_b200_free_pinned.argtypes = [C.c_void_p]
        L.pano_b200_ktimer_launches.restype = C.c_long
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    return np.ascontiguousarray(a, np.uint8)


class Context:
    """One CUDA device + stream + workspaces (pano_b200_ctx)."""

    def __init__(self, device: int = 0):
        L = lib()
        if L.pano_b200_device_count() <= 0:
            raise PanoError("no CUDA device visible: the B200 path cannot run (there is no CPU fallback)")
        h = C.c_void_p()
        rc = L.pano_b200_create(device, C.byref(h))
        if rc != 0 or not h.value:
            raise PanoError(f"pano_b200_create failed ({rc})")
        self.h = h
        self.L = L

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.L.pano_b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise PanoError(f"{what} failed ({rc}): {self.L.pano_b200_last_error(self.h).decode()}")

    PROFILES = {"root": 0, "ex6": 1}

    def set_profile(self, name: str = "root", ransac_seed: int = 666666):
        """Which caller of the hot path is reproduced: "root" (ImageProcess.cpp) or "ex6" (src/ex6/ImageProcess.cpp);
        ransac_seed = the srand() argument of ImageProcess::RANSAC (666666 in root, time(0) in ex6)."""
        self._check(self.L.pano_b200_set_profile(self.h, self.PROFILES[name], C.c_uint(ransac_seed)), "set_profile")
        self.profile = name

    # ---- stages ------------------------------------------------------------------------------------------------
    def project(self, img, want_gray=False):
        img = _u8(img)
        _, h, w = img.shape
        out = np.empty_like(img)
        g = np.empty((h, w), np.uint8) if want_gray else None
        self._check(self.L.pano_b200_project(self.h, _p(img), w, h, _p(out), _p(g) if want_gray else None), "project")
        return (out, g) if want_gray else out

    def gray(self, img):
        img = _u8(img)
        _, h, w = img.shape
        g = np.empty((h, w), np.uint8)
        self._check(self.L.pano_b200_gray(self.h, _p(img), w, h, _p(g)), "gray")
        return g

    def sift_features(self, gray_u8):
        g = _u8(gray_u8)
        h, w = g.shape
        pd, pk, n = C.c_void_p(), C.c_void_p(), C.c_int()
        self._check(self.L.pano_b200_sift_features(self.h, _p(g), w, h, C.byref(pd), C.byref(pk), C.byref(n)), "sift_features")
        n = n.value
        descr = np.ctypeslib.as_array(C.cast(pd, C.POINTER(C.c_float)), (max(n, 1), 128))[:n].copy()
        keys = np.frombuffer(C.string_at(pk, n * KEY_DTYPE.itemsize), KEY_DTYPE).copy()
        self.L.pano_b200_free(pd)
        self.L.pano_b200_free(pk)
        return descr, keys

    def sift_raw(self, im_f32, noctaves=4, nlevels=2):
        im = np.ascontiguousarray(im_f32, np.float32)
        h, w = im.shape
        pk, pa, pd, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int()
        nk = (C.c_int * max(noctaves, 64))()
        self._check(self.L.pano_b200_sift_raw(self.h, _p(im), w, h, noctaves, nlevels, C.byref(pk), C.byref(pa), C.byref(pd),
                                              C.byref(n), nk), "sift_raw")
        n = n.value
        keys = np.frombuffer(C.string_at(pk, n * KEY_DTYPE.itemsize), KEY_DTYPE).copy()
        angles = np.frombuffer(C.string_at(pa, n * 8), np.float64).copy()
        descr = np.frombuffer(C.string_at(pd, n * 512), np.float32).reshape(n, 128).copy()
        for q in (pk, pa, pd):
            self.L.pano_b200_free(q)
        return dict(keys=keys, angles=angles, descr=descr, octave_nkeys=[nk[i] for i in range(noctaves)])

    def sift_octave_dump(self, octave, nlevels=2):
        ow, oh = C.c_int(), C.c_int()
        self._check(self.L.pano_b200_sift_octave_dims(self.h, octave, C.byref(ow), C.byref(oh)), "octave_dims")
        gss = np.empty((nlevels + 3, oh.value, ow.value), np.float32)
        grad = np.empty((nlevels, oh.value, ow.value, 2), np.float32)
        self._check(self.L.pano_b200_sift_octave_dump(self.h, octave, _p(gss), _p(grad)), "octave_dump")
        return gss, grad

    def set_match_mode(self, mode: str = "prefilter"):
        """'prefilter' (default: uint8 SAD pre-filter + exact re-rank, image pairs share one SAD pass), 'full' (exact
        float scan of every pair) or 'prefilter_onedir' (pre-filter, one SAD pass per directed problem)."""
        self._check(self.L.pano_b200_set_match_mode(self.h, {"prefilter": 0, "full": 1, "prefilter_onedir": 2, "prefilter_fullsad": 3}[mode]), "set_match_mode")

    def match_stats(self, reset=False):
        out = (C.c_longlong * 9)()
        self._check(self.L.pano_b200_match_stats_ex(self.h, out, 9, int(reset)), "match_stats")
        return {"queries": out[0], "survivors": out[1], "overflow": out[2], "problems": out[3], "sym_pairs": out[4],
                "group_pairs": out[5], "group_exact": out[6], "group_accepts": out[7], "group_overflow": out[8]}

    def match_idx(self, descA, descB):
        dA = np.ascontiguousarray(descA, np.float32)
        dB = np.ascontiguousarray(descB, np.float32)
        idx = np.empty(len(dB), np.int32)
        n = C.c_int()
        self._check(self.L.pano_b200_match(self.h, _p(dA), len(dA), _p(dB), len(dB), _p(idx), C.byref(n)), "match")
        return idx

    def match_pair(self, descA, descB):
        """-> (getImgPair(A, B) indices [nB], getImgPair(B, A) indices [nA]) from one batch (one SAD pass by default)."""
        dA = np.ascontiguousarray(descA, np.float32)
        dB = np.ascontiguousarray(descB, np.float32)
        ab = np.empty(len(dB), np.int32)
        ba = np.empty(len(dA), np.int32)
        self._check(self.L.pano_b200_match_pair(self.h, _p(dA), len(dA), _p(dB), len(dB), _p(ab), _p(ba)), "match_pair")
        return ab, ba

    def match(self, descA, keysA, descB, keysB):
        """getImgPair: returns (A keypoints, B keypoints) of the matches in B's table order."""
        idx = self.match_idx(descA, descB)
        sel = idx >= 0
        return np.ascontiguousarray(keysA, KEY_DTYPE)[idx[sel]].copy(), np.ascontiguousarray(keysB, KEY_DTYPE)[sel].copy()

    def ransac(self, src, dst, debug=False):
        pairs = np.empty(len(src), PAIR_DTYPE)
        pairs["src"] = src
        pairs["dst"] = dst
        H = np.empty(8, np.float64)
        counts = np.empty(72, np.int32)
        hyps = np.empty((72, 8), np.float64)
        inl = np.empty(len(src), np.int32)
        ninl = C.c_int()
        self._check(self.L.pano_b200_ransac(self.h, _p(pairs), len(pairs), _p(H), _p(counts), _p(hyps), _p(inl), C.byref(ninl)), "ransac")
        if debug:
            return H, counts, hyps, inl[: ninl.value].copy()
        return H

    def plan_canvas(self, dw, dh, fwdH, rw, rh):
        H = np.ascontiguousarray(fwdH, np.float64)
        b = np.empty(4, np.float32)
        s = np.empty(2, np.int32)
        self.L.pano_b200_plan_canvas_ex(self.PROFILES[getattr(self, "profile", "root")], dw, dh, _p(H), rw, rh, _p(b),
                                        _p(s))
        return b, s

    def warp_shift(self, src, H8, offx, offy, prev, ioffx, ioffy, cw, ch):
        H = np.ascontiguousarray(H8, np.float64) if H8 is not None else np.zeros(8)
        a = b = None
        sw = sh = pw = ph = 0
        if src is not None:
            src = _u8(src)
            _, sh, sw = src.shape
            a = np.empty((3, ch, cw), np.uint8)
        if prev is not None:
            prev = _u8(prev)
            _, ph, pw = prev.shape
            b = np.empty((3, ch, cw), np.uint8)
        self._check(self.L.pano_b200_warp_shift(self.h, _p(src) if src is not None else None, sw, sh, _p(H), C.c_float(offx),
                                                C.c_float(offy), _p(prev) if prev is not None else None, pw, ph, int(ioffx),
                                                int(ioffy), cw, ch, _p(a) if a is not None else None,
                                                _p(b) if b is not None else None), "warp_shift")
        return a, b

    def blend(self, a, b):
        a, b = _u8(a), _u8(b)
        _, h, w = a.shape
        out = np.empty_like(a)
        self._check(self.L.pano_b200_blend(self.h, _p(a), _p(b), w, h, _p(out)), "blend")
        return out

    def equalize_mix(self, img):
        img = _u8(img)
        _, h, w = img.shape
        out = np.empty_like(img)
        self._check(self.L.pano_b200_equalize_mix(self.h, _p(img), w, h, _p(out)), "equalize_mix")
        return out

    def color_transfer(self, src, tem):
        """the reference's `transfer tran(src, tem, out)` (Reinhard l-alpha-beta transfer): -> out like src"""
        s, t = _u8(src), _u8(tem)
        out = np.empty_like(s)
        self._check(self.L.pano_b200_color_transfer(self.h, _p(s), s.shape[2], s.shape[1], _p(t), t.shape[2], t.shape[1], _p(out)),
                    "color_transfer")
        return out

    def cimg_blur2(self, planes, deriche=False):
        """get_blur(2, true, true) (Van Vliet; root variant) or get_blur(2) (Deriche; src/ex6)."""
        p = np.ascontiguousarray(planes, np.float32)
        c, h, w = p.shape
        out = np.empty_like(p)
        fn = self.L.pano_b200_cimg_blur2_deriche if deriche else self.L.pano_b200_cimg_blur2
        self._check(fn(self.h, _p(p), w, h, c, _p(out)), "cimg_blur2")
        return out

    def cimg_resize3(self, planes, nw, nh):
        p = np.ascontiguousarray(planes, np.float32)
        c, h, w = p.shape
        out = np.empty((c, nh, nw), np.float32)
        self._check(self.L.pano_b200_cimg_resize3(self.h, _p(p), w, h, c, nw, nh, _p(out)), "cimg_resize3")
        return out

    def stitch_bmp_files(self, paths):
        """ImageProcess(dir, n) on BMP files: decode / encode on the GPU -> bytes of the panorama BMP"""
        blobs = [open(p, "rb").read() for p in paths]
        n = len(blobs)
        bufs = [C.create_string_buffer(b, len(b)) for b in blobs]
        ptrs = (C.c_void_p * n)(*[C.addressof(b) for b in bufs])
        sizes = (C.c_size_t * n)(*[len(b) for b in blobs])
        out, osz = C.c_void_p(), C.c_size_t()
        self._check(self.L.pano_b200_stitch_bmp(self.h, ptrs, sizes, n, C.byref(out), C.byref(osz)), "stitch_bmp")
        data = C.string_at(out, osz.value)
        self.L.pano_b200_free(out)
        return data

    # ---- sharded jobs (dist.py) -------------------------------------------------------------------------------------
    def extract(self, img):
        """readFile body for one image: -> (projected [3][H][W] u8, descr [n][128] f32, keys [n])"""
        img = _u8(img)
        _, h, w = img.shape
        proj = np.empty_like(img)
        pd, pk, n = C.c_void_p(), C.c_void_p(), C.c_int()
        self._check(self.L.pano_b200_extract(self.h, _p(img), w, h, _p(proj), C.byref(pd), C.byref(pk), C.byref(n)), "extract")
        n = n.value
        descr = np.frombuffer(C.string_at(pd, n * 512), np.float32).reshape(n, 128).copy()
        keys = np.frombuffer(C.string_at(pk, n * KEY_DTYPE.itemsize), KEY_DTYPE).copy()
        self.L.pano_b200_free(pd)
        self.L.pano_b200_free(pk)
        return proj, descr, keys

    # ---- sharded job, device-resident exchange (dist.stitch_sharded_device).  Tensors are torch tensors on this
    #      context's device; the library copies device-to-device from / into them. -----------------------------------
    def shard_begin(self, n_global):
        self._check(self.L.pano_b200_shard_begin(self.h, int(n_global)), "shard_begin")

    def shard_extract(self, imgs, slots, staged_ptrs=None, sizes=None):
        """readFile of the images this rank owns.  imgs: planar uint8 host arrays, or staged_ptrs: device pointers of
        inputs already resident in HBM with sizes = [(w, h), ...]."""
        n = len(slots)
        if n == 0:
            return
        if staged_ptrs is None:
            imgs = [_u8(i) for i in imgs]
            ptrs = (C.c_void_p * n)(*[i.ctypes.data for i in imgs])
            ws = (C.c_int * n)(*[i.shape[2] for i in imgs])
            hs = (C.c_int * n)(*[i.shape[1] for i in imgs])
            self._keep = imgs
        else:
            ptrs = (C.c_void_p * n)(*staged_ptrs)
            ws = (C.c_int * n)(*[s[0] for s in sizes])
            hs = (C.c_int * n)(*[s[1] for s in sizes])
        sl = (C.c_int * n)(*[int(x) for x in slots])
        self._check(self.L.pano_b200_shard_extract(self.h, ptrs, ws, hs, sl, n, int(staged_ptrs is not None)), "shard_extract")

    def nfeatures(self, i):
        return int(self.L.pano_b200_stitch_nfeatures(self.h, int(i)))

    def shard_export(self, i, descr_t, keys, proj_t):
        self._check(self.L.pano_b200_shard_export(self.h, int(i), C.c_void_p(descr_t.data_ptr() if descr_t is not None else None),
                                                  _p(keys) if keys is not None else None,
                                                  C.c_void_p(proj_t.data_ptr() if proj_t is not None else None)), "shard_export")

    def shard_import(self, i, w, h, n, descr_t, keys, proj_t):
        keys = np.ascontiguousarray(keys, KEY_DTYPE)
        self._check(self.L.pano_b200_shard_import(self.h, int(i), int(w), int(h), int(n),
                                                  C.c_void_p(descr_t.data_ptr() if n > 0 else None), _p(keys) if n > 0 else None,
                                                  C.c_void_p(proj_t.data_ptr() if proj_t is not None else None)), "shard_import")

    def shard_match(self, I, J, out_t):
        n = len(I)
        if n == 0:
            return
        ia = (C.c_int * n)(*[int(x) for x in I])
        ja = (C.c_int * n)(*[int(x) for x in J])
        self._check(self.L.pano_b200_shard_match(self.h, ia, ja, n, C.c_void_p(out_t.data_ptr())), "shard_match")

    def shard_preset(self, i, j, idx):
        a = np.ascontiguousarray(idx, np.int32)
        self._check(self.L.pano_b200_shard_preset(self.h, int(i), int(j), _p(a), len(a)), "shard_preset")

    def shard_stitch(self, want_output=True, pinned_out=None, pinned_cap=0):
        """rank 0: the sequential part.  -> (panorama or None, info)"""
        ow, oh = C.c_int(), C.c_int()
        if pinned_out is not None:
            self._check(self.L.pano_b200_shard_stitch(self.h, C.c_void_p(pinned_out), C.c_size_t(pinned_cap), C.byref(ow), C.byref(oh)),
                        "shard_stitch")
            pano = None
        else:
            self._check(self.L.pano_b200_shard_stitch(self.h, None, C.c_size_t(0), C.byref(ow), C.byref(oh)), "shard_stitch")
            pano = None
            if want_output:
                pano = np.empty((3, oh.value, ow.value), np.uint8)
                self._check(self.L.pano_b200_result_copy(self.h, _p(pano)), "result_copy")
        buf = C.create_string_buffer(1 << 16)
        self.L.pano_b200_stitch_log(self.h, buf, 1 << 16)
        t = Times()
        self.L.pano_b200_stitch_times(self.h, C.byref(t))
        return pano, dict(log=buf.value.decode(), size=(ow.value, oh.value), times={f[0]: getattr(t, f[0]) for f in Times._fields_})

    # ---- plane-sharded canvas stages (pano_b200_shard_stitch_planes) ------------------------------------------------
    SEAM_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int), C.c_int)

    def shard_stitch_planes(self, first, count, exchange=None):
        """Everything of the stitch loop but the equalisation tail, on `count` colour planes starting at `first`.
        exchange(values | None, is_source) -> 4 ints: called once per edge; the rank carrying plane 0 passes the seam
        statistics in, the others receive them."""
        def _cb(user, p, is_src):
            try:
                out = exchange([p[i] for i in range(4)] if is_src else None, bool(is_src))
                for i in range(4):
                    p[i] = int(out[i])
                return 0
            except Exception:   # noqa: BLE001 -- must not propagate through the C frame
                import traceback
                traceback.print_exc()
                return 1
        cb = self.SEAM_CB(_cb) if exchange is not None else self.SEAM_CB()
        ow, oh = C.c_int(), C.c_int()
        self._check(self.L.pano_b200_shard_stitch_planes(self.h, int(first), int(count), cb, None, C.byref(ow), C.byref(oh)),
                    "shard_stitch_planes")
        buf = C.create_string_buffer(1 << 16)
        self.L.pano_b200_stitch_log(self.h, buf, 1 << 16)
        return dict(log=buf.value.decode(), size=(ow.value, oh.value))

    def shard_plane_export(self, k, out_t):
        self._check(self.L.pano_b200_shard_plane_export(self.h, int(k), C.c_void_p(out_t.data_ptr())), "shard_plane_export")

    def shard_plane_import(self, channel, in_t):
        self._check(self.L.pano_b200_shard_plane_import(self.h, int(channel), C.c_void_p(in_t.data_ptr())), "shard_plane_import")

    def shard_tail(self, want_output=True):
        ow, oh = C.c_int(), C.c_int()
        self._check(self.L.pano_b200_shard_tail(self.h, None, C.c_size_t(0), C.byref(ow), C.byref(oh)), "shard_tail")
        pano = None
        if want_output:
            pano = np.empty((3, oh.value, ow.value), np.uint8)
            self._check(self.L.pano_b200_result_copy(self.h, _p(pano)), "result_copy")
        t = Times()
        self.L.pano_b200_stitch_times(self.h, C.byref(t))
        return pano, dict(size=(ow.value, oh.value), times={f[0]: getattr(t, f[0]) for f in Times._fields_})

    def pairs(self, pairs):
        """Batched independent pairs (pano_b200_pairs): [(img_a, img_b), ...] -> PAIR_RECORD array (pair = position)."""
        from .dist import PAIR_RECORD
        flat = [_u8(im) for ab in pairs for im in ab]
        n = len(flat)
        rec = np.zeros(len(pairs), PAIR_RECORD)
        rec["pair"] = np.arange(len(pairs))
        if n:
            pp = (C.c_void_p * n)(*[im.ctypes.data for im in flat])
            ws = (C.c_int * n)(*[im.shape[2] for im in flat])
            hs = (C.c_int * n)(*[im.shape[1] for im in flat])
            self._check(self.L.pano_b200_pairs(self.h, pp, ws, hs, len(pairs), _p(rec)), "pairs")
        return rec

    def pairs_staged(self, dev_ptrs, sizes):
        """pano_b200_pairs_staged: dev_ptrs = 2 * npairs device pointers (planar RGB in HBM), sizes = [(w, h), ...]"""
        from .dist import PAIR_RECORD
        n = len(dev_ptrs)
        rec = np.zeros(n // 2, PAIR_RECORD)
        rec["pair"] = np.arange(n // 2)
        if n:
            pp = (C.c_void_p * n)(*dev_ptrs)
            ws = (C.c_int * n)(*[s[0] for s in sizes])
            hs = (C.c_int * n)(*[s[1] for s in sizes])
            self._check(self.L.pano_b200_pairs_staged(self.h, pp, ws, hs, n // 2, _p(rec)), "pairs_staged")
        return rec

    def stitch_features(self, projs, feats, match_idx=None):
        """matching() on precomputed projections / feature tables; match_idx: {(i, j): idx array} of preset pairs."""
        n = len(projs)
        projs = [_u8(p) for p in projs]
        descr = [np.ascontiguousarray(f[0], np.float32) for f in feats]
        keys = [np.ascontiguousarray(f[1], KEY_DTYPE) for f in feats]
        pp = (C.c_void_p * n)(*[p.ctypes.data for p in projs])
        ws = (C.c_int * n)(*[p.shape[2] for p in projs])
        hs = (C.c_int * n)(*[p.shape[1] for p in projs])
        pd = (C.c_void_p * n)(*[d.ctypes.data for d in descr])
        pk = (C.c_void_p * n)(*[k.ctypes.data for k in keys])
        nf = (C.c_int * n)(*[len(k) for k in keys])
        keep = []
        pm = None
        if match_idx:
            pm = (C.c_void_p * (n * n))()
            for (i, j), idx in match_idx.items():
                a = np.ascontiguousarray(idx, np.int32)
                assert len(a) == len(keys[j])
                keep.append(a)
                pm[i * n + j] = a.ctypes.data
        out, ow, oh = C.c_void_p(), C.c_int(), C.c_int()
        self._check(self.L.pano_b200_stitch_features(self.h, n, pp, ws, hs, pd, pk, nf, pm, C.byref(out), C.byref(ow), C.byref(oh)),
                    "stitch_features")
        pano = np.frombuffer(C.string_at(out, 3 * ow.value * oh.value), np.uint8).reshape(3, oh.value, ow.value).copy()
        self.L.pano_b200_free(out)
        buf = C.create_string_buffer(1 << 16)
        self.L.pano_b200_stitch_log(self.h, buf, 1 << 16)
        return pano, dict(log=buf.value.decode(), nfeat=[self.L.pano_b200_stitch_nfeatures(self.h, i) for i in range(n)])

    # ---- uint8 / tcgen05 matcher (north-star stage 3; not on the reference-parity path) ---------------------------
    def quantize_u8(self, descr):
        d = np.ascontiguousarray(descr, np.float32)
        out = np.empty(d.shape, np.uint8)
        self._check(self.L.pano_b200_quantize_u8(self.h, _p(d), len(d), _p(out)), "quantize_u8")
        return out

    def match_u8(self, A, B):
        """-> (idx [NB] (-1 = rejected by the ratio rule), d0, d1, nearest)"""
        A = np.ascontiguousarray(A, np.uint8)
        B = np.ascontiguousarray(B, np.uint8)
        idx = np.empty(len(B), np.int32)
        d01 = np.empty((len(B), 3), np.int32)
        n = C.c_int()
        self._check(self.L.pano_b200_match_u8(self.h, _p(A), len(A), _p(B), len(B), _p(idx), _p(d01), C.byref(n)), "match_u8")
        return idx, d01[:, 0].copy(), d01[:, 1].copy(), d01[:, 2].copy()

    def bench_match_u8(self, nA, nB, reps=5, A=None, B=None):
        """ms per repetition of the matcher kernels on resident tables (A, B: row-major u8 tables, or uniform bytes)"""
        ms = C.c_float()
        if A is not None:
            A = np.ascontiguousarray(A, np.uint8)
            B = np.ascontiguousarray(B, np.uint8)
            nA, nB = len(A), len(B)
        self._check(self.L.pano_b200_bench_match_u8(self.h, _p(A) if A is not None else None, nA,
                                                    _p(B) if B is not None else None, nB, reps, C.byref(ms)), "bench_match_u8")
        return ms.value

    def bench_match_u8_peak(self):
        """after bench_match_u8: (ms per repetition of the MMA-only run of the same kernel, K steps of 32 bytes per tile)"""
        ms, ks = C.c_float(), C.c_int()
        self._check(self.L.pano_b200_bench_match_u8_peak(self.h, C.byref(ms), C.byref(ks)), "bench_match_u8_peak")
        return ms.value, ks.value

    # ---- pipeline -------------------------------------------------------------------------------------------------
    def stitch(self, imgs):
        """ImageProcess(dir, n) on in-memory planar RGB images -> (panorama [3][H][W] u8, info dict)."""
        imgs = [_u8(i) for i in imgs]
        n = len(imgs)
        ptrs = (C.c_void_p * n)(*[i.ctypes.data for i in imgs])
        ws = (C.c_int * n)(*[i.shape[2] for i in imgs])
        hs = (C.c_int * n)(*[i.shape[1] for i in imgs])
        out, ow, oh = C.c_void_p(), C.c_int(), C.c_int()
        self._check(self.L.pano_b200_stitch(self.h, ptrs, ws, hs, n, C.byref(out), C.byref(ow), C.byref(oh)), "stitch")
        pano = np.frombuffer(C.string_at(out, 3 * ow.value * oh.value), np.uint8).reshape(3, oh.value, ow.value).copy()
        self.L.pano_b200_free(out)
        buf = C.create_string_buffer(1 << 16)
        self.L.pano_b200_stitch_log(self.h, buf, 1 << 16)
        t = Times()
        self.L.pano_b200_stitch_times(self.h, C.byref(t))
        info = dict(log=buf.value.decode(), nfeat=[self.L.pano_b200_stitch_nfeatures(self.h, i) for i in range(n)],
                    times={f[0]: getattr(t, f[0]) for f in Times._fields_})
        return pano, info


def fnv1a64(buf) -> str:
    """FNV-1a 64 over the planar bytes (hash convention of SURVEY.md 8c)."""
    data = np.ascontiguousarray(buf).tobytes()
    h = 0xCBF29CE484222325
    for b in data:
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"
