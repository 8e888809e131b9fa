"""Diagnostic: wall time + stage times of staged vs host-buffer stitches."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import computervisionimagestich2_b200 as pano
name = sys.argv[1] if len(sys.argv) > 1 else "input2"
imgs, desc, _ = bench.load_workload(name)
L = pano.lib(); ctx = pano.Context(0); n = len(imgs)
ptrs = (C.c_void_p * n)(*[i.ctypes.data for i in imgs])
ws = (C.c_int * n)(*[i.shape[2] for i in imgs]); hs = (C.c_int * n)(*[i.shape[1] for i in imgs])
ctx._check(L.pano_b200_stage_images(ctx.h, ptrs, ws, hs, n), "stage")
ow, oh = C.c_int(), C.c_int()
def times():
    t = pano.Times(); L.pano_b200_stitch_times(ctx.h, C.byref(t))
    return {f[0]: round(getattr(t, f[0]), 2) for f in pano.Times._fields_[:9]}
for r in range(5):
    L.pano_b200_flush_l2(ctx.h)
    t0 = time.perf_counter()
    ctx._check(L.pano_b200_stitch_staged(ctx.h, C.byref(ow), C.byref(oh)), "stitch")
    print("staged", r, round((time.perf_counter() - t0) * 1e3, 2), times(), flush=True)
in_bytes = [3 * i.shape[1] * i.shape[2] for i in imgs]
pin_in = []
for im, b in zip(imgs, in_bytes):
    p = L.pano_b200_alloc_pinned(C.c_size_t(b)); C.memmove(p, im.ctypes.data, b); pin_in.append(p)
out_cap = 3 * ow.value * oh.value
pin_out = L.pano_b200_alloc_pinned(C.c_size_t(out_cap))
pptrs = (C.c_void_p * n)(*pin_in)
for r in range(5):
    L.pano_b200_flush_l2(ctx.h)
    t0 = time.perf_counter()
    ctx._check(L.pano_b200_stitch_into(ctx.h, pptrs, ws, hs, n, C.c_void_p(pin_out), C.c_size_t(out_cap), C.byref(ow), C.byref(oh)), "into")
    print("pinned", r, round((time.perf_counter() - t0) * 1e3, 2), times(), flush=True)
for r in range(3):
    t0 = time.perf_counter()
    ctx._check(L.pano_b200_stitch_into(ctx.h, ptrs, ws, hs, n, C.c_void_p(pin_out), C.c_size_t(out_cap), C.byref(ow), C.byref(oh)), "into")
    print("pageable-in", r, round((time.perf_counter() - t0) * 1e3, 2), times(), flush=True)
