"""Small driver for ncu: stage a workload in HBM, stitch it `reps` times (first = warm-up).

    python tools/profile_run.py [input|input2|synth4k] [reps]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import computervisionimagestich2_b200 as pano  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "input2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
imgs, desc, _ = bench.load_workload(name)
L = pano.lib()
ctx = pano.Context(0)
n = len(imgs)
ptrs = (C.c_void_p * n)(*[i.ctypes.data for i in imgs])
ws = (C.c_int * n)(*[i.shape[2] for i in imgs])
hs = (C.c_int * n)(*[i.shape[1] for i in imgs])
ctx._check(L.pano_b200_stage_images(ctx.h, ptrs, ws, hs, n), "stage")
ow, oh = C.c_int(), C.c_int()
for r in range(reps):
    L.pano_b200_ktimer_reset()
    ctx._check(L.pano_b200_stitch_staged(ctx.h, C.byref(ow), C.byref(oh)), "stitch")
    print(f"rep {r}: {desc}: panorama {ow.value}x{oh.value}, launches {L.pano_b200_ktimer_launches()}")
