"""Per-kernel CUDA-event times of one stitch of a workload (single lane, so kernels do not overlap)."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import computervisionimagestich2_b200 as pano
name = sys.argv[1] if len(sys.argv) > 1 else "input2"
ex6 = name.startswith("ex6_dataset")   # ex6_dataset2 / ex6_dataset3: the src/ex6 sets with the ex6 profile
if ex6:
    from computervisionimagestich2_b200 import bmpio
    k = int(name[-1])
    d = os.path.join(ROOT, "oracle", "_ref", "data", name)
    imgs = [bmpio.load_bmp(os.path.join(d, f"{i + 1}.bmp")) for i in range({2: 18, 3: 11}[k])]
    desc = f"src/ex6/dataset{k} ({len(imgs)} images), ex6 profile"
else:
    imgs, desc, _ = bench.load_workload(name)
if len(sys.argv) > 2:
    imgs = imgs[: int(sys.argv[2])]
L = pano.lib(); ctx = pano.Context(0); n = len(imgs)
if ex6:
    ctx.set_profile("ex6", 666666)
ptrs = (C.c_void_p * n)(*[i.ctypes.data for i in imgs])
ws = (C.c_int * n)(*[i.shape[2] for i in imgs]); hs = (C.c_int * n)(*[i.shape[1] for i in imgs])
ctx._check(L.pano_b200_stage_images(ctx.h, ptrs, ws, hs, n), "stage")
ow, oh = C.c_int(), C.c_int()
L.pano_b200_set_lanes(ctx.h, 1)
for r in range(2):
    L.pano_b200_ktimer_reset()
    L.pano_b200_ktimer_enable(1 if r == 1 else 0)
    L.pano_b200_flush_l2(ctx.h)
    ctx._check(L.pano_b200_stitch_staged(ctx.h, C.byref(ow), C.byref(oh)), "stitch")
buf = C.create_string_buffer(1 << 16)
L.pano_b200_ktimer_report(buf, 1 << 16)
k = json.loads(buf.value.decode())
print(desc, "->", ow.value, "x", oh.value)
for name, v in sorted(k.items(), key=lambda kv: -kv[1]["ms"]):
    gb = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0
    print(f"{name:22s} {v['ms']:9.3f} ms {v['launches']:5d} launches  {gb:9.1f} G(B|op)/s")
print("match stats", ctx.match_stats())
