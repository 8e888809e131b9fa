"""Host timeline of one staged job: PANO_B200_TRACE=1 python tools/trace_run.py [workload] > trace.txt 2>&1"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PANO_B200_TRACE", "1")


def main():
    import bench
    import computervisionimagestich2_b200 as pano
    name = sys.argv[1] if len(sys.argv) > 1 else "synth4k"
    imgs, desc, _ = bench.load_workload(name)
    L = pano.lib(); ctx = pano.Context(0); n = len(imgs)
    ptrs = (C.c_void_p * n)(*[i.ctypes.data for i in imgs])
    ws = (C.c_int * n)(*[i.shape[2] for i in imgs]); hs = (C.c_int * n)(*[i.shape[1] for i in imgs])
    ctx._check(L.pano_b200_stage_images(ctx.h, ptrs, ws, hs, n), "stage")
    ow, oh = C.c_int(), C.c_int()
    if os.environ.get("LANES"):
        L.pano_b200_set_lanes(ctx.h, int(os.environ["LANES"]))
    for r in range(4):
        L.pano_b200_flush_l2(ctx.h)
        print(f"trace === run {r} begin", file=sys.stderr, flush=True)
        t0 = time.perf_counter()
        ctx._check(L.pano_b200_stitch_staged(ctx.h, C.byref(ow), C.byref(oh)), "stitch")
        print(f"trace === run {r} end {(time.perf_counter() - t0) * 1e3:.2f} ms", file=sys.stderr, flush=True)


if __name__ == "__main__":
    main()
