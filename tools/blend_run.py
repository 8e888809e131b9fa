"""One multiband blend of a synthetic canvas pair through pano_b200_blend, with per-kernel CUDA-event times.

    python tools/blend_run.py [W] [H] [reps]      (default 17997 x 2268: the last edge of the 8 x 4K job)
"""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def canvas_pair(w, h, seed=5):
    """a = the new image warped into the right part of the canvas, b = the panorama so far in the left part; black
    borders top / bottom as a warped view leaves them"""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (3, h // 8 + 2, w // 8 + 2), dtype=np.uint8)
    img = np.repeat(np.repeat(base, 8, 1), 8, 2)[:, :h, :w].copy()
    img += rng.integers(0, 8, img.shape, dtype=np.uint8)
    a, b = np.zeros_like(img), np.zeros_like(img)
    m = max(2, h // 40)
    xa, xb = int(w * 0.70), int(w * 0.82)
    a[:, m:h - m, xa:] = np.maximum(img[:, m:h - m, xa:], 1)
    b[:, m // 2:h - m // 2, :xb] = np.maximum(img[:, m // 2:h - m // 2, :xb], 1)
    return a, b


def main():
    import computervisionimagestich2_b200 as pano
    w = int(sys.argv[1]) if len(sys.argv) > 1 else 17997
    h = int(sys.argv[2]) if len(sys.argv) > 2 else 2268
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    a, b = canvas_pair(w, h)
    ctx = pano.Context(0)
    L = pano.lib()
    import hashlib
    for r in range(reps):
        L.pano_b200_ktimer_reset()
        L.pano_b200_ktimer_enable(1 if r == reps - 1 else 0)
        t0 = time.perf_counter()
        out = ctx.blend(a, b)
        dt = (time.perf_counter() - t0) * 1e3
        print(f"rep {r}: blend {w}x{h}: {dt:.2f} ms wall (host buffers), sha256 {hashlib.sha256(out.tobytes()).hexdigest()[:16]}")
    buf = C.create_string_buffer(1 << 16)
    L.pano_b200_ktimer_report(buf, 1 << 16)
    k = json.loads(buf.value.decode())
    for name, v in sorted(k.items(), key=lambda kv: -kv[1]["ms"]):
        gb = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0
        print(f"{name:22s} {v['ms']:9.3f} ms {v['launches']:5d} launches  {gb:9.1f} GB/s")


if __name__ == "__main__":
    main()
