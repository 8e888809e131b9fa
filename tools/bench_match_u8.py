"""Throughput of the uint8 / tcgen05 matcher kernel alone (resident synthetic tables, CUDA events)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import computervisionimagestich2_b200 as pano  # noqa: E402

import numpy as np  # noqa: E402


def sift_like(n, seed):
    """unit-norm, clamped, renormalised non-negative rows (the statistics of SIFT descriptors), VLFeat-quantised"""
    rng = np.random.default_rng(seed)
    x = rng.gamma(0.6, 1.0, (n, 128)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x = np.minimum(x, 0.2)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.minimum(512.0 * x, 255.0).astype(np.uint8)


ctx = pano.Context(0)
out = []
for kind in ("sift_like", "uniform_bytes"):
    for na, nb in ((2263, 2227), (18000, 18000), (72000, 72000), (144000, 72000)):
        if kind == "sift_like":
            ms = ctx.bench_match_u8(0, 0, 5, A=sift_like(na, 1), B=sift_like(nb, 2))
        else:
            ms = ctx.bench_match_u8(na, nb, 5)
        ops = 2.0 * 128 * na * nb
        out.append({"data": kind, "nA": na, "nB": nb, "ms": round(ms, 4), "int8_TOPS": round(ops / (ms * 1e-3) / 1e12, 1),
                    "frac_of_4500_nominal": round(ops / (ms * 1e-3) / 4.5e15, 4)})
        print(out[-1], flush=True)
print(json.dumps(out))
