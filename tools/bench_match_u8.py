"""Throughput of the uint8 / tcgen05 matcher kernel alone (resident synthetic tables, CUDA events)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import computervisionimagestich2_b200 as pano  # noqa: E402

ctx = pano.Context(0)
out = []
for na, nb in ((2263, 2227), (18000, 18000), (36000, 36000), (72000, 72000), (144000, 72000)):
    ms = ctx.bench_match_u8(na, nb, 5)
    ops = 2.0 * 128 * na * nb
    out.append({"nA": na, "nB": nb, "ms": round(ms, 4), "int8_TOPS": round(ops / (ms * 1e-3) / 1e12, 1),
                "frac_of_4500_nominal": round(ops / (ms * 1e-3) / 4.5e15, 4)})
    print(out[-1], flush=True)
print(json.dumps(out))
