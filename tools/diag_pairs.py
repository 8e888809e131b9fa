"""Where the time of one pair job goes (tools/run_pairs.py): wall clock per ABI call, medians over the pairs.
    python tools/diag_pairs.py [npairs=6] [width=1920] [height=1080]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import computervisionimagestich2_b200 as pano  # noqa: E402

npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 6
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
ctx = pano.Context(0)
rows = []
for p in range(npairs + 1):
    a, b = bench.synth_scene_views(2, w, h, seed=20181126 + p)
    t = [time.perf_counter()]
    _, da, ka = ctx.extract(a); t.append(time.perf_counter())
    _, db, kb = ctx.extract(b); t.append(time.perf_counter())
    i01 = ctx.match_idx(da, db); t.append(time.perf_counter())
    i10 = ctx.match_idx(db, da); t.append(time.perf_counter())
    s = i01 >= 0
    ctx.ransac(ka[i01[s]].copy(), kb[s].copy()); t.append(time.perf_counter())
    s = i10 >= 0
    ctx.ransac(kb[i10[s]].copy(), ka[s].copy()); t.append(time.perf_counter())
    if p:
        rows.append(np.diff(t) * 1e3)
m = np.median(np.array(rows), axis=0)
print(json.dumps({"image": [w, h], "nfeat": [len(ka), len(kb)], "ms": dict(zip(
    ["extract_a", "extract_b", "match_ab", "match_ba", "ransac_ab", "ransac_ba"], [round(float(x), 2) for x in m])),
    "ms_per_pair": round(float(m.sum()), 2), "host_cpus": os.cpu_count()}))
