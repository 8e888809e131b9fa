"""Per-kernel split of the uint8 / tcgen05 matcher (main MMA kernel, exact finish kernel, table preparation) on uniform
72k x 72k tables:  python tools/u8split.py   (GPU box)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import computervisionimagestich2_b200 as pano
L = pano.lib(); ctx = pano.Context(0)
L.pano_b200_ktimer_reset(); L.pano_b200_ktimer_enable(1)
ms = ctx.bench_match_u8(72000, 72000, 5)
buf = C.create_string_buffer(1 << 16); L.pano_b200_ktimer_report(buf, 1 << 16)
k = json.loads(buf.value.decode())
print("total ms/rep", ms)
for n, v in k.items(): print(n, v["ms"] / v["launches"], v["launches"])
