#!/usr/bin/env python
"""Stitch the reference's src/ex6 data sets with the ex6 profile and report wall time per job (GPU box).

    python tools/ex6_run.py [--reps 5] [--check]      # --check compares the SHA-256 with tests/golden/anchors.json
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import computervisionimagestich2_b200 as pano  # noqa: E402
from computervisionimagestich2_b200 import bmpio  # noqa: E402

DATA = os.path.join(ROOT, "oracle", "_ref", "data")
SETS = {1: ("Input", 4), 2: ("ex6_dataset2", 18), 3: ("ex6_dataset3", 11)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    anchors = json.load(open(os.path.join(ROOT, "tests", "golden", "anchors.json")))["ex6"]
    ctx = pano.Context(0)
    ctx.set_profile("ex6", anchors["ransac_seed"])
    for k, (name, n) in SETS.items():
        imgs = [bmpio.load_bmp(os.path.join(DATA, name, f"{i + 1}.bmp")) for i in range(n)]
        mpix = sum(i.shape[1] * i.shape[2] for i in imgs) / 1e6
        best, out, info = 1e9, None, None
        for _ in range(args.reps):
            t = time.perf_counter()
            out, info = ctx.stitch(imgs)
            best = min(best, time.perf_counter() - t)
        a = anchors[f"dataset{k}"]
        ok = hashlib.sha256(out.tobytes()).hexdigest() == a["sha256"]
        tm = info["times"]
        print(f"dataset{k}: {n} images {mpix:.2f} Mpixel -> {out.shape[2]}x{out.shape[1]}  best {best * 1e3:.1f} ms "
              f"({mpix / best:.1f} Mpixel/s; reference {a['cpu_seconds']} s on one core = {a['cpu_seconds'] / best:.0f}x)  "
              f"bit-exact {ok}  [sift {tm['sift']:.1f} match {tm['match']:.1f} ransac {tm['ransac']:.1f} warp {tm['warp']:.1f} "
              f"blend {tm['blend']:.1f} tail {tm['tail']:.1f} ms]", flush=True)
        if args.check and not ok:
            sys.exit(1)


if __name__ == "__main__":
    main()
