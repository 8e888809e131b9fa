"""Stitch the same views several times and report which stage first differs (match lists, RANSAC, panorama)."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    import bench
    import computervisionimagestich2_b200 as pano
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    mode = sys.argv[2] if len(sys.argv) > 2 else "prefilter"
    views = bench.synth_scene_views(n, 3840, 2160)
    ctx = pano.Context(0)
    ctx.set_match_mode(mode)
    feats = []
    for v in views:
        p, g = ctx.project(v, want_gray=True)
        d, k = ctx.sift_features(g)
        feats.append((d, k))
        print("features", len(k), hashlib.sha256(d.tobytes()).hexdigest()[:12])
    for rep in range(3):
        for i in range(n):
            for j in range(i + 1, n):
                ab, ba = ctx.match_pair(feats[i][0], feats[j][0])
                print(rep, i, j, (ab >= 0).sum(), (ba >= 0).sum(), hashlib.sha256(ab.tobytes()).hexdigest()[:12], hashlib.sha256(ba.tobytes()).hexdigest()[:12], ctx.match_stats(reset=True))
    for rep in range(3):
        p, info = ctx.stitch(views)
        print("stitch", rep, p.shape, hashlib.sha256(p.tobytes()).hexdigest()[:16], info["log"].replace("\n", " | "))


if __name__ == "__main__":
    main()
