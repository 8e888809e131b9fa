"""GPU tier: whole jobs on the SYNTHETIC workloads of BASELINE.json (tools/synth_scene.py, the SURVEY 8(d) generator) against
the compiled reference, at sizes the CPU reference finishes in about a minute: an 8-view panorama (configs[2] scaled down)
and three 1920x1080 pairs (configs[4]), bit for bit -- stitch order, feature counts, panorama bytes, match counts, RANSAC
coefficients.  The full-size runs are covered by size-independent checks in bench.py (same SHA-256 at every GPU count,
pre-filter == full scan)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_synthetic_panorama_8_views_bit_exact(ctx, ref):
    import synth_scene
    ctx.set_profile("root", 666666)
    views = synth_scene.views(8, 640, 360)
    want, winfo = ref.stitch_mem(views)
    for mode in ("prefilter", "full"):
        ctx.set_match_mode(mode)
        got, info = ctx.stitch(views)
        assert info["log"] == winfo["log"] and info["nfeat"] == winfo["nfeat"], mode
        assert got.shape == want.shape and np.array_equal(got, want), mode
    ctx.set_match_mode("prefilter")
    order = [int(x) for x in winfo["log"].split()]
    assert len(order) == 1 + 2 * 7                       # all eight views were chained: the generator keeps neighbours adjacent


def test_synthetic_1080p_pairs_bit_exact(ctx, ref):
    import synth_scene
    from computervisionimagestich2_b200 import dist as pdist
    ctx.set_profile("root", 666666)
    pairs = [synth_scene.pair(p) for p in range(3)]
    table = ctx.pairs(pairs)
    for rec, (a, b) in zip(table, pairs):
        fa = ref.sift_features(ref.gray(ref.project(a)))
        fb = ref.sift_features(ref.gray(ref.project(b)))
        assert list(rec["nfeat"]) == [len(fa[1]), len(fb[1])]
        for d, (s, t) in enumerate(((fa, fb), (fb, fa))):
            src, dst = ref.match(s[0], s[1], t[0], t[1])
            assert rec["nmatch"][d] == len(src)
            assert bool(rec["has_h"][d]) == (len(src) >= pdist.THRESHOLD)
            if rec["has_h"][d]:
                assert ref.ransac(src, dst).tobytes() == rec["H"][d].tobytes()
        assert rec["has_h"].all()                        # 50 % overlap: both directions are adjacent
