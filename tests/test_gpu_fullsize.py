"""GPU tier, BASELINE.json full sizes (configs[2]: synthetic 3840x2160 views).  The oracle cannot stitch eight 4K views
in test time (hours of scalar matching), so parity at this size is checked stage by stage on bounded pieces -- each
piece still bit-exact against the compiled reference -- plus size-independent properties of the whole job."""
import hashlib
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a.view(np.uint64)


@pytest.fixture(scope="module")
def views():
    import bench
    return bench.synth_scene_views(3, 3840, 2160)


@pytest.fixture(scope="module")
def feats(ctx, views):
    out = []
    for v in views[:2]:
        p, g = ctx.project(v, want_gray=True)
        d, k = ctx.sift_features(g)
        out.append((p, g, d, k))
    return out


def test_4k_projection_and_sift_one_image_bit_exact(ref, views, feats):
    p, g, d, k = feats[0]
    rp = ref.project(views[0])
    assert np.array_equal(p, rp)
    assert np.array_equal(g, ref.gray(rp))
    rd, rk = ref.sift_features(g)
    assert len(k) == len(rk) and len(k) > 10000
    assert np.array_equal(_bits(d), _bits(rd)) and k.tobytes() == rk.tobytes()


def test_4k_matching_against_the_reference_on_a_query_sample(ctx, ref, feats):
    """exact L1 2-NN of 300 queries of view 1 against ALL features of view 0 (the reference's kd-forest, CPU)"""
    (_, _, d0, k0), (_, _, d1, k1) = feats
    sel = np.linspace(0, len(k1) - 1, 300).astype(int)
    idx = ctx.match_idx(d0, d1)
    ra, rb = ref.match(d0, k0, d1[sel], k1[sel])
    got_a = k0[idx[sel][idx[sel] >= 0]]
    got_b = k1[sel][idx[sel] >= 0]
    assert got_a.tobytes() == ra.tobytes() and got_b.tobytes() == rb.tobytes()


def test_4k_blend_and_tail_bit_exact(ctx, ref, feats):
    """one multiband blend + equalisation on a 4K-class canvas (two projected views, half overlapping)"""
    pa, pb = feats[0][0], feats[1][0]
    _, h, w = pa.shape
    cw = w + w // 2
    a = np.zeros((3, h, cw), np.uint8)
    b = np.zeros((3, h, cw), np.uint8)
    a[:, :, :w] = pa
    b[:, :, w // 2:] = pb
    out = ctx.blend(a, b)
    rout = ref.blend(a, b)
    assert np.array_equal(out, rout)
    assert np.array_equal(ctx.equalize_mix(out), ref.equalize_mix(rout))


def test_4k_three_view_job_properties(ctx, views):
    """whole job on 3 x 4K views: deterministic bytes, chain order, canvas at least as large as its inputs"""
    p1, i1 = ctx.stitch(views)
    p2, i2 = ctx.stitch(views)
    assert hashlib.sha256(p1.tobytes()).hexdigest() == hashlib.sha256(p2.tobytes()).hexdigest()
    assert i1["log"] == i2["log"]
    lines = i1["log"].split("\n")
    assert lines[0] == "1" and len([x for x in lines if x.strip()]) == 3   # middle view first, two edges
    assert p1.shape[1] >= 2160 and p1.shape[2] > 3840
    assert min(i1["nfeat"]) > 10000


@pytest.fixture(scope="module")
def view8k():
    import bench
    return bench.synth_scene_views(1, 7680, 4320)[0]


def test_8k_blend_and_tail_bit_exact(ctx, ref, view8k):
    """one multiband blend + equalisation on an 8K-class canvas (11520 x 4320 = 50 Mpixel, 12 pyramid levels, 1.4 GB of
    level-0 planes): 64-bit offsets in every canvas kernel"""
    h, w = view8k.shape[1:]
    cw = w + w // 2
    a = np.zeros((3, h, cw), np.uint8)
    b = np.zeros((3, h, cw), np.uint8)
    a[:, :, :w] = view8k | 1
    b[:, :, w // 2:] = view8k[:, ::-1, :] | 1
    out = ctx.blend(a, b)
    rout = ref.blend(a, b)
    assert np.array_equal(out, rout)
    assert np.array_equal(ctx.equalize_mix(out), ref.equalize_mix(rout))


def test_8k_projection_and_sift_one_image_bit_exact(ctx, ref, view8k):
    """BASELINE.json configs[3] image size (7680x4320, 33 Mpixel): one view through projection + gray + SIFT against the
    compiled reference (about half a minute of CPU).  Exercises 32-bit index limits of the scale-space kernels."""
    v = view8k
    assert v.shape == (3, 4320, 7680)
    p, g = ctx.project(v, want_gray=True)
    rp = ref.project(v)
    assert np.array_equal(p, rp)
    assert np.array_equal(g, ref.gray(rp))
    d, k = ctx.sift_features(g)
    rd, rk = ref.sift_features(g)
    assert len(k) == len(rk) and len(k) > 40000
    assert np.array_equal(_bits(d), _bits(rd)) and k.tobytes() == rk.tobytes()
