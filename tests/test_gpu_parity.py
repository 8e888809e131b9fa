"""GPU tier: the CUDA path through the C ABI against the compiled reference (oracle/_ref), bit-exact."""
import numpy as np
import pytest

from conftest import synth_rgb

pytestmark = pytest.mark.gpu


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a.view(np.uint64)


@pytest.mark.parametrize("name,idx", [("Input", 0), ("Input", 3), ("Input2", 1)])
def test_project_gray(ctx, ref, input_sets, name, idx):
    img = input_sets[name][idx]
    pr, g = ctx.project(img, want_gray=True)
    rp = ref.project(img)
    assert np.array_equal(pr, rp)
    assert np.array_equal(g, ref.gray(rp))
    assert np.array_equal(ctx.gray(img), ref.gray(img))


@pytest.mark.parametrize("w,h", [(37, 53), (64, 64), (131, 40), (40, 131)])
def test_project_odd_sizes(ctx, ref, w, h):
    img = synth_rgb(w, h, w * 1000 + h)
    assert np.array_equal(ctx.project(img), ref.project(img))


@pytest.mark.parametrize("name,idx", [("Input", 0), ("Input2", 2)])
def test_sift_scale_space_and_descriptors(ctx, ref, input_sets, name, idx):
    g = ref.gray(ref.project(input_sets[name][idx])).astype(np.float32)
    rd = ref.sift_dump(g)
    raw = ctx.sift_raw(g)
    assert raw["octave_nkeys"] == [len(o["keys"]) for o in rd]
    for o, r in enumerate(rd):
        gss, grad = ctx.sift_octave_dump(o)
        assert np.array_equal(_bits(gss), _bits(r["gss"])), f"gss octave {o}"
        assert np.array_equal(_bits(grad), _bits(r["grad"])), f"grad octave {o}"
    rk = np.concatenate([o["keys"][o["descr_key"]] for o in rd])
    rdsc = np.concatenate([o["descr"] for o in rd])
    rang = np.concatenate([o["angles"][o["descr_key"], :].reshape(-1, 4)[np.arange(len(o["descr_key"])),
                           _angle_slot(o)] for o in rd])
    wr = np.concatenate([o["descr_written"] for o in rd]).astype(bool)
    assert raw["keys"].tobytes() == rk[wr].tobytes()
    assert np.array_equal(_bits(raw["angles"]), _bits(rang[wr]))
    assert np.array_equal(_bits(raw["descr"]), _bits(rdsc[wr]))


def _angle_slot(o):
    """index (0..3) of each descriptor's angle inside its keypoint's angle list"""
    dk = o["descr_key"]
    slot = np.zeros(len(dk), np.int64)
    for i in range(1, len(dk)):
        slot[i] = slot[i - 1] + 1 if dk[i] == dk[i - 1] else 0
    return slot


@pytest.mark.parametrize("w,h", [(97, 61), (160, 120), (33, 200)])
def test_sift_features_small_random(ctx, ref, w, h):
    g = ref.gray(synth_rgb(w, h, 7 * w + h))
    d, k = ctx.sift_features(g)
    rd, rk = ref.sift_features(g)
    assert len(k) == len(rk)
    assert np.array_equal(_bits(d), _bits(rd))
    assert k.tobytes() == rk.tobytes()


def test_feature_tables_and_matching(ctx, ref, input_sets):
    for name in ("Input", "Input2"):
        feats = []
        for img in input_sets[name]:
            g = ref.gray(ref.project(img))
            d, k = ctx.sift_features(g)
            rd, rk = ref.sift_features(g)
            assert np.array_equal(_bits(d), _bits(rd)) and k.tobytes() == rk.tobytes()
            feats.append((d, k))
        pairs = [(0, 1), (1, 0), (1, 2), (2, 3), (3, 2), (0, 3)] if name == "Input" else [(1, 2), (2, 1), (0, 3)]
        for i, j in pairs:
            ma, mb = ctx.match(feats[i][0], feats[i][1], feats[j][0], feats[j][1])
            ra, rb = ref.match(feats[i][0], feats[i][1], feats[j][0], feats[j][1])
            assert ma.tobytes() == ra.tobytes() and mb.tobytes() == rb.tobytes(), (name, i, j)


def test_ransac_warp_blend_tail(ctx, ref, input_sets):
    imgs = input_sets["Input"]
    proj = [ref.project(i) for i in imgs]
    feats = [ref.sift_features(ref.gray(p)) for p in proj]
    _, h, w = imgs[0].shape
    a, b = ref.match(feats[2][0], feats[2][1], feats[3][0], feats[3][1])   # src=2 (A), dst=3 (B)
    # forward_H = RANSAC(dst->src pairs), backward_H = RANSAC(src->dst pairs)
    Hf, counts, hyps, inl = ctx.ransac(b, a, debug=True)
    assert np.array_equal(_bits(Hf), _bits(ref.ransac(b, a)))
    for k in (0, 17, 71):
        assert len(ref.inliers(b, a, hyps[k])) == counts[k]
    Hb = ctx.ransac(a, b)
    assert np.array_equal(_bits(Hb), _bits(ref.ransac(a, b)))
    bounds, size = ctx.plan_canvas(w, h, Hf, w, h)
    rb = ref.warp_bounds(w, h, Hf)
    assert bounds[0] == min(rb[0], 0) and bounds[1] == min(rb[1], 0)
    cw, ch = int(size[0]), int(size[1])
    wa, sb = ctx.warp_shift(proj[3], Hb, float(bounds[0]), float(bounds[1]), proj[2], int(bounds[0]), int(bounds[1]), cw, ch)
    assert np.array_equal(wa, ref.warp(proj[3], Hb, float(bounds[0]), float(bounds[1]), cw, ch))
    assert np.array_equal(sb, ref.shift(proj[2], int(bounds[0]), int(bounds[1]), cw, ch))
    bl = ctx.blend(wa, sb)
    rbl = ref.blend(wa, sb)
    assert np.array_equal(bl, rbl)
    assert np.array_equal(ctx.equalize_mix(bl), ref.equalize_mix(rbl))


@pytest.mark.parametrize("w,h,c", [(64, 48, 3), (131, 77, 1), (33, 2, 2), (5, 300, 1), (300, 5, 3)])
def test_cimg_primitives(ctx, ref, w, h, c):
    rng = np.random.default_rng(w * h + c)
    p = (rng.random((c, h, w)) * 255).astype(np.float32)
    assert np.array_equal(_bits(ctx.cimg_blur2(p)), _bits(ref.cimg_blur2(p)))
    if w >= 2 and h >= 2:
        assert np.array_equal(_bits(ctx.cimg_resize3(p, w // 2, h // 2)), _bits(ref.cimg_resize3(p, w // 2, h // 2)))
    assert np.array_equal(_bits(ctx.cimg_resize3(p, 2 * w + 1, 2 * h)), _bits(ref.cimg_resize3(p, 2 * w + 1, 2 * h)))


@pytest.mark.parametrize("name", ["Input", "Input2"])
def test_full_panorama(ctx, ref, input_sets, name):
    pano, info = ctx.stitch(input_sets[name])
    rp, rinfo = ref.stitch_mem(input_sets[name])
    assert info["log"] == rinfo["log"]
    assert info["nfeat"] == rinfo["nfeat"]
    assert pano.shape == rp.shape
    assert np.array_equal(pano, rp)
