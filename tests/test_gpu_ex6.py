"""GPU tier: the `src/ex6` profile of the CUDA path through the C ABI against the reference's ex6
variant compiled from its own sources (oracle/_ref/libpano_ref_ex6.so, RANSAC seed pinned) and against the committed
golden anchors.  Bit-exact."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import synth_rgb

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 666666


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ex6():
    from oracle import ref_ex6_api
    if not ref_ex6_api.available():
        pytest.skip("oracle/_ref/libpano_ref_ex6.so not built (make -C oracle ref_ex6 needs /root/reference)")
    ref_ex6_api.set_seed(SEED)
    return ref_ex6_api


@pytest.fixture()
def ctx6(ctx):
    ctx.set_profile("ex6", SEED)
    yield ctx
    ctx.set_profile("root", 666666)


@pytest.fixture(scope="module")
def anchors():
    return json.load(open(os.path.join(HERE, "golden", "anchors.json")))["ex6"]


def _load(k):
    from computervisionimagestich2_b200 import bmpio
    from oracle import ref_ex6_api
    d = ref_ex6_api.dataset_dir(k)
    if not os.path.isdir(d):
        pytest.skip(f"{d} not staged (make -C oracle ref ref_ex6)")
    return [bmpio.load_bmp(os.path.join(d, f"{i + 1}.bmp")) for i in range(ref_ex6_api.DATASET_SIZES[k])]


# lines of 1, < 32, = 32, ragged multiples; more lines than one CTA; a single row / column (one axis skipped)
@pytest.mark.parametrize("c,h,w", [(1, 1, 50), (2, 64, 1), (3, 37, 70), (1, 32, 32), (7, 131, 259), (1, 5, 1000), (2, 700, 33)])
def test_deriche_blur(ctx, ex6, c, h, w):
    rng = np.random.default_rng(c * 100000 + h * 1000 + w)
    p = (rng.random((c, h, w)) * 255).astype(np.float32)
    got = ctx.cimg_blur2(p, deriche=True)
    assert np.array_equal(got.view(np.uint32), ex6.cimg_blur2(p).view(np.uint32))


def test_deriche_blur_signed_and_constant(ctx, ex6):
    rng = np.random.default_rng(7)
    p = ((rng.random((3, 90, 141)) - 0.5) * 512).astype(np.float32)   # Laplacian-like signed input
    p[1] = 17.25
    assert np.array_equal(ctx.cimg_blur2(p, deriche=True).view(np.uint32), ex6.cimg_blur2(p).view(np.uint32))


def _canvas_pair(w, h, seed, cut_a, cut_b):
    rng = np.random.default_rng(seed)
    a = np.zeros((3, h, w), np.uint8)
    b = np.zeros((3, h, w), np.uint8)
    a[:, :, :cut_a] = synth_rgb(cut_a, h, seed) | 1
    b[:, :, cut_b:] = synth_rgb(w - cut_b, h, seed + 1) | 1
    a[1, h // 2, 10:40] = 0   # pixels the 3-channel test skips but the root variant's channel-0 test counts
    b[2, h // 2, cut_b + 5:cut_b + 9] = 0
    return a, b


@pytest.mark.parametrize("w,h,ca,cb", [(420, 300, 260, 180), (300, 420, 200, 90), (1081, 527, 700, 500), (129, 64, 80, 40)])
def test_blend(ctx6, ex6, w, h, ca, cb):
    a, b = _canvas_pair(w, h, w + h, ca, cb)
    assert np.array_equal(ctx6.blend(a, b), ex6.blend(a, b))
    assert np.array_equal(ctx6.blend(b, a), ex6.blend(b, a))   # the other mask branch


def test_tail(ctx6, ex6):
    img = synth_rgb(331, 197, 5)
    assert np.array_equal(ctx6.equalize_mix(img), ex6.tail(img))


@pytest.mark.parametrize("seed", [666666, 1, 1543599813])
def test_ransac_seed(ctx, ex6, ref, input_sets, seed):
    g = [ref.gray(ref.project(im)) for im in input_sets["Input"][2:4]]
    (da, ka), (db, kb) = ctx.sift_features(g[0]), ctx.sift_features(g[1])
    ma, mb = ctx.match(da, ka, db, kb)
    ex6.set_seed(seed)
    ctx.set_profile("ex6", seed)
    try:
        got = ctx.ransac(ma, mb)
        want = ex6.ransac(ma, mb)
    finally:
        ex6.set_seed(SEED)
        ctx.set_profile("root", 666666)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def test_pipeline_dataset1_vs_oracle(ctx6, ex6, anchors):
    imgs = _load(1)
    pano, info = ctx6.stitch(imgs)
    want, winfo = ex6.stitch_mem(imgs)
    assert pano.shape == want.shape and np.array_equal(pano, want)
    assert info["log"] == winfo["log"]
    assert sha(pano) == anchors["dataset1"]["sha256"]


@pytest.mark.parametrize("k", [3, 2])
def test_pipeline_golden(ctx6, anchors, k):
    """18- and 11-image sets (the reference needs 160 s / 50 s of CPU for them): committed anchors only."""
    a = anchors[f"dataset{k}"]
    pano, info = ctx6.stitch(_load(k))
    assert list(pano.shape) == [3, a["height"], a["width"]]
    assert info["nfeat_initial"] == a["nfeat"] if "nfeat_initial" in info else True
    assert info["log"] == a["log"]
    assert sha(pano) == a["sha256"]


def test_two_images_have_no_edge(ctx6):
    """src/ex6/ImageProcess.cpp:152-160: with n = 2 the chain from image 1 has no neighbour list, so the result is the
    tail applied to the projection of image 1."""
    imgs = [synth_rgb(96, 128, 1), synth_rgb(96, 128, 2)]
    pano, _ = ctx6.stitch(imgs)
    assert np.array_equal(pano, ctx6.equalize_mix(ctx6.project(imgs[1])))


def test_landscape_input_is_an_error(ctx6):
    import computervisionimagestich2_b200 as pano
    imgs = [synth_rgb(128, 96, i) for i in range(3)]
    with pytest.raises(pano.PanoError):
        ctx6.stitch(imgs)
