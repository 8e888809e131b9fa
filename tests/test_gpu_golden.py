"""GPU tier: the CUDA path through the C ABI against the COMMITTED golden fixtures (tests/golden) -- needs neither
/root/reference nor oracle/_ref at run time.  Bit-exact."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import synth_rgb

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a.view(np.uint64)


@pytest.fixture(scope="module")
def anchors():
    return json.load(open(os.path.join(HERE, "golden", "anchors.json")))


@pytest.fixture(scope="module")
def small():
    return np.load(os.path.join(HERE, "golden", "small.npz"))


def _load_set(name):
    from computervisionimagestich2_b200 import bmpio
    d = os.path.join(ROOT, "oracle", "_ref", "data", name)
    if not os.path.isdir(d):
        pytest.skip(f"{d} not staged (make -C oracle ref)")
    return [bmpio.load_bmp(os.path.join(d, f"{i}.bmp")) for i in range(1, 5)]


@pytest.mark.parametrize("tag", ["s97x61", "s160x120", "s33x200"])
def test_small_cases(ctx, small, tag):
    img = small[f"{tag}_rgb"]
    assert np.array_equal(ctx.project(img), small[f"{tag}_project"])
    g = ctx.gray(img)
    assert np.array_equal(g, small[f"{tag}_gray"])
    d, k = ctx.sift_features(g)
    assert d.tobytes() == small[f"{tag}_descr"].tobytes() and k.tobytes() == small[f"{tag}_keys"].tobytes()


def test_small_match_ransac(ctx, small, anchors):
    da, ka = ctx.sift_features(small["m_gray_a"])
    db, kb = ctx.sift_features(small["m_gray_b"])
    ma, mb = ctx.match(da, ka, db, kb)
    assert len(ma) == anchors["small_match_count"]
    assert ma.tobytes() == small["m_keys_a"].tobytes() and mb.tobytes() == small["m_keys_b"].tobytes()
    assert np.array_equal(_bits(ctx.ransac(ma, mb)), _bits(small["m_ransac_H"]))


def test_canvas_primitives(ctx, small):
    p = small["cimg_in"]
    assert ctx.cimg_blur2(p).tobytes() == small["cimg_blur2"].tobytes()
    assert ctx.cimg_resize3(p, 32, 24).tobytes() == small["cimg_half"].tobytes()
    assert ctx.cimg_resize3(p, 129, 96).tobytes() == small["cimg_up"].tobytes()
    bl = ctx.blend(small["blend_a"], small["blend_b"])
    assert np.array_equal(bl, small["blend_out"])
    assert np.array_equal(ctx.equalize_mix(bl), small["tail_out"])
    t = synth_rgb(150, 90, 31337)
    a, _ = ctx.warp_shift(t, small["warp_H"], -12.0, -3.0, None, 0, 0, 170, 100)
    assert np.array_equal(a, small["warp_out"])


def test_blend_empty_middle_row_is_an_error(ctx):
    import computervisionimagestich2_b200 as pano
    a = np.zeros((3, 40, 60), np.uint8)
    with pytest.raises(pano.PanoError):
        ctx.blend(a, a)


@pytest.mark.parametrize("name", ["Input", "Input2"])
def test_bundled_sets_stage_hashes_and_panorama(ctx, anchors, small, name):
    A = anchors[name]
    imgs = _load_set(name)
    feats = []
    for i, img in enumerate(imgs):
        assert sha(img) == A["bmp_sha256"][i]
        p, g = ctx.project(img, want_gray=True)
        assert sha(p) == A["project_sha256"][i] and sha(g) == A["gray_sha256"][i]
        d, k = ctx.sift_features(g)
        assert len(k) == A["nfeat"][i]
        assert sha(d) == A["descr_sha256"][i] and sha(k) == A["keys_sha256"][i]
        feats.append((d, k))
    for i in range(4):
        for j in range(4):
            if i == j:
                continue
            a, b = ctx.match(feats[i][0], feats[i][1], feats[j][0], feats[j][1])
            assert len(a) == A["match_counts"][f"{i},{j}"], (i, j)
            if f"{name}_match_{i}_{j}_a" in small:
                assert a.tobytes() == small[f"{name}_match_{i}_{j}_a"].tobytes()
                assert b.tobytes() == small[f"{name}_match_{i}_{j}_b"].tobytes()
    a, b = small[f"{name}_pairs23_src"], small[f"{name}_pairs23_dst"]
    assert [float(x).hex() for x in ctx.ransac(a, b)] == A["ransac_2to3_backward"]
    assert [float(x).hex() for x in ctx.ransac(b, a)] == A["ransac_3to2_forward"]
    pano, info = ctx.stitch(imgs)
    assert info["log"] == A["log"] and info["nfeat"] == A["nfeat_after_stitch"]
    assert list(pano.shape) == A["pano_shape"]
    assert sha(pano) == A["pano_sha256"]


def test_repeated_stitch_is_deterministic(ctx, anchors):
    """Buffers are recycled between jobs and kernels use atomics only for integer counters: same bytes every time."""
    imgs = _load_set("Input")
    h = {sha(ctx.stitch(imgs)[0]) for _ in range(3)}
    assert h == {anchors["Input"]["pano_sha256"]}


# ---- the reference-side bindings: C++ host programs linked against libpano_b200.so -----------------
def _fnv(b):
    h = 0xCBF29CE484222325
    for x in b:
        h = ((h ^ x) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def _example(name):
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, stdout=subprocess.DEVNULL)
    return os.path.join(ROOT, "examples", name)


def test_vl_sift_shim_runs_the_reference_call_sequence(small, tmp_path):
    """examples/sift_features.cpp is ImageProcess::siftAlgorithm compiled against include/vl_b200/sift.h."""
    import subprocess
    exe = _example("sift_features")
    for tag in ("s160x120", "s97x61"):
        g = small[f"{tag}_gray"]
        raw = tmp_path / f"{tag}.raw"
        raw.write_bytes(np.ascontiguousarray(g).tobytes())
        out = subprocess.run([exe, str(raw), str(g.shape[1]), str(g.shape[0])], capture_output=True, text=True, check=True).stdout.split()
        d, k = small[f"{tag}_descr"], small[f"{tag}_keys"]
        assert int(out[0]) == len(k)
        assert out[1] == _fnv(d.tobytes()) and out[2] == _fnv(k.tobytes())


def test_imageprocess_drop_in_main(anchors):
    """examples/main.cpp is the reference's main.cpp on include/pano_b200/ImageProcess.h."""
    import subprocess
    d = os.path.join(ROOT, "oracle", "_ref", "data", "Input")
    if not os.path.isdir(d):
        pytest.skip("bundled inputs not staged")
    out = subprocess.run([_example("pano_main"), d + "/", "4"], capture_output=True, text=True, check=True).stdout
    A = anchors["Input"]
    assert out.splitlines()[:-1] == A["log"].splitlines()
    assert out.splitlines()[-1] == f"panorama {A['pano_shape'][2]}x{A['pano_shape'][1]} fnv1a64 {A['pano_fnv1a64']}"


def test_sharded_job_single_rank_equals_the_fused_pipeline(ctx, anchors):
    """dist.stitch_sharded with one rank: extract -> preset match lists -> pano_b200_stitch_features."""
    from computervisionimagestich2_b200 import dist as pdist
    imgs = _load_set("Input")
    pano, info = pdist.stitch_sharded(ctx, imgs)
    A = anchors["Input"]
    assert info["nfeat"][:0] == [] and info["log"] == A["log"]
    assert list(pano.shape) == A["pano_shape"] and sha(pano) == A["pano_sha256"]


def test_bmp_files_in_bmp_file_out(ctx, anchors, tmp_path):
    """pano_b200_stitch_bmp: GPU BMP decode -> pipeline -> GPU BMP encode; the encoded file decodes to the golden panorama"""
    from computervisionimagestich2_b200 import bmpio
    d = os.path.join(ROOT, "oracle", "_ref", "data", "Input")
    if not os.path.isdir(d):
        pytest.skip("bundled inputs not staged")
    data = ctx.stitch_bmp_files([os.path.join(d, f"{i}.bmp") for i in range(1, 5)])
    p = tmp_path / "pano.bmp"
    p.write_bytes(data)
    pano = bmpio.load_bmp(str(p))
    assert list(pano.shape) == anchors["Input"]["pano_shape"] and sha(pano) == anchors["Input"]["pano_sha256"]
    # and the encoder agrees byte for byte with the host encoder
    q = tmp_path / "host.bmp"
    bmpio.save_bmp(str(q), pano)
    assert q.read_bytes() == data
