"""GPU tier: the drop-in boundary exercised by the reference's OWN source.  oracle/_ref/libpano_ref_shim.so is the
reference's ImageProcess.cpp / Projection.cpp / equalization.cpp compiled against include/vl_b200/compat (the
VLFeat-compatible shim headers) and linked against libpano_b200.so instead of VLFeat: its siftAlgorithm and getImgPair
run unchanged and every vl_sift_* / vl_kdforest_* call lands on the GPU.  Results must equal the all-CPU reference."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def shim(ctx):   # ctx: makes sure a CUDA device is there before the shim library is loaded
    from oracle import ref_shim_api
    if not ref_shim_api.available():
        pytest.skip("oracle/_ref/libpano_ref_shim.so not built (make -C oracle ref_shim needs /root/reference)")
    return ref_shim_api.load()


@pytest.fixture(scope="module")
def anchors():
    return json.load(open(os.path.join(HERE, "golden", "anchors.json")))


def _tables(seed, n, nq, dim=128):
    rng = np.random.default_rng(seed)
    # SIFT-like: non-negative, heavy-tailed, normalised and clamped
    d = rng.gamma(0.6, 1.0, (n + nq, dim)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = np.minimum(d, 0.2).astype(np.float32)
    return d[:n], d[n:]


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("k", [1, 2, 5])
@pytest.mark.parametrize("per_query", [True, False])
def test_kdforest_equals_vlfeat(shim, ref, metric, k, per_query):
    data, queries = _tables(metric * 10 + k, 777, 61)
    gi, gd = shim.kdforest_query(data, queries, k, metric, per_query)
    wi, wd = ref.kdforest_query(data, queries, k, metric, per_query)
    assert np.array_equal(gd.view(np.uint64), wd.view(np.uint64))   # distances: VLFeat's float accumulation, bit for bit
    assert np.array_equal(gi, wi)


def test_kdforest_other_dimensions_and_tiny_tables(shim, ref):
    for n, nq, dim, k in [(2, 5, 128, 2), (130, 3, 7, 2), (1000, 40, 64, 8), (300, 300, 33, 4)]:
        data, queries = _tables(n + dim, n, nq, dim)
        gi, gd = shim.kdforest_query(data, queries, k, 0, False)
        wi, wd = ref.kdforest_query(data, queries, k, 0, False)
        assert np.array_equal(gd.view(np.uint64), wd.view(np.uint64)) and np.array_equal(gi, wi), (n, nq, dim, k)


def test_kdforest_fewer_points_than_neighbours(shim):
    """One data point, two neighbours asked for (getImgPair on a 1-feature image reads an unset second neighbour in
    VLFeat): the shim reports the missing neighbour as index -1 / distance NaN instead of leaving it undefined."""
    data, queries = _tables(3, 1, 4)
    for per_query in (True, False):
        gi, gd = shim.kdforest_query(data, queries, 2, 0, per_query)
        assert np.all(gi[:, 0] == 0) and np.all(gi[:, 1] == -1)
        assert np.all(np.isfinite(gd[:, 0])) and np.all(np.isnan(gd[:, 1]))
        want = np.abs(queries.astype(np.float32) - data[0]).astype(np.float32)
        acc = np.zeros(len(queries), np.float32)
        for d in range(128):
            acc = (acc + want[:, d]).astype(np.float32)
        assert np.array_equal(gd[:, 0].astype(np.float32).view(np.uint32), acc.view(np.uint32))


def test_reference_sift_and_matching_on_the_shim(shim, ref, input_sets):
    imgs = input_sets["Input"][2:4]
    g = [ref.gray(ref.project(im)) for im in imgs]
    fs = [shim.sift_features(x) for x in g]      # the reference's siftAlgorithm, vl_sift_* on the GPU
    fr = [ref.sift_features(x) for x in g]
    for (ds, ks), (dr, kr) in zip(fs, fr):
        assert ds.tobytes() == dr.tobytes() and ks.tobytes() == kr.tobytes()
    ms = shim.match(fs[0][0], fs[0][1], fs[1][0], fs[1][1])   # the reference's getImgPair, vl_kdforest_* on the GPU
    mr = ref.match(fr[0][0], fr[0][1], fr[1][0], fr[1][1])
    assert ms[0].tobytes() == mr[0].tobytes() and ms[1].tobytes() == mr[1].tobytes()


def test_reference_pipeline_on_the_shim(shim, input_sets, anchors):
    pano, info = shim.stitch_mem(input_sets["Input"])
    a = anchors["Input"]
    assert info["nfeat"] == a["nfeat"]
    assert sha(pano) == a["pano_sha256"]
