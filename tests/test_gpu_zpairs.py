"""GPU tier: the batched-pairs job (BASELINE.json configs[4]; dist.pairs_batch with the CUDA Context as the engine)
against the compiled reference on the same pairs: feature counts, match counts and the RANSAC coefficients of every
adjacent direction, bit for bit.  Reference: ImageProcess.cpp:12-23 (readFile), 117-137 / 273-351 (getImgPair),
395-436 (RANSAC)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
PAIR_LIST = [(0, 1), (1, 2), (2, 3), (0, 3), (3, 2)]


def test_pairs_batch_matches_reference(ctx, ref):
    from computervisionimagestich2_b200 import dist as pdist
    ctx.set_profile("root", 666666)
    imgs = [ref.load_bmp(os.path.join(ref.REF_DATA, "Input", f"{i}.bmp")) for i in range(1, 5)]
    table = pdist.pairs_batch(ctx, [(imgs[a], imgs[b]) for a, b in PAIR_LIST], chunk=3)
    # the batched entry (pano_b200_pairs) and the composition of single-stage calls give the same table
    assert pdist.pairs_by_calls(ctx, [(imgs[a], imgs[b]) for a, b in PAIR_LIST]).tobytes() == table.tobytes()
    anchors = json.load(open(os.path.join(HERE, "golden", "anchors.json")))["Input"]
    feats = [ref.sift_features(ref.gray(ref.project(im))) for im in imgs]
    assert [int(p) for p in table["pair"]] == list(range(len(PAIR_LIST)))
    for rec, (a, b) in zip(table, PAIR_LIST):
        assert list(rec["nfeat"]) == [anchors["nfeat"][a], anchors["nfeat"][b]]
        assert list(rec["nmatch"]) == [anchors["match_counts"][f"{a},{b}"], anchors["match_counts"][f"{b},{a}"]]
        for d, (i, j) in enumerate(((a, b), (b, a))):
            assert bool(rec["has_h"][d]) == (rec["nmatch"][d] >= 20)
            if rec["has_h"][d]:
                src, dst = ref.match(feats[i][0], feats[i][1], feats[j][0], feats[j][1])
                assert ref.ransac(src, dst).tobytes() == rec["H"][d].tobytes(), (a, b, d)


def test_pair_job_degenerate_images(ctx, ref):
    """Ragged / empty inputs.  A flat image still yields the reference's single feature (the black border the
    cylindrical projection leaves is an edge); an all-zero image yields none.  Neither reaches the 2-NN query, which
    needs two database rows (ImageProcess.cpp:327), nor RANSAC, which needs four pairs (Q8): no match, no fit, no
    error."""
    from computervisionimagestich2_b200 import dist as pdist
    for value in (128, 0):
        img = np.full((3, 96, 128), value, np.uint8)
        want = len(ref.sift_features(ref.gray(ref.project(img)))[1])
        assert want == (1 if value else 0)
        for rec in (pdist.pair_job(ctx, img, img), ctx.pairs([(img, img)])[0]):
            assert list(rec["nfeat"]) == [want, want] and list(rec["nmatch"]) == [0, 0] and not rec["has_h"].any()
    assert len(ctx.pairs([])) == 0


def parse_pairs_main(text):
    """lines of examples/pairs_main: 'a b nfeat N matches M [8 coefficients]' -> {(a, b): (N, M, H or None)}"""
    out = {}
    for line in text.splitlines():
        t = line.split()
        assert t[2] == "nfeat" and t[4] == "matches" and len(t) in (6, 14), line
        out[(int(t[0]), int(t[1]))] = (int(t[3]), int(t[5]), np.array([float(v) for v in t[6:]]) if len(t) == 14 else None)
    return out


def test_cpp_host_pairs_main(ctx, ref):
    """examples/pairs_main.cpp (C++ host, one pano_b200_pairs call for the chain neighbours of Input) prints what the
    Python binding returns for the same pairs, coefficients bit for bit (%.17g round-trips a double)."""
    import subprocess
    from computervisionimagestich2_b200 import dist as pdist
    d = os.path.join(ref.REF_DATA, "Input")
    exe = os.path.join(os.path.dirname(HERE), "examples", "pairs_main")
    subprocess.run(["make", "-C", os.path.dirname(exe)], check=True, stdout=subprocess.DEVNULL)
    got = parse_pairs_main(subprocess.run([exe, d + "/", "4"], capture_output=True, text=True, check=True).stdout)
    imgs = [ref.load_bmp(os.path.join(d, f"{i}.bmp")) for i in range(1, 5)]
    ctx.set_profile("root", 666666)
    table = ctx.pairs([(imgs[p], imgs[p + 1]) for p in range(3)])
    assert sorted(got) == sorted([(p, p + 1) for p in range(3)] + [(p + 1, p) for p in range(3)])
    for p in range(3):
        for dname, (a, b) in enumerate(((p, p + 1), (p + 1, p))):
            n, m, H = got[(a, b)]
            assert n == table[p]["nfeat"][dname] and m == table[p]["nmatch"][dname]
            assert (H is not None) == bool(table[p]["has_h"][dname])
            if H is not None:
                assert H.tobytes() == table[p]["H"][dname].tobytes()
