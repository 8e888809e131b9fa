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
