"""GPU tier: the matcher's uint8 SAD pre-filter + exact re-rank (match_sad_kernel / match_exact_kernel) returns the
same match list as the full float-L1 scan (match_l1_kernel, the round-1 kernel) and as the reference's getImgPair
(ImageProcess.cpp:273-351) -- bit-identical indices, on real SIFT tables, large synthetic tables, adversarial values
and degenerate sizes.  The CPU tier (tests/test_cpu_prefilter.py) proves the same for the arithmetic itself."""
import os

import numpy as np
import pytest

from test_cpu_prefilter import sift_like, with_matches

pytestmark = pytest.mark.gpu


def both(ctx, A, B):
    ctx.set_match_mode("full")
    full = ctx.match_idx(A, B)
    ctx.set_match_mode("prefilter")
    ctx.match_stats(reset=True)
    pre = ctx.match_idx(A, B)
    st = ctx.match_stats(reset=True)
    assert np.array_equal(full, pre), f"pre-filter changed the match list at {np.nonzero(full != pre)[0][:10]}, {st}"
    return full, st


def test_prefilter_equals_full_scan_large_tables(ctx):
    rng = np.random.default_rng(21)
    A = sift_like(rng, 9001)
    B = with_matches(rng, A, 7003, frac=0.35)
    full, st = both(ctx, A, B)
    assert (full >= 0).sum() > 1500
    assert st["queries"] == len(B) and 0 < st["survivors"] < 0.6 * len(B)
    # no true match between unrelated tables: (almost) everything is rejected by the SAD pass alone
    full, st = both(ctx, A, sift_like(rng, 5000))
    assert st["survivors"] < 0.05 * 5000


def test_prefilter_ragged_and_degenerate_sizes(ctx):
    rng = np.random.default_rng(22)
    A = sift_like(rng, 1500)
    B = with_matches(rng, A, 1100)
    for na, nb in ((2, 1), (2, 300), (3, 257), (65, 256), (127, 1), (129, 513), (1500, 1), (1, 50)):
        both(ctx, A[:na], B[:nb])


def test_prefilter_adversarial_values(ctx):
    rng = np.random.default_rng(23)
    A = sift_like(rng, 2000)
    B = with_matches(rng, A, 1500)
    A2, B2 = A.copy(), B.copy()
    A2[::7] *= 3.0
    B2[::5] *= 3.0
    A2[3] = -A2[3]
    B2[4, :8] = -0.25
    A2[10] = A2[11]
    B2[20] = A2[10]
    A2[50] = 0.0
    B2[60] = 0.0
    both(ctx, A2, B2)
    A3 = A.copy()
    A3[5] = 1e30                                     # unbounded row: every query survives, lists overflow -> full scan
    _, st = both(ctx, A3, B[:300])
    assert st["survivors"] == 300
    both(ctx, np.ascontiguousarray(A * 0.01), np.ascontiguousarray(B * 0.01))   # everything quantises to 0 / 1:
    both(ctx, np.ascontiguousarray(A * 1.9), np.ascontiguousarray(B * 1.9))     # ... long candidate lists, overflow


def test_prefilter_vs_reference_getimgpair(ctx, ref, input_sets):
    ctx.set_match_mode("prefilter")
    tabs = [ref.sift_features(ref.gray(ref.project(img))) for img in input_sets["Input2"][:3]]
    for i, j in ((0, 1), (1, 0), (1, 2), (2, 0)):
        (da, ka), (db, kb) = tabs[i], tabs[j]
        ga, gb = ctx.match(da, ka, db, kb)
        ra, rb = ref.match(da, ka, db, kb)
        assert ga.tobytes() == ra.tobytes() and gb.tobytes() == rb.tobytes()


# ---- both directions of an image pair from one pass over the SAD matrix (match_sad_sym_kernel) --------------------------
def both_pair(ctx, A, B, expect_sym=True):
    ctx.set_match_mode("full")
    fab, fba = ctx.match_idx(A, B), ctx.match_idx(B, A)
    ctx.set_match_mode("prefilter")
    ctx.match_stats(reset=True)
    ab, ba = ctx.match_pair(A, B)
    st = ctx.match_stats(reset=True)
    assert np.array_equal(fab, ab), f"A->B list differs at {np.nonzero(fab != ab)[0][:10]}, {st}"
    assert np.array_equal(fba, ba), f"B->A list differs at {np.nonzero(fba != ba)[0][:10]}, {st}"
    if expect_sym is not None and min(len(A), len(B)) >= 2:
        assert st["sym_pairs"] == (1 if expect_sym else 0), st
        # default mode: symmetric pairs take the grouped pass (or, when its pair queue overflows, are redone without it)
        assert st["group_pairs"] + st["group_overflow"] == st["sym_pairs"], st
    for mode in ("prefilter_onedir", "prefilter_fullsad"):   # one SAD pass per direction / full 128-byte SAD for both
        ctx.set_match_mode(mode)
        ab1, ba1 = ctx.match_pair(A, B)
        assert np.array_equal(fab, ab1) and np.array_equal(fba, ba1), mode
    ctx.set_match_mode("prefilter")
    return ab, ba, st


def test_symmetric_pass_equals_full_scan(ctx):
    rng = np.random.default_rng(31)
    A = sift_like(rng, 6007)
    B = with_matches(rng, A, 5003, frac=0.4)
    ab, ba, st = both_pair(ctx, A, B)
    assert (ab >= 0).sum() > 1000 and (ba >= 0).sum() > 1000
    assert st["queries"] == len(A) + len(B) and st["overflow"] <= 0.01 * st["queries"]
    # unstructured random tables: the grouped bound skips far fewer pairs than on SIFT tables, so the pair queue (sized for
    # 0.5 % of the pairs) overflows and the batch is redone with the full SAD pass -- the fallback is part of the contract
    assert st["group_pairs"] + st["group_overflow"] == 1, st
    both_pair(ctx, A, sift_like(rng, 3001))            # unrelated tables: nothing to match in either direction


def test_symmetric_pass_ragged_sizes(ctx):
    """odd tile tails, partial last blocks of held rows, single-row splits, tiny tables"""
    rng = np.random.default_rng(32)
    A = sift_like(rng, 1400)
    B = with_matches(rng, A, 1300)
    for na, nb in ((2, 2), (2, 301), (3, 257), (65, 256), (127, 2), (129, 513), (1400, 3), (1, 50), (50, 1), (1399, 1299), (64, 64)):
        both_pair(ctx, A[:na], B[:nb], expect_sym=min(na, nb) >= 2)


def test_symmetric_pass_adversarial_and_fallback(ctx):
    rng = np.random.default_rng(33)
    A = sift_like(rng, 1800)
    B = with_matches(rng, A, 1500)
    A2, B2 = A.copy(), B.copy()
    A2[3] = -A2[3]                                     # error bound of this row is large but finite: still <= cap? no ->
    B2[20] = A2[10] = A2[11]                           # duplicates: d0 == d1 == 0
    A2[50] = 0.0
    B2[60] = 0.0
    both_pair(ctx, A2, B2, expect_sym=False)           # the negated row's error bound exceeds the 16-bit cap: one-directional passes
    A3, B3 = A.copy(), B.copy()
    B3[20] = A3[10] = A3[11]
    A3[50] = 0.0
    B3[60] = 0.0
    both_pair(ctx, A3, B3, expect_sym=True)
    _, _, st = both_pair(ctx, np.ascontiguousarray(A * 0.01), np.ascontiguousarray(B * 0.01))   # everything quantises to 0 / 1:
    assert st["group_overflow"] == 1, st                                                       # ... the bound skips nothing -> redo
    both_pair(ctx, np.ascontiguousarray(A * 1.9), np.ascontiguousarray(B * 1.9), expect_sym=None)    # clamped rows: large errors


def test_symmetric_pass_vs_reference_getimgpair(ctx, ref, input_sets):
    """real SIFT tables: the grouped pass proper (no queue overflow), few exact SADs, most matches accepted without float
    arithmetic -- and the reference's match lists bit for bit"""
    ctx.set_match_mode("prefilter")
    tabs = [ref.sift_features(ref.gray(ref.project(img))) for img in input_sets["Input2"][:3]]
    for i, j in ((0, 1), (1, 2), (0, 2)):
        (da, ka), (db, kb) = tabs[i], tabs[j]
        ctx.match_stats(reset=True)
        ab, ba = ctx.match_pair(da, db)
        st = ctx.match_stats(reset=True)
        assert st["group_pairs"] == 1 and st["group_overflow"] == 0, st
        assert 0 < st["group_exact"] < 0.01 * len(da) * len(db), st
        assert st["group_accepts"] <= (ab >= 0).sum() + (ba >= 0).sum(), st
        ra, rb = ref.match(da, ka, db, kb)
        sel = ab >= 0
        assert ka[ab[sel]].tobytes() == ra.tobytes() and kb[sel].tobytes() == rb.tobytes()
        ra, rb = ref.match(db, kb, da, ka)
        sel = ba >= 0
        assert kb[ba[sel]].tobytes() == ra.tobytes() and ka[sel].tobytes() == rb.tobytes()


def test_grouped_pass_dense_queue(ctx):
    """clustered tables: every query fails the grouped bound against the ~50 rows of its own cluster, so a CTA collects
    tens of thousands of pairs and flushes its shared-memory queue many times inside the pass (and falls through to the
    direct global append when a tile adds more than the local buffer holds); the lists must still equal the full scan,
    run after run"""
    rng = np.random.default_rng(35)
    centres = sift_like(rng, 40)
    X = np.repeat(centres, 50, axis=0) + rng.normal(0, 2e-4, (2000, 128)).astype(np.float32)
    Y = np.repeat(centres, 45, axis=0) + rng.normal(0, 2e-4, (1800, 128)).astype(np.float32)
    X = np.ascontiguousarray(np.abs(X), np.float32)
    Y = np.ascontiguousarray(np.abs(Y[rng.permutation(len(Y))]), np.float32)
    ab, ba, st = both_pair(ctx, X, Y)
    assert st["group_pairs"] == 1 and st["group_overflow"] == 0, st
    assert st["group_exact"] >= 40 * 50 * 45, st
    for _ in range(5):
        ab2, ba2 = ctx.match_pair(X, Y)
        assert np.array_equal(ab, ab2) and np.array_equal(ba, ba2)
