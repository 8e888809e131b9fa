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
