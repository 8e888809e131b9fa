"""CPU tier: the matcher's rigorous uint8 pre-filter (csrc/match_device.cuh, driven stage by stage by tests/emul) gives
exactly the match list of the full float scan -- on real descriptor tables, on adversarial tables (values outside the
quantiser's range, negative values, duplicates, near-ties at the 0.5 ratio), and with a tiny candidate capacity that
forces the overflow path.  The bound itself ((*) in match_device.cuh) is checked on random pairs."""
import os

import numpy as np
import pytest

import emul_api as emul

HERE = os.path.dirname(os.path.abspath(__file__))


def sift_like(rng, n, spread=0.6):
    x = rng.gamma(spread, 1.0, (n, 128)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x = np.minimum(x, 0.2)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.ascontiguousarray(x, np.float32)


def with_matches(rng, A, nb, frac=0.4, noise=0.01):
    """queries: a fraction are noisy copies of rows of A (true matches), the rest unrelated"""
    B = sift_like(rng, nb)
    k = int(frac * nb)
    src = rng.choice(len(A), k, replace=False)
    B[:k] = np.abs(A[src] + rng.normal(0, noise, (k, 128)).astype(np.float32))
    return np.ascontiguousarray(B[rng.permutation(nb)], np.float32)


def check(A, B, cand_cap=0):
    full = emul.match_idx(A, B)
    pre, st = emul.match_idx_prefilter(A, B, cand_cap)
    assert np.array_equal(full, pre), f"pre-filter changed the match list: {np.nonzero(full != pre)[0][:10]} {st}"
    return full, st


def test_prefilter_equals_full_scan_on_sift_like_tables():
    rng = np.random.default_rng(7)
    A = sift_like(rng, 700)
    B = with_matches(rng, A, 500)
    full, st = check(A, B)
    assert (full >= 0).sum() > 100                      # the case is not vacuous
    assert st["survivors"] < 0.6 * len(B)               # ... and the filter does reject the non-matches
    assert st["overflow"] == 0 and st["unbounded_rows"] == 0


def test_prefilter_overflow_path_and_tiny_tables():
    rng = np.random.default_rng(8)
    A = sift_like(rng, 300)
    B = with_matches(rng, A, 200)
    _, st = check(A, B, cand_cap=1)                     # every survivor overflows -> full scan of that query
    assert st["overflow"] == st["survivors"] > 0
    for na, nb in ((2, 5), (3, 1), (1, 4), (0, 3), (5, 0)):
        check(A[:na], B[:nb])


def test_prefilter_adversarial_values():
    rng = np.random.default_rng(9)
    A = sift_like(rng, 400)
    B = with_matches(rng, A, 300)
    # values beyond the quantiser's 0..255/512 range, negative values, exact duplicates, an all-zero row
    A2, B2 = A.copy(), B.copy()
    A2[::7] *= 3.0
    B2[::5] *= 3.0
    A2[3] = -A2[3]
    B2[4, :8] = -0.25
    A2[10] = A2[11]
    B2[20] = A2[10]                                     # d0 == d1 == 0 -> ratio NaN -> rejected
    A2[50] = 0.0
    B2[60] = 0.0
    check(A2, B2)
    # huge magnitudes: the error bound is "unbounded", nothing may be rejected early
    A3 = A.copy()
    A3[5] = 1e30
    _, st = check(A3, B)
    assert st["unbounded_rows"] == 1 and st["survivors"] == len(B)
    # whole tables scaled down (everything quantises to 0/1) and up
    check(np.ascontiguousarray(A * 0.01), np.ascontiguousarray(B * 0.01))
    check(np.ascontiguousarray(A * 1.9), np.ascontiguousarray(B * 1.9))


def test_prefilter_near_the_ratio_threshold():
    """queries constructed so that d0 / d1 straddles 0.5 within a few ulps: the decision must come from the exact pass"""
    rng = np.random.default_rng(10)
    A = sift_like(rng, 200)
    B = []
    for i in range(150):
        a0, a1 = A[rng.choice(200, 2, replace=False)]
        t = 1.0 / 3.0 + rng.normal(0, 1e-4)            # a point on the segment a0..a1 with |b-a0| ~ 0.5 |b-a1|
        B.append((1 - t) * a0 + t * a1)
    B = np.ascontiguousarray(np.array(B, np.float32))
    full, st = check(A, B)
    assert 0 < (full >= 0).sum() < len(B)


def test_bound_holds_on_random_pairs():
    """(SAD - ea - eb)(1 - g) <= 512 d_ref(a, b) <= (SAD + ea + eb)(1 + g) for the reference's float-sequential d_ref"""
    rng = np.random.default_rng(11)
    A = sift_like(rng, 64) * rng.uniform(0.2, 2.5, (64, 1)).astype(np.float32)
    B = sift_like(rng, 64) * rng.uniform(0.2, 2.5, (64, 1)).astype(np.float32)
    S, g = 512.0, 2.0 ** -16

    def quant(x):
        s = x * np.float32(S)
        q = np.clip(np.floor(s + np.float32(0.5)), 0, 255)
        e = np.abs(s.astype(np.float64) - q).sum(1)
        return q.astype(np.int64), np.floor(e * (1 + 1e-6) + 1.0).astype(np.int64) + 1

    qa, ea = quant(A)
    qb, eb = quant(B)
    for b in range(len(B)):
        acc = np.zeros(len(A), np.float32)
        for k in range(128):
            acc = acc + np.abs(B[b, k] - A[:, k]).astype(np.float32)
        d = acc.astype(np.float64) * S
        sad = np.abs(qa - qb[b]).sum(1)
        assert np.all((sad - ea - eb[b]) * (1 - g) <= d) and np.all(d <= (sad + ea + eb[b]) * (1 + g))


def test_prefilter_vs_reference_on_input_sets(ref, input_sets):
    """real SIFT tables of the bundled images: identical to the reference's getImgPair through its kd-forest"""
    tabs = []
    for img in input_sets["Input"][:3]:
        d, k = ref.sift_features(ref.gray(ref.project(img)))
        tabs.append((d, k))
    for i, j in ((0, 1), (1, 0), (1, 2), (0, 2)):
        (da, ka), (db, kb) = tabs[i], tabs[j]
        pre, st = emul.match_idx_prefilter(da, db)
        ra, rb = ref.match(da, ka, db, kb)
        qb = np.nonzero(pre >= 0)[0]
        assert ka[pre[qb]].tobytes() == ra.tobytes() and kb[qb].tobytes() == rb.tobytes()
        assert st["overflow"] == 0 and st["survivors"] <= 2 * len(qb) + 16
