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


# ---- grouped pass (second level: 32 group bytes per row, match_device.cuh (**)) ---------------------------------------
def check_grouped(X, Y, sample_rows=0, cand_cap=0):
    fxy, fyx = emul.match_idx(X, Y), emul.match_idx(Y, X)
    gxy, gyx, st = emul.match_pair_grouped(X, Y, sample_rows, cand_cap)
    assert st["unqualified"] == 0
    assert np.array_equal(fxy, gxy), f"grouped pass changed getImgPair(X, Y): {np.nonzero(fxy != gxy)[0][:10]} {st}"
    assert np.array_equal(fyx, gyx), f"grouped pass changed getImgPair(Y, X): {np.nonzero(fyx != gyx)[0][:10]} {st}"
    return fxy, fyx, st


def test_group_bound_holds():
    """SAD(a, b) >= 4 * S(a, b) - 96 for the 2 x 2-cell / same-bin grouping, on arbitrary byte rows"""
    rng = np.random.default_rng(31)
    q = rng.integers(0, 256, (300, 128)).astype(np.int64)
    q[:40] = rng.integers(0, 4, (40, 128))          # tiny values: the slack dominates
    q[40:60] = 255
    d = np.arange(128)
    grp = (d & 7) + 8 * ((((d >> 3) & 3) >> 1) + 2 * ((d >> 5) >> 1))
    assert sorted(np.bincount(grp)) == [4] * 32
    G = np.zeros((len(q), 32), np.int64)
    for k in range(128):
        G[:, grp[k]] += q[:, k]
    assert G.max() <= 1020
    F = G >> 2
    for a in range(0, len(q), 7):
        sad = np.abs(q - q[a]).sum(1)
        S = np.abs(F - F[a]).sum(1)
        assert np.all(sad >= 4 * S - 96)


def test_grouped_equals_full_scan_both_directions():
    rng = np.random.default_rng(32)
    X = sift_like(rng, 900)
    Y = with_matches(rng, X, 700)
    fxy, fyx, st = check_grouped(X, Y)
    assert (fxy >= 0).sum() > 100 and (fyx >= 0).sum() > 100
    assert st["accepts"] > 0.5 * ((fxy >= 0).sum() + (fyx >= 0).sum())   # most true matches need no float arithmetic at all
    assert st["exact"] < 0.2 * st["pairs"]            # the bound does skip (synthetic tables are far less structured than SIFT)
    # a small sample (loose thresholds) and a tiny candidate capacity (overflow -> full scan) change nothing
    check_grouped(X[:300], Y[:250], sample_rows=4, cand_cap=1)
    # unrelated tables: nothing matches, (almost) nothing survives
    _, _, st = check_grouped(X, sift_like(rng, 400))
    assert st["survivors"] < 0.1 * 1300


def test_grouped_adversarial_and_degenerate():
    rng = np.random.default_rng(33)
    X = sift_like(rng, 400)
    Y = with_matches(rng, X, 300)
    X2, Y2 = X.copy(), Y.copy()
    X2[10] = X2[11]                                   # duplicates: d0 == d1 ties
    Y2[20] = X2[10]
    X2[50] = 0.0
    Y2[60] = 0.0
    Y2[4, :8] = -0.02                                 # negative values are clamped by the quantiser, their error is accounted
    check_grouped(X2, Y2)
    check_grouped(np.ascontiguousarray(X * 0.01), np.ascontiguousarray(Y * 0.01))   # everything quantises to 0 / 1
    for nx, ny in ((2, 1), (2, 2), (3, 130), (65, 64), (129, 2), (1, 5), (5, 1)):
        check_grouped(X[:nx], Y[:ny])
    X3 = X.copy()
    X3[5] *= 3.0                                      # a row whose error bound exceeds the cap: the pair does not qualify
    _, _, st = emul.match_pair_grouped(X3, Y)
    assert st["unqualified"] == 1


def test_grouped_vs_reference_on_input_sets(ref, input_sets):
    tabs = [ref.sift_features(ref.gray(ref.project(img))) for img in input_sets["Input"][:3]]
    for i, j in ((0, 1), (1, 2), (0, 2)):
        (dx, kx), (dy, ky) = tabs[i], tabs[j]
        gxy, gyx, st = emul.match_pair_grouped(dx, dy)
        ra, rb = ref.match(dx, kx, dy, ky)
        q = np.nonzero(gxy >= 0)[0]
        assert kx[gxy[q]].tobytes() == ra.tobytes() and ky[q].tobytes() == rb.tobytes()
        ra, rb = ref.match(dy, ky, dx, kx)
        q = np.nonzero(gyx >= 0)[0]
        assert ky[gyx[q]].tobytes() == ra.tobytes() and kx[q].tobytes() == rb.tobytes()
        assert st["exact"] < 0.02 * st["pairs"]       # real SIFT tables: the grouped bound skips almost every row pair


def test_grouped_near_the_ratio_threshold():
    """d0 / d1 straddles 0.5 (and 0.25, where the certain-accept rule stops being certain): such queries must reach the
    exact pass, and the accept rule must never fire on a query the exact rule rejects"""
    rng = np.random.default_rng(34)
    X = sift_like(rng, 260)
    Y = []
    for i in range(200):
        a0, a1 = X[rng.choice(260, 2, replace=False)]
        t = (1.0 / 3.0 if i % 2 else 0.2) + rng.normal(0, 2e-3)
        Y.append((1 - t) * a0 + t * a1)
    Y = np.ascontiguousarray(np.array(Y, np.float32))
    fxy, fyx, st = check_grouped(X, Y)
    assert 0 < (fxy >= 0).sum() < len(Y)
    assert st["accepts"] < (fxy >= 0).sum() + (fyx >= 0).sum()
