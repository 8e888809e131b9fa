"""GPU tier: the uint8 / tcgen05 matcher against its integer oracle (oracle/match_u8_oracle.py) -- exact."""
import os

import numpy as np
import pytest

from oracle import match_u8_oracle as orc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _sift_like(n, seed):
    """unit-norm non-negative rows, then VLFeat quantisation"""
    rng = np.random.default_rng(seed)
    x = rng.gamma(0.6, 1.0, (n, 128)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x = np.minimum(x, 0.2)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def _check(ctx, A, B):
    idx, d0, d1, near = ctx.match_u8(A, B)
    ridx, rd0, rd1, rnear = orc.match_u8(A, B)
    assert np.array_equal(d0, rd0.astype(np.int64).clip(max=2**31 - 1).astype(np.int32))
    if len(A) >= 2:
        assert np.array_equal(d1, rd1.astype(np.int32))
    assert np.array_equal(idx, ridx)
    uniq = rd0 < rd1          # the nearest row is only defined where it is unique
    assert np.array_equal(near[uniq], rnear[uniq])


@pytest.mark.parametrize("na,nb", [(2, 1), (7, 5), (255, 129), (256, 128), (257, 130), (1024, 300), (3000, 2227), (5000, 17)])
def test_random_tables_ragged_sizes(ctx, na, nb):
    rng = np.random.default_rng(na * 7919 + nb)
    _check(ctx, rng.integers(0, 256, (na, 128), dtype=np.uint8), rng.integers(0, 256, (nb, 128), dtype=np.uint8))


def test_extreme_values_and_duplicates(ctx):
    A = np.zeros((600, 128), np.uint8)
    A[1::2] = 255
    A[100] = np.arange(128, dtype=np.uint8)
    A[101] = A[100]                      # exact duplicate rows: d0 == d1 -> rejected by the ratio rule
    B = np.stack([A[100], np.full(128, 255, np.uint8), np.zeros(128, np.uint8), np.full(128, 128, np.uint8)])
    _check(ctx, A, B)


def test_quantisation_matches_vlfeat_convention(ctx):
    x = _sift_like(1000, 3)
    assert np.array_equal(ctx.quantize_u8(x), orc.quantize_u8(x))


def test_planted_neighbours_are_found(ctx):
    a = _sift_like(4000, 11)
    rng = np.random.default_rng(12)
    pick = rng.permutation(4000)[:1500]
    b = a[pick] + rng.normal(0, 0.01, (1500, 128)).astype(np.float32)
    A, B = orc.quantize_u8(a), orc.quantize_u8(np.maximum(b, 0))
    idx, d0, d1, near = ctx.match_u8(A, B)
    ridx, _, _, _ = orc.match_u8(A, B)
    assert np.array_equal(idx, ridx)
    assert (idx == pick).mean() > 0.9


def test_agreement_with_the_exact_float_l1_matcher_on_input(ctx):
    """Not a parity claim: records how far the north-star quantised-L2 matcher is from the reference's float-L1 one."""
    small = np.load(os.path.join(HERE, "golden", "small.npz"))
    da, ka = ctx.sift_features(small["m_gray_a"])
    db, kb = ctx.sift_features(small["m_gray_b"])
    exact = ctx.match_idx(da, db)
    idx, *_ = ctx.match_u8(ctx.quantize_u8(da), ctx.quantize_u8(db))
    both = (exact >= 0) & (idx >= 0)
    assert both.sum() > 0 and np.array_equal(exact[both], idx[both])   # where both accept, they pick the same row
