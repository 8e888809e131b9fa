"""Integer CPU oracle of the uint8 / squared-L2 matcher -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This matcher has no counterpart in the reference (which matches float descriptors under L1, ImageProcess.cpp:273-351);
it is the north-star "int8 GEMM on VLFeat's uint8-quantised descriptors" stage.  Definition (all integer, exact):
    q        = (uint8) min(512 x, 255)           VLFeat's descriptor quantisation (its CLI / MATLAB drivers; float32 math)
    d2(b, a) = sum_k (B[b,k] - A[a,k])^2         = |b|^2 + |a|^2 - 2 b.a
    match    = argmin_a d2, accepted iff 4 * d0 < d1 (d0, d1 the two smallest values; <=> sqrt(d0)/sqrt(d1) < 0.5)
An accepted match has a unique nearest row (d0 < d1), so index tie-breaking cannot matter.
"""
from __future__ import annotations

import numpy as np


def quantize_u8(descr: np.ndarray) -> np.ndarray:
    x = np.float32(512.0) * np.asarray(descr, np.float32)
    x = np.where(x < np.float32(255.0), x, np.float32(255.0))
    return x.astype(np.uint8)   # truncation, as the C cast


def match_u8(A: np.ndarray, B: np.ndarray, block: int = 2048):
    """A [NA][128] u8 database, B [NB][128] u8 queries -> (idx [NB] int32 (-1 = rejected), d0, d1, nearest)."""
    A = np.asarray(A, np.uint8).astype(np.int64)
    B = np.asarray(B, np.uint8).astype(np.int64)
    na = (A * A).sum(1)
    nb = (B * B).sum(1)
    NB = len(B)
    idx = np.full(NB, -1, np.int32)
    d0 = np.zeros(NB, np.int64)
    d1 = np.zeros(NB, np.int64)
    nearest = np.full(NB, -1, np.int32)
    if len(A) == 0:
        return idx, d0, d1, nearest
    for s in range(0, NB, block):
        Bs = B[s:s + block]
        D = nb[s:s + block, None] + na[None, :] - 2 * (Bs @ A.T)
        if len(A) >= 2:
            part = np.partition(D, 1, axis=1)
            d0[s:s + block] = part[:, 0]
            d1[s:s + block] = part[:, 1]
        else:
            d0[s:s + block] = D[:, 0]
            d1[s:s + block] = np.iinfo(np.int32).max
        nearest[s:s + block] = D.argmin(1)
    ok = (len(A) >= 2) & (4 * d0 < d1)
    idx[ok] = nearest[ok]
    return idx, d0, d1, nearest
