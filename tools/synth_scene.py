"""Synthetic panorama workloads (SURVEY.md 8(d) "Concrete inputs"; the reference ships no generator).

One large deterministic scene (seed 20181126) = band-limited fractal noise (5 octaves) + ~200 soft-edged discs and
quadrilaterals per Mpixel with random colours + ~2 % salt texture, uint8.  Views are horizontally translated windows
with 50 % overlap and a per-view jitter of +-8 px vertically, +-0.5 degree rotation and +-3 % gain, so that neighbours
keep well over THRESHOLD = 20 ratio-test matches (ImageProcess.h:18) and non-neighbours (disjoint windows) none.

    views(n, w, h, seed)          -> list of planar uint8 [3][h][w]
    pair(index, w, h)             -> two 50 %-overlap views of the scene seeded with `index` (BASELINE.json configs[4])
"""
from __future__ import annotations

import numpy as np

SEED = 20181126
MARGIN = 48   # scene rows above / below the nominal window: room for the vertical jitter and the rotation


def _fractal(rng, H, W):
    from scipy import ndimage
    acc = np.zeros((H, W), np.float32)
    for octave in range(5):
        s = 4 << octave                       # 4 .. 64 px features
        g = rng.random((H // s + 3, W // s + 3)).astype(np.float32)
        up = ndimage.zoom(g, s, order=1, prefilter=False)[:H, :W]
        acc += up * (s / 64.0)
    acc -= acc.min()
    acc *= 255.0 / max(float(acc.max()), 1e-6)
    return acc


def scene(W, H, seed=SEED):
    """planar float32 [3][H][W] in 0..255"""
    rng = np.random.default_rng(seed)
    sc = np.stack([_fractal(rng, H, W) for _ in range(3)])
    nshape = int(200 * W * H / 1e6)
    ys = rng.integers(0, H, nshape); xs = rng.integers(0, W, nshape); rs = rng.integers(4, 28, nshape)
    kinds = rng.integers(0, 2, nshape); angs = rng.random(nshape) * np.pi; asp = 0.4 + 0.6 * rng.random(nshape)
    cols = rng.integers(0, 256, (nshape, 3)).astype(np.float32)
    for y, x, r, kind, ang, a, c in zip(ys, xs, rs, kinds, angs, asp, cols):
        y0, y1, x0, x1 = max(0, y - r - 2), min(H, y + r + 3), max(0, x - r - 2), min(W, x + r + 3)
        yy, xx = np.mgrid[y0:y1, x0:x1].astype(np.float32)
        dy, dx = yy - y, xx - x
        if kind == 0:      # disc
            d = r - np.sqrt(dx * dx + dy * dy)
        else:              # rotated rectangle (signed distance to the nearest edge)
            u = dx * np.cos(ang) + dy * np.sin(ang)
            v = -dx * np.sin(ang) + dy * np.cos(ang)
            d = np.minimum(r - np.abs(u), a * r - np.abs(v))
        alpha = np.clip(d / 2.0 + 0.5, 0.0, 1.0) * 0.85          # 2-px soft edge
        sc[:, y0:y1, x0:x1] = sc[:, y0:y1, x0:x1] * (1 - alpha) + c[:, None, None] * alpha
    salt = rng.random((H, W)) < 0.02
    sv = rng.integers(0, 2, (H, W)).astype(np.float32) * 255.0
    sc[:, salt] = 0.5 * sc[:, salt] + 0.5 * sv[salt]
    return np.clip(sc, 0, 255)


def cut_view(sc, x0, w, h, rng):
    """window [x0, x0 + w) of the scene with the per-view jitter; bilinear resampling of the rotated window"""
    from scipy import ndimage
    H = sc.shape[1]
    dy = float(rng.uniform(-8, 8)); rot = np.deg2rad(float(rng.uniform(-0.5, 0.5))); gain = float(rng.uniform(0.97, 1.03))
    cy, cx = (H - 1) / 2.0 + dy, x0 + (w - 1) / 2.0
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    yy -= (h - 1) / 2.0; xx -= (w - 1) / 2.0
    sy = cy + yy * np.cos(rot) + xx * np.sin(rot)
    sx = cx - yy * np.sin(rot) + xx * np.cos(rot)
    out = np.empty((3, h, w), np.uint8)
    for c in range(3):
        v = ndimage.map_coordinates(sc[c], [sy, sx], order=1, mode="nearest", prefilter=False)
        out[c] = np.clip(v * gain, 0, 255).astype(np.uint8)
    return out


def views(n, w, h, seed=SEED):
    """n views w x h, 50 % overlap between neighbours, of one scene"""
    step = w // 2
    W = step * (n + 1) + 2 * MARGIN
    sc = scene(W, h + 2 * MARGIN, seed)
    rng = np.random.default_rng(seed + 1)
    return [cut_view(sc, MARGIN + i * step, w, h, rng) for i in range(n)]


def pair(index, w=1920, h=1080):
    """BASELINE.json configs[4]: independent pair `index` (seed = SEED + pair index), 50 % overlap"""
    a, b = views(2, w, h, SEED + 7919 * (index + 1))
    return a, b
