"""Synthetic panorama workloads (SURVEY.md 8(d) "Concrete inputs"; the reference ships no generator).

One large deterministic scene (seed 20181126) = band-limited fractal noise (5 octaves) + ~200 soft-edged discs and
rectangles per Mpixel with random colours + ~2 % salt texture, uint8.  Views are horizontally translated windows
with 50 % overlap and a per-view jitter of +-8 px vertically, +-0.5 degree rotation and +-3 % gain, so that neighbours
keep well over THRESHOLD = 20 ratio-test matches (ImageProcess.h:18) and non-neighbours (disjoint windows) none.

The scene is defined point-wise from global parameters (noise lattices, shape list, per-column-block salt), so any
window of it can be rendered on its own: a rank of a sharded job renders only the views it owns, and neighbouring
views rendered by different processes still see identical scene pixels.

    views(n, w, h, seed, only=None) -> list of planar uint8 [3][h][w] (None for the views not in `only`)
    pair(index, w, h)               -> two 50 %-overlap views of the scene seeded with `index` (BASELINE.json configs[4])
"""
from __future__ import annotations

import numpy as np

SEED = 20181126
MARGIN = 48      # scene rows above / below the nominal window: room for the vertical jitter and the rotation
XPAD = 24        # scene columns left / right of a view's window: room for the rotation
SALT_BLOCK = 256


class Scene:
    """W x H scene; render(x0, x1) returns planar float32 [3][H][x1 - x0] in 0..255."""

    def __init__(self, W, H, seed=SEED):
        self.W, self.H, self.seed = W, H, seed
        rng = np.random.default_rng(seed)
        self.lattice = [[rng.random((H // (4 << o) + 3, W // (4 << o) + 3), dtype=np.float32) for o in range(5)] for _ in range(3)]
        n = int(200 * W * H / 1e6)
        self.sy = rng.integers(0, H, n); self.sx = rng.integers(0, W, n); self.sr = rng.integers(4, 28, n)
        self.kind = rng.integers(0, 2, n); self.ang = rng.random(n) * np.pi; self.asp = 0.4 + 0.6 * rng.random(n)
        self.col = rng.integers(0, 256, (n, 3)).astype(np.float32)

    def _fractal(self, c, x0, x1):
        H = self.H
        xs = np.arange(x0, x1, dtype=np.float32)
        ys = np.arange(H, dtype=np.float32)
        acc = np.zeros((H, x1 - x0), np.float32)
        for o in range(5):
            s = 4 << o
            g = self.lattice[c][o]
            fx, fy = xs / s, ys / s
            i0, j0 = fx.astype(np.int64), fy.astype(np.int64)
            tx, ty = fx - i0, fy - j0
            gx = g[:, i0] * (1 - tx) + g[:, i0 + 1] * tx                       # [lattice rows][window columns]
            acc += (gx[j0] * (1 - ty)[:, None] + gx[j0 + 1] * ty[:, None]) * (s / 64.0)
        # the five octaves sum to 0 .. 1.9375 and concentrate around the middle: stretch about the mean
        return np.clip((acc / 1.9375 - 0.5) * 2.2 + 0.5, 0.0, 1.0) * 255.0

    def render(self, x0, x1):
        x0, x1 = max(0, x0), min(self.W, x1)
        H = self.H
        sc = np.stack([self._fractal(c, x0, x1) for c in range(3)])
        hit = np.nonzero((self.sx + self.sr + 3 > x0) & (self.sx - self.sr - 3 < x1))[0]
        for k in hit:                                                           # global list order = paint order
            y, x, r = int(self.sy[k]), int(self.sx[k]), int(self.sr[k])
            ya, yb, xa, xb = max(0, y - r - 2), min(H, y + r + 3), max(x0, x - r - 2), min(x1, x + r + 3)
            if ya >= yb or xa >= xb:
                continue
            yy, xx = np.mgrid[ya:yb, xa:xb].astype(np.float32)
            dy, dx = yy - y, xx - x
            if self.kind[k] == 0:
                d = r - np.sqrt(dx * dx + dy * dy)
            else:
                ca, sa = np.float32(np.cos(self.ang[k])), np.float32(np.sin(self.ang[k]))
                d = np.minimum(r - np.abs(dx * ca + dy * sa), np.float32(self.asp[k]) * r - np.abs(-dx * sa + dy * ca))
            alpha = np.clip(d / 2.0 + 0.5, 0.0, 1.0) * np.float32(0.85)         # 2-px soft edge
            reg = sc[:, ya:yb, xa - x0:xb - x0]
            sc[:, ya:yb, xa - x0:xb - x0] = reg * (1 - alpha) + self.col[k][:, None, None] * alpha
        for b in range(x0 // SALT_BLOCK, (x1 - 1) // SALT_BLOCK + 1):           # salt: seeded per column block
            rb = np.random.default_rng([self.seed, 7, b])
            m = rb.random((H, SALT_BLOCK), dtype=np.float32) < 0.02
            v = (rb.random((H, SALT_BLOCK), dtype=np.float32) < 0.5).astype(np.float32) * 255.0
            ba, bb = max(x0, b * SALT_BLOCK), min(x1, (b + 1) * SALT_BLOCK)
            mm, vv = m[:, ba - b * SALT_BLOCK:bb - b * SALT_BLOCK], v[:, ba - b * SALT_BLOCK:bb - b * SALT_BLOCK]
            reg = sc[:, :, ba - x0:bb - x0]
            reg[:, mm] = 0.5 * reg[:, mm] + 0.5 * vv[mm]
        return np.clip(sc, 0, 255), x0


def cut_view(scene, xw, w, h, view_seed):
    """window [xw, xw + w) of the scene with the per-view jitter; bilinear resampling of the rotated window"""
    from scipy import ndimage
    rng = np.random.default_rng(view_seed)
    dy = float(rng.uniform(-8, 8)); rot = np.deg2rad(float(rng.uniform(-0.5, 0.5))); gain = float(rng.uniform(0.97, 1.03))
    sc, xa = scene.render(xw - XPAD, xw + w + XPAD)
    H = sc.shape[1]
    cy, cx = (H - 1) / 2.0 + dy, (xw - xa) + (w - 1) / 2.0
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    yy -= (h - 1) / 2.0; xx -= (w - 1) / 2.0
    sy = cy + yy * np.cos(rot) + xx * np.sin(rot)
    sx = cx - yy * np.sin(rot) + xx * np.cos(rot)
    out = np.empty((3, h, w), np.uint8)
    for c in range(3):
        v = ndimage.map_coordinates(sc[c], [sy, sx], order=1, mode="nearest", prefilter=False)
        out[c] = np.clip(v * gain, 0, 255).astype(np.uint8)
    return out


def views(n, w, h, seed=SEED, only=None):
    """n views w x h, 50 % overlap between neighbours, of one scene; only = indices to render (others are None)"""
    step = w // 2
    scene = Scene(step * (n + 1) + 2 * XPAD, h + 2 * MARGIN, seed)
    want = range(n) if only is None else only
    out = [None] * n
    for i in want:
        out[i] = cut_view(scene, XPAD + i * step, w, h, [seed, 1, i])
    return out


def pair(index, w=1920, h=1080):
    """BASELINE.json configs[4]: independent pair `index` (its own scene), 50 % overlap"""
    a, b = views(2, w, h, SEED + 7919 * (index + 1))
    return a, b
