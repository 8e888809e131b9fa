"""24-bpp BMP <-> CImg planar RGB (the container step either side of the path; CImg.h:48395-48546 load, :52614 save).

Host-side convenience for the tests and bench.py; the C++ host (host/ImageProcess.hpp) has its own loader.
"""
from __future__ import annotations

import struct

import numpy as np


def load_bmp(path: str) -> np.ndarray:
    """Returns planar uint8 [3][H][W] (R, G, B), top row first."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:2] != b"BM":
        raise ValueError(f"{path}: not a BMP file")
    off = struct.unpack_from("<I", data, 10)[0]
    hdr = struct.unpack_from("<I", data, 14)[0]
    w, h = struct.unpack_from("<ii", data, 18)
    bpp = struct.unpack_from("<H", data, 28)[0]
    comp = struct.unpack_from("<I", data, 30)[0] if hdr >= 40 else 0
    if bpp != 24 or comp != 0:
        raise ValueError(f"{path}: only uncompressed 24-bpp BMP is on the stitching path (bpp={bpp}, compression={comp})")
    bottom_up = h > 0
    h = abs(h)
    stride = (3 * w + 3) & ~3
    rows = np.frombuffer(data, np.uint8, count=stride * h, offset=off).reshape(h, stride)[:, : 3 * w].reshape(h, w, 3)
    if bottom_up:
        rows = rows[::-1]
    return np.ascontiguousarray(rows[:, :, ::-1].transpose(2, 0, 1))


def save_bmp(path: str, img: np.ndarray) -> None:
    img = np.ascontiguousarray(img, np.uint8)
    _, h, w = img.shape
    stride = (3 * w + 3) & ~3
    rows = np.zeros((h, stride), np.uint8)
    rows[:, : 3 * w] = img.transpose(1, 2, 0)[::-1, :, ::-1].reshape(h, 3 * w)
    with open(path, "wb") as f:
        f.write(b"BM" + struct.pack("<IHHI", 54 + stride * h, 0, 0, 54))
        f.write(struct.pack("<IiiHHIIiiII", 40, w, h, 1, 24, 0, stride * h, 2835, 2835, 0, 0))
        f.write(rows.tobytes())
