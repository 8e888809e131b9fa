"""pytest configuration: the `gpu` marker and shared fixtures.

CPU tier  (-m "not gpu"): oracle vs committed golden vectors, the g++ emulation of the product's device arithmetic vs the
compiled reference, host logic, C-ABI load/symbol checks.  GPU tier (-m gpu): the real kernels through the C ABI vs the
compiled reference (oracle/_ref) and the golden vectors.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    from oracle import ref_api
    if not ref_api.available():
        pytest.skip("oracle/_ref/libpano_ref.so not built (make -C oracle ref needs /root/reference)")
    return ref_api


@pytest.fixture(scope="session")
def ctx():
    import computervisionimagestich2_b200 as pano
    c = pano.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def input_sets(ref):
    out = {}
    for name in ("Input", "Input2"):
        out[name] = [ref.load_bmp(os.path.join(ref.REF_DATA, name, f"{i}.bmp")) for i in range(1, 5)]
    return out


def synth_rgb(w, h, seed):
    """Deterministic textured RGB test image (smooth blobs + noise), planar uint8."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (3, h // 8 + 2, w // 8 + 2)).astype(np.float32)
    img = np.kron(img, np.ones((1, 8, 8), np.float32))[:, :h, :w]
    img = 0.7 * img + 0.3 * rng.integers(0, 256, (3, h, w)).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)
