"""CPU tier: the Reinhard colour transfer's shared per-pixel bodies (csrc/canvas_device.cuh: rgb_to_lab, lab_match,
lab_to_rgb) with serial float plane sums reproduce the reference's `transfer tran(src, tem, out)` (transfer.cpp:4-13,
125-225; oracle/_ref/libpano_ref_transfer.so) bit for bit on the host, where both sides use glibc's logf / pow."""
import os

import numpy as np
import pytest

import emul_api as emul


@pytest.fixture(scope="module")
def tref():
    from oracle import ref_transfer_api
    if not ref_transfer_api.available():
        pytest.skip("oracle/_ref/libpano_ref_transfer.so not built (make -C oracle ref_transfer needs /root/reference)")
    return ref_transfer_api


def test_color_transfer_bodies_equal_reference(tref, input_sets):
    a, b, c = input_sets["Input"][0], input_sets["Input2"][1][:, :300, :400], input_sets["Input"][2]
    black = a.copy()
    black[:, :40, :60] = 0                              # l = m = s = 0 -> the reference substitutes 1 (transfer.cpp:183-185)
    small = np.ascontiguousarray(a[:, :9, :13])         # height <= 16: the reference's non-threaded branch
    for src, tem in ((a, b), (b, a), (black, c), (small, b), (c, small)):
        want = tref.color_transfer(src, tem)
        got = emul.color_transfer(src, tem)
        assert got.shape == want.shape and np.array_equal(got, want)
