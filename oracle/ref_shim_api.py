"""ctypes binding of oracle/_ref/libpano_ref_shim.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

libpano_ref_shim.so is the reference's own ImageProcess.cpp / Projection.cpp / equalization.cpp compiled against the
VLFeat-compatible shim headers (include/vl_b200/compat) and linked against libpano_b200.so INSTEAD of VLFeat
(oracle/Makefile, target ref_shim): the reference's host code, every vl_sift_* / vl_kdforest_* call on the GPU.  It
exports the same harness entry points as libpano_ref.so, so this module is oracle/ref_api.py bound to the other
library.  Needs a CUDA device at run time.
"""
from __future__ import annotations

import importlib.util
import os

from . import ref_api as _ref

SHIM_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libpano_ref_shim.so")


def available() -> bool:
    return os.path.exists(SHIM_SO)


def load():
    """A private copy of the ref_api module whose lib() is libpano_ref_shim.so."""
    spec = importlib.util.spec_from_file_location("oracle._ref_api_on_shim", _ref.__file__)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.REF_SO = SHIM_SO
    return mod
