"""Build libpano_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m computervisionimagestich2_b200.build [--force] [--verbose]

The library is self-contained (cudart linked statically): no torch types or symbols.  -fmad=false and
-ffp-contract=off are part of the numerical contract (bit-exact parity with the CPU reference needs separate
multiply / add roundings), not tuning knobs.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpano_b200.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["sift_kernels.cu", "sift_engine.cu", "match_kernels.cu", "canvas_kernels.cu", "stitcher.cu", "c_api.cu",
           "vl_sift_shim.cu", "vl_kdforest_shim.cu", "match_i8_kernels.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
EXTRA = os.environ.get("PANO_B200_NVCC_EXTRA", "").split()   # e.g. -DPB_DESCR_DEBUG for an instrumented build
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-diag-suppress", "550"]


def _newest_dep() -> float:
    t = os.path.getmtime(__file__)
    for d in (CSRC, os.path.join(HERE, "..", "include"), os.path.join(HERE, "..", "include", "vl_b200")):
        for f in os.listdir(d):
            if f.endswith((".cu", ".cuh", ".h")):
                t = max(t, os.path.getmtime(os.path.join(d, f)))
    return t


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = list(SOURCES)
    missing = [s for s in srcs if not os.path.exists(os.path.join(CSRC, s))]
    if missing:
        raise RuntimeError(f"listed kernel sources are missing from {CSRC}: {missing}")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_dep():
        return OUT
    os.makedirs(OBJ, exist_ok=True)

    def cc(src: str) -> str:
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, *EXTRA, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(cc, srcs))
    cmd = [NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
