"""CPU tier: the C-ABI library builds, loads without a GPU and exports every symbol include/*.h declares; struct
layouts match the reference's (vl/sift.h:19-78); the product fails loudly (no CPU fallback) when no device is visible."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def declared_functions(path):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"static\s+inline[^{;]*\{[^}]*\}", "", src, flags=re.S)  # inline getters are not exports
    return sorted(set(re.findall(r"\b((?:pano_b200|vl_sift|vl_b200|vl_kdforest|vl_kdforestsearcher)_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def so():
    from computervisionimagestich2_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(so):
    lib = C.CDLL(so)
    missing = []
    n = 0
    for hdr in ("pano_b200.h", os.path.join("vl_b200", "sift.h"), os.path.join("vl_b200", "kdtree.h")):
        p = os.path.join(INC, hdr)
        if not os.path.exists(p):
            continue
        for fn in declared_functions(p):
            n += 1
            if not hasattr(lib, fn):
                missing.append(fn)
    assert n >= 30
    assert not missing, f"declared in include/ but not exported: {missing}"


def test_no_torch_or_oracle_dependency(so):
    out = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "torch" not in out and "pano_ref" not in out and "pano_emul" not in out
    syms = subprocess.run(["nm", "-D", "--undefined-only", so], capture_output=True, text=True).stdout
    assert "ref_" not in syms and "emul_" not in syms


def test_headers_compile_as_c_and_layouts_match_vlfeat(tmp_path):
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "pano_b200.h"
#include "vl_b200/sift.h"
#include "vl_b200/kdtree.h"
int main(void) {
  if (sizeof(VlKDForestNeighbor) != 16 || offsetof(VlKDForestNeighbor, index) != 8 || VlDistanceL2 != 1 ||
      VL_TYPE_FLOAT != 1 || VL_ERR_EOF != 5) return 1;   /* vl/kdtree.h:60-63, vl/mathop.h:628-630, vl/generic.h:21,113 */
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(VlSiftKeypoint), sizeof(pano_b200_keypoint), sizeof(pano_b200_pair),
         offsetof(VlSiftFilt, keys), offsetof(VlSiftFilt, nkeys), offsetof(VlSiftFilt, peak_thresh), offsetof(VlSiftFilt, grad_o));
  return 0;
}''')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", INC, str(src), "-o", str(exe)], check=True)
    got = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    # VLFeat 0.9.21 x86-64 layout of VlSiftFilt (vl/sift.h:39-78): 4 doubles, 8 ints, 3 pointers, 2 ints, pointer,
    # double, vl_size, pointer keys @120, nkeys @128, keys_res, 5 doubles @136.., grad pointer, grad_o
    assert got == ["32", "32", "64", "120", "128", "136", "184"]


def test_vlfeat_compat_include_path(tmp_path):
    """A source that keeps the reference's own include lines (ImageProcess.h:33-38) compiles against
    include/vl_b200/compat and links against libpano_b200.so alone."""
    src = tmp_path / "t.cpp"
    src.write_text(r'''
extern "C" {
#include "vl/generic.h"
#include "vl/sift.h"
#include "vl/kdtree.h"
}
int main() {
  assert(sizeof(vl_sift_pix) == 4);
  if (0) {   // link check only: there is no GPU here
    VlSiftFilt* f = vl_sift_new(8, 8, 1, 2, 0);
    vl_sift_delete(f);
    VlKDForest* k = vl_kdforest_new(VL_TYPE_FLOAT, 128, 1, VlDistanceL1);
    VlKDForestNeighbor nb[2];
    float q[128] = {0};
    vl_kdforest_build(k, 1, q);
    VlKDForestSearcher* s = vl_kdforest_new_searcher(k);
    vl_kdforestsearcher_query(s, nb, 2, q);
    vl_kdforestsearcher_delete(s);
    vl_kdforest_delete(k);
  }
  return VL_ERR_OK;
}''')
    exe = tmp_path / "t"
    libdir = os.path.join(ROOT, "computervisionimagestich2_b200")
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-I", os.path.join(INC, "vl_b200", "compat"), str(src), "-o",
                    str(exe), "-L", libdir, "-lpano_b200", f"-Wl,-rpath,{libdir}"], check=True)
    subprocess.run([str(exe)], check=True)


def test_kdforest_fails_loudly_without_a_gpu(so):
    lib = C.CDLL(so)
    lib.vl_kdforest_new.restype = C.c_void_p
    if lib.pano_b200_device_count() > 0:
        pytest.skip("a GPU is visible")
    assert lib.vl_kdforest_new(1, C.c_ulonglong(128), C.c_ulonglong(1), 0) is None


def test_context_fails_loudly_without_a_gpu(so):
    import computervisionimagestich2_b200 as pano
    L = pano.lib()
    if L.pano_b200_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(pano.PanoError):
        pano.Context(0)


def test_plan_canvas_is_host_only(so):
    """pano_b200_plan_canvas is pure host arithmetic (ImageProcess.cpp:206-216, 532-594) and runs without a device."""
    lib = C.CDLL(so)
    H = (C.c_double * 8)(0.9689389863682646, -0.009314512999418417, 0.00017184055549345834, 206.93934141946806,
                         0.0016961781782168055, 1.0005562851266288, -2.318615918566055e-06, 4.828864717610546)
    b = (C.c_float * 4)()
    s = (C.c_int * 2)()
    assert lib.pano_b200_plan_canvas(384, 512, H, 384, 512, b, s) == 0
    assert b[0] <= 0 and b[1] <= 0 and s[0] > 384 and s[1] >= 512


def test_bmp_roundtrip(tmp_path):
    from computervisionimagestich2_b200 import bmpio
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (3, 7, 13), dtype=np.uint8)  # 13*3 bytes per row -> exercises the 4-byte row padding
    p = str(tmp_path / "x.bmp")
    bmpio.save_bmp(p, img)
    assert np.array_equal(bmpio.load_bmp(p), img)
