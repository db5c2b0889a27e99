"""CPU: the C-ABI shared library loads and exports every symbol that
``include/rtb200.h`` declares; struct mirrors in the ctypes binding have the
layout the header gives.  No compute call is made (no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "rtb200.h")


@pytest.fixture(scope="module")
def lib():
    from rt_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rtb200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from rt_b200 import _lib
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "librtb200.so does not export %s" % n
    # ... and the binding declares prototypes for exactly that set
    assert sorted(_lib.SYMBOLS) == names


def test_abi_version_and_constants(lib):
    from rt_b200 import _lib
    src = open(HEADER).read()
    ver = int(re.search(r"#define\s+RTB200_ABI_VERSION\s+(\d+)", src).group(1))
    assert lib.rtb200_abi_version() == ver == _lib.ABI_VERSION
    for name, val in (("MASS", _lib.TERM_MASS), ("SCAT", _lib.TERM_SCAT),
                      ("ICONV", _lib.TERM_ICONV), ("BCONV", _lib.TERM_BCONV), ("ALL", _lib.TERM_ALL)):
        assert int(re.search(r"#define\s+RTB200_TERM_%s\s+(\d+)u" % name, src).group(1)) == val
    assert int(re.search(r"#define\s+RTB200_MAX_NXY\s+(\d+)", src).group(1)) == _lib.MAX_NXY


def test_struct_layout_matches_header(tmp_path):
    """sizeof/offsetof of the three public structs, from a C program compiled
    against the header, equal the ctypes mirrors."""
    from rt_b200 import _lib
    prog = tmp_path / "layout.c"
    prog.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "rtb200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(rtb200_flatmesh), offsetof(rtb200_flatmesh, col_nx),
         offsetof(rtb200_flatmesh, g_pool), offsetof(rtb200_flatmesh, coef),
         offsetof(rtb200_flatmesh, send_src));
  printf("%zu %zu %zu\n", sizeof(rtb200_gmres_params), offsetof(rtb200_gmres_params, rtol),
         offsetof(rtb200_gmres_params, orth));
  printf("%zu %zu %zu\n", sizeof(rtb200_gmres_result), offsetof(rtb200_gmres_result, res_norm),
         offsetof(rtb200_gmres_result, n_apply));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    got = list(map(int, out))
    F, P, R = _lib.FlatMeshC, _lib.GmresParams, _lib.GmresResult
    want = [C.sizeof(F), F.col_nx.offset, F.g_pool.offset, F.coef.offset, F.send_src.offset,
            C.sizeof(P), P.rtol.offset, P.orth.offset,
            C.sizeof(R), R.res_norm.offset, R.n_apply.offset]
    assert got == want


def test_no_gpu_means_loud_failure(lib):
    """Without a CUDA device the product refuses to run instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rt_b200 import configs
    from rt_b200.rt import RTOperator
    mesh = configs.build_mesh([3.0, 2.0], [False, False], [3, 3, 3], nref_ang=2, nref_spt=1)
    prob = configs.two_strong_scatterers()
    with pytest.raises(RuntimeError):
        RTOperator(mesh, prob.kappa, prob.sigma, prob.Phi)
    with pytest.raises(RuntimeError):
        prob.solve(mesh)
    ctx = C.c_void_p()
    assert lib.rtb200_create(0, C.byref(ctx)) != 0          # error code, not a crash


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package refers to it."""
    pkg = os.path.join(ROOT, "radiative-transfer-amr_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
