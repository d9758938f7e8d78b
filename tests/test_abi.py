"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/plaskfem_cuda.h declares; without a GPU it fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import plask_b200
from plask_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    plask_b200.build()
    return _lib.load()


def declared_functions():
    names = set()
    for header in ("plaskfem_cuda.h", "plaskdiff_cuda.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(p(?:fem|diff)_[a-z_0-9]+)\s*\(", text))
    return sorted(names)


def test_header_symbols_exported(lib):
    names = declared_functions()
    assert len(names) >= 40 and sum(n.startswith("pdiff_") for n in names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes prototype"
    assert set(_lib.SYMBOLS) == set(names)


def test_abi_version_and_strerror(lib):
    assert lib.pfem_abi_version() == 1
    assert b"no CPU fallback" in lib.pfem_strerror(_lib.PFEM_ERR_NO_DEVICE)
    assert lib.pfem_strerror(0) == b"ok"


def test_struct_layouts_match_header(lib, tmp_path):
    """sizeof/offsetof of the ABI structs as the C compiler sees the header == the ctypes mirrors."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text(r"""
#include <stdio.h>
#include <stddef.h>
#include "plaskfem_cuda.h"
#include "plaskdiff_cuda.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(pfem_junction), sizeof(pfem_opts), sizeof(pfem_stats),
         offsetof(pfem_opts, outer_tol), offsetof(pfem_stats, maxcur), offsetof(pfem_stats, kernel_launches),
         sizeof(pfem_boundary), offsetof(pfem_boundary, rad_ambient), offsetof(pfem_boundary, verbatim));
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(pdiff_opts), sizeof(pdiff_stats), offsetof(pdiff_opts, lin_tol),
         offsetof(pdiff_opts, verbatim), offsetof(pdiff_stats, lin_relres), offsetof(pdiff_stats, err_log));
  return 0; }
""")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_lib.Junction), ctypes.sizeof(_lib.Opts), ctypes.sizeof(_lib.Stats),
            _lib.Opts.outer_tol.offset, _lib.Stats.maxcur.offset, _lib.Stats.kernel_launches.offset,
            ctypes.sizeof(_lib.Boundary), _lib.Boundary.rad_ambient.offset, _lib.Boundary.verbatim.offset,
            ctypes.sizeof(_lib.DiffOpts), ctypes.sizeof(_lib.DiffStats), _lib.DiffOpts.lin_tol.offset,
            _lib.DiffOpts.verbatim.offset, _lib.DiffStats.lin_relres.offset, _lib.DiffStats.err_log.offset]
    assert got == want


@pytest.mark.skipif(_lib.load().pfem_device_count() > 0 if os.path.exists(_lib.LIB_PATH) else False,
                    reason="a CUDA device is present")
def test_no_device_fails_loudly(lib):
    from plask_b200.fem import DeviceFem
    with pytest.raises(plask_b200.NoDevice):
        DeviceFem(0)
    from plask_b200.solvers import Static3D
    from plask_b200 import configs
    s = Static3D("therm")
    s.problem = configs.config_A(8)
    with pytest.raises(plask_b200.NoDevice):
        s.compute(1)


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "plask_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
