"""patches/plask-algorithm-cuda.diff — the binding of the library into the PLaSK tree — applies cleanly to the reference
sources (build container only: /root/reference is absent on the GPU box), and every adapter entry point the patched solver
code calls exists in include/plaskfem_cuda.hpp."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCH = os.path.join(ROOT, "patches", "plask-algorithm-cuda.diff")
REF = "/root/reference"


def _files():
    return sorted(set(re.findall(r"^\+\+\+ b/(\S+)", open(PATCH).read(), flags=re.M)))


def test_patch_touches_the_documented_files():
    want = {"plask/common/fem/fem_solver.hpp", "python/plask/common/fem/fem.cpp", "plask/common/fem.yml",
            "solvers/thermal/static/therm3d.hpp", "solvers/thermal/static/therm3d.cpp",
            "solvers/thermal/static/therm2d.hpp", "solvers/thermal/static/therm2d.cpp",
            "solvers/electrical/shockley/electr3d.hpp", "solvers/electrical/shockley/electr3d.cpp",
            "solvers/electrical/shockley/electr2d.hpp", "solvers/electrical/shockley/electr2d.cpp",
            "solvers/electrical/shockley/beta.hpp", "solvers/electrical/shockley/python/electr_python.cpp",
            "solvers/thermal/static/CMakeLists.txt", "solvers/electrical/shockley/CMakeLists.txt",
            "solvers/thermal/dynamic/femT3d.hpp", "solvers/thermal/dynamic/femT3d.cpp", "solvers/thermal/dynamic/CMakeLists.txt",
            "solvers/thermal/dynamic/femT2d.hpp", "solvers/thermal/dynamic/femT2d.cpp",
            "solvers/electrical/diffusion/diffusion3d.hpp", "solvers/electrical/diffusion/diffusion3d.cpp",
            "solvers/electrical/diffusion/CMakeLists.txt"}
    assert set(_files()) == want


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree")
def test_patch_applies_to_the_reference(tmp_path):
    r = subprocess.run(["patch", "-p1", "--dry-run", "--batch", "-d", REF, "-i", PATCH], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    # and for real on a scratch copy of the touched files
    for f in _files():
        os.makedirs(os.path.dirname(tmp_path / f), exist_ok=True)
        shutil.copyfile(os.path.join(REF, f), tmp_path / f)
    r = subprocess.run(["patch", "-p1", "--batch", "-d", str(tmp_path), "-i", PATCH], capture_output=True, text=True)
    assert r.returncode == 0 and "FAILED" not in r.stdout, r.stdout + r.stderr
    text = open(tmp_path / "plask/common/fem/fem_solver.hpp").read()
    assert "ALGORITHM_CUDA" in text and '.value("cuda", ALGORITHM_CUDA)' in text
    assert "computeCuda" in open(tmp_path / "solvers/thermal/static/therm3d.cpp").read()
    assert "shockleyParameters" in open(tmp_path / "solvers/electrical/shockley/beta.hpp").read()
    t2 = open(tmp_path / "solvers/thermal/static/therm2d.cpp").read()
    assert "cuda->set_axis_weight(1, emb.radial_weights());" in t2 and "plaskfem::Embedding2D::boundary_mode(" in t2
    assert "if (this->algorithm == ALGORITHM_CUDA) return computeCuda(loops, btemperature, bheatflux, bconvection, bradiation);" in t2
    dyn = open(tmp_path / "solvers/thermal/dynamic/femT3d.cpp").read()
    assert "solve_dynamic" in dyn and "set_capacity" in dyn and "if (algorithm == ALGORITHM_CUDA) return computeCuda(time, btemperature);" in dyn
    d2 = open(tmp_path / "solvers/thermal/dynamic/femT2d.cpp").read()
    assert "if (this->algorithm == ALGORITHM_CUDA) return computeCuda(time, btemperature);" in d2 and "cuda->set_capacity(tables, cpdens);" in d2
    assert "emb.add_dirichlet(bc, cudaNode[r], cond.value);" in d2
    e2h = open(tmp_path / "solvers/electrical/shockley/electr2d.hpp").read()
    # BetaSolver<Geometry2D...> overrides shockleyParameters: its 2-D base must declare the virtual as well
    assert "virtual bool shockleyParameters(" in e2h and "virtual bool shockleyParameters(" in open(tmp_path / "solvers/electrical/shockley/electr3d.hpp").read()
    e2 = open(tmp_path / "solvers/electrical/shockley/electr2d.cpp").read()
    assert "if (this->algorithm == ALGORITHM_CUDA) return computeCuda(loops, vconst);" in e2
    assert "pfem_junction{act.bottom, act.top, act.left, act.right, 0, 1, 1, act.offset, act.height}" in e2
    dif = open(tmp_path / "solvers/electrical/diffusion/diffusion3d.cpp").read()
    assert "computeCuda(loops, act, active, A, B, C, D, J, nmodes, Ps, nrs);" in dif and "#include <plaskdiff_cuda.hpp>" in dif


def test_every_adapter_call_of_the_patch_exists():
    """the patched solver code can only be compiled inside a PLaSK build; what can be checked here is that every member of
    plaskfem:: it uses is declared by the adapter header the plugin would include"""
    added = "\n".join(l[1:] for l in open(PATCH) if l.startswith("+") and not l.startswith("+++"))
    hdr = open(os.path.join(ROOT, "include", "plaskfem_cuda.hpp")).read()
    for name in set(re.findall(r"cuda->(\w+)\(", added)):
        assert re.search(r"\b%s\(" % name, hdr), f"Context::{name} is not declared in plaskfem_cuda.hpp"
    for name in set(re.findall(r"plaskfem::(\w+)", added)):
        assert re.search(r"\b%s\b" % name, hdr), f"plaskfem::{name} is not declared in plaskfem_cuda.hpp"
    for name in set(re.findall(r"\bemb\.(\w+)\(", added)):
        assert re.search(r"\b%s\(" % name, hdr), f"Embedding2D::{name} is not declared in plaskfem_cuda.hpp"
    for name in ("add_node", "node_to_full", "mark_excluded", "PRECOND_MLJ", "Embedding2D", "radial_weights", "lift"):
        assert name in added and name in hdr
    dhdr = open(os.path.join(ROOT, "include", "plaskdiff_cuda.hpp")).read()
    for name in set(re.findall(r"region->(\w+)\(", added)):
        assert re.search(r"\b%s\(" % name, dhdr), f"Region::{name} is not declared in plaskdiff_cuda.hpp"
    for name in set(re.findall(r"plaskdiff::(\w+)", added)):
        assert re.search(r"\b%s\b" % name, dhdr), f"plaskdiff::{name} is not declared in plaskdiff_cuda.hpp"
