"""The C++ host adapter (include/plaskfem_cuda.hpp) compiled and run the way the solver plugin would use it."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "adapter_test.cpp")
LIBDIR = os.path.join(ROOT, "plask_b200")


@pytest.fixture(scope="module")
def adapter_binary():
    import plask_b200
    plask_b200.build()
    out = os.path.join(tempfile.mkdtemp(prefix="pfem_adapter_"), "adapter_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", out,
                           "-L", LIBDIR, "-lplaskfem_cuda", f"-Wl,-rpath,{LIBDIR}"])
    return out


def test_adapter_host_logic(adapter_binary):
    r = subprocess.run([adapter_binary, "host"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "adapter host tests ok" in r.stdout


@pytest.mark.gpu
def test_adapter_manufactured_solution_and_noconv(adapter_binary):
    r = subprocess.run([adapter_binary, "gpu"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "adapter gpu tests ok" in r.stdout


# ---- the Diffusion3D adapter (include/plaskdiff_cuda.hpp) ----------------------------------------------------------------------
DIFF_SRC = os.path.join(ROOT, "tests", "cpp", "diffusion_adapter_test.cpp")


@pytest.fixture(scope="module")
def diffusion_adapter_binary():
    import plask_b200
    plask_b200.build()
    out = os.path.join(tempfile.mkdtemp(prefix="pdiff_adapter_"), "diffusion_adapter_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), DIFF_SRC, "-o", out,
                           "-L", LIBDIR, "-lplaskfem_cuda", f"-Wl,-rpath,{LIBDIR}"])
    return out


def test_diffusion_adapter_host_logic(diffusion_adapter_binary):
    r = subprocess.run([diffusion_adapter_binary, "host"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "diffusion adapter host tests ok" in r.stdout


@pytest.mark.gpu
def test_diffusion_adapter_uniform_case(diffusion_adapter_binary):
    r = subprocess.run([diffusion_adapter_binary, "gpu"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "diffusion adapter gpu tests ok" in r.stdout
