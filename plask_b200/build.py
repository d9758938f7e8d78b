"""In-tree build of libplaskfem_cuda.so with nvcc for sm_100a (cross-compiles without a GPU).

Two translation units — the brick FEM path (plaskfem_cuda.cu, include/plaskfem_cuda.h) and the carrier-diffusion path
(plaskdiff_cuda.cu, include/plaskdiff_cuda.h) — are compiled to objects under plask_b200/_obj/ and linked into ONE shared library;
an object is rebuilt only when one of its own sources is newer."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_INC = os.path.join(os.path.dirname(_HERE), "include")
_OBJ = os.path.join(_HERE, "_obj")
OUT = os.path.join(_HERE, "libplaskfem_cuda.so")
SRC = os.path.join(_CSRC, "plaskfem_cuda.cu")
UNITS = {
    "plaskfem_cuda": [SRC] + [os.path.join(_CSRC, f) for f in (
        "pfem_internal.cuh", "kernels_simple.cuh", "kernels_tiled.cuh", "kernels_tma.cuh", "kernels_fused.cuh", "kernels_surface.cuh",
        "kernels_line.cuh", "kernels_ml.cuh")] + [os.path.join(_INC, "plaskfem_cuda.h")],
    "plaskdiff_cuda": [os.path.join(_CSRC, "plaskdiff_cuda.cu"), os.path.join(_CSRC, "kernels_diffusion.cuh"),
                       os.path.join(_INC, "plaskdiff_cuda.h"), os.path.join(_INC, "plaskfem_cuda.h")],
}
DEPS = [d for deps in UNITS.values() for d in deps]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def nvcc_command(out=OUT, extra=()):
    """one-step build of the whole library (A/B builds with -DPFEM_X=...)"""
    return ["nvcc", *ARCH, "-shared", *extra, "-o", out] + [deps[0] for deps in UNITS.values()]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile the CUDA library if it is missing or older than its sources."""
    os.makedirs(_OBJ, exist_ok=True)
    objs, procs = [], []
    for name, deps in UNITS.items():
        obj = os.path.join(_OBJ, name + ".o")
        objs.append(obj)
        if force or _stale(obj, deps):
            cmd = ["nvcc", *ARCH, "-c", "-o", obj, deps[0]]
            if verbose:
                print(" ".join(cmd))
            procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    if force or procs or _stale(OUT, objs):
        cmd = ["nvcc", *ARCH, "-shared", "-o", OUT, *objs]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return OUT
