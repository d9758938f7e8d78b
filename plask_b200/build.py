"""In-tree build of libplaskfem_cuda.so with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "plaskfem_cuda.cu")
OUT = os.path.join(_HERE, "libplaskfem_cuda.so")
DEPS = [SRC] + [os.path.join(_HERE, "csrc", f) for f in ("pfem_internal.cuh", "kernels_simple.cuh", "kernels_tiled.cuh", "kernels_tma.cuh", "kernels_fused.cuh",
                                                          "kernels_surface.cuh", "kernels_line.cuh", "kernels_ml.cuh")]
DEPS.append(os.path.join(os.path.dirname(_HERE), "include", "plaskfem_cuda.h"))


def nvcc_command(out=OUT, extra=()):
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "-Xcompiler", "-fPIC", "-shared", *extra, "-o", out, SRC]


def build(force=False, verbose=False):
    """Compile the CUDA library if it is missing or older than its sources."""
    if not force and os.path.exists(OUT):
        t = os.path.getmtime(OUT)
        if all(os.path.getmtime(d) <= t for d in DEPS if os.path.exists(d)):
            return OUT
    cmd = nvcc_command()
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT
