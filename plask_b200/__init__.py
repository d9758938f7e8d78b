"""plask_b200 — B200-native (sm_100a, FP64 CUDA) implementation of the 3D finite-element
thermal / electrical solve of PLaSK's thermal.static.Static3D and electrical.shockley.Shockley3D.

Only the hot path lives here: `csrc/` holds the CUDA kernels and the C ABI
(`include/plaskfem_cuda.h`), `fem.py` the ctypes wrapper, `solvers.py` the host-side mirror of
the reference's solver interface.  The CPU oracle under `oracle/` is test infrastructure and is
never imported from this package.
"""
from ._lib import BadInput, ComputationError, NoDevice  # noqa: F401
from .build import build  # noqa: F401
