"""ctypes binding of libplaskfem_cuda.so (the C ABI declared in include/plaskfem_cuda.h).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``plask_b200.build``.
There is no CPU fallback: if the library or a CUDA device is missing every entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PFEM_LIB") or os.path.join(_HERE, "libplaskfem_cuda.so")   # PFEM_LIB: A/B runs of two builds

c_sz = C.c_size_t
c_dp = C.POINTER(C.c_double)


class ComputationError(RuntimeError):
    """plask::ComputationError (plask/exceptions.hpp) — numerical failure inside the solver."""


class BadInput(ValueError):
    """plask::BadInput — wrong argument or call order."""


class NoDevice(RuntimeError):
    """The CUDA algorithm was selected but no CUDA device / library is available."""


class Junction(C.Structure):
    _fields_ = [("bottom", c_sz), ("top", c_sz), ("left", c_sz), ("right", c_sz), ("back", c_sz), ("front", c_sz),
                ("ld", c_sz), ("offset", C.c_ssize_t), ("height", C.c_double)]


class Boundary(C.Structure):
    _fields_ = [("has_flux", C.POINTER(C.c_uint8)), ("flux", c_dp),
                ("has_conv", C.POINTER(C.c_uint8)), ("conv_coeff", c_dp), ("conv_ambient", c_dp),
                ("has_rad", C.POINTER(C.c_uint8)), ("rad_emissivity", c_dp), ("rad_ambient", c_dp),
                ("verbatim", C.c_int), ("mode2d", C.c_int), ("reserved", C.c_int * 2)]


class Opts(C.Structure):
    _fields_ = [("maxit", C.c_int), ("lin_tol", C.c_double), ("precond", C.c_int), ("outer_tol", C.c_double),
                ("loops", C.c_int), ("batch", C.c_int), ("variant", C.c_int), ("reserved", C.c_int * 5)]


class Stats(C.Structure):
    _fields_ = [("outer_loops", C.c_int), ("loopno", C.c_int), ("lin_iters", C.c_longlong), ("last_iters", C.c_int),
                ("converged", C.c_int), ("lin_relres", C.c_double), ("err", C.c_double), ("toterr", C.c_double),
                ("maxval", C.c_double), ("maxcur", C.c_double * 3), ("t_solve_ms", C.c_double),
                ("kernel_launches", C.c_longlong), ("lin_relres_precond", C.c_double), ("reserved", C.c_double * 3)]

    def as_dict(self):
        return dict(outer_loops=self.outer_loops, loopno=self.loopno, lin_iters=self.lin_iters,
                    last_iters=self.last_iters, converged=bool(self.converged), lin_relres=self.lin_relres,
                    err=self.err, toterr=self.toterr, maxval=self.maxval, maxcur=tuple(self.maxcur),
                    t_solve_ms=self.t_solve_ms, kernel_launches=self.kernel_launches,
                    lin_relres_precond=self.lin_relres_precond)


class Dynamic(C.Structure):
    _fields_ = [("time", C.c_double), ("timestep", C.c_double), ("methodparam", C.c_double), ("lumping", C.c_int),
                ("rebuildfreq", C.c_int), ("maxT_log", c_dp), ("maxT_log_len", c_sz), ("reserved", C.c_int * 4)]


class DiffOpts(C.Structure):
    _fields_ = [("loops", C.c_int), ("maxerr", C.c_double), ("maxit", C.c_int), ("lin_tol", C.c_double), ("verbatim", C.c_int),
                ("loop_cap", C.c_int), ("reserved", C.c_int * 6)]


class DiffStats(C.Structure):
    _fields_ = [("loops", C.c_int), ("converged", C.c_int), ("err", C.c_double), ("lin_iters", C.c_longlong),
                ("last_iters", C.c_int), ("lin_relres", C.c_double), ("t_solve_ms", C.c_double),
                ("kernel_launches", C.c_longlong), ("err_log", C.c_double * 64), ("lin_relres_precond", C.c_double), ("reserved", C.c_double * 3)]

    def as_dict(self):
        return dict(loops=self.loops, converged=bool(self.converged), err=self.err, lin_iters=self.lin_iters,
                    last_iters=self.last_iters, lin_relres=self.lin_relres, lin_relres_precond=self.lin_relres_precond,
                    t_solve_ms=self.t_solve_ms, kernel_launches=self.kernel_launches, err_log=list(self.err_log[:min(self.loops, 64)]))


# every symbol include/plaskfem_cuda.h and include/plaskdiff_cuda.h declare: name -> (restype, argtypes)
_vp = C.c_void_p
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_szp = C.POINTER(c_sz)
SYMBOLS = {
    "pfem_abi_version": (C.c_int, []),
    "pfem_device_count": (C.c_int, []),
    "pfem_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "pfem_destroy": (None, [_vp]),
    "pfem_strerror": (C.c_char_p, [C.c_int]),
    "pfem_last_error": (C.c_char_p, [_vp]),
    "pfem_set_layout": (C.c_int, [_vp, C.c_int]),
    "pfem_set_mesh": (C.c_int, [_vp, _szp, c_dp, c_dp, c_dp, _szp]),
    "pfem_set_materials": (C.c_int, [_vp, _u32p, C.c_uint32, C.c_double, C.c_double, C.c_uint32, c_dp, c_dp]),
    "pfem_set_dirichlet": (C.c_int, [_vp, c_sz, _szp, c_dp]),
    "pfem_set_source": (C.c_int, [_vp, c_dp]),
    "pfem_set_boundary": (C.c_int, [_vp, C.POINTER(Boundary)]),
    "pfem_edges2d_host": (C.c_int, [c_sz, c_dp, c_sz, c_dp, C.POINTER(Boundary), c_dp, c_dp, c_dp, c_dp]),
    "pfem_set_field": (C.c_int, [_vp, c_dp]),
    "pfem_fill_field": (C.c_int, [_vp, C.c_double]),
    "pfem_set_elem_temperature": (C.c_int, [_vp, c_dp, C.c_double]),
    "pfem_set_junctions": (C.c_int, [_vp, C.c_uint32, C.POINTER(Junction), _u32p, _u8p, C.c_double, C.c_double, c_sz,
                                     c_dp, c_dp, c_dp, C.c_int]),
    "pfem_slab_blob_size": (c_sz, []),
    "pfem_slab_configure": (C.c_int, [_vp, C.c_int, C.c_int, c_sz, c_sz]),
    "pfem_slab_export": (C.c_int, [_vp, _vp]),
    "pfem_slab_connect": (C.c_int, [_vp, _vp]),
    "pfem_default_opts": (None, [C.POINTER(Opts)]),
    "pfem_solve_thermal": (C.c_int, [_vp, C.POINTER(Opts), C.POINTER(Stats)]),
    "pfem_solve_shockley": (C.c_int, [_vp, C.POINTER(Opts), C.POINTER(Stats)]),
    "pfem_set_capacity": (C.c_int, [_vp, C.c_uint32, C.c_uint32, c_dp]),
    "pfem_set_axis_weight": (C.c_int, [_vp, C.c_int, c_dp]),
    "pfem_solve_dynamic": (C.c_int, [_vp, C.POINTER(Opts), C.POINTER(Dynamic), C.POINTER(Stats)]),
    "pfem_get_field": (C.c_int, [_vp, c_dp]),
    "pfem_interpolate_field": (C.c_int, [_vp, _szp, c_dp, c_dp, c_dp, _szp, c_dp]),
    "pfem_get_elem": (C.c_int, [_vp, C.c_int, _u8p, c_dp]),
    "pfem_get_junction_cond": (C.c_int, [_vp, c_dp]),
    "pfem_get_elem_temperature": (C.c_int, [_vp, c_sz, _szp, c_dp]),
    "pfem_set_noheat": (C.c_int, [_vp, _u8p]),
    "pfem_transfer_temperature": (C.c_int, [_vp, _vp]),
    "pfem_transfer_heat": (C.c_int, [_vp, _vp]),
    "pfem_update_conductivity_thermal": (C.c_int, [_vp]),
    "pfem_update_conductivity_shockley": (C.c_int, [_vp]),
    "pfem_set_conductivity": (C.c_int, [_vp, c_dp]),
    "pfem_apply": (C.c_int, [_vp, c_dp, c_dp, C.c_int]),
    "pfem_apply_precond": (C.c_int, [_vp, C.POINTER(Opts), c_dp, c_dp]),
    "pfem_get_rhs": (C.c_int, [_vp, c_dp]),
    "pfem_get_diag": (C.c_int, [_vp, c_dp]),
    "pfem_solve_linear": (C.c_int, [_vp, C.POINTER(Opts), C.POINTER(Stats)]),
    "pfem_get_info": (C.c_int, [_vp, C.c_int, c_dp]),
    "pfem_bench_pcg": (C.c_int, [_vp, C.POINTER(Opts), C.c_int, C.c_int, c_dp, c_dp, c_dp, C.POINTER(C.c_longlong)]),
    # include/plaskdiff_cuda.h
    "pdiff_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "pdiff_destroy": (None, [_vp]),
    "pdiff_last_error": (C.c_char_p, [_vp]),
    "pdiff_set_mesh": (C.c_int, [_vp, c_sz, c_sz, c_dp, c_dp, C.c_int, _u8p]),
    "pdiff_set_parameters": (C.c_int, [_vp, c_dp, c_dp, c_dp, c_dp]),
    "pdiff_set_current": (C.c_int, [_vp, c_dp]),
    "pdiff_set_modes": (C.c_int, [_vp, c_sz, c_dp, c_dp, c_dp]),
    "pdiff_set_concentration": (C.c_int, [_vp, c_dp]),
    "pdiff_get_concentration": (C.c_int, [_vp, c_dp]),
    "pdiff_default_opts": (None, [C.POINTER(DiffOpts)]),
    "pdiff_compute": (C.c_int, [_vp, C.POINTER(DiffOpts), C.POINTER(DiffStats)]),
    "pdiff_interpolate": (C.c_int, [_vp, c_sz, c_dp, c_dp, C.c_int, c_dp]),
    "pdiff_get_element_matrices": (C.c_int, [_vp, C.c_int, c_dp, c_dp]),
    "pdiff_apply": (C.c_int, [_vp, C.c_int, c_dp, c_dp]),
    "pdiff_get_rhs": (C.c_int, [_vp, C.c_int, c_dp]),
}

PFEM_OK, PFEM_NOT_CONVERGED = 0, 1
PFEM_ERR_CUDA, PFEM_ERR_NO_DEVICE, PFEM_ERR_BAD_INPUT, PFEM_ERR_STATE = -1, -2, -3, -4
PFEM_ERR_NOT_SPD, PFEM_ERR_NOMEM, PFEM_ERR_NAN = -5, -6, -7
ELEM_COND, ELEM_CURRENT, ELEM_HEAT, ELEM_FLUX = 0, 1, 2, 3
MAT_EXCLUDED = 0xFFFFFFFF
LAYOUT_ABI, LAYOUT_VERTICAL_MINOR = 0, 1
INFO_COND_ISO, INFO_ML_LEVELS, INFO_FUSED_CTAS, INFO_DEVICE_BYTES = 0, 1, 2, 3
DIFF_ORDER_01, DIFF_ORDER_10 = 0, 1
DIFF_INTERP_SPLINE, DIFF_INTERP_LINEAR = 0, 1

_lib = None


def load():
    """Load the CUDA library (raises NoDevice if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NoDevice(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a); "
                           "the CUDA algorithm has no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            if os.environ.get("PFEM_LIB") and not hasattr(lib, name):
                continue    # A/B run against an older build
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(ctx, rc, last_error="pfem_last_error"):
    """Map a pfem_status to the exceptions the PLaSK solver would throw."""
    if rc >= 0:
        return rc
    lib = load()
    detail = getattr(lib, last_error)(ctx).decode() if ctx else ""
    msg = f"{lib.pfem_strerror(rc).decode()}: {detail}" if detail else lib.pfem_strerror(rc).decode()
    if rc == PFEM_ERR_NO_DEVICE:
        raise NoDevice(msg)
    if rc in (PFEM_ERR_BAD_INPUT, PFEM_ERR_STATE):
        raise BadInput(msg)
    raise ComputationError(msg)
