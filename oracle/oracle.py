"""CPU oracle for the PLaSK Static3D / Shockley3D hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``plask_b200/`` may import this module.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` use it, and only as the checker / the CPU arm.

It drives ``liboracle.so`` (``fem3d_oracle.c``, our C restatement of the reference loops)
and, when present, ``_ref/libnspcg_ref.so`` — the reference's own vendored NSPCG compiled
from ``/root/reference/extlib/nspcg/nspcg.c`` — plus LAPACK ``dpbtrf/dpbtrs`` from the
OpenBLAS that ships inside scipy (the reference links an external LAPACK,
``plask/common/fem/cholesky_matrix.hpp:24-50``).

The outer loops below follow, statement by statement,
``solvers/thermal/static/therm3d.cpp:281-340`` and
``solvers/electrical/shockley/electr3d.cpp:356-442``.
"""
import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
c_sz = C.c_size_t
c_dp = C.POINTER(C.c_double)


def build(ref=True, quiet=True):
    """Compile liboracle.so and, if /root/reference exists, _ref/libnspcg_ref.so."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", _HERE], stdout=out)
    if ref and os.path.isdir("/root/reference/extlib/nspcg"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=out)


class _Mesh(C.Structure):
    _fields_ = [("n", c_sz * 3), ("ax", c_dp * 3), ("ns", c_sz * 3), ("es", c_sz * 3)]


class _Active(C.Structure):
    _fields_ = [("bottom", c_sz), ("top", c_sz), ("left", c_sz), ("right", c_sz), ("back", c_sz), ("front", c_sz),
                ("ld", c_sz), ("offset", C.c_ssize_t), ("height", C.c_double)]


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _lib = C.CDLL(path)
        _lib.orc_table_at.restype = C.c_double
        _lib.orc_shockley_currents.restype = C.c_double
        _lib.orc_integrate_current.restype = C.c_double
        _lib.orc_total_heat.restype = C.c_double
        _lib.orc_total_energy.restype = C.c_double
        _lib.orc_mesh_nodes.restype = c_sz
        _lib.orc_mesh_elements.restype = c_sz
    return _lib


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libnspcg_ref.so"))


def ref():
    """The reference's own NSPCG (None if oracle/_ref was not built)."""
    global _ref
    if _ref is None and ref_available():
        _ref = C.CDLL(os.path.join(_HERE, "_ref", "libnspcg_ref.so"))
        _ref.ref_nspcg_new.restype = C.c_void_p
        _ref.ref_nspcg_free.argtypes = [C.c_void_p]
    return _ref


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


ORDERS = {"012": (0, 1, 2), "021": (0, 2, 1), "102": (1, 0, 2), "120": (1, 2, 0), "201": (2, 0, 1), "210": (2, 1, 0)}


def optimal_order(n):
    """RectilinearMesh3D::setOptimalIterationOrder, plask/mesh/rectilinear3d.cpp:74-85."""
    for name in ("012", "021", "102", "120", "201", "210"):
        f, s, t = ORDERS[name]
        if n[t] <= n[s] <= n[f]:
            return name
    return "210"


class Mesh:
    """RectangularMesh<3> (plask/mesh/rectilinear3d.hpp): three coordinate axes [um] and an
    iteration order 'major medium minor' (:349)."""

    def __init__(self, ax0, ax1, ax2, order="012"):
        self.axes = [np.ascontiguousarray(a, dtype=np.float64) for a in (ax0, ax1, ax2)]
        self.n = tuple(len(a) for a in self.axes)
        if order == "optimal":
            order = optimal_order(self.n)
        self.order = order
        self.c = _Mesh()
        o = (C.c_int * 3)(*ORDERS[order])
        lib().orc_mesh_init(C.byref(self.c), c_sz(self.n[0]), c_sz(self.n[1]), c_sz(self.n[2]), _p(self.axes[0]),
                            _p(self.axes[1]), _p(self.axes[2]), o)
        self.N = self.n[0] * self.n[1] * self.n[2]
        self.E = (self.n[0] - 1) * (self.n[1] - 1) * (self.n[2] - 1)
        self.ns = tuple(self.c.ns)
        self.es = tuple(self.c.es)
        s = sorted(self.ns)
        self.minor, self.major = s[1], s[2]
        self.icords = np.zeros(14, dtype=np.int32)
        lib().orc_sparse14_offsets(c_sz(self.major), c_sz(self.minor), _p(self.icords, C.c_int))

    @property
    def ref(self):
        return C.byref(self.c)

    def node(self, i0, i1, i2):
        return i0 * self.ns[0] + i1 * self.ns[1] + i2 * self.ns[2]

    def nodes_grid(self):
        """node index array of shape (n0,n1,n2)"""
        i0, i1, i2 = np.meshgrid(*[np.arange(k) for k in self.n], indexing="ij")
        return i0 * self.ns[0] + i1 * self.ns[1] + i2 * self.ns[2]

    def elems_grid(self):
        i0, i1, i2 = np.meshgrid(*[np.arange(k - 1) for k in self.n], indexing="ij")
        return i0 * self.es[0] + i1 * self.es[1] + i2 * self.es[2]


class Tables:
    """Per-material-id conductivity tables on a uniform T grid (sampled by the host from
    material->thermk / material->cond)."""

    def __init__(self, T0, dT, lat, vert):
        self.T0, self.dT = float(T0), float(dT)
        self.lat = np.ascontiguousarray(lat, dtype=np.float64)
        self.vert = np.ascontiguousarray(vert, dtype=np.float64)
        assert self.lat.shape == self.vert.shape and self.lat.ndim == 2
        self.nmat, self.nT = self.lat.shape


def thickness(mesh, elem_mat):
    out = np.empty(mesh.E)
    lib().orc_thickness(mesh.ref, _p(elem_mat, C.c_uint32), _p(out))
    return out


# ----------------------------------------------------------------------------- matrices


class Sparse14:
    """SparseBandMatrix, plask/common/fem/iterative_matrix.hpp:347-486."""

    def __init__(self, mesh):
        self.mesh = mesh
        self.data = np.zeros(14 * mesh.N)
        self.state = None

    def assemble(self, cond, heat, B):
        rc = lib().orc_assemble_sparse14(self.mesh.ref, _p(cond), _p(heat) if heat is not None else None,
                                         _p(self.data), _p(B))
        assert rc == 0

    def add_coo(self, rows, cols, vals):
        lib().orc_add_coo_sparse14(c_sz(self.mesh.N), _p(self.mesh.icords, C.c_int), _p(self.data), c_sz(len(vals)),
                                   _p(rows, c_sz), _p(cols, c_sz), _p(vals))

    def apply_bc(self, B, nodes, values):
        lib().orc_apply_bc_sparse14(c_sz(self.mesh.N), _p(self.mesh.icords, C.c_int), _p(self.data), _p(B),
                                    c_sz(len(nodes)), _p(nodes, c_sz), _p(values))

    def mult(self, x):
        y = np.zeros(self.mesh.N)
        lib().orc_addmult_sparse14(c_sz(self.mesh.N), _p(self.mesh.icords, C.c_int), _p(self.data), _p(x), _p(y))
        return y

    def solve_nspcg(self, B, X, precond="ic", accel="cg", maxit=1000, maxerr=1e-6, nfact=10):
        """SparseMatrix::solverhs, iterative_matrix.hpp:141-339, through the reference's NSPCG."""
        r = ref()
        if r is None:
            raise RuntimeError("oracle/_ref/libnspcg_ref.so not built (needs /root/reference)")
        if self.state is None:
            self.state = C.c_void_p(r.ref_nspcg_new())
        pre = ["rich", "jac", "ljac", "ljacx", "sor", "ssor", "ic", "mic", "lsp", "neu", "lsor", "lssor", "llsp",
               "lneu", "bic", "bicx", "mbic", "mbicx"].index(precond)
        acc = ["cg", "si", "sor", "srcg", "srsi"].index(accel)
        iters, err = C.c_int(0), C.c_double(0)
        tf, tt = C.c_double(0), C.c_double(0)
        rhs = B.copy()
        ier = r.ref_nspcg_solve(self.state, pre, acc, C.c_int(self.mesh.N), C.c_int(self.mesh.major),
                                C.c_int(self.mesh.minor), _p(self.data), _p(self.mesh.icords, C.c_int), _p(X),
                                _p(rhs), C.c_int(maxit), C.c_double(maxerr), C.c_int(nfact), C.byref(iters),
                                C.byref(err), C.byref(tf), C.byref(tt))
        # ier == -7 with a vanishing residual is exact convergence (SURVEY.md §8c gotcha)
        if ier < 0 and ier != -7:
            raise RuntimeError(f"NSPCG error ier={ier}")
        return dict(ier=ier, converged=(ier != 1), iters=iters.value, err=err.value)

    def solve_pcg(self, B, X, maxit=100000, tol=1e-10):
        relres = C.c_double(0)
        it = lib().orc_pcg_jacobi_sparse14(c_sz(self.mesh.N), _p(self.mesh.icords, C.c_int), _p(self.data), _p(B),
                                           _p(X), C.c_int(maxit), C.c_double(tol), C.byref(relres))
        return dict(ier=0 if it >= 0 else it, converged=(it >= 0 and relres.value <= tol * 1.01), iters=it, err=relres.value)

    def __del__(self):
        if self.state is not None and ref() is not None:
            ref().ref_nspcg_free(self.state)


class Dpb:
    """DpbMatrix, plask/common/fem/cholesky_matrix.hpp:58-120 (LAPACK lower band storage)."""

    def __init__(self, mesh):
        self.mesh = mesh
        kd, ld = c_sz(0), c_sz(0)
        lib().orc_dpb_dims(mesh.ref, C.byref(kd), C.byref(ld))
        self.kd, self.ld = kd.value, ld.value
        self.data = np.zeros((mesh.N, self.ld + 1))

    def assemble(self, cond, heat, B):
        lib().orc_assemble_dpb(self.mesh.ref, _p(cond), _p(heat) if heat is not None else None, c_sz(self.ld),
                               _p(self.data), _p(B))

    def add_coo(self, rows, cols, vals):
        lib().orc_add_coo_dpb(c_sz(self.ld), _p(self.data), c_sz(len(vals)), _p(rows, c_sz), _p(cols, c_sz), _p(vals))

    def apply_bc(self, B, nodes, values):
        lib().orc_apply_bc_dpb(c_sz(self.mesh.N), c_sz(self.kd), c_sz(self.ld), _p(self.data), _p(B),
                               c_sz(len(nodes)), _p(nodes, c_sz), _p(values))

    def solve(self, B, X, threads=None):
        """factorize() + solverhs(): dpbtrf('L') + dpbtrs, cholesky_matrix.hpp:90-111."""
        from scipy.linalg import lapack
        ab = self.data.T  # Fortran-ordered (ld+1, N) view, no copy
        c, info = lapack.dpbtrf(ab, lower=1, overwrite_ab=1)
        if info > 0:
            raise RuntimeError(f"leading minor of order {info} of the stiffness matrix is not positive-definite")
        assert info == 0
        x, info = lapack.dpbtrs(c, B, lower=1)
        assert info == 0
        X[:] = x
        return dict(ier=0, converged=True, iters=0, err=0.)


class DpbMasked:
    """DpbMatrix on a RectangularMaskedMesh3D — empty-elements="exclude", the default of the reference's Cholesky path
    (fem_solver.hpp:182-189,219-231): rank = number of masked nodes, band from the masked element spans.  Vectors
    given to / returned from this class are FULL-mesh arrays; entries of nodes outside the masked mesh are left alone
    (the reference has no such entries)."""

    def __init__(self, mesh, included):
        self.mesh = mesh
        self.included = np.ascontiguousarray(included, dtype=np.uint8)
        assert self.included.size == mesh.E
        self.nodemap = np.zeros(mesh.N, dtype=np.uintp)
        lib().orc_masked_nodes.restype = c_sz
        self.Nm = lib().orc_masked_nodes(mesh.ref, _p(self.included, C.c_uint8), _p(self.nodemap, c_sz))
        self.active = self.nodemap != np.uintp(np.iinfo(np.uintp).max)
        kd, ld = c_sz(0), c_sz(0)
        lib().orc_dpb_dims_masked(mesh.ref, _p(self.included, C.c_uint8), _p(self.nodemap, c_sz), C.byref(kd), C.byref(ld))
        self.kd, self.ld = kd.value, ld.value
        self.data = np.zeros((self.Nm, self.ld + 1))
        self.Bm = np.zeros(self.Nm)

    def assemble(self, cond, heat, B):
        lib().orc_assemble_dpb_masked(self.mesh.ref, _p(cond), _p(heat) if heat is not None else None,
                                      _p(self.included, C.c_uint8), _p(self.nodemap, c_sz), c_sz(self.Nm), c_sz(self.ld),
                                      _p(self.data), _p(self.Bm))
        B[:] = 0.
        B[self.active] = self.Bm

    def add_coo(self, rows, cols, vals):
        """convection terms of the kept elements (their nodes all belong to the masked mesh)"""
        none = np.uintp(np.iinfo(np.uintp).max)
        mr = np.ascontiguousarray(self.nodemap[np.asarray(rows, dtype=np.int64)], dtype=np.uintp)
        mc = np.ascontiguousarray(self.nodemap[np.asarray(cols, dtype=np.int64)], dtype=np.uintp)
        assert not (mr == none).any() and not (mc == none).any()
        lib().orc_add_coo_dpb(c_sz(self.ld), _p(self.data), c_sz(len(vals)), _p(mr, c_sz), _p(mc, c_sz), _p(np.ascontiguousarray(vals)))

    def apply_bc(self, B, nodes, values):
        """boundary conditions live on the masked mesh: places outside it hold no nodes"""
        self.Bm[:] = B[self.active]      # load terms added to B since assemble() (heat flux, convection, radiation)
        mn = self.nodemap[np.asarray(nodes, dtype=np.int64)]
        keep = mn != np.uintp(np.iinfo(np.uintp).max)
        mn = np.ascontiguousarray(mn[keep], dtype=np.uintp)
        mv = np.ascontiguousarray(np.asarray(values, dtype=np.float64)[keep])
        lib().orc_apply_bc_dpb(c_sz(self.Nm), c_sz(self.kd), c_sz(self.ld), _p(self.data), _p(self.Bm),
                               c_sz(len(mn)), _p(mn, c_sz), _p(mv))
        B[self.active] = self.Bm

    def solve(self, B, X, threads=None):
        from scipy.linalg import lapack
        ab = self.data.T
        c, info = lapack.dpbtrf(ab, lower=1, overwrite_ab=1)
        if info > 0:
            raise RuntimeError(f"leading minor of order {info} of the stiffness matrix is not positive-definite")
        assert info == 0
        x, info = lapack.dpbtrs(c, self.Bm, lower=1)
        assert info == 0
        X[self.active] = x
        return dict(ier=0, converged=True, iters=0, err=0.)


# ------------------------------------------------------------------------- Static3D
# ------------------------------------------------------------------------- Static3D


class BoundaryTerms:
    """heatflux_boundary / convection_boundary / radiation_boundary of ThermalFem3DSolver (therm3d.hpp:79-82) as the
    per-node optional values `BoundaryConditions::getValue` yields (first matching condition wins,
    plask/mesh/boundary_conditions.hpp:182-186).  Each list holds (nodes, value...) tuples in definition order:
    heatflux (nodes, q [W/m2]); convection (nodes, coeff [W/m2/K], ambient [K]); radiation (nodes, emissivity, ambient)."""

    def __init__(self, N, heatflux=(), convection=(), radiation=()):
        def dense(conds, nval):
            if not conds:
                return None, [np.zeros(1)] * nval
            has = np.zeros(N, dtype=np.uint8)
            vals = [np.zeros(N) for _ in range(nval)]
            for cond in conds:
                nodes = np.asarray(cond[0], dtype=np.int64)
                new = nodes[has[nodes] == 0]
                for k in range(nval):
                    vals[k][new] = cond[1 + k]
                has[new] = 1
            return has, vals
        self.has_flux, (self.flux,) = dense(list(heatflux), 1)
        self.has_conv, (self.coeff, self.camb) = dense(list(convection), 2)
        self.has_rad, (self.emis, self.ramb) = dense(list(radiation), 2)

    def terms(self, mesh, T, B, quirk, included=None):
        """setBoundaries for every element (therm3d.cpp:140-168,242-268): adds the load terms to B, returns COO triplets."""
        def pu(a):
            return _p(a, C.c_uint8) if a is not None else None
        cap = 1 << 16
        while True:
            rows, cols, vals = np.zeros(cap, dtype=np.uintp), np.zeros(cap, dtype=np.uintp), np.zeros(cap)
            B0 = B.copy()
            lib().orc_boundary_terms.restype = c_sz
            nt = lib().orc_boundary_terms(mesh.ref, _p(np.ascontiguousarray(T)), pu(self.has_flux), _p(self.flux),
                                          pu(self.has_conv), _p(self.coeff), _p(self.camb), pu(self.has_rad),
                                          _p(self.emis), _p(self.ramb), C.c_int(1 if quirk else 0), c_sz(cap),
                                          _p(rows, c_sz), _p(cols, c_sz), _p(vals), _p(B0), pu(included))
            if nt <= cap:
                B[:] = B0
                return rows[:nt].copy(), cols[:nt].copy(), vals[:nt].copy()
            cap = int(nt)


class Static3DOracle:
    """ThermalFem3DSolver restated (solvers/thermal/static/therm3d.cpp)."""

    def __init__(self, mesh, elem_mat, tables, dirichlet_nodes, dirichlet_values, heat=None, inittemp=300.,
                 maxerr=0.05, algorithm="cholesky", precond="ic", itmaxerr=1e-6, maxit=1000, nfact=10,
                 boundaries=None, quirk=True, included=None):
        self.mesh = mesh
        # empty-elements="exclude": uint8 [E], 0 = element outside the masked mesh (Cholesky only); None = full mesh
        self.included = None if included is None else np.ascontiguousarray(included, dtype=np.uint8)
        self.boundaries = boundaries   # BoundaryTerms or None (heat flux / convection / radiation, therm3d.cpp:242-268)
        self.quirk = quirk             # True: verbatim local-slot accumulation of setBoundaries (therm3d.cpp:157-162)
        self.elem_mat = np.ascontiguousarray(elem_mat, dtype=np.uint32)
        self.tables = tables
        self.bc_nodes = np.ascontiguousarray(dirichlet_nodes, dtype=np.uintp)
        self.bc_values = np.ascontiguousarray(dirichlet_values, dtype=np.float64)
        self.heat = None if heat is None else np.ascontiguousarray(heat, dtype=np.float64)
        self.inittemp, self.maxerr = inittemp, maxerr
        self.algorithm, self.precond, self.itmaxerr, self.maxit, self.nfact = algorithm, precond, itmaxerr, maxit, nfact
        self.temperatures = np.full(mesh.N, float(inittemp))  # onInitialize, therm3d.cpp:79
        self.loopno = 0
        self.converged = True
        self.history = []
        self.timing = dict(assembly=0., solve=0.)
        self.conds = np.zeros((mesh.E, 2))
        self._A = None

    def _matrix(self):
        if self._A is None:
            if self.included is not None:
                assert self.algorithm == "cholesky", "the masked mesh is restated for the Cholesky path only"
                self._A = DpbMasked(self.mesh, self.included)
                self.temperatures[~self._A.active] = 0.   # nodes outside the masked mesh do not exist
            else:
                self._A = Dpb(self.mesh) if self.algorithm == "cholesky" else Sparse14(self.mesh)
        return self._A

    def set_matrix(self, A, B):
        """setMatrix, therm3d.cpp:170-279 (Dirichlet + volumetric heat only)."""
        t = self.tables
        lib().orc_thermal_conds(self.mesh.ref, _p(self.temperatures), _p(self.elem_mat, C.c_uint32),
                                C.c_uint32(t.nT), C.c_double(t.T0), C.c_double(t.dT), _p(t.lat), _p(t.vert),
                                _p(self.conds))
        if self.included is not None:
            self.conds[self.included == 0] = 0.
        A.assemble(self.conds, self.heat if self.heat is not None else np.zeros(self.mesh.E), B)
        if self.boundaries is not None:
            rows, cols, vals = self.boundaries.terms(self.mesh, self.temperatures, B, self.quirk, self.included)
            A.add_coo(rows, cols, vals)
        A.apply_bc(B, self.bc_nodes, self.bc_values)

    def compute(self, loops=0):
        """compute, therm3d.cpp:281-340."""
        A = self._matrix()
        loop, toterr = 0, 0.
        B = np.zeros(self.mesh.N)
        while True:
            temp0 = self.temperatures.copy()
            t0 = time.perf_counter()
            self.set_matrix(A, B)
            t1 = time.perf_counter()
            if self.algorithm == "cholesky":
                info = A.solve(B, self.temperatures)
            elif self.algorithm == "iterative":
                info = A.solve_nspcg(B, self.temperatures, precond=self.precond, maxit=self.maxit,
                                     maxerr=self.itmaxerr, nfact=self.nfact)
            else:  # 'pcg': oracle's own Jacobi-PCG with a true-residual test
                info = A.solve_pcg(B, self.temperatures, maxit=self.maxit, tol=self.itmaxerr)
            t2 = time.perf_counter()
            self.timing["assembly"] += t1 - t0
            self.timing["solve"] += t2 - t1
            self.converged = info["converged"]
            err, maxT = C.c_double(0), C.c_double(0)
            lib().orc_thermal_error(c_sz(self.mesh.N), _p(self.temperatures), _p(temp0), C.byref(err), C.byref(maxT))
            err, self.maxT = err.value, maxT.value
            toterr = max(toterr, err)
            self.loopno += 1
            loop += 1
            self.history.append(dict(loop=loop, maxT=self.maxT, err=err, iters=info["iters"], lin_err=info["err"]))
            if not ((not self.converged or err > self.maxerr) and (loops == 0 or loop < loops)):
                break
        return toterr

    def heat_fluxes(self):
        """saveHeatFluxes, therm3d.cpp:342-384: thermk is re-evaluated at the CURRENT temperatures (:362-370)."""
        t = self.tables
        lib().orc_thermal_conds(self.mesh.ref, _p(self.temperatures), _p(self.elem_mat, C.c_uint32),
                                C.c_uint32(t.nT), C.c_double(t.T0), C.c_double(t.dT), _p(t.lat), _p(t.vert),
                                _p(self.conds))
        if self.included is not None:
            self.conds[self.included == 0] = 0.   # saveHeatFluxes loops over maskedMesh->elements() (:352)
        flux = np.zeros((self.mesh.E, 3))
        lib().orc_heat_flux(self.mesh.ref, _p(self.temperatures), _p(self.conds), _p(flux))
        return flux


# ------------------------------------------------------------------------- Shockley3D


def table_lookup(tab, mat, T0, dT, T):
    """numpy twin of table_at (fem3d_oracle.c:115-125): linear interpolation, clamped at both ends"""
    tab = np.asarray(tab)
    nT = tab.shape[1]
    t = np.clip((np.asarray(T, dtype=np.float64) - T0) / dT, 0., float(nT - 1))
    i = np.minimum(t.astype(np.int64), nT - 2)
    f = t - i
    a, b = tab[mat, i], tab[mat, i + 1]
    return a + f * (b - a)


class Dynamic3DOracle:
    """DynamicThermalFem3DSolver restated (solvers/thermal/dynamic/femT3d.cpp) as a CORRECTED specification — test infrastructure.

    setMatrix (:127-255) is followed to the letter: A = methodparam*K + C, B = -(1-methodparam)*K + C with the brick stiffness K
    (:186-199, the same as therm3d.cpp:226-237), the element capacity c = cp*dens*0.125e-9*dx*dy*dz/timestep (:176), lumped
    (:207-212) or consistent (:216-231), F = 0.125e-18*dx*dy*dz*heat (:179,204), conductivities and capacities at the mean of the
    8 node temperatures (:163).  The time loop (:271-295) is corrected in two places, because as written it cannot be a parity
    target: (i) `std::swap(temperatures, X)` (:287) replaces the solution by the right-hand side for the iterative algorithm
    (the 2-D twin femT2d.cpp:424-428 has no such line); (ii) only A and F are eliminated on the Dirichlet rows (:236) while
    B.mult keeps them (:279), so a fixed node would receive value + (B T)_r.  Here
        A T^{n+1} = B T^n + F on the free rows,   T^{n+1} = value on the Dirichlet rows
    with the sparse system solved directly (SuperLU).  Pinned by the analytic 1-D cooling test (tests/test_oracle_dynamic.py)."""

    def __init__(self, mesh, elem_mat, tables, cprho, dirichlet_nodes, dirichlet_values, heat=None, inittemp=300.,
                 timestep=0.1, methodparam=0.5, lumping=True, rebuildfreq=0):
        import scipy.sparse as sp  # noqa: F401  (fail early if scipy is missing)
        self.mesh, self.tables = mesh, tables
        self.elem_mat = np.ascontiguousarray(elem_mat, dtype=np.uint32)
        self.cprho = np.ascontiguousarray(cprho, dtype=np.float64)
        self.bc_nodes = np.ascontiguousarray(dirichlet_nodes, dtype=np.int64)
        self.bc_values = np.ascontiguousarray(dirichlet_values, dtype=np.float64)
        self.heat = np.zeros(mesh.E) if heat is None else np.ascontiguousarray(heat, dtype=np.float64)
        self.inittemp, self.timestep, self.methodparam = float(inittemp), float(timestep), float(methodparam)
        self.lumping, self.rebuildfreq = bool(lumping), int(rebuildfreq)
        self.temperatures = np.full(mesh.N, float(inittemp))    # :82
        self.elapstime = 0.          # the reference's clock: advanced by steps*timestep - timestep per call (:293,297)
        self.physical_time = 0.      # time the field really advanced: one timestep per solve
        self.conds = np.zeros((mesh.E, 2))
        self.maxT_log = []
        m = mesh
        eg = m.elems_grid().ravel()
        i0, i1, i2 = np.meshgrid(*[np.arange(k - 1) for k in m.n], indexing="ij")
        base = (i0 * m.ns[0] + i1 * m.ns[1] + i2 * m.ns[2]).ravel()
        # idx[l]: bit 0 -> axis 0, bit 1 -> axis 1, bit 2 -> axis 2 (:143-151)
        self._idx = np.empty((m.E, 8), dtype=np.int64)
        for l in range(8):
            self._idx[eg, l] = base + (l & 1) * m.ns[0] + ((l >> 1) & 1) * m.ns[1] + ((l >> 2) & 1) * m.ns[2]
        d = [np.diff(a) for a in m.axes]
        self._dx = np.empty(m.E); self._dy = np.empty(m.E); self._dz = np.empty(m.E)
        self._dx[eg] = np.broadcast_to(d[0][:, None, None], i0.shape).ravel()
        self._dy[eg] = np.broadcast_to(d[1][None, :, None], i0.shape).ravel()
        self._dz[eg] = np.broadcast_to(d[2][None, None, :], i0.shape).ravel()

    def set_matrix(self):
        """setMatrix, femT3d.cpp:127-255 -> (A, B, F) as scipy CSR / vector, before the Dirichlet elimination"""
        import scipy.sparse as sp
        m, t = self.mesh, self.tables
        lib().orc_thermal_conds(m.ref, _p(self.temperatures), _p(self.elem_mat, C.c_uint32), C.c_uint32(t.nT),
                                C.c_double(t.T0), C.c_double(t.dT), _p(t.lat), _p(t.vert), _p(self.conds))
        dx, dy, dz = self._dx, self._dy, self._dz
        ky = self.conds[:, 0] * 1e-6
        kz = self.conds[:, 1] * 1e-6
        kx = ky / dx * dy * dz                 # :170-172
        ky = ky * dx / dy * dz
        kz = kz * dx * dy / dz
        kv = np.empty((m.E, 8))                # K[i][j] by i ^ j (:186-199)
        kv[:, 0] = (kx + ky + kz) / 9.
        kv[:, 1] = (-2. * kx + ky + kz) / 18.
        kv[:, 2] = (kx - 2. * ky + kz) / 18.
        kv[:, 4] = (kx + ky - 2. * kz) / 18.
        kv[:, 6] = (kx - 2. * ky - 2. * kz) / 36.
        kv[:, 5] = (-2. * kx + ky - 2. * kz) / 36.
        kv[:, 3] = (-2. * kx - 2. * ky + kz) / 36.
        kv[:, 7] = -(kx + ky + kz) / 36.
        temp = self.temperatures[self._idx].sum(axis=1) * 0.125       # :163
        c = table_lookup(self.cprho, self.elem_mat, t.T0, t.dT, temp) * 0.125e-9 * dx * dy * dz / self.timestep   # :176
        f = 0.125e-18 * dx * dy * dz * self.heat                       # :179
        rows, cols, ka, cm = [], [], [], []
        pop = np.array([bin(x).count("1") for x in range(8)])
        for i in range(8):
            for j in range(8):
                rows.append(self._idx[:, i]); cols.append(self._idx[:, j]); ka.append(kv[:, i ^ j])
                if self.lumping:
                    cm.append(c if i == j else np.zeros_like(c))                  # :207-212
                else:
                    cm.append(c * (8 >> pop[i ^ j]) / 27.)                        # :216-231
        rows, cols = np.concatenate(rows), np.concatenate(cols)
        K = sp.coo_matrix((np.concatenate(ka), (rows, cols)), shape=(m.N, m.N)).tocsr()
        Cm = sp.coo_matrix((np.concatenate(cm), (rows, cols)), shape=(m.N, m.N)).tocsr()
        F = np.zeros(m.N)
        np.add.at(F, self._idx.ravel(), np.repeat(f, 8))
        th = self.methodparam
        return (th * K + Cm).tocsr(), (-(1. - th) * K + Cm).tocsr(), F     # :203-204

    def _factor(self):
        import scipy.sparse as sp
        import scipy.sparse.linalg as spl
        A, B, F = self.set_matrix()
        fixed = np.zeros(self.mesh.N, dtype=bool)
        fixed[self.bc_nodes] = True
        xD = np.zeros(self.mesh.N)
        xD[self.bc_nodes] = self.bc_values            # last value wins for duplicated nodes
        free = ~fixed
        Aff = A[free][:, free].tocsc()
        lift = A[free][:, fixed] @ xD[fixed]           # applyBC: F[c] -= A(c,r) * value (:236)
        return dict(solve=spl.factorized(Aff), B=B, F=F, free=free, fixed=fixed, xD=xD, lift=lift)

    def compute(self, time):
        """compute(time), femT3d.cpp:258-305 with the corrected update"""
        sysm = self._factor()
        r = self.rebuildfreq
        tend = time + self.timestep / 2.
        t = 0.
        while t < tend:
            if self.rebuildfreq and r == 0:
                sysm = self._factor()
                r = self.rebuildfreq
            rhs = sysm["B"] @ self.temperatures + sysm["F"]
            free = sysm["free"]
            Tn = np.empty_like(self.temperatures)
            Tn[free] = sysm["solve"](rhs[free] - sysm["lift"])
            Tn[sysm["fixed"]] = sysm["xD"][sysm["fixed"]]
            self.temperatures = Tn
            self.maxT_log.append(float(Tn.max()))
            r -= 1
            self.elapstime += self.timestep
            self.physical_time += self.timestep
            t += self.timestep
        self.elapstime -= self.timestep            # :297 — the loop runs time/timestep + 1 solves, the clock shows `time`
        self.maxT = float(self.temperatures.max())
        return 0.


class Shockley3DOracle:
    """ElectricalFem3DSolver + BetaSolver<Geometry3D> restated
    (solvers/electrical/shockley/electr3d.cpp, beta.hpp)."""

    def __init__(self, mesh, elem_mat, tables, dirichlet_nodes, dirichlet_values, elem_junc=None, elem_role=None,
                 beta=None, js=None, pcond=5., ncond=50., start_cond=(0., 5.), maxerr=0.05, convergence="fast",
                 algorithm="cholesky", precond="ic", itmaxerr=1e-6, maxit=1000, nfact=10, noheat=None, eps=None,
                 included=None):
        self.mesh = mesh
        E = mesh.E
        # empty-elements="exclude": uint8 [E], 0 = element outside the masked mesh (Cholesky only); None = full mesh
        self.included = None if included is None else np.ascontiguousarray(included, dtype=np.uint8)
        self.elem_mat = np.ascontiguousarray(elem_mat, dtype=np.uint32)
        self.tables = tables
        self.bc_nodes = np.ascontiguousarray(dirichlet_nodes, dtype=np.uintp)
        self.bc_values = np.ascontiguousarray(dirichlet_values, dtype=np.float64)
        self.elem_junc = np.zeros(E, dtype=np.uint32) if elem_junc is None else np.ascontiguousarray(elem_junc, dtype=np.uint32)
        self.elem_role = np.zeros(E, dtype=np.uint8) if elem_role is None else np.ascontiguousarray(elem_role, dtype=np.uint8)
        self.noheat = None if noheat is None else np.ascontiguousarray(noheat, dtype=np.uint8)
        self.eps = None if eps is None else np.ascontiguousarray(eps, dtype=np.float64)
        self.pcond, self.ncond, self.maxerr = pcond, ncond, maxerr
        self.stable = 1 if convergence == "stable" else 0
        self.algorithm, self.precond, self.itmaxerr, self.maxit, self.nfact = algorithm, precond, itmaxerr, maxit, nfact
        # setupActiveRegions, electr3d.cpp:89-183
        max_act = int(self.elem_junc.max()) if E else 0
        self.active = (_Active * max(max_act, 1))()
        condsize = c_sz(0)
        self.nact = lib().orc_setup_active(mesh.ref, _p(self.elem_junc, C.c_uint32), self.active, C.c_int(max_act),
                                           C.byref(condsize))
        if self.nact < 0:
            raise ValueError("junction does not have top and bottom edges at constant heights")
        self.junction_conductivity = np.tile(np.asarray(start_cond, dtype=np.float64), (max(condsize.value, 1), 1))
        self.beta, self.js = beta, js  # per junction: float or callable(T)
        # onInitialize, electr3d.cpp:185-193
        self.potential = np.zeros(mesh.N)
        self.current = np.zeros((E, 3))
        self.conds = np.zeros((E, 2))
        self.heat = None
        self.loopno = 0
        self.converged = True
        self.history = []
        self.maxcur = np.zeros(3)
        self.Te = np.full(E, 300.)  # inTemperature default
        self._A = None

    def _matrix(self):
        if self._A is None:
            if self.included is not None:
                assert self.algorithm == "cholesky", "the masked mesh is restated for the Cholesky path only"
                self._A = DpbMasked(self.mesh, self.included)
            else:
                self._A = Dpb(self.mesh) if self.algorithm == "cholesky" else Sparse14(self.mesh)
        return self._A

    def _junction_params(self):
        """beta/js per junction-table entry, evaluated at the mid-plane element temperature
        (temperature[tidx], electr3d.cpp:261-262; python callables electr_python.cpp:103-110)."""
        n = len(self.junction_conductivity)
        bcol, jcol = np.ones(n), np.ones(n)
        for k in range(self.nact):
            a = self.active[k]
            v = (a.top + a.bottom) // 2
            b = self.beta[k] if isinstance(self.beta, (list, tuple)) else self.beta
            j = self.js[k] if isinstance(self.js, (list, tuple)) else self.js
            for t in range(a.left, a.right):
                for l in range(a.back, a.front):
                    T = self.Te[l * self.mesh.es[0] + t * self.mesh.es[1] + v * self.mesh.es[2]]
                    col = a.offset + a.ld * t + l
                    bcol[col] = b(T) if callable(b) else b
                    jcol[col] = j(T) if callable(j) else j
        return bcol, jcol

    def load_conductivity(self):
        t = self.tables
        lib().orc_shockley_load_conds(self.mesh.ref, _p(self.elem_mat, C.c_uint32), _p(self.elem_junc, C.c_uint32),
                                      _p(self.elem_role, C.c_uint8), _p(self.Te), C.c_uint32(t.nT), C.c_double(t.T0),
                                      C.c_double(t.dT), _p(t.lat), _p(t.vert), self.active,
                                      _p(self.junction_conductivity), C.c_double(self.pcond), C.c_double(self.ncond),
                                      _p(self.conds))
        if self.included is not None:
            self.conds[self.included == 0] = 0.   # elements outside the masked mesh do not exist: no current, no heat

    def compute(self, loops=0):
        """compute, electr3d.cpp:356-442."""
        A = self._matrix()
        loop, toterr = 0, 0.
        rhs = np.zeros(self.mesh.N)
        self.load_conductivity()
        bcol, jcol = self._junction_params()
        noactive = 1 if self.nact == 0 else 0
        minj = 100e-7
        self.heat = None
        while True:
            # setMatrix, electr3d.cpp:239-354
            if self.loopno != 0:
                lib().orc_shockley_junction_update(self.mesh.ref, _p(self.elem_junc, C.c_uint32), self.active,
                                                   _p(self.potential), _p(bcol), _p(jcol), C.c_int(self.stable),
                                                   _p(self.conds))
            A.assemble(self.conds, None, rhs)
            A.apply_bc(rhs, self.bc_nodes, self.bc_values)
            if self.algorithm == "cholesky":
                info = A.solve(rhs, self.potential)
            elif self.algorithm == "iterative":
                info = A.solve_nspcg(rhs, self.potential, precond=self.precond, maxit=self.maxit,
                                     maxerr=self.itmaxerr, nfact=self.nfact)
            else:
                info = A.solve_pcg(rhs, self.potential, maxit=self.maxit, tol=self.itmaxerr)
            self.converged = info["converged"]
            mcur = C.c_double(0)
            err = lib().orc_shockley_currents(self.mesh.ref, _p(self.elem_junc, C.c_uint32), C.c_int(noactive),
                                              _p(self.potential), _p(self.conds), _p(self.current), C.byref(mcur),
                                              _p(self.maxcur))
            if (loop != 0 or mcur.value >= minj) and err > toterr:
                toterr = err
            self.loopno += 1
            loop += 1
            self.history.append(dict(loop=loop, mcur=mcur.value, err=err, iters=info["iters"]))
            if not ((not self.converged or err > self.maxerr) and (loops == 0 or loop < loops)):
                break
        lib().orc_shockley_save_conds(self.mesh.ref, self.active, C.c_int(self.nact), _p(self.conds),
                                      _p(self.junction_conductivity))
        return toterr

    def heat_density(self):
        if self.heat is None:
            self.heat = np.zeros(self.mesh.E)
            lib().orc_shockley_heat(self.mesh.ref, _p(self.potential), _p(self.conds),
                                    _p(self.noheat, C.c_uint8) if self.noheat is not None else None, _p(self.heat))
        return self.heat

    def get_total_current(self, nact=0):
        a = self.active[nact]
        level = (a.bottom + a.top) // 2
        return lib().orc_integrate_current(self.mesh.ref, _p(self.elem_junc, C.c_uint32), _p(self.current),
                                           c_sz(level), C.c_int(1))

    def get_total_heat(self):
        return lib().orc_total_heat(self.mesh.ref, _p(self.heat_density()))

    def get_total_energy(self):
        eps = self.eps
        if self.included is not None:   # getTotalEnergy loops over maskedMesh->elements() (electr3d.cpp:577-600)
            eps = np.ascontiguousarray(np.where(self.included != 0, eps, 0.))
        return lib().orc_total_energy(self.mesh.ref, _p(self.potential), _p(eps))

    def get_capacitance(self):
        """getCapacitance, electr3d.cpp:602-610 (exactly two voltage conditions)."""
        vals = np.unique(self.bc_values)
        assert len(vals) == 2
        U = vals[1] - vals[0]
        return 2e12 * self.get_total_energy() / (U * U)


# ------------------------------------------------------------------------- ThermoElectric3D


def midpoints(a):
    """MidpointAxis::at (plask/mesh/axis1d.cpp:51-53)."""
    a = np.asarray(a, dtype=np.float64)
    return (a[:-1] + a[1:]) * 0.5


def interp_linear(src_axes, src_strides, src, dst_axes, dst_strides, dst_size):
    """RectilinearMesh3D::interpolateLinear onto the tensor-product points dst_axes (rectilinear3d.hpp:802-845)."""
    sa = [np.ascontiguousarray(a, dtype=np.float64) for a in src_axes]
    da = [np.ascontiguousarray(a, dtype=np.float64) for a in dst_axes]
    out = np.zeros(dst_size)
    sn = (c_sz * 3)(*[len(a) for a in sa])
    dn = (c_sz * 3)(*[len(a) for a in da])
    ss = (c_sz * 3)(*[int(s) for s in src_strides])
    ds = (c_sz * 3)(*[int(s) for s in dst_strides])
    lib().orc_interp_linear(sn, _p(sa[0]), _p(sa[1]), _p(sa[2]), ss, _p(np.ascontiguousarray(src, dtype=np.float64)),
                            dn, _p(da[0]), _p(da[1]), _p(da[2]), ds, _p(out))
    return out


class ThermoElectric3DOracle:
    """meta.shockley.ThermoElectric3D (solvers/meta/shockley/thermoelectric.py:187-211) over the two oracles:
    electrical.inTemperature = thermal.outTemperature (linear interpolation at the electrical element midpoints,
    therm3d.cpp:385-393, electr3d.cpp:203-205); thermal.inHeat = electrical.outHeat (linear interpolation of the
    element-mesh data at the thermal element midpoints, electr3d.cpp:538-548, therm3d.cpp:179)."""

    def __init__(self, thermal, electrical, tfreq=6):
        self.thermal, self.electrical, self.tfreq = thermal, electrical, tfreq
        self.history = []

    def exchange_temperature(self):
        t, e = self.thermal, self.electrical
        mids = [midpoints(a) for a in e.mesh.axes]
        e.Te = interp_linear(t.mesh.axes, t.mesh.ns, t.temperatures, mids, e.mesh.es, e.mesh.E)
        if t.included is not None:
            # thermal solver on a masked mesh: interpolate(maskedMesh, ...) is NaN outside the kept elements and
            # SafeData<double>(..., 300.) substitutes 300 K (getTemperatures, therm3d.cpp:391-392)
            idx, inside = [], []
            for a in range(3):
                ax = t.mesh.axes[a]
                up = np.searchsorted(ax, mids[a], side="right")
                inside.append((up > 0) & (up < len(ax)))
                idx.append(np.clip(up - 1, 0, len(ax) - 2))
            te = idx[0][:, None, None] * t.mesh.es[0] + idx[1][None, :, None] * t.mesh.es[1] + idx[2][None, None, :] * t.mesh.es[2]
            ok = inside[0][:, None, None] & inside[1][None, :, None] & inside[2][None, None, :] & (t.included[te] != 0)
            ee = e.mesh.elems_grid()
            e.Te[ee[~ok]] = 300.

    def exchange_heat(self):
        t, e = self.thermal, self.electrical
        e.heat = None
        heat = e.heat_density()
        mids = [midpoints(a) for a in t.mesh.axes]
        t.heat = interp_linear([midpoints(a) for a in e.mesh.axes], e.mesh.es, heat, mids, t.mesh.es, t.mesh.E)
        # getHeatDensity returns 0 outside the bounding box of the electrical geometry (electr3d.cpp:545-548; bounds included);
        # the box is taken as the extent of the electrical mesh
        ins = [(m >= e.mesh.axes[a][0]) & (m <= e.mesh.axes[a][-1]) for a, m in enumerate(mids)]
        ok = ins[0][:, None, None] & ins[1][None, :, None] & ins[2][None, None, :]
        t.heat[t.mesh.elems_grid()[~ok]] = 0.

    def compute(self, max_meta_loops=100):
        t, e = self.thermal, self.electrical
        verr, terr = 2. * e.maxerr, 2. * t.maxerr
        n = 0
        while (terr > t.maxerr or verr > e.maxerr) and n < max_meta_loops:
            self.exchange_temperature()
            verr = e.compute(self.tfreq)
            self.exchange_heat()
            terr = t.compute(1)
            n += 1
            self.history.append(dict(verr=verr, terr=terr, maxT=t.maxT, current=e.get_total_current() if e.nact else 0.))
        return n
