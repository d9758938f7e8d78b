"""Host-side mirror of electrical.diffusion.Diffusion3D for the CUDA algorithm (last item).

`Diffusion3D` keeps the names, defaults and semantics of Diffusion3DSolver (solvers/electrical/diffusion/diffusion3d.{hpp,cpp},
python/diffusion.cpp:185-210): `compute(loops=0, shb=False, act=None)`, `maxerr` [%], `inCurrentDensity`, `inTemperature`, `inGain`,
`inWavelength`, `inLightE`, `outCarriersConcentration`, `get_total_burning()`, `get_burning_for_mode(m)`.  As with the other mirrors
the geometry tree is out of scope: the solver is bound to a `DiffusionProblem`, the flat description of what
setupActiveRegions (diffusion3d.cpp:89-178) and ActiveRegion3D (diffusion3d.hpp:37-106) extract from geometry and mesh.

Everything the reference does per element or node inside its loop — element integrals, residual, linear solve, spline evaluation —
runs on the device behind include/plaskdiff_cuda.h; there is no CPU path.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib as L

QE = 1.60217733e-19            # phys::qe   (plask/phys/constants.hpp:31)
C_LIGHT = 299792458.           # phys::c
Z0 = 376.73031346177066        # phys::Z0
H_J = 6.62606957e-34           # phys::h_J
INV_HC = 1.0e-13 / (C_LIGHT * H_J)     # diffusion3d.cpp:22


def _dp(a):
    return a.ctypes.data_as(L.c_dp)


@dataclass
class ActiveRegion:
    """One ActiveRegion3D: `mask[n0-1, n1-1]` = elements of the lateral mesh with role QW / QD / carriers at the height of the
    middle well (diffusion3d.hpp:73-80); `qws` = [(z_bottom, z_top)] of the wells; A, B, C [1/s, cm^3/s, cm^6/s] and D [cm^2/s] are
    either constants, arrays per element [n0-1, n1-1] or callables of the element temperatures (material->A(T) ...)."""
    mask: np.ndarray
    qws: list
    A: object = 3e7
    B: object = 1.7e-10
    C: object = 6e-27
    D: object = 10.
    nr: object = 3.5           # material->Nr(wavelength, T).real() for the SHB terms: constant, array or callable(wavelength, T)

    @property
    def qw_height(self):       # QWheight (diffusion3d.cpp:158-166)
        return float(sum(t - b for b, t in self.qws))

    @property
    def qw_z(self):            # heights of the wells' element centres (QWz)
        return [0.5 * (b + t) for b, t in self.qws]

    @property
    def vert(self):            # ActiveRegion3D::vert(): the middle well (diffusion3d.hpp:72)
        z = self.qw_z
        return z[(len(z) + 1) // 2 - 1]


@dataclass
class DiffusionProblem:
    ax0: np.ndarray            # lateral axes of the solver mesh [um]
    ax1: np.ndarray
    regions: list = field(default_factory=list)
    order: int = L.DIFF_ORDER_01

    @property
    def n(self):
        return len(self.ax0), len(self.ax1)

    def node_points(self, z):
        """mesh2 / mesh3 points in the mesh's own order, [nn, 3]"""
        X, Y = np.meshgrid(self.ax0, self.ax1, indexing="ij")
        if self.order == L.DIFF_ORDER_10:
            X, Y = X.T, Y.T
        return np.stack([X.ravel(), Y.ravel(), np.full(X.size, z)], axis=1)

    def elem_points(self, z):
        xm, ym = 0.5 * (self.ax0[1:] + self.ax0[:-1]), 0.5 * (self.ax1[1:] + self.ax1[:-1])
        X, Y = np.meshgrid(xm, ym, indexing="ij")
        if self.order == L.DIFF_ORDER_10:
            X, Y = X.T, Y.T
        return np.stack([X.ravel(), Y.ravel(), np.full(X.size, z)], axis=1)

    def elem_flat(self, a):
        """[n0-1, n1-1] array -> the mesh's element order"""
        a = np.asarray(a)
        return (a.T if self.order == L.DIFF_ORDER_10 else a).ravel()


class DeviceDiffusion:
    """thin RAII wrapper of a pdiff_ctx"""

    def __init__(self, device=0):
        self.lib = L.load()
        self.ctx = L._vp()
        rc = self.lib.pdiff_create(C.byref(self.ctx), device)
        if rc != 0:
            L.check(None, rc)

    def close(self):
        if self.ctx:
            self.lib.pdiff_destroy(self.ctx)
            self.ctx = L._vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        return L.check(self.ctx, rc, "pdiff_last_error")

    def set_mesh(self, ax0, ax1, order, active):
        ax0, ax1 = np.ascontiguousarray(ax0, dtype=np.float64), np.ascontiguousarray(ax1, dtype=np.float64)
        self.n0, self.n1 = len(ax0), len(ax1)
        self.nn, self.ne = self.n0 * self.n1, (self.n0 - 1) * (self.n1 - 1)
        act = None if active is None else np.ascontiguousarray(active, dtype=np.uint8)
        self._ck(self.lib.pdiff_set_mesh(self.ctx, self.n0, self.n1, _dp(ax0), _dp(ax1), order,
                                         None if act is None else act.ctypes.data_as(L._u8p)))

    def set_parameters(self, A, B, Cc, D):
        arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (self.ne,))) for a in (A, B, Cc, D)]
        self._ck(self.lib.pdiff_set_parameters(self.ctx, *[_dp(a) for a in arrs]))

    def set_current(self, J):
        J = np.ascontiguousarray(np.broadcast_to(np.asarray(J, dtype=np.float64), (self.nn,)))
        self._ck(self.lib.pdiff_set_current(self.ctx, _dp(J)))

    def set_modes(self, P=None, G=None, dG=None):
        if P is None or len(P) == 0:
            self._ck(self.lib.pdiff_set_modes(self.ctx, 0, None, None, None))
            return
        P, G, dG = (np.ascontiguousarray(a, dtype=np.float64) for a in (P, G, dG))
        assert P.shape[1:] == (self.nn, 2) and G.shape == (P.shape[0], self.ne, 2) and dG.shape == G.shape
        self._ck(self.lib.pdiff_set_modes(self.ctx, P.shape[0], _dp(P), _dp(G), _dp(dG)))

    def set_concentration(self, U=None):
        if U is None:
            self._ck(self.lib.pdiff_set_concentration(self.ctx, None))
        else:
            U = np.ascontiguousarray(U, dtype=np.float64)
            assert U.size == 3 * self.nn
            self._ck(self.lib.pdiff_set_concentration(self.ctx, _dp(U)))

    def get_concentration(self):
        U = np.empty(3 * self.nn)
        self._ck(self.lib.pdiff_get_concentration(self.ctx, _dp(U)))
        return U

    def compute(self, loops=0, maxerr=0.05, maxit=20000, lin_tol=1e-12, verbatim=True):
        o, st = L.DiffOpts(), L.DiffStats()
        self.lib.pdiff_default_opts(C.byref(o))
        o.loops, o.maxerr, o.maxit, o.lin_tol, o.verbatim = loops, maxerr, maxit, lin_tol, int(verbatim)
        rc = self._ck(self.lib.pdiff_compute(self.ctx, C.byref(o), C.byref(st)))
        d = st.as_dict()
        d["status"] = rc
        return d

    def interpolate(self, x, y, method=L.DIFF_INTERP_SPLINE):
        x, y = np.ascontiguousarray(x, dtype=np.float64).ravel(), np.ascontiguousarray(y, dtype=np.float64).ravel()
        out = np.empty(x.size)
        self._ck(self.lib.pdiff_interpolate(self.ctx, x.size, _dp(x), _dp(y), method, _dp(out)))
        return out

    def element_matrices(self, verbatim=True):
        K, F = np.empty((self.ne, 12, 12)), np.empty((self.ne, 12))
        self._ck(self.lib.pdiff_get_element_matrices(self.ctx, int(verbatim), _dp(K), _dp(F)))
        return K, F

    def apply(self, v, verbatim=True):
        v = np.ascontiguousarray(v, dtype=np.float64)
        y = np.empty(3 * self.nn)
        self._ck(self.lib.pdiff_apply(self.ctx, int(verbatim), _dp(v), _dp(y)))
        return y

    def rhs(self, verbatim=True):
        F = np.empty(3 * self.nn)
        self._ck(self.lib.pdiff_get_rhs(self.ctx, int(verbatim), _dp(F)))
        return F


def _eval(f, pts, default):
    """a receiver: None -> default, constant, or callable(points[n,3]) -> values"""
    if f is None:
        f = default
    if callable(f):
        return np.asarray(f(pts))
    f = np.asarray(f)
    return np.broadcast_to(f, (len(pts),) + f.shape)


class Diffusion3D:
    def __init__(self, name=""):
        self.id = name
        self.algorithm = "cuda"
        self.device = 0
        self.maxerr = 0.05                     # diffusion3d.cpp:28
        self.maxit, self.lin_tol = 20000, 1e-12
        self.verbatim = True                   # Ug and the burned power as written in the reference (see plaskdiff_cuda.h)
        self.inCurrentDensity = None           # callable(points) -> [n,3] kA/cm^2, or a constant vector
        self.inTemperature = None              # callable(points) -> [n] K; default 300 K
        self.inGain = None                     # callable(points, wavelength, deriv: bool) -> [n,2] (c00, c11)
        self.inWavelength = []                 # [nm] per mode
        self.inLightE = None                   # list (per mode) of callables(points) -> complex [n,3]
        self.noconv = "warning"
        self._problem = None
        self._dev = {}
        self.modesP = {}
        self.loopno = 0
        self.stats = {}
        self.log = []

    @property
    def problem(self):
        return self._problem

    @problem.setter
    def problem(self, p):
        assert isinstance(p, DiffusionProblem)
        self._problem = p
        self.invalidate()

    def invalidate(self):
        """onInvalidate: active.clear() (diffusion3d.cpp:194)"""
        for d in self._dev.values():
            d.close()
        self._dev = {}
        self.modesP = {}
        self.loopno = 0

    def _device(self, act):
        if act not in self._dev:
            p, reg = self._problem, self._problem.regions[act]
            d = DeviceDiffusion(self.device)
            d.set_mesh(p.ax0, p.ax1, p.order, p.elem_flat(reg.mask))
            self._dev[act] = d
        return self._dev[act]

    def _vertical_average(self, recv, pts_of_z, reg, default):
        """ActiveRegion3D::verticallyAverage (diffusion3d.hpp:92-105): mean over the wells' heights"""
        vals = [_eval(recv, pts_of_z(z), default) for z in reg.qw_z]
        return sum(vals) / len(vals)

    def compute(self, loops=0, shb=False, act=None):
        if self.algorithm != "cuda":
            raise L.BadInput(f"{self.id}: algorithm '{self.algorithm}' is not provided by plask_b200; use 'cuda'")
        if self._problem is None:
            raise L.BadInput(f"{self.id}: no problem (geometry and mesh) set")
        if act is None:
            for a in range(len(self._problem.regions)):
                self.compute(loops, shb, a)
            return 0.
        if not 0 <= act < len(self._problem.regions):
            raise L.ComputationError(f"{self.id}: Active region {act} does not exist")
        p, reg = self._problem, self._problem.regions[act]
        d = self._device(act)
        # material parameters at the vertically averaged temperature (diffusion3d.cpp:222-230)
        T = self._vertical_average(self.inTemperature, p.elem_points, reg, 300.)

        def par(v):
            return np.asarray(v(T)) if callable(v) else p.elem_flat(v) if np.ndim(v) == 2 else v
        d.set_parameters(par(reg.A), par(reg.B), par(reg.C), 1e8 * np.asarray(par(reg.D), dtype=float))
        # J = |js j_z| (diffusion3d.cpp:232-238)
        if self.inCurrentDensity is None:
            raise L.BadInput(f"{self.id}: no provider connected to inCurrentDensity")
        j = _eval(self.inCurrentDensity, p.node_points(reg.vert), None)
        d.set_current(np.abs(1e7 / (QE * reg.qw_height) * j[:, 2]))
        if shb:
            nm = len(self.inWavelength)
            if self.inLightE is None or len(self.inLightE) != nm:
                raise L.BadInput(f"{self.id}: number of modes in inWavelength ({nm}) and inLightE "
                                 f"({0 if self.inLightE is None else len(self.inLightE)}) differ")
            P, G, dG = np.zeros((nm, d.nn, 2)), np.zeros((nm, d.ne, 2)), np.zeros((nm, d.ne, 2))
            ep = p.elem_points(reg.vert)
            power = []
            for m, wl in enumerate(self.inWavelength):
                wl = float(np.real(wl))
                E = self._vertical_average(self.inLightE[m], p.node_points, reg, None)
                P[m, :, 0] = (0.5 / Z0) * (np.abs(E[:, 0])**2 + np.abs(E[:, 1])**2)       # diffusion3d.cpp:264-271
                P[m, :, 1] = (0.5 / Z0) * np.abs(E[:, 2])**2
                nr = np.asarray(reg.nr(wl, T)) if callable(reg.nr) else p.elem_flat(reg.nr) if np.ndim(reg.nr) == 2 else reg.nr
                nr = np.broadcast_to(np.asarray(nr, dtype=float), (d.ne,))[:, None]
                g = nr * np.asarray(self.inGain(ep, wl, False))
                dg = nr * np.asarray(self.inGain(ep, wl, True))
                power.append(burned_power(p, reg, P[m], g, self.verbatim))
                G[m], dG[m] = INV_HC * wl * g, INV_HC * wl * dg
            d.set_modes(P, G, dG)
            self.modesP[act] = power
        else:
            d.set_modes()
        st = d.compute(loops, self.maxerr, self.maxit, self.lin_tol, self.verbatim)
        if st["status"] == L.PFEM_NOT_CONVERGED and self.noconv == "error":
            raise L.ComputationError(f"{self.id}: linear solver did not converge in {self.maxit} iterations")
        self.stats = st
        self.loopno += st["loops"]
        for k, e in enumerate(st["err_log"][1:], 1):
            self.log.append(f"Loop {k}({self.loopno - st['loops'] + k}) @ active region {act}: error = {e:g}%")
        return 0.          # the reference returns toterr, which it never updates (diffusion3d.cpp:274,371)

    def get_burning_for_mode(self, mode):
        if self.inLightE is None or mode >= len(self.inLightE):
            raise L.BadInput(f"{self.id}: mode index out of range")
        res = 0.
        for act in range(len(self._problem.regions)):
            if act not in self.modesP or mode >= len(self.modesP[act]):
                raise L.ComputationError(f"{self.id}: SHB not computed for active region {act}")
            res += self.modesP[act][mode]
        return res

    def get_total_burning(self):
        return sum(self.get_burning_for_mode(m) for m in range(len(self.inLightE or [])))

    def outCarriersConcentration(self, points, interpolation="spline"):
        """ConcentrationDataImpl (diffusion3d.cpp:408-484): points [n,3], already wrapped into the mesh by the geometry's symmetry;
        concentration exists only inside the wells' vertical ranges"""
        pts = np.atleast_2d(np.asarray(points, dtype=float))
        out = np.zeros(len(pts))
        if interpolation not in ("spline", "default", "linear"):
            raise L.BadInput(f"{self.id}: interpolation must be 'spline' or 'linear'")
        method = L.DIFF_INTERP_LINEAR if interpolation == "linear" else L.DIFF_INTERP_SPLINE
        for act, reg in enumerate(self._problem.regions):
            if act not in self._dev:
                raise L.ComputationError(f"{self.id}: no value of carriers concentration")
            inside = np.zeros(len(pts), dtype=bool)
            for b, t in reg.qws:
                inside |= (b <= pts[:, 2]) & (pts[:, 2] < t)
            if inside.any():
                out[inside] = self._dev[act].interpolate(pts[inside, 0], pts[inside, 1], method)
        return out


def burned_power(p, reg, P, g, verbatim=True):
    """modesP of one mode (diffusion3d.cpp:295-303): sum over the elements of integrateBilinear(P).g times 1e-13 QWheight.
    It does not involve the unknowns, so it stays on the host.  verbatim=True reproduces the reference as written:
    integrateBilinear (diffusion3d.hpp:170-172) multiplies by Lx*Lx and is handed `Pdata + ie`, i.e. the four CONSECUTIVE nodal
    values starting at the element's index — both in the numbering of the MASKED lateral mesh; verbatim=False takes the four
    corner nodes and Lx*Ly.  P [nn,2] and g [ne,2] are in the full-grid numbering of the ABI."""
    n0, n1 = p.n
    mask = np.asarray(reg.mask, dtype=bool)
    X, Y = np.diff(p.ax0)[:, None] * np.ones((1, n1 - 1)), np.ones((n0 - 1, 1)) * np.diff(p.ax1)[None, :]
    i0, i1 = np.meshgrid(np.arange(n0 - 1), np.arange(n1 - 1), indexing="ij")
    if p.order == L.DIFF_ORDER_10:
        eidx, nidx, s0, s1 = i0 + (n0 - 1) * i1, i0 + n0 * i1, 1, n0
    else:
        eidx, nidx, s0, s1 = (n1 - 1) * i0 + i1, n1 * i0 + i1, n1, 1
    # flatten in the mesh's element order, keep the elements of the masked mesh
    o = np.argsort(eidx.ravel())
    keep = mask.ravel()[o]
    eidx, nidx, X, Y = (a.ravel()[o][keep] for a in (eidx, nidx, X, Y))
    g = np.asarray(g, dtype=float)[eidx]
    P = np.asarray(P, dtype=float)
    if verbatim:
        nact = np.zeros(n0 * n1, dtype=bool)
        for off in (0, s0, s1, s0 + s1):
            nact[nidx + off] = True
        Pm = np.vstack([P[nact], np.zeros((4, 2))])          # nodal values in the masked numbering
        ie = np.arange(len(eidx))                            # element index in the masked numbering
        pe = 0.25 * (Pm[ie] + Pm[ie + 1] + Pm[ie + 2] + Pm[ie + 3]) * (X * X)[:, None]
    else:
        pe = 0.25 * (P[nidx] + P[nidx + s0] + P[nidx + s1] + P[nidx + s0 + s1]) * (X * Y)[:, None]
    return float((pe[:, 0] * g[:, 0] + pe[:, 1] * g[:, 1]).sum() * 1e-13 * reg.qw_height)
