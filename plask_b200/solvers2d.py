"""The two-dimensional solvers on the brick kernels: thermal.static.Static2D / StaticCyl
(solvers/thermal/static/therm2d.cpp), thermal.dynamic.Dynamic2D / DynamicCyl (solvers/thermal/dynamic/femT2d.cpp) and
electrical.shockley.Shockley2D / ShockleyCyl (solvers/electrical/shockley/electr2d.cpp) with algorithm='cuda'.

A rectangular 2-D mesh (x = tran or r, y = vert) is handed to the library as a brick mesh with ONE element layer along a dummy
longitudinal axis and z-invariant data.  For a z-invariant field the brick operator of the layer (thickness d) reduces on each of
its two node planes to d/2 * 1e-6 times the 4-node rectangle operator of therm2d.cpp:206-222 / electr2d.cpp:296-316
(M_z rows sum to 1/2, S_z annihilates constants), and the load vector (0.125e-18 dx dy dz heat against 0.25e-12 w h heat,
therm3d.cpp:222 / therm2d.cpp:210) carries the same factor — so the brick solution restricted to one plane IS the 2-D FEM
solution, Dirichlet rows being fixed on both planes.  The cylindrical solvers multiply every element matrix and load by the
midpoint radius r (therm2d.cpp:353,412-426, electr2d.cpp:219-230): pfem_set_axis_weight on the radial axis.  Currents, heat
fluxes and Joule heat are element gradients and come out identical (the longitudinal component is zero); the integrals
(total current, heat, energy, capacitance) use the 2-D formulas with the extrusion length or 2 pi r.

Boundary conditions of the 2nd / 3rd kind and radiation (therm2d.cpp:138-172, :225-265 Cartesian, :371-413 cylindrical) live on
element EDGES of the 2-D mesh; brick faces do not reproduce them (the cylindrical terms carry r -+ len/6 per node, the
convection matrix terms lack the 1e-6 of the load terms), so the library flattens them itself in its 2-D mode
(pfem_boundary::mode2d), plane by plane with the same d/2 * 1e-6 scale: heatflux_boundary / convection_boundary /
radiation_boundary of Static2D / StaticCyl take 2-D node lists; boundary_verbatim keeps the reference's convection matrix as
written (no 1e-6, a second factor r in the cylindrical solver), False applies the unit factor and the single r."""
from dataclasses import dataclass, field

import numpy as np

from . import _lib as L
from .configs import Problem
from .solvers import Dynamic3D, Shockley3D, Static3D, ThermoElectric3D


@dataclass
class Problem2D:
    """flat arrays of a 2-D problem: node (i0, i1) -> i0 * n1 + i1, element (i0, i1) -> i0 * (n1 - 1) + i1 (x = axis 0 slowest)"""
    name: str
    kind: str                      # 'thermal' | 'shockley'
    x: np.ndarray                  # tran (Cartesian) or r (cylindrical), um
    y: np.ndarray                  # vert, um
    elem_mat: np.ndarray
    T0: float
    dT: float
    tab_lat: np.ndarray
    tab_vert: np.ndarray
    bc_nodes: np.ndarray
    bc_values: np.ndarray
    heat: np.ndarray = None
    cyl: bool = False
    length: float = 1000.          # Cartesian: extrusion length of the geometry, um (geometry->getExtrusion()->getLength())
    inittemp: float = 300.
    maxerr: float = 0.05
    elem_junc: np.ndarray = None
    elem_role: np.ndarray = None
    noheat: np.ndarray = None
    beta: float = 11.
    js: float = 1.
    pcond: float = 5.
    ncond: float = 50.
    start_cond: tuple = (0., 5.)
    tab_cprho: np.ndarray = None   # [nmat][nT] cp(T) * dens(T), J/(m^3 K) (Dynamic2D / DynamicCyl)
    meta: dict = field(default_factory=dict)

    @property
    def n(self):
        return (len(self.x), len(self.y))

    @property
    def N(self):
        return len(self.x) * len(self.y)

    @property
    def E(self):
        return (len(self.x) - 1) * (len(self.y) - 1)


def embed(p2, thickness=1.):
    """Problem2D -> the brick Problem of one element layer: axes (long = [0, thickness], x, y), order 012 (the dummy axis major, the
    vertical axis minor).  Plane 0 of the brick mesh is numbered exactly like the 2-D mesh, plane 1 follows at offset N."""
    N2 = p2.N
    nodes = np.asarray(p2.bc_nodes, dtype=np.int64)
    p = Problem(p2.name, p2.kind, [np.array([0., float(thickness)]), np.asarray(p2.x, dtype=np.float64), np.asarray(p2.y, dtype=np.float64)],
                "012", np.asarray(p2.elem_mat, dtype=np.uint32), p2.T0, p2.dT, p2.tab_lat, p2.tab_vert,
                np.concatenate([nodes, nodes + N2]).astype(np.uintp), np.concatenate([p2.bc_values, p2.bc_values]).astype(np.float64),
                heat=None if p2.heat is None else np.asarray(p2.heat, dtype=np.float64), inittemp=p2.inittemp, maxerr=p2.maxerr)
    p.elem_junc = None if p2.elem_junc is None else np.asarray(p2.elem_junc, dtype=np.uint32)
    p.elem_role = None if p2.elem_role is None else np.asarray(p2.elem_role, dtype=np.uint8)
    p.noheat = None if p2.noheat is None else np.asarray(p2.noheat, dtype=np.uint8)
    p.tab_cprho = p2.tab_cprho
    for k in ("beta", "js", "pcond", "ncond", "start_cond"):
        setattr(p, k, getattr(p2, k))
    p.meta = dict(p2.meta)
    return p


def interpolate2d(p2, field, xq, yq):
    """A nodal field of the 2-D mesh on the product mesh xq x yq (numbering i0 * len(yq) + i1), linear like
    interpolate(mesh, data, dst_mesh, INTERPOLATION_LINEAR) of the 2-D providers (getTemperatures, therm2d.cpp:529-536;
    getVoltage, electr2d.cpp:507-514): points outside the mesh take the value of the nearest edge.  Host side: the providers of the
    2-D solvers hand out a few thousand values, the field is already downloaded."""
    x, y = np.asarray(p2.x, dtype=np.float64), np.asarray(p2.y, dtype=np.float64)
    v = np.asarray(field, dtype=np.float64).reshape(len(x), len(y))

    def weights(a, q):
        q = np.clip(np.asarray(q, dtype=np.float64), a[0], a[-1])
        hi = np.clip(np.searchsorted(a, q, side="right"), 1, len(a) - 1)
        return hi - 1, hi, (q - a[hi - 1]) / (a[hi] - a[hi - 1])
    i0, i1, fx = weights(x, xq)
    j0, j1, fy = weights(y, yq)
    fx, fy = fx[:, None], fy[None, :]
    return ((1. - fx) * ((1. - fy) * v[i0][:, j0] + fy * v[i0][:, j1]) + fx * ((1. - fy) * v[i1][:, j0] + fy * v[i1][:, j1])).ravel()


class _Embedded2D:
    """shared plumbing of the four solvers: `problem` takes a Problem2D, providers return 2-D arrays"""
    cyl = False
    _p2 = None

    @property
    def problem(self):
        return self._p2

    @problem.setter
    def problem(self, p2):
        assert isinstance(p2, Problem2D)
        if bool(p2.cyl) != self.cyl:
            raise L.BadInput(f"{self.id}: the problem is {'cylindrical' if p2.cyl else 'Cartesian'}")
        if p2.cyl and not (p2.x[0] >= 0.):
            raise L.BadInput(f"{self.id}: negative radius in the mesh")
        self._p2 = p2
        self._problem = embed(p2)
        self.invalidate()

    def _new_fem(self):
        self.layout = "abi"          # order 012 already has the vertical axis fastest and the dummy axis slowest
        return super()._new_fem()

    def _weights(self):
        if self.cyl:
            x = np.asarray(self._p2.x, dtype=np.float64)
            self._fem.set_axis_weight(1, 0.5 * (x[1:] + x[:-1]))      # midpoint.rad_r()

    def initialize(self):
        super().initialize()
        self._weights()

    def _plane(self, v):
        return np.asarray(v)[:self._p2.N].copy()

    def _elem_area_weight(self):
        """w * h of every element (times r in the cylindrical case)"""
        p2 = self._p2
        w, h = np.diff(p2.x)[:, None], np.diff(p2.y)[None, :]
        a = w * h
        if self.cyl:
            a = a * (0.5 * (p2.x[1:] + p2.x[:-1]))[:, None]
        return a.ravel()


class Static2D(_Embedded2D, Static3D):
    """thermal.static.Static2D with algorithm='cuda' (therm2d.cpp, Geometry2DCartesian)"""

    def _set_boundary(self, f):
        pass                          # after the element weights, in the library's 2-D mode: initialize below

    def initialize(self):
        super().initialize()          # mesh, materials, Dirichlet, then the radial weights
        self._fem.set_boundary(self.heatflux_boundary, self.convection_boundary, self.radiation_boundary, self.boundary_verbatim,
                               mode2d=2 if self.cyl else 1)

    def outTemperature(self, mesh=None):
        """temperatures on the solver's own mesh, or (mesh = (xq, yq)) interpolated linearly onto a foreign rectangular 2-D mesh"""
        if not self.initialized:
            return np.full(self._p2.N if mesh is None else len(mesh[0]) * len(mesh[1]), float(self.inittemp))
        T = self._plane(Static3D.outTemperature(self))
        return T if mesh is None else interpolate2d(self._p2, T, mesh[0], mesh[1])

    def outHeatFlux(self):
        """(E, 2): (-k_x dT/dx, -k_y dT/dy) at the element midpoints, W/m^2 (saveHeatFluxes, therm2d.cpp:494-527)"""
        self._fem.update_conductivity_thermal()
        return Static3D.outHeatFlux(self)[:, 1:3].copy()

    def outThermalConductivity(self):
        return Static3D.outThermalConductivity(self)


class StaticCyl(Static2D):
    """thermal.static.StaticCyl (therm2d.cpp, Geometry2DCylindrical): x is the radius"""
    cyl = True


class Dynamic2D(_Embedded2D, Dynamic3D):
    """thermal.dynamic.Dynamic2D with algorithm='cuda' (femT2d.cpp, Geometry2DCartesian): the element capacity
    cp dens 0.25e-12 w h / timestep / 1e-9 (:164), lumped or consistent (4/9, 2/9, 1/9: the rows of M (x) M (x) M_z of the brick sum
    to exactly that over the dummy axis), is the brick capacity of the layer times the same d/2 * 1e-6 as K and F; the time loop
    (:415-445) is the one of femT3d.cpp — pfem_solve_dynamic, corrected theta scheme."""

    def outTemperature(self, mesh=None):
        if not self.initialized:
            return np.full(self._p2.N if mesh is None else len(mesh[0]) * len(mesh[1]), float(self.inittemp))
        T = self._plane(Dynamic3D.outTemperature(self))
        return T if mesh is None else interpolate2d(self._p2, T, mesh[0], mesh[1])

    def outHeatFlux(self):
        return Dynamic3D.outHeatFlux(self)[:, 1:3].copy()


class DynamicCyl(Dynamic2D):
    """thermal.dynamic.DynamicCyl (femT2d.cpp:258-388): kx, ky, c and f carry the midpoint radius"""
    cyl = True


class Shockley2D(_Embedded2D, Shockley3D):
    """electrical.shockley.Shockley2D with algorithm='cuda' (electr2d.cpp + beta.hpp, Geometry2DCartesian)"""

    def outVoltage(self, mesh=None):
        """potentials on the solver's own mesh, or (mesh = (xq, yq)) interpolated linearly onto a foreign rectangular 2-D mesh"""
        V = self._plane(Shockley3D.outVoltage(self))
        return V if mesh is None else interpolate2d(self._p2, V, mesh[0], mesh[1])

    def outCurrentDensity(self):
        """(E, 2) kA/cm^2 (electr2d.cpp:399-405)"""
        return Shockley3D.outCurrentDensity(self)[:, 1:3].copy()

    @property
    def maxcur2d(self):
        return tuple(self.maxcur[1:3])

    def integrate_current(self, vindex, onlyactive=False):
        """integrateCurrent (electr2d.cpp:466-496), mA"""
        p2 = self._p2
        cur = self.outCurrentDensity()[:, 1].reshape(p2.n[0] - 1, p2.n[1] - 1)[:, vindex]
        if onlyactive:
            cur = np.where(np.asarray(p2.elem_junc).reshape(p2.n[0] - 1, p2.n[1] - 1)[:, vindex] > 0, cur, 0.)
        x = np.asarray(p2.x, dtype=np.float64)
        if self.cyl:
            return float((cur * (x[1:] ** 2 - x[:-1] ** 2)).sum() * np.pi * 0.01)
        return float((cur * np.diff(x)).sum() * p2.length * 0.01)

    def get_total_heat(self):
        """getTotalHeat (electr2d.cpp:627-649), mW"""
        W = float((self._elem_area_weight() * self.outHeat()).sum())
        return 2e-15 * np.pi * W if self.cyl else self._p2.length * 1e-15 * W

    def get_total_energy(self, eps=None):
        """getTotalEnergy (electr2d.cpp:574-615), J"""
        p2 = self._p2
        eps = np.asarray(p2.meta["eps"] if eps is None else eps, dtype=np.float64)
        V = self.outVoltage().reshape(p2.n)
        ll, lr, ul, ur = V[:-1, :-1], V[1:, :-1], V[:-1, 1:], V[1:, 1:]
        dvx = 0.5e6 * (-ll + lr - ul + ur) / np.diff(p2.x)[:, None]
        dvy = 0.5e6 * (-ll - lr + ul + ur) / np.diff(p2.y)[None, :]
        W = float((self._elem_area_weight() * eps * (dvx * dvx + dvy * dvy).ravel()).sum())
        eps0 = 8.854187817e-12
        return (2. * np.pi if self.cyl else p2.length) * 0.5e-18 * eps0 * W

    def get_capacitance(self):
        vals = np.unique(np.asarray(self._p2.bc_values))
        if len(vals) != 2:
            raise L.BadInput(f"{self.id}: cannot estimate applied voltage (exactly 2 voltage boundary conditions required)")
        U = float(vals[1] - vals[0])
        return 2e12 * self.get_total_energy() / (U * U)


class ShockleyCyl(Shockley2D):
    """electrical.shockley.ShockleyCyl (electr2d.cpp, Geometry2DCylindrical): x is the radius"""
    cyl = True


class ThermoElectric2D(ThermoElectric3D):
    """meta.shockley.ThermoElectric2D (solvers/meta/shockley/thermoelectric.py:187-211 with Static2D + Shockley2D): the loop and the
    device-resident field exchange of ThermoElectric3D — the embedded meshes share the dummy axis, so pfem_transfer_temperature /
    pfem_transfer_heat interpolate bilinearly in (x, y) exactly like getTemperatures / getHeatDensity on 2-D meshes."""
    thermal_solver, electrical_solver = Static2D, Shockley2D


class ThermoElectricCyl(ThermoElectric3D):
    """meta.shockley.ThermoElectricCyl: StaticCyl + ShockleyCyl"""
    thermal_solver, electrical_solver = StaticCyl, ShockleyCyl
