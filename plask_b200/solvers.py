"""Host-side mirror of the reference's solver interface for the CUDA algorithm.

`Static3D` and `Shockley3D` keep the names, defaults and semantics of
thermal.static.Static3D (solvers/thermal/static/therm3d.{hpp,cpp}, python/therm_python.cpp)
and electrical.shockley.Shockley3D (solvers/electrical/shockley/electr3d.{hpp,cpp}, beta.hpp,
python/electr_python.cpp): `compute(loops)`, `maxerr`, `inittemp`, `algorithm`, `iterative.*`,
`beta`/`js`, `pcond`/`ncond`, `convergence`, `inHeat`, `inTemperature`, `outTemperature`,
`outHeatFlux`, `outVoltage`, `outCurrentDensity`, `outHeat`, `get_total_current()` ...
What differs is only where the geometry comes from: PLaSK's geometry tree is out of scope
(SURVEY.md §2.2), so a solver is bound to a `configs.Problem`, the flat arrays the plugin
extracts from geometry/mesh/materials.

`algorithm` must be 'cuda' — the new FemMatrixAlgorithm value; there is no CPU path here.
"""
import numpy as np

from . import _lib as L
from .configs import Problem, setup_active
from .fem import DeviceFem


class IterativeParams:
    """iter_params of FemSolverWithMesh (plask/common/fem/iterative_matrix.hpp:25-94), the subset
    that has a meaning for the CUDA algorithm."""

    def __init__(self):
        self.maxit = 10000          # maxit
        self.maxerr = 1e-8          # here: relative residual ||r||/||b_free|| (north-star tolerance)
        self.preconditioner = "jac"
        self.accelerator = "cg"
        self.noconv = "warning"     # 'error' | 'warning' | 'continue'  (:88-89)
        # outputs (:90-93)
        self.converged = True
        self.iters = 0
        self.err = 0.


class _FemSolver:
    def __init__(self, name=""):
        self.id = name
        self.algorithm = "cuda"
        self.iterative = IterativeParams()
        self.device = 0
        # FemSolverWithMaskedMesh::empty_elements (fem_solver.hpp:152-189): 'include' keeps the full mesh (what the
        # reference does for its iterative algorithm by default), 'exclude' drops the elements of EMPTY material
        # (problem.empty) like the reference's Cholesky default does
        self.empty_elements = "include"
        # device layout: 'abi' keeps the mesh's iteration order, 'vertical-minor' stores the vertical axis fastest (contiguous
        # lines for the 'ljac' preconditioner), 'auto' = 'vertical-minor' when the preconditioner is 'ljac'
        self.layout = "auto"
        self.variant = 3            # 3 production (fused single-kernel iteration); 0/2 two-kernel; 1 simple reference kernels
        self._fem = None
        self._problem = None
        self.initialized = False
        self.log = []
        self.stats = None
        # slab mode (multi-GPU): dict(rank, nranks, own_lo, own_hi, allgather=callable(bytes) -> list of bytes);
        # `problem` is then the LOCAL problem (configs.slab_problem) and compute() is collective
        self.slab = None

    # geometry + mesh + materials, flattened
    @property
    def problem(self):
        return self._problem

    @problem.setter
    def problem(self, p):
        assert isinstance(p, Problem)
        self._problem = p
        self.invalidate()

    def invalidate(self):
        """Solver::invalidate -> onInvalidate (therm3d.cpp:118-122, electr3d.cpp:195-201): drops the
        fields and the device state."""
        if self._fem is not None:
            self._fem.close()
        self._fem = None
        self.initialized = False

    def _elem_materials(self):
        """material id per element; elements outside the masked mesh are marked PFEM_MAT_EXCLUDED"""
        p = self._problem
        if self.empty_elements not in ("include", "exclude", "default"):
            raise L.BadInput(f"{self.id}: empty_elements must be 'include', 'exclude' or 'default'")
        if self.empty_elements != "exclude" or p.empty is None or not np.any(p.empty):
            return p.elem_mat
        m = np.array(p.elem_mat, dtype=np.uint32, copy=True)
        m[np.asarray(p.empty) != 0] = L.MAT_EXCLUDED
        return m

    def _dirichlet(self):
        """boundary conditions live on the masked mesh: a place holds no nodes outside it"""
        p = self._problem
        nodes, vals = np.asarray(p.bc_nodes), np.asarray(p.bc_values)
        if self.empty_elements == "exclude" and p.empty is not None and np.any(p.empty):
            keep = self.masked_nodes()[nodes.astype(np.int64)]
            nodes, vals = nodes[keep], vals[keep]
        return nodes, vals

    def masked_nodes(self):
        """bool [N]: nodes of the masked mesh (RectangularMaskedMesh3D keeps the nodes of the kept elements)"""
        p = self._problem
        keep = np.ones(p.E, dtype=bool) if (self.empty_elements != "exclude" or p.empty is None) else (np.asarray(p.empty) == 0)
        n = p.n
        k3 = keep[np.broadcast_to(p.elem_index_grid(), tuple(k - 1 for k in n))]
        act = np.zeros(n, dtype=bool)
        for d0 in (0, 1):
            for d1 in (0, 1):
                for d2 in (0, 1):
                    act[d0:n[0] - 1 + d0, d1:n[1] - 1 + d1, d2:n[2] - 1 + d2] |= k3
        out = np.zeros(p.N, dtype=bool)
        out[np.broadcast_to(p.node_index_grid(), n).ravel()] = act.ravel()
        return out

    def _new_fem(self):
        f = DeviceFem(self.device)
        lay = self.layout
        if lay == "auto":
            # the warp-per-row line kernel holds up to 512 nodes per line; longer vertical axes keep the mesh order (strided kernel)
            lay = "vertical-minor" if (self.iterative.preconditioner in ("ljac", "mlj") and self._problem.n[2] <= 512) else "abi"
            if self.slab is not None and self.slab["nranks"] > 1 and self._problem.strides[2] == max(self._problem.strides):
                lay = "abi"     # a vertical major axis cannot be cut into slabs in the vertical-minor layout
        if lay not in ("abi", "vertical-minor"):
            raise L.BadInput(f"{self.id}: layout must be 'auto', 'abi' or 'vertical-minor'")
        f.set_layout(L.LAYOUT_VERTICAL_MINOR if lay == "vertical-minor" else L.LAYOUT_ABI)
        return f

    def _setup_slab(self, f):
        """slab mode: tell the context which planes it owns and map the neighbours' memory (collective)"""
        if self.slab is not None:
            sl = self.slab
            f.slab_configure(sl["rank"], sl["nranks"], sl["own_lo"], sl["own_hi"])
            if sl["nranks"] > 1:
                f.slab_connect(sl["allgather"](f.slab_export()))

    def _opts(self, loops, outer_tol):
        if self.algorithm != "cuda":
            raise L.BadInput(f"{self.id}: algorithm '{self.algorithm}' is not provided by plask_b200 "
                             "(cholesky/gauss/iterative live in the reference); use 'cuda'")
        # NSPCG names (iterative_matrix.hpp:27-46): 'jac' point Jacobi, 'ljac' line Jacobi (here: lines along the vertical axis);
        # 'mlj' (new): additive multilevel line preconditioner, the GPU counterpart of the reference's default 'ic' strength
        pre = {"jac": 0, "ljac": 1, "mlj": 2}.get(self.iterative.preconditioner)
        if pre is None or self.iterative.accelerator != "cg":
            raise L.BadInput(f"{self.id}: the CUDA algorithm implements accelerator 'cg' with preconditioner 'jac', 'ljac' or 'mlj'")
        return dict(maxit=int(self.iterative.maxit), lin_tol=float(self.iterative.maxerr), precond=pre,
                    outer_tol=float(outer_tol), loops=int(loops), variant=int(self.variant))

    def _after(self, rc, st):
        self.stats = st
        it = self.iterative
        it.converged, it.iters, it.err = st["converged"], st["last_iters"], st["lin_relres"]
        if rc == L.PFEM_NOT_CONVERGED:
            msg = f"Failed to converge in {it.maxit} iterations (error {it.err})"
            if it.noconv == "error":
                raise L.ComputationError(f"{self.id}: {msg}")
            self.log.append(("warning" if it.noconv == "warning" else "detail", msg))


class Static3D(_FemSolver):
    """thermal.static.Static3D with algorithm='cuda'."""

    def __init__(self, name=""):
        super().__init__(name)
        self.inittemp = 300.     # therm3d.cpp:23
        self.maxerr = 0.05       # therm3d.cpp:24
        self.inHeat = None       # None -> problem.heat; scalar or [E] array (W/m3)
        # boundary conditions of the 2nd / 3rd kind and radiation (therm3d.hpp:79-82), lists in definition order:
        # (nodes, q [W/m2]) / (nodes, coeff [W/m2/K], ambient [K]) / (nodes, emissivity, ambient [K])
        self.heatflux_boundary = []
        self.convection_boundary = []
        self.radiation_boundary = []
        self.boundary_verbatim = True   # reproduce setBoundaries' local-slot accumulation (therm3d.cpp:157-162)
        self.maxT = 0.
        self.loopno = 0

    def initialize(self):
        p = self._problem
        if p is None:
            raise L.BadInput(f"{self.id}: no geometry/mesh (problem) specified")
        f = self._fem = self._new_fem()
        f.set_mesh(p.axes, p.strides)
        self._setup_slab(f)
        f.set_materials(self._elem_materials(), p.T0, p.dT, p.tab_lat, p.tab_vert)
        f.set_field(float(self.inittemp))               # temperatures.reset(size, inittemp), :79
        f.set_dirichlet(*self._dirichlet())
        self._set_boundary(f)
        self.loopno = 0
        self.initialized = True

    def _set_boundary(self, f):
        f.set_boundary(self.heatflux_boundary, self.convection_boundary, self.radiation_boundary, self.boundary_verbatim)

    def compute(self, loops=0):
        """ThermalFem3DSolver::compute (therm3d.cpp:281-340); returns the max loop error."""
        if not self.initialized:
            self.initialize()
        f, p = self._fem, self._problem
        heat = p.heat if self.inHeat is None else self.inHeat
        if isinstance(heat, Shockley3D):
            # inHeat connected to electrical.outHeat: device-to-device, interpolated like getHeatDensity (electr3d.cpp:538-548)
            f.take_heat_from(heat._fem)
        else:
            if heat is not None and np.isscalar(heat):
                heat = np.full(p.E, float(heat))
            f.set_source(heat)
        rc, st = f.solve_thermal(**self._opts(loops, self.maxerr))
        self._after(rc, st)
        self.maxT, self.loopno = st["maxval"], st["loopno"]
        self.log.append(("result", f"Loop {st['outer_loops']}({self.loopno}): max(T) = {self.maxT:.3f} K, "
                                   f"error = {st['err']:g} K"))
        return st["toterr"]

    # providers, on the solver's own mesh (interpolation onto other meshes is a "next" row)
    def outTemperature(self, mesh=None):
        """temperatures on the solver's own mesh, or (mesh = (axes, order)) interpolated linearly onto a foreign rectilinear
        mesh like getTemperatures(dst_mesh, INTERPOLATION_LINEAR) (therm3d.cpp:387-395)"""
        if not self.initialized:
            n = self._problem.N if mesh is None else int(np.prod([len(a) for a in mesh[0]]))
            return np.full(n, float(self.inittemp))
        if mesh is None:
            return self._fem.get_field()
        from .configs import strides_for
        axes, order = mesh
        return self._fem.interpolate_field(axes, strides_for(tuple(len(a) for a in axes), order)[0])

    def outHeatFlux(self):
        return self._fem.get_elem(L.ELEM_FLUX)

    def outThermalConductivity(self):
        self._fem.update_conductivity_thermal()
        return self._fem.get_elem(L.ELEM_COND)


class Dynamic3D(_FemSolver):
    """thermal.dynamic.Dynamic3D with algorithm='cuda' (solvers/thermal/dynamic/femT3d.{hpp,cpp}, python/dynamic_python.cpp:70-86):
    `compute(time)`, `inittemp`, `timestep`, `methodparam`, `lumping`, `rebuildfreq`, `logfreq`, `time` / `elapsed_time`,
    `inHeat`, `outTemperature`, `outHeatFlux`, `outThermalConductivity`.  The update is the corrected theta scheme
    (include/plaskfem_cuda.h, pfem_solve_dynamic)."""

    def __init__(self, name=""):
        super().__init__(name)
        self.inittemp = 300.        # femT3d.cpp:27
        self.methodparam = 0.5      # :28
        self.timestep = 0.1         # ns, :29
        self.lumping = True         # :31
        self.rebuildfreq = 0        # :32
        self.logfreq = 500          # :33
        self.inHeat = None
        self.maxT = 0.
        self._elapstime = 0.
        self.physical_time = 0.

    @property
    def time(self):
        return self._elapstime

    elapsed_time = time

    def initialize(self):
        p = self._problem
        if p is None:
            raise L.BadInput(f"{self.id}: no geometry/mesh (problem) specified")
        f = self._fem = self._new_fem()
        f.set_mesh(p.axes, p.strides)
        self._setup_slab(f)
        f.set_materials(self._elem_materials(), p.T0, p.dT, p.tab_lat, p.tab_vert)
        tab = p.tab_cprho
        if tab is None:
            from .configs import capacity_tables
            tab = capacity_tables(p.T0, p.dT, p.tab_lat.shape[1])[:p.tab_lat.shape[0]]
        f.set_capacity(tab)
        f.set_field(float(self.inittemp))                # temperatures.reset(size, inittemp), :82
        f.set_dirichlet(*self._dirichlet())
        self._elapstime = 0.                              # :76
        self.physical_time = 0.
        self.initialized = True

    def compute(self, time):
        """DynamicThermalFem3DSolver::compute(time) (femT3d.cpp:258-305): advances the temperatures by `time` ns."""
        if not self.initialized:
            self.initialize()
        f, p = self._fem, self._problem
        heat = p.heat if self.inHeat is None else self.inHeat
        if heat is not None and np.isscalar(heat):
            heat = np.full(p.E, float(heat))
        f.set_source(heat)
        if not self.lumping and self.variant == 3:
            raise L.BadInput(f"{self.id}: the consistent capacity matrix (lumping = no) runs with variant = 1")
        o = self._opts(0, 0.)
        rc, st = f.solve_dynamic(time, self.timestep, self.methodparam, self.lumping, self.rebuildfreq, log=bool(self.logfreq), **o)
        self._after(rc, st)
        self.maxT = st["maxval"]
        steps = st["outer_loops"]
        if self.logfreq:                                  # "Time {:.2f} ns: max(T) = {:.3f} K" every logfreq steps (:287-291)
            l = self.logfreq
            for i in range(steps):
                if l == 0:
                    self.log.append(("result", f"Time {self._elapstime + i * self.timestep:.2f} ns: max(T) = {st['maxT_log'][i]:.3f} K"))
                    l = self.logfreq
                l -= 1
        self._elapstime += steps * self.timestep - (self.timestep if steps else 0.)   # `elapstime -= timestep` after the loop (:297)
        self.physical_time += steps * self.timestep   # one timestep per solve: the loop runs time/timestep + 1 of them (:271-272)
        return 0.

    def outTemperature(self, mesh=None):
        if not self.initialized:
            n = self._problem.N if mesh is None else int(np.prod([len(a) for a in mesh[0]]))
            return np.full(n, float(self.inittemp))
        if mesh is None:
            return self._fem.get_field()
        from .configs import strides_for
        axes, order = mesh
        return self._fem.interpolate_field(axes, strides_for(tuple(len(a) for a in axes), order)[0])

    def outHeatFlux(self):
        self._fem.update_conductivity_thermal()           # saveHeatFluxes re-evaluates thermk at the current temperatures (:331-340)
        return self._fem.get_elem(L.ELEM_FLUX)

    def outThermalConductivity(self):
        self._fem.update_conductivity_thermal()
        return self._fem.get_elem(L.ELEM_COND)


class Shockley3D(_FemSolver):
    """electrical.shockley.Shockley3D with algorithm='cuda'."""

    def __init__(self, name=""):
        super().__init__(name)
        self.maxerr = 0.05           # electr3d.cpp:26
        self.pcond, self.ncond = 5., 50.   # :22-23
        self.start_cond = (0., 5.)   # default_junction_conductivity, :25
        self.convergence = "fast"    # :31
        self.beta = None             # float | callable(T) | list per junction  (beta.hpp, electr_python.cpp)
        self.js = None
        self.inTemperature = 300.    # scalar or [E] array
        self.loopno = 0
        self.maxcur = (0., 0., 0.)
        self._acts, self._ncol = [], 0
        self._junc_cond = None

    def initialize(self):
        p = self._problem
        if p is None:
            raise L.BadInput(f"{self.id}: no geometry/mesh (problem) specified")
        if self.beta is None:
            self.beta = p.beta
        if self.js is None:
            self.js = p.js
        f = self._fem = self._new_fem()
        f.set_mesh(p.axes, p.strides)
        self._setup_slab(f)
        f.set_materials(self._elem_materials(), p.T0, p.dT, p.tab_lat, p.tab_vert)
        f.set_field(0.)                                  # potential.reset(size, 0.), electr3d.cpp:190
        f.set_dirichlet(*self._dirichlet())
        f.set_source(None)
        f.set_noheat(p.noheat)
        self._acts, self._ncol = setup_active(p)         # setupActiveRegions, :89-183
        if self._junc_cond is None or len(self._junc_cond) != max(self._ncol, 1):
            self._junc_cond = np.tile(np.asarray(self.start_cond, dtype=np.float64), (max(self._ncol, 1), 1))
        self.loopno = 0
        self.initialized = True

    def invalidate(self):
        super().invalidate()
        self._junc_cond = None      # junction_conductivity.reset(1, default), :200

    def _midplane_elements(self):
        """element index (element order of the mesh) of the mid-plane element of every junction-table entry, -1 = unused entry"""
        es = self._problem.estrides
        out = np.full(max(self._ncol, 1), -1, dtype=np.int64)
        for a in self._acts:
            v = (a["top"] + a["bottom"]) // 2
            t = np.arange(a["left"], a["right"])[:, None]
            l = np.arange(a["back"], a["front"])[None, :]
            out[(a["offset"] + a["ld"] * t + l).ravel()] = (l * es[0] + t * es[1] + v * es[2]).ravel()
        return out

    def _per_junction(self, v, k):
        return v[k] if isinstance(v, (list, tuple)) else v

    def _junction_params(self, Te):
        """beta(T), js(T) per junction-table entry at the mid-plane element temperature."""
        p = self._problem
        es = p.estrides
        n = max(self._ncol, 1)
        bcol, jcol = np.ones(n), np.ones(n)
        for k, a in enumerate(self._acts):
            b, j = self._per_junction(self.beta, k), self._per_junction(self.js, k)
            if b is None:
                raise L.BadInput(f"{self.id}: no beta given for junction {k}")
            v = (a["top"] + a["bottom"]) // 2
            t = np.arange(a["left"], a["right"])[:, None]
            l = np.arange(a["back"], a["front"])[None, :]
            col = (a["offset"] + a["ld"] * t + l).ravel()
            e = (l * es[0] + t * es[1] + v * es[2]).ravel()
            if isinstance(Te, dict):        # temperatures fetched from the device: one value per junction-table entry
                T = Te["per_column"][col]
            else:
                T = np.full(e.size, float(Te)) if np.isscalar(Te) else np.asarray(Te)[e]
            bcol[col] = np.vectorize(b)(T) if callable(b) else b
            jcol[col] = np.vectorize(j)(T) if callable(j) else j
        return bcol, jcol

    def compute(self, loops=0):
        """ElectricalFem3DSolver::compute (electr3d.cpp:356-442)."""
        if not self.initialized:
            self.initialize()
        f, p = self._fem, self._problem
        Te = self.inTemperature
        if isinstance(Te, Static3D):
            # inTemperature connected to thermal.outTemperature: device-to-device, interpolated like getTemperatures
            # (therm3d.cpp:385-393); beta(T)/js(T) callables need the mid-plane values on the host
            if not Te.initialized:
                Te.initialize()
            f.take_temperature_from(Te._fem)
            need_T = any(callable(self._per_junction(v, k)) for k in range(len(self._acts)) for v in (self.beta, self.js))
            if need_T:
                # only the mid-plane element of every junction column is read back (electr3d.cpp:261-262: temperature[tidx])
                elems = self._midplane_elements()
                Tcol = np.full(elems.size, 300.)
                used = elems >= 0
                Tcol[used] = f.get_elem_temperature(elems[used])
                Te = {"per_column": Tcol}
            else:
                Te = 300.
        else:
            f.set_elem_temperature(Te if np.isscalar(Te) else np.asarray(Te, dtype=np.float64))
        bcol, jcol = self._junction_params(Te)
        f.set_junctions(self._acts, p.elem_junc if p.elem_junc is not None else np.zeros(p.E, np.uint32),
                        p.elem_role, self.pcond, self.ncond, self._junc_cond, bcol, jcol,
                        stable=(self.convergence == "stable"))
        rc, st = f.solve_shockley(**self._opts(loops, self.maxerr))
        self._after(rc, st)
        self.loopno, self.maxcur = st["loopno"], st["maxcur"]
        if self._ncol:
            self._junc_cond = f.get_junction_cond()      # saveConductivity, :227-237
        self.log.append(("result", f"Loop {st['outer_loops']}({self.loopno}): max(j@junc) = {st['maxval']:g} kA/cm2, "
                                   f"error = {st['err']:g}%"))
        return st["toterr"]

    # providers on the solver's own meshes
    def outVoltage(self, mesh=None):
        """potential on the solver's own mesh, or interpolated linearly onto a foreign rectilinear mesh (axes, order)"""
        if mesh is None:
            return self._fem.get_field()
        from .configs import strides_for
        axes, order = mesh
        return self._fem.interpolate_field(axes, strides_for(tuple(len(a) for a in axes), order)[0])

    def outCurrentDensity(self):
        return self._fem.get_elem(L.ELEM_CURRENT)

    def outHeat(self):
        return self._fem.get_elem(L.ELEM_HEAT, self._problem.noheat)

    def outConductivity(self):
        if not self.initialized:
            self.initialize()
            f = self._fem
            f.set_elem_temperature(self.inTemperature)
            bcol, jcol = self._junction_params(self.inTemperature)
            f.set_junctions(self._acts, self._problem.elem_junc if self._problem.elem_junc is not None
                            else np.zeros(self._problem.E, np.uint32), self._problem.elem_role, self.pcond, self.ncond,
                            self._junc_cond, bcol, jcol, stable=(self.convergence == "stable"))
        self._fem.update_conductivity_shockley()
        return self._fem.get_elem(L.ELEM_COND)

    def _elem_sizes(self):
        p = self._problem
        d = [np.diff(a) for a in p.axes]
        return d

    def integrate_current(self, vindex, onlyactive=False):
        """integrateCurrent (electr3d.cpp:480-497), mA."""
        p = self._problem
        cur = self.outCurrentDensity()
        es = p.estrides
        d = self._elem_sizes()
        i0 = np.arange(p.n[0] - 1)[:, None]
        i1 = np.arange(p.n[1] - 1)[None, :]
        e = i0 * es[0] + i1 * es[1] + vindex * es[2]
        w = d[0][:, None] * d[1][None, :]
        jz = cur[e, 2]
        if onlyactive:
            jz = np.where(p.elem_junc[e] > 0, jz, 0.)
        return float((jz * w).sum() * 0.01)

    def get_total_current(self, nact=0):
        """getTotalCurrent (electr3d.cpp:499-505)."""
        if nact >= len(self._acts):
            raise L.BadInput(f"{self.id}: wrong active region number")
        a = self._acts[nact]
        return self.integrate_current((a["bottom"] + a["top"]) // 2, True)

    def get_total_heat(self):
        """getTotalHeat (electr3d.cpp:612-624), mW."""
        p = self._problem
        heat = self.outHeat()
        d = self._elem_sizes()
        vol = d[0][:, None, None] * d[1][None, :, None] * d[2][None, None, :]
        return float((1e-15 * vol * heat[p.elem_index_grid()]).sum())


    def get_total_energy(self, eps=None):
        """getTotalEnergy (electr3d.cpp:568-600), J: 0.5 eps0 eps |grad V|^2 over the elements of the (masked) mesh, the
        gradient at the element midpoint.  eps: relative permittivity per element (material->eps(T)); default problem.meta['eps']."""
        p = self._problem
        eps = np.asarray(p.meta["eps"] if eps is None else eps, dtype=np.float64)
        V = self.outVoltage()[np.broadcast_to(p.node_index_grid(), p.n)]
        lo, up = slice(None, -1), slice(1, None)
        c = {(a, b, c_): V[a, b, c_] for a in (lo, up) for b in (lo, up) for c_ in (lo, up)}
        d = self._elem_sizes()
        d0, d1, d2 = d[0][:, None, None], d[1][None, :, None], d[2][None, None, :]
        s = lambda f: sum(f(k) * v for k, v in c.items())
        dvx = -0.25e6 * s(lambda k: 1. if k[0] is up else -1.) / d0
        dvy = -0.25e6 * s(lambda k: 1. if k[1] is up else -1.) / d1
        dvz = -0.25e6 * s(lambda k: 1. if k[2] is up else -1.) / d2
        e3 = eps[p.elem_index_grid()]
        if self.empty_elements == "exclude" and p.empty is not None:
            e3 = np.where(np.asarray(p.empty)[p.elem_index_grid()] != 0, 0., e3)   # loop over maskedMesh->elements()
        epsilon0 = 1. / (4e-7 * np.pi) / 299792458. ** 2
        return float((0.5e-18 * epsilon0 * d0 * d1 * d2 * e3 * (dvx * dvx + dvy * dvy + dvz * dvz)).sum())

    def get_capacitance(self):
        """getCapacitance (electr3d.cpp:602-610), pF: exactly two voltage boundary conditions are required."""
        vals = np.unique(np.asarray(self._problem.bc_values))
        if len(vals) != 2:
            raise L.BadInput(f"{self.id}: cannot estimate applied voltage (exactly 2 voltage boundary conditions required)")
        U = float(vals[1] - vals[0])
        return 2e12 * self.get_total_energy() / (U * U)


class ThermoElectric3D:
    """meta.shockley.ThermoElectric3D (solvers/meta/shockley/thermoelectric.py:160-215) over the two CUDA solvers.

    `thermal.inHeat <- electrical.outHeat` and `electrical.inTemperature <- thermal.outTemperature` are connected
    on the device (pfem_transfer_heat / pfem_transfer_temperature); the loop is the reference's:

        while terr > thermal.maxerr or verr > electrical.maxerr:
            verr = electrical.compute(tfreq)
            terr = thermal.compute(1)
    """

    thermal_solver, electrical_solver = Static3D, Shockley3D

    def __init__(self, name=""):
        self.id = name
        self.thermal = self.thermal_solver(name + "-thermal")
        self.electrical = self.electrical_solver(name + "-electrical")
        self.tfreq = 6                                   # thermoelectric.py:57
        self.thermal.inHeat = self.electrical
        self.electrical.inTemperature = self.thermal
        self.history = []

    def initialize(self):
        if not self.thermal.initialized:
            self.thermal.initialize()
        if not self.electrical.initialized:
            self.electrical.initialize()

    def invalidate(self):
        self.thermal.invalidate()
        self.electrical.invalidate()

    def compute(self, invalidate=True, max_meta_loops=100):
        if invalidate:
            self.invalidate()
        self.initialize()
        t, e = self.thermal, self.electrical
        verr, terr = 2. * e.maxerr, 2. * t.maxerr
        n = 0
        while (terr > t.maxerr or verr > e.maxerr) and n < max_meta_loops:
            verr = e.compute(self.tfreq)
            terr = t.compute(1)
            n += 1
            self.history.append(dict(verr=verr, terr=terr, maxT=t.maxT))
        return n

    def get_total_current(self, nact=0):
        return self.electrical.get_total_current(nact)
