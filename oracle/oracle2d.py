"""CPU restatement of PLaSK's TWO-DIMENSIONAL FEM solvers — TEST INFRASTRUCTURE, never imported by plask_b200/ (SURVEY.md 8f-4).

    Static2D / StaticCyl     thermal.static      solvers/thermal/static/therm2d.cpp
    Dynamic2D / DynamicCyl   thermal.dynamic     solvers/thermal/dynamic/femT2d.cpp (corrected time loop, see Dynamic2DOracle)
    Shockley2D / ShockleyCyl electrical.shockley solvers/electrical/shockley/electr2d.cpp + beta.hpp

It is written independently of the 3-D code paths (oracle.py, fem3d_oracle.c, the CUDA library): 4-node rectangles assembled
entry by entry into a scipy sparse matrix and solved directly (SuperLU) — the stand-in for the reference's Cholesky — so that
the brick-mesh embedding the product uses for these solvers (plask_b200/solvers2d.py: one element layer, z-invariant data,
element weights r for the cylindrical case) is checked against genuinely two-dimensional arithmetic.

Pins: solvers/electrical/shockley/tests/shockley2d.py:40-118 (analytic current, capacitance and heat of Shockley2D and
ShockleyCyl, T-dependent beta) in tests/test_oracle2d_pin.py; Static2D / StaticCyl have no numeric fixture in the reference
(solvers/thermal/static/tests/therm.py checks providers only) and are pinned by analytic solutions — "parity unpinned" by
reference fixtures for the thermal pair.

Mesh convention: axes = [x (tran or r), y (vert)], node (i0, i1) -> i0 * n1 + i1, element (i0, i1) -> i0 * (n1 - 1) + i1."""
import numpy as np

EPS0 = 8.854187817e-12   # F/m  (plask/phys.hpp)


def table_lookup(tab, mat, T0, dT, T):
    """linear interpolation in the per-material tables, clamped at both ends (the host samples material->thermk / ->cond)"""
    tab = np.asarray(tab)
    nT = tab.shape[1]
    t = np.clip((np.asarray(T, dtype=np.float64) - T0) / dT, 0., float(nT - 1))
    i = np.minimum(t.astype(np.int64), nT - 2)
    f = t - i
    a, b = tab[mat, i], tab[mat, i + 1]
    return a + f * (b - a)


class Mesh2D:
    def __init__(self, x, y):
        self.x, self.y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
        self.n = (len(self.x), len(self.y))
        self.N = self.n[0] * self.n[1]
        self.E = (self.n[0] - 1) * (self.n[1] - 1)
        n0, n1 = self.n
        i0, i1 = np.meshgrid(np.arange(n0 - 1), np.arange(n1 - 1), indexing="ij")
        i0, i1 = i0.ravel(), i1.ravel()
        self.ei0, self.ei1 = i0, i1
        # lower-left, lower-right, upper-right, upper-left (the i1, i2, i3, i4 of therm2d.cpp:186-191)
        self.ll = i0 * n1 + i1
        self.lr = (i0 + 1) * n1 + i1
        self.ur = (i0 + 1) * n1 + i1 + 1
        self.ul = i0 * n1 + i1 + 1
        self.w = (self.x[1:] - self.x[:-1])[i0]
        self.h = (self.y[1:] - self.y[:-1])[i1]
        self.rmid = (0.5 * (self.x[1:] + self.x[:-1]))[i0]       # midpoint.rad_r()


def assemble(mesh, kx, ky, f, cyl):
    """setMatrix: therm2d.cpp:186-279 (Cartesian) / :338-420 (cylindrical: K_e and f_e times the midpoint radius); electr2d.cpp:281-316
    + setLocalMatrix (:190-230).  kx, ky: element conductivities already scaled by h/w and w/h.  Returns (A csr, B)."""
    import scipy.sparse as sp
    m = mesh
    r = m.rmid if cyl else 1.
    k_diag = r * (kx + ky) / 3.
    k_hor = r * (-2. * kx + ky) / 6.      # k21 = k43: lower-left <-> lower-right, upper-left <-> upper-right
    k_dia = r * -(kx + ky) / 6.           # k31 = k42
    k_ver = r * (kx - 2. * ky) / 6.       # k41 = k32: lower-left <-> upper-left, lower-right <-> upper-right
    nd = [m.ll, m.lr, m.ur, m.ul]
    kind = [[0, 1, 2, 3], [1, 0, 3, 2], [2, 3, 0, 1], [3, 2, 1, 0]]   # 0 diag, 1 horizontal edge, 2 diagonal, 3 vertical edge
    vals = [k_diag, k_hor, k_dia, k_ver]
    rows, cols, data = [], [], []
    for a in range(4):
        for b in range(4):
            rows.append(nd[a]); cols.append(nd[b]); data.append(vals[kind[a][b]])
    A = sp.coo_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))), shape=(m.N, m.N)).tocsr()
    B = np.zeros(m.N)
    if f is not None:
        fe = (r * f) if cyl else f
        for a in range(4):
            np.add.at(B, nd[a], fe)
    return A, B


def solve_dirichlet(A, B, nodes, values):
    """applyBC (matrix.hpp:111-118) + direct solve: rows of fixed nodes become identity, the columns move to the right-hand side"""
    import scipy.sparse.linalg as spl
    N = A.shape[0]
    fixed = np.zeros(N, dtype=bool)
    fixed[nodes] = True
    xD = np.zeros(N)
    xD[nodes] = values
    free = ~fixed
    X = xD.copy()
    if free.any():
        Aff = A[free][:, free].tocsc()
        rhs = B[free] - A[free][:, fixed] @ xD[fixed]
        X[free] = spl.spsolve(Aff, rhs)
    return X


SB = 5.670373e-8          # W/(m^2 K^4), plask/phys/constants.hpp:41


def _first_wins(N, conds, nval):
    """BoundaryConditionsWithMesh::getValue (plask/mesh/boundary_conditions.hpp:182-186): the first condition naming a node"""
    has = np.zeros(N, dtype=bool)
    vals = np.zeros((nval, N))
    for cond in conds:
        for n in np.asarray(cond[0], dtype=np.int64):
            if not has[n]:
                has[n] = True
                vals[:, n] = cond[1:1 + nval]
    return has, vals


def edge_terms(mesh, T, heatflux=(), convection=(), radiation=(), cyl=False, verbatim=True):
    """The boundary conditions of the 2nd / 3rd kind and radiation of ThermalFem2DSolver::setMatrix, element by element and side by
    side like setBoundaries (therm2d.cpp:138-172) with the lambdas of :225-265 (Cartesian) and :371-413 (cylindrical).
    Returns (K as coo triplets added to the element matrices, F added to the load vector).

    verbatim: the convection matrix terms exactly as the reference adds them — (c1 + c2) len / 6 and / 12 WITHOUT the 1e-6 that
    turns the edge length into metres (the load terms carry it: 0.5e-6 len c T_amb), and in the cylindrical solver added to
    k11 .. k41 BEFORE `A(i, j) += r * kij` (:415-426), i.e. multiplied by the midpoint radius a second time.
    verbatim = False: with the 1e-6 and the single radial factor."""
    m = mesh
    n1 = m.n[1]
    hf, (qf,) = _first_wins(m.N, heatflux, 1)
    hc, (cc, ca) = _first_wins(m.N, convection, 2)
    hr, (re, ra) = _first_wins(m.N, radiation, 2)
    rows, cols, data = [], [], []
    F = np.zeros(m.N)
    kunit = 1. if verbatim else 1e-6
    for e in np.nonzero((hf | hc | hr)[m.ll] | (hf | hc | hr)[m.lr] | (hf | hc | hr)[m.ur] | (hf | hc | hr)[m.ul])[0]:
        i1, i2, i3, i4 = int(m.ll[e]), int(m.lr[e]), int(m.ur[e]), int(m.ul[e])
        w, h, r = float(m.w[e]), float(m.h[e]), float(m.rmid[e])
        lo0, up0 = r - 0.5 * w, r + 0.5 * w                                     # elem.getLower0(), getUpper0()
        sides = [("BOTTOM", i1, i2, w), ("RIGHT", i2, i3, h), ("TOP", i3, i4, w), ("LEFT", i4, i1, h)]

        def radial(side, ia, ib, ln, offdiag=False):
            if not cyl:
                return 1.
            if side == "LEFT":
                return lo0
            if side == "RIGHT":
                return up0
            if offdiag:
                return r
            return r + (-ln / 6. if ia < ib else ln / 6.)

        kfac = kunit * (r if (cyl and verbatim) else 1.)
        for side, ia, ib, ln in sides:
            if hf[ia] and hf[ib]:
                for a, b_ in ((ia, ib), (ib, ia)):
                    F[a] += -0.5e-6 * ln * qf[a] * radial(side, a, b_, ln)
            if hc[ia] and hc[ib]:
                for a, b_ in ((ia, ib), (ib, ia)):
                    if cyl:
                        F[a] += 0.125e-6 * ln * (cc[a] + cc[b_]) * (ca[a] + ca[b_]) * radial(side, a, b_, ln)
                    else:
                        F[a] += 0.5e-6 * ln * cc[a] * ca[a]
                    rows.append(a); cols.append(a); data.append(kfac * (cc[a] + cc[b_]) * ln / 6. * radial(side, a, b_, ln))
                koff = kfac * (cc[ia] + cc[ib]) * ln / 12. * radial(side, ia, ib, ln, offdiag=True)
                rows += [ia, ib]; cols += [ib, ia]; data += [koff, koff]
            if hr[ia] and hr[ib]:
                for a, b_ in ((ia, ib), (ib, ia)):
                    F[a] += -0.5e-6 * ln * re[a] * SB * (T[a] ** 4 - ra[a] ** 4) * radial(side, a, b_, ln)
    return (np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64), np.asarray(data, dtype=np.float64)), F


class Static2DOracle:
    """ThermalFem2DSolver<Cartesian / Cylindrical> (therm2d.cpp): nonlinear loop of compute (:438-492) with boundary conditions of
    the first kind, the volumetric heat source and — heatflux / convection / radiation lists of (nodes, values...) — the
    conditions of the 2nd / 3rd kind and radiation (edge_terms)."""

    heatflux = convection = radiation = ()
    verbatim = True

    def __init__(self, x, y, elem_mat, T0, dT, tab_lat, tab_vert, bc_nodes, bc_values, heat=None, inittemp=300., maxerr=0.05, cyl=False):
        self.mesh = Mesh2D(x, y)
        self.elem_mat = np.asarray(elem_mat, dtype=np.int64)
        self.T0, self.dT, self.tab_lat, self.tab_vert = T0, dT, np.asarray(tab_lat), np.asarray(tab_vert)
        self.bc_nodes, self.bc_values = np.asarray(bc_nodes, dtype=np.int64), np.asarray(bc_values, dtype=np.float64)
        self.heat = np.zeros(self.mesh.E) if heat is None else np.asarray(heat, dtype=np.float64)
        self.maxerr, self.cyl = maxerr, cyl
        self.temperatures = np.full(self.mesh.N, float(inittemp))
        self.history = []
        self.maxT = float(inittemp)
        self.conds = None

    def _conds(self):
        m, T = self.mesh, self.temperatures
        temp = 0.25 * (T[m.ll] + T[m.lr] + T[m.ul] + T[m.ur])                                    # :200
        return (table_lookup(self.tab_lat, self.elem_mat, self.T0, self.dT, temp),
                table_lookup(self.tab_vert, self.elem_mat, self.T0, self.dT, temp))

    def compute(self, loops=0):
        m = self.mesh
        loop, toterr = 0, 0.
        while True:
            kl, kv = self._conds()
            self.conds = np.stack([kl, kv], axis=1)
            kx = kl * m.h / m.w                                                                    # :206-207
            ky = kv * m.w / m.h
            f = 0.25e-12 * m.w * m.h * self.heat                                                   # :210
            A, B = assemble(m, kx, ky, f, self.cyl)
            if self.heatflux or self.convection or self.radiation:
                import scipy.sparse as sp
                (kr, kc, kd), F = edge_terms(m, self.temperatures, self.heatflux, self.convection, self.radiation, self.cyl, self.verbatim)
                A = A + sp.coo_matrix((kd, (kr, kc)), shape=(m.N, m.N)).tocsr()
                B = B + F
            Tn = solve_dirichlet(A, B, self.bc_nodes, self.bc_values)
            err = float(np.abs(Tn - self.temperatures).max())
            self.temperatures = Tn
            self.maxT = float(Tn.max())
            toterr = max(toterr, err)
            loop += 1
            self.history.append(dict(loop=loop, err=err, maxT=self.maxT))
            if not (err > self.maxerr and (loops == 0 or loop < loops)):
                break
        return toterr

    def heat_fluxes(self):
        """saveHeatFluxes (therm2d.cpp:494-527): -k grad T at the element midpoints, W/m^2 (1e6: um -> m)"""
        m, T = self.mesh, self.temperatures
        kl, kv = self._conds()
        gx = 0.5e6 * (-T[m.ll] + T[m.lr] - T[m.ul] + T[m.ur]) / m.w
        gy = 0.5e6 * (-T[m.ll] - T[m.lr] + T[m.ul] + T[m.ur]) / m.h
        return np.stack([-kl * gx, -kv * gy], axis=1)


class Dynamic2DOracle:
    """DynamicThermalFem2DSolver<Cartesian / Cylindrical> (solvers/thermal/dynamic/femT2d.cpp) as a CORRECTED specification.

    setMatrix (:127-255 Cartesian, :258-388 cylindrical) to the letter: A = methodparam K + C, B = -(1 - methodparam) K + C with the
    4-node rectangle K (:176-181; the cylindrical solver multiplies kx, ky, c and f by the midpoint radius, :307-313), the element
    capacity c = cp dens 0.25e-12 w h / timestep / 1e-9 (:164), lumped (diag c, :186-189) or consistent (4/9, 2/9, 1/9 c,
    :212-222), F = 0.25e-12 w h heat (:170), conductivities and capacities at the mean of the four node temperatures (:157).
    The time loop (:415-445) eliminates the Dirichlet rows in A and F only (:247) while B.mult keeps them (:425), so a fixed node
    would receive value + (B T)_r; here, like the 3-D oracle (oracle.Dynamic3DOracle, which documents the same defect of
    femT3d.cpp):  A T' = B T + F on the free rows, T' = value on the Dirichlet rows.  cprho: [nmat][nT] table of cp(T) dens(T)."""

    def __init__(self, x, y, elem_mat, T0, dT, tab_lat, tab_vert, cprho, bc_nodes, bc_values, heat=None, inittemp=300., cyl=False,
                 timestep=0.1, methodparam=0.5, lumping=True, rebuildfreq=0):
        self.mesh = Mesh2D(x, y)
        self.elem_mat = np.asarray(elem_mat, dtype=np.int64)
        self.T0, self.dT, self.tab_lat, self.tab_vert = T0, dT, np.asarray(tab_lat), np.asarray(tab_vert)
        self.cprho = np.asarray(cprho, dtype=np.float64)
        self.bc_nodes, self.bc_values = np.asarray(bc_nodes, dtype=np.int64), np.asarray(bc_values, dtype=np.float64)
        self.heat = np.zeros(self.mesh.E) if heat is None else np.asarray(heat, dtype=np.float64)
        self.cyl = cyl
        self.timestep, self.methodparam, self.lumping, self.rebuildfreq = float(timestep), float(methodparam), bool(lumping), int(rebuildfreq)
        self.temperatures = np.full(self.mesh.N, float(inittemp))
        self.elapstime = 0.
        self.physical_time = 0.
        self.maxT_log = []

    def set_matrix(self):
        import scipy.sparse as sp
        m, T = self.mesh, self.temperatures
        temp = 0.25 * (T[m.ll] + T[m.lr] + T[m.ul] + T[m.ur])
        kl = table_lookup(self.tab_lat, self.elem_mat, self.T0, self.dT, temp)
        kv = table_lookup(self.tab_vert, self.elem_mat, self.T0, self.dT, temp)
        c = table_lookup(self.cprho, self.elem_mat, self.T0, self.dT, temp) * 0.25 * 1e-12 * m.h * m.w / self.timestep / 1e-9
        if self.cyl:
            c = c * m.rmid
        K, F = assemble(m, kl * m.h / m.w, kv * m.w / m.h, 0.25e-12 * m.w * m.h * self.heat, self.cyl)
        nd = [m.ll, m.lr, m.ur, m.ul]
        rows, cols, data = [], [], []
        for a in range(4):
            for b in range(4):
                if self.lumping:
                    wgt = 1. if a == b else 0.
                else:
                    wgt = 4. / 9. if a == b else (1. / 9. if (a + b) % 2 == 0 else 2. / 9.)     # opposite corners 1/9, edges 2/9
                if wgt:
                    rows.append(nd[a]); cols.append(nd[b]); data.append(wgt * c)
        Cm = sp.coo_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))), shape=(m.N, m.N)).tocsr()
        th = self.methodparam
        return (th * K + Cm).tocsr(), (-(1. - th) * K + Cm).tocsr(), F

    def _factor(self):
        import scipy.sparse.linalg as spl
        A, B, F = self.set_matrix()
        fixed = np.zeros(self.mesh.N, dtype=bool)
        fixed[self.bc_nodes] = True
        xD = np.zeros(self.mesh.N)
        xD[self.bc_nodes] = self.bc_values
        free = ~fixed
        return dict(solve=spl.factorized(A[free][:, free].tocsc()), B=B, F=F, free=free, fixed=fixed, xD=xD,
                    lift=A[free][:, fixed] @ xD[fixed])

    def compute(self, time):
        """compute(time), femT2d.cpp:390-452"""
        sysm = self._factor()
        r = self.rebuildfreq
        tend = time + self.timestep / 2.
        t = 0.
        while t < tend:
            if self.rebuildfreq and r == 0:
                sysm = self._factor()
                r = self.rebuildfreq
            rhs = sysm["B"] @ self.temperatures + sysm["F"]
            Tn = np.empty_like(self.temperatures)
            Tn[sysm["free"]] = sysm["solve"](rhs[sysm["free"]] - sysm["lift"])
            Tn[sysm["fixed"]] = sysm["xD"][sysm["fixed"]]
            self.temperatures = Tn
            self.maxT_log.append(float(Tn.max()))
            r -= 1
            self.elapstime += self.timestep
            self.physical_time += self.timestep
            t += self.timestep
        self.elapstime -= self.timestep
        self.maxT = float(self.temperatures.max())
        return 0.


class Shockley2DOracle:
    """ElectricalFem2DSolver<Cartesian / Cylindrical> + BetaSolver (electr2d.cpp, beta.hpp:43-46).  Junctions: `acts` = list of
    dicts(left, right, bottom, top) in element / node-row indices like setupActiveRegions (:61-168); elem_junc[e] = k + 1 for the
    elements of junction k; elem_role: 1 p-contact, 2 n-contact."""

    def __init__(self, x, y, elem_mat, T0, dT, tab_lat, tab_vert, bc_nodes, bc_values, elem_junc, acts, elem_role=None,
                 beta=11., js=1., pcond=5., ncond=50., start_cond=(0., 5.), maxerr=0.05, Te=300., cyl=False, length=1000.,
                 stable=False, noheat=None, eps=None):
        self.mesh = Mesh2D(x, y)
        m = self.mesh
        self.elem_mat = np.asarray(elem_mat, dtype=np.int64)
        self.T0, self.dT, self.tab_lat, self.tab_vert = T0, dT, np.asarray(tab_lat), np.asarray(tab_vert)
        self.bc_nodes, self.bc_values = np.asarray(bc_nodes, dtype=np.int64), np.asarray(bc_values, dtype=np.float64)
        self.elem_junc = np.asarray(elem_junc, dtype=np.int64)
        self.elem_role = np.zeros(m.E, dtype=np.int64) if elem_role is None else np.asarray(elem_role, dtype=np.int64)
        self.acts = [dict(a) for a in acts]
        off = 0
        for a in self.acts:
            a["offset"] = off
            a["height"] = m.y[a["top"]] - m.y[a["bottom"]]                                         # :156
            off += a["right"] - a["left"]                                                          # :157
        self.junction_conductivity = np.tile(np.asarray(start_cond, dtype=np.float64), (max(off, 1), 1))
        self.beta = beta if callable(beta) else (lambda T, b=float(beta): b)
        self.js = js if callable(js) else (lambda T, j=float(js): j)
        self.pcond, self.ncond, self.maxerr, self.cyl, self.length, self.stable = pcond, ncond, maxerr, cyl, length, stable
        self.Te = np.full(m.E, float(Te)) if np.isscalar(Te) else np.asarray(Te, dtype=np.float64)
        self.noheat = np.zeros(m.E, dtype=bool) if noheat is None else np.asarray(noheat).astype(bool)
        self.eps = np.ones(m.E) if eps is None else np.asarray(eps, dtype=np.float64)
        self.potentials = np.zeros(m.N)
        self.currents = np.zeros((m.E, 2))
        self.conds = np.zeros((m.E, 2))
        self.loopno = 0
        self.history = []
        self.maxcur = np.zeros(2)

    def _elem(self, i0, i1):
        return i0 * (self.mesh.n[1] - 1) + i1

    def load_conductivities(self):
        """loadConductivities, electr2d.cpp:330-352"""
        m = self.mesh
        c = table_lookup(self.tab_lat, self.elem_mat, self.T0, self.dT, self.Te)
        d = table_lookup(self.tab_vert, self.elem_mat, self.T0, self.dT, self.Te)
        self.conds = np.stack([c, d], axis=1)
        self.conds[self.elem_role == 1] = self.pcond
        self.conds[self.elem_role == 2] = self.ncond
        for e in np.nonzero(self.elem_junc)[0]:
            a = self.acts[self.elem_junc[e] - 1]
            self.conds[e] = self.junction_conductivity[a["offset"] + m.ei0[e]]                     # :341 (absolute index0)
            if np.isnan(self.conds[e, 1]) or abs(self.conds[e, 1]) < 1e-16:
                self.conds[e, 1] = 1e-16

    def save_conductivities(self):
        """saveConductivities, electr2d.cpp:354-360"""
        for a in self.acts:
            r = (a["top"] + a["bottom"]) // 2
            for i in range(a["left"], a["right"]):
                self.junction_conductivity[a["offset"] + i] = self.conds[self._elem(i, r)]

    def _update_junctions(self):
        """setMatrix :238-262"""
        m, V = self.mesh, self.potentials
        n1 = m.n[1]
        for e in np.nonzero(self.elem_junc)[0]:
            k = self.elem_junc[e] - 1
            a = self.acts[k]
            left, right = m.ei0[e], m.ei0[e] + 1
            U = 0.5 * (V[left * n1 + a["top"]] - V[left * n1 + a["bottom"]] + V[right * n1 + a["top"]] - V[right * n1 + a["bottom"]])
            jy = 0.1 * self.conds[e, 1] * U / a["height"]
            ti = self._elem(m.ei0[e], (a["top"] + a["bottom"]) // 2)
            T = self.Te[ti]
            jy = abs(jy)
            with np.errstate(divide="ignore", invalid="ignore"):
                cond = np.array([0., 10. * jy * a["height"] * self.beta(T) / np.log(1e7 * jy / self.js(T) + 1.)])   # beta.hpp:43-46
            if self.stable:
                cond = 0.5 * (self.conds[e] + cond)
            self.conds[e] = cond
            if np.isnan(self.conds[e, 1]) or abs(self.conds[e, 1]) < 1e-16:
                self.conds[e, 1] = 1e-16

    def compute(self, loops=0):
        """compute, electr2d.cpp:362-440"""
        m = self.mesh
        self.load_conductivities()
        noactive = len(self.acts) == 0
        minj = 100e-7
        loop, toterr = 0, 0.
        while True:
            if self.loopno != 0:
                self._update_junctions()
            kx = self.conds[:, 0] * m.h / m.w
            ky = self.conds[:, 1] * m.w / m.h
            A, B = assemble(m, kx, ky, None, self.cyl)
            self.potentials = V = solve_dirichlet(A, B, self.bc_nodes, self.bc_values)
            dvx = -0.05 * (-V[m.ll] + V[m.lr] - V[m.ul] + V[m.ur]) / m.w
            dvy = -0.05 * (-V[m.ll] - V[m.lr] + V[m.ul] + V[m.ur]) / m.h
            cur = np.stack([self.conds[:, 0] * dvx, self.conds[:, 1] * dvy], axis=1)
            a2 = (cur ** 2).sum(axis=1)
            sel = a2 if noactive else np.where(self.elem_junc > 0, a2, -1.)
            mcur = 0.
            if sel.size and sel.max() > 0.:
                k = int(np.argmax(sel))
                mcur, self.maxcur = float(sel[k]), cur[k].copy()
            err = float(((self.currents - cur) ** 2).sum(axis=1).max())
            self.currents = cur
            mcur = np.sqrt(mcur)
            err = 100. * np.sqrt(err) / max(mcur, minj)
            if (loop != 0 or mcur >= minj) and err > toterr:
                toterr = err
            self.loopno += 1
            loop += 1
            self.history.append(dict(loop=loop, err=err, mcur=mcur))
            if not (err > self.maxerr and (loops == 0 or loop < loops)):
                break
        self.save_conductivities()
        return toterr

    # ---- integrals (electr2d.cpp:466-505, 574-649)
    def integrate_current(self, vindex, onlyactive=False):
        m = self.mesh
        res = 0.
        for i in range(m.n[0] - 1):
            e = self._elem(i, vindex)
            if not onlyactive or self.elem_junc[e]:
                if self.cyl:
                    res += self.currents[e, 1] * (m.x[i + 1] ** 2 - m.x[i] ** 2)
                else:
                    res += self.currents[e, 1] * (m.x[i + 1] - m.x[i])
        return res * np.pi * 0.01 if self.cyl else res * self.length * 0.01

    def get_total_current(self, nact=0):
        a = self.acts[nact]
        return self.integrate_current((a["bottom"] + a["top"]) // 2, True)

    def heat_densities(self):
        m, V = self.mesh, self.potentials
        dvx = 0.5e6 * (-V[m.ll] + V[m.lr] - V[m.ul] + V[m.ur]) / m.w
        dvy = 0.5e6 * (-V[m.ll] - V[m.lr] + V[m.ul] + V[m.ur]) / m.h
        return np.where(self.noheat, 0., self.conds[:, 0] * dvx * dvx + self.conds[:, 1] * dvy * dvy)

    def get_total_heat(self):
        m = self.mesh
        H = self.heat_densities()
        if self.cyl:
            return 2e-15 * np.pi * float((m.w * m.h * m.rmid * H).sum())
        return self.length * 1e-15 * float((m.w * m.h * H).sum())

    def get_total_energy(self):
        m, V = self.mesh, self.potentials
        dvx = 0.5e6 * (-V[m.ll] + V[m.lr] - V[m.ul] + V[m.ur]) / m.w
        dvy = 0.5e6 * (-V[m.ll] - V[m.lr] + V[m.ul] + V[m.ur]) / m.h
        w = self.eps * (dvx * dvx + dvy * dvy)
        if self.cyl:
            return 2. * np.pi * 0.5e-18 * EPS0 * float((m.w * m.h * m.rmid * w).sum())
        return self.length * 0.5e-18 * EPS0 * float((m.w * m.h * w).sum())

    def get_capacitance(self):
        vals = np.unique(self.bc_values)
        assert len(vals) == 2
        U = float(vals[1] - vals[0])
        return 2e12 * self.get_total_energy() / (U * U)


def interp_bilinear(xs, ys, v, xq, yq):
    """interpolateLinear on a rectangular 2-D mesh (plask/mesh/rectangular2d.hpp: points outside are clamped to the edge):
    v[len(xs) * len(ys)] in the numbering i0 * n1 + i1 -> values on the product grid xq x yq, same numbering"""
    xs, ys, xq, yq = (np.asarray(a, dtype=np.float64) for a in (xs, ys, xq, yq))
    v = np.asarray(v, dtype=np.float64).reshape(len(xs), len(ys))

    def axis(a, q):
        if len(a) == 1:
            return np.zeros(len(q), dtype=np.int64), np.zeros(len(q), dtype=np.int64), np.zeros(len(q))
        q = np.clip(q, a[0], a[-1])
        hi = np.clip(np.searchsorted(a, q, side="right"), 1, len(a) - 1)
        lo = hi - 1
        return lo, hi, (q - a[lo]) / (a[hi] - a[lo])
    i0, i1, fx = axis(xs, xq)
    j0, j1, fy = axis(ys, yq)
    fx, fy = fx[:, None], fy[None, :]
    out = ((1. - fx) * (1. - fy) * v[i0][:, j0] + fx * (1. - fy) * v[i1][:, j0] +
           (1. - fx) * fy * v[i0][:, j1] + fx * fy * v[i1][:, j1])
    return out.ravel()


class ThermoElectric2DOracle:
    """meta.shockley.ThermoElectric2D / ThermoElectricCyl (solvers/meta/shockley/thermoelectric.py:187-211) over Static2DOracle and
    Shockley2DOracle: electrical.inTemperature = thermal.outTemperature at the electrical element midpoints (electr2d.cpp:330-335),
    thermal.inHeat = electrical.outHeat — element-mesh data interpolated at the thermal element midpoints, 0 outside the extent of
    the electrical mesh (getHeatDensities, electr2d.cpp:604-618)."""

    def __init__(self, thermal, electrical, tfreq=6):
        self.thermal, self.electrical, self.tfreq = thermal, electrical, tfreq
        self.history = []

    def exchange_temperature(self):
        t, e = self.thermal.mesh, self.electrical.mesh
        mid = lambda a: 0.5 * (a[1:] + a[:-1])
        self.electrical.Te = interp_bilinear(t.x, t.y, self.thermal.temperatures, mid(e.x), mid(e.y))

    def exchange_heat(self):
        t, e = self.thermal.mesh, self.electrical.mesh
        mid = lambda a: 0.5 * (a[1:] + a[:-1])
        xq, yq = mid(t.x), mid(t.y)
        heat = interp_bilinear(mid(e.x), mid(e.y), self.electrical.heat_densities(), xq, yq)
        inside = ((xq >= e.x[0]) & (xq <= e.x[-1]))[:, None] & ((yq >= e.y[0]) & (yq <= e.y[-1]))[None, :]
        self.thermal.heat = np.where(inside.ravel(), heat, 0.)

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
            self.history.append(dict(verr=verr, terr=terr, maxT=t.maxT, current=e.get_total_current()))
        return n
