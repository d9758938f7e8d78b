"""Shared test helpers: build the CPU oracle objects from a plask_b200.configs.Problem."""
import numpy as np

from oracle import oracle as orc
from plask_b200 import configs as cf


def oracle_mesh(p):
    return orc.Mesh(p.axes[0], p.axes[1], p.axes[2], p.order)


def oracle_thermal(p, **kw):
    m = oracle_mesh(p)
    tb = orc.Tables(p.T0, p.dT, p.tab_lat, p.tab_vert)
    return orc.Static3DOracle(m, p.elem_mat, tb, p.bc_nodes, p.bc_values, heat=p.heat, inittemp=p.inittemp,
                              maxerr=p.maxerr, **kw)


def oracle_shockley(p, **kw):
    m = oracle_mesh(p)
    tb = orc.Tables(p.T0, p.dT, p.tab_lat, p.tab_vert)
    kw.setdefault("beta", p.beta)
    kw.setdefault("js", p.js)
    kw.setdefault("maxerr", p.maxerr)
    return orc.Shockley3DOracle(m, p.elem_mat, tb, p.bc_nodes, p.bc_values, elem_junc=p.elem_junc,
                                elem_role=p.elem_role, pcond=p.pcond, ncond=p.ncond, start_cond=p.start_cond,
                                noheat=p.noheat, **kw)


def random_problem(n, order, seed=20261017, nd_frac=0.08, kind="thermal"):
    """Small mesh with jittered spacings, log-uniform conductivities per element (via one material id
    per element would be huge; use 5 ids with anisotropic tables) and a random Dirichlet set."""
    rng = np.random.default_rng(seed)
    axes = []
    for k in n:
        h = 1. + 0.4 * (rng.random(k - 1) - 0.5)
        h *= 10. ** rng.uniform(-1.5, 0.5)
        axes.append(np.concatenate([[0.], np.cumsum(h)]))
    nmat, nT = 5, 41
    T0, dT = 250., 10.
    base = 10. ** rng.uniform(-1, 2.6, size=(nmat, 1))
    T = T0 + dT * np.arange(nT)
    lat = base * (300. / T) ** rng.uniform(0., 1.5, size=(nmat, 1))
    vert = lat * 10. ** rng.uniform(-0.5, 0.5, size=(nmat, 1))
    p = cf.Problem("rand", kind, axes, order, None, T0, dT, lat, vert, None, None)
    p.elem_mat = rng.integers(0, nmat, size=p.E).astype(np.uint32)
    N = p.N
    nd = max(1, int(nd_frac * N))
    nodes = rng.choice(N, size=nd, replace=False)
    p.bc_nodes = nodes.astype(np.uintp)
    p.bc_values = rng.uniform(290., 310., size=nd)
    p.heat = 10. ** rng.uniform(12, 16, size=p.E)
    return p


def shockley3d_reference_problem(order="optimal"):
    """The structure of solvers/electrical/shockley/tests/shockley3d.py:36-60 as a Problem
    (same arrays as tests/test_oracle_pin.py::shockley3d_reference_case)."""
    def divide(edges, k):
        out = [edges[0]]
        for a, b in zip(edges[:-1], edges[1:]):
            out += [a + (b - a) * (i + 1) / k for i in range(k)]
        return np.array(out)

    x = divide([-500., -350., 350., 500.], 3)
    z = divide([0., 1., 301., 301.02, 601.02, 602.02], 2)
    n = (len(x), len(x), len(z))
    if order == "optimal":
        order = cf.optimal_order(n)
    gaas_cond = 1e2 * 1.60217733e-19 * 8000. * 1e16
    sig = np.array([[1e9, 1e9], [gaas_cond, gaas_cond], [0.55e-14, 0.55e-14]])
    p = cf.Problem("shockley3d.py", "shockley", [x, x.copy(), z], order, None, 300., 100., sig, sig.copy(), None, None)
    xm, zm = 0.5 * (x[1:] + x[:-1]), 0.5 * (z[1:] + z[:-1])
    X, Y, Z = np.meshgrid(xm, xm, zm, indexing="ij")
    mat = np.zeros(X.shape, dtype=np.uint32)
    mat[((Z < 1.) | (Z > 601.02)) & ((np.abs(X) > 350.) | (np.abs(Y) > 350.))] = 2
    is_j = (Z > 301.) & (Z < 301.02)
    mat[is_j] = 1
    p.elem_mat = p.to_elem_order(mat, np.uint32)
    p.elem_junc = p.to_elem_order(is_j.astype(np.uint32), np.uint32)
    p.noheat = (p.elem_mat == 2).astype(np.uint8)
    p.empty = p.noheat.copy()                       # air: Material::EMPTY, dropped by empty_elements='exclude'
    p.meta["eps"] = np.array([1., 12.9, 1.])[p.elem_mat]
    ng = np.broadcast_to(p.node_index_grid(), n)
    Xn, Yn = np.meshgrid(x, x, indexing="ij")
    inc = (np.abs(Xn) <= 350. + 1e-9) & (np.abs(Yn) <= 350. + 1e-9)
    top, bot = ng[:, :, -1][inc], ng[:, :, 0][inc]
    p.bc_nodes = np.concatenate([top, bot]).astype(np.uintp)
    p.bc_values = np.concatenate([np.zeros(top.size), np.ones(bot.size)])
    p.beta, p.js, p.maxerr = 10., 1., 1e-3
    return p


def face_nodes(p, axis, side):
    """node numbers of the mesh plane `side` (0 or -1) of physical axis `axis` (RectangularMesh<3>::getLeft/Right/...
    Boundary, plask/mesh/rectangular3d.hpp)"""
    ng = np.broadcast_to(p.node_index_grid(), p.n)
    return np.ascontiguousarray(np.take(ng, side, axis=axis).ravel(), dtype=np.int64)


def slab_problem_1d(n=(4, 4, 17), H=10., k=40., T0=300., order="012", dirichlet="bottom"):
    """uniform material of constant conductivity k on a box of height H um (non-uniform vertical spacing), no heat
    source, T0 on the bottom (or top) plane: every solution driven by conditions on horizontal planes is 1-D."""
    rng = np.random.default_rng(3)
    axes = [np.linspace(0., 3., n[0]), np.linspace(0., 2., n[1])]
    h = 1. + 0.6 * (rng.random(n[2] - 1) - 0.5)
    z = np.concatenate([[0.], np.cumsum(h)])
    axes.append(z * (H / z[-1]))
    tab = np.full((1, 2), float(k))
    p = cf.Problem("slab1d", "thermal", axes, order, None, 200., 1000., tab, tab.copy(), None, None)
    p.elem_mat = np.zeros(p.E, dtype=np.uint32)
    nodes = face_nodes(p, 2, 0 if dirichlet == "bottom" else -1)
    p.bc_nodes = nodes.astype(np.uintp)
    p.bc_values = np.full(nodes.size, float(T0))
    p.heat = np.zeros(p.E)
    p.inittemp = float(T0)
    p.maxerr = 1e-9
    return p


def assembled_csr(A14, mesh):
    """the oracle's SparseBandMatrix (14 diagonals of the upper triangle, iterative_matrix.hpp:372-390) as a scipy CSR matrix"""
    import scipy.sparse as sp
    N = mesh.N
    data = A14.data.reshape(14, N)
    ic = mesh.icords
    rows, cols, vals = [np.arange(N)], [np.arange(N)], [data[0].copy()]
    for i in range(14):
        d = int(ic[i])
        if d == 0:
            continue
        c = np.arange(N - d)
        v = data[i][:N - d]
        nz = v != 0
        rows += [c[nz] + d, c[nz]]
        cols += [c[nz], c[nz] + d]
        vals += [v[nz], v[nz]]
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))


def multilevel_reference(p, A, fixed, c=4):
    """M^-1 = sum_l P_l T_l^-1 P_l^T of kernels_ml.cuh built from the ASSEMBLED matrix with scipy: level 0 = vertical line
    blocks, level l = piecewise-constant aggregates of c^l x c^l lateral columns of free nodes, up to one column.
    Returns a function r -> z (full-mesh vectors in the problem's node numbering)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    n0, n1, n2 = p.n
    N = p.N
    ng = np.broadcast_to(p.node_index_grid(), p.n)
    free = np.ones(N)
    free[np.asarray(fixed, dtype=np.int64)] = 0.
    D = A.diagonal()
    free[D <= 0] = 0.

    def line_solver(P):
        """tridiagonal blocks of P^T A P along the vertical index; columns of P are numbered (aggregate, i2)"""
        Al = (P.T @ A @ P).tocsr()
        m = Al.shape[0]
        d = np.array(Al.diagonal())
        off = np.array(Al.diagonal(1)) if m > 1 else np.zeros(0)
        off = off * ((np.arange(m - 1) + 1) % n2 != 0)
        dead = d <= 0
        d = np.where(dead, 1., d)
        off[dead[:-1] | dead[1:]] = 0.
        lu = spla.splu(sp.diags([d, off, off], [0, 1, -1], format="csc"))
        return lambda r: lu.solve(np.where(dead, 0., r))

    levels = []
    f = 1
    while True:
        a0, a1 = np.arange(n0) // f, np.arange(n1) // f
        m0, m1 = a0.max() + 1, a1.max() + 1
        agg = (a0[:, None, None] * m1 + a1[None, :, None]) * n2 + np.arange(n2)[None, None, :]
        P = sp.csr_matrix((free[ng.ravel()], (ng.ravel(), np.broadcast_to(agg, p.n).ravel())), shape=(N, m0 * m1 * n2))
        levels.append((P, line_solver(P)))
        if m0 * m1 == 1:
            break
        f *= c
    def apply(r):
        z = np.zeros(N)
        for P, s in levels:
            z += P @ s(P.T @ r)
        return z
    apply.nlevels = len(levels)
    return apply


def oracle_dynamic(p, cprho=None, **kw):
    """Dynamic3DOracle (corrected femT3d.cpp) for a thermal Problem; cprho defaults to the problem's / the config tables"""
    m = oracle_mesh(p)
    tb = orc.Tables(p.T0, p.dT, p.tab_lat, p.tab_vert)
    if cprho is None:
        cprho = p.tab_cprho if p.tab_cprho is not None else cf.capacity_tables(p.T0, p.dT, p.tab_lat.shape[1])[:p.tab_lat.shape[0]]
    kw.setdefault("inittemp", p.inittemp)
    return orc.Dynamic3DOracle(m, p.elem_mat, tb, cprho, p.bc_nodes, p.bc_values, heat=p.heat, **kw)


def cooling_problem(nz=41, L=2.0, k=45., cprho=0.327e3 * 5.31749e3, order="012", nxy=(3, 4)):
    """1-D cooling of a slab of thickness L um: bottom plane at 300 K, top insulated, constant k and cp*dens (GaAs.cpp:225,243,249
    at 300 K).  With T(z,0) = 300 + a sin(pi z / 2L) the exact solution is 300 + a sin(pi z / 2L) exp(-alpha (pi/2L)^2 t),
    alpha = k/(cp dens) (in um^2/ns: * 1e3)."""
    axes = [np.linspace(0., 1., nxy[0]), np.linspace(0., 1.5, nxy[1]), np.linspace(0., L, nz)]
    tab = np.full((1, 2), float(k))
    p = cf.Problem("cooling", "thermal", axes, order, None, 200., 1000., tab, tab.copy(), None, None)
    p.elem_mat = np.zeros(p.E, dtype=np.uint32)
    nodes = face_nodes(p, 2, 0)
    p.bc_nodes = nodes.astype(np.uintp)
    p.bc_values = np.full(nodes.size, 300.)
    p.heat = np.zeros(p.E)
    p.inittemp = 300.
    p.tab_cprho = np.full((1, 2), float(cprho))
    p.meta["alpha"] = k / cprho * 1e3
    p.meta["L"] = L
    return p


def cooling_initial(p, a=50.):
    z = np.asarray(p.axes[2])
    ng = np.broadcast_to(p.node_index_grid(), p.n)
    T = np.empty(p.N)
    T[ng] = 300. + a * np.sin(np.pi * z / (2. * p.meta["L"]))[None, None, :]
    return T


def cooling_exact(p, t, a=50.):
    z = np.asarray(p.axes[2])
    L, al = p.meta["L"], p.meta["alpha"]
    return 300. + a * np.sin(np.pi * z / (2. * L)) * np.exp(-al * (np.pi / (2. * L)) ** 2 * t)


# ------------------------------------------------------------------------------------------ 2-D solvers

def _divide(edges, k):
    out = [edges[0]]
    for a, b in zip(edges[:-1], edges[1:]):
        out += [a + (b - a) * (i + 1) / k for i in range(k)]
    return np.array(out)


GAAS_COND = 1e2 * 1.60217733e-19 * 8000. * 1e16      # undoped GaAs at 300 K (like shockley3d_reference_problem)


def shockley2d_reference_problem(cyl=False, nx=2, ny=1):
    """solvers/electrical/shockley/tests/shockley2d.py as a Problem2D: Shockley2D_Test.setUp (:36-55) or, cyl=True, ShockleyCyl_Test.setUp
    (:78-103).  Meshes = DivideGenerator with prediv (2, 1) on the object boundaries (nx, ny refine them for convergence checks)."""
    from plask_b200.solvers2d import Problem2D
    if not cyl:
        x = _divide([0., 1000.], nx)
        y = _divide([0., 300., 300.02, 600.02], ny)
    else:
        x = _divide([0., 400., 600., 1000.], nx)
        y = _divide([0., 300., 300.02, 600.02, 700.02], ny)
    sig = np.array([[1e9, 1e9], [GAAS_COND, GAAS_COND], [0.55e-14, 0.55e-14]])
    xm, ym = 0.5 * (x[1:] + x[:-1]), 0.5 * (y[1:] + y[:-1])
    X, Y = np.meshgrid(xm, ym, indexing="ij")
    mat = np.zeros(X.shape, dtype=np.uint32)
    is_j = (Y > 300.) & (Y < 300.02)
    mat[is_j] = 1
    if cyl:
        mat[(Y > 600.02) & (X > 400.) & (X < 600.)] = 2          # the gap of the shelf: air
    n0, n1 = len(x), len(y)
    ng = np.arange(n0 * n1).reshape(n0, n1)
    if cyl:
        top = ng[(x <= 400. + 1e-9) | (x >= 600. - 1e-9), -1]    # mesh.TopOf(cont): both instances of the contact
    else:
        top = ng[:, -1]
    bot = ng[:, 0]
    p = Problem2D("shockley2d.py", "shockley", x, y, mat.ravel(), 300., 100., sig, sig.copy(),
                  np.concatenate([top, bot]).astype(np.uintp), np.concatenate([np.zeros(top.size), np.ones(bot.size)]), cyl=cyl)
    p.elem_junc = is_j.astype(np.uint32).ravel()
    p.noheat = (mat == 2).astype(np.uint8).ravel()
    p.meta["eps"] = np.array([1., 12.9, 1.])[mat.ravel()]
    p.beta, p.js, p.maxerr = 10., 1., 1e-5
    return p


def active_regions_2d(p2):
    """setupActiveRegions of electr2d.cpp:61-168 for rectangular junctions: one dict(left, right, bottom, top) per junction number"""
    ej = np.asarray(p2.elem_junc).reshape(p2.n[0] - 1, p2.n[1] - 1)
    acts = []
    for k in range(int(ej.max())):
        cols, rows = np.nonzero(ej == k + 1)
        acts.append(dict(left=int(cols.min()), right=int(cols.max()) + 1, bottom=int(rows.min()), top=int(rows.max()) + 1))
    return acts


def oracle_shockley2d(p2, **kw):
    from oracle import oracle2d
    kw.setdefault("beta", p2.beta)
    kw.setdefault("js", p2.js)
    kw.setdefault("maxerr", p2.maxerr)
    return oracle2d.Shockley2DOracle(p2.x, p2.y, p2.elem_mat, p2.T0, p2.dT, p2.tab_lat, p2.tab_vert, p2.bc_nodes, p2.bc_values,
                                     p2.elem_junc, active_regions_2d(p2), elem_role=p2.elem_role, pcond=p2.pcond, ncond=p2.ncond,
                                     start_cond=p2.start_cond, cyl=p2.cyl, length=p2.length, noheat=p2.noheat,
                                     eps=p2.meta.get("eps"), **kw)


def oracle_static2d(p2, **kw):
    from oracle import oracle2d
    return oracle2d.Static2DOracle(p2.x, p2.y, p2.elem_mat, p2.T0, p2.dT, p2.tab_lat, p2.tab_vert, p2.bc_nodes, p2.bc_values,
                                   heat=p2.heat, inittemp=p2.inittemp, maxerr=p2.maxerr, cyl=p2.cyl, **kw)


def oracle_dynamic2d(p2, **kw):
    """Dynamic2DOracle (corrected femT2d.cpp); cp * dens from the problem or the config tables of the thermal ids"""
    from oracle import oracle2d
    cprho = p2.tab_cprho if p2.tab_cprho is not None else cf.capacity_tables(p2.T0, p2.dT, p2.tab_lat.shape[1])[:p2.tab_lat.shape[0]]
    kw.setdefault("inittemp", p2.inittemp)
    return oracle2d.Dynamic2DOracle(p2.x, p2.y, p2.elem_mat, p2.T0, p2.dT, p2.tab_lat, p2.tab_vert, cprho, p2.bc_nodes, p2.bc_values,
                                    heat=p2.heat, cyl=p2.cyl, **kw)


def thermal2d_problem(n=(33, 41), cyl=False, seed=3):
    """layered GaAs / AlGaAs / Cu block with k(T) tables (the thermal ids of configs.thermal_tables), a hot disc / stripe near the
    axis, 300 K on the bottom edge; graded mesh in both directions"""
    from plask_b200.solvers2d import Problem2D
    x = cf.graded_axis(n[0], 0.3, 3.0)
    x = x - x[0]
    y = np.concatenate([[0.], np.cumsum(np.resize([0.07, 0.0795, 0.12, 0.5, 0.03], n[1] - 1))])
    T0, dT, lat, vert = cf.thermal_tables()
    xm, ym = 0.5 * (x[1:] + x[:-1]), 0.5 * (y[1:] + y[:-1])
    X, Y = np.meshgrid(xm, ym, indexing="ij")
    J = np.broadcast_to(np.arange(n[1] - 1)[None, :], X.shape)
    mat = (J % 2).astype(np.uint32)                      # GaAs / AlGaAs pairs
    mat[J < 4] = 5                                       # Cu heat spreader at the bottom
    heat = np.where((X < 0.35 * x[-1]) & (J > (n[1] - 1) // 2) & (J < (n[1] - 1) // 2 + 6), 4e16, 1e13)
    ng = np.arange(n[0] * n[1]).reshape(n)
    p = Problem2D("thermal2d", "thermal", x, y, mat.ravel(), T0, dT, lat, vert, ng[:, 0].astype(np.uintp), np.full(n[0], 300.),
                  heat=heat.ravel(), cyl=cyl)
    return p


def thermoelectric2d_pair(cyl=False, nx=19, thermal_shape=None, voltage=2.2):
    """A small mesa diode for the 2-D meta loop (ThermoElectric2D / ThermoElectricCyl): n-GaAs substrate 2 um, junction 0.02 um (two
    element rows), p-GaAs 1 um, contact on the top over x < 4 um, bottom grounded and held at 300 K.  sigma(T) and k(T) tabulated
    250..600 K, so that the exchanged temperatures matter.  thermal_shape = (n0, n1): the thermal solver on its own, coarser mesh
    that extends 1 um deeper (a heat-spreader the electrical mesh does not have: no Joule heat there)."""
    from plask_b200.solvers2d import Problem2D
    x = cf.graded_axis(nx, 0.25, 1.2)
    x = x - x[0]
    y = np.concatenate([np.linspace(0., 2., 9), [2.01, 2.02], 2.02 + np.linspace(0., 1., 7)[1:]])
    T0, dT, nT = 250., 0.5, 701
    Tg = T0 + dT * np.arange(nT)
    sig = np.stack([5.0e3 * (300. / Tg), np.full(nT, 1.0), 8.0e2 * (300. / Tg) ** 1.5])
    ym = 0.5 * (y[1:] + y[:-1])
    row_mat = np.where(ym < 2., 0, np.where(ym < 2.02, 1, 2)).astype(np.uint32)
    mat = np.tile(row_mat, len(x) - 1)
    n0, n1 = len(x), len(y)
    ng = np.arange(n0 * n1).reshape(n0, n1)
    top = ng[x <= 4.0 + 1e-9, -1]
    bot = ng[:, 0]
    pe = Problem2D("mesa-electrical", "shockley", x, y, mat, T0, dT, sig, sig.copy(),
                   np.concatenate([top, bot]).astype(np.uintp), np.concatenate([np.full(top.size, voltage), np.zeros(bot.size)]), cyl=cyl)
    pe.elem_junc = (mat == 1).astype(np.uint32)
    pe.beta, pe.js, pe.maxerr = 11., 1., 0.05
    if thermal_shape is None:
        xt, yt = x, y
    else:
        xt = np.linspace(0., x[-1], thermal_shape[0])
        yt = np.concatenate([[-1.0, -0.5], np.linspace(0., y[-1], thermal_shape[1] - 2)])
    k = np.stack([45. * (300. / Tg) ** 1.28])
    nt0, nt1 = len(xt), len(yt)
    ngt = np.arange(nt0 * nt1).reshape(nt0, nt1)
    pt = Problem2D("mesa-thermal", "thermal", xt, yt, np.zeros((nt0 - 1) * (nt1 - 1), dtype=np.uint32), T0, dT, k, k.copy(),
                   ngt[:, 0].astype(np.uintp), np.full(nt0, 300.), heat=None, cyl=cyl)
    return pt, pe
