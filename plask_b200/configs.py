"""Synthetic problem configurations of BASELINE.json, as flat arrays.

A `Problem` is exactly what the solver plugin would hand to the C ABI: axis coordinates, the
iteration order (strides), a material id per element, conductivity tables, Dirichlet nodes,
heat per element and (Shockley) the junction description.  Everything is deterministic; the
randomised variants use numpy.random.default_rng(20261017).
"""
from dataclasses import dataclass, field

import numpy as np

from . import materials as M

ORDERS = {"012": (0, 1, 2), "021": (0, 2, 1), "102": (1, 0, 2), "120": (1, 2, 0), "201": (2, 0, 1), "210": (2, 1, 0)}


def optimal_order(n):
    """RectilinearMesh3D::setOptimalIterationOrder (plask/mesh/rectilinear3d.cpp:74-85)."""
    for name in ("012", "021", "102", "120", "201", "210"):
        f, s, t = ORDERS[name]
        if n[t] <= n[s] <= n[f]:
            return name
    return "210"


def strides_for(n, order):
    """node and element strides of the physical axes for ORDER_<major><medium><minor>
    (rectilinear3d.cpp:20-32; the element mesh keeps the order, rectangular3d.cpp:20-22)."""
    mj, md, mn = ORDERS[order]
    ns, es = [0, 0, 0], [0, 0, 0]
    ns[mn], ns[md], ns[mj] = 1, n[mn], n[mn] * n[md]
    es[mn], es[md], es[mj] = 1, n[mn] - 1, (n[mn] - 1) * (n[md] - 1)
    return tuple(ns), tuple(es)


@dataclass
class Problem:
    name: str
    kind: str                      # 'thermal' | 'shockley'
    axes: list
    order: str
    elem_mat: np.ndarray           # uint32 [E], reference element order
    T0: float
    dT: float
    tab_lat: np.ndarray            # [nmat][nT]
    tab_vert: np.ndarray
    bc_nodes: np.ndarray           # uintp
    bc_values: np.ndarray
    heat: np.ndarray = None        # [E] W/m3 (thermal)
    inittemp: float = 300.
    maxerr: float = 0.05
    # Shockley only
    elem_junc: np.ndarray = None   # uint32 [E], 0 = none, k+1 = junction k
    elem_role: np.ndarray = None   # uint8 [E]: 1 p-contact, 2 n-contact
    noheat: np.ndarray = None      # uint8 [E]
    beta: float = 11.
    js: float = 1.
    pcond: float = 5.
    ncond: float = 50.
    start_cond: tuple = (0., 5.)
    Te: float = 300.
    empty: np.ndarray = None       # uint8 [E]: material kind EMPTY (the elements empty-elements="exclude" drops)
    tab_cprho: np.ndarray = None   # [nmat][nT] cp(T)*dens(T), J/(m^3 K) (Dynamic3D); None -> capacity_tables() of the thermal ids
    meta: dict = field(default_factory=dict)

    @property
    def n(self):
        return tuple(len(a) for a in self.axes)

    @property
    def N(self):
        n = self.n
        return n[0] * n[1] * n[2]

    @property
    def E(self):
        n = self.n
        return (n[0] - 1) * (n[1] - 1) * (n[2] - 1)

    @property
    def strides(self):
        return strides_for(self.n, self.order)[0]

    @property
    def estrides(self):
        return strides_for(self.n, self.order)[1]

    def elem_index_grid(self):
        n, es = self.n, self.estrides
        i0, i1, i2 = np.meshgrid(*[np.arange(k - 1) for k in n], indexing="ij", sparse=True)
        return i0 * es[0] + i1 * es[1] + i2 * es[2]

    def node_index_grid(self):
        n, ns = self.n, self.strides
        i0, i1, i2 = np.meshgrid(*[np.arange(k) for k in n], indexing="ij", sparse=True)
        return i0 * ns[0] + i1 * ns[1] + i2 * ns[2]

    def to_elem_order(self, grid3d, dtype=None):
        """(n0-1,n1-1,n2-1) array -> flat array in the reference element order"""
        out = np.empty(self.E, dtype=dtype or grid3d.dtype)
        out[self.elem_index_grid().ravel()] = np.broadcast_to(grid3d, tuple(k - 1 for k in self.n)).ravel()
        return out


# ------------------------------------------------------------------------------ helpers

def graded_axis(n, hmin, hmax):
    """n nodes, symmetric about 0, element sizes growing geometrically from hmin at the centre
    to hmax at the edge (adjacent ratio constant, like DivideGenerator's `gradual` meshes)."""
    ne = n - 1
    half = ne // 2
    r = (hmax / hmin) ** (1. / max(half - 1, 1))
    side = hmin * r ** np.arange(half)
    if ne % 2:
        h = np.concatenate([side[::-1], [hmin], side])
    else:
        h = np.concatenate([side[::-1], side])
    x = np.concatenate([[0.], np.cumsum(h)])
    return x - 0.5 * x[-1]


def graded_run(total, n, first):
    """n element sizes summing to `total`, growing geometrically from about `first`."""
    if n == 1:
        return np.array([total])
    lo, hi = 1.0, 50.0
    for _ in range(200):
        r = 0.5 * (lo + hi)
        s = first * (r ** n - 1) / (r - 1) if abs(r - 1) > 1e-12 else first * n
        if s < total:
            lo = r
        else:
            hi = r
    h = first * r ** np.arange(n)
    return h * (total / h.sum())


D_GAAS, D_ALGAAS = 0.0700, 0.0795   # DBR layer thicknesses, solvers/meta/shockley/tests/thermoelectric.xpl:26-30

# material ids shared by the configurations
GAAS, ALGAAS, QW, ALOX, AU, CU, AIR, NDBR_A, NDBR_B, PDBR_A, PDBR_B, NSUB = range(12)


def thermal_tables(T0=250., dT=0.25, nT=1601):
    models = [M.thermk_GaAs, lambda T: M.thermk_AlGaAs(T, 0.73), M.thermk_GaAs, M.thermk_AlOx, M.thermk_Au,
              M.thermk_Cu, M.thermk_air,
              M.thermk_GaAs, lambda T: M.thermk_AlGaAs(T, 0.73), M.thermk_GaAs, lambda T: M.thermk_AlGaAs(T, 0.73),
              M.thermk_GaAs]
    return M.sample_tables(models, T0, dT, nT)


def capacity_tables(T0=250., dT=0.25, nT=1601):
    """cp(T)*dens(T) per material id of thermal_tables (Dynamic3D, femT3d.cpp:176)"""
    gaas, algaas = M.cpdens_GaAs, lambda T: M.cpdens_AlGaAs(T, 0.73)
    models = [gaas, algaas, gaas, M.cpdens_const(880., 3950.), M.cpdens_const(129., 19300.), M.cpdens_const(385., 8960.),
              M.cpdens_const(1005., 1.2), gaas, algaas, gaas, algaas, gaas]
    return M.sample_table1(models, T0, dT, nT)


def electrical_tables(T0=250., dT=0.25, nT=1601):
    models = [M.cond_GaAs, M.cond_GaAs, M.cond_GaAs, M.cond_AlOx, M.cond_Au, M.cond_Cu, M.cond_air,
              lambda T: M.cond_doped(T, 2e18, 2000., 1.4),    # n-GaAs:Si
              lambda T: M.cond_doped(T, 2e18, 300., 1.4),     # n-AlGaAs:Si
              lambda T: M.cond_doped(T, 2e18, 100., 1.25),    # p-GaAs:C
              lambda T: M.cond_doped(T, 2e18, 40., 1.25),     # p-AlGaAs:C
              lambda T: M.cond_doped(T, 1e18, 2500., 1.4)]    # n-GaAs substrate
    return M.sample_tables(models, T0, dT, nT)


# --------------------------------------------------------------------------- config A

def config_A(n=64, order="optimal", heat=1e15, rows0=None):
    """Static3D on an n^3 GaAs/AlGaAs stack with uniform heat source (BASELINE configs[0]):
    lateral axes uniform over n um, vertical axis = DBR layer interfaces, one element per
    layer; T = 300 K on the bottom plane.  rows0 = (lo, hi): one slab (node rows lo..hi-1 of axis 0) of the mesh."""
    n = (n, n, n) if np.isscalar(n) else tuple(n)
    if rows0 is not None and order == "optimal":
        order = optimal_order(n)
    ax0 = np.linspace(0., float(n[0]), n[0])
    ax1 = np.linspace(0., float(n[1]), n[1])
    if rows0 is not None:
        ax0 = ax0[rows0[0]:rows0[1]]
        n = (len(ax0), n[1], n[2])
    hz = np.where(np.arange(n[2] - 1) % 2 == 0, D_GAAS, D_ALGAAS)
    ax2 = np.concatenate([[0.], np.cumsum(hz)])
    if order == "optimal":
        order = optimal_order(n)
    T0, dT, lat, vert = thermal_tables()
    p = Problem("A", "thermal", [ax0, ax1, ax2], order, None, T0, dT, lat[:2].copy(), vert[:2].copy(), None, None)
    mat = (np.arange(n[2] - 1) % 2).astype(np.uint32)[None, None, :]
    p.elem_mat = p.to_elem_order(mat, np.uint32)
    p.heat = np.full(p.E, float(heat))
    ng = p.node_index_grid()
    bottom = np.broadcast_to(ng, n)[:, :, 0].ravel()
    p.bc_nodes = bottom.astype(np.uintp)
    p.bc_values = np.full(bottom.size, 300.)
    return p


# ----------------------------------------------------------------- VCSEL-like (B, C, D)

def _vcsel_vertical(ne):
    """Vertical layer list (bottom -> top) with exactly `ne` elements:
    Cu sink, n-GaAs substrate, bottom DBR, cavity with QWs, oxide layer, top DBR, Au contact.
    Returns element sizes [um] and a tag per element."""
    n_cav = 11 if ne >= 80 else 3          # QW/barrier elements (role 'active' in config C)
    n_ox = 1
    n_au = 2 if ne >= 40 else 1
    rest = ne - n_cav - n_ox - n_au
    n_sub = max(2, rest // 8)
    n_cu = max(1, rest // 16)
    n_dbr = rest - n_sub - n_cu
    nb = (n_dbr * 5) // 9
    nt = n_dbr - nb
    h, tag = [], []
    for x in graded_run(200., n_cu, 2.)[::-1]:
        h.append(x); tag.append("cu")
    for x in graded_run(100., n_sub, 0.2)[::-1]:
        h.append(x); tag.append("sub")
    for k in range(nb):
        h.append(D_ALGAAS if k % 2 == 0 else D_GAAS); tag.append("nA" if k % 2 == 0 else "nG")
    for k in range(n_cav):
        h.append(0.005); tag.append("qw" if k % 2 else "bar")
    h.append(0.016); tag.append("ox")
    for k in range(nt):
        h.append(D_ALGAAS if k % 2 == 0 else D_GAAS); tag.append("pA" if k % 2 == 0 else "pG")
    for k in range(n_au):
        h.append(0.1); tag.append("au")
    assert len(h) == ne
    return np.array(h), tag


def _vcsel(n, order, kind, r_ap=4., r_mesa=15., hmin=0.25, hmax=4., rows0=None):
    """rows0 = (lo, hi): keep only node rows lo..hi-1 of axis 0 (direct construction of one slab of a large
    mesh without ever building the global arrays; the structure depends on the coordinates only)."""
    n = (n, n, n) if np.isscalar(n) else tuple(n)
    ax0 = graded_axis(n[0], hmin, hmax)
    ax1 = graded_axis(n[1], hmin, hmax)
    if rows0 is not None:
        ax0 = ax0[rows0[0]:rows0[1]]
        n = (len(ax0), n[1], n[2])
    hz, tag = _vcsel_vertical(n[2] - 1)
    ax2 = np.concatenate([[0.], np.cumsum(hz)])
    if order == "optimal":
        order = optimal_order(n)
    xm, ym = 0.5 * (ax0[1:] + ax0[:-1]), 0.5 * (ax1[1:] + ax1[:-1])
    R = np.sqrt(xm[:, None] ** 2 + ym[None, :] ** 2)          # (n0-1, n1-1)
    in_mesa, in_ap = R < r_mesa, R < r_ap
    ring = (R > 1.5 * r_ap) & (R < 0.8 * r_mesa)
    tagv = np.array(tag)
    thermal_id = {"cu": CU, "sub": GAAS, "nA": ALGAAS, "nG": GAAS, "bar": GAAS, "qw": QW, "ox": ALOX, "pA": ALGAAS,
                  "pG": GAAS, "au": AU}
    electr_id = {"cu": CU, "sub": NSUB, "nA": NDBR_B, "nG": NDBR_A, "bar": GAAS, "qw": GAAS, "ox": ALOX, "pA": PDBR_B,
                 "pG": PDBR_A, "au": AU}
    ids = thermal_id if kind == "thermal" else electr_id
    col = np.array([ids[t] for t in tag], dtype=np.uint32)     # per vertical element
    mat = np.broadcast_to(col[None, None, :], (n[0] - 1, n[1] - 1, n[2] - 1)).copy()
    above = np.isin(tagv, ["bar", "qw", "ox", "pA", "pG", "au"])  # the etched mesa
    mat[np.ix_(np.arange(n[0] - 1), np.arange(n[1] - 1), np.where(above)[0])] = np.where(
        in_mesa[:, :, None], mat[:, :, above], AIR)
    kox = np.where(tagv == "ox")[0]
    mat[:, :, kox] = np.where(in_ap[:, :, None], (ALGAAS if kind == "thermal" else PDBR_B),
                              np.where(in_mesa[:, :, None], ALOX, AIR))
    kau = np.where(tagv == "au")[0]
    mat[:, :, kau] = np.where(ring[:, :, None], AU, AIR)
    return n, [ax0, ax1, ax2], order, mat, tagv, dict(in_mesa=in_mesa, in_ap=in_ap, ring=ring)


def config_B(n=256, order="optimal", rows0=None):
    """Static3D, VCSEL-like layered block with nonlinear k(T) (BASELINE configs[1], 256^3)."""
    if rows0 is not None and order == "optimal":
        order = optimal_order((n, n, n) if np.isscalar(n) else tuple(n))
    n, axes, order, mat, tagv, reg = _vcsel(n, order, "thermal", rows0=rows0)
    T0, dT, lat, vert = thermal_tables()
    p = Problem("B", "thermal", axes, order, None, T0, dT, lat, vert, None, None)
    p.elem_mat = p.to_elem_order(mat, np.uint32)
    p.empty = (p.elem_mat == AIR).astype(np.uint8)
    heat = np.zeros(mat.shape)
    mesa_layers = np.isin(tagv, ["bar", "qw", "ox", "pA", "pG"])
    # ~14 mW in the active cylinder (r < r_ap, 55 nm thick) + ~6 mW of Joule heat spread over the mesa:
    # a 20 mW device with a temperature rise of a few tens of K, so k(T) matters but stays inside the tables
    heat[:, :, mesa_layers] = np.where(reg["in_mesa"][:, :, None], 1e12, 0.)
    cav = np.isin(tagv, ["bar", "qw"])
    heat[:, :, cav] = np.where(reg["in_ap"][:, :, None], 5e15, heat[:, :, cav])
    p.heat = p.to_elem_order(heat, np.float64)
    bottom = np.broadcast_to(p.node_index_grid(), n)[:, :, 0].ravel()
    p.bc_nodes = bottom.astype(np.uintp)
    p.bc_values = np.full(bottom.size, 300.)
    return p


def config_C(n=(192, 192, 400), order="optimal", voltage=1.4, rows0=None):
    """Shockley3D with oxide aperture and a nonlinear junction layer (BASELINE configs[2]).
    rows0 = (lo, hi): direct construction of one slab (node rows lo..hi-1 of axis 0), like config_B."""
    nglob = (n, n, n) if np.isscalar(n) else tuple(n)
    if rows0 is not None and order == "optimal":
        order = optimal_order(nglob)
    n, axes, order, mat, tagv, reg = _vcsel(n, order, "shockley", rows0=rows0)
    T0, dT, lat, vert = electrical_tables()
    p = Problem("C", "shockley", axes, order, None, T0, dT, lat, vert, None, None)
    p.elem_mat = p.to_elem_order(mat, np.uint32)
    junc = np.zeros(mat.shape, dtype=np.uint32)
    cav = np.isin(tagv, ["bar", "qw"])
    junc[:, :, cav] = np.where(reg["in_mesa"][:, :, None], 1, 0)
    p.elem_junc = p.to_elem_order(junc, np.uint32)
    p.elem_role = np.zeros(p.E, dtype=np.uint8)
    p.noheat = p.to_elem_order((mat == AIR).astype(np.uint8), np.uint8)
    p.empty = p.noheat.copy()
    p.maxerr = 0.05
    p.beta, p.js = 11., 1.   # thermoelectric.xpl:91
    ng = np.broadcast_to(p.node_index_grid(), n)
    # 0 V on the bottom plane (n side), `voltage` on the top face of the Au ring (p side)
    x, y = axes[0], axes[1]
    Rn = np.sqrt(x[:, None] ** 2 + y[None, :] ** 2)
    # nodes whose four surrounding top-layer elements are all Au
    ring_e = reg["ring"]
    if rows0 is not None:   # the ring of the whole mesh (2-D, cheap), cut to the local rows afterwards
        gx, gy = graded_axis(nglob[0], 0.25, 4.), axes[1]
        gxm, gym = 0.5 * (gx[1:] + gx[:-1]), 0.5 * (gy[1:] + gy[:-1])
        Rg = np.sqrt(gxm[:, None] ** 2 + gym[None, :] ** 2)
        ring_e = (Rg > 1.5 * 4.) & (Rg < 0.8 * 15.)
    ring_n = np.zeros((ring_e.shape[0] + 1, n[1]), dtype=bool)
    ring_n[1:-1, 1:-1] = ring_e[:-1, :-1] & ring_e[1:, :-1] & ring_e[:-1, 1:] & ring_e[1:, 1:]
    if rows0 is not None:
        ring_n = ring_n[rows0[0]:rows0[1]]
    top = ng[:, :, -1][ring_n]
    bottom = ng[:, :, 0].ravel()
    p.bc_nodes = np.concatenate([top, bottom]).astype(np.uintp)
    p.bc_values = np.concatenate([np.full(top.size, float(voltage)), np.zeros(bottom.size)])
    p.meta["Rn"] = Rn
    return p


def config_E(g=1, nz_per_gpu=96, nxy=512, order="012"):
    """Static3D weak-scaling case (BASELINE configs[4]): the config-A stack with the major axis
    extended proportionally to the number of GPUs g.  ORDER_012: the slab axis (major = axis 0) is
    lateral, so every GPU sees the same layer structure; the vertical axis is the minor one."""
    n = (nz_per_gpu * g, nxy, nxy)
    p = config_A(n, order=order)
    p.name = f"E{g}"
    return p


# --------------------------------------------------------------------- junction tables

def setup_active(p):
    """setupActiveRegions (electr3d.cpp:89-183) on the flat description: returns a list of dicts
    (pfem_junction fields) and the junction-table length."""
    n, es = p.n, p.estrides
    junc3 = p.elem_junc[p.elem_index_grid()] if p.elem_junc is not None else None
    acts, tot = [], 0
    if junc3 is None or junc3.max() == 0:
        return acts, 0
    junc3 = np.broadcast_to(junc3, tuple(k - 1 for k in n))
    for num in range(1, int(junc3.max()) + 1):
        m = junc3 == num
        if not m.any():
            acts.append(dict(bottom=0, top=1, left=0, right=0, back=0, front=0, ld=0, offset=0, height=1.))
            continue
        cols = m.any(axis=2)
        vert = m.any(axis=(0, 1))
        ks = np.where(vert)[0]
        bottom, top = int(ks[0]), int(ks[-1]) + 1
        # every active column must span exactly [bottom, top)
        span = m[cols]
        if not (span[:, bottom:top].all() and span.sum(axis=1).max() == top - bottom):
            raise ValueError(f"Junction {num - 1} does not have top and bottom edges at constant heights")
        i0 = np.where(cols.any(axis=1))[0]
        i1 = np.where(cols.any(axis=0))[0]
        back, front, left, right = int(i0[0]), int(i0[-1]) + 1, int(i1[0]), int(i1[-1]) + 1
        ld = front - back
        acts.append(dict(bottom=bottom, top=top, left=left, right=right, back=back, front=front, ld=ld,
                         offset=tot - ld * left - back, height=float(p.axes[2][top] - p.axes[2][bottom])))
        tot += (right - left) * (front - back)
    return acts, tot


# ------------------------------------------------------------------- layer thickness (a2)

def layer_thickness(p, key):
    """ThermalFem3DSolver::onInitialize (therm3d.cpp:81-114): thickness[e] = height of the maximal run of vertically
    adjacent elements of the same material that contains e.  `key`: material identity per element, [E] in element order
    (the plugin compares shared_ptr<Material> with operator==, material.hpp:836-872).  Returns [E] in element order."""
    n = p.n
    ne = tuple(k - 1 for k in n)
    eg = np.broadcast_to(p.elem_index_grid(), ne)
    k3 = np.asarray(key)[eg]                                    # (n0-1, n1-1, n2-1)
    hz = np.diff(np.asarray(p.axes[2], dtype=np.float64))
    brk = np.ones(ne, dtype=bool)                               # True where a new run starts
    brk[:, :, 1:] = k3[:, :, 1:] != k3[:, :, :-1]
    run = np.cumsum(brk, axis=2) - 1                            # run number of every element inside its column
    nrun = int(run.max()) + 1
    th = np.zeros(ne[:2] + (nrun,))
    i0, i1 = np.meshgrid(np.arange(ne[0]), np.arange(ne[1]), indexing="ij")
    for r in range(ne[2]):                                      # sum the element heights per (column, run)
        np.add.at(th, (i0, i1, run[:, :, r]), hz[r])
    t3 = np.take_along_axis(th, run, axis=2)
    return p.to_elem_order(t3, np.float64)


def thickness_material_ids(key, thickness):
    """One table id per distinct (material, layer thickness) pair, so that thermk(T, thickness) can be tabulated per id
    (pfem_set_materials).  Returns (ids [E] uint32, list of (material key, thickness) per id) — the Python twin of
    plaskfem::material_ids (include/plaskfem_cuda.hpp)."""
    key = np.asarray(key)
    pairs = np.stack([key.astype(np.float64), np.asarray(thickness, dtype=np.float64)], axis=1)
    uniq, inv = np.unique(pairs, axis=0, return_inverse=True)
    return inv.astype(np.uint32).ravel(), [(int(a), float(b)) for a, b in uniq]


# ------------------------------------------------------------------------- slab partition

def slab_range(nK, rank, nranks, align=1):
    """Owned node planes [K0, K1) of `rank` along the major axis (balanced, contiguous).  align > 1: every boundary between two
    ranks is a multiple of `align` planes (the multilevel preconditioner wants 16: its lateral aggregates must not straddle slabs)."""
    if align > 1:
        blocks = -(-nK // align)
        b0, b1 = slab_range(blocks, rank, nranks)
        return min(b0 * align, nK), min(b1 * align, nK)
    base, rem = divmod(nK, nranks)
    K0 = rank * base + min(rank, rem)
    return K0, K0 + base + (1 if rank < rem else 0)


def slab_local(nK, rank, nranks, align=1):
    """(lo, hi, own_lo, own_hi): local node planes [lo, hi) of the global major axis = owned planes plus one halo
    plane towards each neighbour, and the owned range in LOCAL plane indices."""
    K0, K1 = slab_range(nK, rank, nranks, align)
    lo, hi = K0 - (1 if rank > 0 else 0), K1 + (1 if rank < nranks - 1 else 0)
    return lo, hi, K0 - lo, K1 - lo


def slab_axis(p, need_vertical_inside=None):
    """Physical axis the slabs cut ("host chooses which physical axis that is").  The major axis of the mesh's own
    iteration order when that is lateral; for a vertical major axis (the order setOptimalIterationOrder gives a tall mesh,
    rectilinear3d.cpp:74-85) with a junction (a12 couples the planes act.bottom / act.top, electr3d.cpp:246-274) or whenever
    `need_vertical_inside` (line preconditioners solve along the vertical lines) the larger lateral axis — the local meshes
    are then handed to the library in an order with that axis major, which is only a host-side renumbering."""
    major = ORDERS[p.order][0]
    if need_vertical_inside is None:
        need_vertical_inside = p.kind == "shockley"
    if major != 2 or not need_vertical_inside:
        return major
    return 0 if p.n[0] >= p.n[1] else 1


def slab_order(order, axis):
    """iteration order of the local meshes: `axis` becomes the major axis, the other two keep their relative order"""
    rest = [a for a in ORDERS[order] if a != axis]
    return f"{axis}{rest[0]}{rest[1]}"


def slab_problem(p, rank, nranks, axis=None, align=1):
    """Cut the local problem of `rank` out of a global Problem (thermal or Shockley) — what the solver plugin would do on each
    process before calling the C ABI in slab mode.  `axis`: the physical axis to cut (default: the major axis of the mesh's
    order; see slab_axis); `align`: see slab_range.  Returns (local Problem, own_lo, own_hi, (lo, hi))."""
    major = ORDERS[p.order][0] if axis is None else int(axis)
    n = p.n
    lo, hi, own_lo, own_hi = slab_local(n[major], rank, nranks, align)
    if hi - lo < 2 or own_hi <= own_lo:
        raise ValueError(f"slab_problem: rank {rank} of {nranks} gets no planes of the {n[major]}-plane axis with align = {align}")
    axes = [a.copy() for a in p.axes]
    axes[major] = axes[major][lo:hi]
    q = Problem(p.name + f"[{rank}/{nranks}]", p.kind, axes, slab_order(p.order, major), None, p.T0, p.dT, p.tab_lat, p.tab_vert, None, None,
                inittemp=p.inittemp, maxerr=p.maxerr)
    esl = [slice(None)] * 3
    esl[major] = slice(lo, hi - 1)
    eg = p.elem_index_grid()
    shape_e = tuple(k - 1 for k in n)

    def cut_elem(a, dtype):
        if a is None:
            return None
        g3 = np.asarray(a)[np.broadcast_to(eg, shape_e)][tuple(esl)]
        return q.to_elem_order(g3, dtype)

    q.elem_mat = cut_elem(p.elem_mat, np.uint32)
    q.heat = cut_elem(p.heat, np.float64)
    q.elem_junc = cut_elem(p.elem_junc, np.uint32)
    q.elem_role = cut_elem(p.elem_role, np.uint8)
    q.noheat = cut_elem(p.noheat, np.uint8)
    q.empty = cut_elem(p.empty, np.uint8)
    q.tab_cprho = p.tab_cprho
    for name in ("beta", "js", "pcond", "ncond", "start_cond", "Te"):
        setattr(q, name, getattr(p, name))
    # Dirichlet nodes inside the local planes (halo planes included), in application order
    q.bc_nodes, keep = slab_local_nodes(p, q, lo, hi, p.bc_nodes, axis=major)
    q.bc_nodes = q.bc_nodes.astype(np.uintp)
    q.bc_values = np.asarray(p.bc_values)[keep].copy()
    return q, own_lo, own_hi, (lo, hi)


def slab_local_nodes(p, q, lo, hi, nodes, axis=None):
    """Global node numbers of `p` -> (node numbers in the local problem `q` holding planes [lo, hi) of the cut axis (default:
    the major axis of p), mask of the entries that lie inside) — for Dirichlet lists and the boundary conditions of the
    2nd / 3rd kind."""
    major = ORDERS[p.order][0] if axis is None else int(axis)
    ns = p.strides
    order3 = sorted(range(3), key=lambda a: -ns[a])            # major, medium, minor of the GLOBAL numbering
    rem = np.asarray(nodes, dtype=np.int64)
    c = [None] * 3
    for a in order3:
        c[a], rem = np.divmod(rem, ns[a])
    keep = (c[major] >= lo) & (c[major] < hi)
    c[major] = c[major] - lo
    qs = q.strides
    return (c[0][keep] * qs[0] + c[1][keep] * qs[1] + c[2][keep] * qs[2]).astype(np.int64), keep


def slab_field_owned(q, field, own_lo, own_hi):
    """Owned part of a local node field as an array shaped (owned planes, medium, minor)."""
    major, medium, minor = ORDERS[q.order]
    n = q.n
    return np.asarray(field).reshape(n[major], n[medium], n[minor])[own_lo:own_hi]


def slab_assemble(p, q_order, parts):
    """Global node field of `p` (in p's own numbering) from the owned parts of all ranks (slab_field_owned arrays, in rank
    order) of local problems numbered in `q_order`."""
    major, medium, minor = ORDERS[q_order]
    g = np.concatenate([np.asarray(a) for a in parts], axis=0)          # (n_major, n_medium, n_minor) of the local order
    g3 = np.transpose(g, np.argsort([major, medium, minor]))            # -> (n0, n1, n2)
    out = np.empty(p.N)
    out[np.broadcast_to(p.node_index_grid(), p.n).ravel()] = g3.ravel()
    return out
