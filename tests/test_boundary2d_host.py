"""Boundary conditions of the 2nd / 3rd kind and radiation of the 2-D thermal solvers (therm2d.cpp:138-172, :225-265, :371-413).

CPU part: (i) the oracle's restatement (oracle2d.edge_terms) against analytic 1-D solutions, (ii) the library's HOST flattening
(pfem_edges2d_host: the very code pfem_set_boundary runs in its 2-D mode, no device involved) against the oracle, entry by entry.
The GPU end-to-end comparison is tests/test_gpu_2d.py."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle2d
from plask_b200 import _lib as L


def _mesh(seed=3, n0=7, n1=6, r0=0.):
    rng = np.random.default_rng(seed)
    x = r0 + np.concatenate([[0.], np.cumsum(rng.uniform(0.4, 2.5, n0 - 1))])
    y = np.concatenate([[0.], np.cumsum(rng.uniform(0.05, 1.5, n1 - 1))])
    return x, y


def _conditions(m, rng, ragged=True):
    """conditions on all four sides with node-wise different values, overlapping at the corners (first definition wins), plus a
    partial edge and an isolated node (an edge needs BOTH nodes)"""
    n0, n1 = m.n
    node = lambda i0, i1: i0 * n1 + i1
    bottom = [node(i, 0) for i in range(n0)]
    top = [node(i, n1 - 1) for i in range(n0)]
    left = [node(0, j) for j in range(n1)]
    right = [node(n0 - 1, j) for j in range(n1)]
    heatflux = [(top, 2.5e5), (right[:3], -1.0e5)]
    convection = [(right, 4.0e3, 295.), (bottom[2:], 150., 310.), (left[1:4], 900., 280.)]
    radiation = [(left, 0.8, 290.), (top[:n0 // 2 + 1], 0.35, 305.)]
    if ragged:
        heatflux.append(([node(2, 2)], 7.0e4))                        # interior single node: no edge is complete
        convection.append(([node(3, 2), node(4, 2)], 50., 300.))      # interior horizontal edge: both adjacent elements see it
    return heatflux, convection, radiation


def _host_terms(x, y, heatflux, convection, radiation, cyl, verbatim):
    lib = L.load()
    N = len(x) * len(y)

    def dense(conds, nval):
        has = np.zeros(N, dtype=np.uint8)
        vals = [np.zeros(N) for _ in range(nval)]
        for cond in conds:
            nodes = np.asarray(cond[0], dtype=np.int64)
            new = nodes[has[nodes] == 0]
            for k in range(nval):
                vals[k][new] = cond[1 + k]
            has[new] = 1
        return has, vals
    hf, (qf,) = dense(heatflux, 1)
    hc, (cc, ca) = dense(convection, 2)
    hr, (re, ra) = dense(radiation, 2)
    b = L.Boundary()
    u8 = lambda a: a.ctypes.data_as(L._u8p)
    dp = lambda a: a.ctypes.data_as(L.c_dp)
    b.has_flux, b.flux = u8(hf), dp(qf)
    b.has_conv, b.conv_coeff, b.conv_ambient = u8(hc), dp(cc), dp(ca)
    b.has_rad, b.rad_emissivity, b.rad_ambient = u8(hr), dp(re), dp(ra)
    b.verbatim, b.mode2d = int(verbatim), 2 if cyl else 1
    x, y = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(y, dtype=np.float64)
    load, rc_, ra4, K = np.zeros(N), np.zeros(N), np.zeros(N), np.zeros(N * N)
    rc = lib.pfem_edges2d_host(len(x), dp(x), len(y), dp(y), C.byref(b), dp(load), dp(rc_), dp(ra4), dp(K))
    assert rc == 0
    return load, rc_, ra4, K.reshape(N, N)


@pytest.mark.parametrize("verbatim", [True, False])
@pytest.mark.parametrize("cyl", [False, True])
def test_host_flattening_equals_oracle(cyl, verbatim):
    x, y = _mesh(r0=0. if cyl else -3.)
    m = oracle2d.Mesh2D(x, y)
    rng = np.random.default_rng(11)
    heatflux, convection, radiation = _conditions(m, rng)
    T = rng.uniform(280., 420., m.N)
    (kr, kc, kd), F = oracle2d.edge_terms(m, T, heatflux, convection, radiation, cyl=cyl, verbatim=verbatim)
    Ko = np.zeros((m.N, m.N))
    np.add.at(Ko, (kr, kc), kd)
    load, rcoef, ramb4, K = _host_terms(x, y, heatflux, convection, radiation, cyl, verbatim)
    Fh = load - rcoef * (T ** 4 - ramb4)
    assert np.abs(Ko).max() > 0 and np.abs(F).max() > 0
    assert np.abs(K - Ko).max() <= 1e-14 * np.abs(Ko).max()
    assert np.abs(K - K.T).max() == 0.
    assert np.abs(Fh - F).max() <= 1e-13 * np.abs(F).max()
    # an interior edge is seen by the two elements that share it: twice the single-element term
    n1 = m.n[1]
    a, b_ = 3 * n1 + 2, 4 * n1 + 2
    ln = x[4] - x[3]
    unit = 1. if verbatim else 1e-6
    if not cyl:
        assert K[a, b_] == pytest.approx(2 * unit * (50. + 50.) * ln / 12., rel=1e-13)


def test_host_flattening_rejects_bad_input():
    lib = L.load()
    b = L.Boundary()
    z = np.zeros(4)
    dp = lambda a: a.ctypes.data_as(L.c_dp)
    b.mode2d = 0
    assert lib.pfem_edges2d_host(2, dp(z), 2, dp(z), C.byref(b), dp(z), dp(z), dp(z), dp(np.zeros(16))) == L.PFEM_ERR_BAD_INPUT
    b.mode2d = 1
    assert lib.pfem_edges2d_host(1, dp(z), 2, dp(z), C.byref(b), dp(z), dp(z), dp(z), dp(np.zeros(16))) == L.PFEM_ERR_BAD_INPUT


def _slab(n0=4, n1=41, H=10., W=3., cyl=False):
    """uniform k, laterally invariant: the solution depends on y only (also in the cylindrical solver: no radial flux)"""
    x = np.linspace(0., W, n0) + (2000. if cyl else 0.)    # a thin shell far from the axis: the radial factors are nearly constant
    y = np.linspace(0., H, n1)
    E = (n0 - 1) * (n1 - 1)
    k = 20.
    tab = np.full((1, 2), k)
    bottom = np.arange(n0) * n1
    top = bottom + n1 - 1
    o = oracle2d.Static2DOracle(x, y, np.zeros(E, dtype=np.int64), 300., 1000., tab, tab, bottom, np.full(n0, 300.), cyl=cyl)
    return o, k, H * 1e-6, top


@pytest.mark.parametrize("cyl", [False, True])
def test_oracle_heatflux_and_corrected_convection_analytic(cyl):
    # heat flux q on the top (W/m^2, the sign of therm2d.cpp:230: F -= q: positive q leaves the body): T_top = T0 - q H / k
    o, k, H, top = _slab(cyl=cyl)
    # (cylindrical: the edge loads carry the exact integral of r N_a, r -+ len/6, the stiffness the midpoint radius — the
    # reference's discretisation is not laterally invariant, by O(len / 6 r): 15 % on the axis, 1e-4 on this shell at r = 2 mm)
    tol = 5e-4 if cyl else 1e-10
    o.heatflux = [(top, -3.0e7)]
    o.compute(1)
    assert o.temperatures[top] - 300. == pytest.approx(3.0e7 * H / k, rel=tol)
    # convection on the top, corrected units: k (T_top - T0) / H = h (T_amb - T_top)
    o, k, H, top = _slab(cyl=cyl)
    h, Ta = 5.0e6, 400.
    o.convection, o.verbatim = [(top, h, Ta)], False
    o.compute(1)
    Ttop = (k / H * 300. + h * Ta) / (k / H + h)
    assert o.temperatures[top] - 300. == pytest.approx(Ttop - 300., rel=tol)
    # verbatim: the matrix term is 1e6 (Cartesian) times too strong against its own load: T_top = (k/H T0 + h Ta) / (k/H + 1e6 h)
    if not cyl:
        o, k, H, top = _slab()
        o.convection = [(top, h, Ta)]
        o.compute(1)
        assert o.temperatures[top] == pytest.approx((k / H * 300. + h * Ta) / (k / H + 1e6 * h), rel=1e-9)


def test_oracle_radiation_fixed_point():
    # radiation evaluated from the previous loop's temperatures (therm2d.cpp:256-262): the loop converges to k (T - T0) / H = -eps SB (T^4 - Ta^4)
    o, k, H, top = _slab()
    eps, Ta = 0.9, 2000.
    o.radiation = [(top, eps, Ta)]
    o.maxerr = 1e-9
    o.compute(200)
    T = o.temperatures[top[0]]
    assert T > 300.
    assert k * (T - 300.) / H == pytest.approx(-eps * oracle2d.SB * (T ** 4 - Ta ** 4), rel=1e-7)


def test_host_flattening_properties():
    """size-independent properties on random meshes and random node sets (hypothesis): the convection matrix is symmetric positive
    semidefinite with support on flagged nodes only; loads are linear in the flux / in coeff * ambient; an edge needs both nodes"""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=25, deadline=None)
    @given(st.integers(0, 2 ** 31 - 1), st.integers(3, 7), st.integers(3, 7), st.booleans(), st.booleans())
    def check(seed, n0, n1, cyl, verbatim):
        rng = np.random.default_rng(seed)
        x = (0. if cyl else -2.) + np.concatenate([[0.], np.cumsum(rng.uniform(0.2, 2., n0 - 1))])
        y = np.concatenate([[0.], np.cumsum(rng.uniform(0.2, 2., n1 - 1))])
        N = n0 * n1
        nodes = np.nonzero(rng.random(N) < 0.6)[0]
        if nodes.size == 0:
            return
        h, Ta, q = rng.uniform(1., 1e4), rng.uniform(250., 400.), rng.uniform(-1e6, 1e6)
        load, _, _, K = _host_terms(x, y, [(nodes, q)], [(nodes, h, Ta)], [], cyl, verbatim)
        assert np.array_equal(K, K.T)
        off = np.ones(N, dtype=bool)
        off[nodes] = False
        assert not K[off].any() and not K[:, off].any() and not load[off].any()
        if K.any():
            assert np.linalg.eigvalsh(K).min() >= -1e-12 * np.abs(K).max()
        load2, _, _, K2 = _host_terms(x, y, [(nodes, 2. * q)], [(nodes, 3. * h, Ta)], [], cyl, verbatim)
        lq, _, _, _ = _host_terms(x, y, [(nodes, q)], [], [], cyl, verbatim)
        lc = load - lq
        assert np.allclose(K2, 3. * K, rtol=1e-13, atol=0.)
        assert np.allclose(load2, 2. * lq + 3. * lc, rtol=1e-12, atol=1e-12 * (np.abs(load).max() + 1e-300))
        # a node whose neighbours along both axes are all unflagged contributes nothing
        g = np.zeros((n0, n1), dtype=bool)
        g.ravel()[nodes] = True
        pad = np.pad(g, 1)
        lonely = g & ~(pad[:-2, 1:-1] | pad[2:, 1:-1] | pad[1:-1, :-2] | pad[1:-1, 2:])
        assert not load[lonely.ravel()].any() and not K[lonely.ravel()].any()

    check()


def test_host_flattening_equals_oracle_on_random_cases():
    """random meshes, random node sets and node-wise random values for all three kinds (several overlapping conditions each: the first
    one naming a node wins), both geometries, verbatim and corrected: library flattening == oracle restatement"""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=30, deadline=None)
    @given(st.integers(0, 2 ** 31 - 1), st.integers(2, 8), st.integers(2, 8), st.booleans(), st.booleans())
    def check(seed, n0, n1, cyl, verbatim):
        rng = np.random.default_rng(seed)
        x = (0. if cyl else -1.) + np.concatenate([[0.], np.cumsum(rng.uniform(0.1, 3., n0 - 1))])
        y = np.concatenate([[0.], np.cumsum(rng.uniform(0.1, 3., n1 - 1))])
        m = oracle2d.Mesh2D(x, y)
        pick = lambda frac: np.nonzero(rng.random(m.N) < frac)[0]
        heatflux = [(pick(0.5), rng.uniform(-1e6, 1e6)), (pick(0.5), rng.uniform(-1e6, 1e6))]
        convection = [(pick(0.4), rng.uniform(1., 1e5), rng.uniform(250., 350.)), (pick(0.6), rng.uniform(1., 1e5), rng.uniform(250., 350.))]
        radiation = [(pick(0.5), rng.uniform(0.1, 1.), rng.uniform(250., 350.)), (pick(0.3), rng.uniform(0.1, 1.), rng.uniform(250., 350.))]
        T = rng.uniform(250., 500., m.N)
        (kr, kc, kd), F = oracle2d.edge_terms(m, T, heatflux, convection, radiation, cyl=cyl, verbatim=verbatim)
        Ko = np.zeros((m.N, m.N))
        np.add.at(Ko, (kr, kc), kd)
        load, rcoef, ramb4, K = _host_terms(x, y, heatflux, convection, radiation, cyl, verbatim)
        Fh = load - rcoef * (T ** 4 - ramb4)
        assert np.abs(K - Ko).max() <= 1e-13 * max(np.abs(Ko).max(), 1e-300)
        assert np.abs(Fh - F).max() <= 1e-12 * max(np.abs(F).max(), 1e-300)

    check()


def test_oracle_cylindrical_side_wall_analytic():
    """radial conduction through a hollow cylinder: T = T_i on the inner wall (r_i), convection (corrected units) or a heat flux on the
    outer wall (r_o) — the RIGHT-side terms of the cylindrical solver carry elem.getUpper0() (therm2d.cpp:375-377, :385-397):
        convection:  T(r) = T_i - (T_i - T_a) ln(r / r_i) / (ln(r_o / r_i) + k / (h r_o))
        heat flux q leaving the wall:  T(r) = T_i - q r_o ln(r / r_i) / k
    (lengths in metres in the formulas).  Second-order convergence of the 4-node elements: 1e-4 of the drop on 201 radial nodes."""
    ri, ro, k = 2.0, 10.0, 30.
    x = np.linspace(ri, ro, 201)
    y = np.linspace(0., 1., 3)
    n0, n1 = len(x), len(y)
    E = (n0 - 1) * (n1 - 1)
    tab = np.full((1, 2), k)
    ng = np.arange(n0 * n1).reshape(n0, n1)
    inner, outer = ng[0, :], ng[-1, :]
    Ti, Ta, h, q = 400., 300., 2.0e6, 5.0e7
    r = np.repeat(x, n1)

    o = oracle2d.Static2DOracle(x, y, np.zeros(E, dtype=np.int64), 300., 1000., tab, tab, inner, np.full(n1, Ti), cyl=True)
    o.convection, o.verbatim = [(outer, h, Ta)], False
    o.compute(1)
    exact = Ti - (Ti - Ta) * np.log(r / ri) / (np.log(ro / ri) + k / (h * ro * 1e-6))
    assert np.abs(o.temperatures - exact).max() <= 1e-4 * (Ti - exact.min())
    assert Ti - exact.min() > 20.

    o = oracle2d.Static2DOracle(x, y, np.zeros(E, dtype=np.int64), 300., 1000., tab, tab, inner, np.full(n1, Ti), cyl=True)
    o.heatflux = [(outer, q)]
    o.compute(1)
    exact = Ti - q * (ro * 1e-6) * np.log(r / ri) / k
    assert np.abs(o.temperatures - exact).max() <= 1e-4 * (Ti - exact.min())
    assert Ti - exact.min() > 20.


def test_oracle_cylindrical_inner_wall_analytic():
    """the LEFT-side terms carry elem.getLower0() (therm2d.cpp:375): heat flux q leaving through the inner wall of a hollow cylinder whose
    outer wall is held at T_o:  T(r) = T_o + q r_i ln(r / r_o) / k"""
    ri, ro, k, To, q = 2.0, 10.0, 30., 400., 2.0e8
    x = np.linspace(ri, ro, 201)
    y = np.linspace(0., 1., 3)
    n0, n1 = len(x), len(y)
    tab = np.full((1, 2), k)
    ng = np.arange(n0 * n1).reshape(n0, n1)
    o = oracle2d.Static2DOracle(x, y, np.zeros((n0 - 1) * (n1 - 1), dtype=np.int64), 300., 1000., tab, tab, ng[-1, :], np.full(n1, To), cyl=True)
    o.heatflux = [(ng[0, :], q)]
    o.compute(1)
    r = np.repeat(x, n1)
    exact = To + q * (ri * 1e-6) * np.log(r / ro) / k
    assert To - exact.min() > 20.
    assert np.abs(o.temperatures - exact).max() <= 1e-4 * (To - exact.min())
