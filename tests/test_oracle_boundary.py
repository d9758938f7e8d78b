"""CPU checks of the oracle's boundary conditions of the 2nd / 3rd kind and radiation (SURVEY.md §8a row a7;
setBoundaries therm3d.cpp:140-168, used in setMatrix :242-268).  The reference has no test for them, so the
restatement is pinned by what the formulas must give: 1-D analytic solutions in the corrected form, equality of the
verbatim and the corrected form where the reference's local-slot accumulation is harmless (z-low sides), and
agreement of Cholesky with the reference's NSPCG on the same assembled system."""
import numpy as np
import pytest

from helpers import oracle_mesh, slab_problem_1d, face_nodes
from oracle import oracle as orc

SB = 5.670373e-8


def _oracle(p, boundaries, quirk, **kw):
    m = oracle_mesh(p)
    tb = orc.Tables(p.T0, p.dT, p.tab_lat, p.tab_vert)
    kw.setdefault("algorithm", "cholesky")
    return orc.Static3DOracle(m, p.elem_mat, tb, p.bc_nodes, p.bc_values, heat=p.heat, inittemp=p.inittemp,
                              maxerr=1e-9, boundaries=boundaries, quirk=quirk, **kw)


def test_convection_top_analytic_corrected():
    """k T' = -h (T(H) - Ta) at the top, T(0) = T0, no source: T linear, T(H) = (k T0/H + h Ta)/(k/H + h)."""
    k, h, Ta, T0 = 40., 2.0e5, 350., 300.
    p = slab_problem_1d(n=(4, 5, 17), H=12., k=k, T0=T0)
    top = face_nodes(p, 2, -1)
    b = orc.BoundaryTerms(p.N, convection=[(top, h, Ta)])
    s = _oracle(p, b, quirk=False)
    s.compute(1)
    z = p.axes[2] * 1e-6
    TH = (k * T0 / z[-1] + h * Ta) / (k / z[-1] + h)
    T = np.broadcast_to(T0 + (TH - T0) * z / z[-1], p.n).ravel()
    got = s.temperatures[np.broadcast_to(p.node_index_grid(), p.n).ravel()]
    assert np.abs(got - T).max() < 1e-9 * 350.


def test_heatflux_top_analytic_corrected():
    """heat flux value q [W/m2] is the flux LEAVING through the side (F = -0.25e-12 A q, therm3d.cpp:244):
    T(z) = T0 - q z / k."""
    k, q, T0 = 40., 3.0e6, 300.
    p = slab_problem_1d(n=(3, 4, 13), H=8., k=k, T0=T0)
    b = orc.BoundaryTerms(p.N, heatflux=[(face_nodes(p, 2, -1), q)])
    s = _oracle(p, b, quirk=False)
    s.compute(1)
    z = p.axes[2] * 1e-6
    T = np.broadcast_to(T0 - q * z / k, p.n).ravel()
    got = s.temperatures[np.broadcast_to(p.node_index_grid(), p.n).ravel()]
    assert np.abs(got - T).max() < 1e-9 * 300.


def test_radiation_top_analytic_corrected():
    """-k T' = eps SB (T(H)^4 - Ta^4): the converged loop satisfies the flux balance."""
    k, eps, Ta, T0 = 0.05, 0.9, 250., 400.
    p = slab_problem_1d(n=(3, 3, 21), H=500., k=k, T0=T0)
    b = orc.BoundaryTerms(p.N, radiation=[(face_nodes(p, 2, -1), eps, Ta)])
    s = _oracle(p, b, quirk=False)
    s.maxerr = 1e-7
    # the explicit (lagged) radiation term of the reference converges only where eps SB 4 T^3 H / k < 1
    s.compute(200)
    ng = np.broadcast_to(p.node_index_grid(), p.n)
    TH = s.temperatures[ng[0, 0, -1]]
    z = p.axes[2] * 1e-6
    flux_cond = k * (T0 - TH) / z[-1]
    flux_rad = eps * SB * (TH ** 4 - Ta ** 4)
    assert abs(flux_cond - flux_rad) < 1e-6 * abs(flux_rad)


@pytest.mark.parametrize("order", ["012", "201"])
def test_verbatim_equals_corrected_on_z_low_sides(order):
    """a condition on the bottom plane only lives on the z-low side {0,1,2,3} of the bottom elements, where the
    local slots of setBoundaries ARE the wall nodes (therm3d.cpp:150,157)"""
    p = slab_problem_1d(n=(5, 4, 9), H=6., k=30., T0=300., order=order, dirichlet="top")
    bot = face_nodes(p, 2, 0)
    b = orc.BoundaryTerms(p.N, heatflux=[(bot, -2.0e6)], radiation=[(bot, 0.7, 280.)])
    a, c = _oracle(p, b, quirk=True), _oracle(p, b, quirk=False)
    a.temperatures[:8] = 300.   # verbatim radiation reads temperatures[0..7] (therm3d.cpp:265); uniform field: same value
    a.compute(1)
    c.compute(1)
    assert np.array_equal(a.temperatures, c.temperatures)
    assert a.temperatures.max() > 300.1


def test_verbatim_convection_factor_on_z_low_side():
    """the verbatim matrix term is a quarter of the consistent one (therm3d.cpp:255): on the bottom plane (slots =
    wall nodes) the 1-D solution obeys  k T'(0) = (h/4) T(0) - h Ta  instead of  h (T(0) - Ta)"""
    k, h, Ta, T0 = 40., 2.0e5, 350., 300.
    p = slab_problem_1d(n=(4, 3, 15), H=12., k=k, T0=T0, dirichlet="top")
    b = orc.BoundaryTerms(p.N, convection=[(face_nodes(p, 2, 0), h, Ta)])
    z = p.axes[2] * 1e-6
    H = z[-1]
    ng = np.broadcast_to(p.node_index_grid(), p.n).ravel()
    for quirk, hk in ((True, h / 4.), (False, h)):
        s = _oracle(p, b, quirk=quirk)
        s.compute(1)
        Tb = (k * T0 / H + h * Ta) / (k / H + hk)      # k (T0 - Tb)/H = hk Tb - h Ta
        T = np.broadcast_to(Tb + (T0 - Tb) * z / H, p.n).ravel()
        assert np.abs(s.temperatures[ng] - T).max() < 1e-9 * 350., quirk


def test_verbatim_differs_on_other_sides():
    """...and on any other side the verbatim form puts the terms on the element's z-low nodes: documented quirk"""
    p = slab_problem_1d(n=(4, 4, 9), H=6., k=30., T0=300.)
    b = orc.BoundaryTerms(p.N, convection=[(face_nodes(p, 2, -1), 1.0e5, 350.)])
    a, c = _oracle(p, b, quirk=True), _oracle(p, b, quirk=False)
    a.compute(1)
    c.compute(1)
    assert np.abs(a.temperatures - c.temperatures).max() > 1e-3


@pytest.mark.parametrize("quirk", [True, False])
def test_cholesky_vs_reference_nspcg_with_boundaries(quirk):
    from plask_b200 import configs as cf
    p = cf.config_A(10)
    b = orc.BoundaryTerms(p.N, convection=[(face_nodes(p, 2, -1), 5.0e4, 320.)], heatflux=[(face_nodes(p, 0, 0), 1.0e5)],
                          radiation=[(face_nodes(p, 1, -1), 0.8, 290.)])
    a = _oracle(p, b, quirk)
    a.maxerr = 1e-6
    a.compute(0)
    c = _oracle(p, b, quirk, algorithm="iterative", precond="ic", itmaxerr=1e-12, maxit=5000)
    c.maxerr = 1e-6
    c.compute(0)
    assert len(a.history) == len(c.history)
    assert np.abs(a.temperatures - c.temperatures).max() < 1e-6
