"""Pins of the 2-D oracle (oracle/oracle2d.py, test infrastructure): the reference's own Shockley2D / ShockleyCyl tests
(solvers/electrical/shockley/tests/shockley2d.py:57-65,105-118: analytic current, capacitance, heat; temperature-dependent beta) with
the reference's assertAlmostEqual places, and — the reference holds no numeric Static2D / StaticCyl fixture — analytic solutions
of the heat equation in a slab and in a cylinder."""
import numpy as np
import pytest

from helpers import oracle_shockley2d, oracle_static2d, shockley2d_reference_problem
from plask_b200.solvers2d import Problem2D, embed

EPS0_PF_UM = 8.854187817e-6   # pF/um, shockley2d.py:21


@pytest.mark.parametrize("cyl", [False, True])
def test_shockley2d_py_testComputations(cyl):
    p2 = shockley2d_reference_problem(cyl)
    o = oracle_shockley2d(p2)
    o.compute(1000 if cyl else 0)
    geo = np.pi if cyl else 1.
    current = 1e-3 * geo * p2.js * (np.exp(p2.beta) - 1.)
    assert abs(o.get_total_current() - current) < 0.5e-3                       # assertAlmostEqual(..., 3)
    assert abs(o.get_capacitance() - EPS0_PF_UM * 12.9 * geo * 1000. ** 2 / 0.02) < 0.5e-2    # places = 2
    assert abs(o.get_total_heat() - current * 1.) < 0.5e-3


def test_shockleycyl_py_testComputationsTemp():
    """beta = log(70 T): at 300 K exp(beta U) = 21000, at 250 K 17500 (shockley2d.py:110-118)"""
    p2 = shockley2d_reference_problem(True)
    o = oracle_shockley2d(p2, beta=lambda T: np.log(T * 70.))
    o.compute(1000)
    assert abs(o.get_total_current() - 1e-3 * np.pi * (21000. - 1.)) < 0.5e-3
    o = oracle_shockley2d(p2, beta=lambda T: np.log(T * 70.), Te=250.)
    o.compute(1000)
    assert abs(o.get_total_current() - 1e-3 * np.pi * (17500. - 1.)) < 0.5e-3


def test_shockley2d_py_testConductivity():
    p2 = shockley2d_reference_problem(False)
    o = oracle_shockley2d(p2)
    o.load_conductivities()
    expect = np.where(np.asarray(p2.elem_junc)[:, None] > 0, np.array([[0., 5.]]), p2.tab_lat[p2.elem_mat, 0][:, None] * np.ones((1, 2)))
    assert np.array_equal(o.conds, expect)


def _uniform(n, cyl, k=44., q=3e15, R=20., H=10.):
    x, y = np.linspace(0., R, n[0]), np.linspace(0., H, n[1])
    tab = np.full((1, 2), k)
    ng = np.arange(n[0] * n[1]).reshape(n)
    return x, y, tab, ng


def test_static2d_slab_parabola_is_nodally_exact():
    """uniform k and heat, bottom at T0, everything else insulated: T(y) = T0 + q (2 H y - y^2) / (2 k), exact for bilinear elements"""
    n, k, q, H = (5, 33), 44., 3e15, 10.
    x, y, tab, ng = _uniform(n, False, k, q, H=H)
    p2 = Problem2D("slab", "thermal", x, y, np.zeros((n[0] - 1) * (n[1] - 1), np.uint32), 250., 100., tab, tab.copy(),
                   ng[:, 0], np.full(n[0], 300.), heat=np.full((n[0] - 1) * (n[1] - 1), q))
    for cyl in (False, True):           # the radial weight cancels for a field that does not depend on r
        p2.cyl = cyl
        o = oracle_static2d(p2)
        o.compute(0)
        ym = y * 1e-6
        exact = 300. + q * (2. * H * 1e-6 * ym - ym * ym) / (2. * k)
        assert np.abs(o.temperatures.reshape(n) - exact[None, :]).max() < 1e-9 * exact.max()


def test_staticcyl_radial_conduction_converges_second_order():
    """uniform heat in a cylinder with the lateral surface at T0: T(r) = T0 + q (R^2 - r^2) / (4 k); bilinear elements with the
    midpoint-radius weight of therm2d.cpp:353 converge with h^2"""
    k, q, R = 44., 3e15, 20.
    errs = []
    for nr in (9, 17, 33):
        n = (nr, 4)
        x, y, tab, ng = _uniform(n, True, k, q, R=R, H=3.)
        p2 = Problem2D("cyl", "thermal", x, y, np.zeros((n[0] - 1) * (n[1] - 1), np.uint32), 250., 100., tab, tab.copy(),
                       ng[-1, :], np.full(n[1], 300.), heat=np.full((n[0] - 1) * (n[1] - 1), q), cyl=True)
        o = oracle_static2d(p2)
        o.compute(0)
        r = x * 1e-6
        exact = 300. + q * ((R * 1e-6) ** 2 - r * r) / (4. * k)
        errs.append(np.abs(o.temperatures.reshape(n) - exact[:, None]).max())
    assert errs[0] / errs[1] > 3. and errs[1] / errs[2] > 3.
    assert errs[2] < 4e-3 * (exact.max() - 300.)      # the axis element (r_mid = h/2) limits the rate to about h^2 log h


def test_embedding_is_consistent():
    """host logic of the product's embedding: plane 0 of the brick mesh is numbered like the 2-D mesh"""
    p2 = shockley2d_reference_problem(True)
    p = embed(p2)
    assert p.n == (2,) + p2.n and p.E == p2.E and p.N == 2 * p2.N
    ng = np.broadcast_to(p.node_index_grid(), p.n)
    assert np.array_equal(ng[0].ravel(), np.arange(p2.N)) and np.array_equal(ng[1].ravel(), np.arange(p2.N) + p2.N)
    assert np.array_equal(np.broadcast_to(p.elem_index_grid(), (1,) + tuple(k - 1 for k in p2.n)).ravel(), np.arange(p2.E))
    assert set(p.bc_nodes.tolist()) == set(p2.bc_nodes.tolist()) | set((np.asarray(p2.bc_nodes) + p2.N).tolist())
