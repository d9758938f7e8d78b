"""Pins of the Diffusion3D oracle (oracle/diffusion_oracle.py) — CPU only.

1. element integrals against the reference's own generated expressions (committed golden vectors made by
   tests/golden/make_golden_diffusion.py from oracle/_ref/libdiffusion_ref.so; compared live too where that library exists);
2. the whole solver against the two analytic cases of the reference's own test
   solvers/electrical/diffusion/tests/diffusion3d.py:86-119 with the reference's tolerances;
3. internal consistency (vectorised assembly, burned power of the mirror)."""
import os

import numpy as np
import pytest

from oracle import diffusion_oracle as d

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "diffusion_elements.npz")

A, B, C, D = 3e7, 1.7e-10, 6e-27, 10.     # diffusion3d.py:30-33
L = 4.0


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_element_integrals_match_reference_golden(gold):
    g = gold
    for i in range(len(g["X"])):
        K, F = d.local_matrix(g["X"][i], g["Y"][i], g["A"][i], g["B"][i], g["C"][i], g["D"][i], g["U"][i], g["J"][i])
        assert _rel(K, g["K"][i]) < 2e-14 and _rel(F, g["F"][i]) < 2e-14
        assert np.abs(K - K.T).max() <= 1e-15 * np.abs(K).max()
        Kb, Fb = d.local_burning(g["X"][i], g["Y"][i], g["G"][i], g["dG"][i], g["Ug"][i], g["P"][i])
        assert _rel(Kb, g["Kb"][i]) < 2e-14 and _rel(Fb, g["Fb"][i]) < 2e-14


@pytest.mark.skipif(not d.ref_available(), reason="oracle/_ref/libdiffusion_ref.so not built (needs /root/reference)")
def test_golden_is_what_the_reference_library_gives(gold):
    g = gold
    for i in (0, 5, 17, 40):
        K, F = d.ref_local_matrix(g["X"][i], g["Y"][i], g["A"][i], g["B"][i], g["C"][i], g["D"][i], g["U"][i], g["J"][i])
        assert np.array_equal(K, g["K"][i]) and np.array_equal(F, g["F"][i])
        Kb, Fb = d.ref_local_burning(g["X"][i], g["Y"][i], g["G"][i], g["dG"][i], g["Ug"][i], g["P"][i])
        assert np.array_equal(Kb, g["Kb"][i]) and np.array_equal(Fb, g["Fb"][i])


def test_element_centre_value_verbatim_and_corrected():
    rng = np.random.default_rng(0)
    U = rng.normal(size=12)
    # on a square element both forms are the Hermite interpolant at the centre
    phi, _, _ = d.basis(0.3, 0.3, np.array([0.5]), np.array([0.5]))
    assert abs(d.element_center(0.3, 0.3, U, True) - U @ phi[:, 0, 0]) < 1e-14
    phi, _, _ = d.basis(0.3, 0.7, np.array([0.5]), np.array([0.5]))
    assert abs(d.element_center(0.3, 0.7, U, False) - U @ phi[:, 0, 0]) < 1e-14
    assert abs(d.element_center(0.3, 0.7, U, True) - U @ phi[:, 0, 0]) > 1e-3


def random_problem(seed=3, n0=9, n1=7, modes=2, holes=0.2):
    rng = np.random.default_rng(seed)
    ax0, ax1 = np.cumsum(rng.uniform(.05, .2, n0)), np.cumsum(rng.uniform(.05, .2, n1))
    act = rng.uniform(size=(n0 - 1, n1 - 1)) > holes
    o = d.Diffusion3DOracle(ax0, ax1, act)
    ne, nn = o.ne, o.nn
    par = dict(A=3e7 * rng.uniform(.5, 2, ne), B=1.7e-10 * rng.uniform(.5, 2, ne), Cc=6e-27 * rng.uniform(.5, 2, ne),
               D=1e9 * rng.uniform(.5, 2, ne), J=rng.uniform(0.2, 1, nn) * 1e30)
    ms = [dict(P=rng.uniform(0, 1, (nn, 2)), G=rng.uniform(0, 1e28, (ne, 2)), dG=rng.uniform(0, 1e9, (ne, 2))) for _ in range(modes)]
    return o, par, ms, rng


def test_vectorised_assembly_equals_element_loop():
    o, par, ms, rng = random_problem()
    o.U = rng.normal(size=3 * o.nn) * 1e18 * np.repeat(o.node_active, 3)
    for vb in (True, False):
        K1, F1 = o.assemble_slow(par["A"], par["B"], par["Cc"], par["D"], par["J"], ms, vb)
        K2, F2 = o.assemble(par["A"], par["B"], par["Cc"], par["D"], par["J"], ms, vb)
        assert abs(K1 - K2).max() < 1e-14 * abs(K1).max() and np.abs(F1 - F2).max() < 1e-14 * np.abs(F1).max()


def quarter_disc(n):
    ax = np.linspace(0, L, n)
    xm = 0.5 * (ax[1:] + ax[:-1])
    return ax, (xm[:, None]**2 + xm[None, :]**2 <= L**2)


def test_uniform_case_of_the_reference_test():
    """diffusion3d.py:86-94: constant current -> uniform concentration, rtol 1e-5 (here on a coarser mesh: the case is exact)"""
    ax, act = quarter_disc(41)
    o = d.Diffusion3DOracle(ax, ax, act)
    n0 = 1.0e19
    J = A * n0 + B * n0**2 + C * n0**3          # = js * j of diffusion3d.py:88
    o.compute(A, B, C, 1e8 * D, J, maxerr=1e-4)
    pts = np.array([(x, y) for x in (0., 1.6, 3.2) for y in (0., 1.6, 3.2)])
    res = o.concentration(pts[:, 0], pts[:, 1])
    ref = n0 * ((pts**2).sum(1) <= L * L)
    np.testing.assert_allclose(res, ref, rtol=1e-5)
    assert o.history[0] == 100. and o.history[-1] < 1e-4


def gaussian_case(ax):
    X, Y = np.meshgrid(ax, ax, indexing="ij")
    nn = 1e19 * (np.exp(-X**2 - Y**2) + 0.5)
    lap = 2e19 * (2 * X**2 - 1) * np.exp(-X**2 - Y**2) + 2e19 * (2 * Y**2 - 1) * np.exp(-X**2 - Y**2)
    return np.abs(-(1e8 * D * lap - A * nn - B * nn**2 - C * nn**3)).ravel()     # = |js * j| of diffusion3d.py:107-111


def test_gaussian_case_of_the_reference_test():
    """diffusion3d.py:113-119 on the reference's own mesh (spacing 0.01 L): rtol 0.5e-3"""
    ax, act = quarter_disc(101)
    o = d.Diffusion3DOracle(ax, ax, act)
    o.compute(A, B, C, 1e8 * D, gaussian_case(ax), maxerr=1e-4)
    pts = np.array([(x, y) for x in (0., 1.6, 3.2) for y in (0., 1.6, 3.2)])
    res = o.concentration(pts[:, 0], pts[:, 1])
    ref = 1e19 * (np.exp(-(pts**2).sum(1)) + 0.5) * ((pts**2).sum(1) <= L * L)
    np.testing.assert_allclose(res, ref, rtol=0.5e-3)


def test_mirror_burned_power_equals_oracle():
    from plask_b200 import diffusion as m
    rng = np.random.default_rng(5)
    ax0, ax1 = np.cumsum(rng.uniform(.1, .2, 7)), np.cumsum(rng.uniform(.1, .2, 6))
    mask = rng.uniform(size=(6, 5)) > 0.25
    reg = m.ActiveRegion(mask, [(0.100, 0.102), (0.103, 0.105), (0.106, 0.108)])
    assert abs(reg.qw_height - 0.006) < 1e-15 and abs(reg.vert - 0.104) < 1e-15
    p = m.DiffusionProblem(ax0, ax1, [reg])
    P, g = rng.uniform(size=(42, 2)), rng.uniform(size=(30, 2))
    for vb in (True, False):
        assert abs(m.burned_power(p, reg, P, g, vb) - d.burned_power(ax0, ax1, mask, P, g, reg.qw_height, vb)) < 1e-28
    # the other iteration order names the same nodes and elements differently
    p10 = m.DiffusionProblem(ax0, ax1, [reg], order=m.L.DIFF_ORDER_10)
    P10 = P.reshape(7, 6, 2).transpose(1, 0, 2).reshape(-1, 2)
    g10 = g.reshape(6, 5, 2).transpose(1, 0, 2).reshape(-1, 2)
    assert abs(m.burned_power(p10, reg, P10, g10, False) - m.burned_power(p, reg, P, g, False)) < 1e-28
