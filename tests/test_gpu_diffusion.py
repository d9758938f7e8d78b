"""GPU parity of the Diffusion3D path (include/plaskdiff_cuda.h, plask_b200/diffusion.py) through the C ABI:
element integrals against the reference's generated expressions (golden vectors) and the oracle, the assembled operator and load
vector, the whole Newton loop against the oracle's band-Cholesky loop, the reference's own analytic cases
(solvers/electrical/diffusion/tests/diffusion3d.py:86-119) with its tolerances, spatial hole burning, both iteration orders,
interpolation and the error paths.  Concentrations are ~1e19 cm^-3, so every comparison is relative."""
import os

import numpy as np
import pytest

from oracle import diffusion_oracle as orc
from plask_b200 import _lib as L
from plask_b200.diffusion import ActiveRegion, DeviceDiffusion, Diffusion3D, DiffusionProblem
from test_oracle_diffusion import A, B, C, D, GOLD, gaussian_case, quarter_disc, random_problem

pytestmark = pytest.mark.gpu
LD = 4.0


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def device_for(o, par, ms=(), order=L.DIFF_ORDER_01):
    """a device context holding the oracle's problem (full-grid ORDER_01 arrays re-ordered when order = ORDER_10)"""
    d = DeviceDiffusion(0)
    n0, n1 = o.n0, o.n1

    def el(a, k=1):
        a = np.asarray(a).reshape((n0 - 1, n1 - 1) + ((k,) if k > 1 else ()))
        return np.ascontiguousarray(np.swapaxes(a, 0, 1) if order == L.DIFF_ORDER_10 else a).reshape((-1,) + ((k,) if k > 1 else ()))

    def nd(a, k=1):
        a = np.asarray(a).reshape((n0, n1) + ((k,) if k > 1 else ()))
        return np.ascontiguousarray(np.swapaxes(a, 0, 1) if order == L.DIFF_ORDER_10 else a).reshape((-1,) + ((k,) if k > 1 else ()))

    d.set_mesh(o.ax0, o.ax1, order, el(o.active))
    d.set_parameters(el(par["A"]), el(par["B"]), el(par["Cc"]), el(par["D"]))
    d.set_current(nd(par["J"]))
    if ms:
        d.set_modes(np.array([nd(m["P"], 2) for m in ms]), np.array([el(m["G"], 2) for m in ms]), np.array([el(m["dG"], 2) for m in ms]))
    d.el, d.nd = el, nd
    return d


def test_element_integrals_match_reference_golden():
    """one-element meshes: K and F of setLocalMatrix + addLocalBurningMatrix as the reference's generated code gives them"""
    g = np.load(GOLD)
    d = DeviceDiffusion(0)
    worst = 0.
    for i in range(len(g["X"])):
        d.set_mesh([0., g["X"][i]], [0., g["Y"][i]], L.DIFF_ORDER_01, None)
        d.set_parameters(g["A"][i], g["B"][i], g["C"][i], g["D"][i])
        d.set_current(g["J"][i])
        d.set_concentration(g["U"][i])
        K, F = d.element_matrices()
        assert _rel(K[0], g["K"][i]) < 1e-13 and _rel(F[0], g["F"][i]) < 1e-13
        d.set_modes(g["P"][i][None], g["G"][i][None, None], g["dG"][i][None, None])
        Kb, Fb = d.element_matrices(verbatim=True)
        # the golden burning terms were made with the verbatim Ug of the same U
        assert _rel(Kb[0] - K[0], g["Kb"][i]) < 1e-9 and _rel(Fb[0], g["F"][i] + g["Fb"][i]) < 1e-13
        worst = max(worst, _rel(K[0], g["K"][i]))
        d.set_modes()
    print("worst relative deviation of K from the reference's closed form:", worst)
    d.close()


@pytest.mark.parametrize("order", [L.DIFF_ORDER_01, L.DIFF_ORDER_10])
@pytest.mark.parametrize("verbatim", [True, False])
def test_operator_and_load_vector_match_oracle(order, verbatim):
    o, par, ms, rng = random_problem(seed=11, n0=13, n1=10)
    o.U = rng.normal(size=3 * o.nn) * 1e18 * np.repeat(o.node_active, 3)
    o.U[0::3] += 5e18 * o.node_active
    d = device_for(o, par, ms, order)
    d.set_concentration(d.nd(o.U, 3).ravel())
    K, F = o.assemble(par["A"], par["B"], par["Cc"], par["D"], par["J"], ms, verbatim)
    Fd = d.rhs(verbatim)
    assert _rel(Fd, d.nd(F, 3).ravel()) < 1e-13
    v = rng.normal(size=3 * o.nn) * np.repeat(o.node_active, 3)
    yd = d.apply(d.nd(v, 3).ravel(), verbatim)
    assert _rel(yd, d.nd(K @ v, 3).ravel()) < 1e-13
    # rows of nodes outside the masked mesh are unit rows
    w = rng.normal(size=3 * o.nn)
    yw = d.apply(d.nd(w, 3).ravel(), verbatim).reshape(-1, 3)
    off = ~d.nd(o.node_active).astype(bool)
    assert np.array_equal(yw[off], d.nd(w, 3)[off])
    Ke, Fe = d.element_matrices(verbatim)
    assert np.all(Ke[~d.el(o.active).astype(bool)] == 0.)
    assert np.abs(Ke - Ke.transpose(0, 2, 1)).max() <= 1e-14 * np.abs(Ke).max()
    d.close()


@pytest.mark.parametrize("order", [L.DIFF_ORDER_01, L.DIFF_ORDER_10])
@pytest.mark.parametrize("shb", [False, True])
def test_newton_loop_matches_oracle(order, shb):
    """compute(): same loop count, same err of every loop, same U as the band-Cholesky loop of the oracle"""
    o, par, ms, rng = random_problem(seed=5, n0=21, n1=17, modes=2)
    ms = ms if shb else []
    for m in ms:                       # keep the linearised matrix positive definite: dG >= 0, moderate burning
        m["G"] = m["G"] * 1e-3
    loops = o.compute(par["A"], par["B"], par["Cc"], par["D"], par["J"], maxerr=1e-6, modes=ms)
    d = device_for(o, par, ms, order)
    st = d.compute(0, 1e-6, lin_tol=1e-13)
    assert st["status"] == 0 and st["converged"] and st["loops"] == loops
    assert st["kernel_launches"] == 1
    for got, want in zip(st["err_log"], o.history):
        if want > 1e-3:                # below that the error of a loop is set by the accuracy of the previous linear solve
            assert abs(got - want) <= 1e-6 * want
    U = d.get_concentration()
    assert _rel(U, d.nd(o.U, 3).ravel()) < 1e-9
    # loops = k stops after k residual evaluations, i.e. k - 1 solves (diffusion3d.cpp:354)
    d.set_concentration(None)
    st3 = d.compute(3, 1e-6, lin_tol=1e-13)
    o3 = orc.Diffusion3DOracle(o.ax0, o.ax1, o.active.reshape(o.n0 - 1, o.n1 - 1))
    o3.compute(par["A"], par["B"], par["Cc"], par["D"], par["J"], loops=3, maxerr=1e-6, modes=ms)
    assert st3["loops"] == 3 and not st3["converged"]
    assert _rel(d.get_concentration(), d.nd(o3.U, 3).ravel()) < 1e-9
    d.close()


def reference_test_solver(n=101):
    """the structure of diffusion3d.py:56-84: quarter of a cylinder of radius L, three 2 nm wells, regular mesh of spacing 0.01 L"""
    ax, mask = quarter_disc(n)
    reg = ActiveRegion(mask, [(0.100, 0.102), (0.103, 0.105), (0.106, 0.108)], A=A, B=B, C=C, D=D)
    s = Diffusion3D("diffusion3d")
    s.problem = DiffusionProblem(ax, ax, [reg])
    s.maxerr = 0.0001
    return s


TEST_POINTS = np.array([(abs(x), abs(y), 0.104) for x in np.linspace(-0.8 * LD, 0.8 * LD, 5) for y in np.linspace(-0.8 * LD, 0.8 * LD, 5)])


def test_reference_uniform_case():
    """diffusion3d.py:86-94 (test_uniform), rtol 1e-5"""
    s = reference_test_solver()
    n = 1.0e19
    j = 1e-7 * (A * n + B * n**2 + C * n**3) * (orc.QE * 0.006)
    s.inCurrentDensity = np.array([0., 0., j])
    s.compute()
    res = s.outCarriersConcentration(TEST_POINTS)
    ref = n * ((TEST_POINTS[:, :2]**2).sum(1) <= LD * LD)
    np.testing.assert_allclose(res, ref, rtol=1e-5)
    assert s.stats["converged"] and s.stats["err"] < 1e-4
    # no concentration outside the wells (ConcentrationDataImpl::at, diffusion3d.cpp:462-484)
    assert s.outCarriersConcentration([[0.5, 0.5, 0.1025]])[0] == 0.
    assert s.outCarriersConcentration([[0.5, 0.5, 0.1035]])[0] > 0.


def test_reference_gaussian_case():
    """diffusion3d.py:96-119 (test_gaussian), rtol 0.5e-3; and the oracle on the same mesh to 1e-8"""
    s = reference_test_solver()

    def n_of(p):
        return 1e19 * (np.exp(-p[:, 0]**2 - p[:, 1]**2) + 0.5) * ((p[:, :2]**2).sum(1) <= LD * LD)

    def j_of(p):
        x, y = p[:, 0], p[:, 1]
        n = 1e19 * (np.exp(-x**2 - y**2) + 0.5)
        lap = 2e19 * (2 * x**2 - 1) * np.exp(-x**2 - y**2) + 2e19 * (2 * y**2 - 1) * np.exp(-x**2 - y**2)
        nj = 1e8 * D * lap - A * n - B * n**2 - C * n**3
        return np.stack([0 * x, 0 * x, -1e-7 * (orc.QE * 0.006) * nj], axis=1)

    s.inCurrentDensity = j_of
    s.compute()
    res = s.outCarriersConcentration(TEST_POINTS)
    np.testing.assert_allclose(res, n_of(TEST_POINTS), rtol=0.5e-3)
    ax, mask = quarter_disc(101)
    o = orc.Diffusion3DOracle(ax, ax, mask)
    o.compute(A, B, C, 1e8 * D, gaussian_case(ax), maxerr=1e-4)
    assert s.stats["loops"] == len(o.history)
    assert _rel(s._dev[0].get_concentration(), o.U) < 1e-8
    lin = s.outCarriersConcentration(TEST_POINTS, "linear")
    np.testing.assert_allclose(lin, n_of(TEST_POINTS), rtol=2e-3)


def test_interpolation_matches_oracle():
    o, par, ms, rng = random_problem(seed=8, n0=12, n1=9, modes=0)
    o.U = rng.normal(size=3 * o.nn) * np.repeat(o.node_active, 3)
    for order in (L.DIFF_ORDER_01, L.DIFF_ORDER_10):
        d = device_for(o, par, (), order)
        d.set_concentration(d.nd(o.U, 3).ravel())
        x = rng.uniform(o.ax0[0] - 0.05, o.ax0[-1] + 0.05, 400)
        y = rng.uniform(o.ax1[0] - 0.05, o.ax1[-1] + 0.05, 400)
        got = d.interpolate(x, y)
        inside = (x >= o.ax0[0]) & (x <= o.ax0[-1]) & (y >= o.ax1[0]) & (y <= o.ax1[-1])
        want = np.where(inside, o.concentration(np.clip(x, o.ax0[0], o.ax0[-1]), np.clip(y, o.ax1[0], o.ax1[-1])), 0.)
        assert np.abs(got - want).max() < 1e-13
        assert np.all(got[~inside] == 0.)
        # at the nodes the interpolant is the nodal value
        X, Y = np.meshgrid(o.ax0, o.ax1, indexing="ij")
        at_nodes = d.interpolate(X.ravel(), Y.ravel())
        touch = o.node_active & (o.concentration(X.ravel(), Y.ravel()) != 0)
        assert np.abs(at_nodes[touch] - o.U[0::3][touch]).max() < 1e-13
        d.close()


def test_shb_through_the_mirror():
    """compute(shb=True): the mirror builds P, G, dG like diffusion3d.cpp:250-271,289-294; against the oracle fed the same numbers"""
    from plask_b200 import diffusion as m
    n = 33
    ax = np.linspace(0., 4., n)
    mask = np.ones((n - 1, n - 1), dtype=bool)
    reg = ActiveRegion(mask, [(0.100, 0.102), (0.103, 0.105), (0.106, 0.108)], A=A, B=B, C=C, D=D, nr=3.5)
    s = Diffusion3D("shb")
    s.problem = DiffusionProblem(ax, ax, [reg])
    s.maxerr = 1e-5
    s.inCurrentDensity = np.array([0., 0., 8.])
    s.inWavelength = [980.]
    s.inLightE = [lambda p: np.stack([4e6 * np.exp(-(p[:, 0]**2 + p[:, 1]**2) / 2.) + 0j, 0 * p[:, 0] + 0j, 0 * p[:, 0] + 0j], axis=1)]
    s.inGain = lambda p, wl, deriv: (np.full((len(p), 2), 2e-16) if deriv else np.full((len(p), 2), 1500.))
    s.compute(shb=True)
    assert s.stats["converged"]
    U_shb = s._dev[0].get_concentration()
    # the same numbers through the oracle
    pts = s.problem.node_points(reg.vert)
    E = s.inLightE[0](pts)
    P = np.stack([(0.5 / m.Z0) * np.abs(E[:, 0])**2, 0 * pts[:, 0]], axis=1)
    G = np.full((s._dev[0].ne, 2), m.INV_HC * 980. * 3.5 * 1500.)
    dG = np.full((s._dev[0].ne, 2), m.INV_HC * 980. * 3.5 * 2e-16)
    o = orc.Diffusion3DOracle(ax, ax, mask)
    J = 1e7 / (m.QE * reg.qw_height) * 8.
    o.compute(A, B, C, 1e8 * D, J, maxerr=1e-5, modes=[dict(P=P, G=G, dG=dG)])
    assert s.stats["loops"] == len(o.history)
    assert _rel(U_shb, o.U) < 1e-8
    # the light burns a hole at the centre
    s2 = Diffusion3D("noshb")
    s2.problem = DiffusionProblem(ax, ax, [reg])
    s2.maxerr = 1e-5
    s2.inCurrentDensity = np.array([0., 0., 8.])
    s2.compute()
    assert s.outCarriersConcentration([[0., 0., 0.104]])[0] < 0.98 * s2.outCarriersConcentration([[0., 0., 0.104]])[0]
    want = orc.burned_power(ax, ax, mask, P, np.full((s._dev[0].ne, 2), 3.5 * 1500.), reg.qw_height, True)
    assert abs(s.get_total_burning() - want) <= 1e-12 * abs(want) and s.get_burning_for_mode(0) == s.get_total_burning()


def test_error_paths():
    d = DeviceDiffusion(0)
    with pytest.raises(L.BadInput):
        d.compute()                                   # no mesh
    with pytest.raises(L.BadInput):
        d.set_mesh([0., 1., 1.], [0., 1.], 0, None)   # axis not increasing
    d.set_mesh(np.linspace(0, 1, 6), np.linspace(0, 1, 5), 0, None)
    with pytest.raises(L.BadInput):
        d.compute()                                   # no parameters
    d.set_parameters(-1e20, 0., 0., 1.)               # negative recombination: not positive definite
    d.set_current(1e20)
    with pytest.raises(L.ComputationError, match="positive definite"):
        d.compute()
    s = Diffusion3D("x")
    s.algorithm = "cholesky"
    with pytest.raises(L.BadInput):
        s.compute()
    d.close()


def test_larger_mesh_converges_and_reports_residual():
    """401 x 401 lateral nodes (482 k unknowns): the whole loop is one kernel launch; the residual of the last linear solve is small"""
    n = 401
    ax = np.linspace(0., 20., n)
    xm = 0.5 * (ax[1:] + ax[:-1])
    mask = xm[:, None]**2 + xm[None, :]**2 <= 20.**2
    d = DeviceDiffusion(0)
    d.set_mesh(ax, ax, 0, mask.ravel())
    d.set_parameters(A, B, C, 1e8 * D)
    X, Y = np.meshgrid(ax, ax, indexing="ij")
    J = 3e27 * (1. + 4. * np.exp(-(X**2 + Y**2) / 16.))
    d.set_current(J.ravel())
    st = d.compute(0, 1e-4)
    assert st["status"] == 0 and st["converged"] and st["lin_relres_precond"] <= 1e-12 and st["lin_relres"] <= 1e-9 and st["kernel_launches"] == 1
    U = d.get_concentration().reshape(n, n, 3)
    unknown = orc.Diffusion3DOracle(ax, ax, mask).node_active.reshape(n, n)
    assert U[0, 0, 0] > U[200, 0, 0] > 0. and np.all(U[~unknown] == 0.) and np.all(U[unknown][:, 0] > 0.)
    print(f"401x401: {st['loops']} loops, {st['lin_iters']} PCG iterations, {st['t_solve_ms']:.1f} ms on the device")
    d.close()
