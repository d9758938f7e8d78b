"""GPU parity of the Static3D nonlinear solve (therm3d.cpp:281-340) against the oracle's
Cholesky (LAPACK dpbtrf/dpbtrs) and the reference's own NSPCG, through the solver mirror.
Tolerance of the north star: max |dT| <= 1e-3 K, relative residual <= 1e-8."""
import numpy as np
import pytest

from helpers import face_nodes as face_nodes_, oracle_thermal, random_problem
from oracle import oracle as orc
from plask_b200 import configs as cf
from plask_b200.solvers import Static3D

pytestmark = pytest.mark.gpu

TOL_T = 1e-3  # K, BASELINE.json north_star


def gpu_solve(p, variant=3, loops=0, lin_tol=1e-10):
    s = Static3D("thermal")
    s.problem = p
    s.inittemp, s.maxerr = p.inittemp, p.maxerr
    s.variant = variant
    s.iterative.maxerr = lin_tol
    s.iterative.maxit = 20000
    err = s.compute(loops)
    return s, err


@pytest.mark.parametrize("variant", [1, 2, 0, 3])
def test_config_A_small_vs_cholesky(variant):
    p = cf.config_A(20)
    o = oracle_thermal(p, algorithm="cholesky")
    o.compute(0)
    s, err = gpu_solve(p, variant)
    T = s.outTemperature()
    assert np.abs(T - o.temperatures).max() <= TOL_T
    assert np.abs(T - o.temperatures).max() <= 1e-6          # in fact much tighter
    assert s.stats["outer_loops"] == len(o.history)
    assert s.iterative.converged and s.iterative.err <= 1e-8
    assert abs(s.maxT - o.maxT) <= TOL_T
    s.invalidate()


@pytest.mark.parametrize("order", ["012", "021", "102", "120", "201", "210"])
def test_config_B_small_all_orders_vs_cholesky(order):
    p = cf.config_B((18, 20, 44), order=order)
    o = oracle_thermal(p, algorithm="cholesky")
    o.compute(0)
    s, err = gpu_solve(p)
    T = s.outTemperature()
    assert np.abs(T - o.temperatures).max() <= TOL_T, np.abs(T - o.temperatures).max()
    assert s.stats["outer_loops"] == len(o.history)
    assert err == pytest.approx(max(h["err"] for h in o.history), abs=TOL_T)
    s.invalidate()


def test_vs_reference_nspcg():
    """the reference's own iterative path (NSPCG cg + ic, tightened maxerr)"""
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built")
    p = cf.config_B(24)
    o = oracle_thermal(p, algorithm="iterative", precond="ic", itmaxerr=1e-10, maxit=5000)
    o.compute(0)
    s, _ = gpu_solve(p)
    assert np.abs(s.outTemperature() - o.temperatures).max() <= TOL_T
    s.invalidate()


@pytest.mark.parametrize("precond", ["jac", "ljac"])
def test_config_A_full_size_vs_reference_nspcg(precond):
    """BASELINE configs[0] at its full size (64^3 GaAs/AlGaAs stack, uniform heat): the reference's iterative path
    (its own NSPCG, cg + ic, tightened maxerr) against the CUDA algorithm, both preconditioners"""
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built")
    p = cf.config_A(64)
    o = oracle_thermal(p, algorithm="iterative", precond="ic", itmaxerr=1e-12, maxit=5000)
    o.compute(0)
    s = Static3D("A64")
    s.problem = p
    s.iterative.preconditioner = precond
    s.iterative.maxerr = 1e-11
    s.iterative.maxit = 200000
    s.compute(0)
    assert s.stats["outer_loops"] == len(o.history)
    assert np.abs(s.outTemperature() - o.temperatures).max() <= TOL_T
    assert s.maxT == pytest.approx(o.maxT, abs=TOL_T)
    s.invalidate()


def test_random_problem_linear_solve_and_flux():
    p = random_problem((23, 17, 29), "120", nd_frac=0.02)
    p.maxerr = 1e-6
    o = oracle_thermal(p, algorithm="cholesky")
    o.compute(3)
    s, _ = gpu_solve(p, 3, loops=3, lin_tol=1e-12)
    T = s.outTemperature()
    assert np.abs(T - o.temperatures).max() <= 1e-6
    # heat flux provider (therm3d.cpp:342-384): conds re-evaluated at the final temperatures
    flux_ref = o.heat_fluxes()
    flux = s.outHeatFlux()
    assert np.abs(flux - flux_ref).max() <= 1e-6 * np.abs(flux_ref).max()
    s.invalidate()


def test_manufactured_parabola():
    """uniform k, uniform heat Q, T0 at the bottom, insulated elsewhere: T(z) = T0 + Q (2 H z - z^2)/(2 k)"""
    n = (6, 5, 41)
    H, k, Q = 10., 45., 1e15
    axes = [np.linspace(0, 3., n[0]), np.linspace(0, 2., n[1]), np.linspace(0, H, n[2])]
    tab = np.full((1, 2), k)
    p = cf.Problem("parabola", "thermal", axes, "012", None, 300., 1000., tab, tab.copy(), None, None)
    p.elem_mat = np.zeros(p.E, dtype=np.uint32)
    p.heat = np.full(p.E, Q)
    ng = np.broadcast_to(p.node_index_grid(), n)
    p.bc_nodes = ng[:, :, 0].ravel().astype(np.uintp)
    p.bc_values = np.full(p.bc_nodes.size, 300.)
    s, _ = gpu_solve(p, 3, lin_tol=1e-12)
    T = s.outTemperature()[ng]
    z = axes[2] * 1e-6
    exact = 300. + Q * (2 * H * 1e-6 * z - z * z) / (2 * k)
    assert np.abs(T - exact[None, None, :]).max() <= 1e-6 * exact.max()   # linear FEM is nodally exact here
    s.invalidate()


def test_warm_start_and_loops_limit():
    p = cf.config_B(20)
    s, _ = gpu_solve(p, 3, loops=1)
    assert s.stats["outer_loops"] == 1 and s.loopno == 1
    s.compute(1)
    assert s.loopno == 2
    o = oracle_thermal(p, algorithm="cholesky")
    o.compute(1); o.compute(1)
    assert np.abs(s.outTemperature() - o.temperatures).max() <= TOL_T
    s.invalidate()


def test_bad_input_is_rejected():
    import plask_b200
    p = cf.config_A(8)
    s = Static3D("t")
    s.problem = p
    s.algorithm = "cholesky"
    with pytest.raises(plask_b200.BadInput):
        s.compute(1)
    s.algorithm = "cuda"
    bad = cf.config_A(8)
    bad.bc_nodes = np.array([10 ** 9], dtype=np.uintp)
    bad.bc_values = np.array([300.])
    s.problem = bad
    with pytest.raises(plask_b200.BadInput):
        s.compute(1)
    s.invalidate()


def test_not_spd_reports_computation_error():
    """negative conductivity -> p.Ap <= 0 -> ComputationError, like NSPCG ier -6/-7
    (iterative_matrix.hpp:285-286)"""
    import plask_b200
    p = cf.config_A(10)
    p.tab_lat = -p.tab_lat
    p.tab_vert = -p.tab_vert
    s = Static3D("t")
    s.problem = p
    with pytest.raises(plask_b200.ComputationError):
        s.compute(1)
    s.invalidate()


def test_full_size_properties_256():
    """BASELINE configs[1] at full size (256^3): the oracle cannot factorise this, so check
    size-independent properties: the linear residual target is met, the solution satisfies the
    discrete maximum principle (T >= 300 with non-negative heat), the energy balance holds
    (heat in == flux through the Dirichlet plane, via the load vector), and both kernel variants
    give the same first-loop field."""
    p = cf.config_B(256)
    s = Static3D("B")
    s.problem = p
    s.iterative.maxerr = 1e-10     # the energy balance below sums the residual over 16.7 M rows
    s.iterative.maxit = 100000
    s.compute(1)
    T = s.outTemperature()
    assert s.iterative.converged and s.iterative.err <= 1e-8
    # (Q1 bricks with large aspect ratios do not give an M-matrix, so small undershoots are legitimate)
    assert T.min() >= 299. and np.isfinite(T).all()
    # energy balance: sum_i (A T)_i over all nodes == 0 for the unconstrained operator, hence the
    # reaction on the Dirichlet rows equals the total load: sum_free b_i = - sum_fixed (A T)_i.
    f = s._fem
    total_heat = float((p.heat * 1e-18 * _elem_volumes(p)).sum())
    f.set_dirichlet(np.zeros(0, dtype=np.uintp), np.zeros(0))   # unconstrained operator
    AT = f.apply(T, variant=3)
    fixed = np.zeros(p.N, dtype=bool); fixed[p.bc_nodes] = True
    reaction = -AT[fixed].sum()
    assert reaction == pytest.approx(total_heat, rel=1e-5)
    s.invalidate()


def _elem_volumes(p):
    d = [np.diff(a) for a in p.axes]
    vol = d[0][:, None, None] * d[1][None, :, None] * d[2][None, None, :]
    out = np.empty(p.E)
    out[p.elem_index_grid().ravel()] = vol.ravel()
    return out


@pytest.mark.parametrize("precond", ["jac", "ljac"])
def test_degenerate_systems(precond):
    """edge cases of the linear system: every node fixed (no free row at all), and a homogeneous problem whose
    solution is the boundary value itself (zero right-hand side after lifting)"""
    rng = np.random.default_rng(9)
    p = cf.config_A(9)
    vals = rng.uniform(290., 310., size=p.N)
    p.bc_nodes = np.arange(p.N, dtype=np.uintp)
    p.bc_values = vals
    s = Static3D("allfixed")
    s.problem = p
    s.iterative.preconditioner = precond
    s.compute(1)
    assert np.array_equal(s.outTemperature(), vals) and s.stats["lin_iters"] == 0 and s.iterative.converged
    s.invalidate()
    p = cf.config_A(9)
    p.heat = np.zeros(p.E)
    s = Static3D("homogeneous")
    s.problem = p
    s.inittemp = 345.
    s.iterative.preconditioner = precond
    s.iterative.maxerr = 1e-12
    err = s.compute(0)
    T = s.outTemperature()
    assert np.abs(T - 300.).max() <= 1e-6 and s.stats["outer_loops"] <= 3 and err >= 44.9   # first loop moves 345 K -> 300 K
    s.invalidate()


@pytest.mark.parametrize("precond", ["jac", "ljac"])
def test_nonlinear_loop_converges_to_the_kirchhoff_solution(precond):
    """the CUDA path against the exact solution of (k(T) T')' = -q with k = k0 (300/T)^a (Kirchhoff transform), second-order
    convergence in h — the same first-principles check tests/test_oracle_pin.py applies to the oracle"""
    from helpers import face_nodes
    k0, a, Tb, q, H = 45., 1.28, 300., 8.0e13, 6.0
    Tt = 250. + 0.05 * np.arange(6001)
    lat = (k0 * (300. / Tt) ** a)[None, :]
    c = k0 * 300. ** a

    def exact(z_um):
        z = z_um * 1e-6
        theta = q * (2. * H * 1e-6 * z - z * z) / 2.
        return (Tb ** (1. - a) + (1. - a) * theta / c) ** (1. / (1. - a))

    errs = []
    for nz in (9, 17, 33):
        axes = [np.linspace(0., 1., 3), np.linspace(0., 1., 3), np.linspace(0., H, nz)]
        p = cf.Problem("kirchhoff", "thermal", axes, "012", None, 250., 0.05, lat, lat.copy(), None, None)
        p.elem_mat = np.zeros(p.E, dtype=np.uint32)
        bot = face_nodes(p, 2, 0)
        p.bc_nodes, p.bc_values = bot.astype(np.uintp), np.full(bot.size, Tb)
        p.heat = np.full(p.E, q)
        s = Static3D("kirchhoff")
        s.problem = p
        s.inittemp, s.maxerr = Tb, 1e-9
        s.iterative.preconditioner = precond
        s.iterative.maxerr = 1e-13
        s.compute(200)
        T = s.outTemperature()[np.broadcast_to(p.node_index_grid(), p.n)][1, 1, :]
        errs.append(np.abs(T - exact(axes[2])).max())
        s.invalidate()
    assert errs[0] / errs[1] == pytest.approx(4., rel=0.25) and errs[1] / errs[2] == pytest.approx(4., rel=0.25), errs
    assert errs[2] < 0.02


@pytest.mark.parametrize("order", ["012", "201"])
def test_overlapping_dirichlet_conditions_first_lifts_last_stays(order):
    """two conditions naming the same nodes with different values (an edge shared by two contacts): setBC handles them one by
    one (iterative_matrix.hpp:462-485) — the free neighbours see the FIRST value, the node itself keeps the LAST"""
    p = cf.config_B((10, 11, 24), order=order)
    top = face_nodes_(p, 2, -1)
    side = face_nodes_(p, 0, 0)
    shared = np.intersect1d(top, side)
    assert shared.size >= 8
    p.bc_nodes = np.concatenate([p.bc_nodes, top, side]).astype(np.uintp)      # bottom 300 K, top 310 K, side 290 K
    p.bc_values = np.concatenate([p.bc_values, np.full(top.size, 310.), np.full(side.size, 290.)])
    o = oracle_thermal(p, algorithm="cholesky")
    o.compute(0)
    assert np.all(o.temperatures[shared] == 290.)
    s = Static3D("dup")
    s.problem = p
    s.iterative.maxerr = 1e-12
    s.iterative.maxit = 100000
    s.compute(0)
    T = s.outTemperature()
    assert np.all(T[shared] == 290.)
    assert np.abs(T - o.temperatures).max() <= 1e-3
    # the other order of the two conditions is a different problem for the neighbours of the shared edge
    q = cf.config_B((10, 11, 24), order=order)
    q.bc_nodes = np.concatenate([q.bc_nodes, side, top]).astype(np.uintp)
    q.bc_values = np.concatenate([q.bc_values, np.full(side.size, 290.), np.full(top.size, 310.)])
    o2 = oracle_thermal(q, algorithm="cholesky")
    o2.compute(0)
    free = np.ones(p.N, dtype=bool)
    free[p.bc_nodes.astype(np.int64)] = False
    assert np.abs(o2.temperatures - o.temperatures)[free].max() > 0.1
    s.invalidate()
