"""GPU parity of Dynamic3D (solvers/thermal/dynamic/femT3d.cpp:127-305, corrected update — DESIGN.md §2)
through the solver mirror and pfem_solve_dynamic, against oracle.Dynamic3DOracle (sparse direct solve of every step) and the
analytic 1-D cooling solution.  Tolerance of the north star: max |dT| <= 1e-3 K."""
import numpy as np
import pytest

from helpers import cooling_exact, cooling_initial, cooling_problem, oracle_dynamic
from plask_b200 import _lib as L
from plask_b200 import configs as cf
from plask_b200.solvers import Dynamic3D

pytestmark = pytest.mark.gpu

TOL_T = 1e-3


def gpu_dynamic(p, precond="jac", variant=3, lin_tol=1e-12, **kw):
    s = Dynamic3D("dynamic")
    s.problem = p
    s.inittemp = p.inittemp
    s.variant = variant
    s.iterative.preconditioner = precond
    s.iterative.maxerr = lin_tol
    s.iterative.maxit = 20000
    s.logfreq = 0
    for k, v in kw.items():
        setattr(s, k, v)
    return s


def heated(shape=(14, 15, 30), order="012", scale=40.):
    p = cf.config_B(shape, order=order)
    p.heat = p.heat * scale
    return p


@pytest.mark.parametrize("precond,variant", [("jac", 3), ("ljac", 3), ("jac", 1)])
@pytest.mark.parametrize("theta", [0.5, 1.0])
def test_lumped_vs_oracle(precond, variant, theta):
    p = heated()
    o = oracle_dynamic(p, timestep=20., methodparam=theta, lumping=True)
    o.compute(200.)
    s = gpu_dynamic(p, precond, variant, timestep=20., methodparam=theta)
    s.compute(200.)
    T = s.outTemperature()
    assert s.time == pytest.approx(o.elapstime)
    assert o.temperatures.max() - 300. > 1.          # the transient is under way, not a flat field
    assert np.abs(T - o.temperatures).max() <= 1e-6
    assert abs(s.maxT - o.maxT) <= 1e-6
    s.invalidate()


@pytest.mark.parametrize("order", ["021", "102", "120", "201", "210"])
def test_lumped_all_orders(order):
    p = heated((12, 13, 26), order=order)
    o = oracle_dynamic(p, timestep=25., methodparam=0.5)
    o.compute(100.)
    s = gpu_dynamic(p, "ljac", timestep=25.)
    s.compute(100.)
    assert np.abs(s.outTemperature() - o.temperatures).max() <= 1e-6
    s.invalidate()


def test_consistent_capacity_vs_oracle():
    p = heated()
    o = oracle_dynamic(p, timestep=20., methodparam=0.5, lumping=False)
    o.compute(100.)
    s = gpu_dynamic(p, "jac", 1, timestep=20., lumping=False)
    s.compute(100.)
    assert np.abs(s.outTemperature() - o.temperatures).max() <= 1e-6
    s.invalidate()
    bad = gpu_dynamic(p, "jac", 3, timestep=20., lumping=False)
    with pytest.raises(L.BadInput):
        bad.compute(20.)
    bad.invalidate()


def test_rebuild_and_log_and_continuation():
    """k(T), cp(T) re-evaluated every 2 steps; two calls of compute continue the same trajectory; the per-step max T log"""
    p = heated(scale=200.)
    o = oracle_dynamic(p, timestep=20., rebuildfreq=2)
    o.compute(100.)
    o.compute(60.)
    s = gpu_dynamic(p, timestep=20., rebuildfreq=2, logfreq=1)
    s.compute(100.)
    n1 = s.stats["outer_loops"]
    s.compute(60.)
    assert n1 == 6 and s.stats["outer_loops"] == 4        # time/timestep + 1 steps per call (femT3d.cpp:271-272)
    assert s.time == pytest.approx(o.elapstime) == pytest.approx(160.)
    assert s.physical_time == pytest.approx(o.physical_time) == pytest.approx(200.)
    assert o.maxT - 300. > 10.
    assert np.abs(s.outTemperature() - o.temperatures).max() <= TOL_T
    assert np.abs(s.outTemperature() - o.temperatures).max() <= 1e-5
    assert any(m.startswith("Time") for _, m in s.log)
    s.invalidate()


def test_cooling_analytic_gpu():
    """the corrected scheme against first principles: Crank-Nicolson cooling of a 1-D slab, error as small as the oracle's"""
    p = cooling_problem(nz=41)
    s = gpu_dynamic(p, "ljac", timestep=4.)
    s.initialize()
    s._fem.set_field(cooling_initial(p))
    s.compute(200.)
    ng = np.broadcast_to(p.node_index_grid(), p.n)
    T = s.outTemperature()
    err = np.abs(T[ng[1, 1, :]] - cooling_exact(p, s.physical_time)).max()
    assert s.time == pytest.approx(200.) and s.physical_time == pytest.approx(204.)   # femT3d.cpp:271-272 against :297
    assert err < 4e-3
    o = oracle_dynamic(p, timestep=4.)
    o.temperatures = cooling_initial(p)
    o.compute(200.)
    assert np.abs(T - o.temperatures).max() <= 1e-7
    s.invalidate()


def test_steady_state_is_static3d_gpu():
    from plask_b200.solvers import Static3D
    p = cf.config_A(16)
    p.tab_lat = np.repeat(p.tab_lat[:, :1], p.tab_lat.shape[1], axis=1)
    p.tab_vert = np.repeat(p.tab_vert[:, :1], p.tab_vert.shape[1], axis=1)
    st = Static3D("static")
    st.problem = p
    st.iterative.maxerr = 1e-12
    st.compute(0)
    s = gpu_dynamic(p, "ljac", timestep=1e5, methodparam=1.0)
    s.compute(2e7)
    rise = st.maxT - 300.
    assert rise > 0.01
    assert np.abs(s.outTemperature() - st.outTemperature()).max() < 1e-6 * (rise + 1.)
    # the providers after a dynamic run read the true (unscaled) conductivities
    assert np.allclose(s.outThermalConductivity(), st.outThermalConductivity())
    assert np.allclose(s.outHeatFlux(), st.outHeatFlux(), atol=1e-9 * np.abs(st.outHeatFlux()).max())
    s.invalidate()
    st.invalidate()


def test_bad_input():
    p = heated((8, 9, 12))
    s = gpu_dynamic(p, "mlj", timestep=10.)
    with pytest.raises(L.BadInput):
        s.compute(10.)
    s.invalidate()
    s = gpu_dynamic(p, timestep=-1.)
    with pytest.raises(L.BadInput):
        s.compute(10.)
    s.invalidate()
