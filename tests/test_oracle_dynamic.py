"""Pins of the Dynamic3D oracle (oracle.Dynamic3DOracle, corrected femT3d.cpp): the reference has no test of its dynamic
solver, so the restatement is pinned from first principles — analytic 1-D cooling (second order in the time step with
methodparam = 0.5, first order with 1), exact energy bookkeeping of the lumped scheme, steady state = Static3D."""
import numpy as np
import pytest

from helpers import cooling_exact, cooling_initial, cooling_problem, oracle_dynamic, oracle_thermal
from plask_b200 import configs as cf


def column(p, T):
    ng = np.broadcast_to(p.node_index_grid(), p.n)
    return T[ng[1, 1, :]]


@pytest.mark.parametrize("lumping", [True, False])
def test_cooling_second_order_in_time(lumping):
    p = cooling_problem(nz=41)
    errs = []
    for dt in (8., 4.):
        o = oracle_dynamic(p, timestep=dt, methodparam=0.5, lumping=lumping)
        o.temperatures = cooling_initial(p)
        o.compute(200.)
        t = o.physical_time                  # time really advanced: the loop of femT3d.cpp:271-272 does time/timestep + 1 steps,
        assert t == pytest.approx(200. + dt) and o.elapstime == pytest.approx(200.)   # while the reference's clock shows `time` (:297)
        errs.append(np.abs(column(p, o.temperatures) - cooling_exact(p, t)).max())
    assert errs[0] < 1.2e-2 and errs[1] < 4e-3
    assert errs[1] < errs[0] / 2.5           # better than first order; limited by the O(h^2) space error at the fine step


def test_backward_euler_first_order():
    p = cooling_problem(nz=41)
    errs = []
    for dt in (4., 2.):
        o = oracle_dynamic(p, timestep=dt, methodparam=1.0)
        o.temperatures = cooling_initial(p)
        o.compute(100.)
        errs.append(np.abs(column(p, o.temperatures) - cooling_exact(p, o.physical_time)).max())
    assert 1.6 < errs[0] / errs[1] < 2.4


def test_lumped_energy_balance():
    """insulated box, lumped capacity: sum_n C_nn (T^{n+1} - T^n) = sum F exactly every step (K 1 = 0)"""
    p = cf.config_B((10, 11, 22))
    p.bc_nodes, p.bc_values = np.zeros(0, dtype=np.uintp), np.zeros(0)
    o = oracle_dynamic(p, timestep=5., methodparam=0.5, lumping=True)
    A, B, F = o.set_matrix()
    cdiag = np.asarray((A - B * 0).diagonal())   # noqa: F841 (shape check only)
    mass = np.asarray((A + B).sum(axis=1)).ravel() / 2. if False else None
    T0 = o.temperatures.copy()
    o.compute(20.)
    steps = len(o.maxT_log)
    Cm = (A - o.methodparam * (A - B))           # C = A - theta K with K = A - B
    dE = float(np.asarray(Cm.sum(axis=1)).ravel() @ (o.temperatures - T0))
    assert dE == pytest.approx(steps * F.sum(), rel=1e-9)


def test_steady_state_is_static3d():
    """constant coefficients: the time loop converges to the solution of K T = F with the same Dirichlet set"""
    p = cf.config_A(8)
    p.tab_lat = np.repeat(p.tab_lat[:, :1], p.tab_lat.shape[1], axis=1)
    p.tab_vert = np.repeat(p.tab_vert[:, :1], p.tab_vert.shape[1], axis=1)
    s = oracle_thermal(p, algorithm="cholesky")
    s.compute(0)
    o = oracle_dynamic(p, timestep=2000., methodparam=1.0)
    o.compute(200000.)
    assert np.abs(o.temperatures - s.temperatures).max() < 1e-6 * (s.temperatures.max() - 300. + 1.)


def test_rebuild_changes_the_result():
    """k(T) matters: re-evaluating it every step differs from freezing it at the start, by much less than the rise itself"""
    p = cf.config_B((8, 9, 20))
    p.heat = p.heat * 40.
    a = oracle_dynamic(p, timestep=20., rebuildfreq=0)
    b = oracle_dynamic(p, timestep=20., rebuildfreq=1)
    a.compute(400.)
    b.compute(400.)
    rise = a.temperatures.max() - 300.
    d = np.abs(a.temperatures - b.temperatures).max()
    assert rise > 5. and 1e-6 < d < 0.2 * rise
