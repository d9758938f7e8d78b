"""GPU parity of the 2-D solvers: Static2D / StaticCyl (therm2d.cpp) and Shockley2D / ShockleyCyl (electr2d.cpp)
run by the brick kernels through a one-layer embedding (plask_b200/solvers2d.py, pfem_set_axis_weight for the radial weight),
against the independent two-dimensional oracle (oracle/oracle2d.py) and the reference's own analytic pins
(solvers/electrical/shockley/tests/shockley2d.py).  North star: <= 1e-3 K, <= 1e-6 V."""
import numpy as np
import pytest

from helpers import (oracle_dynamic2d, oracle_shockley2d, oracle_static2d, shockley2d_reference_problem, thermal2d_problem,
                     thermoelectric2d_pair)
from plask_b200 import _lib as L
from plask_b200.solvers2d import (Dynamic2D, DynamicCyl, Shockley2D, ShockleyCyl, Static2D, StaticCyl, ThermoElectric2D,
                                  ThermoElectricCyl)

pytestmark = pytest.mark.gpu

EPS0_PF_UM = 8.854187817e-6


def thermal(p2, precond="jac", variant=3):
    s = (StaticCyl if p2.cyl else Static2D)("therm2d")
    s.problem = p2
    s.variant = variant
    s.iterative.preconditioner = precond
    s.iterative.maxerr = 1e-11
    s.iterative.maxit = 100000
    return s


def electrical(p2, precond="jac", **kw):
    e = (ShockleyCyl if p2.cyl else Shockley2D)("electr2d")
    e.problem = p2
    e.beta, e.js, e.maxerr = p2.beta, p2.js, p2.maxerr
    e.iterative.preconditioner = precond
    e.iterative.maxerr = 1e-13
    e.iterative.maxit = 200000
    for k, v in kw.items():
        setattr(e, k, v)
    return e


@pytest.mark.parametrize("cyl", [False, True])
@pytest.mark.parametrize("precond,variant", [("jac", 3), ("ljac", 3), ("mlj", 3), ("jac", 1)])
def test_static2d_vs_oracle(cyl, precond, variant):
    p2 = thermal2d_problem(cyl=cyl)
    o = oracle_static2d(p2)
    o.compute(0)
    s = thermal(p2, precond, variant)
    err = s.compute(0)
    T = s.outTemperature()
    assert o.maxT - 300. > 5.
    assert s.stats["outer_loops"] == len(o.history)
    assert np.abs(T - o.temperatures).max() <= 1e-3
    assert np.abs(T - o.temperatures).max() <= 1e-5
    assert err == pytest.approx(max(h["err"] for h in o.history), abs=1e-5)
    assert abs(s.maxT - o.maxT) <= 1e-5
    # both node planes of the embedding hold the same field
    full = s._fem.get_field()
    assert np.abs(full[:p2.N] - full[p2.N:]).max() <= 1e-6
    F = s.outHeatFlux()
    Fo = o.heat_fluxes()
    assert np.abs(F - Fo).max() <= 1e-6 * np.abs(Fo).max()
    s.invalidate()


def boundary_conditions_2d(p2):
    """heat flux into the top edge over the hot region, convection on the outer (right) edge and on the rest of the top, radiation
    towards a hot ambient on the top: node-wise conditions, the corner node shared by two of them (the first definition wins)"""
    n0, n1 = p2.n
    ng = np.arange(n0 * n1).reshape(n0, n1)
    top, right = ng[:, n1 - 1], ng[n0 - 1, :]
    hot = top[:n0 // 3 + 1]
    return dict(heatflux=[(hot, -2.0e8)], convection=[(right[4:], 2.0e6, 320.), (top[n0 // 3:], 6.0e5, 295.)],
                radiation=[(top, 0.8, 2500.)])


@pytest.mark.parametrize("cyl", [False, True])
@pytest.mark.parametrize("verbatim,precond", [(False, "jac"), (False, "ljac"), (False, "mlj"), (True, "jac"), (True, "ljac")])
def test_static2d_boundary_conditions_vs_oracle(cyl, verbatim, precond):
    """therm2d.cpp:138-172, :225-265 (Cartesian), :371-413 (cylindrical) through pfem_boundary::mode2d against the 2-D oracle;
    verbatim = the convection matrix as the reference writes it (no 1e-6, second factor r), corrected = with them"""
    p2 = thermal2d_problem(cyl=cyl)
    conds = boundary_conditions_2d(p2)
    o = oracle_static2d(p2)
    o.heatflux, o.convection, o.radiation, o.verbatim = conds["heatflux"], conds["convection"], conds["radiation"], verbatim
    o.compute(0)
    plain = oracle_static2d(p2)
    plain.compute(0)
    assert np.abs(o.temperatures - plain.temperatures).max() > 1.           # the conditions matter
    s = thermal(p2, precond)
    s.heatflux_boundary, s.convection_boundary, s.radiation_boundary = conds["heatflux"], conds["convection"], conds["radiation"]
    s.boundary_verbatim = verbatim
    s.compute(0)
    T = s.outTemperature()
    assert s.stats["outer_loops"] == len(o.history)
    assert np.abs(T - o.temperatures).max() <= 1e-3
    assert np.abs(T - o.temperatures).max() <= 1e-5
    full = s._fem.get_field()
    assert np.abs(full[:p2.N] - full[p2.N:]).max() <= 1e-6
    s.invalidate()


def test_static2d_boundary_mode_needs_the_embedding():
    from helpers import face_nodes
    from plask_b200 import configs
    from plask_b200.fem import DeviceFem
    p = configs.config_B((6, 6, 20))
    f = DeviceFem(0)
    f.set_mesh(p.axes, p.strides)
    with pytest.raises(L.BadInput):
        f.set_boundary([], [(face_nodes(p, 2, -1), 1e4, 300.)], [], False, mode2d=1)
    f.close()


@pytest.mark.parametrize("cyl", [False, True])
@pytest.mark.parametrize("precond,variant,lumping", [("jac", 3, True), ("ljac", 3, True), ("jac", 1, True), ("jac", 1, False)])
def test_dynamic2d_vs_oracle(cyl, precond, variant, lumping):
    """femT2d.cpp through the one-layer embedding (pfem_solve_dynamic) against the 2-D oracle: k(T), cp(T) rebuilt every 3 steps,
    two calls of compute continue one trajectory"""
    p2 = thermal2d_problem((21, 26), cyl=cyl)
    o = oracle_dynamic2d(p2, timestep=0.5, methodparam=0.5, lumping=lumping, rebuildfreq=3)
    o.compute(3.)
    o.compute(2.)
    s = (DynamicCyl if cyl else Dynamic2D)("dyn2d")
    s.problem = p2
    s.variant = variant
    s.iterative.preconditioner = precond
    s.iterative.maxerr = 1e-12
    s.iterative.maxit = 100000
    s.timestep, s.methodparam, s.lumping, s.rebuildfreq = 0.5, 0.5, lumping, 3
    s.compute(3.)
    s.compute(2.)
    T = s.outTemperature()
    assert o.temperatures.max() - 300. > 1.
    assert s.time == pytest.approx(o.elapstime)
    assert np.abs(T - o.temperatures).max() <= 1e-6
    assert abs(s.maxT - o.maxT) <= 1e-6
    full = s._fem.get_field()
    assert np.abs(full[:p2.N] - full[p2.N:]).max() <= 1e-6
    assert s.outHeatFlux().shape == (p2.E, 2)
    s.invalidate()


@pytest.mark.parametrize("cyl", [False, True])
def test_shockley2d_py_through_the_gpu(cyl):
    """shockley2d.py testComputations with the reference's own tolerances, and the oracle field by field"""
    p2 = shockley2d_reference_problem(cyl)
    o = oracle_shockley2d(p2)
    o.compute(1000 if cyl else 0)
    # maxerr = 1e-5 % of shockley2d.py:52 is a current change of 1e-7: in the 1e9 S/m contacts that is a potential noise of 1e-15 V,
    # which a direct solver just reaches and no iterative solver does — the CUDA path runs the loop count of the direct solve
    e = electrical(p2)
    e.compute(len(o.history))
    geo = np.pi if cyl else 1.
    current = 1e-3 * geo * p2.js * (np.exp(p2.beta) - 1.)
    assert abs(e.get_total_current() - current) < 0.5e-3
    assert abs(e.get_capacitance() - EPS0_PF_UM * 12.9 * geo * 1000. ** 2 / 0.02) < 0.5e-2
    assert abs(e.get_total_heat() - current) < 0.5e-3
    assert 100 < len(o.history) < 200 and e.stats["outer_loops"] == len(o.history)
    assert np.abs(e.outVoltage() - o.potentials).max() <= 1e-6
    # element by element where the current is not a difference of potentials at the 1e-15 V level (the 1e9 S/m contacts)
    soft = o.conds[:, 1] < 1e6
    assert soft.sum() >= p2.n[0] - 1
    assert np.allclose(e.outCurrentDensity()[soft], o.currents[soft], rtol=1e-6, atol=1e-9 * np.abs(o.currents).max())
    assert np.allclose(e.outHeat()[soft], o.heat_densities()[soft], rtol=1e-6, atol=1e-9 * np.abs(o.heat_densities()).max())
    assert np.abs(e.outCurrentDensity() - o.currents).max() <= 2e-3 * np.abs(o.currents).max()
    assert e.stats["maxval"] == pytest.approx(o.history[-1]["mcur"], rel=1e-6)      # max |j| at the junction
    assert e.stats["err"] < 0.5                                                      # loop error floor set by the contact currents, %
    e.invalidate()


def test_shockleycyl_temperature_dependent_beta():
    """shockley2d.py:110-118"""
    p2 = shockley2d_reference_problem(True)
    e = electrical(p2, beta=lambda T: np.log(T * 70.))
    e.compute(200)
    assert abs(e.get_total_current() - 1e-3 * np.pi * (21000. - 1.)) < 0.5e-3
    e.inTemperature = 250.
    e.compute(200)
    assert abs(e.get_total_current() - 1e-3 * np.pi * (17500. - 1.)) < 0.5e-3
    e.invalidate()


@pytest.mark.parametrize("cyl", [False, True])
def test_refined_junction_problem_vs_oracle(cyl):
    """the same structures on finer meshes (lateral current spreading under the ring contact in the cylindrical case), line-Jacobi"""
    p2 = shockley2d_reference_problem(cyl, nx=6, ny=5)
    p2.maxerr = 1e-3
    o = oracle_shockley2d(p2)
    o.compute(8)
    e = electrical(p2, "ljac")
    e.compute(8)
    assert np.abs(e.outVoltage() - o.potentials).max() <= 1e-6
    assert abs(e.get_total_current() - o.get_total_current()) <= 1e-6 * abs(o.get_total_current())
    e.invalidate()


def test_axis_weight_rejects_bad_input_and_boundary_terms():
    p2 = thermal2d_problem((9, 11), cyl=True)
    s = thermal(p2)
    s.initialize()
    with pytest.raises(L.BadInput):
        s._fem.set_axis_weight(1, -np.ones(p2.n[0] - 1))
    with pytest.raises(L.BadInput):
        s._fem.set_boundary([], [(np.arange(4), 1e4, 300.)], [], False)
    s._fem.set_axis_weight(1, None)          # weights removed: boundary terms are accepted again
    s._fem.set_boundary([], [], [], False)
    s.invalidate()


@pytest.mark.parametrize("cyl", [False, True])
@pytest.mark.parametrize("thermal_shape", [None, (12, 14)])
def test_thermoelectric2d_meta_loop_vs_oracles(cyl, thermal_shape):
    """meta.shockley.ThermoElectric2D / ThermoElectricCyl (thermoelectric.py:187-211): sigma(T), k(T) and a temperature-dependent
    junction beta(T), the fields exchanged on the device — same mesh, and a coarser thermal mesh that reaches below the electrical one"""
    from oracle import oracle2d
    pt, pe = thermoelectric2d_pair(cyl, thermal_shape=thermal_shape)
    beta = lambda T: 11. * 300. / T
    te = (ThermoElectricCyl if cyl else ThermoElectric2D)("te2d")
    te.thermal.problem, te.electrical.problem = pt, pe
    te.tfreq = 4
    for s in (te.thermal, te.electrical):
        s.iterative.maxerr, s.iterative.maxit = 1e-13, 200000
    te.electrical.beta, te.electrical.js, te.electrical.maxerr = beta, pe.js, pe.maxerr
    te.thermal.maxerr = pt.maxerr
    n = te.compute(max_meta_loops=5)
    o = oracle2d.ThermoElectric2DOracle(oracle_static2d(pt), oracle_shockley2d(pe, beta=beta), tfreq=4)
    no = o.compute(max_meta_loops=5)
    assert n == no == 5
    T, V = te.thermal.outTemperature(), te.electrical.outVoltage()
    assert o.thermal.maxT > 310.
    assert np.abs(T - o.thermal.temperatures).max() <= 1e-3
    assert np.abs(V - o.electrical.potentials).max() <= 1e-6
    assert te.get_total_current() == pytest.approx(o.electrical.get_total_current(), rel=1e-6)
    for h, g in zip(te.history, o.history):
        assert h["terr"] == pytest.approx(g["terr"], abs=1e-3)
        assert h["verr"] == pytest.approx(g["verr"], rel=1e-3, abs=1e-6)
    te.invalidate()
