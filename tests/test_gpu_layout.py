"""Internal layout (pfem_set_layout): with PFEM_LAYOUT_VERTICAL_MINOR the fields are stored with the vertical axis fastest
whatever the mesh's iteration order; every result that crosses the ABI must be unchanged.  Checked against the oracle
(operator level, nonlinear solves) and against the ABI layout (all providers), for the six iteration orders."""
import numpy as np
import pytest

from helpers import face_nodes, oracle_mesh, oracle_shockley, oracle_thermal, random_problem
from oracle import oracle as orc
from plask_b200 import _lib as L
from plask_b200 import configs as cf
from plask_b200.fem import DeviceFem
from plask_b200.solvers import Shockley3D, Static3D, ThermoElectric3D

pytestmark = pytest.mark.gpu

ORDERS = ["012", "021", "102", "120", "201", "210"]


def _conds(p):
    return dict(convection=[(face_nodes(p, 2, -1), 4.0e4, 310.), (face_nodes(p, 0, 0), 9.0e4, 295.)],
                heatflux=[(face_nodes(p, 1, -1), -3.0e5)], radiation=[(face_nodes(p, 1, 0), 0.85, 285.)])


@pytest.mark.parametrize("order", ORDERS)
def test_operator_rhs_diag_vertical_minor(order):
    p = random_problem((9, 7, 11), order)
    rng = np.random.default_rng(5)
    T = rng.uniform(280., 420., size=p.N)
    conds = _conds(p)
    m = oracle_mesh(p)
    tb = orc.Tables(p.T0, p.dT, p.tab_lat, p.tab_vert)
    s = orc.Static3DOracle(m, p.elem_mat, tb, p.bc_nodes, p.bc_values, heat=p.heat, algorithm="iterative",
                           boundaries=orc.BoundaryTerms(p.N, **conds), quirk=False)
    s.temperatures[:] = T
    A, B = orc.Sparse14(m), np.zeros(p.N)
    s.set_matrix(A, B)
    f = DeviceFem(0)
    f.set_layout(L.LAYOUT_VERTICAL_MINOR)
    f.set_mesh(p.axes, p.strides)
    f.set_materials(p.elem_mat, p.T0, p.dT, p.tab_lat, p.tab_vert)
    f.set_field(T)
    assert np.array_equal(f.get_field(), T)                 # node transfer round trip
    f.set_dirichlet(p.bc_nodes, p.bc_values)
    f.set_source(p.heat)
    f.set_boundary(verbatim=False, **conds)
    f.update_conductivity_thermal()
    assert np.array_equal(f.get_elem(0), s.conds)           # element transfer + same table arithmetic
    v = rng.standard_normal(p.N)
    scale = np.abs(A.data[:p.N]).max() * np.abs(v).max()
    for variant in (1, 2, 0, 3):
        assert np.abs(f.apply(v, variant=variant) - A.mult(v)).max() <= 2e-14 * scale, variant
    assert np.abs(f.get_rhs() - B).max() <= 1e-13 * np.abs(B).max()
    assert np.abs(f.get_diag() - A.data[:p.N]).max() <= 1e-14 * np.abs(A.data[:p.N]).max()
    with pytest.raises(L.BadInput):
        f.set_layout(L.LAYOUT_ABI)                          # not after set_mesh
    f.close()


@pytest.mark.parametrize("precond", ["jac", "ljac"])
@pytest.mark.parametrize("order", ORDERS)
def test_thermal_vertical_minor_vs_cholesky(order, precond):
    p = cf.config_B((14, 16, 40), order=order)
    o = oracle_thermal(p, algorithm="cholesky")
    o.compute(0)
    s = Static3D("lay")
    s.problem = p
    s.layout = "vertical-minor"
    s.iterative.preconditioner = precond
    s.iterative.maxerr, s.iterative.maxit = 1e-11, 100000
    s.compute(0)
    assert s.stats["outer_loops"] == len(o.history)
    assert np.abs(s.outTemperature() - o.temperatures).max() <= 1e-3
    ref = o.heat_fluxes()
    assert np.abs(s.outHeatFlux() - ref).max() <= 1e-6 * np.abs(ref).max()
    s.invalidate()


@pytest.mark.parametrize("order", ["012", "120", "201"])
def test_shockley_vertical_minor_all_providers(order):
    p = cf.config_C((20, 22, 52), order=order)
    LOOPS = 6
    res = {}
    for lay in ("abi", "vertical-minor"):
        e = Shockley3D(lay)
        e.problem = p
        e.layout = lay
        e.beta, e.js, e.maxerr = p.beta, p.js, p.maxerr
        e.iterative.maxerr, e.iterative.maxit = 1e-13, 200000
        e.compute(LOOPS)
        res[lay] = dict(V=e.outVoltage(), j=e.outCurrentDensity(), h=e.outHeat(), c=e.outConductivity(), jc=e._junc_cond.copy(),
                        I=e.get_total_current(), err=e.stats["err"], maxcur=np.array(e.maxcur))
        e.invalidate()
    a, b = res["abi"], res["vertical-minor"]
    o = oracle_shockley(p, algorithm="cholesky")
    o.compute(LOOPS)
    assert np.abs(b["V"] - o.potential).max() <= 1e-6
    assert np.abs(a["V"] - b["V"]).max() <= 1e-9
    assert np.abs(a["j"] - b["j"]).max() <= 1e-7 * np.abs(a["j"]).max()
    assert np.abs(a["h"] - b["h"]).max() <= 1e-6 * np.abs(a["h"]).max()
    assert np.array_equal(a["c"][p.elem_junc == 0], b["c"][p.elem_junc == 0])
    assert np.abs(a["jc"] - b["jc"]).max() <= 1e-7 * np.abs(a["jc"]).max()
    assert b["I"] == pytest.approx(a["I"], rel=1e-7)
    assert np.allclose(a["maxcur"], b["maxcur"], rtol=1e-6, atol=1e-12)


def test_thermoelectric_mixed_layouts():
    """field exchange between a thermal context in one layout and an electrical context in the other"""
    outs = []
    for lt, le in (("abi", "abi"), ("vertical-minor", "abi"), ("abi", "vertical-minor")):
        te = ThermoElectric3D("te")
        te.thermal.problem = cf.config_B((16, 16, 44), order="201")
        te.electrical.problem = cf.config_C((16, 16, 44), order="012")
        te.thermal.layout, te.electrical.layout = lt, le
        for s in (te.thermal, te.electrical):
            s.iterative.maxerr, s.iterative.maxit = 1e-12, 200000
        te.compute(invalidate=False, max_meta_loops=2)
        outs.append((te.thermal.outTemperature(), te.electrical.outVoltage()))
        te.invalidate()
    for T, V in outs[1:]:
        assert np.abs(T - outs[0][0]).max() <= 1e-7
        assert np.abs(V - outs[0][1]).max() <= 1e-9
