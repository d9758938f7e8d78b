"""Host-side logic of the solver mirror that needs no GPU: the integrals it computes from downloaded fields
(getTotalEnergy / getCapacitance, electr3d.cpp:568-610) and the masked-mesh bookkeeping, against the oracle."""
import numpy as np
import pytest

from helpers import oracle_shockley, shockley3d_reference_problem
from plask_b200 import _lib as L
from plask_b200 import configs as cf
from plask_b200.solvers import Shockley3D, Static3D


def _mirror_with_field(p, potential, empty_elements="include"):
    e = Shockley3D("host")
    e._problem = p                       # no device: only the host-side formulas are exercised
    e.empty_elements = empty_elements
    e.outVoltage = lambda: potential
    return e


def test_energy_and_capacitance_formula_vs_oracle():
    p = shockley3d_reference_problem()
    o = oracle_shockley(p, algorithm="cholesky", eps=p.meta["eps"])
    o.compute(20)
    e = _mirror_with_field(p, o.potential)
    assert e.get_total_energy() == pytest.approx(o.get_total_energy(), rel=1e-12)
    assert e.get_capacitance() == pytest.approx(o.get_capacitance(), rel=1e-12)
    oi = oracle_shockley(p, algorithm="cholesky", eps=p.meta["eps"], included=(p.empty == 0).astype(np.uint8))
    oi.compute(20)
    ei = _mirror_with_field(p, oi.potential, "exclude")
    assert ei.get_total_energy() == pytest.approx(oi.get_total_energy(), rel=1e-12)
    p.bc_values = np.concatenate([p.bc_values, [0.5]])
    p.bc_nodes = np.concatenate([p.bc_nodes, p.bc_nodes[:1]])
    with pytest.raises(L.BadInput):
        e.get_capacitance()              # exactly two voltage conditions are required (electr3d.cpp:603-605)


@pytest.mark.parametrize("order", ["012", "201"])
def test_masked_nodes_match_the_masked_mesh(order):
    from oracle import oracle as orc
    from helpers import oracle_mesh
    p = cf.config_C((12, 14, 40), order=order)
    s = Static3D("m")
    s._problem = p
    s.empty_elements = "exclude"
    ref = orc.DpbMasked(oracle_mesh(p), (p.empty == 0).astype(np.uint8))
    assert np.array_equal(s.masked_nodes(), ref.active)
    m = s._elem_materials()
    assert np.all(m[p.empty != 0] == L.MAT_EXCLUDED) and np.array_equal(m[p.empty == 0], p.elem_mat[p.empty == 0])
    nodes, vals = s._dirichlet()
    assert np.all(ref.active[nodes.astype(np.int64)]) and len(nodes) <= len(p.bc_nodes)
    s.empty_elements = "include"
    assert s.masked_nodes().all() and s._elem_materials() is p.elem_mat
    s.empty_elements = "sometimes"
    with pytest.raises(L.BadInput):
        s._elem_materials()


@pytest.mark.parametrize("cyl", [False, True])
def test_2d_mirror_integrals_vs_oracle2d(cyl):
    """integrateCurrent / getTotalHeat / getTotalEnergy / getCapacitance of the 2-D solvers (electr2d.cpp:466-649) as the mirror computes
    them from downloaded fields — here the oracle's fields are injected, no device is involved — and the bookkeeping of embed()"""
    from helpers import oracle_shockley2d, shockley2d_reference_problem
    from plask_b200.solvers2d import Shockley2D, ShockleyCyl, embed
    p2 = shockley2d_reference_problem(cyl, nx=3, ny=2)
    o = oracle_shockley2d(p2)
    o.compute(30)
    e = (ShockleyCyl if cyl else Shockley2D)("host2d")
    e._p2 = p2
    e._problem = embed(p2)
    e.outVoltage = lambda mesh=None: o.potentials
    e.outCurrentDensity = lambda: o.currents
    e.outHeat = lambda: o.heat_densities()
    vindex = int(np.nonzero(np.asarray(p2.elem_junc).reshape(p2.n[0] - 1, p2.n[1] - 1).any(axis=0))[0].min())   # bottom row of the junction
    assert e.integrate_current(vindex, True) == pytest.approx(o.integrate_current(vindex, True), rel=1e-12)
    assert e.integrate_current(vindex, False) == pytest.approx(o.integrate_current(vindex, False), rel=1e-12)
    assert e.get_total_heat() == pytest.approx(o.get_total_heat(), rel=1e-12)
    assert e.get_total_energy() == pytest.approx(o.get_total_energy(), rel=1e-12)
    assert e.get_capacitance() == pytest.approx(o.get_capacitance(), rel=1e-12)
    # the embedding: plane 0 keeps the 2-D numbering, plane 1 follows at offset N, every condition of the first kind on both planes
    p3 = e._problem
    assert p3.n == (2, p2.n[0], p2.n[1]) and p3.N == 2 * p2.N and p3.E == p2.E
    assert np.array_equal(p3.bc_nodes[:len(p2.bc_nodes)], p2.bc_nodes) and np.array_equal(p3.bc_nodes[len(p2.bc_nodes):], p2.bc_nodes + p2.N)
    assert np.array_equal(p3.elem_mat, p2.elem_mat) and np.array_equal(p3.elem_junc, p2.elem_junc)
    with pytest.raises(L.BadInput):
        (Shockley2D if cyl else ShockleyCyl)("wrong").problem = p2       # Cartesian problem on the cylindrical solver and vice versa


def test_2d_provider_interpolation_vs_oracle2d():
    """outTemperature / outVoltage of the 2-D mirror on a foreign rectangular mesh (host-side bilinear interpolation, clamped outside
    like interpolateLinear) against the oracle's independent implementation and against scipy"""
    from scipy.interpolate import RegularGridInterpolator
    from oracle import oracle2d
    from helpers import thermal2d_problem
    from plask_b200.solvers2d import interpolate2d
    p2 = thermal2d_problem((9, 12))
    rng = np.random.default_rng(2)
    f = rng.normal(size=p2.N)
    xq = np.concatenate([[p2.x[0] - 1.], rng.uniform(p2.x[0], p2.x[-1], 9), [p2.x[-1], p2.x[-1] + 2.]])
    yq = np.concatenate([[p2.y[0] - 0.3], rng.uniform(p2.y[0], p2.y[-1], 7), [p2.y[0], p2.y[-1] + 1.]])
    got = interpolate2d(p2, f, xq, yq)
    assert np.abs(got - oracle2d.interp_bilinear(p2.x, p2.y, f, xq, yq)).max() <= 1e-13
    X, Y = np.meshgrid(np.clip(xq, p2.x[0], p2.x[-1]), np.clip(yq, p2.y[0], p2.y[-1]), indexing="ij")
    ref = RegularGridInterpolator((p2.x, p2.y), f.reshape(p2.n))(np.stack([X, Y], axis=-1)).ravel()
    assert np.abs(got - ref).max() <= 1e-13
    assert np.abs(interpolate2d(p2, f, p2.x, p2.y) - f).max() == 0.       # the own mesh: identity
