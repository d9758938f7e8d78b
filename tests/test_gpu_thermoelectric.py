"""ThermoElectric3D meta loop (solvers/meta/shockley/thermoelectric.py:207-211) with the device-resident field
exchange against the two coupled oracles.  Tolerances: 1e-3 K, 1e-6 V (north star); the
interpolated element temperatures are checked bit-exactly through the conductivities they select."""
import numpy as np
import pytest

from helpers import oracle_shockley, oracle_thermal
from oracle import oracle as orc
from plask_b200 import _lib as L
from plask_b200 import configs as cf
from plask_b200.fem import DeviceFem
from plask_b200.solvers import ThermoElectric3D

pytestmark = pytest.mark.gpu


def make_pair(nt, ne, order_t="optimal", order_e="optimal"):
    pt = cf.config_B(nt, order=order_t)
    pt.heat = None
    pe = cf.config_C(ne, order=order_e)
    return pt, pe


def run_both(pt, pe, meta_loops, tfreq):
    te = ThermoElectric3D("te")
    te.thermal.problem, te.electrical.problem = pt, pe
    te.tfreq = tfreq
    for s in (te.thermal, te.electrical):
        s.iterative.maxerr, s.iterative.maxit = 1e-12, 100000
    te.electrical.beta, te.electrical.js, te.electrical.maxerr = pe.beta, pe.js, pe.maxerr
    te.thermal.maxerr = pt.maxerr
    n = te.compute(max_meta_loops=meta_loops)
    ot = oracle_thermal(pt, algorithm="cholesky")
    oe = oracle_shockley(pe, algorithm="cholesky")
    o = orc.ThermoElectric3DOracle(ot, oe, tfreq=tfreq)
    no = o.compute(max_meta_loops=meta_loops)
    return te, o, n, no


@pytest.mark.parametrize("order", ["optimal", "012", "120"])
def test_same_mesh_meta_loop_vs_oracles(order):
    pt, pe = make_pair((20, 22, 52), (20, 22, 52), order, order)
    te, o, n, no = run_both(pt, pe, 3, 4)
    assert n == no == 3
    T, V = te.thermal.outTemperature(), te.electrical.outVoltage()
    assert np.abs(T - o.thermal.temperatures).max() <= 1e-3, np.abs(T - o.thermal.temperatures).max()
    assert np.abs(V - o.electrical.potential).max() <= 1e-6
    assert te.thermal.maxT == pytest.approx(o.thermal.maxT, abs=1e-3)
    assert te.thermal.maxT > 300.5                      # the Joule heat really arrived in the thermal solver
    assert te.get_total_current() == pytest.approx(o.electrical.get_total_current(), rel=1e-6)
    for h, g in zip(te.history, o.history):
        assert h["terr"] == pytest.approx(g["terr"], abs=1e-3)
        assert h["verr"] == pytest.approx(g["verr"], rel=1e-3, abs=1e-6)
    te.invalidate()


def test_different_meshes_meta_loop_vs_oracles():
    """thermal and electrical solvers on different meshes (thermoelectric.py:132-138): both exchanges interpolate"""
    pt, pe = make_pair((16, 18, 44), (20, 22, 52))
    te, o, n, no = run_both(pt, pe, 2, 3)
    assert n == no == 2
    T, V = te.thermal.outTemperature(), te.electrical.outVoltage()
    assert np.abs(T - o.thermal.temperatures).max() <= 1e-3
    assert np.abs(V - o.electrical.potential).max() <= 1e-6
    te.invalidate()


@pytest.mark.parametrize("nt,ne", [((9, 8, 12), (9, 8, 12)), ((9, 8, 12), (7, 12, 14)), ((6, 5, 10), (13, 12, 17))])
def test_temperature_exchange_bit_exact(nt, ne):
    """T_elem = interpolateLinear(T) at the element midpoints selects exactly the oracle's conductivities"""
    rng = np.random.default_rng(20261017)
    pt, pe = make_pair(nt, ne, "102", "021")
    T = 300. + 60. * rng.random(pt.N)
    ft, fe = DeviceFem(0), DeviceFem(0)
    ft.set_mesh(pt.axes, pt.strides)
    ft.set_field(T)
    fe.set_mesh(pe.axes, pe.strides)
    fe.set_materials(pe.elem_mat, pe.T0, pe.dT, pe.tab_lat, pe.tab_vert)
    fe.take_temperature_from(ft)
    fe.update_conductivity_shockley()
    cond = fe.get_elem(L.ELEM_COND)
    mt, me = orc.Mesh(*pt.axes, pt.order), orc.Mesh(*pe.axes, pe.order)
    Te = orc.interp_linear(mt.axes, mt.ns, T, [orc.midpoints(a) for a in me.axes], me.es, me.E)
    oe = oracle_shockley(pe, algorithm="cholesky")
    oe.elem_junc[:] = 0
    oe.Te = Te
    oe.load_conductivity()
    assert np.array_equal(cond, oe.conds)
    if nt == ne:   # same mesh: the midpoint value is the mean of the 8 corners up to rounding
        ng = np.broadcast_to(pt.node_index_grid(), pt.n)
        T3 = T[ng]
        mean8 = sum(T3[a:pt.n[0] - 1 + a, b:pt.n[1] - 1 + b, c:pt.n[2] - 1 + c] for a in (0, 1) for b in (0, 1) for c in (0, 1)) / 8
        assert np.abs(Te[np.broadcast_to(pe.elem_index_grid(), mean8.shape)] - mean8).max() <= 1e-9
    ft.close(); fe.close()


def test_heat_without_electrical_solution_is_an_error():
    pt, pe = make_pair((8, 8, 12), (8, 8, 12))
    te = ThermoElectric3D("te")
    te.thermal.problem, te.electrical.problem = pt, pe
    te.initialize()
    with pytest.raises(L.BadInput):
        te.thermal.compute(1)      # NoValue("heat density") in the reference (electr3d.cpp:539)
    te.invalidate()


@pytest.mark.parametrize("order_src,order_dst,layout", [("012", "201", "abi"), ("210", "012", "vertical-minor"), ("120", "120", "abi")])
def test_provider_on_foreign_mesh(order_src, order_dst, layout):
    """outTemperature(mesh) with linear interpolation (getTemperatures, therm3d.cpp:387-395 ->
    RectilinearMesh3D::interpolateLinear, rectilinear3d.hpp:802-845): bit-equal to the oracle's restatement, including
    points outside the source mesh (constant continuation) and points that coincide with source nodes"""
    from oracle import oracle as orc
    from plask_b200.solvers import Static3D
    p = cf.config_B((14, 16, 40), order=order_src)
    s = Static3D("prov")
    s.problem = p
    s.layout = layout
    s.iterative.maxerr = 1e-10
    s.compute(1)
    T = s.outTemperature()
    rng = np.random.default_rng(4)
    axes = []
    for a, m in zip(p.axes, (9, 11, 23)):
        lo, hi = a[0] - 0.1 * (a[-1] - a[0]), a[-1] + 0.1 * (a[-1] - a[0])
        pts = np.sort(np.concatenate([rng.uniform(lo, hi, size=m - 3), a[[0, len(a) // 2, -1]]]))
        axes.append(pts)
    n = tuple(len(a) for a in axes)
    dst_strides = cf.strides_for(n, order_dst)[0]
    ref = orc.interp_linear(p.axes, p.strides, T, axes, dst_strides, int(np.prod(n)))
    got = s.outTemperature((axes, order_dst))
    assert np.array_equal(got, ref)
    assert got.min() >= T.min() - 1e-9 and got.max() <= T.max() + 1e-9      # interpolation::trilinear rounds at 1e-13
    s.invalidate()


def test_no_heat_outside_the_electrical_domain():
    """thermal mesh larger than the electrical one (heat sink / substrate margins in ThermoElectric3D): getHeatDensity is 0
    outside the bounding box of the electrical geometry (electr3d.cpp:545-548), not the constant continuation of its rim"""
    pt, pe = make_pair((26, 28, 52), (20, 22, 52))
    assert pt.axes[0][-1] > pe.axes[0][-1] + 1. and pt.axes[1][0] < pe.axes[1][0] - 1.
    te, o, n, no = run_both(pt, pe, 2, 3)
    assert n == no == 2
    mt = o.thermal.mesh
    mids = [orc.midpoints(a) for a in mt.axes]
    outside = ~((mids[0] >= pe.axes[0][0]) & (mids[0] <= pe.axes[0][-1]))[:, None, None] | \
              ~((mids[1] >= pe.axes[1][0]) & (mids[1] <= pe.axes[1][-1]))[None, :, None] | np.zeros((1, 1, len(mids[2])), dtype=bool)
    assert outside.sum() > 1000 and np.all(o.thermal.heat[mt.elems_grid()[outside]] == 0.)
    assert o.thermal.heat.max() > 0.
    T, V = te.thermal.outTemperature(), te.electrical.outVoltage()
    assert np.abs(T - o.thermal.temperatures).max() <= 1e-3
    assert np.abs(V - o.electrical.potential).max() <= 1e-6
    te.invalidate()


def test_temperature_from_a_masked_thermal_mesh_is_300K_outside():
    """thermal solver with empty-elements='exclude': SafeData substitutes 300 K where the masked mesh has no element
    (getTemperatures, therm3d.cpp:391-392), so the electrical conductivities there are those of 300 K"""
    rng = np.random.default_rng(5)
    pt, pe = make_pair((12, 13, 40), (9, 11, 44), "012", "102")
    T = 320. + 40. * rng.random(pt.N)
    mat = np.array(pt.elem_mat, dtype=np.uint32, copy=True)
    mat[np.asarray(pt.empty) != 0] = L.MAT_EXCLUDED
    assert (mat == L.MAT_EXCLUDED).sum() > 100
    ft, fe = DeviceFem(0), DeviceFem(0)
    ft.set_mesh(pt.axes, pt.strides)
    ft.set_materials(mat, pt.T0, pt.dT, pt.tab_lat, pt.tab_vert)
    ft.set_field(T)
    fe.set_mesh(pe.axes, pe.strides)
    fe.set_materials(pe.elem_mat, pe.T0, pe.dT, pe.tab_lat, pe.tab_vert)
    fe.take_temperature_from(ft)
    fe.update_conductivity_shockley()
    cond = fe.get_elem(L.ELEM_COND)
    ot = oracle_thermal(pt, algorithm="cholesky", included=(np.asarray(pt.empty) == 0).astype(np.uint8))
    # pfem_get_field / the exchange see 0 on the nodes outside the masked mesh (they do not exist in the reference)
    ot._matrix()
    ot.temperatures[:] = np.where(ot._A.active, T, 0.)
    oe = oracle_shockley(pe, algorithm="cholesky")
    oe.elem_junc[:] = 0
    o = orc.ThermoElectric3DOracle(ot, oe)
    o.exchange_temperature()
    assert (oe.Te == 300.).sum() > 50
    oe.load_conductivity()
    assert np.array_equal(cond, oe.conds)
    ft.close(); fe.close()


def test_temperature_dependent_junction_in_the_meta_loop():
    """beta(T), js(T) as callables with electrical.inTemperature connected to the thermal solver on the device
    (electr_python.cpp:103-110 evaluates them at temperature[tidx], the mid-plane element of the junction column,
    electr3d.cpp:261-262): only those elements are read back from the device"""
    pt, pe = make_pair((16, 18, 44), (20, 22, 52))
    beta = lambda T: 11. * (300. / T) ** 0.7        # noqa: E731
    js = lambda T: 1. * np.exp((T - 300.) / 40.)    # noqa: E731
    te = ThermoElectric3D("te-betaT")
    te.thermal.problem, te.electrical.problem = pt, pe
    te.tfreq = 3
    for s in (te.thermal, te.electrical):
        s.iterative.maxerr, s.iterative.maxit = 1e-12, 100000
    te.electrical.beta, te.electrical.js, te.electrical.maxerr = beta, js, pe.maxerr
    te.thermal.maxerr = pt.maxerr
    n = te.compute(max_meta_loops=3)
    ot = oracle_thermal(pt, algorithm="cholesky")
    oe = oracle_shockley(pe, algorithm="cholesky", beta=beta, js=js)
    o = orc.ThermoElectric3DOracle(ot, oe, tfreq=3)
    no = o.compute(max_meta_loops=3)
    assert n == no == 3
    assert o.thermal.maxT > 300.5                      # the junction temperature really moved, so beta(T) != beta(300)
    assert np.abs(te.thermal.outTemperature() - o.thermal.temperatures).max() <= 1e-3
    assert np.abs(te.electrical.outVoltage() - o.electrical.potential).max() <= 1e-6
    assert te.get_total_current() == pytest.approx(o.electrical.get_total_current(), rel=1e-6)
    # against constant parameters the current differs visibly
    oc = oracle_shockley(pe, algorithm="cholesky", beta=11., js=1.)
    oc.compute(3)
    assert abs(oc.get_total_current() - o.electrical.get_total_current()) > 1e-3 * abs(oc.get_total_current())
    te.invalidate()


def test_get_elem_temperature_gathers_the_exchanged_field():
    rng = np.random.default_rng(8)
    pt, pe = make_pair((9, 8, 12), (7, 12, 14), "102", "201")
    T = 300. + 60. * rng.random(pt.N)
    ft, fe = DeviceFem(0), DeviceFem(0)
    ft.set_mesh(pt.axes, pt.strides)
    ft.set_field(T)
    fe.set_layout(L.LAYOUT_VERTICAL_MINOR)
    fe.set_mesh(pe.axes, pe.strides)
    fe.take_temperature_from(ft)
    mt, me = orc.Mesh(*pt.axes, pt.order), orc.Mesh(*pe.axes, pe.order)
    Te = orc.interp_linear(mt.axes, mt.ns, T, [orc.midpoints(a) for a in me.axes], me.es, me.E)
    pick = rng.choice(pe.E, size=200, replace=False)
    assert np.array_equal(fe.get_elem_temperature(pick), Te[pick])
    ft.close(); fe.close()
