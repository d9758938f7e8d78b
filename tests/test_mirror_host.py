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
