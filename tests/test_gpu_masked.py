"""GPU parity on masked meshes — empty-elements="exclude" (FemSolverWithMaskedMesh::setupMaskedMesh,
fem_solver.hpp:182-189; the default of the reference's Cholesky path): elements of EMPTY material are marked
PFEM_MAT_EXCLUDED, nodes that touch no kept element drop out.  Pinned by the reference's own
shockley3d.py:71-73 (testComputationsExcluded) and compared with the oracle's Cholesky on the compressed
(RectangularMaskedMesh3D) numbering."""
import numpy as np
import pytest

from helpers import oracle_shockley, oracle_thermal, shockley3d_reference_problem
from plask_b200 import configs as cf
from plask_b200.solvers import Shockley3D, Static3D

pytestmark = pytest.mark.gpu


def _included(p):
    return (np.asarray(p.empty) == 0).astype(np.uint8)


def test_shockley3d_reference_case_excluded():
    p = shockley3d_reference_problem()
    e = Shockley3D("electrical3d")
    e.problem = p
    e.empty_elements = "exclude"
    e.beta, e.js, e.maxerr = 10., 1., 1e-5
    e.iterative.maxerr = 1e-13
    e.iterative.maxit = 100000
    e.compute(1000)
    S = 1e6
    correct_current = 1e-9 * S * 1. * (np.exp(10.) - 1)
    assert abs(e.get_total_current()) == pytest.approx(correct_current, abs=0.5e-3)       # shockley3d.py:64-65,71-73
    assert e.get_total_heat() == pytest.approx(correct_current * 1., abs=0.5e-3)         # :68-69
    assert e.get_capacitance() == pytest.approx(8.854187817e-6 * 12.9 * S / 0.02, abs=0.5e-2)   # :66-67, masked elements only
    act = e.masked_nodes()
    assert act.sum() == 764 and p.N == 1100
    V = e.outVoltage()
    assert np.all(V[~act] == 0.)
    # excluded elements carry no current and no heat, their conductivity reads 0
    excl = np.asarray(p.empty) != 0
    assert np.all(e.outCurrentDensity()[excl] == 0.) and np.all(e.outHeat()[excl] == 0.)
    assert np.all(e.outConductivity()[excl] == 0.)
    e.invalidate()


@pytest.mark.parametrize("order", ["201", "012"])
def test_shockley3d_excluded_vs_masked_cholesky(order):
    p = shockley3d_reference_problem(order=order)
    LOOPS = 12
    o = oracle_shockley(p, algorithm="cholesky", included=_included(p))
    o.compute(LOOPS)
    e = Shockley3D("e")
    e.problem = p
    e.empty_elements = "exclude"
    e.beta, e.js, e.maxerr = p.beta, p.js, p.maxerr
    e.iterative.maxerr = 1e-13
    e.iterative.maxit = 100000
    e.compute(LOOPS)
    act = e.masked_nodes()
    assert np.array_equal(act, o._matrix().active)
    dV = np.abs(e.outVoltage() - o.potential)[act].max()
    assert dV <= 1e-6, dV
    assert e.stats["err"] == pytest.approx(o.history[-1]["err"], rel=1e-3)   # 1e9 S/m conductor against a uS junction: the loop error amplifies 1e-12 V
    e.invalidate()


@pytest.mark.parametrize("order", ["012", "210"])
def test_config_C_small_excluded_vs_masked_cholesky(order):
    p = cf.config_C((20, 22, 52), order=order)
    assert p.empty.sum() > 0
    LOOPS = 8
    o = oracle_shockley(p, algorithm="cholesky", included=_included(p))
    o.compute(LOOPS)
    e = Shockley3D("C")
    e.problem = p
    e.empty_elements = "exclude"
    e.beta, e.js, e.maxerr = p.beta, p.js, p.maxerr
    e.iterative.maxerr = 1e-13
    e.iterative.maxit = 200000
    e.compute(LOOPS)
    act = e.masked_nodes()
    dV = np.abs(e.outVoltage() - o.potential)[act].max()
    assert dV <= 1e-6, dV
    e.invalidate()


@pytest.mark.parametrize("order", ["012", "120"])
def test_config_B_small_excluded_vs_masked_cholesky(order):
    p = cf.config_B((18, 20, 44), order=order)
    assert p.empty.sum() > 0
    o = oracle_thermal(p, algorithm="cholesky", included=_included(p))
    o.compute(0)
    s = Static3D("B")
    s.problem = p
    s.empty_elements = "exclude"
    s.iterative.maxerr = 1e-11
    s.iterative.maxit = 100000
    s.compute(0)
    act = s.masked_nodes()
    assert np.array_equal(act, o._matrix().active)
    assert s.stats["outer_loops"] == len(o.history)
    dT = np.abs(s.outTemperature() - o.temperatures)[act].max()
    assert dT <= 1e-3, dT
    assert s.maxT == pytest.approx(o.maxT, abs=1e-6)
    flux = s.outHeatFlux()
    assert np.all(flux[np.asarray(p.empty) != 0] == 0.)
    ref = o.heat_fluxes()
    keep = np.asarray(p.empty) == 0
    assert np.abs(flux - ref)[keep].max() <= 1e-6 * np.abs(ref).max()
    s.invalidate()


@pytest.mark.parametrize("quirk", [False, True])
def test_excluded_elements_add_no_boundary_terms(quirk):
    """convection / radiation / heat flux on a surface INSIDE the mesh where kept elements meet excluded (air) ones: the
    reference calls setBoundaries only for elements of the masked mesh (therm3d.cpp:186-268), so the air side adds nothing"""
    from helpers import face_nodes, oracle_thermal
    from oracle import oracle as orc
    from plask_b200.solvers import Static3D
    p = cf.config_B((14, 16, 40))
    inc = (p.empty == 0).astype(np.uint8)
    # every node that belongs to both a kept and an excluded element: the surface of the structure inside the mesh
    n = p.n
    eg = np.broadcast_to(p.elem_index_grid(), tuple(k - 1 for k in n))
    kept3 = inc[eg].astype(bool)
    touch_kept, touch_excl = np.zeros(n, dtype=bool), np.zeros(n, dtype=bool)
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                sl = (slice(a, n[0] - 1 + a), slice(b, n[1] - 1 + b), slice(c, n[2] - 1 + c))
                touch_kept[sl] |= kept3
                touch_excl[sl] |= ~kept3
    surf = np.broadcast_to(p.node_index_grid(), n)[touch_kept & touch_excl].astype(np.int64)
    assert surf.size > 50
    conds = dict(convection=[(surf, 6.0e4, 305.)], radiation=[(surf, 0.7, 295.)])
    o = oracle_thermal(p, algorithm="cholesky", included=inc, boundaries=orc.BoundaryTerms(p.N, **conds), quirk=quirk)
    o.compute(0)
    s = Static3D("masked-boundary")
    s.problem = p
    s.empty_elements = "exclude"
    s.convection_boundary, s.radiation_boundary, s.boundary_verbatim = conds["convection"], conds["radiation"], quirk
    s.iterative.maxerr = 1e-11
    s.iterative.maxit = 100000
    s.compute(0)
    act = s.masked_nodes()
    assert s.stats["outer_loops"] == len(o.history)
    assert np.abs(s.outTemperature() - o.temperatures)[act].max() <= 1e-3
    s.invalidate()
    if not quirk:
        # and it matters: with the terms of the excluded elements counted as well (every inner face twice) the field differs
        class AllElements(orc.BoundaryTerms):
            def terms(self, mesh, T, B, quirk, included=None):
                return super().terms(mesh, T, B, quirk, None)
        o2 = oracle_thermal(p, algorithm="cholesky", included=inc, boundaries=AllElements(p.N, **conds), quirk=False)
        o2.compute(0)
        assert np.abs(o2.temperatures - o.temperatures)[act].max() > 0.05
