"""GPU parity of the boundary conditions of the 2nd / 3rd kind and radiation (setBoundaries
therm3d.cpp:140-168, used in setMatrix :242-268) against the oracle, through the C ABI: operator / load vector /
diagonal against the assembled matrix, nonlinear solves against Cholesky, in the verbatim and the corrected form."""
import numpy as np
import pytest

from helpers import face_nodes, oracle_mesh, random_problem, slab_problem_1d
from oracle import oracle as orc
from plask_b200 import _lib as L
from plask_b200 import configs as cf
from plask_b200.fem import DeviceFem
from plask_b200.solvers import Static3D

pytestmark = pytest.mark.gpu


def _conditions(p, rng=None):
    """conditions on all six sides, overlapping at the edges (first one wins), values varying from node to node"""
    rng = rng or np.random.default_rng(17)
    top, bot = face_nodes(p, 2, -1), face_nodes(p, 2, 0)
    left, right = face_nodes(p, 0, 0), face_nodes(p, 0, -1)
    back, front = face_nodes(p, 1, 0), face_nodes(p, 1, -1)
    conv = [(top, 4.0e4, 310.), (left, 9.0e4, 295.)]
    flux = [(right, -3.0e5), (bot[: bot.size // 2], 1.0e5)]
    rad = [(front, 0.85, 285.), (back, 0.3, 330.), (top, 0.5, 300.)]
    return dict(heatflux=flux, convection=conv, radiation=rad)


def _oracle(p, conds, quirk, **kw):
    m = oracle_mesh(p)
    tb = orc.Tables(p.T0, p.dT, p.tab_lat, p.tab_vert)
    b = orc.BoundaryTerms(p.N, **conds)
    kw.setdefault("algorithm", "cholesky")
    return orc.Static3DOracle(m, p.elem_mat, tb, p.bc_nodes, p.bc_values, heat=p.heat, inittemp=p.inittemp,
                              maxerr=p.maxerr, boundaries=b, quirk=quirk, **kw)


@pytest.mark.parametrize("quirk", [True, False])
@pytest.mark.parametrize("order", ["012", "120", "201"])
def test_operator_rhs_diag_with_boundary_terms(order, quirk):
    p = random_problem((9, 7, 11), order)
    rng = np.random.default_rng(5)
    T = rng.uniform(280., 420., size=p.N)
    conds = _conditions(p)
    s = _oracle(p, conds, quirk, algorithm="iterative")
    s.temperatures[:] = T
    A = orc.Sparse14(s.mesh)
    B = np.zeros(p.N)
    s.set_matrix(A, B)
    f = DeviceFem(0)
    f.set_mesh(p.axes, p.strides)
    f.set_materials(p.elem_mat, p.T0, p.dT, p.tab_lat, p.tab_vert)
    f.set_field(T)
    f.set_dirichlet(p.bc_nodes, p.bc_values)
    f.set_source(p.heat)
    f.set_boundary(verbatim=quirk, **conds)
    f.update_conductivity_thermal()
    v = rng.standard_normal(p.N)
    q_ref = A.mult(v)
    scale = np.abs(A.data[:p.N]).max() * np.abs(v).max()
    for variant in (1, 2, 0, 3):
        q = f.apply(v, variant=variant)
        assert np.abs(q - q_ref).max() <= 2e-14 * scale, (variant, np.abs(q - q_ref).max() / scale)
    b = f.get_rhs()
    assert np.abs(b - B).max() <= 1e-13 * np.abs(B).max()
    d = f.get_diag()
    assert np.abs(d - A.data[:p.N]).max() <= 1e-14 * np.abs(A.data[:p.N]).max()
    # removing the conditions restores the plain system
    f.set_boundary()
    s0 = _oracle(p, {}, quirk, algorithm="iterative")
    s0.boundaries = None
    s0.temperatures[:] = T
    A0, B0 = orc.Sparse14(s0.mesh), np.zeros(p.N)
    s0.set_matrix(A0, B0)
    assert np.abs(f.get_rhs() - B0).max() <= 1e-13 * np.abs(B0).max()
    assert np.abs(f.apply(v) - A0.mult(v)).max() <= 2e-14 * scale
    f.close()


@pytest.mark.parametrize("precond", ["jac", "ljac"])
@pytest.mark.parametrize("quirk", [True, False])
@pytest.mark.parametrize("order", ["012", "210"])
def test_nonlinear_solve_vs_cholesky(order, quirk, precond):
    p = cf.config_B((14, 16, 40), order=order)
    conds = _conditions(p)
    o = _oracle(p, conds, quirk)
    o.compute(0)
    s = Static3D("bc")
    s.problem = p
    s.iterative.preconditioner = precond
    s.heatflux_boundary, s.convection_boundary, s.radiation_boundary = conds["heatflux"], conds["convection"], conds["radiation"]
    s.boundary_verbatim = quirk
    s.iterative.maxerr = 1e-11
    s.iterative.maxit = 100000
    s.compute(0)
    assert s.stats["outer_loops"] == len(o.history)
    dT = np.abs(s.outTemperature() - o.temperatures).max()
    assert dT <= 1e-3, dT          # north-star tolerance; observed ~1e-7
    assert s.iterative.err <= 1e-8
    s.invalidate()


def test_convection_only_no_dirichlet():
    """no Dirichlet node at all: the stiffness part alone is singular, only the convection matrix makes the system
    definite — exercises the hand-over of the alpha / breakdown step from k_fpcg to k_surf_iter"""
    p = cf.config_A(12)
    p.bc_nodes = np.zeros(0, dtype=np.uintp)
    p.bc_values = np.zeros(0)
    conds = dict(convection=[(face_nodes(p, 2, 0), 2.0e5, 300.), (face_nodes(p, 0, -1), 5.0e4, 320.)])
    for quirk, precond in ((True, "jac"), (False, "jac"), (False, "ljac")):
        o = _oracle(p, conds, quirk)
        o.compute(0)
        s = Static3D("conv")
        s.problem = p
        s.iterative.preconditioner = precond
        s.convection_boundary = conds["convection"]
        s.boundary_verbatim = quirk
        s.iterative.maxerr = 1e-11
        s.iterative.maxit = 100000
        s.compute(0)
        dT = np.abs(s.outTemperature() - o.temperatures).max()
        assert dT <= 1e-3, (quirk, dT)
        s.invalidate()


def test_analytic_convection_slab_corrected():
    k, h, Ta, T0 = 40., 2.0e5, 350., 300.
    p = slab_problem_1d(n=(40, 9, 33), H=12., k=k, T0=T0)
    s = Static3D("slab")
    s.problem = p
    s.inittemp = T0
    s.convection_boundary = [(face_nodes(p, 2, -1), h, Ta)]
    s.boundary_verbatim = False
    s.iterative.maxerr = 1e-12
    s.compute(1)
    z = p.axes[2] * 1e-6
    TH = (k * T0 / z[-1] + h * Ta) / (k / z[-1] + h)
    T = np.broadcast_to(T0 + (TH - T0) * z / z[-1], p.n).ravel()
    got = s.outTemperature()[np.broadcast_to(p.node_index_grid(), p.n).ravel()]
    assert np.abs(got - T).max() < 1e-7
    s.invalidate()


def test_two_kernel_variants_reject_convection():
    p = cf.config_A(8)
    s = Static3D("v")
    s.problem = p
    s.convection_boundary = [(face_nodes(p, 2, -1), 1e4, 300.)]
    s.variant = 1
    with pytest.raises(L.BadInput):
        s.compute(1)
    s.invalidate()
