"""GPU parity of the matrix-free operator, load vector, diagonal and conductivity update against
the oracle's assembled SparseBandMatrix (iterative_matrix.hpp:347-486) — through the C ABI."""
import numpy as np
import pytest

from helpers import oracle_mesh, random_problem
from oracle import oracle as orc
from plask_b200.fem import DeviceFem

pytestmark = pytest.mark.gpu

ORDERS = ["012", "021", "102", "120", "201", "210"]


def _setup(p):
    f = DeviceFem(0)
    f.set_mesh(p.axes, p.strides)
    f.set_materials(p.elem_mat, p.T0, p.dT, p.tab_lat, p.tab_vert)
    return f


def _oracle_system(p, T):
    m = oracle_mesh(p)
    tb = orc.Tables(p.T0, p.dT, p.tab_lat, p.tab_vert)
    s = orc.Static3DOracle(m, p.elem_mat, tb, p.bc_nodes, p.bc_values, heat=p.heat, algorithm="iterative")
    s.temperatures[:] = T
    A = orc.Sparse14(m)
    B = np.zeros(m.N)
    s.set_matrix(A, B)
    return s, A, B


@pytest.mark.parametrize("order", ORDERS)
@pytest.mark.parametrize("n", [(7, 9, 11), (34, 5, 19)])
def test_operator_rhs_diag_all_orders(order, n):
    p = random_problem(n, order)
    rng = np.random.default_rng(5)
    T = rng.uniform(280., 420., size=p.N)
    s, A, B = _oracle_system(p, T)
    f = _setup(p)
    f.set_field(T)
    f.set_dirichlet(p.bc_nodes, p.bc_values)
    f.set_source(p.heat)
    f.update_conductivity_thermal()
    # conductivities are bit-identical (same table arithmetic, same corner order)
    assert np.array_equal(f.get_elem(0), s.conds)
    # q = A p on random p, both kernel variants
    v = rng.standard_normal(p.N)
    q_ref = A.mult(v)
    scale = np.abs(A.data[:p.N]).max() * np.abs(v).max()
    for variant in (1, 2, 0, 3):
        q = f.apply(v, variant=variant)
        assert np.abs(q - q_ref).max() <= 2e-14 * scale, (variant, np.abs(q - q_ref).max() / scale)
    # load vector after applyBC and the diagonal
    b = f.get_rhs()
    assert np.abs(b - B).max() <= 1e-13 * np.abs(B).max()
    d = f.get_diag()
    assert np.abs(d - A.data[:p.N]).max() <= 1e-14 * np.abs(A.data[:p.N]).max()
    f.close()


@pytest.mark.parametrize("n,order", [((65, 33, 20), "012"), ((20, 70, 37), "201"), ((3, 3, 3), "012"), ((2, 2, 2), "210"),
                                     ((2, 40, 3), "102"), ((130, 4, 6), "210")])
def test_operator_tile_edges(n, order):
    """sizes that straddle the 32/64-wide tiles, degenerate 2-node axes"""
    p = random_problem(n, order, nd_frac=0.03)
    rng = np.random.default_rng(7)
    T = rng.uniform(290., 350., size=p.N)
    s, A, B = _oracle_system(p, T)
    f = _setup(p)
    f.set_field(T)
    f.set_dirichlet(p.bc_nodes, p.bc_values)
    f.set_source(p.heat)
    f.update_conductivity_thermal()
    v = rng.standard_normal(p.N)
    q_ref = A.mult(v)
    scale = np.abs(A.data[:p.N]).max() * np.abs(v).max()
    for variant in (1, 2, 0, 3):
        q = f.apply(v, variant=variant)
        assert np.abs(q - q_ref).max() <= 2e-14 * scale, (variant, np.abs(q - q_ref).max() / scale)
    f.close()


def test_operator_linearity_and_symmetry_large():
    """size-independent properties at a size the oracle does not assemble: <u, A v> == <A u, v>,
    A(const) == 0 on a pure-Neumann problem."""
    from plask_b200 import configs as cf
    p = cf.config_B(96)
    f = _setup(p)
    f.set_field(300.)
    f.set_source(None)
    f.update_conductivity_thermal()
    rng = np.random.default_rng(11)
    u, v = rng.standard_normal(p.N), rng.standard_normal(p.N)
    for variant in (1, 2, 0, 3):
        Au, Av = f.apply(u, variant=variant), f.apply(v, variant=variant)
        assert abs(u @ Av - v @ Au) <= 1e-12 * (np.abs(u @ Av) + np.linalg.norm(Au) * np.linalg.norm(v) * 1e-3)
        ones = f.apply(np.ones(p.N), variant=variant)
        assert np.abs(ones).max() <= 1e-12 * np.abs(Au).max()
    a0, a1, a2, a3 = f.apply(u, variant=0), f.apply(u, variant=1), f.apply(u, variant=2), f.apply(u, variant=3)
    assert np.abs(a0 - a1).max() <= 1e-13 * np.abs(a1).max()
    assert np.abs(a2 - a1).max() <= 1e-13 * np.abs(a1).max()
    assert np.abs(a3 - a1).max() <= 1e-13 * np.abs(a1).max()
    f.close()
