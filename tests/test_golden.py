"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the reference's own
NSPCG + LAPACK Cholesky on reference-assembled matrices, and shockley3d.py's analytic values).

CPU (`not gpu`): the oracle still reproduces them — the fixtures pin the oracle against drift and against the
reference's NSPCG where /root/reference is absent.
GPU: the CUDA path against the same vectors, through the solver mirror / C ABI.
Tolerances: north star — 1e-3 K, 1e-6 V; operator-level quantities to round-off (relative 1e-12)."""
import os

import numpy as np
import pytest

from helpers import oracle_shockley, oracle_thermal, random_problem, shockley3d_reference_problem
from oracle import oracle as orc
from plask_b200 import configs as cf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

THERMAL = {
    "static3d_A16": lambda: cf.config_A(16),
    "static3d_B_14x16x40": lambda: cf.config_B((14, 16, 40)),
    "static3d_rand_11x9x13_order120": lambda: random_problem((11, 9, 13), "120"),
}


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


# ------------------------------------------------------------------------------ CPU: oracle vs fixtures

@pytest.mark.parametrize("name", sorted(THERMAL))
def test_oracle_reproduces_thermal_fixture(name):
    g, p = gold(name), THERMAL[name]()
    assert tuple(g["n"]) == p.n and str(g["order"]) == p.order
    assert np.abs(g["T_cholesky"] - g["T_nspcg"]).max() <= 1e-6      # the two reference solvers agree with each other
    o = oracle_thermal(p, algorithm="cholesky")
    toterr = o.compute(0)
    assert len(o.history) == int(g["loops"])
    assert np.abs(o.temperatures - g["T_cholesky"]).max() <= 1e-9
    assert toterr == pytest.approx(float(g["toterr"]), abs=1e-9)
    f = oracle_thermal(p, algorithm="iterative")
    A, B = orc.Sparse14(f.mesh), np.zeros(f.mesh.N)
    f.set_matrix(A, B)
    assert np.array_equal(f.conds, g["conds0"])
    assert np.array_equal(A.mult(g["p"]), g["q"])                    # same arithmetic, same order -> bit-identical
    assert np.array_equal(B, g["rhs"])
    # the oracle's own Jacobi-PCG (the algorithm the CUDA path implements) reaches the same temperatures
    # (to the north-star tolerance: its plain ||r||/||b|| stop leaves ~1e-4 K on the badly scaled random case)
    j = oracle_thermal(p, algorithm="pcg", itmaxerr=1e-12, maxit=100000)
    j.compute(0)
    assert np.abs(j.temperatures - g["T_nspcg"]).max() <= 1e-3


def test_oracle_reproduces_shockley_fixtures():
    g, p = gold("shockley3d_py"), shockley3d_reference_problem()
    o = oracle_shockley(p, algorithm="cholesky", eps=p.meta["eps"])
    o.compute(25)
    assert np.abs(o.potential - g["V_cholesky"]).max() <= 1e-10
    assert np.abs(g["V_cholesky"] - g["V_nspcg"]).max() <= 1e-9
    assert o.get_total_current() == pytest.approx(float(g["total_current"]), rel=1e-9)
    # shockley3d.py:64-69 analytic values are stored next to the fields
    assert float(g["analytic_current"]) == pytest.approx(1e-9 * 1e6 * (np.exp(10.) - 1))
    g, p = gold("shockley_C_20x22x52"), cf.config_C((20, 22, 52))
    o = oracle_shockley(p, algorithm="cholesky")
    o.compute(8)
    assert np.abs(o.potential - g["V_cholesky"]).max() <= 1e-10
    assert np.abs(o.junction_conductivity - g["junction_conductivity"]).max() <= 1e-9 * np.abs(g["junction_conductivity"]).max()


# ------------------------------------------------------------------------------ GPU: CUDA path vs fixtures

@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(THERMAL))
def test_cuda_thermal_vs_fixture(name):
    from plask_b200.fem import DeviceFem
    from plask_b200.solvers import Static3D
    g, p = gold(name), THERMAL[name]()
    s = Static3D(name)
    s.problem = p
    s.inittemp, s.maxerr = p.inittemp, p.maxerr
    s.iterative.maxerr, s.iterative.maxit = 1e-11, 50000
    toterr = s.compute(0)
    T = s.outTemperature()
    assert np.abs(T - g["T_cholesky"]).max() <= 1e-3                # north-star tolerance (K)
    assert np.abs(T - g["T_nspcg"]).max() <= 1e-3
    assert np.abs(T - g["T_cholesky"]).max() <= 1e-6                # observed: orders of magnitude tighter
    assert s.stats["outer_loops"] == int(g["loops"])
    assert toterr == pytest.approx(float(g["toterr"]), abs=1e-3)
    assert s.iterative.err <= 1e-8                                   # relative residual of the last solve
    flux = s.outHeatFlux()
    assert np.abs(flux - g["flux"]).max() <= 1e-6 * np.abs(g["flux"]).max()
    s.invalidate()
    # operator level: first-loop matrix (conductivities at inittemp)
    f = DeviceFem(0)
    f.set_mesh(p.axes, p.strides)
    f.set_materials(p.elem_mat, p.T0, p.dT, p.tab_lat, p.tab_vert)
    f.set_field(float(p.inittemp))
    f.set_dirichlet(p.bc_nodes, p.bc_values)
    f.set_source(p.heat)
    f.update_conductivity_thermal()
    assert np.array_equal(f.get_elem(0), g["conds0"])                # conductivities: bit-identical
    scale = np.abs(g["diag"]).max() * np.abs(g["p"]).max()
    for variant in (1, 2, 0, 3):
        assert np.abs(f.apply(g["p"], variant) - g["q"]).max() <= 1e-12 * scale, variant
    assert np.abs(f.get_rhs() - g["rhs"]).max() <= 1e-12 * np.abs(g["rhs"]).max()
    assert np.abs(f.get_diag() - g["diag"]).max() <= 1e-13 * np.abs(g["diag"]).max()
    f.close()


@pytest.mark.gpu
def test_cuda_shockley_vs_fixtures():
    from plask_b200.solvers import Shockley3D
    for name, p, loops in (("shockley3d_py", shockley3d_reference_problem(), 25), ("shockley_C_20x22x52", cf.config_C((20, 22, 52)), 8)):
        g = gold(name)
        s = Shockley3D(name)
        s.problem = p
        s.beta, s.js, s.maxerr = p.beta, p.js, p.maxerr
        s.iterative.maxerr, s.iterative.maxit = 1e-12, 50000
        s.compute(loops)
        assert np.abs(s.outVoltage() - g["V_cholesky"]).max() <= 1e-6   # north-star tolerance (V)
        jc = g["junction_conductivity"]
        assert np.abs(s._junc_cond - jc).max() <= 1e-6 * np.abs(jc).max()
        assert s.get_total_current() == pytest.approx(float(g["total_current"]), rel=1e-6)
        heat = s.outHeat()
        assert np.abs(heat - g["heat"]).max() <= 1e-5 * np.abs(g["heat"]).max()
        s.invalidate()


# ------------------------------------------------------------------------------ boundary terms and masked meshes

def _boundary_conditions(p):
    from helpers import face_nodes
    top, bot = face_nodes(p, 2, -1), face_nodes(p, 2, 0)
    return dict(convection=[(top, 4.0e4, 310.), (face_nodes(p, 0, 0), 9.0e4, 295.)],
                heatflux=[(face_nodes(p, 0, -1), -3.0e5), (bot[: bot.size // 2], 1.0e5)],
                radiation=[(face_nodes(p, 1, -1), 0.85, 285.), (face_nodes(p, 1, 0), 0.3, 330.), (top, 0.5, 300.)])


@pytest.mark.parametrize("tag,quirk", [("verbatim", True), ("corrected", False)])
def test_oracle_reproduces_boundary_fixture(tag, quirk):
    g, p = gold("static3d_B_boundary_14x16x40"), cf.config_B((14, 16, 40))
    assert np.abs(g[f"T_cholesky_{tag}"] - g[f"T_nspcg_{tag}"]).max() <= 1e-6     # LAPACK and the reference NSPCG agree
    o = oracle_thermal(p, algorithm="cholesky", boundaries=orc.BoundaryTerms(p.N, **_boundary_conditions(p)), quirk=quirk)
    o.compute(0)
    assert len(o.history) == int(g[f"loops_{tag}"])
    assert np.abs(o.temperatures - g[f"T_cholesky_{tag}"]).max() <= 1e-9


def test_oracle_reproduces_masked_fixtures():
    g, p = gold("shockley3d_py_excluded"), shockley3d_reference_problem()
    o = oracle_shockley(p, algorithm="cholesky", eps=p.meta["eps"], included=(p.empty == 0).astype(np.uint8))
    o.compute(25)
    act = g["masked_nodes"]
    assert np.array_equal(o._matrix().active, act)
    assert np.abs(o.potential - g["V_cholesky"])[act].max() <= 1e-10
    assert o.get_capacitance() == pytest.approx(float(g["capacitance"]), rel=1e-9)
    g, p = gold("static3d_B_excluded_14x16x40"), cf.config_B((14, 16, 40))
    o = oracle_thermal(p, algorithm="cholesky", included=(p.empty == 0).astype(np.uint8))
    o.compute(0)
    assert len(o.history) == int(g["loops"])
    assert np.abs(o.temperatures - g["T_cholesky"])[g["masked_nodes"]].max() <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("precond", ["jac", "ljac"])
@pytest.mark.parametrize("tag,quirk", [("verbatim", True), ("corrected", False)])
def test_cuda_boundary_vs_fixture(tag, quirk, precond):
    from plask_b200.solvers import Static3D
    g, p = gold("static3d_B_boundary_14x16x40"), cf.config_B((14, 16, 40))
    c = _boundary_conditions(p)
    s = Static3D(tag)
    s.problem = p
    s.heatflux_boundary, s.convection_boundary, s.radiation_boundary = c["heatflux"], c["convection"], c["radiation"]
    s.boundary_verbatim = quirk
    s.iterative.preconditioner = precond
    s.iterative.maxerr, s.iterative.maxit = 1e-11, 100000
    s.compute(0)
    T = s.outTemperature()
    assert s.stats["outer_loops"] == int(g[f"loops_{tag}"])
    assert np.abs(T - g[f"T_cholesky_{tag}"]).max() <= 1e-3          # north-star tolerance (K)
    assert np.abs(T - g[f"T_nspcg_{tag}"]).max() <= 1e-3
    s.invalidate()


@pytest.mark.gpu
@pytest.mark.parametrize("precond", ["jac", "ljac"])
def test_cuda_masked_vs_fixtures(precond):
    from plask_b200.solvers import Shockley3D, Static3D
    g, p = gold("shockley3d_py_excluded"), shockley3d_reference_problem()
    e = Shockley3D("excluded")
    e.problem = p
    e.empty_elements = "exclude"
    e.beta, e.js, e.maxerr = p.beta, p.js, p.maxerr
    e.iterative.preconditioner = precond
    e.iterative.maxerr, e.iterative.maxit = 1e-13, 100000
    e.compute(25)
    act = g["masked_nodes"]
    assert np.array_equal(e.masked_nodes(), act)
    assert np.abs(e.outVoltage() - g["V_cholesky"])[act].max() <= 1e-6   # north-star tolerance (V)
    assert e.get_total_current() == pytest.approx(float(g["total_current"]), rel=1e-5)
    assert e.get_total_heat() == pytest.approx(float(g["total_heat"]), rel=1e-5)
    e.invalidate()
    g, p = gold("static3d_B_excluded_14x16x40"), cf.config_B((14, 16, 40))
    s = Static3D("excluded")
    s.problem = p
    s.empty_elements = "exclude"
    s.iterative.preconditioner = precond
    s.iterative.maxerr, s.iterative.maxit = 1e-11, 100000
    s.compute(0)
    assert s.stats["outer_loops"] == int(g["loops"])
    assert np.abs(s.outTemperature() - g["T_cholesky"])[g["masked_nodes"]].max() <= 1e-3
    assert s.maxT == pytest.approx(float(g["maxT"]), abs=1e-6)
    s.invalidate()
