"""BASELINE configs B and C at their FULL size against the reference's own NSPCG (tests/golden/make_golden_full.py: config B
256^3 = 16.8 M nodes to convergence, config C 192x192x400 = 14.7 M nodes for its first 4 nonlinear loops; cg + ic with the
linear tolerance tightened to 1e-10; 35 / 71 CPU-minutes on one core).  The fixtures hold a strided sample of the nodes (every
8th plane of each axis, the central column, the extrema), the loop history and the iteration counts.
North star: max |dT| <= 1e-3 K, max |dV| <= 1e-6 V."""
import os

import numpy as np
import pytest

from plask_b200 import configs as cf
from plask_b200.solvers import Shockley3D, Static3D

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixture(name):
    path = os.path.join(HERE, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} has not been generated (tests/golden/make_golden_full.py)")
    return np.load(path)


def test_fixtures_are_consistent():
    """CPU: the committed samples describe the configs as plask_b200.configs builds them today"""
    for name, mk in (("full_B_256.npz", lambda: (256, 256, 256)), ("full_C_192x192x400.npz", lambda: (192, 192, 400))):
        path = os.path.join(HERE, name)
        if not os.path.exists(path):
            continue
        g = np.load(path)
        n = mk()
        assert tuple(int(v) for v in g["n"]) == n
        assert str(g["order"]) == cf.optimal_order(n)
        assert g["nodes"].max() < n[0] * n[1] * n[2] and g["nodes"].size > 30000
        assert float(g["itmaxerr"]) <= 1e-10 and str(g["precond"]) == "ic"


@pytest.mark.gpu
@pytest.mark.parametrize("precond", ["mlj", "ljac"])
def test_config_B_256_vs_reference_nspcg(precond):
    g = _fixture("full_B_256.npz")
    p = cf.config_B(256)
    assert p.order == str(g["order"])
    s = Static3D("B")
    s.problem = p
    s.iterative.preconditioner = precond
    s.iterative.maxerr = 1e-10 if precond == "mlj" else 1e-9
    s.iterative.maxit = 200000
    s.compute(0)
    T = s.outTemperature()
    assert s.stats["outer_loops"] == int(g["loops"])          # same Picard loop count as the reference
    d = np.abs(T[g["nodes"]] - g["T"])
    assert d.max() <= 1e-3, (d.max(), int(g["nodes"][np.argmax(d)]))
    assert abs(float(T.max()) - float(g["maxT"])) <= 1e-3
    assert abs(float(T.min()) - float(g["minT"])) <= 1e-3
    assert abs(float(T.mean()) - float(g["mean"])) <= 1e-4
    s.invalidate()


@pytest.mark.gpu
def test_config_B_256_first_loop_all_preconditioners():
    """the three iteration kernels (fused Jacobi, line-Jacobi, multilevel) give the same first-loop field at full size"""
    p = cf.config_B(256)
    T = {}
    for pre in ("jac", "ljac", "mlj"):
        s = Static3D("B1")
        s.problem = p
        s.iterative.preconditioner = pre
        s.iterative.maxerr = 1e-10
        s.iterative.maxit = 200000
        s.compute(1)
        T[pre] = s.outTemperature().copy()
        assert s.iterative.converged
        s.invalidate()
    assert np.abs(T["jac"] - T["ljac"]).max() <= 1e-4
    assert np.abs(T["mlj"] - T["ljac"]).max() <= 1e-4


@pytest.mark.gpu
def test_config_C_full_vs_reference_nspcg():
    g = _fixture("full_C_192x192x400.npz")
    loops = int(g["loops"])
    p = cf.config_C()
    assert p.order == str(g["order"])
    e = Shockley3D("C")
    e.problem = p
    e.beta, e.js, e.maxerr = p.beta, p.js, p.maxerr
    e.iterative.preconditioner = "mlj"
    e.iterative.maxerr = 1e-11
    e.iterative.maxit = 200000
    e.compute(loops)
    V = e.outVoltage()
    d = np.abs(V[g["nodes"]] - g[f"V_loop{loops}"])
    assert d.max() <= 1e-6, (d.max(), int(g["nodes"][np.argmax(d)]))
    assert abs(e.get_total_current() - float(g["total_current"])) <= 1e-6 * abs(float(g["total_current"])) + 1e-9
    assert abs(float(V.min()) - float(g["Vmin"])) <= 1e-6 and abs(float(V.max()) - float(g["Vmax"])) <= 1e-6
    s = e.stats
    assert abs(s["err"] - float(g["loop_err"][-1])) <= 1e-3 * float(g["loop_err"][-1])      # loop error of the last loop [%]
    e.invalidate()


@pytest.mark.gpu
@pytest.mark.parametrize("precond", ["jac", "ljac", "mlj"])
def test_config_A_64_vs_cholesky(precond):
    """BASELINE configs[0] at its full size against the reference's DIRECT algorithm (DpbMatrix: LAPACK dpbtrf + dpbtrs on 8.7 GB
    of band storage, cholesky_matrix.hpp:90-111; tests/golden/make_golden_A64.py, timings in profiles/r02_cpu_A_64.jsonl)"""
    g = _fixture("full_A_64_cholesky.npz")
    p = cf.config_A(64)
    assert tuple(int(v) for v in g["n"]) == p.n
    s = Static3D("A")
    s.problem = p
    s.iterative.preconditioner = precond
    s.iterative.maxerr = 1e-11
    s.iterative.maxit = 200000
    s.compute(0)
    T = s.outTemperature()
    assert s.stats["outer_loops"] == int(g["loops"])
    d = np.abs(T[g["nodes"]] - g["values"])
    assert d.max() <= 1e-3, d.max()
    assert d.max() <= 1e-5          # in fact: the Cholesky and the PCG solutions agree far below the north-star tolerance
    assert abs(s.maxT - float(g["maxT"])) <= 1e-5
    s.invalidate()
