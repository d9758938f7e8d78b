"""Generate the committed golden fixtures of tests/golden/*.npz.

    python tests/golden/make_golden.py

Run in the build container, where /root/reference exists and oracle/_ref (the reference's own NSPCG,
extlib/nspcg/nspcg.c compiled by oracle/Makefile) is built.  PLaSK itself cannot be built or imported here
(SURVEY.md §8c), so the vectors come from
  * the reference's NSPCG driven exactly as SparseMatrix::solverhs drives it (iterative_matrix.hpp:141-339)
    on the matrix assembled like setMatrix does (therm3d.cpp:170-279, electr3d.cpp:281-344)  -> *_nspcg
  * LAPACK dpbtrf/dpbtrs on the same matrix in DpbMatrix storage (cholesky_matrix.hpp:90-111)  -> *_cholesky
  * the analytic values of solvers/electrical/shockley/tests/shockley3d.py:62-83.
The fixtures travel to the GPU box (which has no /root/reference); tests/test_golden.py checks the oracle
(CPU) and the CUDA path (GPU) against them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import oracle_shockley, oracle_thermal, random_problem, shockley3d_reference_problem  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from plask_b200 import configs as cf  # noqa: E402

CASES = {
    "static3d_A16": lambda: cf.config_A(16),
    "static3d_B_14x16x40": lambda: cf.config_B((14, 16, 40)),
    "static3d_rand_11x9x13_order120": lambda: random_problem((11, 9, 13), "120"),
}


def thermal_case(name, p):
    assert orc.ref_available(), "oracle/_ref must be built (needs /root/reference)"
    o = oracle_thermal(p, algorithm="cholesky")
    toterr = o.compute(0)
    r = oracle_thermal(p, algorithm="iterative", precond="ic", itmaxerr=1e-12, maxit=5000)
    r.compute(0)
    # operator / rhs / diagonal of the FIRST loop's matrix (conductivities at inittemp)
    f = oracle_thermal(p, algorithm="iterative")
    A = orc.Sparse14(f.mesh)
    B = np.zeros(f.mesh.N)
    f.set_matrix(A, B)
    rng = np.random.default_rng(20261017)
    pvec = rng.standard_normal(f.mesh.N)
    q = A.mult(pvec)
    diag = A.data[:f.mesh.N].copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"),
                        T_cholesky=o.temperatures, T_nspcg=r.temperatures, toterr=toterr, maxT=o.maxT,
                        loops=len(o.history), loop_err=np.array([h["err"] for h in o.history]),
                        nspcg_loops=len(r.history), conds0=f.conds, p=pvec, q=q, rhs=B, diag=diag,
                        flux=o.heat_fluxes(), n=np.array(p.n), order=p.order)
    print(f"{name}: N={p.N} loops={len(o.history)} maxT={o.maxT:.6f} |T_chol-T_nspcg|={np.abs(o.temperatures - r.temperatures).max():.2e}")


def shockley_cases():
    p = shockley3d_reference_problem()
    # a fixed number of loops on both sides (the loop count to convergence depends on round-off noise)
    o = oracle_shockley(p, algorithm="cholesky", eps=p.meta["eps"])
    o.compute(25)
    r = oracle_shockley(p, algorithm="iterative", precond="ic", itmaxerr=1e-12, maxit=5000, eps=p.meta["eps"])
    r.compute(25)
    S = 1e6
    np.savez_compressed(os.path.join(HERE, "shockley3d_py.npz"), V_cholesky=o.potential, V_nspcg=r.potential,
                        current=o.current, heat=o.heat_density(), total_current=o.get_total_current(),
                        total_heat=o.get_total_heat(), capacitance=o.get_capacitance(), loops=len(o.history),
                        analytic_current=1e-9 * S * 1. * (np.exp(10.) - 1), analytic_capacitance=8.854187817e-6 * 12.9 * S / 0.02,
                        junction_conductivity=o.junction_conductivity, n=np.array(p.n), order=p.order)
    print(f"shockley3d_py: loops={len(o.history)} I={o.get_total_current():.6f} mA (analytic {1e-9 * S * (np.exp(10.) - 1):.6f}) "
          f"|V_chol-V_nspcg|={np.abs(o.potential - r.potential).max():.2e}")
    p = cf.config_C((20, 22, 52))
    o = oracle_shockley(p, algorithm="cholesky")
    o.compute(8)
    np.savez_compressed(os.path.join(HERE, "shockley_C_20x22x52.npz"), V_cholesky=o.potential, current=o.current,
                        heat=o.heat_density(), total_current=o.get_total_current(), loops=len(o.history),
                        loop_err=np.array([h["err"] for h in o.history]), junction_conductivity=o.junction_conductivity,
                        n=np.array(p.n), order=p.order)
    print(f"shockley_C_20x22x52: loops={len(o.history)} I={o.get_total_current():.6e} mA")


def boundary_conditions(p):
    """conditions of the 2nd / 3rd kind and radiation on all six sides of config B (the same set tests/test_golden.py builds)"""
    from helpers import face_nodes
    top, bot = face_nodes(p, 2, -1), face_nodes(p, 2, 0)
    return dict(convection=[(top, 4.0e4, 310.), (face_nodes(p, 0, 0), 9.0e4, 295.)],
                heatflux=[(face_nodes(p, 0, -1), -3.0e5), (bot[: bot.size // 2], 1.0e5)],
                radiation=[(face_nodes(p, 1, -1), 0.85, 285.), (face_nodes(p, 1, 0), 0.3, 330.), (top, 0.5, 300.)])


def boundary_cases():
    """setBoundaries (therm3d.cpp:140-168,242-268), verbatim and corrected: Cholesky and the reference NSPCG on the same matrix"""
    p = cf.config_B((14, 16, 40))
    out = {}
    for tag, quirk in (("verbatim", True), ("corrected", False)):
        b = orc.BoundaryTerms(p.N, **boundary_conditions(p))
        o = oracle_thermal(p, algorithm="cholesky", boundaries=b, quirk=quirk)
        o.compute(0)
        r = oracle_thermal(p, algorithm="iterative", precond="ic", itmaxerr=1e-12, maxit=5000, boundaries=b, quirk=quirk)
        r.compute(0)
        out[f"T_cholesky_{tag}"], out[f"T_nspcg_{tag}"], out[f"loops_{tag}"] = o.temperatures, r.temperatures, len(o.history)
        print(f"static3d_B_boundary {tag}: loops={len(o.history)} maxT={o.maxT:.6f} |T_chol-T_nspcg|={np.abs(o.temperatures - r.temperatures).max():.2e}")
    np.savez_compressed(os.path.join(HERE, "static3d_B_boundary_14x16x40.npz"), n=np.array(p.n), order=p.order, **out)


def masked_cases():
    """empty-elements="exclude": Cholesky on the RectangularMaskedMesh3D numbering (the reference's default Cholesky mesh)"""
    p = shockley3d_reference_problem()
    inc = (p.empty == 0).astype(np.uint8)
    o = oracle_shockley(p, algorithm="cholesky", eps=p.meta["eps"], included=inc)
    o.compute(25)
    np.savez_compressed(os.path.join(HERE, "shockley3d_py_excluded.npz"), V_cholesky=o.potential, masked_nodes=o._matrix().active,
                        total_current=o.get_total_current(), total_heat=o.get_total_heat(), capacitance=o.get_capacitance(),
                        loops=len(o.history), n=np.array(p.n), order=p.order)
    print(f"shockley3d_py_excluded: loops={len(o.history)} I={o.get_total_current():.6f} mA, masked nodes {int(o._matrix().active.sum())} of {p.N}")
    p = cf.config_B((14, 16, 40))
    inc = (p.empty == 0).astype(np.uint8)
    o = oracle_thermal(p, algorithm="cholesky", included=inc)
    o.compute(0)
    np.savez_compressed(os.path.join(HERE, "static3d_B_excluded_14x16x40.npz"), T_cholesky=o.temperatures, masked_nodes=o._matrix().active,
                        loops=len(o.history), maxT=o.maxT, n=np.array(p.n), order=p.order)
    print(f"static3d_B_excluded: loops={len(o.history)} maxT={o.maxT:.6f}, masked nodes {int(o._matrix().active.sum())} of {p.N}")


if __name__ == "__main__":
    orc.build(ref=True, quiet=True)
    which = sys.argv[1:] or ["thermal", "shockley", "boundary", "masked"]
    if "thermal" in which:
        for name, mk in CASES.items():
            thermal_case(name, mk())
    if "shockley" in which:
        shockley_cases()
    if "boundary" in which:
        boundary_cases()
    if "masked" in which:
        masked_cases()
