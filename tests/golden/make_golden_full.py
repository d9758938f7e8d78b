"""Full-size golden fixtures: BASELINE configs B (Static3D 256^3) and C (Shockley3D 192x192x400) solved on the CPU by the
reference's own NSPCG (oracle/_ref = /root/reference/extlib/nspcg/nspcg.c driven like SparseMatrix::solverhs,
iterative_matrix.hpp:141-339) on the matrix assembled like setMatrix (therm3d.cpp:170-279, electr3d.cpp:281-344).

    python tests/golden/make_golden_full.py B          # tight linear tolerance -> tests/golden/full_B_256.npz
    python tests/golden/make_golden_full.py C [loops]  # first `loops` nonlinear loops -> tests/golden/full_C_192x192x400.npz
    python tests/golden/make_golden_full.py Bdefault   # PLaSK's default parameters (cg+ic, maxerr 1e-6): time to solution only
    python tests/golden/make_golden_full.py Cdefault

Runs in the build container only (needs /root/reference for oracle/_ref), single thread like the reference (NSPCG and
the assembly are serial), about 6 GB and tens of minutes per case.  The full fields (134 MB / 118 MB) cannot be committed;
the fixture keeps a strided sample of the nodes (every `STEP`-th node plane along each axis plus the central column and the
hottest / extreme nodes), the extrema, the loop history and the iteration counts.  tests/test_golden_full.py compares the
CUDA path with these samples at the same full size (north star: <= 1e-3 K, <= 1e-6 V).
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import oracle_shockley, oracle_thermal  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from plask_b200 import configs as cf  # noqa: E402

STEP = 8


def sample_nodes(p, field, extra=()):
    """node numbers of the committed sample: every STEP-th plane of each axis (+ the last), the central column, extras"""
    n = p.n
    idx = [np.unique(np.concatenate([np.arange(0, k, STEP), [k - 1]])) for k in n]
    ng = p.node_index_grid()
    ng = np.broadcast_to(ng, n)
    lattice = ng[np.ix_(*idx)].ravel()
    column = ng[n[0] // 2, n[1] // 2, :].ravel()
    nodes = np.unique(np.concatenate([lattice, column, np.asarray(extra, dtype=np.int64)]))
    return nodes.astype(np.int64), field[nodes].copy()


def log(path, rec):
    with open(path, "a") as f:
        f.write(json.dumps(rec) + "\n")
    print(rec, flush=True)


def case_B(tight):
    p = cf.config_B(256)
    tag = "B_256" if tight else "B_256_default"
    logf = os.path.join(ROOT, "profiles", f"r02_cpu_full_{tag}.jsonl")
    open(logf, "w").close()
    kw = dict(itmaxerr=1e-10, maxit=20000) if tight else dict(itmaxerr=1e-6, maxit=20000)
    o = oracle_thermal(p, algorithm="iterative", precond="ic", **kw)
    t0 = time.perf_counter()
    # the loop of compute() one step at a time, so that every loop is logged as it finishes
    loops = 0
    while True:
        err = o.compute(1)
        h = o.history[-1]
        loops += 1
        log(logf, dict(case=tag, loop=loops, err=h["err"], maxT=h["maxT"], iters=h["iters"], lin_err=h["lin_err"],
                       converged=bool(o.converged), elapsed_s=time.perf_counter() - t0,
                       assembly_s=o.timing["assembly"], solve_s=o.timing["solve"]))
        if o.converged and err <= p.maxerr:
            break
        if loops >= 50:
            break
    T = o.temperatures
    total = time.perf_counter() - t0
    iters = int(sum(h["iters"] for h in o.history))
    log(logf, dict(case=tag, done=True, loops=loops, total_iters=iters, total_s=total, assembly_s=o.timing["assembly"],
                   solve_s=o.timing["solve"], N=p.N, dof_iter_per_s=p.N * iters / o.timing["solve"], cores=1,
                   host=os.uname().nodename, cpu="Intel Xeon (Sapphire Rapids class, KVM guest), 8 vCPU, 1 used"))
    if tight:
        nodes, vals = sample_nodes(p, T, extra=[int(np.argmax(T)), int(np.argmin(T))])
        np.savez_compressed(os.path.join(HERE, "full_B_256.npz"), nodes=nodes, T=vals, maxT=T.max(), minT=T.min(),
                            argmax=int(np.argmax(T)), loops=loops, loop_err=np.array([h["err"] for h in o.history]),
                            loop_maxT=np.array([h["maxT"] for h in o.history]), iters=np.array([h["iters"] for h in o.history]),
                            n=np.array(p.n), order=p.order, itmaxerr=kw["itmaxerr"], mean=T.mean(), l2=np.sqrt((T * T).mean()),
                            precond="ic", total_s=total)


def case_C(tight, loops_cap):
    p = cf.config_C()
    tag = "C_192x192x400" if tight else "C_192x192x400_default"
    logf = os.path.join(ROOT, "profiles", f"r02_cpu_full_{tag}.jsonl")
    open(logf, "w").close()
    kw = dict(itmaxerr=1e-10, maxit=50000) if tight else dict(itmaxerr=1e-6, maxit=50000)
    o = oracle_shockley(p, algorithm="iterative", precond="ic", **kw)
    t0 = time.perf_counter()
    # ONE compute(loops_cap) call: compute() brackets its loops with loadConductivity / saveConductivity (electr3d.cpp:378,435),
    # so k calls of compute(1) are not the same computation as compute(k)
    o.compute(loops_cap)
    loops = len(o.history)
    for h in o.history:
        log(logf, dict(case=tag, loop=h["loop"], err=h["err"], mcur=h["mcur"], iters=h["iters"]))
    snaps = {loops: o.potential.copy()}
    total = time.perf_counter() - t0
    iters = int(sum(h["iters"] for h in o.history))
    log(logf, dict(case=tag, done=True, loops=loops, total_iters=iters, total_s=total, N=p.N, cores=1,
                   total_current_mA=o.get_total_current(), host=os.uname().nodename))
    if tight:
        V = o.potential
        out = {}
        nodes = None
        for k, v in snaps.items():
            nodes, vals = sample_nodes(p, v, extra=[int(np.argmax(V)), int(np.argmin(V))])
            out[f"V_loop{k}"] = vals
        np.savez_compressed(os.path.join(HERE, "full_C_192x192x400.npz"), nodes=nodes, loops=loops,
                            loop_err=np.array([h["err"] for h in o.history]), loop_mcur=np.array([h["mcur"] for h in o.history]),
                            iters=np.array([h["iters"] for h in o.history]), Vmin=V.min(), Vmax=V.max(),
                            total_current=o.get_total_current(), n=np.array(p.n), order=p.order, itmaxerr=kw["itmaxerr"],
                            junction_conductivity=o.junction_conductivity[::97].copy(), precond="ic", total_s=total, **out)


if __name__ == "__main__":
    orc.build(ref=True, quiet=True)
    assert orc.ref_available(), "oracle/_ref must be built (needs /root/reference)"
    which = sys.argv[1] if len(sys.argv) > 1 else "B"
    if which == "B":
        case_B(True)
    elif which == "Bdefault":
        case_B(False)
    elif which == "C":
        case_C(True, int(sys.argv[2]) if len(sys.argv) > 2 else 4)
    elif which == "Cdefault":
        case_C(False, int(sys.argv[2]) if len(sys.argv) > 2 else 200)
    else:
        raise SystemExit("usage: make_golden_full.py B|Bdefault|C|Cdefault [loops]")
