"""BASELINE configs[0] at its full size on the CPU: Static3D 64^3 GaAs/AlGaAs stack with uniform heat, algorithm = cholesky
(DpbMatrix: LAPACK dpbtrf('L') + dpbtrs, cholesky_matrix.hpp:90-111, band kd = n_minor (n_medium + 1) + 1 from
fem_solver.hpp:220-229 — 8.7 GB of band storage) against algorithm = iterative (the reference's NSPCG cg+ic with PLaSK's
defaults, and with a tightened tolerance).  Timings go to profiles/r02_cpu_A_64.jsonl, a strided sample of the Cholesky
solution to tests/golden/full_A_64_cholesky.npz (tests/test_golden_full.py compares the CUDA path with it).

    python tests/golden/make_golden_A64.py        # build container only (oracle/_ref needs /root/reference); ~10 GB, minutes
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
sys.path.insert(0, HERE)

from helpers import oracle_thermal  # noqa: E402
from make_golden_full import sample_nodes  # noqa: E402
from plask_b200 import configs as cf  # noqa: E402

LOG = os.path.join(ROOT, "profiles", "r02_cpu_A_64.jsonl")


def log(rec):
    with open(LOG, "a") as f:
        f.write(json.dumps(rec) + "\n")
    print(rec, flush=True)


def run(p, tag, **kw):
    o = oracle_thermal(p, **kw)
    t0 = time.perf_counter()
    o.compute(0)
    total = time.perf_counter() - t0
    log(dict(case="A_64", leg=tag, loops=len(o.history), iters=[int(h["iters"]) for h in o.history], maxT=o.maxT,
             total_s=total, assembly_s=o.timing["assembly"], solve_s=o.timing["solve"], N=p.N, **kw))
    return o


if __name__ == "__main__":
    open(LOG, "w").close()
    p = cf.config_A(64)
    threads = os.cpu_count()
    try:
        from threadpoolctl import threadpool_info
        threads = max(i["num_threads"] for i in threadpool_info() if i.get("user_api") == "blas")
    except Exception:
        pass
    log(dict(case="A_64", note="LAPACK = scipy's bundled OpenBLAS", blas_threads=threads, host_cpus=os.cpu_count(),
             cpu="Intel Xeon (Sapphire Rapids class, KVM guest), 8 vCPU"))
    it = run(p, "iterative cg+ic, PLaSK defaults (maxerr 1e-6)", algorithm="iterative", precond="ic", itmaxerr=1e-6, maxit=1000)
    it10 = run(p, "iterative cg+ic, maxerr 1e-10", algorithm="iterative", precond="ic", itmaxerr=1e-10, maxit=5000)
    ch = run(p, "cholesky (dpbtrf + dpbtrs, band storage 8.7 GB)", algorithm="cholesky")
    d_def = float(np.abs(it.temperatures - ch.temperatures).max())
    d_10 = float(np.abs(it10.temperatures - ch.temperatures).max())
    log(dict(case="A_64", max_abs_iterative_default_minus_cholesky_K=d_def, max_abs_iterative_tight_minus_cholesky_K=d_10))
    T = ch.temperatures
    nodes, vals = sample_nodes(p, T, extra=[int(np.argmax(T)), int(np.argmin(T))])
    np.savez_compressed(os.path.join(HERE, "full_A_64_cholesky.npz"), nodes=nodes, values=vals, maxT=T.max(), minT=T.min(),
                        argmax=int(np.argmax(T)), loops=len(ch.history), errs=np.array([h["err"] for h in ch.history]),
                        n=np.array(p.n), iterative_default_minus_cholesky=d_def, iterative_tight_minus_cholesky=d_10)
    print("wrote full_A_64_cholesky.npz,", nodes.size, "nodes")
