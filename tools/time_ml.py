"""Multilevel preconditioner at full size: iteration counts and time to solution of the BASELINE configs with
jac / ljac / mlj, and the per-iteration cost.   usage: python tools/time_ml.py [B|C] [n ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from plask_b200 import configs as cf  # noqa: E402
from plask_b200.solvers import Shockley3D, Static3D  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "B"
pres = os.environ.get("PRES", "ljac,mlj").split(",")
if which == "B":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    p = cf.config_B(n, order="012")
    for pre in pres:
        s = Static3D("B")
        s.problem = p
        s.iterative.preconditioner = pre
        s.iterative.maxerr = float(os.environ.get("TOL", "1e-8"))
        s.iterative.maxit = 100000
        t0 = time.perf_counter()
        s.compute(0)
        t1 = time.perf_counter()
        st = s.stats
        print(json.dumps(dict(config="B", n=n, precond=pre, loops=st["outer_loops"], iters=st["lin_iters"], device_ms=st["t_solve_ms"],
                              wall_s=t1 - t0, maxT=st["maxval"], relres=st["lin_relres"], launches=st["kernel_launches"],
                              ms_per_iter=st["t_solve_ms"] / max(st["lin_iters"], 1))), flush=True)
        s.invalidate()
else:
    n = tuple(int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (192, 192, 400)
    p = cf.config_C(n, order=os.environ.get("ORDER", "optimal"))
    for pre in pres:
        e = Shockley3D("C")
        e.problem = p
        e.iterative.preconditioner = pre
        e.iterative.maxerr = float(os.environ.get("TOL", "1e-8"))
        e.iterative.maxit = 200000
        t0 = time.perf_counter()
        e.compute(int(os.environ.get("LOOPS", "0")))
        t1 = time.perf_counter()
        st = e.stats
        print(json.dumps(dict(config="C", n=n, order=p.order, precond=pre, loops=st["outer_loops"], iters=st["lin_iters"], device_ms=st["t_solve_ms"],
                              wall_s=t1 - t0, err=st["err"], relres=st["lin_relres"], current_mA=e.get_total_current(),
                              ms_per_iter=st["t_solve_ms"] / max(st["lin_iters"], 1))), flush=True)
        e.invalidate()
