"""Full-size runs of the BASELINE.json configurations that bench.py does not time (they are parity-test cases, not the
bench workload): prints one JSON line per run with loops, PCG iterations, device time and DOF*iter/s.

    python tools/run_configs.py --config C                       # Shockley3D 192x192x400, one GPU
    python tools/run_configs.py --config B --boundary            # config B with convection/radiation on the outer faces
    torchrun --nproc-per-node 8 tools/run_configs.py --config D  # ThermoElectric3D 464^3 (~100 M nodes) in slab mode
    torchrun --nproc-per-node 8 tools/run_configs.py --config E  # Static3D weak scaling case, ~25 M nodes per GPU
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from plask_b200 import configs as cf  # noqa: E402
from plask_b200.solvers import Shockley3D, Static3D, ThermoElectric3D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C")
    ap.add_argument("--mesh", dest="n", type=int, nargs="*", default=None)
    ap.add_argument("--boundary", action="store_true")
    ap.add_argument("--lin-tol", type=float, default=1e-8)
    ap.add_argument("--meta-loops", type=int, default=100)
    ap.add_argument("--precond", default="jac", help="jac | ljac (line-Jacobi along the vertical axis) | mlj (multilevel line preconditioner)")
    ap.add_argument("--order", default="optimal")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    allgather = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")

        def allgather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out

    def slab_of(n0):
        lo, hi, own_lo, own_hi = cf.slab_local(n0, rank, world, 16 if a.precond == "mlj" else 1)   # mlj: slab boundaries at multiples of 16 planes
        return (lo, hi), dict(rank=rank, nranks=world, own_lo=own_lo, own_hi=own_hi, allgather=allgather)

    def tune(s):
        s.device = local
        s.iterative.maxerr = a.lin_tol
        s.iterative.maxit = 200000
        s.iterative.preconditioner = a.precond

    out = dict(config=a.config, n_gpus=world, preconditioner=a.precond)
    t0 = time.perf_counter()
    if a.config == "C":
        n = tuple(a.n) if a.n else (192, 192, 400)
        p = cf.config_C(n, order=a.order)
        e = Shockley3D("C")
        tune(e)
        if world > 1:   # the host cuts a lateral axis (the junction and the vertical lines stay inside every slab)
            axis = cf.slab_axis(p, need_vertical_inside=True)
            q, own_lo, own_hi, _ = cf.slab_problem(p, rank, world, axis=axis, align=16 if a.precond == "mlj" else 1)
            e.problem = q
            e.slab = dict(rank=rank, nranks=world, own_lo=own_lo, own_hi=own_hi, allgather=allgather)
            e.beta, e.js, e.maxerr = p.beta, p.js, p.maxerr
            out["slab_axis"] = axis
        else:
            e.problem = p
        t1 = time.perf_counter()
        e.compute(0)
        t2 = time.perf_counter()
        st = e.stats
        I = e.get_total_current() if world == 1 else float("nan")
        out.update(mesh=n, dof=p.N, order=p.order, outer_loops=st["outer_loops"], pcg_iterations=st["lin_iters"],
                   device_ms=st["t_solve_ms"], wall_s=t2 - t1, setup_s=t1 - t0, err_percent=st["err"], lin_relres=st["lin_relres"],
                   max_j_kAcm2=st["maxval"], total_current_mA=I, dof_iter_per_s=p.N * st["lin_iters"] / (st["t_solve_ms"] * 1e-3))
    elif a.config == "B":
        n = tuple(a.n) if a.n else (256, 256, 256)
        p = cf.config_B(n, order=a.order)
        s = Static3D("B")
        tune(s)
        s.problem = p
        if a.boundary:
            from helpers import face_nodes
            s.convection_boundary = [(face_nodes(p, 2, -1), 2.0e3, 300.), (face_nodes(p, 0, 0), 1.0e3, 300.), (face_nodes(p, 0, -1), 1.0e3, 300.)]
            s.radiation_boundary = [(face_nodes(p, 1, 0), 0.9, 300.), (face_nodes(p, 1, -1), 0.9, 300.)]
            s.boundary_verbatim = False
        t1 = time.perf_counter()
        s.compute(0)
        t2 = time.perf_counter()
        st = s.stats
        out.update(mesh=n, dof=p.N, order=p.order, boundary=bool(a.boundary), outer_loops=st["outer_loops"], pcg_iterations=st["lin_iters"],
                   device_ms=st["t_solve_ms"], wall_s=t2 - t1, setup_s=t1 - t0, maxT=st["maxval"], lin_relres=st["lin_relres"],
                   kernel_launches=st["kernel_launches"], dof_iter_per_s=p.N * st["lin_iters"] / (st["t_solve_ms"] * 1e-3))
    elif a.config == "D":
        n = tuple(a.n) if a.n else (464, 464, 464)
        rows, slab = slab_of(n[0])
        pt = cf.config_B(n, order="012", rows0=rows if world > 1 else None)
        pe = cf.config_C(n, order="012", rows0=rows if world > 1 else None)
        te = ThermoElectric3D("D")
        for s, p in ((te.thermal, pt), (te.electrical, pe)):
            tune(s)
            s.problem = p
            if world > 1:
                s.slab = slab
        t1 = time.perf_counter()
        loops = te.compute(invalidate=False, max_meta_loops=a.meta_loops)
        t2 = time.perf_counter()
        N = n[0] * n[1] * n[2]
        out.update(mesh=n, dof=N, order="012", meta_loops=loops, wall_s=t2 - t1, setup_s=t1 - t0,
                   maxT=te.thermal.maxT, max_j_kAcm2=te.electrical.stats["maxval"], history=te.history[-3:],
                   thermal_last=dict(pcg=te.thermal.stats["lin_iters"], ms=te.thermal.stats["t_solve_ms"], relres=te.thermal.stats["lin_relres"]),
                   electrical_last=dict(pcg=te.electrical.stats["lin_iters"], ms=te.electrical.stats["t_solve_ms"], relres=te.electrical.stats["lin_relres"]))
    elif a.config == "E":
        nz, nxy = (a.n + [96, 512])[:2] if a.n else (96, 512)
        n = (nz * world, nxy, nxy)
        rows, slab = slab_of(n[0])
        # layered stack: every slab has the same structure; heat scaled with the stack height so that the temperature
        # rise stays inside the k(T) tables (config A's 1e15 W/m3 is quoted for 63 layers)
        pg = cf.config_A(n, order="012", rows0=rows if world > 1 else None, heat=1e14 * (63. / (nxy - 1)) ** 2)
        s = Static3D("E")
        tune(s)
        s.problem = pg
        if world > 1:
            s.slab = slab
        t1 = time.perf_counter()
        s.compute(0)
        t2 = time.perf_counter()
        st = s.stats
        N = n[0] * n[1] * n[2]
        out.update(mesh=n, dof=N, order="012", outer_loops=st["outer_loops"], pcg_iterations=st["lin_iters"], device_ms=st["t_solve_ms"],
                   wall_s=t2 - t1, setup_s=t1 - t0, maxT=st["maxval"], lin_relres=st["lin_relres"],
                   dof_iter_per_s=N * st["lin_iters"] / (st["t_solve_ms"] * 1e-3))
    else:
        raise SystemExit(f"unknown config {a.config}")
    if rank == 0:
        print(json.dumps(out, default=lambda o: o.tolist() if hasattr(o, "tolist") else float(o)))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
