"""Timing of the Diffusion3D path on the GPU next to the reference's algorithm on the CPU (band Cholesky dpbtrf/dpbtrs per loop, through
oracle/diffusion_oracle.py — test infrastructure used here as the CPU arm, like bench.py's cpu_baseline):

    python tools/time_diffusion.py [--cpu-max N]      -> one JSON line per mesh size

Workload: quarter disc of radius R with a Gaussian current profile (the reference's own test structure scaled up), GaAs-like A, B, C, D,
maxerr 1e-4 %, from U = 0."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from plask_b200.diffusion import DeviceDiffusion  # noqa: E402

A, B, C, D = 3e7, 1.7e-10, 6e-27, 10.


def case(n, h=0.04):
    R = h * (n - 1)
    ax = np.linspace(0., R, n)
    xm = 0.5 * (ax[1:] + ax[:-1])
    mask = xm[:, None]**2 + xm[None, :]**2 <= R * R
    X, Y = np.meshgrid(ax, ax, indexing="ij")
    J = 3e27 * (1. + 4. * np.exp(-(X**2 + Y**2) / (0.2 * R)**2))
    return ax, mask, J.ravel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="101,201,401,801,1601")
    ap.add_argument("--cpu-max", type=int, default=201)
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    for n in [int(s) for s in args.sizes.split(",")]:
        ax, mask, J = case(n)
        d = DeviceDiffusion(0)
        d.set_mesh(ax, ax, 0, mask.ravel())
        d.set_parameters(A, B, C, 1e8 * D)
        d.set_current(J)
        best = None
        for _ in range(args.repeat):
            d.set_concentration(None)
            t0 = time.perf_counter()
            st = d.compute(0, 1e-4)
            wall = time.perf_counter() - t0
            if best is None or st["t_solve_ms"] < best["t_solve_ms"]:
                best = dict(st, wall_ms=1e3 * wall)
        U = d.get_concentration()
        out = dict(n=n, unknowns=int(3 * n * n), elements=int(mask.sum()), loops=best["loops"], pcg_iterations=best["lin_iters"],
                   gpu_ms=round(best["t_solve_ms"], 3), gpu_wall_ms=round(best["wall_ms"], 3),
                   us_per_pcg_iteration=round(1e3 * best["t_solve_ms"] / max(best["lin_iters"], 1), 2), err=best["err"],
                   launches=best["kernel_launches"])
        if n <= args.cpu_max:
            from oracle import diffusion_oracle as orc
            o = orc.Diffusion3DOracle(ax, ax, mask)
            t0 = time.perf_counter()
            loops = o.compute(A, B, C, 1e8 * D, J, maxerr=1e-4)
            cpu = time.perf_counter() - t0
            out.update(cpu_s=round(cpu, 2), cpu_loops=loops, cpu_threads=os.cpu_count(),
                       max_rel_diff=float(np.abs(U - o.U).max() / np.abs(o.U).max()), speedup=round(cpu / (1e-3 * best["wall_ms"]), 1))
        print(json.dumps(out), flush=True)
        d.close()


if __name__ == "__main__":
    main()
