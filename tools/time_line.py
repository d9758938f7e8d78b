"""Per-kernel timing of the line-Jacobi PCG iteration (line solve | operator step) next to the fused Jacobi iteration.
usage: python tools/time_line.py [n] [orders...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from plask_b200 import configs as cf  # noqa: E402
from plask_b200.fem import DeviceFem  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
orders = sys.argv[2:] or ["012", "201"]
for order in orders:
    p = cf.config_B(n, order=order)
    f = DeviceFem(0)
    f.set_mesh(p.axes, p.strides)
    f.set_materials(p.elem_mat, p.T0, p.dT, p.tab_lat, p.tab_vert)
    f.set_field(300.)
    f.set_dirichlet(p.bc_nodes, p.bc_values)
    f.set_source(p.heat)
    f.update_conductivity_thermal()
    pres = [int(a) for a in os.environ["PRE"].split(",")] if os.environ.get("PRE") else ((0, 1, 2) if order == "012" else (0, 1))
    for pre in pres:
        f.bench_pcg(20, precond=pre)
        a = f.bench_pcg(200, precond=pre)
        b = f.bench_pcg(100, split_timing=True, precond=pre)
        t0 = time.perf_counter()
        rc, st = f.solve_linear(precond=pre, maxit=1, lin_tol=1e-30)
        t1 = time.perf_counter()
        print(f"order {order} precond {pre}: graph {a['ms'] / 200:.4f} ms/iter; split: first kernel {b['apply_ms'] / 100:.4f} ms, "
              f"second {b['update_ms'] / 100:.4f} ms; prepare + 1 iteration {st['t_solve_ms']:.2f} ms (wall {1e3 * (t1 - t0):.2f})", flush=True)
    f.close()
