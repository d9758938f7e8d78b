"""single-process, single-GPU emulation of slab mode with threads (debug)"""
import os, sys, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from plask_b200 import configs as cf
from plask_b200.fem import DeviceFem

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = (8 * world + 3, 14, 12)
p = cf.config_B(n)
parts = [cf.slab_problem(p, r, world) for r in range(world)]
fems = []
for r, (q, ol, oh, _) in enumerate(parts):
    f = DeviceFem(0)
    f.set_mesh(q.axes, q.strides)
    f.slab_configure(r, world, ol, oh)
    fems.append(f)
for f in fems:
    f.slab_connect_local(fems)
for f, (q, ol, oh, _) in zip(fems, parts):
    f.set_materials(q.elem_mat, q.T0, q.dT, q.tab_lat, q.tab_vert)
    f.set_field(300.)
    f.set_dirichlet(q.bc_nodes, q.bc_values)
    f.set_source(q.heat)
res = [None] * world
def run(r):
    f = fems[r]
    f.update_conductivity_thermal()
    res[r] = f.solve_linear(maxit=20000, lin_tol=1e-10)
ths = [threading.Thread(target=run, args=(r,)) for r in range(world)]
[t.start() for t in ths]; [t.join() for t in ths]
print([ (rc, st["lin_iters"], st["lin_relres"]) for rc, st in res])
one = DeviceFem(0)
one.set_mesh(p.axes, p.strides); one.set_materials(p.elem_mat, p.T0, p.dT, p.tab_lat, p.tab_vert); one.set_field(300.)
one.set_dirichlet(p.bc_nodes, p.bc_values); one.set_source(p.heat); one.update_conductivity_thermal()
rc, st = one.solve_linear(maxit=20000, lin_tol=1e-10)
print("single", rc, st["lin_iters"], st["lin_relres"])
T1 = one.get_field().reshape(n)
T = np.concatenate([cf.slab_field_owned(q, f.get_field(), ol, oh) for f, (q, ol, oh, _) in zip(fems, parts)], axis=0)
print("max diff", np.abs(T - T1).max(), "per plane", np.abs(T - T1).max(axis=(1, 2)))
