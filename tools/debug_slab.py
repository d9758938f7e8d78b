"""2-process debug of slab mode: torchrun --nproc-per-node 2 tools/debug_slab.py"""
import os, sys
import numpy as np
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from plask_b200 import configs as cf
from plask_b200.fem import DeviceFem

def ag(b):
    out = [None] * dist.get_world_size(); dist.all_gather_object(out, b); return out

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = (8 * world + 3, 14, 12)
p = cf.config_B(n)
q, ol, oh, (lo, hi) = cf.slab_problem(p, rank, world)
f = DeviceFem(rank)
f.set_mesh(q.axes, q.strides)
f.slab_configure(rank, world, ol, oh)
f.slab_connect(ag(f.slab_export()))
f.set_materials(q.elem_mat, q.T0, q.dT, q.tab_lat, q.tab_vert)
f.set_field(300.)
f.set_dirichlet(q.bc_nodes, q.bc_values)
f.set_source(q.heat)
f.update_conductivity_thermal()
for maxit in (1, 2, 3, 5, 20000):
    f.set_field(300.)
    rc, st = f.solve_linear(maxit=maxit, lin_tol=1e-10)
    T = cf.slab_field_owned(q, f.get_field(), ol, oh)
    parts = ag(T)
    if rank == 0:
        one = DeviceFem(0)
        one.set_mesh(p.axes, p.strides); one.set_materials(p.elem_mat, p.T0, p.dT, p.tab_lat, p.tab_vert); one.set_field(300.)
        one.set_dirichlet(p.bc_nodes, p.bc_values); one.set_source(p.heat); one.update_conductivity_thermal()
        rc1, st1 = one.solve_linear(maxit=maxit, lin_tol=1e-10)
        T1 = one.get_field().reshape(n)
        Ts = np.concatenate(parts, axis=0)
        print(f"maxit {maxit}: slab iters {st['lin_iters']} relres {st['lin_relres']:.3e} | single iters {st1['lin_iters']} relres {st1['lin_relres']:.3e} | "
              f"max diff {np.abs(Ts - T1).max():.3e} per plane {np.array2string(np.abs(Ts - T1).max(axis=(1, 2)), precision=1)}", flush=True)
        one.close()
    dist.barrier()
f.close()
dist.destroy_process_group()
