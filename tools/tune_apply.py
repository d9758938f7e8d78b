"""Sweep tile configurations of the operator kernel on config B (run on the GPU box).
usage: python tools/tune_apply.py [n] [variant]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plask_b200 import configs
from plask_b200.fem import DeviceFem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
p = configs.config_B(n)
rng = np.random.default_rng(1)
v = rng.standard_normal(p.N)


def setup():
    f = DeviceFem(0)
    f.set_mesh(p.axes, p.strides)
    f.set_materials(p.elem_mat, p.T0, p.dT, p.tab_lat, p.tab_vert)
    f.set_field(300.)
    f.set_dirichlet(p.bc_nodes, p.bc_values)
    f.set_source(p.heat)
    f.update_conductivity_thermal()
    return f


os.environ.pop("PFEM_TILE", None)
f = setup()
ref = f.apply(v, variant=1)
r1 = f.bench_pcg(20, split_timing=True, variant=1)
print(json.dumps(dict(tile="simple", apply_ms=r1["apply_ms"] / 20, update_ms=r1["update_ms"] / 20)))
f.close()
cfgs = sys.argv[3:] or ["32,16,2,0,3", "32,16,2,0,2", "32,16,1,0,3", "32,16,1,0,2", "32,16,4,0,3", "32,32,2,0,2",
                        "32,32,4,0,2", "64,16,2,0,2", "64,16,4,0,2", "64,8,2,0,3", "64,8,1,0,3", "32,8,1,0,4", "32,8,2,0,4",
                        "32,16,2,16,3", "32,16,2,32,3", "32,16,2,64,3", "32,16,2,128,3", "32,16,1,32,3", "64,8,1,32,3"]
for c in cfgs:
    os.environ["PFEM_TILE"] = c
    try:
        f = setup()
        q = f.apply(v, variant=variant)
        err = float(np.abs(q - ref).max() / np.abs(ref).max())
        r = f.bench_pcg(20, split_timing=True, variant=variant)
        g = f.bench_pcg(50, split_timing=False, variant=variant)
        print(json.dumps(dict(tile=c, apply_ms=r["apply_ms"] / 20, update_ms=r["update_ms"] / 20, iter_ms_graph=g["ms"] / 50,
                              relerr=err, gbs_apply=56 * p.N / (r["apply_ms"] / 20 * 1e-3) / 1e9)))
        f.close()
    except Exception as ex:
        print(json.dumps(dict(tile=c, error=str(ex))))
