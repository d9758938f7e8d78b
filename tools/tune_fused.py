"""Sweep tile configurations of the fused single-kernel PCG iteration on config B (run on the GPU box).
usage: python tools/tune_fused.py [n] [cfg ...]   cfg = "tj,rj,ns,minb[,lk]" """
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plask_b200 import configs
from plask_b200.fem import DeviceFem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
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


os.environ.pop("PFEM_FUSED_TILE", None)
f = setup()
ref = f.apply(v, variant=1)
f.close()
cfgs = sys.argv[2:] or ["8,1,3,2", "8,1,2,2", "8,2,3,2", "8,2,2,2", "8,2,2,3", "16,2,2,1", "16,2,2,2", "16,1,2,1", "16,4,2,1",
                        "8,1,3,2,16", "8,1,3,2,32", "8,1,3,2,64", "8,1,3,2,128", "8,1,3,2,256"]
for c in cfgs:
    os.environ["PFEM_FUSED_TILE"] = c
    try:
        f = setup()
        q = f.apply(v, variant=3)
        err = float(np.abs(q - ref).max() / np.abs(ref).max())
        g = f.bench_pcg(50, split_timing=False, variant=3)
        g = f.bench_pcg(100, split_timing=False, variant=3)
        ms = g["ms"] / 100
        print(json.dumps(dict(tile=c, iter_ms=ms, relerr=err, gdofs=p.N / ms / 1e6, gbs88=88 * p.N / (ms * 1e-3) / 1e9)), flush=True)
        f.close()
    except Exception as ex:
        print(json.dumps(dict(tile=c, error=str(ex))), flush=True)
