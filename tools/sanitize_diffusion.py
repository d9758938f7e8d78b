"""Small Diffusion3D run for compute-sanitizer (memcheck): masked mesh, both iteration orders, hole burning, hooks, interpolation.
    compute-sanitizer --tool memcheck python tools/sanitize_diffusion.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plask_b200.diffusion import DeviceDiffusion  # noqa: E402

rng = np.random.default_rng(2)
n0, n1 = 21, 17
ax0, ax1 = np.cumsum(rng.uniform(.05, .2, n0)), np.cumsum(rng.uniform(.05, .2, n1))
act = rng.uniform(size=(n0 - 1, n1 - 1)) > 0.2
for order in (0, 1):
    d = DeviceDiffusion(0)
    d.set_mesh(ax0, ax1, order, (act.T if order else act).ravel())
    ne, nn = d.ne, d.nn
    d.set_parameters(3e7, 1.7e-10, 6e-27, 1e9)
    d.set_current(rng.uniform(0.2, 1, nn) * 1e30)
    d.set_modes(rng.uniform(0, 1, (2, nn, 2)), rng.uniform(0, 1e25, (2, ne, 2)), rng.uniform(0, 1e9, (2, ne, 2)))
    st = d.compute(0, 1e-6)
    assert st["converged"], st
    K, F = d.element_matrices()
    y = d.apply(rng.normal(size=3 * nn))
    f = d.rhs()
    v = d.interpolate(rng.uniform(ax0[0] - .1, ax0[-1] + .1, 500), rng.uniform(ax1[0] - .1, ax1[-1] + .1, 500))
    lin = d.interpolate(rng.uniform(ax0[0], ax0[-1], 100), rng.uniform(ax1[0], ax1[-1], 100), 1)
    print(f"order {order}: {st['loops']} loops, {st['lin_iters']} PCG iterations, |U| max {np.abs(d.get_concentration()).max():.3e}")
    d.close()
print("sanitize diffusion ok")
