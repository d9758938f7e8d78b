"""Dynamic2DOracle (oracle/oracle2d.py, corrected femT2d.cpp) pinned from first principles on the CPU: the analytic 1-D cooling of a
slab (the pin of the 3-D dynamic oracle, tests/test_oracle_dynamic.py) in both geometries, second-order convergence of the
Crank-Nicolson scheme, lumped against consistent capacity, and agreement with the independent 3-D oracle on the one-layer
embedding the product uses (plask_b200/solvers2d.embed)."""
import numpy as np
import pytest

from helpers import oracle_dynamic, oracle_dynamic2d
from plask_b200.solvers2d import Problem2D, embed


def cooling2d(ny=41, L=2.0, k=45., cprho=0.327e3 * 5.31749e3, nx=4, cyl=False):
    x = np.linspace(0., 1.5, nx) + (3.0 if cyl else 0.)
    y = np.linspace(0., L, ny)
    tab = np.full((1, 2), float(k))
    ng = np.arange(nx * ny).reshape(nx, ny)
    p = Problem2D("cooling2d", "thermal", x, y, np.zeros((nx - 1) * (ny - 1), dtype=np.uint32), 200., 1000., tab, tab.copy(),
                  ng[:, 0].astype(np.uintp), np.full(nx, 300.), heat=np.zeros((nx - 1) * (ny - 1)), cyl=cyl)
    p.tab_cprho = np.full((1, 2), float(cprho))
    p.meta.update(alpha=k / cprho * 1e3, L=L)
    return p


def initial(p, a=50.):
    return np.tile(300. + a * np.sin(np.pi * p.y / (2. * p.meta["L"])), len(p.x))


def exact(p, t, a=50.):
    L, al = p.meta["L"], p.meta["alpha"]
    return np.tile(300. + a * np.sin(np.pi * p.y / (2. * L)) * np.exp(-al * (np.pi / (2. * L)) ** 2 * t), len(p.x))


@pytest.mark.parametrize("cyl", [False, True])
@pytest.mark.parametrize("lumping", [True, False])
def test_cooling_analytic(cyl, lumping):
    p = cooling2d(cyl=cyl)
    errs = []
    for dt in (4.0, 2.0):
        o = oracle_dynamic2d(p, timestep=dt, methodparam=0.5, lumping=lumping)
        o.temperatures = initial(p)
        o.compute(40.)
        assert o.physical_time == pytest.approx(40. + dt)
        errs.append(np.abs(o.temperatures - exact(p, o.physical_time)).max())
    assert errs[1] < 0.05                      # of an amplitude of ~ 20 K left
    if lumping:
        assert errs[0] / errs[1] > 1.5          # spatial + temporal error, both second order: halving dt alone still helps


@pytest.mark.parametrize("cyl", [False, True])
@pytest.mark.parametrize("lumping", [True, False])
def test_embedding_agrees_with_the_3d_oracle(cyl, lumping):
    """Cartesian: the brick mesh of one layer with z-invariant data IS the 2-D problem (no weights involved); cylindrical: the 2-D
    oracle with r = const far from the axis tends to the Cartesian one"""
    from helpers import thermal2d_problem
    p2 = thermal2d_problem((9, 12), cyl=False)
    o2 = oracle_dynamic2d(p2, timestep=0.5, methodparam=0.5, lumping=lumping, rebuildfreq=3)
    o2.compute(4.)
    if not cyl:
        p3 = embed(p2)
        o3 = oracle_dynamic(p3, timestep=0.5, methodparam=0.5, lumping=lumping, rebuildfreq=3)
        o3.compute(4.)
        assert o2.temperatures.max() - 300. > 1.
        assert np.abs(o3.temperatures[:p2.N] - o2.temperatures).max() <= 1e-9
        assert np.abs(o3.temperatures[p2.N:] - o2.temperatures).max() <= 1e-9
    else:
        pc = thermal2d_problem((9, 12), cyl=True)
        pc.x = pc.x + 1e7                        # a thin shell at r = 10 m: every element has the same weight to 1e-6
        oc = oracle_dynamic2d(pc, timestep=0.5, methodparam=0.5, lumping=lumping, rebuildfreq=3)
        oc.compute(4.)
        assert np.abs(oc.temperatures - o2.temperatures).max() <= 1e-4 * (o2.temperatures.max() - 300.)


def test_cylinder_radial_cooling_bessel():
    """solid cylinder of radius R, wall held at 300 K, T(r, 0) = 300 + a J0(l1 r / R): the exact solution is the same profile times
    exp(-alpha l1^2 t / R^2) (l1 = first zero of J0).  Exercises the radial weights of K AND of the capacity matrix (c r, femT2d.cpp:
    307-313) — the slab test above cannot see a wrong weight on c."""
    from scipy.special import j0, jn_zeros
    from plask_b200.solvers2d import Problem2D
    R, k, cprho, a = 3.0, 45., 0.327e3 * 5.31749e3, 40.
    nx, ny = 121, 3
    x, y = np.linspace(0., R, nx), np.linspace(0., 0.5, ny)
    tab = np.full((1, 2), k)
    ng = np.arange(nx * ny).reshape(nx, ny)
    p = Problem2D("bessel", "thermal", x, y, np.zeros((nx - 1) * (ny - 1), dtype=np.uint32), 200., 1000., tab, tab.copy(),
                  ng[-1, :].astype(np.uintp), np.full(ny, 300.), heat=np.zeros((nx - 1) * (ny - 1)), cyl=True)
    p.tab_cprho = np.full((1, 2), cprho)
    l1 = jn_zeros(0, 1)[0]
    alpha = k / cprho * 1e3                       # um^2 / ns
    profile = np.repeat(j0(l1 * x / R), ny)
    for lumping in (True, False):
        o = oracle_dynamic2d(p, timestep=2.0, methodparam=0.5, lumping=lumping)
        o.temperatures = 300. + a * profile
        o.compute(60.)
        exact = 300. + a * profile * np.exp(-alpha * (l1 / R) ** 2 * o.physical_time)
        left = a * np.exp(-alpha * (l1 / R) ** 2 * o.physical_time)
        assert 0.2 * a < left < 0.8 * a           # a good part of the transient has happened, a good part is left
        assert np.abs(o.temperatures - exact).max() <= 2e-3 * left, lumping
