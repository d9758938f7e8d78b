"""CPU checks of the 2-D meta-loop oracle (oracle2d.ThermoElectric2DOracle): its bilinear interpolation against scipy, the exchange on
identical meshes (temperature = mean of the four element nodes, heat = identity), zero heat outside the electrical mesh, and
convergence of the coupled loop on the mesa diode the GPU test uses."""
import numpy as np
import pytest

from helpers import oracle_shockley2d, oracle_static2d, thermoelectric2d_pair
from oracle import oracle2d


def test_interp_bilinear_against_scipy():
    from scipy.interpolate import RegularGridInterpolator
    rng = np.random.default_rng(5)
    xs, ys = np.cumsum(rng.uniform(0.1, 1., 7)), np.cumsum(rng.uniform(0.1, 1., 9))
    v = rng.normal(size=(7, 9))
    xq, yq = rng.uniform(xs[0] - 0.5, xs[-1] + 0.5, 11), rng.uniform(ys[0] - 0.5, ys[-1] + 0.5, 13)
    got = oracle2d.interp_bilinear(xs, ys, v.ravel(), xq, yq).reshape(11, 13)
    ref = RegularGridInterpolator((xs, ys), v)
    X, Y = np.meshgrid(np.clip(xq, xs[0], xs[-1]), np.clip(yq, ys[0], ys[-1]), indexing="ij")      # clamped to the edge outside
    assert np.abs(got - ref(np.stack([X, Y], axis=-1))).max() <= 1e-13


@pytest.mark.parametrize("cyl", [False, True])
def test_exchange_on_identical_meshes(cyl):
    pt, pe = thermoelectric2d_pair(cyl)
    ot, oe = oracle_static2d(pt), oracle_shockley2d(pe)
    o = oracle2d.ThermoElectric2DOracle(ot, oe, tfreq=2)
    rng = np.random.default_rng(1)
    ot.temperatures = rng.uniform(300., 400., ot.mesh.N)
    o.exchange_temperature()
    m = ot.mesh
    assert np.abs(oe.Te - 0.25 * (ot.temperatures[m.ll] + ot.temperatures[m.lr] + ot.temperatures[m.ul] + ot.temperatures[m.ur])).max() <= 1e-11
    oe.compute(2)
    o.exchange_heat()
    assert np.abs(ot.heat - oe.heat_densities()).max() <= 1e-12 * np.abs(ot.heat).max()
    assert ot.heat.max() > 0.


def test_coupled_loop_converges_and_spreader_gets_no_heat():
    pt, pe = thermoelectric2d_pair(False, thermal_shape=(12, 14))
    ot, oe = oracle_static2d(pt), oracle_shockley2d(pe, beta=lambda T: 11. * 300. / T)
    o = oracle2d.ThermoElectric2DOracle(ot, oe, tfreq=4)
    n = o.compute(max_meta_loops=20)
    assert n < 20 and o.history[-1]["terr"] <= pt.maxerr and o.history[-1]["verr"] <= pe.maxerr
    assert ot.maxT > 330.
    ym = 0.5 * (pt.y[1:] + pt.y[:-1])
    below = np.tile(ym < 0., len(pt.x) - 1)
    assert below.any() and np.all(ot.heat[below] == 0.) and ot.heat[~below].max() > 0.
