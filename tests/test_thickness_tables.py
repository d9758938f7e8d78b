"""Row a2 of SURVEY.md §8 (layer thickness, therm3d.cpp:81-114) and the tabulation of thermk(T, thickness):
  * the thickness restatements — the oracle's C (orc_thickness), the C++ host adapter (plaskfem::layer_thickness, through
    tests/cpp/adapter_test.cpp) and the Python mirror (configs.layer_thickness) — agree on random material stacks in all
    six iteration orders;
  * a thickness-dependent conductivity (GaN, materials/semiconductors35/nitrides/GaN.cpp:38-44) runs end to end through the
    (material, thickness) -> table id path and matches the oracle's Cholesky solve (GPU);
  * the error of the 0.25 K tables against the analytic thermk(T) is bounded far below the 1e-3 K parity tolerance."""
import numpy as np
import pytest

from helpers import oracle_mesh, oracle_thermal, random_problem
from oracle import oracle as orc
from plask_b200 import configs as cf
from plask_b200 import materials as M


def _random_stack(n, order, seed):
    rng = np.random.default_rng(seed)
    p = random_problem(n, order, seed=seed)
    # materials in vertical runs of random length per column block (so that runs, single layers and whole columns occur)
    ne = tuple(k - 1 for k in n)
    mat = np.zeros(ne, dtype=np.uint32)
    for i0 in range(ne[0]):
        for i1 in range(ne[1]):
            r = 0
            while r < ne[2]:
                L = int(rng.integers(1, 5))
                mat[i0, i1, r:r + L] = rng.integers(0, 3)
                r += L
    p.elem_mat = p.to_elem_order(mat, np.uint32)
    return p


@pytest.mark.parametrize("order", ["012", "021", "102", "120", "201", "210"])
def test_mirror_thickness_equals_oracle(order):
    p = _random_stack((5, 6, 17), order, seed=7)
    want = orc.thickness(oracle_mesh(p), p.elem_mat)
    got = cf.layer_thickness(p, p.elem_mat)
    assert np.array_equal(np.isnan(want), np.isnan(got)) and not np.isnan(want).any()
    assert np.abs(want - got).max() <= 1e-12 * np.abs(want).max()


def test_cpp_adapter_thickness_equals_oracle(tmp_path):
    """plaskfem::layer_thickness + material_ids (what the plugin compiles) on the same stack as the oracle"""
    import os
    import subprocess
    from test_adapter_cpp import LIBDIR, ROOT, SRC
    import plask_b200
    plask_b200.build()
    exe = str(tmp_path / "adapter_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", LIBDIR,
                           "-lplaskfem_cuda", f"-Wl,-rpath,{LIBDIR}"])
    for order in ("012", "201", "120"):
        p = _random_stack((5, 6, 17), order, seed=11)
        fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
        with open(fin, "wb") as f:
            np.array(p.n, dtype=np.uint64).tofile(f)
            for a in p.axes:
                np.asarray(a, dtype=np.float64).tofile(f)
            p.elem_mat.astype(np.uint32).tofile(f)
        r = subprocess.run([exe, "thickness", order, fin, fout], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stdout + r.stderr
        raw = np.fromfile(fout, dtype=np.uint8)
        th = raw[:8 * p.E].view(np.float64)
        ids = raw[8 * p.E:].view(np.uint32)
        want = orc.thickness(oracle_mesh(p), p.elem_mat)
        assert np.abs(th - want).max() <= 1e-12 * np.abs(want).max()
        # ids distinguish exactly the (material, thickness) pairs
        pid, pairs = cf.thickness_material_ids(p.elem_mat, want)
        assert len(np.unique(ids)) == len(pairs)
        assert len(set(zip(ids.tolist(), pid.tolist()))) == len(pairs)


def _gan_problem(n=(12, 10, 30)):
    """GaN layers of different thickness (k depends on the layer thickness, GaN.cpp:38-44) between GaAs spacers"""
    p = cf.config_A(n, order="optimal", heat=3e15)
    ne = tuple(k - 1 for k in p.n)
    mat = np.zeros(ne, dtype=np.uint32)             # 0 GaAs, 1 GaN
    layers = [(2, 3), (6, 10), (13, 14), (15, 22), (24, 26)]   # GaN runs [lo, hi) of 1, 4, 1, 7, 2 elements, all of different height
    for lo, hi in layers:
        mat[:, :, lo:hi] = 1
    mat[: ne[0] // 2, :, 15:22] = 0                 # the thick run exists in half of the columns only
    key = p.to_elem_order(mat, np.uint32)
    return p, key


def _tables_for(p, key, thickness, dT=0.25):
    ids, pairs = cf.thickness_material_ids(key, thickness)
    T0, nT = 250., int(round(350. / dT)) + 1
    models = [(lambda T, t=t: M.thermk_GaN(T, t)) if m == 1 else M.thermk_GaAs for m, t in pairs]
    _, _, lat, vert = M.sample_tables(models, T0, dT, nT)
    return ids, pairs, T0, dT, lat, vert


def test_thickness_ids_are_distinct_per_layer():
    p, key = _gan_problem()
    th = cf.layer_thickness(p, key)
    assert np.abs(th - orc.thickness(oracle_mesh(p), key)).max() <= 1e-12
    ids, pairs, *_ = _tables_for(p, key, th)
    gan = [t for m, t in pairs if m == 1]
    assert len(gan) == 5 and len(set(round(t, 9) for t in gan)) == 5      # five GaN thicknesses -> five table ids
    k300 = [float(M.thermk_GaN(300., t)[0]) for t in sorted(gan)]
    assert all(b > a for a, b in zip(k300, k300[1:]))                      # thicker layer, higher conductivity


@pytest.mark.gpu
def test_thickness_dependent_conductivity_end_to_end():
    """GPU path: ids from the MIRROR's thickness; oracle: ids from ITS OWN orc_thickness; same analytic model behind both"""
    from plask_b200.solvers import Static3D
    p, key = _gan_problem()
    ids_g, _, T0, dT, lat, vert = _tables_for(p, key, cf.layer_thickness(p, key))
    p.elem_mat, p.T0, p.dT, p.tab_lat, p.tab_vert = ids_g, T0, dT, lat, vert
    s = Static3D("gan")
    s.problem = p
    s.iterative.maxerr = 1e-11
    s.iterative.maxit = 100000
    s.compute(0)
    q, _ = _gan_problem()
    ids_o, _, T0, dT, lat, vert = _tables_for(q, key, orc.thickness(oracle_mesh(q), key))
    q.elem_mat, q.T0, q.dT, q.tab_lat, q.tab_vert = ids_o, T0, dT, lat, vert
    o = oracle_thermal(q, algorithm="cholesky")
    o.compute(0)
    assert s.stats["outer_loops"] == len(o.history)
    assert np.abs(s.outTemperature() - o.temperatures).max() <= 1e-3
    assert o.maxT - 300. > 1.         # the problem is not trivial
    # and the thickness matters: with the bulk value for every GaN layer the field differs by far more than the tolerance
    r, _ = _gan_problem()
    ids_b, pairs = cf.thickness_material_ids(key, np.zeros(r.E))
    _, _, lat, vert = M.sample_tables([(lambda T: M.thermk_GaN(T, 1e4)) if m == 1 else M.thermk_GaAs for m, t in pairs], T0, dT, lat.shape[1])
    r.elem_mat, r.T0, r.dT, r.tab_lat, r.tab_vert = ids_b, T0, dT, lat, vert
    ob = oracle_thermal(r, algorithm="cholesky")
    ob.compute(0)
    assert np.abs(ob.temperatures - o.temperatures).max() > 0.05
    s.invalidate()


def test_table_interpolation_error_is_far_below_the_parity_tolerance():
    """the 0.25 K tables against the analytic thermk(T) / cond(T): relative error of the linear interpolation
    <= 1e-6 for every material of the configs, and the converged Static3D field moves by < 1e-4 K when the tables are
    refined 64 times (oracle Cholesky; the reference evaluates the analytic formula)."""
    T0, dT, lat, _ = cf.thermal_tables()
    Tm = T0 + dT * (np.arange(lat.shape[1] - 1) + 0.5)          # interval midpoints: where linear interpolation errs most
    models = [M.thermk_GaAs, lambda T: M.thermk_AlGaAs(T, 0.73), M.thermk_AlOx, M.thermk_Au, M.thermk_Cu, M.thermk_air]
    rows = [0, 1, 3, 4, 5, 6]
    for f, r in zip(models, rows):
        exact = f(Tm)[0]
        interp = 0.5 * (lat[r, :-1] + lat[r, 1:])
        assert np.abs(interp / exact - 1.).max() <= 1e-6
    T0e, dTe, late, _ = cf.electrical_tables()
    Tme = T0e + dTe * (np.arange(late.shape[1] - 1) + 0.5)
    for f, r in ((M.cond_GaAs, 0), (M.cond_Au, 4), (lambda T: M.cond_doped(T, 2e18, 2000., 1.4), 7)):
        assert np.abs(0.5 * (late[r, :-1] + late[r, 1:]) / f(Tme)[0] - 1.).max() <= 2e-6
    p = cf.config_B((14, 16, 40))
    o = oracle_thermal(p, algorithm="cholesky")
    o.compute(0)
    fine = cf.config_B((14, 16, 40))
    fine.T0, fine.dT, fine.tab_lat, fine.tab_vert = cf.thermal_tables(250., 0.25 / 64, 64 * 1600 + 1)
    of = oracle_thermal(fine, algorithm="cholesky")
    of.compute(0)
    assert o.maxT - 300. > 5.
    assert np.abs(o.temperatures - of.temperatures).max() <= 1e-4
