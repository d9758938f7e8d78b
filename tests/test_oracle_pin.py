"""Pin the CPU oracle against the only results the reference's own tests hold for this path:
solvers/electrical/shockley/tests/shockley3d.py:36-83 (analytic current, capacitance, heat)."""
import numpy as np
import pytest

from oracle import oracle as orc


def shockley3d_reference_case(order="optimal"):
    """The structure of shockley3d.py:36-60 restated as flat arrays.

    Stack from the bottom: contact 700x700x1, conductor 1000x1000x300, GaAs junction 0.02
    (role 'active'), conductor 300, contact 700x700x1; DivideGenerator(prediv=(3,3,2),
    gradual=False) splits every geometry interval into 3,3,2 equal parts.
    Material ids: 0 = Conductor (cond 1e9, eps 1), 1 = GaAs (junction), 2 = air.
    """
    def divide(edges, k):
        out = [edges[0]]
        for a, b in zip(edges[:-1], edges[1:]):
            out += [a + (b - a) * (i + 1) / k for i in range(k)]
        return np.array(out)

    x = divide([-500., -350., 350., 500.], 3)
    z = divide([0., 1., 301., 301.02, 601.02, 602.02], 2)
    mesh = orc.Mesh(x, x.copy(), z, order)
    xm, zm = 0.5 * (x[1:] + x[:-1]), 0.5 * (z[1:] + z[:-1])
    X, Y, Z = np.meshgrid(xm, xm, zm, indexing="ij")
    mat = np.zeros(X.shape, dtype=np.uint32)
    junc = np.zeros(X.shape, dtype=np.uint32)
    in_contact_layer = (Z < 1.) | (Z > 601.02)
    outside = (np.abs(X) > 350.) | (np.abs(Y) > 350.)
    mat[in_contact_layer & outside] = 2
    is_j = (Z > 301.) & (Z < 301.02)
    mat[is_j] = 1
    junc[is_j] = 1
    eg = mesh.elems_grid()
    elem_mat = np.zeros(mesh.E, dtype=np.uint32); elem_mat[eg.ravel()] = mat.ravel()
    elem_junc = np.zeros(mesh.E, dtype=np.uint32); elem_junc[eg.ravel()] = junc.ravel()
    gaas_cond = 1e2 * 1.60217733e-19 * 8000. * 1e16  # GaAs.cpp:234-237 at 300 K (unused: junction)
    sig = np.array([[1e9, 1e9], [gaas_cond, gaas_cond], [0.55e-14, 0.55e-14]])  # air.cpp:43-46
    tables = orc.Tables(300., 100., sig, sig)
    eps = np.array([1., 12.9, 1.])[elem_mat]  # GaAs.cpp:299-301; Conductor eps = 1 (shockley3d.py:31-32)
    noheat = (elem_mat == 2).astype(np.uint8)  # Material::EMPTY, electr3d.cpp:472
    ng = mesh.nodes_grid()
    Xn, Yn = np.meshgrid(x, x, indexing="ij")
    inc = (np.abs(Xn) <= 350. + 1e-9) & (np.abs(Yn) <= 350. + 1e-9)
    top_nodes = ng[:, :, -1][inc]
    bot_nodes = ng[:, :, 0][inc]
    # voltage_boundary: TopOf(contact, top) = 0 V first, BottomOf(contact, bottom) = 1 V second
    nodes = np.concatenate([top_nodes, bot_nodes]).astype(np.uintp)
    values = np.concatenate([np.zeros(len(top_nodes)), np.ones(len(bot_nodes))])
    return dict(mesh=mesh, elem_mat=elem_mat, elem_junc=elem_junc, tables=tables, eps=eps, noheat=noheat,
                nodes=nodes, values=values)


@pytest.mark.parametrize("algorithm,precond", [("cholesky", None), ("iterative", "ic"), ("iterative", "jac"), ("pcg", None)])
def test_shockley3d_analytic(algorithm, precond):
    if algorithm == "iterative" and not orc.ref_available():
        pytest.skip("oracle/_ref not built")
    c = shockley3d_reference_case()
    assert c["mesh"].order == "201" and c["mesh"].n == (10, 10, 11)
    kw = dict(itmaxerr=1e-6, maxit=1000)
    maxerr = 1e-3                                            # shockley3d.py:56
    if algorithm == "pcg":
        # A noise-free solver takes ever smaller steps towards the fixed point (rate ~0.9/loop), so
        # the step-size criterion 1e-3 % stops it ~1e-4 away from the analytic current; the
        # reference's solvers only get closer because their round-off noise keeps the loop going.
        kw = dict(itmaxerr=1e-14, maxit=200000)
        maxerr = 1e-5
    s = orc.Shockley3DOracle(c["mesh"], c["elem_mat"], c["tables"], c["nodes"], c["values"], elem_junc=c["elem_junc"],
                             beta=10., js=1., maxerr=maxerr, algorithm=algorithm, precond=precond or "ic",
                             noheat=c["noheat"], eps=c["eps"], **kw)
    s.compute(1000)
    S = 1e6
    correct_current = 1e-9 * S * 1. * (np.exp(10.) - 1)
    assert abs(s.get_total_current()) == pytest.approx(correct_current, abs=0.5e-3)       # shockley3d.py:64-65
    capacitance = 8.854187817e-6 * 12.9 * S / 0.02
    assert s.get_capacitance() == pytest.approx(capacitance, abs=0.5e-2)                  # :66-67
    assert s.get_total_heat() == pytest.approx(correct_current * 1., abs=0.5e-3)         # :68-69


def test_shockley3d_beta_of_T():
    """shockley3d.py:75-83: beta as a python function of T, inTemperature 300 then 250."""
    c = shockley3d_reference_case()
    s = orc.Shockley3DOracle(c["mesh"], c["elem_mat"], c["tables"], c["nodes"], c["values"], elem_junc=c["elem_junc"],
                             beta=lambda T: np.log(T * 70), js=1., maxerr=1e-3, algorithm="cholesky",
                             noheat=c["noheat"], eps=c["eps"])
    s.compute(1000)
    assert abs(s.get_total_current()) == pytest.approx(1e-9 * 1e6 * (21000 - 1), abs=0.5e-3)
    s.Te[:] = 250.
    s.compute(1000)
    assert abs(s.get_total_current()) == pytest.approx(1e-9 * 1e6 * (17500 - 1), abs=0.5e-3)


def test_shockley3d_conductivity():
    """shockley3d.py:85-91: outConductivity equals material cond(300) / (0,5) in the junction."""
    c = shockley3d_reference_case()
    s = orc.Shockley3DOracle(c["mesh"], c["elem_mat"], c["tables"], c["nodes"], c["values"], elem_junc=c["elem_junc"],
                             beta=10., js=1.)
    s.load_conductivity()
    expect = np.where(c["elem_junc"][:, None] > 0, np.array([0., 5.]), c["tables"].lat[c["elem_mat"], 0][:, None])
    assert np.array_equal(s.conds, expect)


def test_shockley3d_analytic_excluded():
    """shockley3d.py:71-73 (testComputationsExcluded): empty_elements='exclude' — the masked mesh drops the air
    elements beside the contacts; same analytic current, capacitance and heat.  (This is also the default mesh of the
    reference's Cholesky path, fem_solver.hpp:183-187.)"""
    c = shockley3d_reference_case()
    included = (c["elem_mat"] != 2).astype(np.uint8)
    s = orc.Shockley3DOracle(c["mesh"], c["elem_mat"], c["tables"], c["nodes"], c["values"], elem_junc=c["elem_junc"],
                             beta=10., js=1., maxerr=1e-3, algorithm="cholesky", noheat=c["noheat"], eps=c["eps"],
                             included=included)
    s.compute(1000)
    A = s._matrix()
    assert A.Nm == 764 and c["mesh"].N == 1100          # 4 x 2 x 42 corner-region nodes of the two contact layers drop out
    S = 1e6
    correct_current = 1e-9 * S * 1. * (np.exp(10.) - 1)
    assert abs(s.get_total_current()) == pytest.approx(correct_current, abs=0.5e-3)
    assert s.get_capacitance() == pytest.approx(8.854187817e-6 * 12.9 * S / 0.02, abs=0.5e-2)
    assert s.get_total_heat() == pytest.approx(correct_current * 1., abs=0.5e-3)
    # against the full mesh (air kept, sigma = 5.5e-15 S/m): same potential on the masked nodes
    f = orc.Shockley3DOracle(c["mesh"], c["elem_mat"], c["tables"], c["nodes"], c["values"], elem_junc=c["elem_junc"],
                             beta=10., js=1., maxerr=1e-3, algorithm="cholesky", noheat=c["noheat"], eps=c["eps"])
    f.compute(len(s.history))
    assert np.abs(s.potential - f.potential)[A.active].max() < 1e-9


def test_brick_stiffness_is_the_exact_galerkin_integral():
    """The closed-form 8x8 brick matrix of setMatrix (therm3d.cpp:226-237, electr3d.cpp:326-337) restated in
    oracle/fem3d_oracle.c equals the Galerkin integral of grad N_i . k grad N_j over the brick with trilinear shape functions
    (2x2x2 Gauss quadrature is exact for it), and the load vector is the exact integral of a constant source — an
    independent, first-principles pin of the assembly for Static3D, which has no reference test."""
    dx, dy, dz = 0.7, 1.9, 0.013                 # um; strongly anisotropic like a thin epitaxial layer
    k_lat, k_vert = 44.0, 3.5                    # W/(m K)
    mesh = orc.Mesh(np.array([0., dx]), np.array([0., dy]), np.array([0., dz]), "012")
    A = orc.Sparse14(mesh)
    B = np.zeros(8)
    heat = np.array([2.5e15])
    A.assemble(np.array([[k_lat, k_vert]]), heat, B)
    K = np.column_stack([A.mult(np.eye(8)[:, j].copy()) for j in range(8)])
    # reference node numbering of the element: bit 0 = axis 0 upper, bit 1 = axis 1, bit 2 = axis 2 (therm3d.cpp:190-197)
    ng = mesh.nodes_grid()
    idx = [ng[l & 1, (l >> 1) & 1, (l >> 2) & 1] for l in range(8)]
    g = 1. / np.sqrt(3.)
    Kq = np.zeros((8, 8))
    for gx in (-g, g):
        for gy in (-g, g):
            for gz in (-g, g):
                grads = np.zeros((8, 3))
                for l in range(8):
                    sx, sy, sz = (1 if l & 1 else -1), (1 if l & 2 else -1), (1 if l & 4 else -1)
                    grads[l] = [sx * (1 + sy * gy) * (1 + sz * gz) / 8. * 2. / dx,
                                sy * (1 + sx * gx) * (1 + sz * gz) / 8. * 2. / dy,
                                sz * (1 + sx * gx) * (1 + sy * gy) / 8. * 2. / dz]
                kk = np.array([k_lat, k_lat, k_vert]) * 1e-6            # W/(um K): coordinates are in um (:215)
                Kq += (grads * kk) @ grads.T * (dx * dy * dz / 8.)
    Kref = K[np.ix_(idx, idx)]
    assert np.abs(Kref - Kq).max() <= 1e-13 * np.abs(Kq).max()
    assert np.allclose(B[idx], 0.125e-18 * dx * dy * dz * heat[0], rtol=1e-15)   # int N_i f dV = f V / 8, um^3 -> m^3
    assert abs(Kref.sum()) <= 1e-12 * np.abs(Kq).max() and np.allclose(Kref, Kref.T)


def test_nonlinear_loop_converges_to_the_kirchhoff_solution():
    """Static3D's loop (k evaluated at the 8-node average temperature of the previous loop, therm3d.cpp:204-213,311-334) against
    the exact solution of  (k(T) T')' = -q,  T(0) = Tb,  T'(H) = 0  with  k = k0 (300/T)^a  (Kirchhoff transform):
    theta(T) = int_Tb^T k dT = q (2 H z - z^2) / 2.  The discretisation error must fall like h^2."""
    from helpers import face_nodes
    from plask_b200 import configs as cf
    k0, a, Tb, q, H = 45., 1.28, 300., 8.0e13, 6.0                  # GaAs thermk (materials/GaAs.cpp), W/m3, um
    Tt = 250. + 0.05 * np.arange(6001)
    lat = (k0 * (300. / Tt) ** a)[None, :]
    c = k0 * 300. ** a

    def exact(z_um):
        z = z_um * 1e-6
        theta = q * (2. * H * 1e-6 * z - z * z) / 2.
        return (Tb ** (1. - a) + (1. - a) * theta / c) ** (1. / (1. - a))

    errs = []
    for nz in (9, 17, 33):
        axes = [np.linspace(0., 1., 3), np.linspace(0., 1., 3), np.linspace(0., H, nz)]
        p = cf.Problem("kirchhoff", "thermal", axes, "012", None, 250., 0.05, lat, lat.copy(), None, None)
        p.elem_mat = np.zeros(p.E, dtype=np.uint32)
        bot = face_nodes(p, 2, 0)
        p.bc_nodes, p.bc_values = bot.astype(np.uintp), np.full(bot.size, Tb)
        p.heat = np.full(p.E, q)
        m = orc.Mesh(*axes, "012")
        o = orc.Static3DOracle(m, p.elem_mat, orc.Tables(250., 0.05, lat, lat), p.bc_nodes, p.bc_values, heat=p.heat,
                               inittemp=Tb, maxerr=1e-9, algorithm="cholesky")
        o.compute(200)
        T = o.temperatures[np.broadcast_to(p.node_index_grid(), p.n)][1, 1, :]
        errs.append(np.abs(T - exact(axes[2])).max())
        assert T.max() > Tb + q * (H * 1e-6) ** 2 / (2. * k0) + 1.     # k(T) matters: well above the constant-k solution (332 K)
    assert errs[0] / errs[1] == pytest.approx(4., rel=0.25) and errs[1] / errs[2] == pytest.approx(4., rel=0.25), errs
    assert errs[2] < 0.02
