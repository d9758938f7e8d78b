// adapter_test.cpp — exercises include/plaskfem_cuda.hpp the way the solver plugin would.
//   adapter_test host   host-side logic only (no GPU): thickness, ids, junction detection, BC flattening,
//                       NoDevice mapping when no CUDA device is usable
//   adapter_test gpu    additionally a Static3D solve with a manufactured solution (uniform k, uniform heat,
//                       Dirichlet bottom: T(z) = T0 + q (2Hz - z^2) / (2k), nodally exact for trilinear bricks,
//                       SURVEY.md §8c(ii)) and a two-material k(T) problem run through Context::solve
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "plaskfem_cuda.hpp"

using namespace plaskfem;

static bool near(double a, double b) { return std::fabs(a - b) <= 1e-13 * (std::fabs(a) + std::fabs(b)); }
#define REQUIRE(c) do { if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static Mesh make_mesh(size_t n0, size_t n1, size_t n2, IterationOrder o) {
    Mesh m;
    for (size_t i = 0; i < n0; ++i) m.axis[0].push_back(0.5 * i + 0.01 * i * i);
    for (size_t i = 0; i < n1; ++i) m.axis[1].push_back(0.7 * i);
    for (size_t i = 0; i < n2; ++i) m.axis[2].push_back(0.1 * i + 0.02 * i * i);
    m.order = o;
    return m;
}

static int host_tests() {
    Mesh m = make_mesh(5, 4, 9, ORDER_012);
    size_t s[3], es[3];
    m.strides(s); m.elem_strides(es);
    REQUIRE(s[2] == 1 && s[1] == 9 && s[0] == 36);
    REQUIRE(es[2] == 1 && es[1] == 8 && es[0] == 24);
    Mesh o = make_mesh(5, 4, 9, ORDER_012);
    o.set_optimal_order();               // sizes 5,4,9 -> largest axis slowest: major 2, medium 0, minor 1
    REQUIRE(o.order == ORDER_201);

    // layers: rows 0-2 material 1, rows 3-4 material 2, rows 5-7 material 1; one column differs
    auto key = [](size_t i0, size_t i1, size_t r) -> int {
        if (i0 == 1 && i1 == 2) return r < 4 ? 7 : 1;
        return (r < 3 || r >= 5) ? 1 : 2;
    };
    std::vector<double> th = layer_thickness(m, key);
    const double* z = m.axis[2].data();
    REQUIRE(std::fabs(th[m.elem(0, 0, 1)] - (z[3] - z[0])) < 1e-15);
    REQUIRE(std::fabs(th[m.elem(0, 0, 3)] - (z[5] - z[3])) < 1e-15);
    REQUIRE(std::fabs(th[m.elem(3, 2, 7)] - (z[8] - z[5])) < 1e-15);
    REQUIRE(std::fabs(th[m.elem(1, 2, 2)] - (z[4] - z[0])) < 1e-15);
    REQUIRE(std::fabs(th[m.elem(1, 2, 6)] - (z[8] - z[4])) < 1e-15);
    std::vector<int> keys(m.elements());
    for (size_t i0 = 0; i0 < 4; ++i0) for (size_t i1 = 0; i1 < 3; ++i1) for (size_t r = 0; r < 8; ++r) keys[m.elem(i0, i1, r)] = key(i0, i1, r);
    std::vector<size_t> reps;
    std::vector<uint32_t> ids = material_ids(keys, th, &reps);
    // (1, z3-z0), (2, z5-z3), (1, z8-z5), (7, z4-z0), (1, z8-z4)
    REQUIRE(reps.size() == 5);
    REQUIRE(ids[m.elem(0, 0, 0)] == ids[m.elem(3, 1, 2)] && ids[m.elem(0, 0, 0)] != ids[m.elem(0, 0, 6)]);

    // junction: rows 3..4 active in columns lon 1..2, tra 0..2
    auto jn = [](size_t lon, size_t tra, size_t ver) -> size_t { return (lon >= 1 && lon <= 2 && ver >= 3 && ver < 5) ? 1 : 0; (void)tra; };
    size_t condsize = 0;
    std::vector<pfem_junction> act = setup_active_regions(m, jn, &condsize, "test");
    REQUIRE(act.size() == 1);
    REQUIRE(act[0].bottom == 3 && act[0].top == 5 && act[0].back == 1 && act[0].front == 3 && act[0].left == 0 && act[0].right == 3);
    REQUIRE(act[0].ld == 2 && condsize == 6 && act[0].offset == -1);
    REQUIRE(std::fabs(act[0].height - (z[5] - z[3])) < 1e-15);
    bool thrown = false;
    try {
        auto bad = [](size_t lon, size_t, size_t ver) -> size_t { return (ver >= (lon == 0 ? 2u : 3u) && ver < 5) ? 1 : 0; };
        setup_active_regions(m, bad, &condsize, "test");
    } catch (const ComputationError& e) { thrown = strstr(e.what(), "constant heights") != nullptr; }
    REQUIRE(thrown);

    Dirichlet bc;
    std::vector<size_t> bottom;
    for (size_t i0 = 0; i0 < 5; ++i0) for (size_t i1 = 0; i1 < 4; ++i1) bottom.push_back(m.node(i0, i1, 0));
    bc.add(bottom, 300.);
    REQUIRE(bc.node.size() == 20 && bc.value[19] == 300.);

    // conditions of the 2nd / 3rd kind: getValue semantics, the FIRST condition naming a node wins
    NodeConditions<2> conv;
    REQUIRE(conv.empty());
    conv.add(m.size(), bottom, {1e4, 300.});
    std::vector<size_t> edge = {m.node(0, 0, 0), m.node(0, 0, 1)};
    conv.add(m.size(), edge, {5e4, 350.});
    REQUIRE(!conv.empty() && conv.has[m.node(0, 0, 0)] == 1 && conv.v[0][m.node(0, 0, 0)] == 1e4 && conv.v[1][m.node(0, 0, 1)] == 350.);
    REQUIRE(conv.has[m.node(1, 1, 1)] == 0);
    bool outside = false;
    try { std::vector<size_t> far = {m.size()}; conv.add(m.size(), far, {1., 1.}); } catch (const BadInput&) { outside = true; }
    REQUIRE(outside);

    // masked mesh numbering: keep the elements with i2 >= 2 or i0 < 2 -> nodes (i0 > 2, any i1, i2 < 2) drop out
    {
        auto inc = [](size_t i0, size_t, size_t i2) { return i2 >= 2 || i0 < 2; };
        MaskedNumbering mn(m, inc);
        size_t ne = 0, nn = 0;
        for (size_t i0 = 0; i0 + 1 < 5; ++i0) for (size_t i1 = 0; i1 + 1 < 4; ++i1) for (size_t i2 = 0; i2 + 1 < m.n(2); ++i2) ne += inc(i0, i1, i2);
        for (size_t i0 = 0; i0 < 5; ++i0) for (size_t i1 = 0; i1 < 4; ++i1) for (size_t i2 = 0; i2 < m.n(2); ++i2) nn += !(i0 > 2 && i2 < 2);
        REQUIRE(mn.full_of_elem.size() == ne && mn.full_of_node.size() == nn);
        REQUIRE(mn.node_of_full[m.node(4, 1, 0)] == MaskedNumbering::NONE && mn.node_of_full[m.node(2, 1, 0)] != MaskedNumbering::NONE);
        for (size_t k = 1; k < mn.full_of_node.size(); ++k) REQUIRE(mn.full_of_node[k] > mn.full_of_node[k - 1]);   // full-mesh order
        std::vector<uint32_t> ids = mn.mark_excluded(std::vector<uint32_t>(m.elements(), 7));
        REQUIRE(ids[m.elem(3, 0, 0)] == PFEM_MAT_EXCLUDED && ids[m.elem(1, 0, 0)] == 7 && ids[m.elem(3, 0, 2)] == 7);
        std::vector<double> full(m.size()), masked(nn), back(m.size());
        for (size_t i = 0; i < full.size(); ++i) full[i] = 1. + (double)i;
        mn.nodes_to_masked(full.data(), masked.data());
        mn.nodes_to_full(masked.data(), back.data(), -1.);
        REQUIRE(back[m.node(4, 1, 0)] == -1. && back[m.node(1, 1, 1)] == full[m.node(1, 1, 1)]);
    }

    Tables t = sample_tables(2, [](uint32_t id, double T) { return std::make_pair(10. * (id + 1) * 300. / T, 5. * (id + 1)); }, 250., 0.5, 701);
    REQUIRE(t.lat.size() == 1402 && std::fabs(t.lat[701 + 100] - 20. * 300. / 300.) < 1e-12);

    // the 2-D solvers' bookkeeping and the library's host-side flattening of the 2-D edge conditions
    {
        const std::vector<double> x = {0., 1., 3.}, y = {0., 0.5, 1.0, 2.0};
        Embedding2D emb(x, y);
        REQUIRE(emb.plane() == 12 && emb.elements() == 6 && emb.mesh.size() == 24 && emb.mesh.elements() == 6);
        REQUIRE(emb.mesh.node(0, 2, 1) == emb.node(2, 1) && emb.mesh.node(1, 2, 1) == emb.node(2, 1) + emb.plane());
        REQUIRE(emb.mesh.elem(0, 1, 2) == emb.elem(1, 2));
        std::vector<double> w = emb.radial_weights();
        REQUIRE(w.size() == 2 && w[0] == 0.5 && w[1] == 2.0);
        Dirichlet bc;
        emb.add_dirichlet(bc, emb.node(1, 0), 300.);
        REQUIRE(bc.node.size() == 2 && bc.node[0] == 4 && bc.node[1] == 16 && bc.value[1] == 300.);
        std::vector<double> f2(12);
        for (size_t i = 0; i < 12; ++i) f2[i] = double(i);
        std::vector<double> f3 = emb.lift(f2.data()), back(12);
        REQUIRE(f3.size() == 24 && f3[5] == 5. && f3[17] == 5.);
        emb.restrict_to_plane(f3.data(), back.data());
        REQUIRE(back == f2);
        bool threw = false;
        try { Embedding2D bad({0.}, y); } catch (const BadInput&) { threw = true; }
        REQUIRE(threw);
        REQUIRE(Embedding2D::boundary_mode(false) == 1 && Embedding2D::boundary_mode(true) == 2);

        // convection on the top edge (nodes (i0, 3)): one edge per element column, therm2d.cpp:236-246 (Cartesian) / :385-397 (cylindrical)
        const size_t N = emb.plane();
        NodeConditions<2> cv;
        for (size_t i0 = 0; i0 < 3; ++i0) cv.add_node(N, emb.node(i0, 3), {100., 300.});
        pfem_boundary b;
        memset(&b, 0, sizeof b);
        b.has_conv = cv.has.data(); b.conv_coeff = cv.v[0].data(); b.conv_ambient = cv.v[1].data();
        std::vector<double> load(N), rc(N), ra(N), K(N * N);
        b.verbatim = 0; b.mode2d = 1;
        REQUIRE(pfem_edges2d_host(3, x.data(), 4, y.data(), &b, load.data(), rc.data(), ra.data(), K.data()) == PFEM_OK);
        const size_t a = emb.node(0, 3), c = emb.node(1, 3), d = emb.node(2, 3);
        REQUIRE(near(load[a], 0.5e-6 * 1. * 100. * 300.) && near(load[c], 0.5e-6 * (1. + 2.) * 100. * 300.) && near(load[d], 0.5e-6 * 2. * 100. * 300.));
        REQUIRE(near(K[a * N + a], 1e-6 * 200. * 1. / 6.) && near(K[a * N + c], 1e-6 * 200. * 1. / 12.) && near(K[c * N + d], 1e-6 * 200. * 2. / 12.));
        REQUIRE(near(K[c * N + c], 1e-6 * 200. * (1. + 2.) / 6.) && K[a * N + d] == 0. && load[emb.node(1, 2)] == 0.);
        b.verbatim = 1; b.mode2d = 2;      // cylindrical, as written: no 1e-6, r - len/6 on the inner node, a second factor r_mid
        REQUIRE(pfem_edges2d_host(3, x.data(), 4, y.data(), &b, load.data(), rc.data(), ra.data(), K.data()) == PFEM_OK);
        REQUIRE(near(K[a * N + a], 0.5 * (200. * 1. / 6.) * (0.5 - 1. / 6.)) && near(K[a * N + c], 0.5 * (200. * 1. / 12.) * 0.5));
        REQUIRE(near(load[a], 0.125e-6 * 1. * 200. * 600. * (0.5 - 1. / 6.)));
        b.mode2d = 0;
        REQUIRE(pfem_edges2d_host(3, x.data(), 4, y.data(), &b, load.data(), rc.data(), ra.data(), K.data()) == PFEM_ERR_BAD_INPUT);
    }

    if (pfem_device_count() == 0) {
        bool nodev = false;
        try { Context c(0, "nodev"); } catch (const NoDevice&) { nodev = true; }
        REQUIRE(nodev);   // no CPU fallback: creating a context without a device must throw
    }
    printf("adapter host tests ok\n");
    return 0;
}

static int gpu_tests() {
    // ---- manufactured solution
    for (int ord = 0; ord < 6; ++ord) {
        Mesh m = make_mesh(9, 7, 21, (IterationOrder)ord);
        const size_t N = m.size(), E = m.elements();
        const double k = 44., q = 3e15, T0 = 300.;
        Context c(0, "manufactured");
        c.set_mesh(m);
        std::vector<uint32_t> ids(E, 0);
        Tables t = sample_tables(1, [&](uint32_t, double) { return std::make_pair(k, k); }, 250., 1., 400);
        c.set_materials(ids, t);
        c.fill_field(T0);
        Dirichlet bc;
        std::vector<size_t> bottom;
        for (size_t i0 = 0; i0 < m.n(0); ++i0) for (size_t i1 = 0; i1 < m.n(1); ++i1) bottom.push_back(m.node(i0, i1, 0));
        bc.add(bottom, T0);
        c.set_dirichlet(bc);
        std::vector<double> heat(E, q);
        c.set_source(heat.data());
        IterParams ip;
        ip.maxerr = 1e-12; ip.maxit = 20000;
        ip.preconditioner = (ord & 1) ? IterParams::PRECOND_LJAC : IterParams::PRECOND_JAC;   // both preconditioners
        Context::LoopResult r = c.solve(true, ip, 0.05, 0);
        REQUIRE(ip.converged && r.loops >= 1);
        std::vector<double> T(N);
        c.get_field(T.data());
        const double H = (m.axis[2].back() - m.axis[2].front()) * 1e-6;
        double maxd = 0.;
        for (size_t i0 = 0; i0 < m.n(0); ++i0) for (size_t i1 = 0; i1 < m.n(1); ++i1) for (size_t i2 = 0; i2 < m.n(2); ++i2) {
            const double z = (m.axis[2][i2] - m.axis[2][0]) * 1e-6;
            const double ex = T0 + q * (2. * H * z - z * z) / (2. * k);
            maxd = std::fmax(maxd, std::fabs(T[m.node(i0, i1, i2)] - ex));
        }
        printf("order %d: loops %d, PCG iterations %lld, max|T - T_exact| = %.3e K (max T %.3f)\n", ord, r.loops, r.lin_iters, maxd, r.maxval);
        REQUIRE(maxd < 1e-6);
    }
    // ---- convection on the top plane, corrected form: T linear, k (T(H) - T0)/H = h (Ta - T(H))
    {
        Mesh m = make_mesh(6, 5, 17, ORDER_012);
        const double k = 40., h = 2e5, Ta = 350., T0 = 300.;
        Context c(0, "convection");
        c.set_mesh(m);
        std::vector<uint32_t> ids(m.elements(), 0);
        c.set_materials(ids, sample_tables(1, [&](uint32_t, double) { return std::make_pair(k, k); }, 250., 1., 400));
        c.fill_field(T0);
        Dirichlet bc;
        std::vector<size_t> bottom, top;
        for (size_t i0 = 0; i0 < m.n(0); ++i0) for (size_t i1 = 0; i1 < m.n(1); ++i1) {
            bottom.push_back(m.node(i0, i1, 0));
            top.push_back(m.node(i0, i1, m.n(2) - 1));
        }
        bc.add(bottom, T0);
        c.set_dirichlet(bc);
        c.set_source(nullptr);
        NodeConditions<1> none1;
        NodeConditions<2> conv, none2;
        conv.add(m.size(), top, {h, Ta});
        c.set_boundary(none1, conv, none2, false);
        IterParams ip;
        ip.maxerr = 1e-12; ip.maxit = 20000;
        c.solve(true, ip, 0.05, 1);
        REQUIRE(ip.converged);
        std::vector<double> T(m.size());
        c.get_field(T.data());
        const double H = (m.axis[2].back() - m.axis[2].front()) * 1e-6;
        const double TH = (k * T0 / H + h * Ta) / (k / H + h);
        double maxd = 0.;
        for (size_t i0 = 0; i0 < m.n(0); ++i0) for (size_t i1 = 0; i1 < m.n(1); ++i1) for (size_t i2 = 0; i2 < m.n(2); ++i2) {
            const double z = (m.axis[2][i2] - m.axis[2][0]) * 1e-6;
            maxd = std::fmax(maxd, std::fabs(T[m.node(i0, i1, i2)] - (T0 + (TH - T0) * z / H)));
        }
        printf("convection: max|T - T_exact| = %.3e K (T(H) = %.4f)\n", maxd, TH);
        REQUIRE(maxd < 1e-6);
    }
    // ---- device-resident field exchange (ThermoElectric meta loop): heat taken from the electrical context on the device
    //      equals heat downloaded and passed back through set_source
    {
        Mesh m = make_mesh(8, 7, 15, ORDER_012);
        const size_t E = m.elements();
        std::vector<size_t> bottom, top;
        for (size_t i0 = 0; i0 < m.n(0); ++i0) for (size_t i1 = 0; i1 < m.n(1); ++i1) {
            bottom.push_back(m.node(i0, i1, 0));
            top.push_back(m.node(i0, i1, m.n(2) - 1));
        }
        std::vector<uint32_t> ids(E, 0);
        Context el(0, "electrical");
        el.set_layout(PFEM_LAYOUT_VERTICAL_MINOR);
        el.set_mesh(m);
        el.set_materials(ids, sample_tables(1, [](uint32_t, double T) { return std::make_pair(1e4 * 300. / T, 2e4 * 300. / T); }, 250., 1., 400));
        el.fill_field(0.);
        Dirichlet bv;
        bv.add(bottom, 0.);
        bv.add(top, 0.3);
        el.set_dirichlet(bv);
        el.set_source(nullptr);
        el.set_elem_temperature(nullptr, 300.);
        el.set_noheat(std::vector<uint8_t>());
        IterParams ip;
        ip.maxerr = 1e-12; ip.maxit = 20000;
        el.solve(false, ip, 0.05, 1);
        REQUIRE(ip.converged);
        std::vector<double> heat(E);
        el.get_elem(PFEM_ELEM_HEAT, heat.data());
        double hmax = 0.;
        for (double h : heat) hmax = std::fmax(hmax, h);
        REQUIRE(hmax > 0.);
        std::vector<double> T[2];
        for (int via_device = 0; via_device < 2; ++via_device) {
            Context th(0, "thermal");
            th.set_mesh(m);
            th.set_materials(ids, sample_tables(1, [](uint32_t, double) { return std::make_pair(44., 44.); }, 250., 1., 400));
            th.fill_field(300.);
            Dirichlet bt;
            bt.add(bottom, 300.);
            th.set_dirichlet(bt);
            if (via_device) th.take_heat_from(el); else th.set_source(heat.data());
            IterParams it;
            it.maxerr = 1e-12; it.maxit = 20000;
            it.preconditioner = IterParams::PRECOND_LJAC;
            th.solve(true, it, 0.05, 1);
            REQUIRE(it.converged);
            T[via_device].resize(m.size());
            th.get_field(T[via_device].data());
            if (via_device) el.take_temperature_from(th);     // and back: temperatures at the electrical element midpoints
        }
        double maxd = 0., maxT = 0.;
        for (size_t i = 0; i < m.size(); ++i) { maxd = std::fmax(maxd, std::fabs(T[0][i] - T[1][i])); maxT = std::fmax(maxT, T[1][i]); }
        printf("field exchange: max|T_device - T_host| = %.3e K, max T = %.4f K\n", maxd, maxT);
        REQUIRE(maxd < 1e-9 && maxT > 300.);
    }
    // ---- noconv policy
    {
        Mesh m = make_mesh(9, 7, 21, ORDER_012);
        Context c(0, "noconv");
        c.set_mesh(m);
        std::vector<uint32_t> ids(m.elements(), 0);
        c.set_materials(ids, sample_tables(1, [](uint32_t, double T) { return std::make_pair(45. * std::pow(300. / T, 1.28), 45. * std::pow(300. / T, 1.28)); }));
        c.fill_field(300.);
        Dirichlet bc;
        std::vector<size_t> bottom;
        for (size_t i0 = 0; i0 < m.n(0); ++i0) for (size_t i1 = 0; i1 < m.n(1); ++i1) bottom.push_back(m.node(i0, i1, 0));
        bc.add(bottom, 300.);
        c.set_dirichlet(bc);
        std::vector<double> heat(m.elements(), 1e16);
        c.set_source(heat.data());
        IterParams ip;
        ip.maxit = 3; ip.maxerr = 1e-12;
        ip.no_convergence_behavior = IterParams::NO_CONVERGENCE_ERROR;
        bool thrown = false;
        try { c.solve(true, ip, 0.05, 1); } catch (const ComputationError& e) { thrown = strstr(e.what(), "Failed to converge") != nullptr; }
        REQUIRE(thrown && !ip.converged && ip.iters == 3);
        int warned = 0;
        ip.no_convergence_behavior = IterParams::NO_CONVERGENCE_WARNING;
        c.solve(true, ip, 0.05, 1, [&](int lvl, const std::string&) { if (lvl == 1) ++warned; });
        REQUIRE(warned == 1);
        bool bad = false;
        try { c.set_source(nullptr); IterParams z; z.maxit = 0; c.solve(true, z, 0.05, 1); } catch (const BadInput&) { bad = true; }
        REQUIRE(bad);
    }
    // ---- Dynamic3D: 1-D cooling, bottom at 300 K, top insulated, T(z,0) = 300 + a sin(pi z / 2L):
    //      T(z,t) = 300 + a sin(pi z / 2L) exp(-alpha (pi/2L)^2 t), Crank-Nicolson, lumped capacity
    {
        Mesh m;
        const size_t nz = 41;
        const double L = 2.0, k = 45., cprho = 0.327e3 * 5.31749e3, a = 50.;
        for (size_t i = 0; i < 3; ++i) m.axis[0].push_back(0.5 * i);
        for (size_t i = 0; i < 4; ++i) m.axis[1].push_back(0.5 * i);
        for (size_t i = 0; i < nz; ++i) m.axis[2].push_back(L * i / (nz - 1));
        m.order = ORDER_012;
        Context c(0, "dynamic");
        c.set_mesh(m);
        std::vector<uint32_t> ids(m.elements(), 0);
        Tables t = sample_tables(1, [&](uint32_t, double) { return std::make_pair(k, k); }, 250., 1., 400);
        c.set_materials(ids, t);
        c.set_capacity(t, std::vector<double>(t.nT, cprho));
        std::vector<double> T(m.size());
        for (size_t i0 = 0; i0 < m.n(0); ++i0) for (size_t i1 = 0; i1 < m.n(1); ++i1) for (size_t i2 = 0; i2 < nz; ++i2)
            T[m.node(i0, i1, i2)] = 300. + a * std::sin(M_PI * m.axis[2][i2] / (2. * L));
        c.set_field(T.data());
        Dirichlet bc;
        std::vector<size_t> bottom;
        for (size_t i0 = 0; i0 < m.n(0); ++i0) for (size_t i1 = 0; i1 < m.n(1); ++i1) bottom.push_back(m.node(i0, i1, 0));
        bc.add(bottom, 300.);
        c.set_dirichlet(bc);
        c.set_source(nullptr);
        IterParams ip;
        ip.maxerr = 1e-10; ip.maxit = 5000; ip.preconditioner = IterParams::PRECOND_LJAC;
        double elapsed = 0.;
        int lines = 0;
        Context::TimeResult r = c.solve_dynamic(ip, 200., 4., 0.5, true, 0, 10, elapsed, [&](int lvl, const std::string& s) { if (lvl == 3 && s.rfind("Time", 0) == 0) ++lines; });
        printf("Dynamic3D: steps %d, elapsed %.3f, log lines %d, converged %d, err %.3e\n", r.steps, elapsed, lines, (int)ip.converged, ip.err);
        REQUIRE(r.steps == 51 && std::fabs(elapsed - 200.) < 1e-9 && lines == 5 && ip.converged);
        c.get_field(T.data());
        const double alpha = k / cprho * 1e3, rate = alpha * (M_PI / (2. * L)) * (M_PI / (2. * L));
        double maxd = 0.;
        for (size_t i2 = 0; i2 < nz; ++i2)
            maxd = std::fmax(maxd, std::fabs(T[m.node(1, 1, i2)] - (300. + a * std::sin(M_PI * m.axis[2][i2] / (2. * L)) * std::exp(-rate * (elapsed + 4.)))));   // 51 solves of 4 ns
        printf("Dynamic3D cooling: %d steps, %lld PCG iterations, max|T - T_exact| = %.3e K\n", r.steps, r.lin_iters, maxd);
        REQUIRE(maxd < 4e-3);
    }
    printf("adapter gpu tests ok\n");
    return 0;
}

// adapter_test thickness <order> <in> <out>: layer_thickness + material_ids on a stack given by the test
// (in: n[3] uint64, the three axes, elem_mat[E] uint32; out: thickness[E] double, ids[E] uint32), compared with the oracle's
// orc_thickness by tests/test_thickness_tables.py
static int thickness_tool(const char* order, const char* fin, const char* fout) {
    static const char* names[6] = {"012", "021", "102", "120", "201", "210"};
    static const IterationOrder orders[6] = {ORDER_012, ORDER_021, ORDER_102, ORDER_120, ORDER_201, ORDER_210};
    Mesh m;
    int oi = -1;
    for (int k = 0; k < 6; ++k) if (!strcmp(order, names[k])) oi = k;
    REQUIRE(oi >= 0);
    m.order = orders[oi];
    FILE* f = fopen(fin, "rb");
    REQUIRE(f != nullptr);
    uint64_t n[3];
    REQUIRE(fread(n, 8, 3, f) == 3);
    for (int a = 0; a < 3; ++a) { m.axis[a].resize(n[a]); REQUIRE(fread(m.axis[a].data(), 8, n[a], f) == n[a]); }
    std::vector<uint32_t> mat(m.elements());
    REQUIRE(fread(mat.data(), 4, mat.size(), f) == mat.size());
    fclose(f);
    std::vector<double> th = layer_thickness(m, [&](size_t i0, size_t i1, size_t r) { return mat[m.elem(i0, i1, r)]; });
    std::vector<uint32_t> ids = material_ids(mat, th);
    f = fopen(fout, "wb");
    REQUIRE(f != nullptr);
    fwrite(th.data(), 8, th.size(), f);
    fwrite(ids.data(), 4, ids.size(), f);
    fclose(f);
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 4 && !strcmp(argv[1], "thickness")) return thickness_tool(argv[2], argv[3], argv[4]);
    int rc = host_tests();
    if (rc) return rc;
    if (argc > 1 && !strcmp(argv[1], "gpu")) rc = gpu_tests();
    return rc;
}
