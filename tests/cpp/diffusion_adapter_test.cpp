// diffusion_adapter_test.cpp — exercises include/plaskdiff_cuda.hpp the way Diffusion3DSolver::compute would.
//   diffusion_adapter_test host   masked numbering, burned power (verbatim and corrected), NoDevice mapping
//   diffusion_adapter_test gpu    the uniform case of solvers/electrical/diffusion/tests/diffusion3d.py:86-94 on a quarter disc,
//                                 all arrays in the MASKED numbering the solver holds
#include <cmath>
#include <cstdio>
#include <cstring>

#include "plaskdiff_cuda.hpp"

using namespace plaskdiff;

#define REQUIRE(c) do { if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static int host_tests() {
    // 4 x 3 nodes, elements (0,0) and (2,1) excluded
    auto inc = [](size_t i0, size_t i1) { return !((i0 == 0 && i1 == 0) || (i0 == 2 && i1 == 1)); };
    MaskedNumbering2D a(4, 3, PDIFF_ORDER_01, inc), b(4, 3, PDIFF_ORDER_10, inc);
    REQUIRE(a.elements() == 4 && b.elements() == 4);
    REQUIRE(a.nodes() == 10 && b.nodes() == 10);                      // node (0,0) and node (3,2) touch no kept element
    REQUIRE(a.node_of_full[a.node(0, 0)] == MaskedNumbering2D::NONE && a.node_of_full[a.node(3, 2)] == MaskedNumbering2D::NONE);
    REQUIRE(a.node_of_full[a.node(0, 1)] == 0 && a.node_of_full[a.node(3, 1)] == 9);
    REQUIRE(b.node_of_full[b.node(1, 0)] == 0 && b.node_of_full[b.node(2, 2)] == 9);   // axis 0 fastest
    REQUIRE(a.elem_of_full[a.elem(0, 1)] == 0 && a.elem_of_full[a.elem(2, 0)] == 3);
    double m[10], back[10];
    for (int i = 0; i < 10; ++i) m[i] = 1. + i;
    auto full = a.nodes_to_full(m, 1, -1.);
    REQUIRE(full.size() == 12 && full[a.node(0, 0)] == -1. && full[a.node(0, 1)] == 1.);
    a.nodes_to_masked(full.data(), back);
    for (int i = 0; i < 10; ++i) REQUIRE(back[i] == m[i]);

    // burned power: 2 elements, 6 nodes in a row of the masked numbering
    const double P[12] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12}, g[4] = {2, 3, 5, 7}, X[2] = {0.5, 0.25}, Y[2] = {2., 4.};
    const size_t corner[8] = {0, 1, 2, 3, 2, 3, 4, 5};
    const double vb = burned_power(2, 6, P, g, X, Y, corner, 0.006, true);
    // element 0: P[0..3] -> c00 1+3+5+7 = 16, c11 2+4+6+8 = 20, * 0.25 * 0.25 ; element 1: P[1..4] -> 3+5+7+9 = 24, 28, * 0.25 * 0.0625
    const double want_vb = ((16 * 0.0625) * 2 + (20 * 0.0625) * 3 + (24 * 0.015625) * 5 + (28 * 0.015625) * 7) * 1e-13 * 0.006;
    REQUIRE(std::fabs(vb - want_vb) <= 1e-15 * want_vb);
    const double co = burned_power(2, 6, P, g, X, Y, corner, 0.006, false);
    const double want_co = ((16 * 0.25) * 2 + (20 * 0.25) * 3 + ((5 + 7 + 9 + 11) * 0.25) * 5 + ((6 + 8 + 10 + 12) * 0.25) * 7) * 1e-13 * 0.006;
    REQUIRE(std::fabs(co - want_co) <= 1e-15 * want_co);

    if (pfem_device_count() == 0) {
        bool thrown = false;
        try {
            Region r("diff", {0., 1.}, {0., 1.}, PDIFF_ORDER_01, [](size_t, size_t) { return true; });
        } catch (const NoDevice&) { thrown = true; }
        REQUIRE(thrown);
    }
    printf("diffusion adapter host tests ok\n");
    return 0;
}

static int gpu_tests() {
    const double A = 3e7, B = 1.7e-10, C = 6e-27, D = 10., L = 4.0, n0c = 1.0e19;
    const size_t n = 61;
    std::vector<double> ax(n);
    for (size_t i = 0; i < n; ++i) ax[i] = L * i / (n - 1);
    auto inside = [&](size_t i0, size_t i1) {
        const double x = 0.5 * (ax[i0] + ax[i0 + 1]), y = 0.5 * (ax[i1] + ax[i1 + 1]);
        return x * x + y * y <= L * L;
    };
    for (int order : {PDIFF_ORDER_01, PDIFF_ORDER_10}) {
        Region reg("diffusion3d", ax, ax, order, inside);
        const auto& num = reg.numbering();
        const size_t ne = num.elements(), nn = num.nodes();
        REQUIRE(ne < (n - 1) * (n - 1) && nn < n * n);
        std::vector<double> a(ne, A), b(ne, B), c(ne, C), d(ne, 1e8 * D), J(nn, A * n0c + B * n0c * n0c + C * n0c * n0c * n0c);
        reg.set_parameters(a.data(), b.data(), c.data(), d.data());
        reg.set_current(J.data());
        pdiff_stats st;
        const int rc = reg.compute(0, 1e-4, st);
        REQUIRE(rc == PFEM_OK && st.converged && st.err < 1e-4 && st.kernel_launches == 1);
        std::vector<double> U(3 * nn);
        reg.get_U(U.data());
        for (size_t k = 0; k < nn; ++k) {
            REQUIRE(std::fabs(U[3 * k] / n0c - 1.) < 1e-7);
            REQUIRE(std::fabs(U[3 * k + 1]) < 1e-6 * n0c && std::fabs(U[3 * k + 2]) < 1e-6 * n0c);
        }
        const double x[3] = {0.3, 3.9, 3.2}, y[3] = {0.2, 0.1, 3.2};
        double out[3];
        reg.interpolate(3, x, y, true, out);
        REQUIRE(std::fabs(out[0] / n0c - 1.) < 1e-7 && std::fabs(out[1] / n0c - 1.) < 1e-7 && out[2] == 0.);   // (3.2, 3.2) is outside
        // a second call starts from the converged U: one residual evaluation, no solve
        REQUIRE(reg.compute(0, 1e-4, st) == PFEM_OK && st.loops == 1 && st.lin_iters == 0);
        printf("order %d: %d loops, %lld PCG iterations, %.2f ms\n", order, st.loops, st.lin_iters, st.t_solve_ms);
    }
    bool thrown = false;
    try {
        Region bad("diff", {0., 1., 0.5}, {0., 1.}, PDIFF_ORDER_01, [](size_t, size_t) { return true; });
    } catch (const BadInput&) { thrown = true; }
    REQUIRE(thrown);
    printf("diffusion adapter gpu tests ok\n");
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    if (host_tests()) return 1;
    if (!std::strcmp(argv[1], "gpu")) return gpu_tests();
    return 0;
}
