// oracle/diffusion_ref_shim.cpp — TEST INFRASTRUCTURE, not the product.
//
// Compiles the REFERENCE's own code-generated element integrals of Diffusion3DSolver where they lie:
//   $(REFERENCE)/solvers/electrical/diffusion/diffusion3d-eval.ipp       (body of setLocalMatrix,        diffusion3d.cpp:196-199)
//   $(REFERENCE)/solvers/electrical/diffusion/diffusion3d-eval-shb.ipp   (body of addLocalBurningMatrix,  diffusion3d.cpp:201-204)
// into oracle/_ref/libdiffusion_ref.so (`make ref`).  Nothing of the reference is copied: the two files are #included through the
// -I path of the Makefile; this shim only supplies the names their expressions use (`e`, `K(i,j)`, `F[]`, `U[]`, `J[]`, `P[]`,
// `G`, `dG`, `Ug`), i.e. ElementParams3D (diffusion3d.hpp:108-141), FemMatrix::operator() and Tensor2<double>.
//
// The element handed in is the unit of the comparison: local nodes 0..3 = n00 (lo,lo), n01 (lo,up), n10 (up,lo), n11 (up,up) and the
// unknown 3*node + c with c = 0 value, 1 d/dy ("i01"), 2 d/dx ("i10") — the numbering ElementParams3D builds.
#include <cstddef>

namespace {
struct Elem {
    std::size_t n00, n01, n10, n11, i00, i01, i10, i02, i03, i12, i20, i21, i30, i22, i23, i32;
    double X, Y;
    Elem(double X, double Y)
        : n00(0), n01(1), n10(2), n11(3), i00(0), i01(1), i10(2), i02(3), i03(4), i12(5), i20(6), i21(7), i30(8), i22(9), i23(10),
          i32(11), X(X), Y(Y) {}
};
// symmetric matrix: both (r,c) and (c,r) name the entry of the upper triangle, like DpbMatrix::index (cholesky_matrix.hpp:58-66)
struct Sym12 {
    double* a;
    double& operator()(std::size_t r, std::size_t c) { return r <= c ? a[12 * r + c] : a[12 * c + r]; }
};
struct T2 {
    double c00, c11;
};
}  // namespace

extern "C" {

// K[144] (row-major, upper triangle filled, mirrored below on return) and F[12] are ACCUMULATED into, like the reference does
void dref_local_matrix(double X, double Y, double A, double B, double C, double D, const double* U, const double* J, double* Kout,
                       double* F) {
    const Elem e(X, Y);
    Sym12 K{Kout};
#include "diffusion3d-eval.ipp"
    for (int r = 0; r < 12; ++r)
        for (int c = 0; c < r; ++c) Kout[12 * r + c] = Kout[12 * c + r];
}

// P: 4 nodal (c00,c11) pairs in local node order; G, dG: (c00,c11)
void dref_local_burning(double X, double Y, const double* G2, const double* dG2, double Ug, const double* P8, double* Kout, double* F) {
    const Elem e(X, Y);
    Sym12 K{Kout};
    const T2 G{G2[0], G2[1]}, dG{dG2[0], dG2[1]};
    const T2 P[4] = {{P8[0], P8[1]}, {P8[2], P8[3]}, {P8[4], P8[5]}, {P8[6], P8[7]}};
#include "diffusion3d-eval-shb.ipp"
    for (int r = 0; r < 12; ++r)
        for (int c = 0; c < r; ++c) Kout[12 * r + c] = Kout[12 * c + r];
}
}
