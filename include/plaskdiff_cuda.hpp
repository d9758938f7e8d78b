// plaskdiff_cuda.hpp — header-only C++ host adapter over the C ABI of the Diffusion3D path (plaskdiff_cuda.h).
//
// This is the code a maintainer adds on the plugin side of solvers/electrical/diffusion/diffusion3d.cpp: it turns
// what Diffusion3DSolver::compute already holds for one active region — the masked lateral mesh, A/B/C/D per element of the MASKED
// mesh, J per node of the MASKED mesh, active.U in the masked numbering (diffusion3d.cpp:208-240) — into the full-grid arrays of the
// ABI and back, and maps the status codes to the exceptions the solver throws.
#ifndef PLASKDIFF_CUDA_HPP
#define PLASKDIFF_CUDA_HPP

#include <cstdint>
#include <limits>
#include <string>
#include <vector>

#include "plaskdiff_cuda.h"
#include "plaskfem_cuda.hpp"

namespace plaskdiff {

using plaskfem::BadInput;
using plaskfem::ComputationError;
using plaskfem::NoDevice;

// RectangularMaskedMesh2D (the lateral mesh of ActiveRegion3D, diffusion3d.hpp:73-80) keeps the elements selected by the predicate
// and their nodes, both numbered in the order of the full mesh.  The ABI speaks the full grid; these are the two index maps.
struct MaskedNumbering2D {
    static constexpr size_t NONE = std::numeric_limits<size_t>::max();
    size_t n0, n1;
    int order;   // PDIFF_ORDER_01 / PDIFF_ORDER_10
    std::vector<uint8_t> elem_active;                 // full element numbering, for pdiff_set_mesh
    std::vector<size_t> node_of_full, elem_of_full;   // full index -> masked index or NONE
    std::vector<size_t> full_of_node, full_of_elem;   // masked index -> full index

    size_t node(size_t i0, size_t i1) const { return order == PDIFF_ORDER_01 ? i0 * n1 + i1 : i1 * n0 + i0; }
    size_t elem(size_t i0, size_t i1) const { return order == PDIFF_ORDER_01 ? i0 * (n1 - 1) + i1 : i1 * (n0 - 1) + i0; }

    // included(i0, i1): the element's midpoint has the role QW / QD / carriers at the height of the middle well
    template <typename Included>
    MaskedNumbering2D(size_t n0_, size_t n1_, int order_, Included included) : n0(n0_), n1(n1_), order(order_) {
        if (n0 < 2 || n1 < 2) throw BadInput("MaskedNumbering2D: the lateral mesh needs at least 2 x 2 nodes");
        elem_active.assign((n0 - 1) * (n1 - 1), 0);
        std::vector<uint8_t> used(n0 * n1, 0);
        for (size_t i0 = 0; i0 + 1 < n0; ++i0)
            for (size_t i1 = 0; i1 + 1 < n1; ++i1)
                if (included(i0, i1)) {
                    elem_active[elem(i0, i1)] = 1;
                    used[node(i0, i1)] = used[node(i0 + 1, i1)] = used[node(i0, i1 + 1)] = used[node(i0 + 1, i1 + 1)] = 1;
                }
        node_of_full.assign(used.size(), NONE);
        elem_of_full.assign(elem_active.size(), NONE);
        for (size_t i = 0; i < used.size(); ++i)
            if (used[i]) { node_of_full[i] = full_of_node.size(); full_of_node.push_back(i); }
        for (size_t e = 0; e < elem_active.size(); ++e)
            if (elem_active[e]) { elem_of_full[e] = full_of_elem.size(); full_of_elem.push_back(e); }
    }
    size_t nodes() const { return full_of_node.size(); }
    size_t elements() const { return full_of_elem.size(); }

    // masked <-> full, nc interleaved components per entry; entries outside the masked mesh get `fill`
    std::vector<double> nodes_to_full(const double* masked, int nc = 1, double fill = 0.) const {
        std::vector<double> full(node_of_full.size() * nc, fill);
        for (size_t k = 0; k < full_of_node.size(); ++k)
            for (int c = 0; c < nc; ++c) full[full_of_node[k] * nc + c] = masked[k * nc + c];
        return full;
    }
    std::vector<double> elems_to_full(const double* masked, int nc = 1, double fill = 0.) const {
        std::vector<double> full(elem_of_full.size() * nc, fill);
        for (size_t k = 0; k < full_of_elem.size(); ++k)
            for (int c = 0; c < nc; ++c) full[full_of_elem[k] * nc + c] = masked[k * nc + c];
        return full;
    }
    void nodes_to_masked(const double* full, double* masked, int nc = 1) const {
        for (size_t k = 0; k < full_of_node.size(); ++k)
            for (int c = 0; c < nc; ++c) masked[k * nc + c] = full[full_of_node[k] * nc + c];
    }
};

// modesP of one mode (diffusion3d.cpp:295-303) in the MASKED numbering the solver holds: P[node][2], g[elem][2] = nrs * gain,
// X[elem] = element size along axis 0 (Y only enters the corrected form).  verbatim = true is the reference as written:
// integrateBilinear(Lx, Ly, Pdata + ie) = 0.25 (P[ie] + P[ie+1] + P[ie+2] + P[ie+3]) Lx Lx (diffusion3d.hpp:170-172) — four consecutive
// nodal entries starting at the element index; verbatim = false takes the element's corner nodes (corner[elem][4]) and Lx Ly.
inline double burned_power(size_t nelem, size_t nnode, const double* P, const double* g, const double* X, const double* Y,
                           const size_t* corner, double qw_height, bool verbatim) {
    double tot = 0.;
    for (size_t e = 0; e < nelem; ++e) {
        double p0 = 0., p1 = 0.;
        if (verbatim) {
            for (size_t k = e; k < e + 4 && k < nnode; ++k) { p0 += P[2 * k]; p1 += P[2 * k + 1]; }
            p0 *= 0.25 * X[e] * X[e]; p1 *= 0.25 * X[e] * X[e];
        } else {
            for (int l = 0; l < 4; ++l) { p0 += P[2 * corner[4 * e + l]]; p1 += P[2 * corner[4 * e + l] + 1]; }
            p0 *= 0.25 * X[e] * Y[e]; p1 *= 0.25 * X[e] * Y[e];
        }
        tot += p0 * g[2 * e] + p1 * g[2 * e + 1];
    }
    return tot * 1e-13 * qw_height;
}

// One active region on the device (what `active.U` plus the matrix K and the vectors F, resid are in the reference).
class Region {
    pdiff_ctx* ctx_ = nullptr;
    std::string id_;
    MaskedNumbering2D num_;

    void check(int rc) const {
        if (rc >= 0) return;
        const std::string msg = id_ + ": " + pfem_strerror(rc) + ": " + pdiff_last_error(ctx_);
        switch (rc) {
            case PFEM_ERR_NO_DEVICE: throw NoDevice(msg);
            case PFEM_ERR_BAD_INPUT:
            case PFEM_ERR_STATE: throw BadInput(msg);
            default: throw ComputationError(msg);
        }
    }

  public:
    // axes of solver->getMesh()->lon() / tran(), the iteration order of the lateral mesh and the QW predicate of ActiveRegion3D
    template <typename Included>
    Region(const std::string& solver_id, const std::vector<double>& ax0, const std::vector<double>& ax1, int order, Included included,
           int device = 0)
        : id_(solver_id), num_(ax0.size(), ax1.size(), order, included) {
        int rc = pdiff_create(&ctx_, device);
        if (rc != PFEM_OK) { ctx_ = nullptr; throw NoDevice(id_ + ": " + pfem_strerror(rc)); }
        try {
            check(pdiff_set_mesh(ctx_, ax0.size(), ax1.size(), ax0.data(), ax1.data(), order, num_.elem_active.data()));
        } catch (...) {
            pdiff_destroy(ctx_);
            ctx_ = nullptr;
            throw;
        }
    }
    Region(const Region&) = delete;
    Region& operator=(const Region&) = delete;
    ~Region() { if (ctx_) pdiff_destroy(ctx_); }

    const MaskedNumbering2D& numbering() const { return num_; }

    // A, B, C, D (D already 1e8 * material->D(T)) per element and J per node of the MASKED mesh (diffusion3d.cpp:222-238)
    void set_parameters(const double* A, const double* B, const double* C, const double* D) {
        check(pdiff_set_parameters(ctx_, num_.elems_to_full(A).data(), num_.elems_to_full(B).data(), num_.elems_to_full(C).data(),
                                   num_.elems_to_full(D, 1, 1.).data()));
    }
    void set_current(const double* J) { check(pdiff_set_current(ctx_, num_.nodes_to_full(J).data())); }
    // modes: Ps[m] per masked node (c00, c11); G, dG per masked element, already factor * nrs * gain (diffusion3d.cpp:289-294)
    void set_modes(const std::vector<const double*>& P, const std::vector<const double*>& G, const std::vector<const double*>& dG) {
        std::vector<double> p, g, dg;
        for (size_t m = 0; m < P.size(); ++m) {
            auto a = num_.nodes_to_full(P[m], 2), b = num_.elems_to_full(G[m], 2), c = num_.elems_to_full(dG[m], 2);
            p.insert(p.end(), a.begin(), a.end()); g.insert(g.end(), b.begin(), b.end()); dg.insert(dg.end(), c.begin(), c.end());
        }
        check(pdiff_set_modes(ctx_, P.size(), p.data(), g.data(), dg.data()));
    }
    // active.U in the masked numbering, 3 per node
    void set_U(const double* U) { check(pdiff_set_concentration(ctx_, U ? num_.nodes_to_full(U, 3).data() : nullptr)); }
    void get_U(double* U) {
        std::vector<double> full(3 * num_.node_of_full.size());
        check(pdiff_get_concentration(ctx_, full.data()));
        num_.nodes_to_masked(full.data(), U, 3);
    }
    // the while(true) of compute(): returns the statistics; a linear solve that hit maxit is reported in the return value
    // (PFEM_NOT_CONVERGED) for the solver's `noconv` policy, everything else throws
    int compute(unsigned loops, double maxerr, pdiff_stats& st, bool verbatim = true, int maxit = 20000, double lin_tol = 1e-12) {
        pdiff_opts o;
        pdiff_default_opts(&o);
        o.loops = (int)loops; o.maxerr = maxerr; o.maxit = maxit; o.lin_tol = lin_tol; o.verbatim = verbatim ? 1 : 0;
        const int rc = pdiff_compute(ctx_, &o, &st);
        check(rc);
        return rc;
    }
    // ConcentrationDataImpl: lateral points already wrapped by InterpolationFlags; spline = INTERPOLATION_SPLINE / DEFAULT
    void interpolate(size_t npts, const double* x, const double* y, bool spline, double* out) {
        check(pdiff_interpolate(ctx_, npts, x, y, spline ? PDIFF_INTERP_SPLINE : PDIFF_INTERP_LINEAR, out));
    }
};

}  // namespace plaskdiff
#endif
