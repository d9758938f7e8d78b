/* plaskdiff_cuda.h — C ABI of the carrier-diffusion part of libplaskfem_cuda.so: the B200 (sm_100a, FP64) replacement for the
 * hot path of PLaSK's electrical.diffusion.Diffusion3D (the last "next" row).
 *
 * What it replaces (paths relative to the PLaSK source tree): the body of the while(true) loop of
 *   Diffusion3DSolver::compute(loops, shb, act)          solvers/electrical/diffusion/diffusion3d.cpp:283-366
 * for ONE active region, i.e.
 *   - K->clear(); F.fill(0.); setLocalMatrix(...) for every element      diffusion3d.cpp:284-289, 196-199 (diffusion3d-eval.ipp)
 *   - addLocalBurningMatrix(...) for every mode and element (shb)        diffusion3d.cpp:293-308, 201-204 (diffusion3d-eval-shb.ipp)
 *   - the residual  err = 100 |K U - F| / |F|                             diffusion3d.cpp:341-349
 *   - K->solve(F, active.U)                                               diffusion3d.cpp:358 (DpbMatrix / DgbMatrix / SparseFreeMatrix,
 *                                                                          diffusion3d.cpp:275-279)
 * and the spline evaluation of outCarriersConcentration                   diffusion3d.cpp:420-458.
 * The plugin keeps what is geometry: setupActiveRegions (diffusion3d.cpp:89-178), the vertical averaging of inTemperature /
 * inLightE over the quantum wells (diffusion3d.hpp:92-105), the material calls A(T), B(T), C(T), D(T), Nr (diffusion3d.cpp:222-230,
 * 258-262) and the receivers inCurrentDensity / inGain (diffusion3d.cpp:232-238, 295-296).  INTEGRATION.md 10 shows the binding.
 *
 * The unknowns are those of the reference: three per node of the lateral mesh — value, d/dy, d/dx (ElementParams3D,
 * diffusion3d.hpp:108-141) — on the 12-function Hermite element.  The element integrals are evaluated on the device by 7x7
 * Gauss-Legendre quadrature, which is exact for them (the reference has them in closed form; tests compare the two to 1e-13).
 * The linear system is solved by a block-Jacobi (3x3 per node) preconditioned conjugate gradient kept on the device for the whole
 * whole loop of compute() (ONE cooperative launch per call: assembly, residual, decision and the solves of all Newton loops).
 *
 * Conventions: as plaskfem_cuda.h (status codes = pfem_status, host pointers copied during the call, FP64, no CPU fallback, not
 * thread-safe per context).  Arrays use the FULL lateral grid numbering of RectangularMesh2D(lon, tran) — node = index_f(i0, i1)
 * of the given iteration order (rectangular2d.cpp:21-33), element likewise on (n0-1) x (n1-1) — NOT the masked numbering of
 * RectangularMaskedMesh2D: elements outside the active region are flagged in `elem_active`, nodes that touch no active element are
 * not unknowns and read back as 0 (plaskdiff::MaskedNumbering2D in plaskdiff_cuda.hpp maps between the two numberings).
 */
#ifndef PLASKDIFF_CUDA_H
#define PLASKDIFF_CUDA_H

#include <stddef.h>
#include <stdint.h>

#include "plaskfem_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pdiff_ctx pdiff_ctx;

/* ---- life cycle ---------------------------------------------------------------------------------------------------------- */
int pdiff_create(pdiff_ctx** ctx, int device);
void pdiff_destroy(pdiff_ctx* ctx);
const char* pdiff_last_error(const pdiff_ctx* ctx);

/* ---- problem description -------------------------------------------------------------------------------------------------- */

/* Lateral mesh of the active region (ActiveRegion3D::mesh2, diffusion3d.hpp:73-80): n0 x n1 nodes, coordinates in um;
 * order = PDIFF_ORDER_01 (axis 1 fastest: the default of RectangularMesh2D(axis0, axis1), rectangular2d.hpp:313) or PDIFF_ORDER_10.
 * elem_active[(n0-1)(n1-1)] != 0 marks the elements of the masked mesh (role QW / QD / carriers at the well's height); NULL = all.
 * Resets everything else in the context (U = 0, like `if (!active.U) active.U.reset(N, 0.)`, diffusion3d.cpp:220). */
#define PDIFF_ORDER_01 0
#define PDIFF_ORDER_10 1
int pdiff_set_mesh(pdiff_ctx* ctx, size_t n0, size_t n1, const double* ax0, const double* ax1, int order, const uint8_t* elem_active);

/* Per element: A [1/s], B [cm^3/s], C [cm^6/s], D [um^2/s] (= 1e8 * material->D(T), diffusion3d.cpp:226-229). */
int pdiff_set_parameters(pdiff_ctx* ctx, const double* A, const double* B, const double* C, const double* D);

/* Per node: J = |js * j.c2| with js = 1e7 / (qe * QWheight) (diffusion3d.cpp:232-238). */
int pdiff_set_current(pdiff_ctx* ctx, const double* J);

/* Spatial hole burning (compute(loops, shb = true)): for mode m = 0..nmodes-1
 *   P[m][node][2]  = (c00, c11) of Ps[m] (diffusion3d.cpp:264-271),
 *   G[m][elem][2]  = factor * nrs[m][e] * gain[e]   (c00, c11),  dG likewise from dgdn   (diffusion3d.cpp:289-294).
 * nmodes = 0 switches it off.  The burned power modesP (diffusion3d.cpp:295,303) does not involve the unknowns and stays with the
 * host (plaskdiff::burned_power). */
int pdiff_set_modes(pdiff_ctx* ctx, size_t nmodes, const double* P, const double* G, const double* dG);

/* The unknowns active.U, 3 per node of the full grid (value, d/dy, d/dx); NULL in set = zeros. */
int pdiff_set_concentration(pdiff_ctx* ctx, const double* U);
int pdiff_get_concentration(pdiff_ctx* ctx, double* U);

/* ---- the loop ------------------------------------------------------------------------------------------------------------- */
typedef struct {
    int loops;        /* maximum number of loops, 0 = until err < maxerr (compute(loops, ...)) */
    double maxerr;    /* [%] Diffusion3DSolver::maxerr, default 0.05 */
    int maxit;        /* PCG iterations per linear solve */
    double lin_tol;   /* a linear solve stops at sqrt(r.M^-1 r / F.M^-1 F) <= lin_tol, r = F - K U, M = the nodal 3x3 blocks of K */
    int verbatim;     /* 1: Ug as written in diffusion3d.cpp:296-300 (X and Y exchanged between the slope unknowns), 0: the Hermite
                         interpolant at the element centre.  Identical on square elements. */
    int loop_cap;     /* with loops = 0: give up (PFEM_NOT_CONVERGED) after this many loops; the reference would not return.  Default 10000 */
    int reserved[6];
} pdiff_opts;

typedef struct {
    int loops;              /* loops run by this call (the `loop` counter) */
    int converged;          /* err < maxerr reached */
    double err;             /* last err [%] */
    long long lin_iters;    /* PCG iterations, all loops */
    int last_iters;
    double lin_relres;      /* |F - K U| / |F| at the end of the last linear solve (recurrence residual) */
    double t_solve_ms;      /* device time of the call */
    long long kernel_launches;
    double err_log[64];     /* err of loop 1, 2, ... (first 64) */
    double lin_relres_precond; /* the same in the norm of the stopping rule */
    double reserved[3];
} pdiff_stats;

void pdiff_default_opts(pdiff_opts* opts);
/* Returns PFEM_OK, PFEM_NOT_CONVERGED (a linear solve hit maxit, or loop_cap), PFEM_ERR_NOT_SPD (p.Kp <= 0 or a nodal block that is
 * not positive definite: dpbtrf would fail too), PFEM_ERR_NAN, ... */
int pdiff_compute(pdiff_ctx* ctx, const pdiff_opts* opts, pdiff_stats* stats);

/* ---- results -------------------------------------------------------------------------------------------------------------- */

/* outCarriersConcentration at npts lateral points (already wrapped into the mesh by the geometry's symmetry flags, already
 * restricted to the quantum wells' vertical range — ConcentrationDataImpl::at, diffusion3d.cpp:462-484).
 * method PDIFF_INTERP_SPLINE: the Hermite interpolant (diffusion3d.cpp:420-458); PDIFF_INTERP_LINEAR: bilinear interpolation of the
 * nodal values (the `else` branch, :460-474).  Points outside the mesh or inside an inactive element give 0. */
#define PDIFF_INTERP_SPLINE 0
#define PDIFF_INTERP_LINEAR 1
int pdiff_interpolate(pdiff_ctx* ctx, size_t npts, const double* x, const double* y, int method, double* out);

/* ---- parity hooks (tests; they assemble at the current U with the given `verbatim`) ----------------------------------------- */
/* K[elem][12][12], F[elem][12] of every element (zeros for inactive ones) in the local numbering of ElementParams3D:
 * local node n00, n01, n10, n11 (first digit = axis 0 side), unknown 3 * local node + (0 value, 1 d/dy, 2 d/dx). */
int pdiff_get_element_matrices(pdiff_ctx* ctx, int verbatim, double* K, double* F);
/* y = K v and the load vector F of the assembled system at the current U (3 per node; rows of non-unknowns: y = v, F = 0). */
int pdiff_apply(pdiff_ctx* ctx, int verbatim, const double* v, double* y);
int pdiff_get_rhs(pdiff_ctx* ctx, int verbatim, double* F);

#ifdef __cplusplus
}
#endif
#endif
