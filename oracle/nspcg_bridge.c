/* nspcg_bridge.c — drives the REFERENCE's own NSPCG (extlib/nspcg/nspcg.c, compiled in
 * place from /root/reference into oracle/_ref/) exactly the way PLaSK does in
 * SparseMatrix::solverhs, plask/common/fem/iterative_matrix.hpp:141-339.
 * TEST INFRASTRUCTURE; never linked into the product. */
#include <stdlib.h>
#include <string.h>

typedef int (*nspcg_fn)();
extern int nspcg_(nspcg_fn precon, nspcg_fn accel, int* ndim, int* mdim, int* n, int* maxnz, double* coef, int* jcoef,
                  int* p, int* ip, double* u, double* ubar, double* rhs, double* wksp, int* iwksp, int* nw, int* inw,
                  int* iparm, double* rparm, int* ier);
extern int dfault_(int* iparm, double* rparm);
extern int rich2_(), jac2_(), ljac2_(), ljacx2_(), sor2_(), ssor2_(), ic2_(), mic2_(), lsp2_(), neu2_(), lsor2_(),
    lssor2_(), llsp2_(), lneu2_(), bic2_(), bicx2_(), mbic2_(), mbicx2_();
extern int cg_(), si_(), sor_(), srcg_(), srsi_();

/* persistent state of one SparseBandMatrix: workspace and the `ifact` countdown
 * (iterative_matrix.hpp:106-110,148-150) */
typedef struct {
    int ifact;
    int nw, inw;
    double* wksp;
    int* iwksp;
} ref_nspcg_state;

ref_nspcg_state* ref_nspcg_new(void) {
    ref_nspcg_state* s = (ref_nspcg_state*)calloc(1, sizeof(ref_nspcg_state));
    s->ifact = 1;
    return s;
}
void ref_nspcg_free(ref_nspcg_state* s) {
    if (!s) return;
    free(s->wksp);
    free(s->iwksp);
    free(s);
}

/* precond: index into IterativeMatrixParams::Preconditioner (iterative_matrix.hpp:53-72);
 * accel: 0 cg, 1 si (others unused by this path).  coef = 14*n doubles (symmetric
 * diagonal storage, nstore=2), jcoef = the 14 offsets.  u: initial guess in, solution out.
 * Returns NSPCG's ier; *iters / *err receive iparm.itmax / rparm.zeta on exit
 * (iterative_matrix.hpp:335-336). */
int ref_nspcg_solve(ref_nspcg_state* s, int precond, int accel, int n, int major, int minor, double* coef, int* jcoef,
                    double* u, double* rhs, int itmax, double zeta, int nfact, int* iters, double* err,
                    double* timfac, double* timtot) {
    static nspcg_fn const pre[18] = {rich2_, jac2_,  ljac2_, ljacx2_, sor2_,  ssor2_, ic2_,   mic2_,  lsp2_,
                                     neu2_,  lsor2_, lssor2_, llsp2_, lneu2_, bic2_,  bicx2_, mbic2_, mbicx2_};
    static nspcg_fn const acc[5] = {cg_, si_, sor_, srcg_, srsi_};
    int iparm[30];
    double rparm[30];
    memset(iparm, 0, sizeof iparm);
    memset(rparm, 0, sizeof rparm);
    dfault_(iparm, rparm);
    int mdim = 14, ndim = n, maxnz = 14;
    iparm[11] = 2;                              /* nstore */
    iparm[1] = itmax;                           /* itmax  */
    iparm[17] = 0;                              /* ipropa */
    iparm[14] = (--s->ifact) ? 0 : 1;           /* ifact  */
    if (s->ifact <= 0) s->ifact = nfact;
    rparm[0] = zeta;
    iparm[18] = minor - 1;                      /* kblsz, iterative_matrix.hpp:388 */
    iparm[19] = major - minor - 1;              /* nbl2d, :389 */
    iparm[2] = -1;                              /* level (release build) */
    int kblsz = minor - 1;
    size_t default_nw = 3 * (size_t)n + 2 * (size_t)itmax + (size_t)n * maxnz + (kblsz > 1 ? kblsz : 1);
    size_t default_inw = maxnz + ((2 * n > maxnz * maxnz + maxnz) ? 2 * n : maxnz * maxnz + maxnz);
    if ((size_t)s->nw < default_nw) {
        s->nw = (int)default_nw;
        free(s->wksp);
        s->wksp = (double*)malloc((size_t)s->nw * sizeof(double));
        iparm[14] = 1;
    }
    if ((size_t)s->inw < default_inw) {
        s->inw = (int)default_inw;
        free(s->iwksp);
        s->iwksp = (int*)malloc((size_t)s->inw * sizeof(int));
        iparm[14] = 1;
    }
    int ier = 0;
    for (;;) {
        nspcg_(pre[precond], acc[accel], &ndim, &mdim, &n, &maxnz, coef, jcoef, NULL, NULL, u, NULL, rhs, s->wksp,
               s->iwksp, &s->nw, &s->inw, iparm, rparm, &ier);
        if (ier == -2 && s->nw) {
            free(s->wksp);
            s->wksp = (double*)malloc((size_t)s->nw * sizeof(double));
            iparm[14] = 1;
        } else if (ier == -3 && s->inw) {
            free(s->iwksp);
            s->iwksp = (int*)malloc((size_t)s->inw * sizeof(int));
            iparm[14] = 1;
        } else
            break;
    }
    if (iters) *iters = iparm[1];
    if (err) *err = rparm[0];
    if (timfac) *timfac = rparm[12];
    if (timtot) *timtot = rparm[13];
    return ier;
}
