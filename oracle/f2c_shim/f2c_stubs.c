/* The handful of libf2c runtime helpers that the reference's nspcg.c links against
 * (SURVEY.md §8c), written from their documented f2c semantics.  TEST INFRASTRUCTURE
 * (part of oracle/_ref); never linked into the product.  The formatted-I/O trio is
 * only reached when iparm.level >= 0; PLaSK sets -1 in release builds
 * (plask/common/fem/iterative_matrix.hpp:169-173), and so does nspcg_bridge.c. */
#include <math.h>
#include "libf2c/f2c.h"

double d_lg10(doublereal* x) { return log10(*x); }
double d_sign(doublereal* a, doublereal* b) { double x = fabs(*a); return *b >= 0 ? x : -x; }
integer i_sign(integer* a, integer* b) { integer x = *a >= 0 ? *a : -*a; return *b >= 0 ? x : -x; }
double pow_dd(doublereal* a, doublereal* b) { return pow(*a, *b); }
double pow_di(doublereal* a, integer* n) {
    double x = *a, r = 1.;
    long k = *n;
    if (k < 0) { k = -k; x = 1. / x; }
    for (; k; k >>= 1, x *= x) if (k & 1) r *= x;
    return r;
}
integer s_wsfe(cilist* a) { (void)a; return 0; }
integer do_fio(integer* n, char* p, ftnlen l) { (void)n; (void)p; (void)l; return 0; }
integer e_wsfe(void) { return 0; }
