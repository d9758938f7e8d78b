/* Minimal f2c type shim used ONLY to compile the reference's vendored NSPCG
 * translation (extlib/nspcg/nspcg.c, compiled in place from /root/reference)
 * into oracle/_ref/.  TEST INFRASTRUCTURE — never linked into the product.
 *
 * Written from scratch: it declares just the names nspcg.c uses.  INTEGER is
 * 32-bit here because PLaSK passes plain `int` to NSPCG
 * (extlib/nspcg/nspcg.hpp:349, plask/common/fem/iterative_matrix.hpp:255) and
 * the Linux build compiles nspcg.f with default INTEGER*4.
 */
#ifndef ORACLE_F2C_SHIM_H
#define ORACLE_F2C_SHIM_H

typedef int integer;
typedef int logical;
typedef float real;
typedef double doublereal;
typedef int ftnlen;
typedef int ftnint;
typedef int flag;

#define TRUE_ (1)
#define FALSE_ (0)

#ifndef abs
#define abs(x) ((x) >= 0 ? (x) : -(x))
#endif
#define dabs(x) (doublereal) abs(x)
#ifndef min
#define min(a, b) ((a) <= (b) ? (a) : (b))
#endif
#ifndef max
#define max(a, b) ((a) >= (b) ? (a) : (b))
#endif
#define dmin(a, b) (doublereal) min(a, b)
#define dmax(a, b) (doublereal) max(a, b)

/* formatted-write control block (only reached when iparm.level >= 0) */
typedef struct {
    flag cierr;
    ftnint ciunit;
    flag ciend;
    char* cifmt;
    ftnint cirec;
} cilist;

#ifdef __cplusplus
typedef int (*U_fp)(...);
typedef int (*S_fp)(...);
typedef doublereal (*D_fp)(...);
typedef integer (*I_fp)(...);
typedef logical (*L_fp)(...);
#else
typedef int (*U_fp)();
typedef int (*S_fp)();
typedef doublereal (*D_fp)();
typedef integer (*I_fp)();
typedef logical (*L_fp)();
#endif
#define VOID void

#endif
