// kernels_ml.cuh — additive multilevel line preconditioner (pfem_opts::precond = 2, mirror: iterative.preconditioner = 'mlj').
//
// The reference's default preconditioner is IC(0) (iterative_matrix.hpp:73, nspcg.f:1697 `ic2`): a sequential
// factorisation that needs about 10x fewer CG iterations than line-Jacobi on layered devices.  Its strength does not
// come from the vertical couplings — the line blocks already invert those exactly — but from carrying information
// LATERALLY across the mesh in one application.  The GPU counterpart does that with a hierarchy of lateral aggregates:
//
//     M^-1 = sum_{l=0..L}  P_l T_l^-1 P_l^T
//
// level 0: the mesh itself, T_0 = tridiagonal blocks of A along the vertical mesh lines (the `ljac` preconditioner);
// level l: piecewise-constant aggregation P_l of 4^l x 4^l lateral columns of free nodes (vertical resolution kept),
//          T_l = vertical tridiagonal blocks of the Galerkin operator P_l^T A P_l; the top level is ONE column, i.e. the
//          1-D vertical problem of the whole device, solved exactly.
// Every term is symmetric positive (semi)definite, so plain PCG applies and — unlike a multiplicative cycle — no coarse
// OPERATOR is ever needed: a level is two arrays of line factors.  r'.z = sum_l r_l.z_l with r_l = P_l^T r', so the CG
// scalars come out of the level solves.  Measured on the first-loop systems (tools/study_multilevel.py): config B 48^3
// 272 iterations with line blocks -> 66; config C 408 -> 94.
//
// Per PCG iteration (vertical axis = I, the fastest axis; pfem_set_layout(PFEM_LAYOUT_VERTICAL_MINOR) guarantees that):
//   k_line_ml<FINE>   r' = r - alpha q, z_0 = T_0^-1 r', r_1 = P_1^T r'       (one block per 4x4 aggregate: the restriction
//                                                                              is a fixed-order sum in shared memory)
//   k_line_ml (x L)   z_l = T_l^-1 r_l, r_{l+1} = sum of 4x4 level-l lines, rho += r_l.z_l; the last one finalises beta / stop
//   k_fpcg<MODE 3>    p' = mask (z_0 + z_1(parent) + z_2(parent) + z_top) + beta p, x' = x + alpha p, q' = M A p'  (the coarse
//                     values are gathered by the operator kernel itself, L2-resident; there is no separate prolongation kernel)
#pragma once
#include "pfem_internal.cuh"

namespace pfem {

#define PFEM_ML_MAXL 3          // levels: 4x4 aggregates, 16x16 aggregates, one column
#define PFEM_ML_C 4             // aggregation factor per level and lateral axis
#define PFEM_ML_SHIFT 2

// lattice of one level: rows (j, k) of nI nodes with pitch sJ; level 0 is the mesh
struct LineDom {
    int nI, nJ, nK;
    idx_t sJ, sK;
    // slab mode (fine level only): rows k outside [k0, k1) belong to the neighbours and are skipped; row k lies in the level-1
    // aggregate (k + koff) >> 2.  koff = 16 - kown0 puts one empty 16-plane aggregate in front of the owned planes, so the halo
    // planes fall into aggregates of their own whose coarse z stays 0.  Single GPU: k0 = 0, k1 = nK, koff = 0.
    int k0, k1, koff;
};

struct MLDev {
    int nlev;                               // coarse levels 1..nlev, the LAST one is the top level = one column (0: not set up)
    int nagg;                               // aggregated levels with more than one column: nlev - 1 (at most 2: 4x4 and 16x16)
    LineDom dom[PFEM_ML_MAXL + 1];          // [0] = fine mesh
    double* r[PFEM_ML_MAXL + 1];            // restricted residual of the aggregated level l >= 1
    double* z[PFEM_ML_MAXL + 1];            // z_l, l >= 1 (the last aggregated level also receives z_top)
    double* ll[PFEM_ML_MAXL + 1];           // L D L^T factors of T_l, l >= 1 (before k_ml_factor: off-diagonals / diagonals)
    double* ld[PFEM_ML_MAXL + 1];
    double* part;                           // block totals for the top-level residual
    double* zero;                           // a zero row
};

// ---- set-up: tridiagonal blocks of the Galerkin operators -------------------------------------------------------
// One thread per fine node (i, j, k), block = 32 nodes along I x one 4x4 aggregate.  For every pair (n, n') of free nodes
// coupled by A with n' in the same plane or the plane above (di = 0, +1) the entry A[n, n'] (brick matrix,
// therm3d.cpp:227-237) is added to the level where n and n' first share an aggregate; a prefix sum over the levels and a
// fixed-order sum over the 16 nodes of the level-1 aggregate give, per level-1 aggregate and level l,
//     S[2(l-1)  ] = sum_{n in agg1, n' in agg_l(n), same i} A[n,n']      S[2(l-1)+1] = the same with n' at i+1.
// The diagonal entry itself is taken from the Jacobi diagonal (1/dinv), which includes the convection face terms.
__global__ void __launch_bounds__(32 * PFEM_ML_C * PFEM_ML_C)
k_ml_rowsums(const Grid g, const double* __restrict__ cl, const double* __restrict__ cv, const double* __restrict__ dinv,
             const int nlev, const int nJ1, double* __restrict__ S, const idx_t slen, const int koff) {
    __shared__ double sbuf[PFEM_ML_C * PFEM_ML_C][33];
    const int i = blockIdx.x * 32 + threadIdx.x;
    const int j = blockIdx.y * PFEM_ML_C + threadIdx.y, k = blockIdx.z * PFEM_ML_C + threadIdx.z - koff;
    const bool valid = i < g.nI && j < g.nJ && k >= g.kown0 && k < g.kown1;   // rows of this rank (all rows on one GPU)
    double a0[PFEM_ML_MAXL], a1[PFEM_ML_MAXL];
#pragma unroll
    for (int l = 0; l < PFEM_ML_MAXL; ++l) a0[l] = a1[l] = 0.;
    const idx_t n = valid ? i + g.sJ * j + g.sK * (idx_t)k : 0;
    const double dn = valid ? dinv[n] : 0.;
    if (valid && dn != 0.) {
        a0[0] = 1. / dn;
#pragma unroll
        for (int ek = 0; ek < 2; ++ek)
#pragma unroll
            for (int ej = 0; ej < 2; ++ej)
#pragma unroll
                for (int ei = 0; ei < 2; ++ei) {
                    // element whose lowest corner is (i-1+ei, j-1+ej, k-1+ek); this node is its local node a
                    const idx_t slot = n + (ei - 1) + g.sJ * (ej - 1) + g.sK * (ek - 1);
                    double kI, kJ, kK, kv[8];
                    elem_conductances(g, cl[slot], cv[slot], i - 1 + ei, j - 1 + ej, k - 1 + ek, kI, kJ, kK);
                    elem_matrix8(kI, kJ, kK, kv);
                    const int a = (1 - ei) | ((1 - ej) << 1) | ((1 - ek) << 2);
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        if (b == a) continue;
                        const int di = (b & 1) - (a & 1);
                        if (di < 0) continue;
                        const int jj = j - 1 + ej + ((b >> 1) & 1), kk = k - 1 + ek + ((b >> 2) & 1);
                        const idx_t nn = slot + (b & 1) + g.sJ * ((b >> 1) & 1) + g.sK * ((b >> 2) & 1);
                        if (dinv[nn] == 0.) continue;      // fixed / inactive / outside: no column in P
                        const double v = kv[a ^ b];
                        // first level on which (j, k) and (jj, kk) share an aggregate
                        int l = 1;
                        while (l < nlev && (((j ^ jj) | ((k + koff) ^ (kk + koff))) >> (PFEM_ML_SHIFT * l)) != 0) ++l;
#pragma unroll
                        for (int m = 0; m < PFEM_ML_MAXL; ++m)
                            if (m == l - 1) { if (di == 0) a0[m] += v; else a1[m] += v; }
                    }
                }
    }
    // prefix over the levels: a pair that shares an aggregate on level l shares one on every higher level
#pragma unroll
    for (int l = 1; l < PFEM_ML_MAXL; ++l) { a0[l] += a0[l - 1]; a1[l] += a1[l - 1]; }
    const int t = threadIdx.y + PFEM_ML_C * threadIdx.z;
    const idx_t dst = ((idx_t)blockIdx.z * nJ1 + blockIdx.y) * g.sJ + i;
#pragma unroll
    for (int l = 0; l < PFEM_ML_MAXL; ++l) {
        if (l >= nlev) break;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            __syncthreads();
            sbuf[t][threadIdx.x] = c ? a1[l] : a0[l];
            __syncthreads();
            if (t == 0 && i < g.sJ) {
                double s = 0.;
#pragma unroll
                for (int u = 0; u < PFEM_ML_C * PFEM_ML_C; ++u) s += sbuf[u][threadIdx.x];
                S[(idx_t)(2 * l + c) * slen + dst] = s;
            }
        }
    }
}

// level l >= 2: sum the level-1 partial sums of the f x f level-1 aggregates inside one level-l aggregate (fixed order)
__global__ void k_ml_gather(const LineDom d1, const LineDom dl, const int f, const double* __restrict__ S0, const double* __restrict__ S1,
                            double* __restrict__ diag, double* __restrict__ off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int J = blockIdx.y, K = blockIdx.z;
    if (i >= dl.sJ) return;
    double s0 = 0., s1 = 0.;
    const int j1 = min((J + 1) * f, d1.nJ), k1 = min((K + 1) * f, d1.nK);
    for (int k = K * f; k < k1; ++k)
        for (int j = J * f; j < j1; ++j) {
            const idx_t n = ((idx_t)k * d1.nJ + j) * d1.sJ + i;
            s0 += S0[n]; s1 += S1[n];
        }
    const idx_t n = ((idx_t)K * dl.nJ + J) * dl.sJ + i;
    diag[n] = s0; off[n] = s1;
}

// L D L^T of the tridiagonal blocks of one coarse level, in place: on entry ld = diagonal, ll = coupling of node i with i+1;
// on exit ld = 1/d (0 for rows without free nodes, whose z stays 0), ll[i] = multiplier of row i-1 in row i (like k_line_factor).
__global__ void k_ml_factor(const LineDom d, double* __restrict__ ll, double* __restrict__ ld, Scalars* sc) {
    const idx_t line = blockIdx.x * (idx_t)blockDim.x + threadIdx.x;
    if (line >= (idx_t)d.nJ * d.nK) return;
    const idx_t base = line * d.sJ;
    double dprev = 1., bprev = 0.;
    bool live_prev = false;
    for (int i = 0; i < d.sJ; ++i) {
        if (i >= d.nI) { ll[base + i] = 0.; ld[base + i] = 0.; continue; }
        const double diag = ld[base + i], b = ll[base + i];
        const bool live = diag > 0.;
        double l = 0., dd = diag;
        if (live && live_prev) { l = bprev / dprev; dd -= l * bprev; }
        if (live && !(dd > 0.)) { sc->neg_diag = 1; dd = diag; l = 0.; }   // cannot happen: T_l is a Galerkin block of an SPD matrix
        ll[base + i] = l;
        ld[base + i] = live ? 1. / dd : 0.;
        dprev = dd; bprev = b; live_prev = live;
    }
}

// ---- iteration ---------------------------------------------------------------------------------------------------

// Finalise rho / beta / stopping test once all levels have added their r_l.z_l.  mode 1: rho -> sc->bz; mode 2: nothing (test hook).
// ||z||^2 of the stopping test comes from the operator kernel of the PREVIOUS iteration (it is the only kernel that forms
// the complete z); a test that lags by one iteration can only delay the stop by one iteration.
__device__ __forceinline__ void ml_finalize(Scalars* sc, const double rho, const double rr, const int mode) {
    if (mode == 1) { sc->bz = rho; return; }
    if (mode == 2) return;
    const double rho_old = sc->rho;
    const int first = (sc->launch == 0);
    sc->rho_prev = rho_old; sc->rho = rho; sc->rr = rr;
    sc->beta = (!first && rho_old > 0.) ? rho / rho_old : 0.;
    const int launch = sc->launch + 1;
    sc->launch = launch;
    const int it = sc->bench ? launch : launch - 1;
    sc->iter = it;
    if (!sc->bench) {
        if (!(rr == rr) || !(rho == rho)) { sc->done = 2; sc->status = -2; }
        else if (rr <= sc->tol2 * sc->bb && rho <= sc->tol2 * sc->bz && sc->zz <= sc->tol2 * sc->xx) { sc->done = 2; sc->status = 1; }
        else if (it >= sc->maxit) { sc->done = 2; sc->status = 2; }
    }
}

// One tridiagonal solve T z = r by a warp, operands in the coalesced layout (lane holds the double2 at 2*lane + 64*c):
// transposition into the segment layout through the padded row `tb`, forward and backward sweep as warp scans over
// affine maps (same scheme as k_line_I).  Returns z in the coalesced layout and this lane's part of r.z (and of r.r).
template <int SEG>
__device__ __forceinline__ void warp_line_solve(double* tb, const int lane, const double2 (&vr)[SEG / 2], const double2 (&vl)[SEG / 2],
                                                const double2 (&vd)[SEG / 2], double2 (&vz)[SEG / 2], double& rz, double& rr) {
    const int so = lane * SEG + lane;
    double rp[SEG], l[SEG + 1], wv[SEG];
#pragma unroll
    for (int c = 0; c < SEG / 2; ++c) {
        const int i = 2 * lane + 64 * c;
        tb[i + i / SEG] = vr[c].x; tb[i + 1 + (i + 1) / SEG] = vr[c].y;
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < SEG; ++e) rp[e] = tb[so + e];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < SEG / 2; ++c) {
        const int i = 2 * lane + 64 * c;
        tb[i + i / SEG] = vl[c].x; tb[i + 1 + (i + 1) / SEG] = vl[c].y;
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < SEG; ++e) l[e] = tb[so + e];
    l[SEG] = (lane < 31) ? tb[so + SEG + 1] : 0.;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < SEG / 2; ++c) {
        const int i = 2 * lane + 64 * c;
        tb[i + i / SEG] = vd[c].x; tb[i + 1 + (i + 1) / SEG] = vd[c].y;
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < SEG; ++e) wv[e] = tb[so + e];
    __syncwarp();
    // forward sweep y_e = r_e - l_e y_{e-1}: affine map of the segment, inclusive warp scan, replay
    double A = 1., B = 0.;
#pragma unroll
    for (int e = 0; e < SEG; ++e) { B = fma(-l[e], B, rp[e]); A = -l[e] * A; }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double Ap = __shfl_up_sync(0xffffffffu, A, o), Bp = __shfl_up_sync(0xffffffffu, B, o);
        if (lane >= o) { B = fma(A, Bp, B); A *= Ap; }
    }
    double y = __shfl_up_sync(0xffffffffu, B, 1);
    if (lane == 0) y = 0.;
#pragma unroll
    for (int e = 0; e < SEG; ++e) { y = fma(-l[e], y, rp[e]); wv[e] *= y; }
    // backward sweep z_e = w_e - l_{e+1} z_{e+1}
    A = 1.; B = 0.;
#pragma unroll
    for (int e = SEG - 1; e >= 0; --e) { B = fma(-l[e + 1], B, wv[e]); A = -l[e + 1] * A; }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double An = __shfl_down_sync(0xffffffffu, A, o), Bn = __shfl_down_sync(0xffffffffu, B, o);
        if (lane + o < 32) { B = fma(A, Bn, B); A *= An; }
    }
    double z = __shfl_down_sync(0xffffffffu, B, 1);
    if (lane == 31) z = 0.;
#pragma unroll
    for (int e = SEG - 1; e >= 0; --e) {
        z = fma(-l[e + 1], z, wv[e]);
        tb[so + e] = z;
        rz = fma(rp[e], z, rz);
        rr = fma(rp[e], rp[e], rr);
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < SEG / 2; ++c) {
        const int i = 2 * lane + 64 * c;
        vz[c] = make_double2(tb[i + i / SEG], tb[i + 1 + (i + 1) / SEG]);
    }
    __syncwarp();
}

// The top level (ONE column = the 1-D vertical problem of the whole device) rides on the last aggregated level's kernel:
// every block also totals the residual rows it handles, the block that finishes last adds the totals in block order, solves
// the top line and finalises (z_top depends on i only: the operator kernel loads it once per thread).
struct MLTop {
    double* part;        // [gridDim.x][sJ] block totals; null: this launch is not the last aggregated level
    const double* ll;    // factors of the top line
    const double* ld;
    double* z;           // [sJ] z_top
};

// One level of the preconditioner.  One block (8 warps) per 4x4 aggregate of rows: warp w solves rows
// (j = 4J + (w & 3), k = 4K + 2 (w >> 2) + {0,1}) and accumulates their residuals; the block then adds the 8 partial rows
// in a fixed order into the residual row of the parent aggregate (rc_out, null on the last aggregated level).
// FINE: level 0 — r' = r - alpha q is formed and stored, the sums start the accumulators; otherwise r_in is the level's residual.
template <int SEG, bool FINE>
__global__ void __launch_bounds__(256)
k_line_ml(const LineDom d, const double* __restrict__ r_in, const double* __restrict__ q_in, const double* __restrict__ ll,
          const double* __restrict__ ld, double* __restrict__ r_out, double* __restrict__ z_out, double* __restrict__ rc_out,
          const int nJc, Scalars* sc, double* partials, const int mode, const MLTop top, const int pf) {
    constexpr int ROW = 32 * SEG, ROWP = ROW + ROW / SEG, NE = (ROW + 255) / 256;
    __shared__ double sh[32 * 2];
    __shared__ int sh_flag;
    __shared__ double tbuf[8][ROWP];   // per-warp transposition row; afterwards the warp's partial residual row
    if (mode == 0 && sc->done) return;
    const double alpha = (FINE && mode == 0) ? sc->alpha : 0.;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* const tb = tbuf[w];
    const int naJ = (d.nJ + PFEM_ML_C - 1) / PFEM_ML_C, naK = (d.nK + d.koff + PFEM_ML_C - 1) / PFEM_ML_C;
    double acc[2] = {0., 0.};
    double tot[NE];                    // block total of the residual rows (elements threadIdx.x + 256 e)
#pragma unroll
    for (int e = 0; e < NE; ++e) tot[e] = 0.;
    for (int agg = blockIdx.x; agg < naJ * naK; agg += gridDim.x) {
        const int J = agg % naJ, K = agg / naJ;
        double2 rs[SEG / 2];
#pragma unroll
        for (int c = 0; c < SEG / 2; ++c) rs[c] = make_double2(0., 0.);
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
            const int j = J * PFEM_ML_C + (w & 3), k = K * PFEM_ML_C + 2 * (w >> 2) + t - d.koff;
            if (j >= d.nJ || k < d.k0 || k >= d.k1) continue;     // warp-uniform
            const idx_t base = d.sJ * j + d.sK * (idx_t)k;
            double2 vr[SEG / 2], vl[SEG / 2], vd[SEG / 2], vz[SEG / 2];
#pragma unroll
            for (int c = 0; c < SEG / 2; ++c) {
                const int i = 2 * lane + 64 * c;
                vr[c] = vl[c] = vd[c] = make_double2(0., 0.);
                if (i < d.sJ) {
                    vr[c] = *reinterpret_cast<const double2*>(r_in + base + i);
                    vl[c] = *reinterpret_cast<const double2*>(ll + base + i);
                    vd[c] = *reinterpret_cast<const double2*>(ld + base + i);
                    if (FINE && mode == 0) {
                        const double2 qv = *reinterpret_cast<const double2*>(q_in + base + i);
                        vr[c].x = fma(-alpha, qv.x, vr[c].x); vr[c].y = fma(-alpha, qv.y, vr[c].y);
                        *reinterpret_cast<double2*>(r_out + base + i) = vr[c];
                    }
                    rs[c].x += vr[c].x; rs[c].y += vr[c].y;
                }
            }
            if (pf) {   // the row this warp solves next (t = 1 of this aggregate, or t = 0 of the block's next aggregate): into L2 meanwhile
                const int an = t ? agg + (int)gridDim.x : agg;
                const int jn = (an % naJ) * PFEM_ML_C + (w & 3), kn = (an / naJ) * PFEM_ML_C + 2 * (w >> 2) + (1 - t) - d.koff;
                if (an < naJ * naK && jn < d.nJ && kn >= d.k0 && kn < d.k1) {
                    const idx_t bn = d.sJ * jn + d.sK * (idx_t)kn;
                    prefetch_row_l2(r_in + bn, lane, (int)d.sJ);
                    prefetch_row_l2(ll + bn, lane, (int)d.sJ);
                    prefetch_row_l2(ld + bn, lane, (int)d.sJ);
                    if (FINE && mode == 0) prefetch_row_l2(q_in + bn, lane, (int)d.sJ);
                }
            }
            double rr = 0.;
            warp_line_solve<SEG>(tb, lane, vr, vl, vd, vz, acc[0], rr);
            if (FINE) acc[1] += rr;
#pragma unroll
            for (int c = 0; c < SEG / 2; ++c) {
                const int i = 2 * lane + 64 * c;
                if (i < d.sJ) *reinterpret_cast<double2*>(z_out + base + i) = vz[c];
            }
        }
        if (rc_out || top.part) {
#pragma unroll
            for (int c = 0; c < SEG / 2; ++c) { tb[2 * lane + 64 * c] = rs[c].x; tb[2 * lane + 64 * c + 1] = rs[c].y; }
            __syncthreads();
            const idx_t cb = ((idx_t)K * nJc + J) * d.sJ;
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                const int i = threadIdx.x + 256 * e;
                if (i < d.sJ) {
                    double s = tbuf[0][i];
#pragma unroll
                    for (int u = 1; u < 8; ++u) s += tbuf[u][i];
                    if (rc_out) rc_out[cb + i] = s;
                    tot[e] += s;
                }
            }
            __syncthreads();
        }
    }
    if (top.part) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const int i = threadIdx.x + 256 * e;
            if (i < d.sJ) top.part[(idx_t)blockIdx.x * d.sJ + i] = tot[e];
        }
    }
    if (grid_reduce<2, false>(acc, partials, &sc->ticket[3], sh, &sh_flag)) {
        double rho_top = 0.;
        bool summed = false;   // acc already holds the totals of all levels and ranks
        if (top.part) {
            // residual of the top level = sum of all rows of this level, block totals added in block order
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                const int i = threadIdx.x + 256 * e;
                if (i < ROW) {
                    double s = 0.;
                    if (i < d.sJ)
                        for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(top.part + (idx_t)b * d.sJ + i);
                    tbuf[1][i] = s;
                }
            }
            __syncthreads();
            // slab mode: the top level is the 1-D problem of the WHOLE device — its residual is the sum over the ranks, every rank
            // solves the same line; the scalars of the lower levels travel with it (acc: r.z of this rank's levels, |r'|^2)
            if (sc->comm) {
                if (threadIdx.x == 0) { acc[0] += FINE ? 0. : sc->ml_rho; acc[1] = FINE ? acc[1] : sc->ml_rr; }
                rank_allreduce_vec<2>(acc, tbuf[1], (int)d.sJ, sc->comm, sh);
                summed = true;
            }
            if (w == 0) {
                double2 vr[SEG / 2], vl[SEG / 2], vd[SEG / 2], vz[SEG / 2];
#pragma unroll
                for (int c = 0; c < SEG / 2; ++c) {
                    const int i = 2 * lane + 64 * c;
                    vr[c] = make_double2(tbuf[1][i], tbuf[1][i + 1]);
                    vl[c] = vd[c] = make_double2(0., 0.);
                    if (i < d.sJ) { vl[c] = *reinterpret_cast<const double2*>(top.ll + i); vd[c] = *reinterpret_cast<const double2*>(top.ld + i); }
                }
                double rr = 0.;
                warp_line_solve<SEG>(tbuf[0], lane, vr, vl, vd, vz, rho_top, rr);
                rho_top = warp_sum(rho_top);
#pragma unroll
                for (int c = 0; c < SEG / 2; ++c) {
                    const int i = 2 * lane + 64 * c;
                    if (i < d.sJ) *reinterpret_cast<double2*>(top.z + i) = vz[c];
                }
            }
        }
        if (threadIdx.x == 0) {
            const double rho = (FINE || summed ? acc[0] : sc->ml_rho + acc[0]) + rho_top;
            const double rr = FINE || summed ? acc[1] : sc->ml_rr;
            sc->ml_rho = rho; sc->ml_rr = rr;
            if (top.part) ml_finalize(sc, rho, rr, mode);
        }
    }
}

// What k_fpcg<MODE 3> adds to z_0 on the fly: z of the parent aggregates on the (up to) two aggregated levels,
// z_a[((k >> sh_a) * nJ_a + (j >> sh_a)) * sJ + i], plus z_top[i].  Unused slots point to a zero row with shift 31.
struct CoarseAdd {
    const double* z1;
    int nJ1, sh1;
    const double* z2;
    int nJ2, sh2;
    const double* zt;    // z_top [sJ]
    int koff;            // plane k lies in the level-1 aggregate row (k + koff) >> sh1 (slab mode: see LineDom)
};

// Slab mode: the operator kernel forms p' on its halo planes from z of the neighbour's boundary plane.  z_0 + z_1 + z_2 of the
// first / last owned plane goes into the neighbour's halo plane of z_0 (whose own coarse aggregates are empty); z_top is the same
// on all ranks and is added by the receiver like on every other plane.  Followed by k_rank_barrier.
__global__ void k_ml_halo(const Grid g, const double* __restrict__ z0, const CoarseAdd ca, double* z_lo, double* z_hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= g.nI) return;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        double* dst = side ? z_hi : z_lo;
        if (!dst) continue;
        const int k = side ? g.kown1 - 1 : g.kown0, kc = k + ca.koff;
        const idx_t n = i + g.sJ * j + g.sK * (idx_t)k;
        dst[g.sJ * j + i] = z0[n] + ca.z1[((idx_t)(kc >> ca.sh1) * ca.nJ1 + (j >> ca.sh1)) * g.sJ + i] +
                            ca.z2[((idx_t)(kc >> ca.sh2) * ca.nJ2 + (j >> ca.sh2)) * g.sJ + i];
    }
}

// Set-up in slab mode: diagonal and coupling of the top line are sums over the ranks (one block, before k_ml_factor of the top level)
__global__ void k_ml_top_allreduce(double* ld, double* ll, const int n, Scalars* sc) {
    __shared__ double sh[PFEM_COMM_NV];
    __shared__ double buf[PFEM_COMM_VEC];
    for (int i = threadIdx.x; i < n; i += blockDim.x) { buf[i] = ld[i]; buf[n + i] = ll[i]; }
    double v[1] = {0.};
    rank_allreduce_vec<1>(v, buf, 2 * n, sc->comm, sh);
    for (int i = threadIdx.x; i < n; i += blockDim.x) { ld[i] = buf[i]; ll[i] = buf[n + i]; }
}

__global__ void k_pupdate_plain(idx_t N, const double* __restrict__ r, const double* __restrict__ dinv, double* __restrict__ z) {
    for (idx_t n = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; n < N; n += (idx_t)gridDim.x * blockDim.x) z[n] = dinv[n] * r[n];
}
// test hook (pfem_apply_precond): the complete z = z_0 + coarse correction on the fine lattice
__global__ void k_ml_prolong_add(const Grid g, const double* __restrict__ z0, const CoarseAdd ca, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, k = blockIdx.z;
    if (i >= g.nI) return;
    const idx_t n = i + g.sJ * j + g.sK * (idx_t)k;
    const int kc = k + ca.koff;
    out[n] = z0[n] + ca.z1[((idx_t)(kc >> ca.sh1) * ca.nJ1 + (j >> ca.sh1)) * g.sJ + i] + ca.z2[((idx_t)(kc >> ca.sh2) * ca.nJ2 + (j >> ca.sh2)) * g.sJ + i] + ca.zt[i];
}

}  // namespace pfem
