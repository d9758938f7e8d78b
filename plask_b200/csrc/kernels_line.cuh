// kernels_line.cuh — line-Jacobi preconditioner along the vertical axis (pfem_opts::precond = 1; NSPCG's "ljac",
// extlib/nspcg/nspcg.f:1593 ff., selectable in the reference through iter_params.preconditioner,
// plask/common/fem/iterative_matrix.hpp:27-46).
//
// M = the tridiagonal blocks of A along the mesh lines of the physical vertical axis (a principal sub-matrix of A for
// every line, hence SPD); z = M^-1 r is one L D L^T solve per line.  VCSEL-like meshes have layers of 5-80 nm against
// lateral steps of 0.25-4 um, so the vertical couplings dominate A by (h_lat/h_vert)^2 = 10..10^5 and point Jacobi leaves
// a condition number that grows with that ratio; the line solve removes it.
//
// PCG iteration m with this preconditioner is two kernels (two reductions, classic PCG — the beta prediction of k_fpcg
// needs M^-1 q, which is another line solve):
//   k_line_I / k_line_strided   r' = r - alpha q ;  z = M^-1 r' ;  rho = r'.z, |r'|^2, |z|^2 -> beta, stopping test
//   k_fpcg (Scalars::line = 1)  x' = x + alpha p ;  p' = z + beta p ;  q' = M_D A p' ;  p'.q' -> alpha
// Traffic: (r, q, l, d^-1 | r', z) + (z, p, mask, x, c_lat, c_vert | p', q', x') = 15 words = 120 B / DOF / iteration.
//
// Lines along I (vertical axis = minor axis, the default order 012 of explicit meshes) are rows of the lattice: one WARP
// per row, every lane owns SEG consecutive nodes, and the two bidiagonal recurrences are solved with a warp scan over
// affine maps (y_out = A y_in + B), so forward and backward sweep happen in registers in one pass over the row.
// Lines along J or K: one thread per line, coalesced across I, forward sweep parked in the z array.
#pragma once
#include "pfem_internal.cuh"

namespace pfem {

// L D L^T of the line blocks.  One thread per line; `sL` = lattice stride along the line, `nL` its length, lines are
// enumerated by (a, b) with strides sA, sB and counts nA, nB.  diag = 1/dinv_jacobi (boundary-face terms included),
// off-diagonal = sum over the 4 elements around the edge of (kI+kJ+kK - 3 k_line)/18 (therm3d.cpp:229-231).
// Fixed / empty rows (dinv == 0) become identity rows.  Outputs: ll[n] = l (coupling to the previous node of the line,
// 0 for the first), ld[n] = 1/d, lmask[n] = 1 for free rows, 0 otherwise.
__global__ void k_line_factor(const Grid g, const double* __restrict__ cl, const double* __restrict__ cv,
                              const double* __restrict__ dinv, double* __restrict__ ll, double* __restrict__ ld,
                              double* __restrict__ lmask, Scalars* sc) {
    const int vd = g.vdim;
    const int nL = vd == 0 ? g.nI : vd == 1 ? g.nJ : g.nK;
    const int nA = vd == 0 ? g.nJ : g.nI;                 // fastest remaining axis first
    const int nB = vd == 2 ? g.nJ : g.nK;
    const idx_t sL = vd == 0 ? 1 : vd == 1 ? g.sJ : g.sK;
    const idx_t sA = vd == 0 ? g.sJ : 1;
    const idx_t sB = vd == 2 ? g.sJ : g.sK;
    const idx_t line = blockIdx.x * (idx_t)blockDim.x + threadIdx.x;
    if (line >= (idx_t)nA * nB) return;
    const int a = (int)(line % nA), b = (int)(line / nA);
    // lattice coordinates of the line start and the two transverse coordinates
    int c[3];
    c[vd] = 0;
    if (vd == 0) { c[1] = a; c[2] = b; } else if (vd == 1) { c[0] = a; c[2] = b; } else { c[0] = a; c[1] = b; }
    const idx_t n0 = sA * a + sB * b;
    // transverse axes t1, t2 (index-space axes other than vd) and their strides
    const int t1 = vd == 0 ? 1 : 0, t2 = vd == 2 ? 1 : 2;
    const idx_t st[3] = {1, g.sJ, g.sK};
    double dprev = 1., bprev = 0.;   // d_{m-1}, b_{m-1} (coupling m-1 <-> m)
    bool fprev = true;
    for (int m = 0; m < nL; ++m) {
        const idx_t n = n0 + sL * m;
        const double dj = dinv[n];
        const bool fixed = (dj == 0.);
        // coupling between node m and m+1 along the line: the 4 elements around that edge
        double bnext = 0.;
        if (m + 1 < nL) {
            c[vd] = m;
#pragma unroll
            for (int o2 = -1; o2 <= 0; ++o2)
#pragma unroll
                for (int o1 = -1; o1 <= 0; ++o1) {
                    int e[3] = {c[0], c[1], c[2]};
                    e[t1] += o1; e[t2] += o2;
                    const idx_t slot = n + st[t1] * o1 + st[t2] * o2;
                    double kI, kJ, kK;
                    elem_conductances(g, cl[slot], cv[slot], e[0], e[1], e[2], kI, kJ, kK);
                    const double kl = vd == 0 ? kI : vd == 1 ? kJ : kK;
                    bnext += ((kI + kJ + kK) - 3. * kl) * (1. / 18.);
                }
        }
        double l = 0., d;
        if (fixed) d = 1.;
        else {
            d = 1. / dj;
            if (!fprev && m > 0) { l = bprev / dprev; d -= l * bprev; }
        }
        if (!(d > 0.)) { sc->neg_diag = 1; d = 1.; l = 0.; }   // cannot happen for an SPD line block
        ll[n] = l;
        ld[n] = 1. / d;
        lmask[n] = fixed ? 0. : 1.;
        dprev = d; bprev = fixed ? 0. : bnext; fprev = fixed;
    }
}

// What the last block of a line kernel does with the sums.  mode 0: PCG iteration; mode 1: rho -> sc->bz (norm of the lifted
// rhs in the metric of this preconditioner).
__device__ __forceinline__ void line_finalize(Scalars* sc, const double rho, const double rr, const double zz, const int mode) {
    if (mode == 1) { sc->bz = rho; return; }
    const double rho_old = sc->rho;
    const int first = (sc->launch == 0);
    sc->rho_prev = rho_old; sc->rho = rho; sc->rr = rr; sc->zz = zz;
    sc->beta = (!first && rho_old > 0.) ? rho / rho_old : 0.;
    const int launch = sc->launch + 1;
    sc->launch = launch;
    const int it = sc->bench ? launch : launch - 1;   // like k_fpcg: launch m has applied m-1 updates to r
    sc->iter = it;
    if (!sc->bench) {
        // done = 2: the operator kernel of this iteration still has to apply the pending x update, then it sets done = 1
        if (!(rr == rr) || !(rho == rho)) { sc->done = 2; sc->status = -2; }
        else if (rr <= sc->tol2 * sc->bb && rho <= sc->tol2 * sc->bz && zz <= sc->tol2 * sc->xx) { sc->done = 2; sc->status = 1; }
        else if (it >= sc->maxit) { sc->done = 2; sc->status = 2; }
    }
}

// ---- lines along I: one warp per lattice row ---------------------------------------------------------------
// Global accesses are coalesced (lane reads the double2 at 2*lane + 64*c); a padded per-warp shared-memory row turns
// that layout into the segment layout of the scan (lane owns nodes lane*SEG .. lane*SEG+SEG-1) and back.
template <int SEG>
__global__ void __launch_bounds__(256)
k_line_I(const Grid g, const double* __restrict__ r_in, const double* __restrict__ q_in, const double* __restrict__ ll,
         const double* __restrict__ ld, double* __restrict__ r_out, double* __restrict__ z_out, Scalars* sc, double* partials,
         const int mode, const PeerOut po, const int pf) {
    constexpr int ROW = 32 * SEG, ROWP = ROW + ROW / SEG;   // one pad per segment: conflict-free segment reads
    __shared__ double sh[32 * 3];
    __shared__ int sh_flag;
    __shared__ double tbuf[8][ROWP];
    if (mode == 0 && sc->done) return;
    const double alpha = mode == 0 ? sc->alpha : 0.;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* const tb = tbuf[wib];
    const idx_t warp = (blockIdx.x * (idx_t)blockDim.x + threadIdx.x) >> 5;
    const idx_t nwarps = ((idx_t)gridDim.x * blockDim.x) >> 5;
    const idx_t rows = (idx_t)g.nJ * (g.kown1 - g.kown0);
    const int i0 = lane * SEG;
    const int so = i0 + lane;   // padded offset of the own segment
    double acc[3] = {0., 0., 0.};
    for (idx_t row = warp; row < rows; row += nwarps) {
        const int j = (int)(row % g.nJ), k = g.kown0 + (int)(row / g.nJ);
        const idx_t base = g.sJ * j + g.sK * (idx_t)k;
        double rp[SEG], l[SEG + 1], w[SEG];
        // all global loads of the row first (coalesced, in flight together); r' = r - alpha q is stored right away
        double2 vr[SEG / 2], vl[SEG / 2], vd[SEG / 2];
#pragma unroll
        for (int c = 0; c < SEG / 2; ++c) {
            const int i = 2 * lane + 64 * c;
            vr[c] = vl[c] = vd[c] = make_double2(0., 0.);
            if (i < g.sJ) {   // pads inside the pitch hold zeros
                vr[c] = *reinterpret_cast<const double2*>(r_in + base + i);
                vl[c] = *reinterpret_cast<const double2*>(ll + base + i);
                vd[c] = *reinterpret_cast<const double2*>(ld + base + i);
                if (mode == 0) {
                    const double2 qv = *reinterpret_cast<const double2*>(q_in + base + i);
                    vr[c].x = fma(-alpha, qv.x, vr[c].x); vr[c].y = fma(-alpha, qv.y, vr[c].y);
                    *reinterpret_cast<double2*>(r_out + base + i) = vr[c];
                }
            }
        }
        if (pf && row + pf * nwarps < rows) {   // the row this warp solves next (pf = 1; 2 = the one after): into L2 while this one is solved
            const idx_t rn = row + pf * nwarps;
            const idx_t bn = g.sJ * (idx_t)(rn % g.nJ) + g.sK * (idx_t)(g.kown0 + rn / g.nJ);
            prefetch_row_l2(r_in + bn, lane, (int)g.sJ);
            prefetch_row_l2(ll + bn, lane, (int)g.sJ);
            prefetch_row_l2(ld + bn, lane, (int)g.sJ);
            if (mode == 0) prefetch_row_l2(q_in + bn, lane, (int)g.sJ);
        }
        // transposition into the segment layout, one array at a time through the padded row buffer
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
        l[SEG] = (lane < 31) ? tb[so + SEG + 1] : 0.;   // first l of the next segment
        __syncwarp();
#pragma unroll
        for (int c = 0; c < SEG / 2; ++c) {
            const int i = 2 * lane + 64 * c;
            tb[i + i / SEG] = vd[c].x; tb[i + 1 + (i + 1) / SEG] = vd[c].y;
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < SEG; ++e) w[e] = tb[so + e];
        __syncwarp();
        // forward sweep y_e = r'_e - l_e y_{e-1}: affine map of the segment, inclusive warp scan, replay
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
        for (int e = 0; e < SEG; ++e) { y = fma(-l[e], y, rp[e]); w[e] *= y; }   // w = D^-1 y
        // backward sweep z_e = w_e - l_{e+1} z_{e+1}
        A = 1.; B = 0.;
#pragma unroll
        for (int e = SEG - 1; e >= 0; --e) { B = fma(-l[e + 1], B, w[e]); A = -l[e + 1] * A; }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double An = __shfl_down_sync(0xffffffffu, A, o), Bn = __shfl_down_sync(0xffffffffu, B, o);
            if (lane + o < 32) { B = fma(A, Bn, B); A *= An; }
        }
        double z = __shfl_down_sync(0xffffffffu, B, 1);
        if (lane == 31) z = 0.;
#pragma unroll
        for (int e = SEG - 1; e >= 0; --e) {
            z = fma(-l[e + 1], z, w[e]);
            tb[so + e] = z;
            acc[0] = fma(rp[e], z, acc[0]);
            acc[1] = fma(rp[e], rp[e], acc[1]);
            acc[2] = fma(z, z, acc[2]);
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < SEG / 2; ++c) {
            const int i = 2 * lane + 64 * c;
            if (i < g.sJ) {
                const double2 zv = make_double2(tb[i + i / SEG], tb[i + 1 + (i + 1) / SEG]);
                *reinterpret_cast<double2*>(z_out + base + i) = zv;
                // slab mode: z of the first / last owned plane also goes into the neighbour's halo plane (NVLink peer store)
                if (po.z_lo && k == g.kown0) *reinterpret_cast<double2*>(po.z_lo + g.sJ * j + i) = zv;
                if (po.z_hi && k == g.kown1 - 1) *reinterpret_cast<double2*>(po.z_hi + g.sJ * j + i) = zv;
            }
        }
        __syncwarp();
    }
    if (grid_reduce<3, false>(acc, partials, &sc->ticket[3], sh, &sh_flag, po.z_lo != nullptr || po.z_hi != nullptr)) {
        if (sc->comm) rank_allreduce<3, false>(acc, sc->comm, sh);   // also orders the z halo stores before the operator step
        if (threadIdx.x == 0) line_finalize(sc, acc[0], acc[1], acc[2], mode);
    }
}

// ---- lines along J or K: one thread per line, coalesced across I ---------------------------------------------
// Parallelism is only nI x (other axis) threads, so every thread keeps U planes of loads in flight: batches of U nodes are
// loaded, swept and stored.  rho = y.D^-1 y comes out of the forward sweep (r'.M^-1 r' = y^T D^-1 y).
__global__ void __launch_bounds__(128)
k_line_strided(const Grid g, const double* r_in, const double* __restrict__ q_in, const double* __restrict__ ll,
               const double* __restrict__ ld, double* r_out, double* z_out, Scalars* sc, double* partials, const int mode,
               const PeerOut po) {
    constexpr int U = 8;
    __shared__ double sh[32 * 3];
    __shared__ int sh_flag;
    if (mode == 0 && sc->done) return;
    const double alpha = mode == 0 ? sc->alpha : 0.;
    // vd = 1: lines along J, one per (i, owned plane k);  vd = 2: lines along K, one per (i, j), never in slab mode;
    // vd = 0: lines along I, one per (j, owned plane k) — the uncoalesced fallback for lines longer than k_line_I holds
    const int vd = g.vdim;
    const int nL = vd == 0 ? g.nI : vd == 1 ? g.nJ : g.nK;
    const int nA = vd == 0 ? g.nJ : g.nI;
    const int b0 = vd == 2 ? 0 : g.kown0;
    const int nB = vd == 2 ? g.nJ : g.kown1 - g.kown0;
    const idx_t sL = vd == 0 ? 1 : vd == 1 ? g.sJ : g.sK, sA = vd == 0 ? g.sJ : 1, sB = vd == 2 ? g.sJ : g.sK;
    const idx_t lines = (idx_t)nA * nB;
    double acc[3] = {0., 0., 0.};
    for (idx_t line = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; line < lines; line += (idx_t)gridDim.x * blockDim.x) {
        const int bi = b0 + (int)(line / nA);
        const idx_t n0 = sA * (line % nA) + sB * bi;
        double* const zp = (vd != 2 && po.z_lo && bi == g.kown0) ? po.z_lo - sB * bi
                         : (vd != 2 && po.z_hi && bi == g.kown1 - 1) ? po.z_hi - sB * bi : nullptr;   // halo copy of this plane
        double* const zp2 = (vd != 2 && po.z_lo && po.z_hi && bi == g.kown0 && bi == g.kown1 - 1) ? po.z_hi - sB * bi : nullptr;
        double y = 0.;
        for (int m0 = 0; m0 < nL; m0 += U) {
            double rp[U], lv[U], dv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int m = m0 + u;
                if (m < nL) {
                    const idx_t n = n0 + sL * m;
                    rp[u] = r_in[n]; lv[u] = ll[n]; dv[u] = ld[n];
                    if (mode == 0) rp[u] = fma(-alpha, q_in[n], rp[u]);
                } else { rp[u] = 0.; lv[u] = 0.; dv[u] = 0.; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                y = fma(-lv[u], y, rp[u]);
                acc[0] = fma(y * y, dv[u], acc[0]);
                acc[1] = fma(rp[u], rp[u], acc[1]);
                lv[u] = y;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int m = m0 + u;
                if (m < nL) {
                    const idx_t n = n0 + sL * m;
                    if (mode == 0) r_out[n] = rp[u];
                    z_out[n] = lv[u];
                }
            }
        }
        double z = 0., lnext = 0.;
        for (int m1 = nL; m1 > 0; m1 -= U) {
            double yv[U], lv[U], dv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int m = m1 - 1 - u;
                if (m >= 0) {
                    const idx_t n = n0 + sL * m;
                    yv[u] = z_out[n]; lv[u] = ll[n]; dv[u] = ld[n];
                } else { yv[u] = 0.; lv[u] = 0.; dv[u] = 0.; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                z = fma(-lnext, z, yv[u] * dv[u]);
                lnext = lv[u];
                yv[u] = z;
                acc[2] = fma(z, z, acc[2]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int m = m1 - 1 - u;
                if (m >= 0) {
                    z_out[n0 + sL * m] = yv[u];
                    if (zp) zp[n0 + sL * m] = yv[u];
                    if (zp2) zp2[n0 + sL * m] = yv[u];
                }
            }
        }
    }
    if (grid_reduce<3, false>(acc, partials, &sc->ticket[3], sh, &sh_flag, po.z_lo != nullptr || po.z_hi != nullptr)) {
        if (sc->comm) rank_allreduce<3, false>(acc, sc->comm, sh);   // also orders the z halo stores before the operator step
        if (threadIdx.x == 0) line_finalize(sc, acc[0], acc[1], acc[2], mode);
    }
}

}  // namespace pfem
