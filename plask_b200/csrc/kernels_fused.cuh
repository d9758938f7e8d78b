// kernels_fused.cuh — the production kernel: ONE kernel per PCG iteration.
//
// Classic Jacobi-PCG (extlib/nspcg/nspcg.f:9281-9337) needs two grid-wide reductions per iteration
// (p.Ap for alpha, r.z for beta), hence two kernels and 14 FP64 words of HBM traffic per DOF.
// Here iteration m is one launch that, per node,
//     r' = r - alpha q          x' = x + alpha p         (update with the alpha of launch m-1)
//     z  = D^-1 r'              p' = z + beta~ p          (new search direction)
//     q' = M A p'                                         (matrix-free 27-point brick operator)
// and reduces  p'.q', r'.z, q'.z, q'.D^-1 q', r'.r', z.z, x'.x'  in one deterministic grid reduction.
// The last CTA then has the EXACT rho = r'.z, alpha' = rho / p'.q', the stopping test, and the
// PREDICTED next rho~ = rho - 2 alpha' q'.z + alpha'^2 q'.D^-1 q'  (= r''.D^-1 r'' expanded), which
// gives beta~' = rho~/rho for the next launch without another pass over the vectors.  Only beta uses
// the predicted value; the next launch recomputes rho exactly, so alpha is always the exact line
// search along p' and the energy norm decreases monotonically whatever the rounding in beta~.
// Traffic: reads r, q, p, D^-1, x, c_lat, c_vert; writes r', p', q', x'  = 11 words = 88 B / DOF.
// r, q, p are double buffered (halo nodes of the old vectors are read by neighbouring CTAs).
//
// Data movement: per step of the march along the major axis one elected thread issues TMA box loads
// (cp.async.bulk.tensor.3d, tile + halo, OOB zero fill) of the next node plane of r, q, p, D^-1 and
// of the next element layer of c_lat, c_vert into an NS-deep shared-memory ring, completion on
// mbarriers.  Phase 1 turns a landed stage into the p' plane (tile + 1-node halo) and the per-element
// stencil coefficients {kI+kJ, kJ-2kI, kI-2kJ, kK}; phase 2 gathers the operator with the plane below
// kept in registers.  One __syncthreads per step.
//
// Stencil algebra (K_e = kI S(x)M(x)M + kJ M(x)S(x)M + kK M(x)M(x)S, therm3d.cpp:226-237, all k /36):
// the layer between node planes a (below) and b (above) contributes  2La + Lb - C  to plane a and
// La + 2Lb + C to plane b, where for a plane X with window X[row][col] around the node (centre 11)
//     LX = 2S X11 + A0 X10 + A1 X12 + B0 X01 + B1 X21 - sum_e (kI+kJ)_e X_corner(e)
//     C  = 4T d11 + 2 (P0 d10 + P1 d12 + Q0 d01 + Q1 d21) + sum_e kK_e d_corner(e),   d = b - a
// with S, T the sums of (kI+kJ), kK over the 4 elements of the layer around the node, A/P the sums
// over the two elements left/right (of kJ-2kI, kK), B/Q over the two elements below/above (kI-2kJ, kK).
#pragma once
#include <cuda.h>

#include "kernels_tma.cuh"

namespace pfem {

template <int TJ>
struct FusedTile {
    static constexpr int TI = 32, HX = 2;
    static constexpr int PW = TI + 2 * HX, PH = TJ + 2;   // raw TMA box (doubles)
    static constexpr int BOX = PW * PH;
    static constexpr int BOXP = (BOX + 15) / 16 * 16;
    static constexpr int PWP = TI + 2;                   // p' plane pitch, rows PH
    static constexpr int PLANE = (PWP * PH + 15) / 16 * 16;
    static constexpr int CW = TI + 1, CH = TJ + 1;       // coefficient layer, 4 doubles per element
    static constexpr int LAYER = (CW * CH * 4 + 15) / 16 * 16;
    static constexpr int NRED = 7;
    static constexpr size_t smem_bytes(int ns, bool fused) {
        return 128 + sizeof(double) * ((size_t)ns * (fused ? 6 : 4) * BOXP + 2 * PLANE + 2 * LAYER + 32 * NRED) + 16 * 8 + 16;
    }
};

struct __align__(16) Coef4 { double sij, ui, uj, kk; };

template <int TJ, int RJ, int NS, int MINB, int VDIM, bool FUSED>
__global__ void __launch_bounds__(32 * (TJ / RJ), MINB)
k_fpcg(const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_q,
       const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_d,
       const __grid_constant__ CUtensorMap tm_cl, const __grid_constant__ CUtensorMap tm_cv, const Grid g, const int lk,
       double* __restrict__ r_out, double* __restrict__ q_out, double* __restrict__ p_out, double* __restrict__ x,
       Scalars* sc, double* partials) {
    typedef FusedTile<TJ> T;
    constexpr int TI = T::TI, HX = T::HX, PW = T::PW, PH = T::PH, BOX = T::BOX, BOXP = T::BOXP, PWP = T::PWP;
    constexpr int PLANE = T::PLANE, CW = T::CW, CH = T::CH, LAYER = T::LAYER, NRED = T::NRED;
    constexpr int NT = TI * (TJ / RJ);
    constexpr int NBN = FUSED ? 4 : 2;   // node boxes per stage: r q p d | p d
    constexpr int NB = NBN + 2;          // + c_lat, c_vert
    constexpr int B_P = FUSED ? 2 : 0, B_D = FUSED ? 3 : 1, B_CL = NBN, B_CV = NBN + 1;
    constexpr int NRING = 2 * PWP + 2 * TJ;     // halo ring of the p' plane
    constexpr int NERING = TI + TJ + 1;         // halo row/column of the coefficient layer

    extern __shared__ __align__(128) unsigned char smem_dyn[];
    double* const sRaw = reinterpret_cast<double*>(smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u));  // [NS][NB][BOXP]
    double* const sP = sRaw + (size_t)NS * NB * BOXP;   // [2][PLANE]
    double* const sC = sP + 2 * PLANE;                  // [2][LAYER]
    double* const sRed = sC + 2 * LAYER;                // [32*NRED]
    uint64_t* const bars = reinterpret_cast<uint64_t*>(sRed + 32 * NRED);  // [NS]
    int* const sh_flag = reinterpret_cast<int*>(bars + 16);

    if (FUSED && sc->done) return;
    const double alpha = FUSED ? sc->alpha : 0.;
    const double beta = FUSED ? sc->beta : 0.;

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = tx + TI * ty;
    const int i0 = blockIdx.x * TI, j0 = blockIdx.y * TJ;
    const int k0 = blockIdx.z * lk;
    const int k1 = min(k0 + lk, g.nK);
    const int nsteps = k1 - k0 + 1;   // items t = 0 .. nsteps: node planes k0-1 .. k1, element layers k0-1 .. k1-1
    const int jl0 = ty * RJ;

    // ---- per-thread constants -----------------------------------------------------------
    const int i = i0 + tx;
    const bool vi = i < g.nI;
    constexpr double s36 = 1e-6 / 36.;
    // element weights: kI = cI wI hk, kJ = cJ wJ hk, kK = cK wK rk   (therm3d.cpp:215-220)
    double wI[RJ], wJ[RJ], wK[RJ];
    bool vj[RJ];
    {
        const int ei = min(i, g.nI - 1);
        const double hi = g.hI[ei], ri = g.rI[ei];
#pragma unroll
        for (int rr = 0; rr < RJ; ++rr) {
            const int j = j0 + jl0 + rr;
            vj[rr] = vi && j < g.nJ;
            const int ej = min(j, g.nJ - 1);
            const double hj = g.hJ[ej];
            wI[rr] = s36 * hj * ri;
            wJ[rr] = s36 * hi * g.rJ[ej];
            wK[rr] = s36 * hi * hj;
        }
    }
    // halo ring of the node plane handled by this thread (entry h = tid), offsets into raw box / p' plane
    int ring_raw = -1, ring_pl = 0;
    if (tid < NRING) {
        int jj, ii;
        if (tid < PWP) { jj = 0; ii = tid; }
        else if (tid < 2 * PWP) { jj = PH - 1; ii = tid - PWP; }
        else { const int h = tid - 2 * PWP; jj = 1 + (h >> 1); ii = (h & 1) ? PWP - 1 : 0; }
        ring_raw = jj * PW + ii + (HX - 1);
        ring_pl = jj * PWP + ii;
    }
    // halo row / column of the element layer handled by this thread (entry e = NT-1-tid)
    int er_raw = -1, er_c = 0;
    double wIr = 0., wJr = 0., wKr = 0.;
    if (NT - 1 - tid < NERING) {
        const int e = NT - 1 - tid;
        int ejl, eil;                       // layer coordinates: element (j0-1+ejl, i0-1+eil)
        if (e <= TJ) { ejl = e; eil = 0; } else { ejl = 0; eil = e - TJ; }
        er_raw = ejl * PW + eil + (HX - 1);
        er_c = ejl * CW + eil;
        const int ei = min(max(i0 - 1 + eil, -1), g.nI - 1), ej = min(max(j0 - 1 + ejl, -1), g.nJ - 1);
        const double hi = g.hI[ei], hj = g.hJ[ej];
        wIr = s36 * hj * g.rI[ei];
        wJr = s36 * hi * g.rJ[ej];
        wKr = s36 * hi * hj;
    }

    auto issue = [&](int t) {
        const int st = t % NS;
        double* dst = sRaw + (size_t)st * NB * BOXP;
        uint64_t* bar = &bars[st];
        mbar_expect_tx(bar, (uint32_t)((NBN + (t > 0 ? 2 : 0)) * BOX * sizeof(double)));
        const int P = k0 - 1 + t;
        if (FUSED) {
            tma_load_3d(dst, &tm_r, bar, i0 - HX, j0 - 1, P);
            tma_load_3d(dst + BOXP, &tm_q, bar, i0 - HX, j0 - 1, P);
        }
        tma_load_3d(dst + B_P * BOXP, &tm_p, bar, i0 - HX, j0 - 1, P);
        tma_load_3d(dst + B_D * BOXP, &tm_d, bar, i0 - HX, j0 - 1, P);
        if (t > 0) {
            tma_load_3d(dst + B_CL * BOXP, &tm_cl, bar, i0 - HX, j0 - 1, P - 1);
            tma_load_3d(dst + B_CV * BOXP, &tm_cv, bar, i0 - HX, j0 - 1, P - 1);
        }
    };

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int t = 0; t < NS && t <= nsteps; ++t) issue(t);

    // ---- state carried along the march ----------------------------------------------------
    double wa[RJ + 2][3];          // p' window of the plane below (registers)
    double carry[RJ];              // contribution of the layer below to the current plane
    double za[RJ], da[RJ];         // z and D^-1 of the own nodes of the plane below
#pragma unroll
    for (int y = 0; y < RJ + 2; ++y)
#pragma unroll
        for (int c = 0; c < 3; ++c) wa[y][c] = 0.;
#pragma unroll
    for (int rr = 0; rr < RJ; ++rr) { carry[rr] = 0.; za[rr] = 0.; da[rr] = 0.; }
    double red[NRED];              // pq, rz, qz, qdq, rr, zz, xx
#pragma unroll
    for (int a = 0; a < NRED; ++a) red[a] = 0.;

    // x of the own nodes, fetched one step ahead
    double xn[RJ];
#pragma unroll
    for (int rr = 0; rr < RJ; ++rr) xn[rr] = 0.;
    auto fetch_x = [&](int P) {
        if (!FUSED) return;
        if (P >= k0 && P < k1) {
#pragma unroll
            for (int rr = 0; rr < RJ; ++rr)
                if (vj[rr]) xn[rr] = x[i + g.sJ * (j0 + jl0 + rr) + g.sK * P];
        }
    };
    fetch_x(k0 - 1);

    for (int t = 0; t <= nsteps; ++t) {
        const int st = t % NS;
        const int P = k0 - 1 + t;                 // node plane of this step
        const bool own_plane = (P >= k0 && P < k1);
        const double* raw = sRaw + (size_t)st * NB * BOXP;
        double* sPb = sP + (t & 1) * PLANE;
        double* sCb = sC + (t & 1) * LAYER;
        double xcur[RJ];
#pragma unroll
        for (int rr = 0; rr < RJ; ++rr) xcur[rr] = xn[rr];
        fetch_x(P + 1);
        const int L = P - 1;                      // element layer of this step (t >= 1)
        const int Lc = min(max(L, -1), g.nK - 1);
        const double hk = g.hK[Lc], rk = g.rK[Lc];
        mbar_wait(&bars[st], (uint32_t)((t / NS) & 1));

        // ---------------- phase 1: p' plane and coefficient layer -------------------------
        double zb[RJ], db[RJ];
#pragma unroll
        for (int rr = 0; rr < RJ; ++rr) {
            const int ro = (jl0 + rr + 1) * PW + tx + HX;
            double pn;
            if (FUSED) {
                const double r0 = raw[ro], q0 = raw[BOXP + ro], p0 = raw[B_P * BOXP + ro], dd = raw[B_D * BOXP + ro];
                const double rn = fma(-alpha, q0, r0);
                const double z = dd * rn;
                pn = fma(beta, p0, z);
                zb[rr] = z; db[rr] = dd;
                if (own_plane && vj[rr]) {
                    const idx_t n = i + g.sJ * (j0 + jl0 + rr) + g.sK * P;
                    const double xv = fma(alpha, p0, xcur[rr]);
                    r_out[n] = rn;
                    p_out[n] = pn;
                    x[n] = xv;
                    red[1] = fma(rn, z, red[1]);
                    red[4] = fma(rn, rn, red[4]);
                    red[5] = fma(z, z, red[5]);
                    red[6] = fma(xv, xv, red[6]);
                }
            } else {
                pn = raw[B_P * BOXP + ro];
                zb[rr] = 0.; db[rr] = raw[B_D * BOXP + ro];
            }
            sPb[(jl0 + rr + 1) * PWP + tx + 1] = pn;
            if (t > 0) {
                const double a = raw[B_CL * BOXP + ro], b = raw[B_CV * BOXP + ro];
                const double kI = ((VDIM == 0 ? b : a) * wI[rr]) * hk;
                const double kJ = ((VDIM == 1 ? b : a) * wJ[rr]) * hk;
                const double kK = ((VDIM == 2 ? b : a) * wK[rr]) * rk;
                Coef4 c;
                c.sij = kI + kJ; c.ui = fma(-2., kI, kJ); c.uj = fma(-2., kJ, kI); c.kk = kK;
                reinterpret_cast<Coef4*>(sCb)[(jl0 + rr + 1) * CW + tx + 1] = c;
            }
        }
        if (ring_raw >= 0) {
            double pn;
            if (FUSED) {
                const double r0 = raw[ring_raw], q0 = raw[BOXP + ring_raw], p0 = raw[B_P * BOXP + ring_raw],
                             dd = raw[B_D * BOXP + ring_raw];
                pn = fma(beta, p0, dd * fma(-alpha, q0, r0));
            } else {
                pn = raw[B_P * BOXP + ring_raw];
            }
            sPb[ring_pl] = pn;
        }
        if (er_raw >= 0 && t > 0) {
            const double a = raw[B_CL * BOXP + er_raw], b = raw[B_CV * BOXP + er_raw];
            const double kI = ((VDIM == 0 ? b : a) * wIr) * hk;
            const double kJ = ((VDIM == 1 ? b : a) * wJr) * hk;
            const double kK = ((VDIM == 2 ? b : a) * wKr) * rk;
            Coef4 c;
            c.sij = kI + kJ; c.ui = fma(-2., kI, kJ); c.uj = fma(-2., kJ, kI); c.kk = kK;
            reinterpret_cast<Coef4*>(sCb)[er_c] = c;
        }
        __syncthreads();
        if (tid == 0 && t + NS <= nsteps) issue(t + NS);

        // ---------------- phase 2: gather layer L between planes a (registers) and b --------
        double wb[RJ + 2][3];
#pragma unroll
        for (int y = 0; y < RJ + 2; ++y)
#pragma unroll
            for (int c = 0; c < 3; ++c) wb[y][c] = sPb[(jl0 + y) * PWP + tx + c];
        if (t > 0) {
            double dw[RJ + 2][3];
#pragma unroll
            for (int y = 0; y < RJ + 2; ++y)
#pragma unroll
                for (int c = 0; c < 3; ++c) dw[y][c] = wb[y][c] - wa[y][c];
            double La[RJ], Lb[RJ], Cc[RJ], S[RJ], Tt[RJ], A0[RJ], A1[RJ], P0[RJ], P1[RJ];
#pragma unroll
            for (int rr = 0; rr < RJ; ++rr) { La[rr] = Lb[rr] = Cc[rr] = S[rr] = Tt[rr] = A0[rr] = A1[rr] = P0[rr] = P1[rr] = 0.; }
#pragma unroll
            for (int ey = 0; ey <= RJ; ++ey) {          // element row between window rows ey and ey+1
                const Coef4 e0 = reinterpret_cast<const Coef4*>(sCb)[(jl0 + ey) * CW + tx];
                const Coef4 e1 = reinterpret_cast<const Coef4*>(sCb)[(jl0 + ey) * CW + tx + 1];
                const double Bs = e0.uj + e1.uj, Qs = e0.kk + e1.kk, Ss = e0.sij + e1.sij;
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    // side 0: node rr = ey (centre row ey+1, other row ey); side 1: node rr = ey-1 (centre ey, other ey+1)
                    const int rr = side ? ey - 1 : ey;
                    if (rr < 0 || rr >= RJ) continue;
                    const int yo = side ? ey + 1 : ey;
                    La[rr] = fma(Bs, wa[yo][1], La[rr]); La[rr] = fma(-e0.sij, wa[yo][0], La[rr]); La[rr] = fma(-e1.sij, wa[yo][2], La[rr]);
                    Lb[rr] = fma(Bs, wb[yo][1], Lb[rr]); Lb[rr] = fma(-e0.sij, wb[yo][0], Lb[rr]); Lb[rr] = fma(-e1.sij, wb[yo][2], Lb[rr]);
                    Cc[rr] = fma(2. * Qs, dw[yo][1], Cc[rr]); Cc[rr] = fma(e0.kk, dw[yo][0], Cc[rr]); Cc[rr] = fma(e1.kk, dw[yo][2], Cc[rr]);
                    S[rr] += Ss; Tt[rr] += Qs;
                    A0[rr] += e0.ui; A1[rr] += e1.ui; P0[rr] += e0.kk; P1[rr] += e1.kk;
                }
            }
#pragma unroll
            for (int rr = 0; rr < RJ; ++rr) {
                const int yc = rr + 1;
                const double S2 = 2. * S[rr];
                double la = fma(S2, wa[yc][1], La[rr]); la = fma(A0[rr], wa[yc][0], la); la = fma(A1[rr], wa[yc][2], la);
                double lb = fma(S2, wb[yc][1], Lb[rr]); lb = fma(A0[rr], wb[yc][0], lb); lb = fma(A1[rr], wb[yc][2], lb);
                double cc = fma(4. * Tt[rr], dw[yc][1], Cc[rr]);
                cc = fma(2. * P0[rr], dw[yc][0], cc); cc = fma(2. * P1[rr], dw[yc][2], cc);
                const double lo = fma(2., la, lb) - cc;
                const double hi = fma(2., lb, la) + cc;
                const int Pa = P - 1;
                if (Pa >= k0 && vj[rr]) {       // finalise the plane below (always < k1 here)
                    const double qv = (da[rr] == 0.) ? 0. : carry[rr] + lo;
                    q_out[i + g.sJ * (j0 + jl0 + rr) + g.sK * Pa] = qv;
                    red[0] = fma(wa[yc][1], qv, red[0]);
                    if (FUSED) {
                        red[2] = fma(qv, za[rr], red[2]);
                        red[3] = fma(qv * qv, da[rr], red[3]);
                    }
                }
                carry[rr] = hi;
            }
        }
#pragma unroll
        for (int y = 0; y < RJ + 2; ++y)
#pragma unroll
            for (int c = 0; c < 3; ++c) wa[y][c] = wb[y][c];
#pragma unroll
        for (int rr = 0; rr < RJ; ++rr) { za[rr] = zb[rr]; da[rr] = db[rr]; }
    }
    if (!FUSED) return;
    if (grid_reduce<NRED, false>(red, partials, &sc->ticket[0], sRed, sh_flag)) {
        if (tid == 0) {
            const double pq = red[0], rho = red[1], qz = red[2], qdq = red[3], rr = red[4], zz = red[5], xx = red[6];
            const double rho_old = sc->rho;
            sc->rho_prev = rho_old; sc->rho = rho; sc->pq = pq; sc->rr = rr; sc->zz = zz; sc->xx = xx;
            const int launch = sc->launch + 1;
            sc->launch = launch;
            const int it = sc->bench ? launch : launch - 1;   // launch m has applied m-1 updates to x
            sc->iter = it;
            bool stop = false;
            if (!sc->bench) {
                if (!(rr == rr) || !(pq == pq)) { sc->done = 1; sc->status = -2; stop = true; }
                else if (rr <= sc->tol2 * sc->bb && rho <= sc->tol2 * sc->bz && zz <= sc->tol2 * xx) { sc->done = 1; sc->status = 1; stop = true; }
                else if (it >= sc->maxit) { sc->done = 1; sc->status = 2; stop = true; }
                else if (!(pq > 0.)) { sc->done = 1; sc->status = -1; stop = true; }
            }
            if (!stop) {
                const double al = (pq > 0.) ? rho / pq : 0.;
                const double rho_next = fma(al * al, qdq, fma(-2. * al, qz, rho));
                sc->alpha = al;
                sc->beta = (rho_next > 0. && rho > 0.) ? rho_next / rho : 0.;
            }
        }
    }
}

// ---------------------------------------------------------------------- host side -------

struct FusedPlan {
    bool valid;
    int tj, rj, ns, minb;
    int lk, tilesI, tilesJ, chunksK;
    CUtensorMap m_r[2], m_q[2], m_p[2], m_d, m_cl, m_cv;
    char why[160];
};

static inline FusedPlan make_fused_plan(const Grid& g, int sm_count, double* const r[2], double* const q[2], double* const p[2],
                                        double* dinv, double* cl, double* cv) {
    FusedPlan f;
    memset(&f, 0, sizeof(f));
    f.tj = 8; f.rj = 1; f.ns = 3; f.minb = 2;
    int lk = 0;
    const char* env = getenv("PFEM_FUSED_TILE");  // "tj,rj,ns,minb,lk" for tuning runs
    if (env) {
        int a, b, c, d, e;
        int got = sscanf(env, "%d,%d,%d,%d,%d", &a, &b, &c, &d, &e);
        if (got >= 4) { f.tj = a; f.rj = b; f.ns = c; f.minb = d; }
        if (got >= 5) lk = e;
    }
    f.tilesI = (g.nI + 31) / 32;
    f.tilesJ = (g.nJ + f.tj - 1) / f.tj;
    if (lk <= 0) {
        // CTAs resident per wave: 2 per SM; aim for >= 4 waves but keep >= 16 planes per CTA
        const long long tiles = (long long)f.tilesI * f.tilesJ;
        const long long want = 8LL * sm_count;
        long long chunks = (want + tiles - 1) / tiles;
        if (chunks < 1) chunks = 1;
        lk = (int)((g.nK + chunks - 1) / chunks);
        if (lk < 16) lk = 16;
        if (lk > g.nK) lk = g.nK;
    }
    f.lk = lk;
    f.chunksK = (g.nK + lk - 1) / lk;
    if ((g.sJ * 8) % 16 != 0 || (g.sK * 8) % 16 != 0) { snprintf(f.why, sizeof f.why, "row pitch is not a multiple of 16 bytes"); return f; }
    const int bw = 32 + 4, bh = f.tj + 2;
    bool ok = true;
    for (int b = 0; b < 2; ++b)
        ok = ok && make_lattice_map(&f.m_r[b], r[b], g.nI, g.nJ, g.nK, g.sJ, g.sK, bw, bh) &&
             make_lattice_map(&f.m_q[b], q[b], g.nI, g.nJ, g.nK, g.sJ, g.sK, bw, bh) &&
             make_lattice_map(&f.m_p[b], p[b], g.nI, g.nJ, g.nK, g.sJ, g.sK, bw, bh);
    ok = ok && make_lattice_map(&f.m_d, dinv, g.nI, g.nJ, g.nK, g.sJ, g.sK, bw, bh) &&
         make_lattice_map(&f.m_cl, cl, g.nI - 1, g.nJ - 1, g.nK - 1, g.sJ, g.sK, bw, bh) &&
         make_lattice_map(&f.m_cv, cv, g.nI - 1, g.nJ - 1, g.nK - 1, g.sJ, g.sK, bw, bh);
    if (!ok) { snprintf(f.why, sizeof f.why, "cuTensorMapEncodeTiled failed or is unavailable"); return f; }
    f.valid = true;
    return f;
}

template <int TJ, int RJ, int NS, int MINB, int VDIM, bool FUSED>
static inline cudaError_t launch_fused_inst(const FusedPlan& f, const Grid& g, int par, double* r_out, double* q_out, double* p_out,
                                            double* x, Scalars* sc, double* partials, cudaStream_t st) {
    const size_t smem = FusedTile<TJ>::smem_bytes(NS, FUSED);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_fpcg<TJ, RJ, NS, MINB, VDIM, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    dim3 grid(f.tilesI, f.tilesJ, f.chunksK), block(32, TJ / RJ, 1);
    k_fpcg<TJ, RJ, NS, MINB, VDIM, FUSED><<<grid, block, smem, st>>>(f.m_r[par], f.m_q[par], f.m_p[par], f.m_d, f.m_cl, f.m_cv, g, f.lk,
                                                                r_out, q_out, p_out, x, sc, partials);
    return cudaGetLastError();
}

template <int TJ, int RJ, int NS, int MINB, bool FUSED>
static inline cudaError_t launch_fused_vdim(const FusedPlan& f, const Grid& g, int par, double* r_out, double* q_out, double* p_out,
                                            double* x, Scalars* sc, double* partials, cudaStream_t st) {
    switch (g.vdim) {
        case 0: return launch_fused_inst<TJ, RJ, NS, MINB, 0, FUSED>(f, g, par, r_out, q_out, p_out, x, sc, partials, st);
        case 1: return launch_fused_inst<TJ, RJ, NS, MINB, 1, FUSED>(f, g, par, r_out, q_out, p_out, x, sc, partials, st);
        default: return launch_fused_inst<TJ, RJ, NS, MINB, 2, FUSED>(f, g, par, r_out, q_out, p_out, x, sc, partials, st);
    }
}

// par: which of the double buffers holds the INPUT vectors r, q, p (outputs go to the other one).
// FUSED = false: plain q_out = M A p with p = p[par] (tests, pfem_apply).
template <bool FUSED>
static inline cudaError_t launch_fused_dispatch(const FusedPlan& f, const Grid& g, int par, double* r_out, double* q_out,
                                                double* p_out, double* x, Scalars* sc, double* partials, cudaStream_t st) {
#define PFEM_FUSED_CASE(TJ, RJ, NS, MINB) \
    if (f.tj == TJ && f.rj == RJ && f.ns == NS && f.minb == MINB) return launch_fused_vdim<TJ, RJ, NS, MINB, FUSED>(f, g, par, r_out, q_out, p_out, x, sc, partials, st);
    PFEM_FUSED_CASE(8, 1, 3, 2)
    PFEM_FUSED_CASE(8, 1, 2, 2)
    PFEM_FUSED_CASE(8, 2, 3, 2)
    PFEM_FUSED_CASE(8, 2, 2, 2)
    PFEM_FUSED_CASE(8, 2, 2, 3)
    PFEM_FUSED_CASE(16, 2, 2, 1)
    PFEM_FUSED_CASE(16, 2, 2, 2)
    PFEM_FUSED_CASE(16, 1, 2, 1)
    PFEM_FUSED_CASE(16, 4, 2, 1)
#undef PFEM_FUSED_CASE
    return cudaErrorInvalidConfiguration;
}

}  // namespace pfem
