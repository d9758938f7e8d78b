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

#include <algorithm>
#include <functional>
#include <type_traits>
#include <vector>

#include "kernels_tma.cuh"
#include "kernels_ml.cuh"

#ifndef PFEM_X
#define PFEM_X 45   // experiment mask of k_fpcg: 1 = phase-1 loads hoisted (-1.7 %), 2 = x fetched two planes ahead in registers (+1.5 %: off),
                    // 4 = vertical stiffness per element (-0.5 %), 8 = x of the own tile through the TMA stage instead of a global load (-3 %),
                    // 32 = isotropic coefficient layer from per-thread constants (-0.4 ... -1.0 % inside the power-capped bench);
                    // tried and dropped: a branch-free phase 2 that guards only the stores (+1.1 %); TMA L2 prefetch
                    // (cp.async.bulk.prefetch.tensor) of the boxes 1 / 2 / 4 steps beyond the shared-memory ring: 0.227 -> 0.295 / 0.299 / 0.318 ms
                    // per iteration in the graph (the prefetches queue in front of the next step's real loads; profiles/r02_prefetch_ab.md)
#endif

namespace pfem {

// Chunks of the march along K: chunk c of every tile covers the owned planes [off[c], off[c+1]) (relative to kown0).
// The lengths DECREASE with c (make_fused_plan): the hardware hands CTAs to free SM slots in blockIdx order, so long chunks go
// first and short ones fill the tail of the launch.
#define PFEM_MAX_CHUNKS 32
struct ChunkTab {
    int n, lkmax;
    int off[PFEM_MAX_CHUNKS + 1];
};

template <int TJ>
struct FusedTile {
    static constexpr int TI = 32, HX = 2;
    static constexpr int PW = TI + 2 * HX, PH = TJ + 2;   // raw TMA box (doubles)
    static constexpr int BOX = PW * PH;
    static constexpr int BOXP = (BOX + 15) / 16 * 16;
    static constexpr int PWP = TI + 2;                   // p' plane pitch, rows PH
    static constexpr int PLANE = (PWP * PH + 15) / 16 * 16;
    static constexpr int CW = TI + 1, CH = TJ + 1;       // coefficient layer: two double2 planes per element,
    static constexpr int CHALF = (CW * CH * 2 + 15) / 16 * 16;   // (sij, ui) then (uj, kk): conflict-free LDS.128
    static constexpr int LAYER = 2 * CHALF;
    static constexpr int NRED = 7;
    static constexpr size_t smem_bytes(int ns, int mode, int lk, bool iso = false) {
        return 128 + sizeof(double) * ((size_t)ns * ((mode == 1 ? 6 : mode >= 2 ? 5 : 4) - (iso ? 1 : 0)) * BOXP + (((PFEM_X & 9) == 9) && mode >= 1 ? (size_t)ns * TI * TJ : 0) +
                                       2 * PLANE + 2 * LAYER + 32 * NRED + 2 * (size_t)(lk + 2)) +
               16 * 8 + 16;
    }
};


// MODE 0: plain q = M A p (tests, pfem_apply);  1: the fused Jacobi-PCG iteration;  3: MODE 2 with z = z_0 + z_1(parent aggregate)
// (multilevel preconditioner, kernels_ml.cuh; vertical axis = I only);  2: operator step of the line-Jacobi
// iteration (kernels_line.cuh): boxes z, p, mask instead of r, q, p, D^-1 — p' = mask z + beta p, x' = x + alpha p, q' = M A p'.
// ISO: c_lat == c_vert in every element (bulk materials: thermk returns the same lateral and vertical value, e.g. GaAs.cpp:225-228):
// the c_vert layer is not loaded at all — 10 instead of 11 words per DOF.
// MASS: the operator is A + diag(mass) — the time step of Dynamic3D with a lumped capacity matrix (femT3d.cpp:203-212;
// the conductivities arrive scaled by methodparam); mass is read for the own nodes only, one step ahead of its use.
template <int TJ, int RJ, int NS, int MINB, int VDIM, int MODE, bool ISO = false, bool MASS = false>
__global__ void __launch_bounds__(32 * (TJ / RJ), MINB)
k_fpcg(const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_q,
       const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_d,
       const __grid_constant__ CUtensorMap tm_cl, const __grid_constant__ CUtensorMap tm_cv, const __grid_constant__ CUtensorMap tm_x,
       const Grid g, const ChunkTab ck,
       double* __restrict__ r_out, double* __restrict__ q_out, double* __restrict__ p_out, double* __restrict__ x,
       Scalars* sc, double* partials, const PeerOut po, const CoarseAdd ca, const double* __restrict__ mass) {
    typedef FusedTile<TJ> T;
    constexpr int TI = T::TI, HX = T::HX, PW = T::PW, PH = T::PH, BOX = T::BOX, BOXP = T::BOXP, PWP = T::PWP;
    constexpr int PLANE = T::PLANE, CW = T::CW, CHALF = T::CHALF, LAYER = T::LAYER, NRED = T::NRED;
    constexpr int NT = TI * (TJ / RJ);
    constexpr bool FUSED = MODE >= 1, LINE = MODE >= 2, MLZ = MODE == 3;
    constexpr int NBN = LINE ? 3 : FUSED ? 4 : 2;   // node boxes per stage: r q p d | z p mask | p d
    constexpr int NCB = ISO ? 1 : 2;      // coefficient boxes: c_lat (, c_vert)
    constexpr int NB = NBN + NCB;
    constexpr int B_P = LINE ? 1 : FUSED ? 2 : 0, B_D = B_P + 1, B_CL = NBN, B_CV = NBN + 1;
    constexpr int NRING = 2 * PWP + 2 * TJ;     // halo ring of the p' plane
    constexpr int NERING = TI + TJ + 1;         // halo row/column of the coefficient layer

    extern __shared__ __align__(128) unsigned char smem_dyn[];
    double* const sRaw = reinterpret_cast<double*>(smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u));  // [NS][NB][BOXP]
    constexpr bool XTMA = ((PFEM_X & 9) == 9) && FUSED;   // x of the own tile arrives with the stage (box TI x TJ, no halo) instead of a per-thread
    constexpr int XBOX = TI * TJ;                  // global load one step ahead, whose HBM latency under load exceeds a step (ncu: long_scoreboard)
    double* const sX = sRaw + (size_t)NS * NB * BOXP;   // [NS][XBOX] (XTMA)
    double* const sP = sX + (XTMA ? (size_t)NS * XBOX : 0);   // [2][PLANE]
    double* const sC = sP + 2 * PLANE;                  // [2][LAYER]
    double* const sRed = sC + 2 * LAYER;                // [32*NRED]
    uint64_t* const bars = reinterpret_cast<uint64_t*>(sRed + 32 * NRED);  // [NS]
    int* const sh_flag = reinterpret_cast<int*>(bars + 16);
    double* const sHK = reinterpret_cast<double*>(bars + 18);   // [lk+2] hK and [lk+2] 1/hK of the element layers of this chunk

    // Programmatic dependent launch: the next launch of the stream may become resident as soon as every CTA of this grid has
    // passed this point and SM slots free up; everything up to pdl_wait() below touches only data that is constant during a
    // linear solve (mesh spacings, D^-1 / mask, conductivities), so that prologue runs under the tail of the previous launch
    // (its last CTAs, the grid reduction and — in slab mode — the cross-rank scalar exchange).
    if (FUSED) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = tx + TI * ty;
    const int i0 = blockIdx.x * TI, j0 = blockIdx.y * TJ;
    const int lk = ck.lkmax;
    const int k0 = g.kown0 + ck.off[blockIdx.z];   // owned planes only (slab mode: halo planes belong to the neighbours)
    const int k1 = g.kown0 + ck.off[blockIdx.z + 1];
    const int nsteps = k1 - k0 + 1;   // items t = 0 .. nsteps: node planes k0-1 .. k1, element layers k0-1 .. k1-1
    const int jl0 = ty * RJ;
    // slab mode: the first / last owned plane is also stored into the neighbour's halo plane (NVLink peer stores)
    const bool push_lo = FUSED && po.p_lo != nullptr, push_hi = FUSED && po.p_hi != nullptr;
    const bool slab = push_lo || push_hi;

    // ---- per-thread constants -----------------------------------------------------------
    const int i = i0 + tx;
    const bool vi = i < g.nI;
    constexpr double s36 = 1e-6 / 36.;
    // element weights: kI = cI wI hk, kJ = cJ wJ hk, kK = cK wK rk   (therm3d.cpp:215-220)
    constexpr bool ISOW = ISO && ((PFEM_X & 33) == 33);   // 6 instead of 9 FP64 operations per element for the coefficient layer
    double wI[RJ], wJ[RJ], wK[RJ], wU[RJ];
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
            if (ISOW) {   // isotropic conductivity: the three stencil coefficients are (c hk) times constants of the thread
                const double a = wI[rr], b = wJ[rr];
                wI[rr] = a + b;              // kI + kJ
                wJ[rr] = fma(-2., a, b);     // kJ - 2 kI
                wU[rr] = fma(-2., b, a);     // kI - 2 kJ
            }
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
    // multilevel preconditioner (MODE 3, vertical axis = I): z = z_0 + z_1(parent) + z_2(parent).  The coarse values of a node
    // change only every 4th / 16th plane of the march: they are kept in registers and re-fetched one step AHEAD of the plane
    // that needs them, so the L2 latency of the gather never sits on the critical path of a step.
    int zj_own[RJ], zj_ring = 0, zi_own = 0, zi_ring = 0;
    double zc[RJ], zcr = 0., zt_own = 0., zt_ring = 0.;
    if (MLZ) {
        zi_own = min(i, g.nI - 1);
#pragma unroll
        for (int rr = 0; rr < RJ; ++rr) { zj_own[rr] = min(j0 + jl0 + rr, g.nJ - 1); zc[rr] = 0.; }
        if (ring_raw >= 0) {
            const int jj = ring_pl / PWP, ii = ring_pl % PWP;
            zj_ring = min(max(j0 - 1 + jj, 0), g.nJ - 1);
            zi_ring = min(max(i0 - 1 + ii, 0), g.nI - 1);
        }
    }
    auto load_zc = [&](const int P) {   // coarse correction of node plane P
        const int Pc = min(max(P, 0), g.nK - 1) + ca.koff;
        const double* const a1 = ca.z1 + (idx_t)(Pc >> ca.sh1) * ca.nJ1 * g.sJ;
        const double* const a2 = ca.z2 + (idx_t)(Pc >> ca.sh2) * ca.nJ2 * g.sJ;
#pragma unroll
        for (int rr = 0; rr < RJ; ++rr)
            zc[rr] = (__ldg(a1 + (idx_t)(zj_own[rr] >> ca.sh1) * g.sJ + zi_own) + __ldg(a2 + (idx_t)(zj_own[rr] >> ca.sh2) * g.sJ + zi_own)) + zt_own;
        if (ring_raw >= 0) zcr = (__ldg(a1 + (idx_t)(zj_ring >> ca.sh1) * g.sJ + zi_ring) + __ldg(a2 + (idx_t)(zj_ring >> ca.sh2) * g.sJ + zi_ring)) + zt_ring;
    };
    // halo row / column of the element layer handled by this thread (entry e = NT-1-tid)
    int er_raw = -1, er_c = 0;
    double wIr = 0., wJr = 0., wKr = 0., wUr = 0.;
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
        if (ISOW) { const double a = wIr, b = wJr; wIr = a + b; wJr = fma(-2., a, b); wUr = fma(-2., b, a); }
    }

    // part 1: boxes that are constant during a linear solve (D^-1 or mask, conductivities) + the expected byte count of the
    // whole stage; part 2: the iteration vectors.  issue() = both.
    auto issue_const = [&](int t) {
        const int st = t % NS;
        double* dst = sRaw + (size_t)st * NB * BOXP;
        uint64_t* bar = &bars[st];
        mbar_expect_tx(bar, (uint32_t)(((NBN + (t > 0 ? NCB : 0)) * BOX + (XTMA ? XBOX : 0)) * sizeof(double)));
        const int P = k0 - 1 + t;
        tma_load_3d(dst + B_D * BOXP, &tm_d, bar, i0 - HX, j0 - 1, P);
        if (t > 0) {
            tma_load_3d(dst + B_CL * BOXP, &tm_cl, bar, i0 - HX, j0 - 1, P - 1);
            if (!ISO) tma_load_3d(dst + B_CV * BOXP, &tm_cv, bar, i0 - HX, j0 - 1, P - 1);
        }
    };
    auto issue_vec = [&](int t) {
        const int st = t % NS;
        double* dst = sRaw + (size_t)st * NB * BOXP;
        uint64_t* bar = &bars[st];
        const int P = k0 - 1 + t;
        if (FUSED) tma_load_3d(dst, &tm_r, bar, i0 - HX, j0 - 1, P);
        if (FUSED && !LINE) tma_load_3d(dst + BOXP, &tm_q, bar, i0 - HX, j0 - 1, P);
        tma_load_3d(dst + B_P * BOXP, &tm_p, bar, i0 - HX, j0 - 1, P);
        if (XTMA) tma_load_3d(sX + (size_t)st * XBOX, &tm_x, bar, i0, j0, P);
    };
    auto issue = [&](int t) { issue_const(t); issue_vec(t); };

    // spacings of the element layers k0-2 .. k1-1 (entry t = layer of step t), staged once: a global load per
    // step would sit on the critical path of every warp
    for (int t = tid; t <= nsteps; t += NT) {
        const int Lc = min(max(k0 - 2 + t, -1), g.nK - 1);
        sHK[t] = g.hK[Lc];
        sHK[lk + 2 + t] = g.rK[Lc];
    }
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int t = 0; t < NS && t <= nsteps; ++t) issue_const(t);
    if (FUSED) asm volatile("griddepcontrol.wait;" ::: "memory");   // the previous launch has completed and its writes are visible
    if (tid == 0)
        for (int t = 0; t < NS && t <= nsteps; ++t) issue_vec(t);
    const int done_in = FUSED ? sc->done : 0;
    if (done_in == 1) {   // a finished solve (done == 2, line-Jacobi PCG: the pending x update is still to be applied):
        if (tid == 0)     // let the boxes in flight land before the shared memory is given back
            for (int t = 0; t < NS && t <= nsteps; ++t) mbar_wait(&bars[t], 0u);
        return;
    }
    const double alpha = FUSED ? sc->alpha : 0.;
    const double beta = FUSED ? sc->beta : 0.;

    // ---- state carried along the march ----------------------------------------------------
    // The p' windows and (z, D^-1) of the own nodes ping-pong between two register sets so that
    // "plane above" becomes "plane below" without moves: step(t) reads set A (plane below), fills set B.
    double w0[RJ + 2][3], w1[RJ + 2][3];
    double zd0[2 * RJ], zd1[2 * RJ];   // z[rr], D^-1[rr]
    double carry[RJ];                  // contribution of the layer below to the current plane
#pragma unroll
    for (int y = 0; y < RJ + 2; ++y)
#pragma unroll
        for (int c = 0; c < 3; ++c) { w0[y][c] = 0.; w1[y][c] = 0.; }
#pragma unroll
    for (int rr = 0; rr < RJ; ++rr) { carry[rr] = 0.; zd0[rr] = zd0[RJ + rr] = zd1[rr] = zd1[RJ + rr] = 0.; }
    double red[NRED];              // pq, rz, qz, qdq, rr, zz, xx
#pragma unroll
    for (int a = 0; a < NRED; ++a) red[a] = 0.;

    // lattice index of the own nodes in the plane of the current step (advanced by sK per step)
    idx_t nown[RJ];
#pragma unroll
    for (int rr = 0; rr < RJ; ++rr) nown[rr] = i + g.sJ * (j0 + jl0 + rr) + g.sK * (idx_t)(k0 - 1);
    const idx_t sK = g.sK;

    // x of the own nodes, fetched one step ahead
    double xn[RJ], xn2[RJ];   // PFEM_X & 2: two planes ahead (an HBM round trip under load is longer than one step)
#pragma unroll
    for (int rr = 0; rr < RJ; ++rr) xn[rr] = xn2[rr] = 0.;
    double mprev[RJ];   // MASS: capacity diagonal of the own nodes of the plane below (loaded when that plane was the current one)
#pragma unroll
    for (int rr = 0; rr < RJ; ++rr) mprev[rr] = 0.;

    // OWN: plane P = k0-1+t is owned (store r', p', x');  NEXT_OWN: plane P+1 is owned (prefetch x)
    // GATHER: t >= 1 (element layer P-1 exists);  FINAL: plane P-1 is owned (store q', dots)
    auto step = [&](auto own_c, auto next_own_c, auto gather_c, auto final_c, const int t, double (&wa)[RJ + 2][3],
                    double (&wb)[RJ + 2][3], double (&zda)[2 * RJ], double (&zdb)[2 * RJ]) {
        constexpr bool OWN = decltype(own_c)::value, NEXT_OWN = decltype(next_own_c)::value;
        constexpr bool GATHER = decltype(gather_c)::value, FINAL = decltype(final_c)::value;
        const int st = t % NS;
        const double* raw = sRaw + (size_t)st * NB * BOXP;
        double* sPb = sP + (t & 1) * PLANE;
        double* sCb = sC + (t & 1) * LAYER;
        double xcur[RJ];
        if (FUSED && OWN && !XTMA) {
#pragma unroll
            for (int rr = 0; rr < RJ; ++rr) xcur[rr] = xn[rr];
        }
        if (XTMA) {
        } else if (PFEM_X & 2) {
            if (FUSED) {
#pragma unroll
                for (int rr = 0; rr < RJ; ++rr) xn[rr] = xn2[rr];
                if (t + 2 <= nsteps - 1) {
#pragma unroll
                    for (int rr = 0; rr < RJ; ++rr)
                        if (vj[rr]) xn2[rr] = x[nown[rr] + 2 * sK];
                }
                if (t == 0 && NEXT_OWN) {   // start-up: the first owned plane has not been fetched by an earlier step
#pragma unroll
                    for (int rr = 0; rr < RJ; ++rr)
                        if (vj[rr]) xn[rr] = x[nown[rr] + sK];
                }
            }
        } else if (FUSED && NEXT_OWN) {
#pragma unroll
            for (int rr = 0; rr < RJ; ++rr)
                if (vj[rr]) xn[rr] = x[nown[rr] + sK];
        }
        double mcur[RJ];   // MASS: capacity diagonal of plane P (used by the next step)
        if (MASS) {
#pragma unroll
            for (int rr = 0; rr < RJ; ++rr) mcur[rr] = (OWN && vj[rr]) ? mass[nown[rr]] : 0.;
        }
        const double hk = sHK[t], rk = sHK[lk + 2 + t];   // element layer L = P - 1 of this step (t >= 1)
        mbar_wait(&bars[st], (uint32_t)((t / NS) & 1));

#if PFEM_X & 1
        // ---------------- phase 1: p' plane and coefficient layer -------------------------
        // all shared-memory loads of the phase are issued before the first store: the compiler cannot prove that the raw boxes
        // and the p' / coefficient planes do not alias and would otherwise serialise load -> use -> store per node
        double l_r[RJ], l_q[RJ], l_p[RJ], l_d[RJ], l_a[RJ], l_b[RJ];
        double g_r = 0., g_q = 0., g_p = 0., g_d = 0., h_a = 0., h_b = 0.;
#pragma unroll
        for (int rr = 0; rr < RJ; ++rr) {
            const int ro = (jl0 + rr + 1) * PW + tx + HX;
            l_r[rr] = FUSED ? raw[ro] : 0.;
            l_q[rr] = (FUSED && !LINE) ? raw[BOXP + ro] : 0.;
            l_p[rr] = raw[B_P * BOXP + ro];
            l_d[rr] = raw[B_D * BOXP + ro];
            if (GATHER) { l_a[rr] = raw[B_CL * BOXP + ro]; l_b[rr] = ISO ? l_a[rr] : raw[B_CV * BOXP + ro]; }
            if (XTMA && OWN) xcur[rr] = sX[(size_t)st * XBOX + (jl0 + rr) * TI + tx];
        }
        if (ring_raw >= 0) {
            g_p = raw[B_P * BOXP + ring_raw];
            if (FUSED) { g_r = raw[ring_raw]; g_q = LINE ? 0. : raw[BOXP + ring_raw]; g_d = raw[B_D * BOXP + ring_raw]; }
        }
        if (GATHER && er_raw >= 0) { h_a = raw[B_CL * BOXP + er_raw]; h_b = ISO ? h_a : raw[B_CV * BOXP + er_raw]; }
#pragma unroll
        for (int rr = 0; rr < RJ; ++rr) {
            double pn;
            if (FUSED) {
                const double r0 = l_r[rr], q0 = l_q[rr], p0 = l_p[rr], dd = l_d[rr];
                const double rn = MLZ ? r0 + zc[rr] : LINE ? r0 : fma(-alpha, q0, r0);
                const double z = dd * rn;
                pn = fma(beta, p0, z);
                zdb[rr] = z; zdb[RJ + rr] = dd;
                if (OWN) {
                    if (vj[rr]) {
                        const idx_t n = nown[rr];
                        const double xv = fma(alpha, p0, xcur[rr]);
                        if (!LINE) r_out[n] = rn;   // line-Jacobi iteration: r' is written by the line kernel
                        p_out[n] = pn;
                        x[n] = xv;
                        if (slab) {
                            const int P = k0 - 1 + t;
                            const idx_t nip = n - sK * P;   // offset inside the plane
                            if (push_lo && P == g.kown0) { if (!LINE) po.r_lo[nip] = rn; po.p_lo[nip] = pn; }
                            if (push_hi && P == g.kown1 - 1) { if (!LINE) po.r_hi[nip] = rn; po.p_hi[nip] = pn; }
                        }
                        if (!LINE) {
                            red[1] = fma(rn, z, red[1]);
                            red[4] = fma(rn, rn, red[4]);
                            red[5] = fma(z, z, red[5]);
                        }
                        if (MLZ) red[5] = fma(z, z, red[5]);
                        red[6] = fma(xv, xv, red[6]);
                    }
                }
            } else {
                pn = l_p[rr];
                zdb[rr] = 0.; zdb[RJ + rr] = l_d[rr];
            }
            sPb[(jl0 + rr + 1) * PWP + tx + 1] = pn;
            if (GATHER) {
                const double a = l_a[rr], b = l_b[rr];
                const int co = (jl0 + rr + 1) * CW + tx + 1;
                if (ISOW) {
                    const double ah = a * hk;
                    reinterpret_cast<double2*>(sCb)[co] = make_double2(ah * wI[rr], ah * wJ[rr]);
                    reinterpret_cast<double2*>(sCb + CHALF)[co] = make_double2(ah * wU[rr], (a * rk) * wK[rr]);
                } else {
                    const double kI = ((VDIM == 0 ? b : a) * wI[rr]) * hk;
                    const double kJ = ((VDIM == 1 ? b : a) * wJ[rr]) * hk;
                    const double kK = ((VDIM == 2 ? b : a) * wK[rr]) * rk;
                    reinterpret_cast<double2*>(sCb)[co] = make_double2(kI + kJ, fma(-2., kI, kJ));
                    reinterpret_cast<double2*>(sCb + CHALF)[co] = make_double2(fma(-2., kJ, kI), kK);
                }
            }
        }
        if (ring_raw >= 0) {
            double pn;
            if (FUSED) pn = fma(beta, g_p, g_d * (MLZ ? g_r + zcr : LINE ? g_r : fma(-alpha, g_q, g_r)));
            else pn = g_p;
            sPb[ring_pl] = pn;
        }
        if (GATHER && er_raw >= 0) {
            const double a = h_a, b = h_b;
            if (ISOW) {
                const double ah = a * hk;
                reinterpret_cast<double2*>(sCb)[er_c] = make_double2(ah * wIr, ah * wJr);
                reinterpret_cast<double2*>(sCb + CHALF)[er_c] = make_double2(ah * wUr, (a * rk) * wKr);
            } else {
                const double kI = ((VDIM == 0 ? b : a) * wIr) * hk;
                const double kJ = ((VDIM == 1 ? b : a) * wJr) * hk;
                const double kK = ((VDIM == 2 ? b : a) * wKr) * rk;
                reinterpret_cast<double2*>(sCb)[er_c] = make_double2(kI + kJ, fma(-2., kI, kJ));
                reinterpret_cast<double2*>(sCb + CHALF)[er_c] = make_double2(fma(-2., kJ, kI), kK);
            }
        }
#else
        // ---------------- phase 1: p' plane and coefficient layer -------------------------
#pragma unroll
        for (int rr = 0; rr < RJ; ++rr) {
            const int ro = (jl0 + rr + 1) * PW + tx + HX;
            double pn;
            if (FUSED) {
                const double r0 = raw[ro], q0 = LINE ? 0. : raw[BOXP + ro], p0 = raw[B_P * BOXP + ro], dd = raw[B_D * BOXP + ro];
                const double rn = MLZ ? r0 + zc[rr] : LINE ? r0 : fma(-alpha, q0, r0);
                const double z = dd * rn;
                pn = fma(beta, p0, z);
                zdb[rr] = z; zdb[RJ + rr] = dd;
                if (OWN) {
                    if (vj[rr]) {
                        const idx_t n = nown[rr];
                        const double xv = fma(alpha, p0, xcur[rr]);
                        if (!LINE) r_out[n] = rn;   // line-Jacobi iteration: r' is written by the line kernel
                        p_out[n] = pn;
                        x[n] = xv;
                        if (slab) {
                            const int P = k0 - 1 + t;
                            const idx_t nip = n - sK * P;   // offset inside the plane
                            if (push_lo && P == g.kown0) { if (!LINE) po.r_lo[nip] = rn; po.p_lo[nip] = pn; }
                            if (push_hi && P == g.kown1 - 1) { if (!LINE) po.r_hi[nip] = rn; po.p_hi[nip] = pn; }
                        }
                        if (!LINE) {
                            red[1] = fma(rn, z, red[1]);
                            red[4] = fma(rn, rn, red[4]);
                            red[5] = fma(z, z, red[5]);
                        }
                        if (MLZ) red[5] = fma(z, z, red[5]);
                        red[6] = fma(xv, xv, red[6]);
                    }
                }
            } else {
                pn = raw[B_P * BOXP + ro];
                zdb[rr] = 0.; zdb[RJ + rr] = raw[B_D * BOXP + ro];
            }
            sPb[(jl0 + rr + 1) * PWP + tx + 1] = pn;
            if (GATHER) {
                const double a = raw[B_CL * BOXP + ro], b = ISO ? a : raw[B_CV * BOXP + ro];
                const double kI = ((VDIM == 0 ? b : a) * wI[rr]) * hk;
                const double kJ = ((VDIM == 1 ? b : a) * wJ[rr]) * hk;
                const double kK = ((VDIM == 2 ? b : a) * wK[rr]) * rk;
                const int co = (jl0 + rr + 1) * CW + tx + 1;
                reinterpret_cast<double2*>(sCb)[co] = make_double2(kI + kJ, fma(-2., kI, kJ));
                reinterpret_cast<double2*>(sCb + CHALF)[co] = make_double2(fma(-2., kJ, kI), kK);
            }
        }
        if (ring_raw >= 0) {
            double pn;
            if (FUSED) {
                const double r0 = raw[ring_raw], q0 = LINE ? 0. : raw[BOXP + ring_raw], p0 = raw[B_P * BOXP + ring_raw],
                             dd = raw[B_D * BOXP + ring_raw];
                pn = fma(beta, p0, dd * (MLZ ? r0 + zcr : LINE ? r0 : fma(-alpha, q0, r0)));
            } else {
                pn = raw[B_P * BOXP + ring_raw];
            }
            sPb[ring_pl] = pn;
        }
        if (GATHER && er_raw >= 0) {
            const double a = raw[B_CL * BOXP + er_raw], b = ISO ? a : raw[B_CV * BOXP + er_raw];
            const double kI = ((VDIM == 0 ? b : a) * wIr) * hk;
            const double kJ = ((VDIM == 1 ? b : a) * wJr) * hk;
            const double kK = ((VDIM == 2 ? b : a) * wKr) * rk;
            reinterpret_cast<double2*>(sCb)[er_c] = make_double2(kI + kJ, fma(-2., kI, kJ));
            reinterpret_cast<double2*>(sCb + CHALF)[er_c] = make_double2(fma(-2., kJ, kI), kK);
        }
#endif
        __syncthreads();
        if (tid == 0 && t + NS <= nsteps) issue(t + NS);
        if (MLZ) {   // the next plane lies in another level-1 aggregate: fetch its coarse correction now, use it next step
            const int Pn = min(max(k0 + t, 0), g.nK - 1), Pc = min(max(k0 - 1 + t, 0), g.nK - 1);
            if (((Pn + ca.koff) >> ca.sh1) != ((Pc + ca.koff) >> ca.sh1)) load_zc(Pn);
        }

        // ---------------- phase 2: gather layer L between planes a (registers) and b --------
#pragma unroll
        for (int y = 0; y < RJ + 2; ++y)
#pragma unroll
            for (int c = 0; c < 3; ++c) wb[y][c] = sPb[(jl0 + y) * PWP + tx + c];
        if (GATHER) {
            // coefficients of the RJ+1 element rows x 2 element columns around the own nodes
            double e_sij[RJ + 1][2], e_ui[RJ + 1][2], e_uj[RJ + 1][2], e_kk[RJ + 1][2];
#pragma unroll
            for (int ey = 0; ey <= RJ; ++ey) {          // element row between window rows ey and ey+1
                const int co = (jl0 + ey) * CW + tx;
#pragma unroll
                for (int si = 0; si < 2; ++si) {
                    const double2 ea = reinterpret_cast<const double2*>(sCb)[co + si];
                    const double2 eb = reinterpret_cast<const double2*>(sCb + CHALF)[co + si];
                    e_sij[ey][si] = ea.x; e_ui[ey][si] = ea.y; e_uj[ey][si] = eb.x; e_kk[ey][si] = eb.y;
                }
            }
            double dw[RJ + 2][3];   // b - a
#pragma unroll
            for (int y = 0; y < RJ + 2; ++y)
#pragma unroll
                for (int c = 0; c < 3; ++c) dw[y][c] = wb[y][c] - wa[y][c];
            double Bs[RJ + 1], Ss[RJ + 1];   // sums over the two elements of a row
#pragma unroll
            for (int ey = 0; ey <= RJ; ++ey) {
                Bs[ey] = e_uj[ey][0] + e_uj[ey][1];
                Ss[ey] = e_sij[ey][0] + e_sij[ey][1];
            }
#if PFEM_X & 4
            double mw[RJ + 2][2];
#pragma unroll
            for (int y = 0; y < RJ + 2; ++y) { mw[y][0] = fma(2., dw[y][1], dw[y][0]); mw[y][1] = fma(2., dw[y][1], dw[y][2]); }
#else
            double Qs[RJ + 1];
#pragma unroll
            for (int ey = 0; ey <= RJ; ++ey) Qs[ey] = e_kk[ey][0] + e_kk[ey][1];
#endif
#pragma unroll
            for (int rr = 0; rr < RJ; ++rr) {
                const int y0 = rr, yc = rr + 1, y2 = rr + 2;   // window rows; element rows rr (below) and rr+1 (above)
                const double S2 = 2. * (Ss[rr] + Ss[rr + 1]);
                const double A0 = e_ui[rr][0] + e_ui[rr + 1][0], A1 = e_ui[rr][1] + e_ui[rr + 1][1];
                // in-plane stiffness of planes a and b: three independent chains each
                double la = fma(A1, wa[yc][2], fma(A0, wa[yc][0], S2 * wa[yc][1]));
                double la0 = fma(-e_sij[rr][1], wa[y0][2], fma(-e_sij[rr][0], wa[y0][0], Bs[rr] * wa[y0][1]));
                double la2 = fma(-e_sij[rr + 1][1], wa[y2][2], fma(-e_sij[rr + 1][0], wa[y2][0], Bs[rr + 1] * wa[y2][1]));
                double lb = fma(A1, wb[yc][2], fma(A0, wb[yc][0], S2 * wb[yc][1]));
                double lb0 = fma(-e_sij[rr][1], wb[y0][2], fma(-e_sij[rr][0], wb[y0][0], Bs[rr] * wb[y0][1]));
                double lb2 = fma(-e_sij[rr + 1][1], wb[y2][2], fma(-e_sij[rr + 1][0], wb[y2][0], Bs[rr + 1] * wb[y2][1]));
                la += la0 + la2;
                lb += lb0 + lb2;
                // vertical stiffness M9 (b - a) on the difference window
#if PFEM_X & 4
                // element by element: kK_e (4 d11 + 2 d_iedge + 2 d_jedge + d_corner) = kK_e (2 m[yc][s] + m[yo][s]) with the row
                // combinations m[y][s] = 2 d[y][1] + d[y][2s] shared by the nodes of the thread
                const double cc = fma(e_kk[rr][0], fma(2., mw[yc][0], mw[y0][0]), e_kk[rr][1] * fma(2., mw[yc][1], mw[y0][1])) +
                                  fma(e_kk[rr + 1][0], fma(2., mw[yc][0], mw[y2][0]), e_kk[rr + 1][1] * fma(2., mw[yc][1], mw[y2][1]));
#else
                const double T4 = 4. * (Qs[rr] + Qs[rr + 1]);
                const double P0 = 2. * (e_kk[rr][0] + e_kk[rr + 1][0]), P1 = 2. * (e_kk[rr][1] + e_kk[rr + 1][1]);
                const double Q0 = 2. * Qs[rr], Q1 = 2. * Qs[rr + 1];
                double cc = fma(P1, dw[yc][2], fma(P0, dw[yc][0], T4 * dw[yc][1]));
                const double cc0 = fma(e_kk[rr][1], dw[y0][2], fma(e_kk[rr][0], dw[y0][0], Q0 * dw[y0][1]));
                const double cc2 = fma(e_kk[rr + 1][1], dw[y2][2], fma(e_kk[rr + 1][0], dw[y2][0], Q1 * dw[y2][1]));
                cc += cc0 + cc2;
#endif
                const double lo = fma(2., la, lb) - cc;
                const double hi = fma(2., lb, la) + cc;
                if (FINAL) {
                    if (vj[rr]) {       // finalise the plane below
                        const double da = zda[RJ + rr];
                        const double qv = (da == 0.) ? 0. : MASS ? fma(mprev[rr], wa[yc][1], carry[rr] + lo) : carry[rr] + lo;
                        q_out[nown[rr] - sK] = qv;
                        if (slab && !LINE) {   // the line-Jacobi iteration reads q on owned rows only
                            const int Pa = k0 - 2 + t;
                            const idx_t nip = nown[rr] - sK - sK * Pa;
                            if (push_lo && Pa == g.kown0) po.q_lo[nip] = qv;
                            if (push_hi && Pa == g.kown1 - 1) po.q_hi[nip] = qv;
                        }
                        red[0] = fma(wa[yc][1], qv, red[0]);
                        if (FUSED && !LINE) {
                            red[2] = fma(qv, zda[rr], red[2]);
                            red[3] = fma(qv * qv, da, red[3]);
                        }
                    }
                }
                carry[rr] = hi;
            }
        }
#pragma unroll
        for (int rr = 0; rr < RJ; ++rr) nown[rr] += sK;
        if (MASS) {
#pragma unroll
            for (int rr = 0; rr < RJ; ++rr) mprev[rr] = mcur[rr];
        }
    };

    if (MLZ) {   // written by the level kernels of this iteration: after pdl wait
        zt_own = __ldg(ca.zt + zi_own);
        if (ring_raw >= 0) zt_ring = __ldg(ca.zt + zi_ring);
        load_zc(k0 - 1);
    }
    {
        typedef std::true_type Y;
        typedef std::false_type N;
        // t = 0: halo plane k0-1;  t = 1: first owned plane;  t = 2 .. nsteps-1: owned planes, all flags on;
        // t = nsteps: halo plane k1.  nsteps >= 2.  Even t fills w0 from w1, odd t fills w1 from w0.
        step(N(), Y(), N(), N(), 0, w1, w0, zd1, zd0);
        if (nsteps == 2) {
            step(Y(), N(), Y(), N(), 1, w0, w1, zd0, zd1);
        } else {
            step(Y(), Y(), Y(), N(), 1, w0, w1, zd0, zd1);
            int t = 2;
            for (; t + 1 < nsteps - 1; t += 2) {
                step(Y(), Y(), Y(), Y(), t, w1, w0, zd1, zd0);
                step(Y(), Y(), Y(), Y(), t + 1, w0, w1, zd0, zd1);
            }
            // one or two owned planes left; the last owned plane (t = nsteps-1) has no owned successor
            if (t < nsteps - 1) { step(Y(), Y(), Y(), Y(), t, w1, w0, zd1, zd0); ++t; }
            if (t & 1) step(Y(), N(), Y(), Y(), t, w0, w1, zd0, zd1);
            else step(Y(), N(), Y(), Y(), t, w1, w0, zd1, zd0);
        }
        if (nsteps & 1) step(N(), N(), Y(), Y(), nsteps, w0, w1, zd0, zd1);
        else step(N(), N(), Y(), Y(), nsteps, w1, w0, zd1, zd0);
    }
    if (!FUSED) return;
    // system-scope fence only in the CTAs that stored into a neighbour's halo planes (first / last chunk); the others publish
    // their partials with a device-scope fence, the last CTA's own system fence before the flag covers them by cumulativity
    const bool cta_pushes = (push_lo && k0 == g.kown0) || (push_hi && k1 == g.kown1);
    if (grid_reduce<NRED, false>(red, partials, &sc->ticket[0], sRed, sh_flag, cta_pushes)) {
        if (sc->comm) rank_allreduce<NRED, false>(red, sc->comm, sRed);
        if (tid == 0 && LINE) {
            // line-Jacobi PCG: rho, beta and the stopping test belong to the line kernel; here only alpha = rho / p'.q'
            const double pq = red[0];
            sc->pq = pq; sc->xx = red[6];
            if (MLZ) sc->zz = red[5];   // multilevel: the complete z exists only here (stopping test of the next iteration)
            if (sc->done == 2) sc->done = 1;
            else if (sc->surf) {}   // convection terms: k_surf_iter adds p'.S p' and computes alpha
            else if (!sc->bench && !(pq > 0.)) { sc->done = 1; sc->status = (pq == pq) ? -1 : -2; }
            else sc->alpha = (pq > 0.) ? sc->rho / pq : 0.;
        } else if (tid == 0) {
            const double pq = red[0], rho = red[1], qz = red[2], qdq = red[3], rr = red[4], zz = red[5], xx = red[6];
            const double rho_old = sc->rho;
            sc->rho_prev = rho_old; sc->rho = rho; sc->pq = pq; sc->rr = rr; sc->zz = zz; sc->xx = xx;
            const int surf = sc->surf;   // boundary-face matrix terms: k_surf_iter adds S p' to q' and takes over from here
            const int launch = sc->launch + 1;
            sc->launch = launch;
            const int it = sc->bench ? launch : launch - 1;   // launch m has applied m-1 updates to x
            sc->iter = it;
            bool stop = false;
            if (!sc->bench) {
                if (sc->comm && sc->comm->timeout) { sc->done = 1; sc->status = -3; stop = true; }
                else if (!(rr == rr) || !(pq == pq)) { sc->done = 1; sc->status = -2; stop = true; }
                else if (rr <= sc->tol2 * sc->bb && rho <= sc->tol2 * sc->bz && zz <= sc->tol2 * xx) { sc->done = 1; sc->status = 1; stop = true; }
                else if (it >= sc->maxit) { sc->done = 1; sc->status = 2; stop = true; }
                else if (!surf && !(pq > 0.)) { sc->done = 1; sc->status = -1; stop = true; }
            }
            if (surf) { sc->qz = qz; sc->qdq = qdq; }
            else if (!stop) {
                const double al = (pq > 0.) ? rho / pq : 0.;
                const double rho_next = fma(al * al, qdq, fma(-2. * al, qz, rho));
                sc->alpha = al;
                sc->beta = (rho_next > 0. && rho > 0.) ? rho_next / rho : 0.;
            }
        }
    }
}

// ---------------------------------------------------------------------- host side -------

static inline bool ck_ok(const int* v, int n, int lmax) {
    if (n < 1 || n > PFEM_MAX_CHUNKS) return false;
    for (int c = 0; c < n; ++c) if (v[c] < 1 || v[c] > lmax) return false;
    return true;
}

// Model of one launch: `tiles` CTAs per chunk index, handed out in blockIdx order (chunk-major) to `resident` SM slots, a CTA
// with l owned planes taking l + 2 (halo planes) + 1.5 (start-up) plane steps.  Returns the makespan in plane steps.
static inline double chunk_makespan(const int* len, int n, int tiles, int resident, std::vector<double>& heap) {
    heap.assign((size_t)resident, 0.);
    for (int c = 0; c < n; ++c) {
        const double cost = len[c] + 3.5;
        for (int t = 0; t < tiles; ++t) {
            std::pop_heap(heap.begin(), heap.end(), std::greater<double>());
            heap.back() += cost;
            std::push_heap(heap.begin(), heap.end(), std::greater<double>());
        }
    }
    return *std::max_element(heap.begin(), heap.end());
}

// Chunk lengths along K.  Uniform chunks leave the last wave of CTAs partly empty (256^3 on 148 x 3 slots: 1280 CTAs = 2.9
// waves, 8 % of the SM time idle in ncu); lengths that decrease with the chunk index let the short chunks fill that tail.
// Geometric start sequences, then a deterministic hill climb on the modelled makespan.  fixed_lk > 0: uniform chunks of that length.
static inline void plan_chunks(int nown, int tiles, int resident, int fixed_lk, ChunkTab& ck) {
    const int LMAX = 510;   // <= 8 KB of staged layer spacings per CTA
    int best[PFEM_MAX_CHUNKS], nbest = 0;
    auto finish = [&]() {
        ck.n = nbest; ck.lkmax = 0; ck.off[0] = 0;
        for (int c = 0; c < nbest; ++c) { ck.off[c + 1] = ck.off[c] + best[c]; if (best[c] > ck.lkmax) ck.lkmax = best[c]; }
    };
    const char* env = getenv("PFEM_FUSED_CHUNKS");   // "l0,l1,..." for tuning runs (must sum to the owned planes)
    if (env && !fixed_lk) {
        int sum = 0; nbest = 0;
        for (const char* s = env; *s && nbest < PFEM_MAX_CHUNKS;) {
            char* e; long v = strtol(s, &e, 10);
            if (e == s || v <= 0) break;
            best[nbest++] = (int)v; sum += (int)v; s = (*e == ',') ? e + 1 : e;
        }
        if (sum == nown && ck_ok(best, nbest, LMAX)) { finish(); return; }
    }
    if (fixed_lk > 0 || getenv("PFEM_FUSED_UNIFORM")) {
        int l = fixed_lk;
        if (l <= 0) {   // round-1 rule: whole waves against the two re-staged halo planes
            double be = -1.;
            for (int c = 1; c <= nown; ++c) {
                const int lc = (nown + c - 1) / c;
                if (lc < 8 && c > 1) break;
                const long long ctas = (long long)tiles * ((nown + lc - 1) / lc), waves = (ctas + resident - 1) / resident;
                const double eff = (double)ctas / (double)(waves * resident) * (double)lc / (double)(lc + 2);
                if (eff > be + 1e-9) { be = eff; l = lc; }
            }
        }
        if (l > LMAX) l = LMAX;
        if ((nown + l - 1) / l > PFEM_MAX_CHUNKS) l = (nown + PFEM_MAX_CHUNKS - 1) / PFEM_MAX_CHUNKS;
        nbest = 0;
        for (int k = 0; k < nown; k += l) best[nbest++] = (nown - k < l) ? nown - k : l;
        finish();
        return;
    }
    std::vector<double> heap;
    double bestT = 1e300;
    int cur[PFEM_MAX_CHUNKS];
    const int lmin = nown >= 32 ? 4 : 1;
    auto normalise = [&](int* v, int n) {   // decreasing order
        std::sort(v, v + n, std::greater<int>());
    };
    auto consider = [&](int* v, int n) {
        if (!ck_ok(v, n, LMAX)) return false;
        const double T = chunk_makespan(v, n, tiles, resident, heap);
        if (T < bestT - 1e-9) { bestT = T; nbest = n; memcpy(best, v, sizeof(int) * n); return true; }
        return false;
    };
    const int nmaxc = std::min(PFEM_MAX_CHUNKS, std::max(1, nown / lmin));
    const int nminc = (nown + LMAX - 1) / LMAX;
    static const double ratios[] = {1.0, 0.85, 0.7, 0.55, 0.4, 0.3};
    for (int n = nminc; n <= nmaxc; ++n)
        for (double rho : ratios) {
            if (rho < 1. && n > 12) break;
            double w[PFEM_MAX_CHUNKS], sw = 0., x = 1.;
            for (int c = 0; c < n; ++c) { w[c] = x; sw += x; x *= rho; }
            int sum = 0;
            for (int c = 0; c < n; ++c) { cur[c] = std::max(lmin, (int)(nown * w[c] / sw)); sum += cur[c]; }
            if (sum > nown) { cur[0] -= sum - nown; if (cur[0] < lmin) continue; }
            for (int c = 0; sum < nown; c = (c + 1) % n) { ++cur[c]; ++sum; }
            normalise(cur, n);
            consider(cur, n);
        }
    if (nbest == 0) { best[0] = nown; nbest = 1; if (!ck_ok(best, 1, LMAX)) { plan_chunks(nown, tiles, resident, std::min(LMAX, nown), ck); return; } }
    // hill climb: move d planes from chunk a to chunk b
    unsigned lcg = 12345u;
    const int budget = tiles * (long long)nbest > 20000 ? 60 : 400;
    for (int it = 0; it < budget && nbest > 1; ++it) {
        lcg = lcg * 1664525u + 1013904223u;
        const int a = (int)((lcg >> 8) % (unsigned)nbest);
        lcg = lcg * 1664525u + 1013904223u;
        int b = (int)((lcg >> 8) % (unsigned)nbest);
        if (a == b) b = (b + 1) % nbest;
        lcg = lcg * 1664525u + 1013904223u;
        const int d = 1 << ((lcg >> 8) % 4u);
        memcpy(cur, best, sizeof(int) * nbest);
        if (cur[a] - d < lmin) continue;
        cur[a] -= d; cur[b] += d;
        normalise(cur, nbest);
        consider(cur, nbest);
    }
    finish();
}

struct FusedPlan {
    bool valid;
    bool iso;            // launch the ISO instantiation (set per launch by the host: c_lat == c_vert everywhere)
    const double* mass;  // non-null: launch the MASS instantiation (Dynamic3D time step, lumped capacity diagonal)
    bool pdl;            // launch the iteration kernels with programmatic stream serialization (PFEM_NO_PDL=1 turns it off)
    int tj, rj, ns, minb;
    int lk, tilesI, tilesJ, chunksK;
    ChunkTab ck;
    CUtensorMap m_r[2], m_q[2], m_p[2], m_d, m_cl, m_cv, m_x;
    char why[160];
};

static inline FusedPlan make_fused_plan(const Grid& g, int sm_count, double* const r[2], double* const q[2], double* const p[2],
                                        double* dinv, double* cl, double* cv, double* x) {
    FusedPlan f;
    memset(&f, 0, sizeof(f));
    f.tj = 8; f.rj = 2; f.ns = 2; f.minb = 3;   // measured best on B200 at 256^3 (tools/tune_fused.py)
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
    plan_chunks(g.kown1 - g.kown0, f.tilesI * f.tilesJ, sm_count * f.minb, lk, f.ck);
    f.lk = f.ck.lkmax;
    f.chunksK = f.ck.n;
    if (getenv("PFEM_DEBUG_PLAN")) {
        fprintf(stderr, "[pfem] fused plan: %d x %d tiles, %d chunks:", f.tilesI, f.tilesJ, f.ck.n);
        for (int c = 0; c < f.ck.n; ++c) fprintf(stderr, " %d", f.ck.off[c + 1] - f.ck.off[c]);
        fprintf(stderr, "\n");
    }
    if ((g.sJ * 8) % 16 != 0 || (g.sK * 8) % 16 != 0) { snprintf(f.why, sizeof f.why, "row pitch is not a multiple of 16 bytes"); return f; }
    const int bw = 32 + 4, bh = f.tj + 2;
    bool ok = true;
    for (int b = 0; b < 2; ++b)
        ok = ok && make_lattice_map(&f.m_r[b], r[b], g.nI, g.nJ, g.nK, g.sJ, g.sK, bw, bh) &&
             make_lattice_map(&f.m_q[b], q[b], g.nI, g.nJ, g.nK, g.sJ, g.sK, bw, bh) &&
             make_lattice_map(&f.m_p[b], p[b], g.nI, g.nJ, g.nK, g.sJ, g.sK, bw, bh);
    ok = ok && make_lattice_map(&f.m_d, dinv, g.nI, g.nJ, g.nK, g.sJ, g.sK, bw, bh) &&
         make_lattice_map(&f.m_cl, cl, g.nI - 1, g.nJ - 1, g.nK - 1, g.sJ, g.sK, bw, bh) &&
         make_lattice_map(&f.m_cv, cv, g.nI - 1, g.nJ - 1, g.nK - 1, g.sJ, g.sK, bw, bh) &&
         make_lattice_map(&f.m_x, x, g.nI, g.nJ, g.nK, g.sJ, g.sK, 32, f.tj);
    if (!ok) { snprintf(f.why, sizeof f.why, "cuTensorMapEncodeTiled failed or is unavailable"); return f; }
    f.pdl = getenv("PFEM_NO_PDL") == nullptr;
    f.valid = true;
    return f;
}

template <int TJ, int RJ, int NS, int MINB, int VDIM, int MODE, bool ISO, bool MASS = false>
static inline cudaError_t launch_fused_inst2(const FusedPlan& f, const Grid& g, int par, double* r_out, double* q_out, double* p_out,
                                            double* x, Scalars* sc, double* partials, const PeerOut& po, cudaStream_t st, const CoarseAdd& ca) {
    const size_t smem = FusedTile<TJ>::smem_bytes(NS, MODE, f.lk, ISO);
    static size_t attr_done[64] = {};   // the opt-in is per function AND per device (contexts may live on different GPUs)
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_done[dev & 63] < smem) {
        cudaError_t e = cudaFuncSetAttribute(k_fpcg<TJ, RJ, NS, MINB, VDIM, MODE, ISO, MASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done[dev & 63] = smem;
    }
    dim3 grid(f.tilesI, f.tilesJ, f.chunksK), block(32, TJ / RJ, 1);
    if (MODE >= 1 && f.pdl) {   // programmatic dependent launch (see the kernel prologue); also valid inside stream capture
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, k_fpcg<TJ, RJ, NS, MINB, VDIM, MODE, ISO, MASS>, f.m_r[par], f.m_q[par], f.m_p[par], f.m_d, f.m_cl, f.m_cv, f.m_x, g,
                                  f.ck, r_out, q_out, p_out, x, sc, partials, po, ca, f.mass);
    }
    k_fpcg<TJ, RJ, NS, MINB, VDIM, MODE, ISO, MASS><<<grid, block, smem, st>>>(f.m_r[par], f.m_q[par], f.m_p[par], f.m_d, f.m_cl, f.m_cv, f.m_x, g, f.ck,
                                                                      r_out, q_out, p_out, x, sc, partials, po, ca, f.mass);
    return cudaGetLastError();
}

// the ISO instantiation exists for the production tiles and the iteration modes only
template <int TJ, int RJ, int NS, int MINB, int VDIM, int MODE>
static inline cudaError_t launch_fused_inst(const FusedPlan& f, const Grid& g, int par, double* r_out, double* q_out, double* p_out,
                                            double* x, Scalars* sc, double* partials, const PeerOut& po, cudaStream_t st, const CoarseAdd& ca) {
    if constexpr ((MODE == 1 || MODE == 2) && TJ == 8 && RJ == 2 && NS == 2 && MINB == 3) {   // Dynamic3D: general-conductivity kernel + capacity diagonal
        if (f.mass) return launch_fused_inst2<TJ, RJ, NS, MINB, VDIM, MODE, false, true>(f, g, par, r_out, q_out, p_out, x, sc, partials, po, st, ca);
    }
    if (f.mass) return cudaErrorInvalidConfiguration;
    if constexpr (MODE >= 1 && TJ == 8 && RJ == 2 && MINB == 3) {
        if (f.iso) return launch_fused_inst2<TJ, RJ, NS, MINB, VDIM, MODE, true>(f, g, par, r_out, q_out, p_out, x, sc, partials, po, st, ca);
    }
    return launch_fused_inst2<TJ, RJ, NS, MINB, VDIM, MODE, false>(f, g, par, r_out, q_out, p_out, x, sc, partials, po, st, ca);
}

template <int TJ, int RJ, int NS, int MINB, int MODE>
static inline cudaError_t launch_fused_vdim(const FusedPlan& f, const Grid& g, int par, double* r_out, double* q_out, double* p_out,
                                            double* x, Scalars* sc, double* partials, const PeerOut& po, cudaStream_t st, const CoarseAdd& ca) {
    if constexpr (MODE == 3) {   // multilevel preconditioner: vertical axis = I only
        if (g.vdim != 0) return cudaErrorInvalidConfiguration;
        return launch_fused_inst<TJ, RJ, NS, MINB, 0, MODE>(f, g, par, r_out, q_out, p_out, x, sc, partials, po, st, ca);
    } else {
        switch (g.vdim) {
            case 0: return launch_fused_inst<TJ, RJ, NS, MINB, 0, MODE>(f, g, par, r_out, q_out, p_out, x, sc, partials, po, st, ca);
            case 1: return launch_fused_inst<TJ, RJ, NS, MINB, 1, MODE>(f, g, par, r_out, q_out, p_out, x, sc, partials, po, st, ca);
            default: return launch_fused_inst<TJ, RJ, NS, MINB, 2, MODE>(f, g, par, r_out, q_out, p_out, x, sc, partials, po, st, ca);
        }
    }
}

// par: which of the double buffers holds the INPUT vectors r, q, p (outputs go to the other one).
// MODE 0: plain q_out = M A p with p = p[par] (tests, pfem_apply); 2: line-Jacobi operator step (f.m_r = z, f.m_d = mask).
template <int MODE>
static inline cudaError_t launch_fused_dispatch(const FusedPlan& f, const Grid& g, int par, double* r_out, double* q_out,
                                                double* p_out, double* x, Scalars* sc, double* partials, const PeerOut& po,
                                                cudaStream_t st, const CoarseAdd ca = CoarseAdd{nullptr, 0, 0, nullptr, 0, 0, nullptr, 0}) {
#define PFEM_FUSED_CASE(TJ, RJ, NS, MINB) \
    if (f.tj == TJ && f.rj == RJ && f.ns == NS && f.minb == MINB) return launch_fused_vdim<TJ, RJ, NS, MINB, MODE>(f, g, par, r_out, q_out, p_out, x, sc, partials, po, st, ca);
    PFEM_FUSED_CASE(8, 2, 2, 3)    // production tile (tools/tune_fused.py); the others are kept for PFEM_FUSED_TILE tuning runs
    PFEM_FUSED_CASE(8, 2, 3, 3)
    PFEM_FUSED_CASE(16, 2, 2, 2)
#undef PFEM_FUSED_CASE
    return cudaErrorInvalidConfiguration;
}

}  // namespace pfem
