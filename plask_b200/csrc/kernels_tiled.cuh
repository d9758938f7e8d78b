// kernels_tiled.cuh — production operator kernel of the PCG iteration.
//
//   p_new = D^-1 r + beta p_old          (fused, recomputed on the tile halo)
//   q     = M A p_new                    (matrix-free 27-point brick operator, FP64)
//   p.q   -> alpha = rho / p.q           (warp shuffles + deterministic last-block reduction)
//
// One CTA owns a TJ x TI tile of the (medium, minor) plane and marches through LK node planes
// of the major axis.  Per element layer it stages into shared memory
//   * the new node plane of p_new (tile + 1-node halo), computed on the fly from r, D^-1, p_old
//   * the three directional conductances k_I,k_J,k_K (already divided by 36) of the layer's
//     elements (tile + 1-element halo), computed once per CTA from (c_lat, c_vert) and the
//     1-D mesh spacings,
// and every thread gathers the contribution of the layer to its node(s) in the plane below
// (finalised and stored) and in the plane above (carried in a register).  Using the tensor
// structure K_e = kI S(x)M(x)M + kJ M(x)S(x)M + kK M(x)M(x)S (S = [1,-1;-1,1], M = [2,1;1,2]/6,
// therm3d.cpp:226-237) the layer between planes a (below) and b (above) contributes
//   to plane a:  L(2a + b) - C(b - a)        to plane b:  L(a + 2b) + C(b - a)
// with the in-plane 9-point operators L = sum_e kI tI + kJ tJ and C = sum_e kK m.
//
// Algorithmic traffic per node: reads r, D^-1, p_old, c_lat, c_vert; writes p_new, q = 7 words.
// p is double-buffered (p_old and p_new are different arrays) because halo nodes of p_old are
// read by neighbouring CTAs.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pfem_internal.cuh"

namespace pfem {

struct TiledPlan {
    bool valid;
    int ti, tj, rj;  // tile width (I), tile height (J), rows per thread
    int lk;          // node planes per CTA
    int tilesI, tilesJ, chunksK;
    size_t smem;
};

template <int TI, int TJ>
struct TileSmem {
    static constexpr int PW = TI + 2;  // node plane pitch
    static constexpr int PH = TJ + 2;
    static constexpr int CW = TI + 1;  // element plane pitch
    static constexpr int CH = TJ + 1;
    static constexpr size_t bytes = sizeof(double) * (size_t)(2 * PW * PH + 3 * CW * CH) + 64 * sizeof(double);
};

// FUSED: p_new = dinv*r + beta*p_old is computed while staging (and stored for owned nodes) and
//        the p.q reduction / alpha update is performed.  !FUSED: p is read from pin as is, q = M A p.
template <int TI, int TJ, int RJ, bool FUSED>
__global__ void __launch_bounds__(TI*(TJ / RJ))
k_apply_tiled(const Grid g, const int lk, const double* __restrict__ cl, const double* __restrict__ cv,
              const double* __restrict__ r, const double* __restrict__ dinv, const double* __restrict__ pin,
              double* __restrict__ pout, double* __restrict__ q, Scalars* sc, double* partials) {
    typedef TileSmem<TI, TJ> S;
    constexpr int NT = TI * (TJ / RJ);
    extern __shared__ double smem[];
    double* sP = smem;                          // [2][PH][PW] node planes (ping-pong)
    double* sC = smem + 2 * S::PW * S::PH;      // [3][CH][CW] kI,kJ,kK of the current layer
    double* sRed = sC + 3 * S::CW * S::CH;      // 64 doubles for reductions
    __shared__ int sh_flag;

    if (FUSED && sc->done) return;
    const double beta = FUSED ? sc->beta : 0.;

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = tx + TI * ty;
    const int i0 = blockIdx.x * TI, j0 = blockIdx.y * TJ;
    const int k0 = blockIdx.z * lk;
    const int k1 = min(k0 + lk, g.nK);

    // stage node plane kp (tile + halo) into buffer `buf`
    auto stage_plane = [&](int kp, int buf) {
        double* dst = sP + buf * S::PW * S::PH;
        const bool kin = (kp >= 0 && kp < g.nK);
        for (int m = tid; m < S::PW * S::PH; m += NT) {
            const int jj = m / S::PW, ii = m - jj * S::PW;
            const int i = i0 + ii - 1, j = j0 + jj - 1;
            double v = 0.;
            if (kin && i >= 0 && i < g.nI && j >= 0 && j < g.nJ) {
                const idx_t n = i + g.sJ * j + g.sK * kp;
                if (FUSED) {
                    v = dinv[n] * r[n] + beta * pin[n];
                    if (ii >= 1 && ii <= TI && jj >= 1 && jj <= TJ && kp >= k0 && kp < k1) pout[n] = v;
                } else {
                    v = pin[n];
                }
            }
            dst[m] = v;
        }
    };
    // stage the conductances of element layer ek (tile + halo) scaled by 1/36
    auto stage_layer = [&](int ek) {
        const bool kin = (ek >= 0 && ek < g.nK - 1);
        const double hk = kin ? g.hK[ek] : 1., rk = kin ? g.rK[ek] : 1.;
        const double s = 1e-6 / 36.;
        for (int m = tid; m < S::CW * S::CH; m += NT) {
            const int jj = m / S::CW, ii = m - jj * S::CW;
            const int ei = i0 + ii - 1, ej = j0 + jj - 1;
            double kI = 0., kJ = 0., kK = 0.;
            if (kin && ei >= 0 && ei < g.nI - 1 && ej >= 0 && ej < g.nJ - 1) {
                const idx_t slot = ei + g.sJ * ej + g.sK * ek;
                const double a = cl[slot], b = cv[slot];
                const double hi = g.hI[ei], hj = g.hJ[ej];
                const double cI = (g.vdim == 0 ? b : a) * s, cJ = (g.vdim == 1 ? b : a) * s, cK = (g.vdim == 2 ? b : a) * s;
                kI = cI * (hj * hk) * g.rI[ei];
                kJ = cJ * (hi * hk) * g.rJ[ej];
                kK = cK * (hi * hj) * rk;
            }
            sC[m] = kI;
            sC[S::CW * S::CH + m] = kJ;
            sC[2 * S::CW * S::CH + m] = kK;
        }
    };

    double carry[RJ];
#pragma unroll
    for (int rr = 0; rr < RJ; ++rr) carry[rr] = 0.;
    double dot = 0.;

    // layer L lies between node planes L (buffer L&1 ... see below) and L+1
    stage_plane(k0 - 1, 0);
    int cur = 0;  // buffer holding plane L
    for (int L = k0 - 1; L < k1; ++L) {
        stage_plane(L + 1, cur ^ 1);
        stage_layer(L);
        __syncthreads();
        const double* Pa = sP + cur * S::PW * S::PH;        // plane L   (below)
        const double* Pb = sP + (cur ^ 1) * S::PW * S::PH;  // plane L+1 (above)
#pragma unroll
        for (int rr = 0; rr < RJ; ++rr) {
            const int jl = ty * RJ + rr;  // local node row 0..TJ-1; smem row jl+1, column tx+1
            double a[3][3], b[3][3];
#pragma unroll
            for (int dj = 0; dj < 3; ++dj)
#pragma unroll
                for (int di = 0; di < 3; ++di) {
                    a[dj][di] = Pa[(jl + dj) * S::PW + tx + di];
                    b[dj][di] = Pb[(jl + dj) * S::PW + tx + di];
                }
            double lo = 0., hi = 0., cc = 0.;  // L(2a+b), L(a+2b), C(b-a)
#pragma unroll
            for (int sj = 0; sj < 2; ++sj)
#pragma unroll
                for (int si = 0; si < 2; ++si) {
                    // element on side (sj,si): rows jl+sj (elements) = node rows {jl+sj, jl+sj+1} in smem terms
                    const int ce = (jl + sj) * S::CW + tx + si;
                    const double kI = sC[ce], kJ = sC[S::CW * S::CH + ce], kK = sC[2 * S::CW * S::CH + ce];
                    const int on = sj ? 2 : 0;  // other row in the 3x3 window
                    const int cn = si ? 2 : 0;  // neighbour column
                    // tI = 2 (u[1][1]-u[1][cn]) + (u[on][1]-u[on][cn]);  tJ = 2 (u[1][1]-u[on][1]) + (u[1][cn]-u[on][cn])
                    const double tIa = 2. * (a[1][1] - a[1][cn]) + (a[on][1] - a[on][cn]);
                    const double tIb = 2. * (b[1][1] - b[1][cn]) + (b[on][1] - b[on][cn]);
                    const double tJa = 2. * (a[1][1] - a[on][1]) + (a[1][cn] - a[on][cn]);
                    const double tJb = 2. * (b[1][1] - b[on][1]) + (b[1][cn] - b[on][cn]);
                    const double m = 4. * (b[1][1] - a[1][1]) + 2. * ((b[1][cn] - a[1][cn]) + (b[on][1] - a[on][1])) +
                                     (b[on][cn] - a[on][cn]);
                    lo += kI * (2. * tIa + tIb) + kJ * (2. * tJa + tJb);
                    hi += kI * (tIa + 2. * tIb) + kJ * (tJa + 2. * tJb);
                    cc += kK * m;
                }
            if (L >= k0) {
                const int i = i0 + tx, j = j0 + jl;
                if (i < g.nI && j < g.nJ) {
                    const idx_t n = i + g.sJ * j + g.sK * L;
                    const double qv = (dinv[n] == 0.) ? 0. : carry[rr] + lo - cc;
                    q[n] = qv;
                    dot += a[1][1] * qv;
                }
            }
            carry[rr] = hi + cc;
        }
        __syncthreads();
        cur ^= 1;
    }
    if (!FUSED) return;
    double v[1] = {dot};
    if (grid_reduce<1, false>(v, partials, &sc->ticket[0], sRed, &sh_flag)) {
        if (tid == 0) {
            sc->pq = v[0];
            if (v[0] > 0.) sc->alpha = sc->rho / v[0];
            else {
                sc->alpha = 0.;
                if (!sc->bench) { sc->done = 1; sc->status = (v[0] == v[0]) ? -1 : -2; }
            }
        }
    }
}

// ---------------------------------------------------------------------- host side -------

static inline TiledPlan make_tiled_plan(const Grid& g, int sm_count) {
    TiledPlan p;
    memset(&p, 0, sizeof(p));
    p.ti = 32; p.tj = 16; p.rj = 2;
    const char* env = getenv("PFEM_TILE");  // "ti,tj,rj,lk" for tuning runs
    int lk = 0;
    if (env) {
        int a, b, c, d;
        if (sscanf(env, "%d,%d,%d,%d", &a, &b, &c, &d) == 4) { p.ti = a; p.tj = b; p.rj = c; lk = d; }
    }
    p.tilesI = (g.nI + p.ti - 1) / p.ti;
    p.tilesJ = (g.nJ + p.tj - 1) / p.tj;
    if (lk <= 0) {
        // enough CTAs for ~8 per SM, but planes per CTA >= 8 to bound the (lk+2)/lk re-staging
        const long long tiles = (long long)p.tilesI * p.tilesJ;
        const long long want = 8LL * sm_count;
        long long chunks = (want + tiles - 1) / tiles;
        if (chunks < 1) chunks = 1;
        lk = (int)((g.nK + chunks - 1) / chunks);
        if (lk < 8) lk = 8;
        if (lk > g.nK) lk = g.nK;
    }
    p.lk = lk;
    p.chunksK = (g.nK + lk - 1) / lk;
    p.valid = true;
    return p;
}

template <int TI, int TJ, int RJ, bool FUSED>
static inline cudaError_t launch_tiled_inst(const TiledPlan& p, const Grid& g, const double* cl, const double* cv,
                                            const double* r, const double* dinv, const double* pin, double* pout,
                                            double* q, Scalars* sc, double* partials, cudaStream_t st) {
    const size_t smem = TileSmem<TI, TJ>::bytes;
    static bool attr_done[64] = {};   // per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        cudaFuncSetAttribute(k_apply_tiled<TI, TJ, RJ, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done[dev & 63] = true;
    }
    dim3 grid(p.tilesI, p.tilesJ, p.chunksK), block(TI, TJ / RJ, 1);
    k_apply_tiled<TI, TJ, RJ, FUSED><<<grid, block, smem, st>>>(g, p.lk, cl, cv, r, dinv, pin, pout, q, sc, partials);
    return cudaGetLastError();
}

template <bool FUSED>
static inline cudaError_t launch_tiled_dispatch(const TiledPlan& p, const Grid& g, const double* cl, const double* cv,
                                                const double* r, const double* dinv, const double* pin, double* pout,
                                                double* q, Scalars* sc, double* partials, cudaStream_t st) {
#define PFEM_TILE_CASE(TI, TJ, RJ)                                                                          \
    if (p.ti == TI && p.tj == TJ && p.rj == RJ)                                                             \
        return launch_tiled_inst<TI, TJ, RJ, FUSED>(p, g, cl, cv, r, dinv, pin, pout, q, sc, partials, st);
    PFEM_TILE_CASE(32, 16, 2)
    PFEM_TILE_CASE(32, 16, 1)
    PFEM_TILE_CASE(32, 8, 1)
    PFEM_TILE_CASE(32, 8, 2)
    PFEM_TILE_CASE(32, 16, 4)
    PFEM_TILE_CASE(32, 32, 4)
    PFEM_TILE_CASE(64, 8, 2)
    PFEM_TILE_CASE(64, 16, 2)
    PFEM_TILE_CASE(64, 16, 4)
#undef PFEM_TILE_CASE
    return cudaErrorInvalidConfiguration;
}

// fused CG operator step: pin = p_old, pout = p_new
static inline cudaError_t launch_apply_tiled(const TiledPlan& p, const Grid& g, const double* cl, const double* cv,
                                             const double* r, const double* dinv, const double* pin, double* pout,
                                             double* q, Scalars* sc, double* partials, cudaStream_t st) {
    return launch_tiled_dispatch<true>(p, g, cl, cv, r, dinv, pin, pout, q, sc, partials, st);
}
// plain q = M A p
static inline cudaError_t launch_apply_tiled_plain(const TiledPlan& p, const Grid& g, const double* cl, const double* cv,
                                                   const double* dinv, const double* pin, double* q, cudaStream_t st) {
    return launch_tiled_dispatch<false>(p, g, cl, cv, nullptr, dinv, pin, nullptr, q, nullptr, nullptr, st);
}

}  // namespace pfem
