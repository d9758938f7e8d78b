// kernels_tma.cuh — production operator kernel: TMA-staged, mbarrier-pipelined version of the
// fused CG operator step (same mathematics as kernels_tiled.cuh, see the header there).
//
//   p_new = D^-1 r + beta p_old ;  q = M A p_new ;  p.q -> alpha
//
// Data movement (sm_100a): one elected thread issues cp.async.bulk.tensor.3d (TMA) box loads of
// the next node plane of r, D^-1, p_old and of the next element layer of c_lat, c_vert — tile plus
// a one-node halo, out-of-range parts zero-filled by the TMA unit — into a ring of NS shared-memory
// stages, each completing on its own mbarrier (complete_tx).  All threads then (1) fuse the p-update
// and the element conductances k_I,k_J,k_K/36 into compact shared-memory planes and (2) gather the
// 27-point operator for their nodes.  Loads for step s+NS-1 are in flight while step s computes, so
// HBM latency is hidden without spending registers.
//
// TMA constraint met here (measured with tools/tma_probe.cu on a B200): the start address of a box
// must be 16-byte aligned, i.e. with 8-byte elements the innermost box coordinate must be EVEN —
// an odd one faults with "illegal instruction".  Tiles start at multiples of TI (even), so the box
// carries a TWO-node halo on the low-I side and on the high-I side (HX = 2, box width TI + 4), of
// which one column each side is used; rows (J) and planes (K) have no such constraint.
#pragma once
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pfem_internal.cuh"

namespace pfem {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA: global (tensor map, element coordinates c0 fastest) -> shared, completion on an mbarrier
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

template <int TI, int TJ>
struct TmaTile {
    static constexpr int HX = 2;                        // halo columns on each side of the tile (see header)
    static constexpr int PW = TI + 2 * HX, PH = TJ + 2;
    static constexpr int BOX = PW * PH;                 // doubles moved per box
    static constexpr int BOXP = (BOX + 15) / 16 * 16;   // 128-byte aligned box slot
    static constexpr size_t smem_bytes(int ns, bool fused) {
        return 128 /*align slack*/ + sizeof(double) * ((size_t)ns * (fused ? 5 : 3) * BOXP + 5 * BOXP + 64) + 16 * 8 + 16;
    }
};

// FUSED: node boxes p_old, r, D^-1; plain: node box p only.
template <int TI, int TJ, int RJ, int NS, bool FUSED>
__global__ void __launch_bounds__(TI*(TJ / RJ))
k_apply_tma(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_r,
            const __grid_constant__ CUtensorMap tm_d, const __grid_constant__ CUtensorMap tm_cl,
            const __grid_constant__ CUtensorMap tm_cv, const Grid g, const int lk, const double* __restrict__ dinv,
            double* __restrict__ pout, double* __restrict__ q, Scalars* sc, double* partials) {
    typedef TmaTile<TI, TJ> T;
    constexpr int NT = TI * (TJ / RJ);
    constexpr int NBN = FUSED ? 3 : 1;  // node boxes per stage
    constexpr int NB = NBN + 2;         // + c_lat, c_vert
    constexpr int PW = T::PW, BOX = T::BOX, BOXP = T::BOXP, HX = T::HX;
    extern __shared__ unsigned char smem_dyn[];
    double* base = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
    double* sRaw = base;                          // [NS][NB][BOXP]
    double* sP = sRaw + (size_t)NS * NB * BOXP;   // [2][BOXP]  p_new planes (ping-pong)
    double* sC = sP + 2 * BOXP;                   // [3][BOXP]  kI,kJ,kK of the current layer
    double* sRed = sC + 3 * BOXP;                 // 64
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 64);  // [NS]
    int* sh_flag = reinterpret_cast<int*>(bars + 16);

    if (FUSED && sc->done) return;
    const double beta = FUSED ? sc->beta : 0.;

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = tx + TI * ty;
    const int i0 = blockIdx.x * TI, j0 = blockIdx.y * TJ;
    const int k0 = blockIdx.z * lk;
    const int k1 = min(k0 + lk, g.nK);
    const int nsteps = k1 - k0 + 1;  // element layers k0-1 .. k1-1

    // item t: node plane k0-1+t and (t > 0) element layer k0-2+t, into stage t % NS
    auto issue = [&](int t) {
        const int st = t % NS;
        double* dst = sRaw + (size_t)st * NB * BOXP;
        uint64_t* bar = &bars[st];
        mbar_expect_tx(bar, (uint32_t)((NBN + (t > 0 ? 2 : 0)) * BOX * sizeof(double)));
        const int P = k0 - 1 + t;
        tma_load_3d(dst, &tm_p, bar, i0 - HX, j0 - 1, P);
        if (FUSED) {
            tma_load_3d(dst + BOXP, &tm_r, bar, i0 - HX, j0 - 1, P);
            tma_load_3d(dst + 2 * BOXP, &tm_d, bar, i0 - HX, j0 - 1, P);
        }
        if (t > 0) {
            tma_load_3d(dst + NBN * BOXP, &tm_cl, bar, i0 - HX, j0 - 1, P - 1);
            tma_load_3d(dst + (NBN + 1) * BOXP, &tm_cv, bar, i0 - HX, j0 - 1, P - 1);
        }
    };
    // p_new of node plane P from stage st into plane buffer `buf`; owned nodes are stored to pout
    auto make_plane = [&](int st, int buf, int P) {
        const double* raw = sRaw + (size_t)st * NB * BOXP;
        double* dst = sP + buf * BOXP;
        const bool own_k = (P >= k0 && P < k1);
        for (int m = tid; m < BOX; m += NT) {
            double v;
            if (FUSED) {
                v = raw[2 * BOXP + m] * raw[BOXP + m] + beta * raw[m];
                if (own_k) {
                    const int jj = m / PW, ii = m - jj * PW;
                    const int i = i0 + ii - HX, j = j0 + jj - 1;
                    if (ii >= HX && ii < HX + TI && jj >= 1 && jj <= TJ && i < g.nI && j < g.nJ) pout[i + g.sJ * j + g.sK * P] = v;
                }
            } else {
                v = raw[m];
            }
            dst[m] = v;
        }
    };
    // conductances of element layer ek (already zero where the TMA box was out of range)
    auto make_layer = [&](int st, int ek) {
        const double* raw = sRaw + (size_t)st * NB * BOXP + NBN * BOXP;
        const int ekc = min(max(ek, -1), g.nK - 1);
        const double hk = g.hK[ekc], rk = g.rK[ekc];
        const double s = 1e-6 / 36.;
        for (int m = tid; m < BOX; m += NT) {
            const int jj = m / PW, ii = m - jj * PW;
            const int ei = min(max(i0 + ii - HX, -1), g.nI - 1), ej = min(j0 + jj - 1, g.nJ - 1);  // guards of hI/hJ: -1 .. n-1
            const double a = raw[m], b = raw[BOXP + m];
            const double hi = g.hI[ei], hj = g.hJ[ej];
            const double cI = (g.vdim == 0 ? b : a) * s, cJ = (g.vdim == 1 ? b : a) * s, cK = (g.vdim == 2 ? b : a) * s;
            sC[m] = cI * (hj * hk) * g.rI[ei];
            sC[BOXP + m] = cJ * (hi * hk) * g.rJ[ej];
            sC[2 * BOXP + m] = cK * (hi * hj) * rk;
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

    double carry[RJ];
#pragma unroll
    for (int rr = 0; rr < RJ; ++rr) carry[rr] = 0.;
    double dot = 0.;

    mbar_wait(&bars[0], 0);
    make_plane(0, 0, k0 - 1);
    __syncthreads();
    if (tid == 0 && NS <= nsteps) issue(NS);
    int cur = 0;
    for (int s = 0; s < nsteps; ++s) {
        const int t = s + 1, L = k0 - 1 + s, st = t % NS;
        mbar_wait(&bars[st], (uint32_t)((t / NS) & 1));
        make_plane(st, cur ^ 1, L + 1);
        make_layer(st, L);
        __syncthreads();
        if (tid == 0 && t + NS <= nsteps) issue(t + NS);
        const double* Pa = sP + cur * BOXP;        // plane L   (below)
        const double* Pb = sP + (cur ^ 1) * BOXP;  // plane L+1 (above)
#pragma unroll
        for (int rr = 0; rr < RJ; ++rr) {
            const int jl = ty * RJ + rr;
            const int i = i0 + tx, j = j0 + jl;
            const bool store = (L >= k0) && i < g.nI && j < g.nJ;
            const idx_t n = i + g.sJ * j + g.sK * L;
            const double dn = store ? dinv[n] : 0.;  // issued early; only the mask is needed
            double a[3][3], b[3][3];
#pragma unroll
            for (int dj = 0; dj < 3; ++dj)
#pragma unroll
                for (int di = 0; di < 3; ++di) {
                    a[dj][di] = Pa[(jl + dj) * PW + tx + di + (HX - 1)];
                    b[dj][di] = Pb[(jl + dj) * PW + tx + di + (HX - 1)];
                }
            double lo = 0., hi = 0., cc = 0.;
#pragma unroll
            for (int sj = 0; sj < 2; ++sj)
#pragma unroll
                for (int si = 0; si < 2; ++si) {
                    const int ce = (jl + sj) * PW + tx + si + (HX - 1);
                    const double kI = sC[ce], kJ = sC[BOXP + ce], kK = sC[2 * BOXP + ce];
                    const int on = sj ? 2 : 0;
                    const int cn = si ? 2 : 0;
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
            if (store) {
                const double qv = (dn == 0.) ? 0. : carry[rr] + lo - cc;
                q[n] = qv;
                dot += a[1][1] * qv;
            }
            carry[rr] = hi + cc;
        }
        __syncthreads();
        cur ^= 1;
    }
    if (!FUSED) return;
    double v[2] = {dot, 0.};
    if (grid_reduce<2, false>(v, partials, &sc->ticket[0], sRed, sh_flag)) {
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

struct TmaPlan {
    bool valid;
    int ti, tj, rj, ns;
    int lk, tilesI, tilesJ, chunksK;
    CUtensorMap m_p[2], m_r, m_d, m_cl, m_cv;  // p ping-pong buffers, residual, D^-1, conductivities
    char why[160];
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 3-D FP64 tensor map over a lattice array: extents (d0,d1,d2) elements, row pitch sJ, plane pitch sK
static inline bool make_lattice_map(CUtensorMap* m, double* basep, idx_t d0, idx_t d1, idx_t d2, idx_t sJ, idx_t sK, int bw,
                                    int bh) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)sJ * 8, (cuuint64_t)sK * 8};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, basep, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static inline TmaPlan make_tma_plan(const Grid& g, int sm_count, double* p0, double* p1, double* r, double* dinv, double* cl,
                                    double* cv) {
    TmaPlan p;
    memset(&p, 0, sizeof(p));
    p.ti = 32; p.tj = 16; p.rj = 2; p.ns = 3;
    int lk = 0;
    const char* env = getenv("PFEM_TILE");  // "ti,tj,rj,lk,ns" for tuning runs
    if (env) {
        int a, b, c, d, e;
        int got = sscanf(env, "%d,%d,%d,%d,%d", &a, &b, &c, &d, &e);
        if (got >= 4) { p.ti = a; p.tj = b; p.rj = c; lk = d; }
        if (got >= 5) p.ns = e;
    }
    p.tilesI = (g.nI + p.ti - 1) / p.ti;
    p.tilesJ = (g.nJ + p.tj - 1) / p.tj;
    if (lk <= 0) {
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
    if ((g.sJ * 8) % 16 != 0 || (g.sK * 8) % 16 != 0) { snprintf(p.why, sizeof p.why, "row pitch is not a multiple of 16 bytes"); return p; }
    const int bw = p.ti + 4, bh = p.tj + 2;   // TmaTile::PW x PH
    bool ok = make_lattice_map(&p.m_p[0], p0, g.nI, g.nJ, g.nK, g.sJ, g.sK, bw, bh) &&
              make_lattice_map(&p.m_p[1], p1, g.nI, g.nJ, g.nK, g.sJ, g.sK, bw, bh) &&
              make_lattice_map(&p.m_r, r, g.nI, g.nJ, g.nK, g.sJ, g.sK, bw, bh) &&
              make_lattice_map(&p.m_d, dinv, g.nI, g.nJ, g.nK, g.sJ, g.sK, bw, bh) &&
              make_lattice_map(&p.m_cl, cl, g.nI - 1, g.nJ - 1, g.nK - 1, g.sJ, g.sK, bw, bh) &&
              make_lattice_map(&p.m_cv, cv, g.nI - 1, g.nJ - 1, g.nK - 1, g.sJ, g.sK, bw, bh);
    if (!ok) { snprintf(p.why, sizeof p.why, "cuTensorMapEncodeTiled failed or is unavailable"); return p; }
    p.valid = true;
    return p;
}

template <int TI, int TJ, int RJ, int NS, bool FUSED>
static inline cudaError_t launch_tma_inst(const TmaPlan& p, const Grid& g, const CUtensorMap& mp, const double* dinv,
                                          double* pout, double* q, Scalars* sc, double* partials, cudaStream_t st) {
    const size_t smem = TmaTile<TI, TJ>::smem_bytes(NS, FUSED);
    static bool attr_done[64] = {};   // per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(k_apply_tma<TI, TJ, RJ, NS, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done[dev & 63] = true;
    }
    dim3 grid(p.tilesI, p.tilesJ, p.chunksK), block(TI, TJ / RJ, 1);
    k_apply_tma<TI, TJ, RJ, NS, FUSED><<<grid, block, smem, st>>>(mp, p.m_r, p.m_d, p.m_cl, p.m_cv, g, p.lk, dinv, pout, q, sc, partials);
    return cudaGetLastError();
}

template <bool FUSED>
static inline cudaError_t launch_tma_dispatch(const TmaPlan& p, const Grid& g, const CUtensorMap& mp, const double* dinv,
                                              double* pout, double* q, Scalars* sc, double* partials, cudaStream_t st) {
#define PFEM_TMA_CASE(TI, TJ, RJ, NS) \
    if (p.ti == TI && p.tj == TJ && p.rj == RJ && p.ns == NS) return launch_tma_inst<TI, TJ, RJ, NS, FUSED>(p, g, mp, dinv, pout, q, sc, partials, st);
    PFEM_TMA_CASE(32, 16, 2, 3)
    PFEM_TMA_CASE(32, 16, 2, 2)
    PFEM_TMA_CASE(32, 16, 1, 3)
    PFEM_TMA_CASE(32, 16, 1, 2)
    PFEM_TMA_CASE(32, 16, 4, 3)
    PFEM_TMA_CASE(32, 32, 2, 2)
    PFEM_TMA_CASE(32, 32, 4, 2)
    PFEM_TMA_CASE(64, 16, 2, 2)
    PFEM_TMA_CASE(64, 16, 4, 2)
    PFEM_TMA_CASE(64, 8, 2, 3)
    PFEM_TMA_CASE(64, 8, 1, 3)
    PFEM_TMA_CASE(32, 8, 1, 4)
    PFEM_TMA_CASE(32, 8, 2, 4)
#undef PFEM_TMA_CASE
    return cudaErrorInvalidConfiguration;
}

}  // namespace pfem
