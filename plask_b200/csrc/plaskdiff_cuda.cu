// plaskdiff_cuda.cu — host side of the Diffusion3D path of libplaskfem_cuda.so (C ABI: include/plaskdiff_cuda.h).
// Second translation unit of the library; shares nothing with plaskfem_cuda.cu but the status codes.
#include "../../include/plaskdiff_cuda.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "kernels_diffusion.cuh"

using namespace pdiff;

struct pdiff_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    bool has_mesh = false, has_par = false, has_cur = false;
    int order = 0;
    size_t n0 = 0, n1 = 0, nn = 0, ne = 0;
    int guard = 0, grid = 0;
    std::vector<double> ax0, ax1;
    std::vector<uint8_t> eact_h, nact_h;   // lattice
    Problem P{};
    // device allocations
    double *d_ax0 = nullptr, *d_ax1 = nullptr, *d_h0 = nullptr, *d_h1 = nullptr;
    uint8_t *d_eact_base = nullptr, *d_nact = nullptr;
    double *d_par = nullptr, *d_J = nullptr, *d_modes = nullptr, *d_Ke = nullptr, *d_St = nullptr, *d_Mi = nullptr, *d_vec = nullptr, *d_Fe = nullptr,
           *d_part = nullptr, *d_tmp = nullptr;
    Control* d_ctl = nullptr;
    size_t modes_cap = 0, tmp_cap = 0;
    long long launches = 0;
};

namespace {

int fail(pdiff_ctx* c, int st, const std::string& msg) {
    if (c) c->err = msg;
    return st;
}
#define PD_CUDA(call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? PFEM_ERR_NOMEM : PFEM_ERR_CUDA,             \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                               \
    } while (0)

// 7-point Gauss-Legendre on [0,1] by Newton iteration on P_7 (long double)
void gauss7(long double x[7], long double w[7]) {
    const int n = 7;
    for (int i = 0; i < n; ++i) {
        long double t = cosl(3.14159265358979323846264338327950288L * (i + 0.75L) / (n + 0.5L));
        long double dp = 0;
        for (int it = 0; it < 100; ++it) {
            long double p0 = 1, p1 = t;
            for (int k = 2; k <= n; ++k) { const long double p2 = ((2 * k - 1) * t * p1 - (k - 1) * p0) / k; p0 = p1; p1 = p2; }
            dp = n * (t * p1 - p0) / (t * t - 1);
            const long double dt = p1 / dp;
            t -= dt;
            if (fabsl(dt) < 1e-19L) break;
        }
        x[n - 1 - i] = 0.5L * (t + 1);
        w[n - 1 - i] = 1 / ((1 - t * t) * dp * dp);   // = 0.5 * 2 / ((1 - t^2) P'^2)
    }
}

void hermite(long double t, long double H[4], long double dH[4]) {   // unit interval
    H[0] = 1 - 3 * t * t + 2 * t * t * t;  dH[0] = -6 * t + 6 * t * t;
    H[1] = t * (1 - t) * (1 - t);          dH[1] = (1 - t) * (1 - 3 * t);
    H[2] = 3 * t * t - 2 * t * t * t;      dH[2] = 6 * t - 6 * t * t;
    H[3] = t * t * (t - 1);                dH[3] = t * (3 * t - 2);
}

int init_tables(pdiff_ctx* ctx) {
    long double gx[7], gw[7];
    gauss7(gx, gw);
    static double phi[12][NQ2], w[NQ2], bl[4][NQ2], kx[144], ky[144];
    long double px[12][NQ2], py[12][NQ2];
    for (int qx = 0; qx < 7; ++qx)
        for (int qy = 0; qy < 7; ++qy) {
            const int q = qx * 7 + qy;
            long double Hx[4], dHx[4], Hy[4], dHy[4];
            hermite(gx[qx], Hx, dHx);
            hermite(gx[qy], Hy, dHy);
            w[q] = (double)(gw[qx] * gw[qy]);
            for (int n = 0; n < 4; ++n) {
                const int a = (n >> 1) * 2, b = (n & 1) * 2;   // Hermite value index along axis 0 / axis 1
                const int ia[3] = {a, a, a + 1}, ib[3] = {b, b + 1, b};   // value, d/dy, d/dx
                for (int c = 0; c < 3; ++c) {
                    phi[3 * n + c][q] = (double)(Hx[ia[c]] * Hy[ib[c]]);
                    px[3 * n + c][q] = dHx[ia[c]] * Hy[ib[c]];
                    py[3 * n + c][q] = Hx[ia[c]] * dHy[ib[c]];
                }
                bl[n][q] = (double)(((n >> 1) ? gx[qx] : 1 - gx[qx]) * ((n & 1) ? gx[qy] : 1 - gx[qy]));
            }
        }
    for (int r = 0; r < 12; ++r)
        for (int c = 0; c < 12; ++c) {
            long double sx = 0, sy = 0;
            for (int qx = 0; qx < 7; ++qx)
                for (int qy = 0; qy < 7; ++qy) {
                    const int q = qx * 7 + qy;
                    sx += gw[qx] * gw[qy] * px[r][q] * px[c][q];
                    sy += gw[qx] * gw[qy] * py[r][q] * py[c][q];
                }
            kx[12 * r + c] = (double)sx;
            ky[12 * r + c] = (double)sy;
        }
    PD_CUDA(cudaMemcpyToSymbol(c_phi, phi, sizeof(phi)));
    PD_CUDA(cudaMemcpyToSymbol(c_w, w, sizeof(w)));
    PD_CUDA(cudaMemcpyToSymbol(c_bl, bl, sizeof(bl)));
    PD_CUDA(cudaMemcpyToSymbol(c_kx, kx, sizeof(kx)));
    PD_CUDA(cudaMemcpyToSymbol(c_ky, ky, sizeof(ky)));
    return PFEM_OK;
}

void free_mesh(pdiff_ctx* c) {
    void* ptrs[] = {c->d_ax0, c->d_ax1, c->d_h0, c->d_h1, c->d_eact_base, c->d_nact, c->d_par, c->d_J, c->d_modes, c->d_Ke, c->d_St, c->d_Mi,
                    c->d_vec, c->d_Fe, c->d_part, c->d_tmp};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    c->d_ax0 = c->d_ax1 = c->d_h0 = c->d_h1 = c->d_par = c->d_J = c->d_modes = c->d_Ke = c->d_St = c->d_Mi = c->d_vec = c->d_Fe = c->d_part =
        c->d_tmp = nullptr;
    c->d_eact_base = c->d_nact = nullptr;
    c->modes_cap = c->tmp_cap = 0;
    c->has_mesh = c->has_par = c->has_cur = false;
    c->P = Problem{};
}

// compact element index (ABI) -> lattice slot
inline size_t elem_slot(const pdiff_ctx* c, size_t i0, size_t i1) { return i0 * (size_t)c->P.s0 + i1 * (size_t)c->P.s1; }
inline size_t elem_abi(const pdiff_ctx* c, size_t i0, size_t i1) {
    return c->order == PDIFF_ORDER_01 ? i0 * (c->n1 - 1) + i1 : i1 * (c->n0 - 1) + i0;
}

int upload_elem(pdiff_ctx* ctx, const double* src, size_t ncomp, double* dst) {   // compact [ne][ncomp] -> lattice [NL][ncomp]
    std::vector<double> tmp((size_t)ctx->P.NL * ncomp, 0.);
    for (size_t i0 = 0; i0 + 1 < ctx->n0; ++i0)
        for (size_t i1 = 0; i1 + 1 < ctx->n1; ++i1)
            for (size_t k = 0; k < ncomp; ++k) tmp[elem_slot(ctx, i0, i1) * ncomp + k] = src[elem_abi(ctx, i0, i1) * ncomp + k];
    PD_CUDA(cudaMemcpyAsync(dst, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PD_CUDA(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

int need_tmp(pdiff_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->tmp_cap) return PFEM_OK;
    if (ctx->d_tmp) cudaFree(ctx->d_tmp);
    ctx->d_tmp = nullptr;
    ctx->tmp_cap = 0;
    PD_CUDA(cudaMalloc(&ctx->d_tmp, bytes));
    ctx->tmp_cap = bytes;
    return PFEM_OK;
}

int plain_grid(size_t items) { return (int)std::min<size_t>((items + 127) / 128, 148 * 16); }

// Ke, F, M^-1 at the current U (the hooks; the compute kernel does the same inside its loop)
int assemble_now(pdiff_ctx* ctx, int verbatim) {
    if (!ctx->has_mesh || !ctx->has_par || !ctx->has_cur) return fail(ctx, PFEM_ERR_STATE, "mesh, parameters and current must be set first");
    k_diff_assemble<<<plain_grid(12 * (size_t)ctx->P.NLp), 128, 0, ctx->stream>>>(ctx->P, verbatim, ctx->d_Fe);
    k_diff_gather<<<plain_grid(ctx->P.NL), 128, 0, ctx->stream>>>(ctx->P, ctx->d_Fe);
    ctx->launches += 2;
    PD_CUDA(cudaGetLastError());
    return PFEM_OK;
}

// interleaved ABI vector [node][3] <-> SoA [3][NLp]
void to_soa(const pdiff_ctx* c, const double* v, std::vector<double>& out) {
    out.assign(3 * (size_t)c->P.NLp, 0.);
    for (size_t n = 0; n < c->nn; ++n)
        for (int k = 0; k < 3; ++k) out[(size_t)k * c->P.NLp + n] = v[3 * n + k];
}
void from_soa(const pdiff_ctx* c, const std::vector<double>& in, double* v) {
    for (size_t n = 0; n < c->nn; ++n)
        for (int k = 0; k < 3; ++k) v[3 * n + k] = in[(size_t)k * c->P.NLp + n];
}

}  // namespace

extern "C" {

int pdiff_create(pdiff_ctx** out, int device) {
    if (!out) return PFEM_ERR_BAD_INPUT;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return PFEM_ERR_NO_DEVICE;
    }
    pdiff_ctx* ctx = new pdiff_ctx;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreate(&ctx->stream) != cudaSuccess) {
        delete ctx;
        return PFEM_ERR_CUDA;
    }
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    if (!coop) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return PFEM_ERR_NO_DEVICE;
    }
    int st = init_tables(ctx);
    if (st != PFEM_OK || cudaMalloc(&ctx->d_ctl, sizeof(Control)) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return st != PFEM_OK ? st : PFEM_ERR_NOMEM;
    }
    *out = ctx;
    return PFEM_OK;
}

void pdiff_destroy(pdiff_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    free_mesh(ctx);
    if (ctx->d_ctl) cudaFree(ctx->d_ctl);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* pdiff_last_error(const pdiff_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int pdiff_set_mesh(pdiff_ctx* ctx, size_t n0, size_t n1, const double* ax0, const double* ax1, int order, const uint8_t* elem_active) {
    if (!ctx) return PFEM_ERR_BAD_INPUT;
    if (n0 < 2 || n1 < 2 || !ax0 || !ax1 || (order != PDIFF_ORDER_01 && order != PDIFF_ORDER_10))
        return fail(ctx, PFEM_ERR_BAD_INPUT, "mesh needs at least 2 x 2 nodes and a valid order");
    if (n0 * n1 > (size_t)150000000) return fail(ctx, PFEM_ERR_BAD_INPUT, "lateral mesh too large (32-bit lattice indices)");
    for (size_t i = 1; i < n0; ++i)
        if (!(ax0[i] > ax0[i - 1])) return fail(ctx, PFEM_ERR_BAD_INPUT, "axis 0 is not strictly increasing");
    for (size_t i = 1; i < n1; ++i)
        if (!(ax1[i] > ax1[i - 1])) return fail(ctx, PFEM_ERR_BAD_INPUT, "axis 1 is not strictly increasing");
    PD_CUDA(cudaSetDevice(ctx->device));
    free_mesh(ctx);
    ctx->n0 = n0; ctx->n1 = n1; ctx->nn = n0 * n1; ctx->ne = (n0 - 1) * (n1 - 1); ctx->order = order;
    ctx->ax0.assign(ax0, ax0 + n0);
    ctx->ax1.assign(ax1, ax1 + n1);
    Problem& P = ctx->P;
    P.n0 = (int)n0; P.n1 = (int)n1;
    P.s0 = order == PDIFF_ORDER_01 ? (int)n1 : 1;
    P.s1 = order == PDIFF_ORDER_01 ? 1 : (int)n0;
    P.NL = (int)ctx->nn;
    P.NLp = (P.NL + 31) / 32 * 32;
    ctx->guard = std::max(P.s0, P.s1) + 2;

    std::vector<double> h0(n0 - 1), h1(n1 - 1);
    for (size_t i = 0; i + 1 < n0; ++i) h0[i] = ax0[i + 1] - ax0[i];
    for (size_t i = 0; i + 1 < n1; ++i) h1[i] = ax1[i + 1] - ax1[i];
    ctx->eact_h.assign(ctx->nn, 0);
    ctx->nact_h.assign(ctx->nn, 0);
    for (size_t i0 = 0; i0 + 1 < n0; ++i0)
        for (size_t i1 = 0; i1 + 1 < n1; ++i1)
            if (!elem_active || elem_active[elem_abi(ctx, i0, i1)]) {
                const size_t e = elem_slot(ctx, i0, i1);
                ctx->eact_h[e] = 1;
                ctx->nact_h[e] = ctx->nact_h[e + P.s0] = ctx->nact_h[e + P.s1] = ctx->nact_h[e + P.s0 + P.s1] = 1;
            }

    const size_t NLp = P.NLp, g = ctx->guard;
    PD_CUDA(cudaMalloc(&ctx->d_ax0, n0 * sizeof(double)));
    PD_CUDA(cudaMalloc(&ctx->d_ax1, n1 * sizeof(double)));
    PD_CUDA(cudaMalloc(&ctx->d_h0, (n0 - 1) * sizeof(double)));
    PD_CUDA(cudaMalloc(&ctx->d_h1, (n1 - 1) * sizeof(double)));
    PD_CUDA(cudaMalloc(&ctx->d_eact_base, NLp + 2 * g));
    PD_CUDA(cudaMalloc(&ctx->d_nact, NLp));
    PD_CUDA(cudaMalloc(&ctx->d_par, 4 * NLp * sizeof(double)));
    PD_CUDA(cudaMalloc(&ctx->d_J, NLp * sizeof(double)));
    PD_CUDA(cudaMalloc(&ctx->d_Ke, 144 * NLp * sizeof(double)));
    PD_CUDA(cudaMalloc(&ctx->d_St, 81 * NLp * sizeof(double)));
    PD_CUDA(cudaMalloc(&ctx->d_Mi, 6 * NLp * sizeof(double)));
    PD_CUDA(cudaMalloc(&ctx->d_vec, 7 * 3 * NLp * sizeof(double)));
    PD_CUDA(cudaMalloc(&ctx->d_Fe, 12 * NLp * sizeof(double)));
    PD_CUDA(cudaMemsetAsync(ctx->d_eact_base, 0, NLp + 2 * g, ctx->stream));
    PD_CUDA(cudaMemsetAsync(ctx->d_nact, 0, NLp, ctx->stream));
    PD_CUDA(cudaMemsetAsync(ctx->d_par, 0, 4 * NLp * sizeof(double), ctx->stream));
    PD_CUDA(cudaMemsetAsync(ctx->d_J, 0, NLp * sizeof(double), ctx->stream));
    PD_CUDA(cudaMemsetAsync(ctx->d_Ke, 0, 144 * NLp * sizeof(double), ctx->stream));
    PD_CUDA(cudaMemsetAsync(ctx->d_St, 0, 81 * NLp * sizeof(double), ctx->stream));
    PD_CUDA(cudaMemsetAsync(ctx->d_Mi, 0, 6 * NLp * sizeof(double), ctx->stream));
    PD_CUDA(cudaMemsetAsync(ctx->d_vec, 0, 7 * 3 * NLp * sizeof(double), ctx->stream));
    PD_CUDA(cudaMemsetAsync(ctx->d_Fe, 0, 12 * NLp * sizeof(double), ctx->stream));
    PD_CUDA(cudaMemcpyAsync(ctx->d_ax0, ax0, n0 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PD_CUDA(cudaMemcpyAsync(ctx->d_ax1, ax1, n1 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PD_CUDA(cudaMemcpyAsync(ctx->d_h0, h0.data(), h0.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PD_CUDA(cudaMemcpyAsync(ctx->d_h1, h1.data(), h1.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PD_CUDA(cudaMemcpyAsync(ctx->d_eact_base + g, ctx->eact_h.data(), ctx->nn, cudaMemcpyHostToDevice, ctx->stream));
    PD_CUDA(cudaMemcpyAsync(ctx->d_nact, ctx->nact_h.data(), ctx->nn, cudaMemcpyHostToDevice, ctx->stream));
    PD_CUDA(cudaStreamSynchronize(ctx->stream));

    P.h0 = ctx->d_h0; P.h1 = ctx->d_h1;
    P.eact = ctx->d_eact_base + g;
    P.nact = ctx->d_nact;
    P.A = ctx->d_par; P.B = ctx->d_par + NLp; P.C = ctx->d_par + 2 * NLp; P.D = ctx->d_par + 3 * NLp;
    P.J = ctx->d_J;
    P.nmodes = 0; P.P = P.G = P.dG = nullptr;
    P.Ke = ctx->d_Ke; P.St = ctx->d_St; P.Mi = ctx->d_Mi;
    double* v = ctx->d_vec;
    P.F = v; P.U = v + 3 * NLp; P.r = v + 6 * NLp; P.z = v + 9 * NLp; P.q = v + 12 * NLp; P.p0 = v + 15 * NLp; P.p1 = v + 18 * NLp;

    // persistent grid: as many CTAs as are co-resident (cooperative launch), but no more than the node loop can use
    int per_sm = 0, sms = 0;
    PD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_diff_compute, 128, 0));
    PD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    if (per_sm < 1) return fail(ctx, PFEM_ERR_CUDA, "k_diff_compute does not fit on an SM");
    ctx->grid = (int)std::min<size_t>((size_t)per_sm * sms, (ctx->nn + 127) / 128);
    PD_CUDA(cudaMalloc(&ctx->d_part, 2 * 4 * (size_t)ctx->grid * sizeof(double)));
    P.part = ctx->d_part;
    ctx->has_mesh = true;
    return PFEM_OK;
}

int pdiff_set_parameters(pdiff_ctx* ctx, const double* A, const double* B, const double* C, const double* D) {
    if (!ctx) return PFEM_ERR_BAD_INPUT;
    if (!ctx->has_mesh) return fail(ctx, PFEM_ERR_STATE, "pdiff_set_mesh first");
    if (!A || !B || !C || !D) return fail(ctx, PFEM_ERR_BAD_INPUT, "null parameter array");
    PD_CUDA(cudaSetDevice(ctx->device));
    const double* src[4] = {A, B, C, D};
    for (int k = 0; k < 4; ++k) {
        int st = upload_elem(ctx, src[k], 1, ctx->d_par + (size_t)k * ctx->P.NLp);
        if (st != PFEM_OK) return st;
    }
    ctx->has_par = true;
    return PFEM_OK;
}

int pdiff_set_current(pdiff_ctx* ctx, const double* J) {
    if (!ctx) return PFEM_ERR_BAD_INPUT;
    if (!ctx->has_mesh) return fail(ctx, PFEM_ERR_STATE, "pdiff_set_mesh first");
    if (!J) return fail(ctx, PFEM_ERR_BAD_INPUT, "null current array");
    PD_CUDA(cudaSetDevice(ctx->device));
    PD_CUDA(cudaMemcpyAsync(ctx->d_J, J, ctx->nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PD_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->has_cur = true;
    return PFEM_OK;
}

int pdiff_set_modes(pdiff_ctx* ctx, size_t nmodes, const double* Pm, const double* G, const double* dG) {
    if (!ctx) return PFEM_ERR_BAD_INPUT;
    if (!ctx->has_mesh) return fail(ctx, PFEM_ERR_STATE, "pdiff_set_mesh first");
    if (nmodes == 0) {
        ctx->P.nmodes = 0;
        return PFEM_OK;
    }
    if (!Pm || !G || !dG) return fail(ctx, PFEM_ERR_BAD_INPUT, "null mode array");
    PD_CUDA(cudaSetDevice(ctx->device));
    const size_t NL = ctx->P.NL, per = 2 * NL, need = 3 * nmodes * per * sizeof(double);
    if (need > ctx->modes_cap) {
        if (ctx->d_modes) cudaFree(ctx->d_modes);
        ctx->d_modes = nullptr;
        ctx->modes_cap = 0;
        PD_CUDA(cudaMalloc(&ctx->d_modes, need));
        ctx->modes_cap = need;
    }
    double* dP = ctx->d_modes;
    double* dGm = dP + nmodes * per;
    double* ddG = dGm + nmodes * per;
    PD_CUDA(cudaMemcpyAsync(dP, Pm, nmodes * per * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    for (size_t m = 0; m < nmodes; ++m) {
        int st = upload_elem(ctx, G + m * 2 * ctx->ne, 2, dGm + m * per);
        if (st == PFEM_OK) st = upload_elem(ctx, dG + m * 2 * ctx->ne, 2, ddG + m * per);
        if (st != PFEM_OK) return st;
    }
    PD_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->P.nmodes = (int)nmodes; ctx->P.P = dP; ctx->P.G = dGm; ctx->P.dG = ddG;
    return PFEM_OK;
}

int pdiff_set_concentration(pdiff_ctx* ctx, const double* U) {
    if (!ctx) return PFEM_ERR_BAD_INPUT;
    if (!ctx->has_mesh) return fail(ctx, PFEM_ERR_STATE, "pdiff_set_mesh first");
    PD_CUDA(cudaSetDevice(ctx->device));
    std::vector<double> soa;
    if (U) {
        to_soa(ctx, U, soa);
        for (size_t n = 0; n < ctx->nn; ++n)   // nodes outside the masked mesh are not unknowns
            if (!ctx->nact_h[n]) soa[n] = soa[ctx->P.NLp + n] = soa[2 * (size_t)ctx->P.NLp + n] = 0.;
    } else {
        soa.assign(3 * (size_t)ctx->P.NLp, 0.);
    }
    PD_CUDA(cudaMemcpyAsync(ctx->P.U, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PD_CUDA(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

int pdiff_get_concentration(pdiff_ctx* ctx, double* U) {
    if (!ctx) return PFEM_ERR_BAD_INPUT;
    if (!ctx->has_mesh) return fail(ctx, PFEM_ERR_STATE, "pdiff_set_mesh first");
    if (!U) return fail(ctx, PFEM_ERR_BAD_INPUT, "null output");
    PD_CUDA(cudaSetDevice(ctx->device));
    std::vector<double> soa(3 * (size_t)ctx->P.NLp);
    PD_CUDA(cudaMemcpyAsync(soa.data(), ctx->P.U, soa.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PD_CUDA(cudaStreamSynchronize(ctx->stream));
    from_soa(ctx, soa, U);
    return PFEM_OK;
}

void pdiff_default_opts(pdiff_opts* o) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->loops = 0;
    o->maxerr = 0.05;      // Diffusion3DSolver::maxerr default (diffusion3d.cpp:23)
    o->maxit = 20000;
    o->lin_tol = 1e-12;
    o->verbatim = 1;
    o->loop_cap = 10000;
}

int pdiff_compute(pdiff_ctx* ctx, const pdiff_opts* opts, pdiff_stats* stats) {
    if (!ctx) return PFEM_ERR_BAD_INPUT;
    if (!ctx->has_mesh || !ctx->has_par || !ctx->has_cur) return fail(ctx, PFEM_ERR_STATE, "mesh, parameters and current must be set first");
    pdiff_opts o;
    if (opts) o = *opts; else pdiff_default_opts(&o);
    if (o.loops < 0 || o.maxit < 1 || !(o.lin_tol > 0.)) return fail(ctx, PFEM_ERR_BAD_INPUT, "bad options");
    PD_CUDA(cudaSetDevice(ctx->device));
    Control h{};
    h.loops = o.loops; h.maxerr = o.maxerr; h.maxit = o.maxit; h.lin_tol = o.lin_tol; h.verbatim = o.verbatim;
    h.loop_cap = o.loop_cap > 0 ? o.loop_cap : 10000;
    PD_CUDA(cudaMemcpyAsync(ctx->d_ctl, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    cudaEvent_t e0, e1;
    PD_CUDA(cudaEventCreate(&e0));
    PD_CUDA(cudaEventCreate(&e1));
    PD_CUDA(cudaEventRecord(e0, ctx->stream));
    Problem P = ctx->P;
    Control* ctl = ctx->d_ctl;
    double* Fe = ctx->d_Fe;
    void* args[] = {&P, &ctl, &Fe};
    cudaError_t le = cudaLaunchCooperativeKernel((const void*)k_diff_compute, dim3(ctx->grid), dim3(128), args, 0, ctx->stream);
    ctx->launches += 1;
    if (le != cudaSuccess) {
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return fail(ctx, PFEM_ERR_CUDA, std::string("cooperative launch of k_diff_compute: ") + cudaGetErrorString(le));
    }
    PD_CUDA(cudaEventRecord(e1, ctx->stream));
    PD_CUDA(cudaMemcpyAsync(&h, ctx->d_ctl, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    PD_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->loops = h.loops_done; stats->converged = h.converged; stats->err = h.err; stats->lin_iters = h.lin_iters;
        stats->last_iters = h.last_iters; stats->lin_relres = h.lin_relres; stats->lin_relres_precond = h.lin_relres_precond; stats->t_solve_ms = ms; stats->kernel_launches = 1;
        for (int i = 0; i < 64 && i < h.loops_done; ++i) stats->err_log[i] = h.err_log[i];
    }
    if (h.status == -5) return fail(ctx, PFEM_ERR_NOT_SPD, "p.Kp <= 0 in the conjugate gradient: the linearised diffusion matrix is not positive definite");
    if (h.status == -7) return fail(ctx, PFEM_ERR_NAN, "non-finite value in the diffusion loop (|F| = 0 or bad input)");
    if (h.status == 2) ctx->err = "loop cap reached without err < maxerr";
    return h.status >= 1 ? PFEM_NOT_CONVERGED : PFEM_OK;
}

int pdiff_interpolate(pdiff_ctx* ctx, size_t npts, const double* x, const double* y, int method, double* out) {
    if (!ctx) return PFEM_ERR_BAD_INPUT;
    if (!ctx->has_mesh) return fail(ctx, PFEM_ERR_STATE, "pdiff_set_mesh first");
    if (method != PDIFF_INTERP_SPLINE && method != PDIFF_INTERP_LINEAR) return fail(ctx, PFEM_ERR_BAD_INPUT, "unknown interpolation method");
    if (npts == 0) return PFEM_OK;
    if (!x || !y || !out) return fail(ctx, PFEM_ERR_BAD_INPUT, "null point array");
    PD_CUDA(cudaSetDevice(ctx->device));
    int st = need_tmp(ctx, 3 * npts * sizeof(double));
    if (st != PFEM_OK) return st;
    double *dx = ctx->d_tmp, *dy = dx + npts, *dout = dy + npts;
    PD_CUDA(cudaMemcpyAsync(dx, x, npts * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PD_CUDA(cudaMemcpyAsync(dy, y, npts * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_diff_interp<<<(unsigned)((npts + 127) / 128), 128, 0, ctx->stream>>>(ctx->P, ctx->d_ax0, ctx->d_ax1, (int)npts, dx, dy, method, dout);
    ctx->launches += 1;
    PD_CUDA(cudaGetLastError());
    PD_CUDA(cudaMemcpyAsync(out, dout, npts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PD_CUDA(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

int pdiff_get_element_matrices(pdiff_ctx* ctx, int verbatim, double* K, double* F) {
    if (!ctx) return PFEM_ERR_BAD_INPUT;
    int st = assemble_now(ctx, verbatim);
    if (st != PFEM_OK) return st;
    const size_t NLp = ctx->P.NLp;
    std::vector<double> hK(144 * NLp), hF(12 * NLp);
    PD_CUDA(cudaMemcpyAsync(hK.data(), ctx->d_Ke, hK.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PD_CUDA(cudaMemcpyAsync(hF.data(), ctx->d_Fe, hF.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PD_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t i0 = 0; i0 + 1 < ctx->n0; ++i0)
        for (size_t i1 = 0; i1 + 1 < ctx->n1; ++i1) {
            const size_t e = elem_slot(ctx, i0, i1), a = elem_abi(ctx, i0, i1);
            const bool act = ctx->eact_h[e];
            if (K)
                for (int k = 0; k < 144; ++k) K[a * 144 + k] = act ? hK[k * NLp + e] : 0.;
            if (F)
                for (int k = 0; k < 12; ++k) F[a * 12 + k] = act ? hF[k * NLp + e] : 0.;
        }
    return PFEM_OK;
}

int pdiff_apply(pdiff_ctx* ctx, int verbatim, const double* v, double* y) {
    if (!ctx) return PFEM_ERR_BAD_INPUT;
    if (!v || !y) return fail(ctx, PFEM_ERR_BAD_INPUT, "null vector");
    int st = assemble_now(ctx, verbatim);
    if (st != PFEM_OK) return st;
    const size_t NLp = ctx->P.NLp;
    st = need_tmp(ctx, 6 * NLp * sizeof(double));
    if (st != PFEM_OK) return st;
    std::vector<double> soa;
    to_soa(ctx, v, soa);
    double *dv = ctx->d_tmp, *dy = dv + 3 * NLp;
    PD_CUDA(cudaMemcpyAsync(dv, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_diff_apply<<<plain_grid(ctx->P.NL), 128, 0, ctx->stream>>>(ctx->P, dv, dy);
    ctx->launches += 1;
    PD_CUDA(cudaGetLastError());
    PD_CUDA(cudaMemcpyAsync(soa.data(), dy, soa.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PD_CUDA(cudaStreamSynchronize(ctx->stream));
    from_soa(ctx, soa, y);
    return PFEM_OK;
}

int pdiff_get_rhs(pdiff_ctx* ctx, int verbatim, double* F) {
    if (!ctx) return PFEM_ERR_BAD_INPUT;
    if (!F) return fail(ctx, PFEM_ERR_BAD_INPUT, "null output");
    int st = assemble_now(ctx, verbatim);
    if (st != PFEM_OK) return st;
    std::vector<double> soa(3 * (size_t)ctx->P.NLp);
    PD_CUDA(cudaMemcpyAsync(soa.data(), ctx->P.F, soa.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PD_CUDA(cudaStreamSynchronize(ctx->stream));
    from_soa(ctx, soa, F);
    return PFEM_OK;
}

}  // extern "C"
