// kernels_surface.cuh — boundary conditions of the 2nd / 3rd kind and radiation.
//
// The reference adds them element by element inside setMatrix (setBoundaries, therm3d.cpp:140-168, used at
// :242-268): heat flux and radiation only touch the load vector, convection adds a face mass matrix to K as well.
// They live on mesh faces, i.e. on O(N^(2/3)) nodes, so they stay OUT of the volume kernel: the host flattens them
// (pfem_set_boundary, plaskfem_cuda.cu) into a row list `Surf`, and
//   * k_surf_rad   evaluates the radiation loads from the temperatures the loop starts with (therm3d.cpp:262-267),
//   * k_surf_rhs   builds the effective load  f + load - S in  for the rows (prepare phase, lifted rhs / r0),
//   * k_surf_diag  adds diag(S) to the Jacobi preconditioner,
//   * k_surf_add   q += S p (plain operator, pfem_apply),
//   * k_surf_iter  runs after k_fpcg in every PCG iteration when S != 0: q' += S p' on the rows, corrects the three
//                  dot products that involve q' (p'.q', q'.z, q'.D^-1 q') and then does the alpha / beta-prediction step
//                  that k_fpcg's last CTA does otherwise.  The convergence test does not involve q', k_fpcg keeps it.
#pragma once
#include "pfem_internal.cuh"

namespace pfem {

__global__ void k_surf_rad(const Surf s, const double* __restrict__ T) {
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < s.nrows; row += gridDim.x * blockDim.x) {
        double v = 0.;
        for (int k = s.radptr[row]; k < s.radptr[row + 1]; ++k) {
            double t = T[s.rad_src[k]];
            t = t * t;
            v -= s.rad_coef[k] * (t * t - s.rad_amb4[k]);   // - 0.25e-12 * area * eps * SB * (T^4 - Ta^4), therm3d.cpp:266
        }
        s.radv[row] = v;
    }
}

// fS[node] = f[node] + load[row] - sum_c S[row,c] in[c]   on the surface rows (fS is a copy of f elsewhere)
__global__ void k_surf_rhs(const Surf s, const double* __restrict__ f, const double* __restrict__ in, double* __restrict__ fS) {
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < s.nrows; row += gridDim.x * blockDim.x) {
        const idx_t n = s.node[row];
        double v = f[n] + (s.lconst[row] + s.radv[row]);
        double sq = 0.;
        for (int k = s.kptr[row]; k < s.kptr[row + 1]; ++k) sq = fma(s.kval[k], in[s.kcol[k]], sq);
        fS[n] = v - sq;
    }
}

__global__ void k_surf_diag(const Surf s, double* __restrict__ dinv) {
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < s.nrows; row += gridDim.x * blockDim.x) {
        const idx_t n = s.node[row];
        const double di = dinv[n];
        if (di == 0.) continue;   // Dirichlet row
        double sd = 0.;
        for (int k = s.kptr[row]; k < s.kptr[row + 1]; ++k)
            if (s.kcol[k] == n) sd += s.kval[k];
        if (sd != 0.) dinv[n] = 1. / (1. / di + sd);
    }
}

// q[row] += sum_c S[row,c] p[c] on the free rows (p is already masked: 0 on Dirichlet nodes)
__global__ void k_surf_add(const Surf s, const double* __restrict__ dinv, const double* __restrict__ p, double* __restrict__ q) {
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < s.nrows; row += gridDim.x * blockDim.x) {
        const idx_t n = s.node[row];
        if (dinv[n] == 0.) continue;
        double sq = 0.;
        for (int k = s.kptr[row]; k < s.kptr[row + 1]; ++k) sq = fma(s.kval[k], p[s.kcol[k]], sq);
        q[n] += sq;
    }
}

// After k_fpcg (same stream): p, q, r are that launch's OUTPUT buffers.
__global__ void __launch_bounds__(256)
k_surf_iter(const Surf s, const Grid g, const double* __restrict__ p, double* __restrict__ q, const double* __restrict__ r,
            const double* __restrict__ dinv, Scalars* sc, double* partials, const PeerOut po) {
    __shared__ double sh[32 * 3];
    __shared__ int sh_flag;
    if (sc->done) return;
    const bool slab = po.q_lo != nullptr || po.q_hi != nullptr;
    double c[3] = {0., 0., 0.};
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < s.nrows; row += gridDim.x * blockDim.x) {
        const int ka = s.kptr[row], kb = s.kptr[row + 1];
        if (ka == kb) continue;
        const idx_t n = s.node[row];
        const int P = (int)(n / g.sK);
        if (P < g.kown0 || P >= g.kown1) continue;   // slab mode: the owner of the plane handles the row
        const double d = dinv[n];
        if (d == 0.) continue;
        double sq = 0.;
        for (int k = ka; k < kb; ++k) sq = fma(s.kval[k], p[s.kcol[k]], sq);
        const double q0 = q[n], qn = q0 + sq;
        q[n] = qn;
        if (slab) {
            const idx_t nip = n - g.sK * P;
            if (po.q_lo && P == g.kown0) po.q_lo[nip] = qn;
            if (po.q_hi && P == g.kown1 - 1) po.q_hi[nip] = qn;
        }
        const double z = d * r[n];
        c[0] = fma(p[n], sq, c[0]);
        c[1] = fma(sq, z, c[1]);
        c[2] = fma(d, fma(2. * q0, sq, sq * sq), c[2]);
    }
    if (grid_reduce<3, false>(c, partials, &sc->ticket[2], sh, &sh_flag, slab)) {
        if (sc->comm) rank_allreduce<3, false>(c, sc->comm, sh);
        if (threadIdx.x == 0) {
            const double pq = sc->pq + c[0], qz = sc->qz + c[1], qdq = sc->qdq + c[2], rho = sc->rho;
            sc->pq = pq;
            if (!sc->bench && !(pq > 0.)) { sc->done = 1; sc->status = (pq == pq) ? -1 : -2; }
            else if (sc->line) sc->alpha = (pq > 0.) ? rho / pq : 0.;   // line-Jacobi PCG: beta belongs to the line kernel
            else {
                const double al = (pq > 0.) ? rho / pq : 0.;
                const double rho_next = fma(al * al, qdq, fma(-2. * al, qz, rho));
                sc->alpha = al;
                sc->beta = (rho_next > 0. && rho > 0.) ? rho_next / rho : 0.;
            }
        }
    }
}

}  // namespace pfem
