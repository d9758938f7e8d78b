// kernels_simple.cuh — straightforward one-thread-per-node / per-element FP64 kernels.
// They are the device-side reference for the tiled production kernels (kernels_tiled.cuh)
// and implement every non-iteration step of the path (conductivity updates, load vector,
// diagonal, loop-error reductions, post-processing).
#pragma once
#include "pfem_internal.cuh"

namespace pfem {

#define PFEM_NODE_BLOCK_X 64
#define PFEM_NODE_BLOCK_Y 4

// ---------------------------------------------------------------- operator (simple) -----

// MODE 0: out = M (A in),  partial in.out -> pq, alpha          (CG: q = A p)
// MODE 1: out = M (f - A in)                                    (initial residual)
// MODE 2: out = M (f - A in),  partials out.out -> bb, out.D^-1 out -> bz   (norms of the lifted rhs)
// MODE 3: out = M (A in), no reduction, ignores sc->done        (tests: plain operator)
// M masks Dirichlet rows (dinv == 0).  `in` must be 0 on Dirichlet nodes for MODE 0/3.
// Dynamic3D (solvers/thermal/dynamic/femT3d.cpp:176,205-231): the operator of a time step is theta*K + C/dt.  The caller
// passes theta-scaled conductivities; `mass` (node array, may be null) is the lumped capacity diagonal, `cmass` (element
// lattice, may be null) the element capacity c of the consistent matrix C_e = c * (8,4,2,1)/27 by the number of differing axes.
template <int MODE>
__global__ void __launch_bounds__(PFEM_NODE_BLOCK_X* PFEM_NODE_BLOCK_Y)
k_apply_simple(const Grid g, const double* __restrict__ cl, const double* __restrict__ cv,
               const double* __restrict__ in, const double* __restrict__ dinv, const double* __restrict__ f,
               double* __restrict__ out, Scalars* sc, double* partials, const double* __restrict__ mass = nullptr,
               const double* __restrict__ cmass = nullptr) {
    __shared__ double sh[64];
    __shared__ int sh_flag;
    if (MODE == 0 && sc->done) return;
    const int i = blockIdx.x * PFEM_NODE_BLOCK_X + threadIdx.x;
    const int j = blockIdx.y * PFEM_NODE_BLOCK_Y + threadIdx.y;
    const int k = blockIdx.z;
    double dot[2] = {0., 0.};
    if (i < g.nI && j < g.nJ) {
        const idx_t n = i + g.sJ * j + g.sK * k;
        double P[3][3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int a = 0; a < 3; ++a) P[c][b][a] = in[n + (a - 1) + g.sJ * (b - 1) + g.sK * (c - 1)];
        double acc = 0.;
#pragma unroll
        for (int dk = -1; dk <= 0; ++dk)
#pragma unroll
            for (int dj = -1; dj <= 0; ++dj)
#pragma unroll
                for (int di = -1; di <= 0; ++di) {
                    const idx_t slot = n + di + g.sJ * dj + g.sK * dk;
                    double kI, kJ, kK, kv[8];
                    elem_conductances(g, cl[slot], cv[slot], i + di, j + dj, k + dk, kI, kJ, kK);
                    elem_matrix8(kI, kJ, kK, kv);
                    if (cmass) {
                        const double c = cmass[slot] * (1. / 27.);
#pragma unroll
                        for (int x = 0; x < 8; ++x) kv[x] += c * (double)(8 >> __popc(x));
                    }
                    // this node is local corner (-di,-dj,-dk) of the element
#pragma unroll
                    for (int l = 0; l < 8; ++l) {
                        const int bi = l & 1, bj = (l >> 1) & 1, bk = (l >> 2) & 1;
                        const int x = ((-di) ^ bi) | (((-dj) ^ bj) << 1) | (((-dk) ^ bk) << 2);
                        acc += kv[x] * P[dk + bk + 1][dj + bj + 1][di + bi + 1];
                    }
                }
        if (mass) acc = fma(mass[n], P[1][1][1], acc);
        const double dn = dinv[n];
        const bool fixed = (dn == 0.);
        double o;
        if (MODE == 0 || MODE == 3) o = fixed ? 0. : acc;
        else o = fixed ? 0. : f[n] - acc;
        out[n] = o;
        if (MODE == 0) dot[0] = P[1][1][1] * o;
        if (MODE == 2 && k >= g.kown0 && k < g.kown1) { dot[0] = o * o; dot[1] = o * o * dn; }
    }
    if (MODE == 1 || MODE == 3) return;
    if (grid_reduce<2, false>(dot, partials, &sc->ticket[0], sh, &sh_flag)) {
        if (MODE == 2 && sc->comm) rank_allreduce<2, false>(dot, sc->comm, sh);
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            if (MODE == 0) {
                sc->pq = dot[0];
                if (dot[0] > 0.) sc->alpha = sc->rho / dot[0];
                else {
                    sc->alpha = 0.;
                    if (!sc->bench) { sc->done = 1; sc->status = (dot[0] == dot[0]) ? -1 : -2; }
                }
            } else { sc->bb = dot[0]; sc->bz = dot[1]; }
        }
    }
}

// Diagonal of the eliminated matrix: sum over the 8 adjacent elements of (kI+kJ+kK)/9
// (therm3d.cpp:227); dinv = 0 marks Dirichlet rows (and rows with an empty diagonal).
__global__ void k_diag(const Grid g, const double* __restrict__ cl, const double* __restrict__ cv,
                       const uint8_t* __restrict__ fixed, double* __restrict__ dinv, Scalars* sc,
                       const double* __restrict__ mass = nullptr, const double* __restrict__ cmass = nullptr) {
    const int i = blockIdx.x * PFEM_NODE_BLOCK_X + threadIdx.x;
    const int j = blockIdx.y * PFEM_NODE_BLOCK_Y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= g.nI || j >= g.nJ) return;
    const idx_t n = i + g.sJ * j + g.sK * k;
    double d = 0.;
#pragma unroll
    for (int dk = -1; dk <= 0; ++dk)
#pragma unroll
        for (int dj = -1; dj <= 0; ++dj)
#pragma unroll
            for (int di = -1; di <= 0; ++di) {
                const idx_t slot = n + di + g.sJ * dj + g.sK * dk;
                double kI, kJ, kK;
                elem_conductances(g, cl[slot], cv[slot], i + di, j + dj, k + dk, kI, kJ, kK);
                d += (kI + kJ + kK) * (1. / 9.);
                if (cmass) d += cmass[slot] * (8. / 27.);
            }
    if (mass) d += mass[n];
    if (!fixed[n] && d < 0.) sc->neg_diag = 1;
    dinv[n] = (fixed[n] || !(d > 0.)) ? 0. : 1. / d;
}

// Load vector: B[node] += 0.125e-18*dx*dy*dz*heat[e] over the 8 adjacent elements
// (therm3d.cpp:223,274).  heat lives on the padded element lattice (0 in padding).
// mat != nullptr: elements marked PFEM_MAT_EXCLUDED (outside the masked mesh) carry no load.
__global__ void k_load_vector(const Grid g, const double* __restrict__ heat, const uint32_t* __restrict__ mat, double* __restrict__ f) {
    const int i = blockIdx.x * PFEM_NODE_BLOCK_X + threadIdx.x;
    const int j = blockIdx.y * PFEM_NODE_BLOCK_Y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= g.nI || j >= g.nJ) return;
    const idx_t n = i + g.sJ * j + g.sK * k;
    double s = 0.;
#pragma unroll
    for (int dk = -1; dk <= 0; ++dk)
#pragma unroll
        for (int dj = -1; dj <= 0; ++dj)
#pragma unroll
            for (int di = -1; di <= 0; ++di) {
                const idx_t slot = n + di + g.sJ * dj + g.sK * dk;
                if (mat && mat[slot] == PFEM_MAT_EXCLUDED) continue;
                s += 0.125e-18 * g.hI[i + di] * g.hJ[j + dj] * g.hK[k + dk] * heat[slot];
            }
    f[n] = s;
}

// Masked mesh: a node belongs to the masked mesh iff one of its (up to 8) adjacent elements is kept.  The other nodes
// do not exist in the reference (RectangularMaskedMesh3D); here they are rows with an empty diagonal (dinv = 0) whose
// field value is pinned to 0 so that they contribute nothing to any norm, maximum or gradient.
__global__ void k_mark_inactive(const Grid g, const uint32_t* __restrict__ mat, uint8_t* __restrict__ inactive) {
    const int i = blockIdx.x * PFEM_NODE_BLOCK_X + threadIdx.x;
    const int j = blockIdx.y * PFEM_NODE_BLOCK_Y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= g.nI || j >= g.nJ) return;
    const idx_t n = i + g.sJ * j + g.sK * k;
    bool any = false;
#pragma unroll
    for (int dk = -1; dk <= 0; ++dk)
#pragma unroll
        for (int dj = -1; dj <= 0; ++dj)
#pragma unroll
            for (int di = -1; di <= 0; ++di) {
                const int ei = i + di, ej = j + dj, ek = k + dk;
                if (ei < 0 || ej < 0 || ek < 0 || ei >= g.nI - 1 || ej >= g.nJ - 1 || ek >= g.nK - 1) continue;
                if (mat[n + di + g.sJ * dj + g.sK * dk] != PFEM_MAT_EXCLUDED) any = true;
            }
    inactive[n] = any ? 0 : 1;
}
__global__ void k_zero_inactive(idx_t N, const uint8_t* __restrict__ inactive, double* __restrict__ x) {
    for (idx_t n = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; n < N; n += (idx_t)gridDim.x * blockDim.x)
        if (inactive[n]) x[n] = 0.;
}

// ------------------------------------------------------------ vector kernels (PCG) ------

// p = dinv*r + beta*p   (simple variant only; the tiled operator kernel fuses this)
__global__ void k_pupdate(idx_t N, const double* __restrict__ r, const double* __restrict__ dinv,
                          double* __restrict__ p, const Scalars* sc) {
    if (sc->done) return;
    const double beta = sc->beta;
    for (idx_t n = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; n < N; n += (idx_t)gridDim.x * blockDim.x)
        p[n] = dinv[n] * r[n] + beta * p[n];
}

// x += alpha p; r -= alpha q; rho = sum r^2 dinv; rr = sum r^2; zz = sum (dinv r)^2; xx = sum x^2;
// then (last block) the scalar recurrences of CG (itcg, extlib/nspcg/nspcg.f:9281-9337) and the
// stopping test: ||r|| <= tol ||b||, ||r||_D^-1 <= tol ||b||_D^-1 and ||D^-1 r|| <= tol ||x||
// (the last is NSPCG's pseudo-residual test #2, nspcg.f:26487-26518, without the eigenvalue
// estimate; it is what makes nearly floating nodes — air pockets — converge too).
// INIT = true: no update, only the two reductions and the initial scalars.
template <bool INIT>
__global__ void __launch_bounds__(256)
k_update(idx_t N, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
         const double* __restrict__ q, const double* __restrict__ dinv, Scalars* sc, double* partials) {
    __shared__ double sh[128];
    __shared__ int sh_flag;
    if (!INIT && sc->done) return;
    const double alpha = INIT ? 0. : sc->alpha;
    double v[4] = {0., 0., 0., 0.};
    const idx_t nthreads = (idx_t)gridDim.x * blockDim.x;
    const idx_t t = blockIdx.x * (idx_t)blockDim.x + threadIdx.x;
    const idx_t N2 = N >> 1;  // arrays are 16-byte aligned (guard band is a multiple of 16 doubles)
    for (idx_t m = t; m < N2; m += nthreads) {
        double2 rr = reinterpret_cast<const double2*>(r)[m];
        const double2 dd = reinterpret_cast<const double2*>(dinv)[m];
        double2 xx = reinterpret_cast<double2*>(x)[m];
        if (!INIT) {
            const double2 pp = reinterpret_cast<const double2*>(p)[m];
            const double2 qq = reinterpret_cast<const double2*>(q)[m];
            xx.x += alpha * pp.x; xx.y += alpha * pp.y;
            rr.x -= alpha * qq.x; rr.y -= alpha * qq.y;
            reinterpret_cast<double2*>(x)[m] = xx;
            reinterpret_cast<double2*>(r)[m] = rr;
        }
        const double zx = rr.x * dd.x, zy = rr.y * dd.y;
        v[0] += rr.x * zx + rr.y * zy;
        v[1] += rr.x * rr.x + rr.y * rr.y;
        v[2] += zx * zx + zy * zy;
        v[3] += xx.x * xx.x + xx.y * xx.y;
    }
    if ((N & 1) && t == 0) {
        const idx_t n = N - 1;
        double rn = r[n], xn = x[n];
        if (!INIT) { xn += alpha * p[n]; x[n] = xn; rn -= alpha * q[n]; r[n] = rn; }
        const double zn = rn * dinv[n];
        v[0] += rn * zn;
        v[1] += rn * rn;
        v[2] += zn * zn;
        v[3] += xn * xn;
    }
    if (grid_reduce<4, false>(v, partials, &sc->ticket[1], sh, &sh_flag)) {
        if (INIT && sc->comm) rank_allreduce<4, false>(v, sc->comm, sh);
        if (threadIdx.x == 0) {
            if (INIT) {
                sc->rho = v[0]; sc->rho_prev = v[0]; sc->rr = v[1]; sc->zz = v[2]; sc->xx = v[3];
                sc->beta = 0.; sc->alpha = 0.; sc->pq = 0.; sc->iter = 0; sc->launch = 0; sc->status = 0; sc->done = 0;
                if (!(v[1] == v[1])) { sc->done = 1; sc->status = -2; }
                else if (!sc->bench && v[1] <= sc->tol2 * sc->bb && v[0] <= sc->tol2 * sc->bz && v[2] <= sc->tol2 * v[3]) { sc->done = 1; sc->status = 1; }
            } else {
                const double rho_old = sc->rho;
                sc->rho_prev = rho_old; sc->rho = v[0]; sc->rr = v[1]; sc->zz = v[2]; sc->xx = v[3];
                sc->beta = (rho_old > 0.) ? v[0] / rho_old : 0.;
                const int it = sc->iter + 1;
                sc->iter = it;
                if (!sc->bench) {
                    if (!(v[1] == v[1])) { sc->done = 1; sc->status = -2; }
                    else if (v[1] <= sc->tol2 * sc->bb && v[0] <= sc->tol2 * sc->bz && v[2] <= sc->tol2 * v[3]) { sc->done = 1; sc->status = 1; }
                    else if (it >= sc->maxit) { sc->done = 1; sc->status = 2; }
                }
            }
        }
    }
}

// ---- slab mode helpers ---------------------------------------------------------------------
// copy my first / last owned plane into the halo plane of the lower / upper neighbour (peer-mapped memory)
__global__ void k_halo_push(idx_t len, const double* __restrict__ src_lo, double* __restrict__ dst_lo,
                            const double* __restrict__ src_hi, double* __restrict__ dst_hi) {
    for (idx_t n = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; n < len; n += (idx_t)gridDim.x * blockDim.x) {
        if (dst_lo) dst_lo[n] = src_lo[n];
        if (dst_hi) dst_hi[n] = src_hi[n];
    }
}
// cross-rank barrier (also publishes the halo stores of the preceding kernel); ORs a per-rank flag
__global__ void k_rank_barrier(Scalars* sc, int flag_in) {
    __shared__ double sh[8];
    double v[1] = {(double)flag_in};
    __threadfence_system();
    rank_allreduce<1, true>(v, sc->comm, sh);
    if (threadIdx.x == 0) sc->red[3] = v[0];
}

// q = fixed ? x : 0  (the Dirichlet-only vector used to lift the boundary values into the rhs)
__global__ void k_dirichlet_only(idx_t N, const double* __restrict__ x, const uint8_t* __restrict__ fixed,
                                 double* __restrict__ out) {
    for (idx_t n = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; n < N; n += (idx_t)gridDim.x * blockDim.x)
        out[n] = fixed[n] ? x[n] : 0.;
}
// out = fixed ? a : b
__global__ void k_select_fixed(idx_t N, const uint8_t* __restrict__ fixed, const double* __restrict__ a,
                               const double* __restrict__ b, double* __restrict__ out) {
    for (idx_t n = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; n < N; n += (idx_t)gridDim.x * blockDim.x)
        out[n] = fixed[n] ? a[n] : b[n];
}
// The host passes a de-duplicated list (last value wins, like B[r] = val in setBC), so a plain
// scatter is race free.  x and/or fixed may be null.
__global__ void k_scatter_dirichlet(size_t nd, const idx_t* __restrict__ node, const double* __restrict__ value,
                                    double* __restrict__ x, uint8_t* __restrict__ fixed) {
    for (size_t m = blockIdx.x * (size_t)blockDim.x + threadIdx.x; m < nd; m += (size_t)gridDim.x * blockDim.x) {
        if (x) x[node[m]] = value[m];
        if (fixed) fixed[node[m]] = 1;
    }
}
// fill the true nodes of a pitched lattice array (pad entries stay 0)
__global__ void k_fill_nodes(const Grid g, double* __restrict__ x, double v) {
    const int i = blockIdx.x * PFEM_NODE_BLOCK_X + threadIdx.x;
    const int j = blockIdx.y * PFEM_NODE_BLOCK_Y + threadIdx.y;
    const int k = blockIdx.z;
    if (i < g.nI && j < g.nJ) x[i + g.sJ * j + g.sK * k] = v;
}
__global__ void k_diag_from_dinv(idx_t N, const double* __restrict__ dinv, const uint8_t* __restrict__ fixed,
                                 double* __restrict__ d) {
    for (idx_t n = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; n < N; n += (idx_t)gridDim.x * blockDim.x)
        d[n] = fixed[n] ? 1. : (dinv[n] != 0. ? 1. / dinv[n] : 0.);
}

// ------------------------------------------------- element lattice <-> compact order ----

// Compact (ABI order: the element mesh keeps the node iteration order of the ABI, e = e_minor + n_minor'*(e_medium +
// n_medium'*e_major)) <-> padded node-lattice slots.  g.abi_dim says which index-space axis each ABI axis is.
__device__ __forceinline__ idx_t compact_to_slot(const Grid& g, idx_t e) {
    const idx_t en[3] = {g.nI - 1, g.nJ - 1, g.nK - 1}, st[3] = {1, g.sJ, g.sK};
    const idx_t n0 = en[g.abi_dim[0]], n1 = en[g.abi_dim[1]];
    const idx_t m0 = e % n0, t = e / n0;
    const idx_t m1 = t % n1, m2 = t / n1;
    return m0 * st[g.abi_dim[0]] + m1 * st[g.abi_dim[1]] + m2 * st[g.abi_dim[2]];
}
__device__ __forceinline__ idx_t slot_to_compact(const Grid& g, int i, int j, int k) {
    const idx_t en[3] = {g.nI - 1, g.nJ - 1, g.nK - 1};
    const int ii[3] = {i, j, k};
    return ii[g.abi_dim[0]] + en[g.abi_dim[0]] * (ii[g.abi_dim[1]] + en[g.abi_dim[1]] * (idx_t)ii[g.abi_dim[2]]);
}
// dense ABI node array <-> pitched lattice when the internal layout differs from the ABI's iteration order
__global__ void k_node_expand(const Grid g, const double* __restrict__ dense, double* __restrict__ lat) {
    const int i = blockIdx.x * PFEM_NODE_BLOCK_X + threadIdx.x;
    const int j = blockIdx.y * PFEM_NODE_BLOCK_Y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= g.nI || j >= g.nJ) return;
    lat[i + g.sJ * j + g.sK * k] = dense[i * g.abi_ns[0] + j * g.abi_ns[1] + k * g.abi_ns[2]];
}
__global__ void k_node_compact(const Grid g, const double* __restrict__ lat, double* __restrict__ dense) {
    const int i = blockIdx.x * PFEM_NODE_BLOCK_X + threadIdx.x;
    const int j = blockIdx.y * PFEM_NODE_BLOCK_Y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= g.nI || j >= g.nJ) return;
    dense[i * g.abi_ns[0] + j * g.abi_ns[1] + k * g.abi_ns[2]] = lat[i + g.sJ * j + g.sK * k];
}
template <typename T, int NC>
__global__ void k_elem_expand(const Grid g, const T* __restrict__ src, T* __restrict__ dst0, T* __restrict__ dst1,
                              T* __restrict__ dst2) {
    for (idx_t e = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; e < g.E; e += (idx_t)gridDim.x * blockDim.x) {
        const idx_t slot = compact_to_slot(g, e);
        dst0[slot] = src[e * NC];
        if (NC > 1) dst1[slot] = src[e * NC + 1];
        if (NC > 2) dst2[slot] = src[e * NC + 2];
    }
}
template <typename T, int NC>
__global__ void k_elem_compact(const Grid g, const T* __restrict__ src0, const T* __restrict__ src1,
                               const T* __restrict__ src2, T* __restrict__ dst) {
    for (idx_t e = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; e < g.E; e += (idx_t)gridDim.x * blockDim.x) {
        const idx_t slot = compact_to_slot(g, e);
        dst[e * NC] = src0[slot];
        if (NC > 1) dst[e * NC + 1] = src1[slot];
        if (NC > 2) dst[e * NC + 2] = src2[slot];
    }
}

// Decode helper for per-element kernels launched over the node lattice: returns false for
// padding slots.  (pi[a] = element index along PHYSICAL axis a.)
__device__ __forceinline__ bool elem_slot(const Grid& g, int i, int j, int k, int (&pi)[3]) {
    if (i >= g.nI - 1 || j >= g.nJ - 1 || k >= g.nK - 1) return false;
    const int ii[3] = {i, j, k};
    pi[0] = ii[g.dim_of_phys[0]];
    pi[1] = ii[g.dim_of_phys[1]];
    pi[2] = ii[g.dim_of_phys[2]];
    return true;
}

// Linear table interpolation, clamped (same arithmetic as oracle table_at()).
__device__ __forceinline__ double table_at(const double* __restrict__ tab, uint32_t mat, uint32_t nT, double T0,
                                           double dT, double T) {
    double t = (T - T0) / dT;
    if (!(t > 0.)) t = 0.;
    const double tmax = (double)(nT - 1);
    if (t > tmax) t = tmax;
    uint32_t i = (uint32_t)t;
    if (i > nT - 2) i = nT - 2;
    const double f = t - (double)i;
    const double* row = tab + (size_t)mat * nT;
    const double a = row[i], b = row[i + 1];
    return __dadd_rn(a, __dmul_rn(f, __dsub_rn(b, a)));  // no FMA contraction: bit-equal to the oracle
}

// ------------------------------------------------------- conductivity updates -----------

// Thermal: T_mean = 0.125 * sum of the 8 node temperatures in the reference's corner order,
// conds = table(T_mean)   (therm3d.cpp:208-213)
__global__ void k_cond_thermal(const Grid g, const double* __restrict__ T, const uint32_t* __restrict__ mat,
                               uint32_t nT, double T0, double dT, const double* __restrict__ tab_lat,
                               const double* __restrict__ tab_vert, double* __restrict__ cl, double* __restrict__ cv) {
    const int i = blockIdx.x * PFEM_NODE_BLOCK_X + threadIdx.x;
    const int j = blockIdx.y * PFEM_NODE_BLOCK_Y + threadIdx.y;
    const int k = blockIdx.z;
    int pi[3];
    if (!elem_slot(g, i, j, k, pi)) return;
    const idx_t n = i + g.sJ * j + g.sK * k;
    double temp = 0.;
#pragma unroll
    for (int l = 0; l < 8; ++l)
        temp = __dadd_rn(temp, T[n + ((l & 1) ? g.ps[0] : 0) + ((l & 2) ? g.ps[1] : 0) + ((l & 4) ? g.ps[2] : 0)]);
    temp *= 0.125;
    const uint32_t m = mat[n];
    if (m == PFEM_MAT_EXCLUDED) { cl[n] = 0.; cv[n] = 0.; return; }   // element outside the masked mesh
    cl[n] = table_at(tab_lat, m, nT, T0, dT, temp);
    cv[n] = table_at(tab_vert, m, nT, T0, dT, temp);
}

// ------------------------------------------------------- Dynamic3D (femT3d.cpp) -----------

// Element capacity c = cp(T_e)*dens(T_e) * 0.125e-9 * dx*dy*dz / timestep (femT3d.cpp:176) with T_e the mean of the 8 node
// temperatures (:163) and cp*dens from the per-material table; elements outside the masked mesh carry none.
// lumped != 0: the value each of the 8 nodes receives on its diagonal (:207-212); else the c of the consistent matrix
// (:216-231), whose entries are c*(8,4,2,1)/27.
__global__ void k_elem_capacity(const Grid g, const double* __restrict__ T, const uint32_t* __restrict__ mat, uint32_t nT,
                                double T0, double dT, const double* __restrict__ tab_cprho, double inv_timestep,
                                double* __restrict__ ce) {
    const int i = blockIdx.x * PFEM_NODE_BLOCK_X + threadIdx.x;
    const int j = blockIdx.y * PFEM_NODE_BLOCK_Y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= g.nI || j >= g.nJ) return;
    const idx_t n = i + g.sJ * j + g.sK * k;
    int pi[3];
    if (!elem_slot(g, i, j, k, pi)) { ce[n] = 0.; return; }
    const uint32_t m = mat[n];
    if (m == PFEM_MAT_EXCLUDED) { ce[n] = 0.; return; }
    double temp = 0.;
#pragma unroll
    for (int l = 0; l < 8; ++l)
        temp = __dadd_rn(temp, T[n + ((l & 1) ? g.ps[0] : 0) + ((l & 2) ? g.ps[1] : 0) + ((l & 4) ? g.ps[2] : 0)]);
    temp *= 0.125;
    ce[n] = table_at(tab_cprho, m, nT, T0, dT, temp) * 0.125e-9 * g.hI[i] * g.hJ[j] * g.hK[k] * inv_timestep;
}

// lumped capacity diagonal: mass[node] = sum of c over the 8 adjacent elements (femT3d.cpp:207-212)
__global__ void k_mass_lumped(const Grid g, const double* __restrict__ ce, double* __restrict__ mass) {
    const int i = blockIdx.x * PFEM_NODE_BLOCK_X + threadIdx.x;
    const int j = blockIdx.y * PFEM_NODE_BLOCK_Y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= g.nI || j >= g.nJ) return;
    const idx_t n = i + g.sJ * j + g.sK * k;
    double s = 0.;
#pragma unroll
    for (int dk = -1; dk <= 0; ++dk)
#pragma unroll
        for (int dj = -1; dj <= 0; ++dj)
#pragma unroll
            for (int di = -1; di <= 0; ++di) s += ce[n + di + g.sJ * dj + g.sK * dk];
    mass[n] = s;
}

// out = a + b: the right-hand side of a time step, B T + F = M (F - K T) + M (A T) with B = A - K (femT3d.cpp:203-204,279-280)
__global__ void k_axpby(idx_t N, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
    for (idx_t n = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; n < N; n += (idx_t)gridDim.x * blockDim.x) out[n] = a[n] + b[n];
}

__global__ void k_scale2(idx_t N, double s, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ ao,
                         double* __restrict__ bo) {
    for (idx_t n = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; n < N; n += (idx_t)gridDim.x * blockDim.x) {
        ao[n] = s * a[n];
        bo[n] = s * b[n];
    }
}

struct JunctionDev {
    idx_t bottom, top, left, right, back, front, ld, offset;
    double height;
};

// loadConductivity (electr3d.cpp:203-225)
__global__ void k_cond_shockley(const Grid g, const uint32_t* __restrict__ mat, const uint32_t* __restrict__ junc,
                                const uint8_t* __restrict__ role, const double* __restrict__ Te, uint32_t nT, double T0,
                                double dT, const double* __restrict__ tab_lat, const double* __restrict__ tab_vert,
                                const JunctionDev* __restrict__ act, const double* __restrict__ junc_cond, double pcond,
                                double ncond, double* __restrict__ cl, double* __restrict__ cv) {
    const int i = blockIdx.x * PFEM_NODE_BLOCK_X + threadIdx.x;
    const int j = blockIdx.y * PFEM_NODE_BLOCK_Y + threadIdx.y;
    const int k = blockIdx.z;
    int pi[3];
    if (!elem_slot(g, i, j, k, pi)) return;
    const idx_t n = i + g.sJ * j + g.sK * k;
    const uint32_t actn = junc ? junc[n] : 0u;
    double c0, c1;
    if (mat[n] == PFEM_MAT_EXCLUDED) {   // element outside the masked mesh
        c0 = c1 = 0.;
    } else if (actn) {
        const JunctionDev a = act[actn - 1];
        const idx_t col = a.offset + a.ld * pi[1] + pi[0];
        c0 = junc_cond[2 * col];
        c1 = junc_cond[2 * col + 1];
        if (isnan(c1) || fabs(c1) < 1e-16) c1 = 1e-16;
    } else if (role && role[n] == 1) {
        c0 = c1 = pcond;
    } else if (role && role[n] == 2) {
        c0 = c1 = ncond;
    } else {
        const double T = Te[n];
        c0 = table_at(tab_lat, mat[n], nT, T0, dT, T);
        c1 = table_at(tab_vert, mat[n], nT, T0, dT, T);
    }
    cl[n] = c0;
    cv[n] = c1;
}

// Junction update (electr3d.cpp:246-274) with the Shockley law of beta.hpp:43-46.
__global__ void k_junction_update(const Grid g, const uint32_t* __restrict__ junc, const JunctionDev* __restrict__ act,
                                  const double* __restrict__ phi, const double* __restrict__ beta_col,
                                  const double* __restrict__ js_col, int stable, double* __restrict__ cl,
                                  double* __restrict__ cv) {
    const int i = blockIdx.x * PFEM_NODE_BLOCK_X + threadIdx.x;
    const int j = blockIdx.y * PFEM_NODE_BLOCK_Y + threadIdx.y;
    const int k = blockIdx.z;
    int pi[3];
    if (!elem_slot(g, i, j, k, pi)) return;
    const idx_t n = i + g.sJ * j + g.sK * k;
    const uint32_t nact = junc[n];
    if (!nact) return;
    const JunctionDev a = act[nact - 1];
    const idx_t b0 = (idx_t)pi[0] * g.ps[0], f0 = b0 + g.ps[0], l1 = (idx_t)pi[1] * g.ps[1], r1 = l1 + g.ps[1];
    const idx_t zb = a.bottom * g.ps[2], zt = a.top * g.ps[2];
    const double U = 0.25 * (-phi[b0 + l1 + zb] - phi[f0 + l1 + zb] - phi[b0 + r1 + zb] - phi[f0 + r1 + zb] +
                             phi[b0 + l1 + zt] + phi[f0 + l1 + zt] + phi[b0 + r1 + zt] + phi[f0 + r1 + zt]);
    double jy = 0.1 * cv[n] * U / a.height;
    const idx_t col = a.offset + a.ld * pi[1] + pi[0];
    jy = fabs(jy);
    double c0 = 0., c1 = 10. * jy * a.height * beta_col[col] / log(1e7 * jy / js_col[col] + 1.);
    if (stable) { c0 = 0.5 * (cl[n] + c0); c1 = 0.5 * (cv[n] + c1); }
    if (isnan(c1) || fabs(c1) < 1e-16) c1 = 1e-16;
    cl[n] = c0;
    cv[n] = c1;
}

// saveConductivity (electr3d.cpp:227-237): junction table <- conds of the mid-plane element
__global__ void k_junction_save(const Grid g, const JunctionDev* __restrict__ act, int nact,
                                const double* __restrict__ cl, const double* __restrict__ cv,
                                double* __restrict__ junc_cond) {
    for (int a = 0; a < nact; ++a) {
        const JunctionDev A = act[a];
        const idx_t v = (A.top + A.bottom) / 2;
        const idx_t nl = A.front - A.back, nt = A.right - A.left;
        for (idx_t m = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; m < nl * nt; m += (idx_t)gridDim.x * blockDim.x) {
            const idx_t t = A.left + m / nl, l = A.back + m % nl;
            const idx_t slot = l * g.ps[0] + t * g.ps[1] + v * g.ps[2];
            const idx_t col = A.offset + A.ld * t + l;
            junc_cond[2 * col] = cl[slot];
            junc_cond[2 * col + 1] = cv[slot];
        }
    }
}

// ----------------------------------------------------- loop errors & post-processing ----

// err = max|T - T_prev|, maxT (therm3d.cpp:318-325); also T_prev <- T for the next loop is done
// by the caller with a device copy.
__global__ void __launch_bounds__(256)
k_thermal_error(idx_t N, const double* __restrict__ T, const double* __restrict__ Tp, Scalars* sc, double* partials) {
    __shared__ double sh[64];
    __shared__ int sh_flag;
    double v[2] = {0., 0.};
    for (idx_t n = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; n < N; n += (idx_t)gridDim.x * blockDim.x) {
        const double t = T[n];
        v[0] = fmax(v[0], fabs(Tp[n] - t));
        v[1] = fmax(v[1], t);
    }
    if (grid_reduce<2, true>(v, partials, &sc->ticket[2], sh, &sh_flag)) {
        if (sc->comm) rank_allreduce<2, true>(v, sc->comm, sh);
        if (threadIdx.x == 0) { sc->red[0] = v[0]; sc->red[1] = v[1]; }
    }
}

// Mean gradient of a nodal field over one brick, physical axis order, exactly the sign
// pattern of electr3d.cpp:399-410 / therm3d.cpp:372-381 (sum of 4 edge differences).
__device__ __forceinline__ void brick_edge_sums(const Grid& g, const double* __restrict__ u, idx_t n, double& s0,
                                                double& s1, double& s2) {
    const double lll = u[n], ull = u[n + g.ps[0]], lul = u[n + g.ps[1]], uul = u[n + g.ps[0] + g.ps[1]];
    const double llu = u[n + g.ps[2]], ulu = u[n + g.ps[0] + g.ps[2]], luu = u[n + g.ps[1] + g.ps[2]],
                 uuu = u[n + g.ps[0] + g.ps[1] + g.ps[2]];
    s0 = -lll - llu - lul - luu + ull + ulu + uul + uuu;
    s1 = -lll - llu + lul + luu - ull - ulu + uul + uuu;
    s2 = -lll + llu - lul + luu - ull + ulu - uul + uuu;
}
__device__ __forceinline__ void brick_sizes(const Grid& g, const int (&pi)[3], double& d0, double& d1, double& d2) {
    const double* h[3] = {g.uI, g.uJ, g.uK};
    d0 = h[g.dim_of_phys[0]][pi[0]];
    d1 = h[g.dim_of_phys[1]][pi[1]];
    d2 = h[g.dim_of_phys[2]][pi[2]];
}

// Element current densities + loop error (electr3d.cpp:387-425).  red[0] = max |dj|^2,
// red[1] = max |j|^2 over junction elements (all if noactive); maxcur = j at that element
// (first element in reference order among equal maxima).
__global__ void __launch_bounds__(PFEM_NODE_BLOCK_X* PFEM_NODE_BLOCK_Y)
k_currents(const Grid g, const double* __restrict__ phi, const double* __restrict__ cl, const double* __restrict__ cv,
           const uint32_t* __restrict__ junc, int noactive, double* __restrict__ c0, double* __restrict__ c1,
           double* __restrict__ c2, Scalars* sc, double* partials, long long* partial_idx) {
    __shared__ double sh[64];
    __shared__ int sh_flag;
    __shared__ long long sh_idx[32];
    const int i = blockIdx.x * PFEM_NODE_BLOCK_X + threadIdx.x;
    const int j = blockIdx.y * PFEM_NODE_BLOCK_Y + threadIdx.y;
    const int k = blockIdx.z;
    int pi[3];
    double v[2] = {0., 0.};
    long long my_idx = 0x7fffffffffffffffLL;
    double my_cur = -1.;
    if (elem_slot(g, i, j, k, pi)) {
        const idx_t n = i + g.sJ * j + g.sK * k;
        double s0, s1, s2, d0, d1, d2;
        brick_edge_sums(g, phi, n, s0, s1, s2);
        brick_sizes(g, pi, d0, d1, d2);
        const double a = cl[n], b = cv[n];
        const double j0 = -0.025 * a * s0 / d0, j1 = -0.025 * a * s1 / d1, j2 = -0.025 * b * s2 / d2;
        if (noactive || junc[n]) {
            my_cur = j0 * j0 + j1 * j1 + j2 * j2;
            my_idx = slot_to_compact(g, i, j, k);
            v[1] = my_cur;
        }
        const double e0 = c0[n] - j0, e1 = c1[n] - j1, e2 = c2[n] - j2;
        v[0] = e0 * e0 + e1 * e1 + e2 * e2;
        c0[n] = j0; c1[n] = j1; c2[n] = j2;
    }
    // arg-max: block max of v[1], then the smallest reference element index that attains it
    const int tid = threadIdx.x + blockDim.x * threadIdx.y;
    const unsigned int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    const unsigned int nblk = gridDim.x * gridDim.y * gridDim.z;
    double bv[2] = {v[0], v[1]};
    block_reduce<2, true>(bv, sh);
    __shared__ double sh_max;
    if (tid == 0) sh_max = bv[1];
    __syncthreads();
    long long cand = (my_cur == sh_max) ? my_idx : 0x7fffffffffffffffLL;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { long long t = __shfl_down_sync(0xffffffffu, cand, o); cand = t < cand ? t : cand; }
    if ((tid & 31) == 0) sh_idx[tid >> 5] = cand;
    __syncthreads();
    if (tid == 0) {
        long long best = sh_idx[0];
        for (int w = 1; w < (int)((blockDim.x * blockDim.y + 31) >> 5); ++w) best = sh_idx[w] < best ? sh_idx[w] : best;
        partials[(size_t)bid * 2] = bv[0];
        partials[(size_t)bid * 2 + 1] = bv[1];
        partial_idx[bid] = best;
        __threadfence();
        sh_flag = (atomicAdd(&sc->ticket[3], 1u) == nblk - 1);
    }
    __syncthreads();
    if (!sh_flag) return;
    __threadfence();
    // last block: serial-in-thread strided scan, then block combine
    double m0 = 0., m1 = -1.;
    long long mi = 0x7fffffffffffffffLL;
    for (unsigned int b = tid; b < nblk; b += blockDim.x * blockDim.y) {
        const double x0 = __ldcg(partials + (size_t)b * 2), x1 = __ldcg(partials + (size_t)b * 2 + 1);
        const long long xi = __ldcg(partial_idx + b);
        m0 = fmax(m0, x0);
        if (x1 > m1 || (x1 == m1 && xi < mi)) { m1 = x1; mi = xi; }
    }
    double w[2] = {m0, m1};
    block_reduce<2, true>(w, sh);
    if (tid == 0) sh_max = w[1];
    __syncthreads();
    cand = (m1 == sh_max) ? mi : 0x7fffffffffffffffLL;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { long long t = __shfl_down_sync(0xffffffffu, cand, o); cand = t < cand ? t : cand; }
    if ((tid & 31) == 0) sh_idx[tid >> 5] = cand;
    __syncthreads();
    if (tid == 0) {
        long long best = sh_idx[0];
        for (int q = 1; q < (int)((blockDim.x * blockDim.y + 31) >> 5); ++q) best = sh_idx[q] < best ? sh_idx[q] : best;
        sc->red[0] = w[0];
        sc->red[1] = fmax(w[1], 0.);
        sc->argidx = best;
        sc->ticket[3] = 0u;
    }
}

// maxcur = current at the arg-max element (compact index -> slot).  One warp.
// Slab mode: the loop error and max |j|^2 become maxima over the ranks; among ranks that attain the maximum the
// lowest one wins (slabs are ordered along the major axis, so this is still "first element in reference order")
// and its current vector is broadcast.  Element layers shared with a neighbour are evaluated by both ranks from
// identical data, which is harmless for maxima.
__global__ void k_fetch_maxcur(const Grid g, const double* __restrict__ c0, const double* __restrict__ c1,
                               const double* __restrict__ c2, Scalars* sc) {
    __shared__ double sh[8];
    const long long e = sc->argidx;
    double cur[3] = {0., 0., 0.};
    const bool have = !(e < 0 || e >= g.E);
    if (have) {
        const idx_t slot = compact_to_slot(g, e);
        cur[0] = c0[slot]; cur[1] = c1[slot]; cur[2] = c2[slot];
    }
    Comm* cm = sc->comm;
    if (cm) {
        const double mine = sc->red[1];
        double v[2] = {sc->red[0], mine};
        rank_allreduce<2, true>(v, cm, sh);
        if (threadIdx.x == 0) { sh[4] = v[0]; sh[5] = v[1]; }
        __syncthreads();
        const double g0 = sh[4], g1 = sh[5];
        double w[1] = {(have && mine == g1) ? -(double)cm->rank : -1e9};
        rank_allreduce<1, true>(w, cm, sh);
        if (threadIdx.x == 0) sh[6] = w[0];
        __syncthreads();
        const bool winner = (sh[6] == -(double)cm->rank);
        double b[3] = {winner ? cur[0] : 0., winner ? cur[1] : 0., winner ? cur[2] : 0.};
        rank_allreduce<3, false>(b, cm, sh);
        if (threadIdx.x == 0) { cur[0] = b[0]; cur[1] = b[1]; cur[2] = b[2]; sc->red[0] = g0; sc->red[1] = g1; }
    }
    if (threadIdx.x == 0) { sc->maxcur[0] = cur[0]; sc->maxcur[1] = cur[1]; sc->maxcur[2] = cur[2]; }
}

// Joule heat (electr3d.cpp:444-478) and heat flux (therm3d.cpp:342-384); o0..o2 padded outputs.
template <bool HEAT>
__global__ void k_gradient_fields(const Grid g, const double* __restrict__ u, const double* __restrict__ cl,
                                  const double* __restrict__ cv, const uint8_t* __restrict__ noheat,
                                  double* __restrict__ o0, double* __restrict__ o1, double* __restrict__ o2) {
    const int i = blockIdx.x * PFEM_NODE_BLOCK_X + threadIdx.x;
    const int j = blockIdx.y * PFEM_NODE_BLOCK_Y + threadIdx.y;
    const int k = blockIdx.z;
    int pi[3];
    if (!elem_slot(g, i, j, k, pi)) return;
    const idx_t n = i + g.sJ * j + g.sK * k;
    double s0, s1, s2, d0, d1, d2;
    brick_edge_sums(g, u, n, s0, s1, s2);
    brick_sizes(g, pi, d0, d1, d2);
    const double a = cl[n], b = cv[n];
    if (HEAT) {
        const double dvx = -0.25e6 * s0 / d0, dvy = -0.25e6 * s1 / d1, dvz = -0.25e6 * s2 / d2;
        o0[n] = (noheat && noheat[n]) ? 0. : a * dvx * dvx + a * dvy * dvy + b * dvz * dvz;
    } else {
        o0[n] = -0.25e6 * a * s0 / d0;
        o1[n] = -0.25e6 * a * s1 / d1;
        o2[n] = -0.25e6 * b * s2 / d2;
    }
}

// ----------------------------------------------------- provider interpolation ----------
// RectilinearMesh3D::interpolateLinear (plask/mesh/rectilinear3d.hpp:802-845) of a source lattice array onto the
// element midpoints of the destination mesh.  The per-axis bracketing (prepareInterpolationForAxis,
// plask/mesh/axis1d.cpp:99-156) is tabulated by the host per destination element index of each PHYSICAL axis:
// lo/hi source indices, their (possibly faked) coordinates and the point coordinate.  Arithmetic in the order of
// interpolation::trilinear (plask/utils/interpolation.hpp:31-71), no FMA contraction: bit-equal to the oracle.
struct InterpAxis {
    const int *ilo, *ihi;
    const double *lo, *hi, *pt;
};
// What the providers do with points the plain interpolation does not cover:
//   bbox     getHeatDensity returns 0 outside the bounding box of the source geometry (electr3d.cpp:545-548); the box is
//            taken as the extent of the source node axes (bb[2a], bb[2a+1]), bounds included like Box3D::contains
//   src_mat  source on a masked mesh: interpolate(maskedMesh, ...) yields NaN outside the kept elements and SafeData
//            substitutes `fill` (getTemperatures, therm3d.cpp:391-392: 300 K); src_mat = material ids on the source lattice
struct InterpOutside {
    int bbox;
    double bb[6];
    const uint32_t* src_mat;
    double fill;
};
__global__ void k_interp_to_elems(const Grid gd, const idx_t ss0, const idx_t ss1, const idx_t ss2,
                                  const double* __restrict__ src, const InterpAxis a0, const InterpAxis a1,
                                  const InterpAxis a2, double* __restrict__ dst, const InterpOutside out) {
    const int i = blockIdx.x * PFEM_NODE_BLOCK_X + threadIdx.x;
    const int j = blockIdx.y * PFEM_NODE_BLOCK_Y + threadIdx.y;
    const int k = blockIdx.z;
    int pi[3];
    if (!elem_slot(gd, i, j, k, pi)) return;
    if (out.bbox) {
        const double q0 = a0.pt[pi[0]], q1 = a1.pt[pi[1]], q2 = a2.pt[pi[2]];
        if (!(q0 >= out.bb[0] && q0 <= out.bb[1] && q1 >= out.bb[2] && q1 <= out.bb[3] && q2 >= out.bb[4] && q2 <= out.bb[5])) {
            dst[i + gd.sJ * j + gd.sK * k] = 0.;
            return;
        }
    }
    if (out.src_mat) {
        // the source element that holds the point: lo index on every axis, none when the point is outside the source axes
        const int e0 = a0.ilo[pi[0]], e1 = a1.ilo[pi[1]], e2 = a2.ilo[pi[2]];
        const bool inside = e0 != a0.ihi[pi[0]] && e1 != a1.ihi[pi[1]] && e2 != a2.ihi[pi[2]];
        if (!inside || out.src_mat[e0 * ss0 + e1 * ss1 + e2 * ss2] == PFEM_MAT_EXCLUDED) {
            dst[i + gd.sJ * j + gd.sK * k] = out.fill;
            return;
        }
    }
    const idx_t l0 = a0.ilo[pi[0]] * ss0, h0 = a0.ihi[pi[0]] * ss0;
    const idx_t l1 = a1.ilo[pi[1]] * ss1, h1 = a1.ihi[pi[1]] * ss1;
    const idx_t l2 = a2.ilo[pi[2]] * ss2, h2 = a2.ihi[pi[2]] * ss2;
    const double back = a0.lo[pi[0]], front = a0.hi[pi[0]], px = a0.pt[pi[0]];
    const double left = a1.lo[pi[1]], right = a1.hi[pi[1]], py = a1.pt[pi[1]];
    const double bottom = a2.lo[pi[2]], top = a2.hi[pi[2]], pz = a2.pt[pi[2]];
    const double dxh = __dsub_rn(front, px), dxl = __dsub_rn(px, back);
    const double dyh = __dsub_rn(right, py), dyl = __dsub_rn(py, left);
    const double wy = __dsub_rn(right, left), wx = __dsub_rn(front, back);
    auto plane = [&](idx_t z) {
        const double b = __dadd_rn(__dmul_rn(src[l0 + l1 + z], dxh), __dmul_rn(src[h0 + l1 + z], dxl));
        const double t = __dadd_rn(__dmul_rn(src[l0 + h1 + z], dxh), __dmul_rn(src[h0 + h1 + z], dxl));
        return __ddiv_rn(__ddiv_rn(__dadd_rn(__dmul_rn(b, dyh), __dmul_rn(t, dyl)), wy), wx);
    };
    const double lo = plane(l2), hi = plane(h2);
    const double w = __ddiv_rn(__dsub_rn(pz, bottom), __dsub_rn(top, bottom));
    dst[i + gd.sJ * j + gd.sK * k] = __dadd_rn(lo, __dmul_rn(w, __dsub_rn(hi, lo)));
}

// The same interpolation onto the tensor-product points of a target mesh given by its axes (providers on a foreign mesh:
// getTemperatures / getVoltage with INTERPOLATION_LINEAR, therm3d.cpp:387-395); out is dense with the target's own strides.
__global__ void k_interp_to_points(const int n0, const int n1, const int n2, const idx_t os0, const idx_t os1, const idx_t os2,
                                   const idx_t ss0, const idx_t ss1, const idx_t ss2, const double* __restrict__ src,
                                   const InterpAxis a0, const InterpAxis a1, const InterpAxis a2, double* __restrict__ out) {
    const idx_t total = (idx_t)n0 * n1 * n2;
    for (idx_t m = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; m < total; m += (idx_t)gridDim.x * blockDim.x) {
        const int i2 = (int)(m % n2), i1 = (int)((m / n2) % n1), i0 = (int)(m / ((idx_t)n2 * n1));
        const idx_t l0 = a0.ilo[i0] * ss0, h0 = a0.ihi[i0] * ss0;
        const idx_t l1 = a1.ilo[i1] * ss1, h1 = a1.ihi[i1] * ss1;
        const idx_t l2 = a2.ilo[i2] * ss2, h2 = a2.ihi[i2] * ss2;
        const double back = a0.lo[i0], front = a0.hi[i0], px = a0.pt[i0];
        const double left = a1.lo[i1], right = a1.hi[i1], py = a1.pt[i1];
        const double bottom = a2.lo[i2], top = a2.hi[i2], pz = a2.pt[i2];
        const double dxh = __dsub_rn(front, px), dxl = __dsub_rn(px, back);
        const double dyh = __dsub_rn(right, py), dyl = __dsub_rn(py, left);
        const double wy = __dsub_rn(right, left), wx = __dsub_rn(front, back);
        auto plane = [&](idx_t z) {
            const double b = __dadd_rn(__dmul_rn(src[l0 + l1 + z], dxh), __dmul_rn(src[h0 + l1 + z], dxl));
            const double t = __dadd_rn(__dmul_rn(src[l0 + h1 + z], dxh), __dmul_rn(src[h0 + h1 + z], dxl));
            return __ddiv_rn(__ddiv_rn(__dadd_rn(__dmul_rn(b, dyh), __dmul_rn(t, dyl)), wy), wx);
        };
        const double lo = plane(l2), hi = plane(h2);
        const double w = __ddiv_rn(__dsub_rn(pz, bottom), __dsub_rn(top, bottom));
        out[i0 * os0 + i1 * os1 + i2 * os2] = __dadd_rn(lo, __dmul_rn(w, __dsub_rn(hi, lo)));
    }
}

// out[m] = a[slot[m]]
__global__ void k_gather(const size_t n, const idx_t* __restrict__ slot, const double* __restrict__ a, double* __restrict__ out) {
    for (size_t m = blockIdx.x * (size_t)blockDim.x + threadIdx.x; m < n; m += (size_t)gridDim.x * blockDim.x) out[m] = a[slot[m]];
}

}  // namespace pfem
