// kernels_diffusion.cuh — device side of the Diffusion3D path (include/plaskdiff_cuda.h).
//
// One persistent cooperative kernel (`k_diff_compute`) runs the WHOLE loop of Diffusion3DSolver::compute
// (solvers/electrical/diffusion/diffusion3d.cpp:283-366) on the device: assemble the 12x12 Hermite element matrices and the load
// vector at the current U (setLocalMatrix + addLocalBurningMatrix, :196-204), build the 3x3 nodal blocks of the preconditioner, form
// err = 100 |K U - F| / |F| (:341-349), decide, and solve K U' = F (:358) by block-Jacobi PCG — grid barriers instead of launches,
// no host round trip between the loops.  The system is small (3 unknowns per lateral node, 10^4 .. 10^6 nodes) and lives in L2; what
// a launch-per-step design would pay is latency, which is why the loop is kept inside one kernel.
//
// Layout.  Lateral nodes on the lattice NL = n0*n1 in the ABI's own order (node strides s0, s1); an element sits at the lattice slot
// of its lowest-corner node (slots of the last row / column are padding, never active).  Vectors are SoA: v[c*NLp + node], c = 0
// value, 1 d/dy, 2 d/dx.  Element matrices: Ke[(12 r + c)*NLp + e] — consecutive threads = consecutive elements, every access
// coalesced, the tables of basis values are warp-uniform constant-memory reads.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pdiff {
namespace cg = cooperative_groups;

constexpr int NQ = 7, NQ2 = 49;

// Tables on the unit square at the 7x7 Gauss-Legendre points (filled by the host in init_tables, plaskdiff_cuda.cu):
// local function l = 3*node + comp, node n00, n01, n10, n11 (first digit = axis 0), comp 0 value, 1 d/dy, 2 d/dx
// (ElementParams3D, diffusion3d.hpp:108-141); physical function = scale * unit function with scale 1, Y, X.
__constant__ double c_phi[12][NQ2];   // unit-square functions at the points
__constant__ double c_w[NQ2];         // weights (sum = 1)
__constant__ double c_bl[4][NQ2];     // bilinear nodal functions at the points
__constant__ double c_kx[144];        // int d/dtx phi_r d/dtx phi_c   (times D Y/X in the element)
__constant__ double c_ky[144];        // int d/dty phi_r d/dty phi_c   (times D X/Y)

struct Problem {
    int n0, n1;                 // nodes along the physical axes 0, 1
    int s0, s1;                 // node strides of the axes in the lattice (one of them is 1)
    int NL, NLp;                // lattice size, padded to a multiple of 32
    const double *h0, *h1;      // element sizes along axis 0 [n0-1], axis 1 [n1-1]
    const uint8_t* eact;        // [NL] (+ guard bands of zeros on both sides)
    const uint8_t* nact;        // [NL] node is an unknown
    const double *A, *B, *C, *D;   // per element (lattice)
    const double* J;            // per node
    int nmodes;
    const double *P, *G, *dG;   // [m][NL][2]
    double* Ke;                 // [144][NLp]
    double* St;                 // [81][NLp] nodal stencil of K (see gather_nodes)
    double* Mi;                 // [6][NLp] inverse of the nodal 3x3 blocks (xx, xy, xz, yy, yz, zz)
    double *F, *U, *r, *z, *q, *p0, *p1;   // [3][NLp]
    double* part;               // [grid][4] block partial sums
};

struct Control {
    int loops;          // 0 = until converged
    double maxerr;      // [%]
    int maxit;
    double lin_tol;
    int verbatim;
    int loop_cap;       // safety net for loops = 0
    // results
    int loops_done, converged, status;   // status: 0 ok, 1 a linear solve hit maxit, 2 loop cap, -5 not SPD, -7 non-finite
    long long lin_iters;
    int last_iters;
    double lin_relres, lin_relres_precond, err;
    double err_log[64];
};

__device__ __forceinline__ void elem_coords(const Problem& P, int e, int& i0, int& i1) {
    if (P.s1 == 1) { i0 = e / P.n1; i1 = e - i0 * P.n1; } else { i1 = e / P.n0; i0 = e - i1 * P.n0; }
}

// Rows `row` of the element matrix of element e and its load entry: setLocalMatrix + addLocalBurningMatrix by quadrature.
__device__ __forceinline__ void assemble_row(const Problem& P, int e, int row, int verbatim, double* Krow, double& Frow) {
    int i0, i1;
    elem_coords(P, e, i0, i1);
    const double X = P.h0[i0], Y = P.h1[i1];
    const int nd[4] = {e, e + P.s1, e + P.s0, e + P.s0 + P.s1};
    double Uh[12], sc[12];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        sc[3 * n] = 1.; sc[3 * n + 1] = Y; sc[3 * n + 2] = X;
#pragma unroll
        for (int c = 0; c < 3; ++c) Uh[3 * n + c] = P.U[c * P.NLp + nd[n]] * sc[3 * n + c];
    }
    const double A = P.A[e], B = P.B[e], C = P.C[e], D = P.D[e];
    double Jn[4], Sn[4] = {0., 0., 0., 0.}, Tn[4] = {0., 0., 0., 0.};
#pragma unroll
    for (int n = 0; n < 4; ++n) Jn[n] = P.J[nd[n]];
    double ug = 0.;
    if (P.nmodes > 0) {
        // Ug (diffusion3d.cpp:296-300) is a property of the element, the nodal P.dG and P.G add up over the modes
        const double* u = Uh;   // scaled: slope unknowns already carry their own X or Y
        const double dy = u[1] - u[4] + u[7] - u[10], dx = u[2] + u[5] - u[8] - u[11];
        ug = verbatim ? 0.25 * (u[0] + u[3] + u[6] + u[9] + 0.25 * (X / Y * dy + Y / X * dx))
                      : 0.25 * (u[0] + u[3] + u[6] + u[9] + 0.25 * (dx + dy));
        for (int m = 0; m < P.nmodes; ++m) {
            const double2 g = reinterpret_cast<const double2*>(P.G)[(size_t)m * P.NL + e];
            const double2 dg = reinterpret_cast<const double2*>(P.dG)[(size_t)m * P.NL + e];
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const double2 p = reinterpret_cast<const double2*>(P.P)[(size_t)m * P.NL + nd[n]];
                Sn[n] += p.x * dg.x + p.y * dg.y;
                Tn[n] += p.x * g.x + p.y * g.y;
            }
        }
    }
    double acc[12], fa = 0.;
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = 0.;
    for (int q = 0; q < NQ2; ++q) {
        double u = 0.;
#pragma unroll
        for (int j = 0; j < 12; ++j) u = fma(Uh[j], c_phi[j][q], u);
        const double jb = Jn[0] * c_bl[0][q] + Jn[1] * c_bl[1][q] + Jn[2] * c_bl[2][q] + Jn[3] * c_bl[3][q];
        double c = A + 2. * B * u + 3. * C * u * u;
        double f = jb + B * u * u + 2. * C * u * u * u;
        if (P.nmodes > 0) {
            const double s = Sn[0] * c_bl[0][q] + Sn[1] * c_bl[1][q] + Sn[2] * c_bl[2][q] + Sn[3] * c_bl[3][q];
            const double t = Tn[0] * c_bl[0][q] + Tn[1] * c_bl[1][q] + Tn[2] * c_bl[2][q] + Tn[3] * c_bl[3][q];
            c += s;
            f += ug * s - t;
        }
        const double wp = c_w[q] * c_phi[row][q];
        const double t = wp * c;
        fa = fma(wp, f, fa);
#pragma unroll
        for (int j = 0; j < 12; ++j) acc[j] = fma(t, c_phi[j][q], acc[j]);
    }
    const int rc = row % 3;
    const double area = X * Y, dyx = D * Y / X, dxy = D * X / Y, sr = rc == 0 ? 1. : rc == 1 ? Y : X;
#pragma unroll
    for (int j = 0; j < 12; ++j) Krow[j] = sr * sc[j] * (area * acc[j] + dyx * c_kx[12 * row + j] + dxy * c_ky[12 * row + j]);
    Frow = sr * area * fa;
}

// all rows of all active elements -> Ke; items are (row, element) pairs, element fastest
__device__ void assemble_all(const Problem& P, int verbatim, double* Fe /* [12][NLp] or null */, int gtid, int gsize) {
    const int total = 12 * P.NLp;
    for (int item = gtid; item < total; item += gsize) {
        const int row = item / P.NLp, e = item - row * P.NLp;
        if (e >= P.NL || !P.eact[e]) continue;
        double Krow[12], Frow;
        assemble_row(P, e, row, verbatim, Krow, Frow);
#pragma unroll
        for (int j = 0; j < 12; ++j) P.Ke[(size_t)(12 * row + j) * P.NLp + e] = Krow[j];
        if (Fe) Fe[(size_t)row * P.NLp + e] = Frow;
    }
}

__device__ __forceinline__ int elem_off(const Problem& P, int l) { return (l & 1 ? P.s1 : 0) + (l & 2 ? P.s0 : 0); }

// Per node: the load vector, the 9 x (3x3) blocks of the node's rows of K (the nodal stencil St[(27 c + 3 nb + c2)*NLp + n],
// neighbour nb = 3 (d0 + 1) + (d1 + 1)) and the inverse of the diagonal block — gathered from the <= 4 adjacent elements, no atomics,
// bit-reproducible.  The iteration applies K from the stencil: 81 + 27 loads per node instead of 144 + 48 through the element
// matrices, 648 B per node instead of 1 152 B per element.  Nodes that are not unknowns get a unit row.
__device__ void gather_nodes(const Problem& P, const double* Fe, int gtid, int gsize) {
    for (int n = gtid; n < P.NL; n += gsize) {
        double f[3] = {0., 0., 0.};
        double m[6] = {0., 0., 0., 0., 0., 0.};
        const bool act = P.nact[n];
        bool ea[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) ea[l] = act && P.eact[n - elem_off(P, l)];
#pragma unroll
        for (int l = 0; l < 4; ++l)
            if (ea[l]) {
#pragma unroll
                for (int c = 0; c < 3; ++c) f[c] += Fe[(size_t)(3 * l + c) * P.NLp + n - elem_off(P, l)];
            }
#pragma unroll
        for (int nb = 0; nb < 9; ++nb) {
            const int d0 = nb / 3 - 1, d1 = nb % 3 - 1;
            double b[9] = {0., 0., 0., 0., 0., 0., 0., 0., 0.};
#pragma unroll
            for (int l = 0; l < 4; ++l) {      // the node is local node l of element n - off(l); the neighbour is its local node l2
                const int a0 = (l >> 1) + d0, a1 = (l & 1) + d1;
                if (a0 < 0 || a0 > 1 || a1 < 0 || a1 > 1) continue;
                const int l2 = 2 * a0 + a1;
                if (!ea[l]) continue;
                const double* Kb = P.Ke + (size_t)(12 * 3 * l + 3 * l2) * P.NLp + (n - elem_off(P, l));
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int c2 = 0; c2 < 3; ++c2) b[3 * c + c2] += Kb[(size_t)(12 * c + c2) * P.NLp];
            }
            if (nb == 4) {
                if (!act) b[0] = b[4] = b[8] = 1.;
                m[0] = b[0]; m[1] = b[1]; m[2] = b[2]; m[3] = b[4]; m[4] = b[5]; m[5] = b[8];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int c2 = 0; c2 < 3; ++c2) P.St[(size_t)(27 * c + 3 * nb + c2) * P.NLp + n] = b[3 * c + c2];
        }
        if (act) {   // inverse of the symmetric 3x3 diagonal block by cofactors
            const double c00 = m[3] * m[5] - m[4] * m[4], c01 = m[2] * m[4] - m[1] * m[5], c02 = m[1] * m[4] - m[2] * m[3];
            const double det = m[0] * c00 + m[1] * c01 + m[2] * c02, id = 1. / det;
            const double i0 = c00 * id, i1 = c01 * id, i2 = c02 * id;
            const double i3 = (m[0] * m[5] - m[2] * m[2]) * id, i4 = (m[1] * m[2] - m[0] * m[4]) * id, i5 = (m[0] * m[3] - m[1] * m[1]) * id;
            m[0] = i0; m[1] = i1; m[2] = i2; m[3] = i3; m[4] = i4; m[5] = i5;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) P.F[c * P.NLp + n] = f[c];
#pragma unroll
        for (int k = 0; k < 6; ++k) P.Mi[k * P.NLp + n] = m[k];
    }
}

// y = (K v)(node n) from the nodal stencil; `getv(node, c)` supplies v.  Stencil entries towards nodes that do not exist are zero; the
// index is clamped so that the (unused) load stays inside the vector.
template <class GetV> __device__ __forceinline__ void apply_node(const Problem& P, int n, GetV getv, double y[3]) {
    y[0] = y[1] = y[2] = 0.;
    const double* S = P.St + n;
#pragma unroll
    for (int nb = 0; nb < 9; ++nb) {
        const int m = min(max(n + (nb / 3 - 1) * P.s0 + (nb % 3 - 1) * P.s1, 0), P.NL - 1);
#pragma unroll
        for (int c2 = 0; c2 < 3; ++c2) {
            const double v = getv(m, c2);
            y[0] = fma(S[(size_t)(3 * nb + c2) * P.NLp], v, y[0]);
            y[1] = fma(S[(size_t)(27 + 3 * nb + c2) * P.NLp], v, y[1]);
            y[2] = fma(S[(size_t)(54 + 3 * nb + c2) * P.NLp], v, y[2]);
        }
    }
}

// deterministic grid reduction of NV values: block partials -> P.part, grid barrier, every block re-sums all partials.  The caller
// alternates between two halves of the partial buffer: a block can only write the partials of reduction k+2 after it has passed the
// barrier of reduction k+1, which every block reaches after its reads of reduction k.
template <int NV> __device__ __forceinline__ void grid_sum(const Problem& P, cg::grid_group& grid, double (&v)[NV], double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh[k * 32 + warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = lane < nw ? sh[k * 32 + lane] : 0.;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) P.part[(size_t)blockIdx.x * 4 + k] = x;
        }
    }
    grid.sync();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = 0.;
            for (int b = lane; b < (int)gridDim.x; b += 32) x += P.part[(size_t)b * 4 + k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) sh[k] = x;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = sh[k];
    __syncthreads();   // sh is reused by the next reduction
}

__global__ void __launch_bounds__(128, 4) k_diff_compute(Problem P, Control* ctl, double* Fe) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[4 * 32];
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    const int loops = ctl->loops, maxit = ctl->maxit, verbatim = ctl->verbatim, loop_cap = ctl->loop_cap;
    const double maxerr = ctl->maxerr, lin_tol = ctl->lin_tol;
    const int NLp = P.NLp;
    double* const part0 = P.part;
    int flip = 0;
    auto next_part = [&]() { P.part = part0 + (size_t)(flip ^= 1) * 4 * gridDim.x; };

    int loop = 0, status = 0, last_iters = 0, converged = 0;
    long long lin_iters = 0;
    double err = 0., relres = 0., relres_p = 0.;
    while (true) {
        // ---- K->clear(); F.fill(0.); setLocalMatrix / addLocalBurningMatrix for every element (diffusion3d.cpp:284-308)
        assemble_all(P, verbatim, Fe, gtid, gsize);
        grid.sync();
        gather_nodes(P, Fe, gtid, gsize);
        grid.sync();
        // ---- resid = K U - F, err = 100 sqrt(|resid|^2 / |F|^2)  (:341-349); r = -resid is the initial PCG residual, z = M^-1 r
        double s3[4] = {0., 0., 0., 0.};   // |r|^2, |F|^2, r.z, F.M^-1 F
        for (int n = gtid; n < P.NL; n += gsize) {
            double y[3];
            apply_node(P, n, [&](int nb, int c) { return P.U[c * NLp + nb]; }, y);
            const double r0 = P.F[n] - y[0], r1 = P.F[NLp + n] - y[1], r2 = P.F[2 * NLp + n] - y[2];
            const double f0 = P.F[n], f1 = P.F[NLp + n], f2 = P.F[2 * NLp + n];
            const double m0 = P.Mi[n], m1 = P.Mi[NLp + n], m2 = P.Mi[2 * NLp + n], m3 = P.Mi[3 * NLp + n], m4 = P.Mi[4 * NLp + n],
                         m5 = P.Mi[5 * NLp + n];
            const double z0 = m0 * r0 + m1 * r1 + m2 * r2, z1 = m1 * r0 + m3 * r1 + m4 * r2, z2 = m2 * r0 + m4 * r1 + m5 * r2;
            P.r[n] = r0; P.r[NLp + n] = r1; P.r[2 * NLp + n] = r2;
            P.z[n] = z0; P.z[NLp + n] = z1; P.z[2 * NLp + n] = z2;
            s3[0] += r0 * r0 + r1 * r1 + r2 * r2;
            s3[1] += f0 * f0 + f1 * f1 + f2 * f2;
            s3[2] += r0 * z0 + r1 * z1 + r2 * z2;
            s3[3] += f0 * (m0 * f0 + m1 * f1 + m2 * f2) + f1 * (m1 * f0 + m3 * f1 + m4 * f2) + f2 * (m2 * f0 + m4 * f1 + m5 * f2);
        }
        next_part();
        grid_sum<4>(P, grid, s3, sh);
        const double ff = s3[1], fz = s3[3];
        err = 100. * sqrt(s3[0] / ff);
        if (gtid == 0 && loop < 64) ctl->err_log[loop] = err;
        ++loop;
        if (!(err == err)) { status = -7; break; }
        if (err < maxerr) { converged = 1; break; }
        if (loops != 0 && loop >= loops) break;
        if (loop >= loop_cap) { status = 2; break; }   // loops = 0 and no convergence: the reference would spin for ever
        // the nodal blocks must be positive definite (they are principal blocks of K): F.M^-1 F > 0, r.M^-1 r >= 0
        if (!(s3[3] > 0.) || s3[2] < 0.) { status = -5; break; }

        // ---- K->solve(F, U) (:358): block-Jacobi PCG from the current U.  p' = z + beta p is formed on the fly for the neighbours
        // (double-buffered p), so an iteration needs two grid barriers, not three.
        // Stopping rule: r.M^-1 r <= lin_tol^2 F.M^-1 F.  During the first loops the unknowns (and with them the entries of F and
        // of the matrix) span many orders of magnitude across the region; the plain |r| / |F| would be satisfied while the part of
        // the region with small entries is still unsolved, and the Newton path would leave the one of the reference's direct solver.
        double rz = s3[2], beta = 0., rr = s3[0];
        int it = 0;
        const double tol2 = lin_tol * lin_tol * fz;
        while (rz > tol2 && it < maxit) {
            const double* pold = (it & 1) ? P.p1 : P.p0;
            double* pnew = (it & 1) ? P.p0 : P.p1;
            double s1[1] = {0.};
            for (int n = gtid; n < P.NL; n += gsize) {
                double y[3];
                auto getp = [&](int nb, int c) { return fma(beta, pold[c * NLp + nb], P.z[c * NLp + nb]); };
                apply_node(P, n, getp, y);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double pv = getp(n, c);
                    pnew[c * NLp + n] = pv;
                    P.q[c * NLp + n] = y[c];
                    s1[0] = fma(pv, y[c], s1[0]);
                }
            }
            next_part();
            grid_sum<1>(P, grid, s1, sh);
            const double pq = s1[0];
            if (!(pq > 0.)) { status = (pq == pq) ? -5 : -7; break; }
            const double alpha = rz / pq;
            double s2[2] = {0., 0.};
            for (int n = gtid; n < P.NL; n += gsize) {
                double r[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    P.U[c * NLp + n] = fma(alpha, pnew[c * NLp + n], P.U[c * NLp + n]);
                    r[c] = fma(-alpha, P.q[c * NLp + n], P.r[c * NLp + n]);
                    P.r[c * NLp + n] = r[c];
                }
                const double m0 = P.Mi[n], m1 = P.Mi[NLp + n], m2 = P.Mi[2 * NLp + n], m3 = P.Mi[3 * NLp + n], m4 = P.Mi[4 * NLp + n],
                             m5 = P.Mi[5 * NLp + n];
                const double z0 = m0 * r[0] + m1 * r[1] + m2 * r[2], z1 = m1 * r[0] + m3 * r[1] + m4 * r[2],
                             z2 = m2 * r[0] + m4 * r[1] + m5 * r[2];
                P.z[n] = z0; P.z[NLp + n] = z1; P.z[2 * NLp + n] = z2;
                s2[0] += r[0] * z0 + r[1] * z1 + r[2] * z2;
                s2[1] += r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
            }
            next_part();
            grid_sum<2>(P, grid, s2, sh);
            beta = s2[0] / rz;
            rz = s2[0];
            rr = s2[1];
            ++it;
        }
        lin_iters += it;
        last_iters = it;
        relres = sqrt(rr / ff);
        relres_p = sqrt(rz / fz);
        if (status < 0) break;
        if (it == 0) { status = 2; break; }   // the linear residual is below lin_tol although err >= maxerr: no further progress possible
        if (rz > tol2) status = 1;   // maxit reached: keep going like noconv = warning (iterative_matrix.hpp:299-314)
        grid.sync();                 // U complete before the next assembly reads the neighbours
    }
    if (gtid == 0) {
        ctl->loops_done = loop; ctl->converged = converged; ctl->status = status; ctl->lin_iters = lin_iters;
        ctl->last_iters = last_iters; ctl->lin_relres = relres; ctl->lin_relres_precond = relres_p; ctl->err = err;
    }
}

// ---- parity hooks: the same device functions, one plain launch each ---------------------------------------------------------
__global__ void k_diff_assemble(Problem P, int verbatim, double* Fe) {
    assemble_all(P, verbatim, Fe, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void k_diff_gather(Problem P, const double* Fe) {
    gather_nodes(P, Fe, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void k_diff_apply(Problem P, const double* v, double* y) {
    const int gsize = gridDim.x * blockDim.x;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < P.NL; n += gsize) {
        double t[3];
        apply_node(P, n, [&](int nb, int c) { return v[c * P.NLp + nb]; }, t);
#pragma unroll
        for (int c = 0; c < 3; ++c) y[c * P.NLp + n] = t[c];
    }
}

// outCarriersConcentration at lateral points (diffusion3d.cpp:420-474)
__global__ void k_diff_interp(Problem P, const double* ax0, const double* ax1, int npts, const double* x, const double* y, int method,
                              double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const double px = x[i], py = y[i];
    double res = 0.;
    if (px >= ax0[0] && px <= ax0[P.n0 - 1] && py >= ax1[0] && py <= ax1[P.n1 - 1]) {
        auto find = [](const double* a, int n, double v) {   // last i <= n-2 with a[i] <= v
            int lo = 0, hi = n - 1;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (a[mid] <= v) lo = mid; else hi = mid; }
            return lo;
        };
        const int i0 = find(ax0, P.n0, px), i1 = find(ax1, P.n1, py);
        const int e = i0 * P.s0 + i1 * P.s1;
        if (P.eact[e]) {
            const double X = P.h0[i0], Y = P.h1[i1];
            const double tx = (px - ax0[i0]) / X, ty = (py - ax1[i1]) / Y;
            const int nd[4] = {e, e + P.s1, e + P.s0, e + P.s0 + P.s1};
            if (method == 1) {
                res = (1 - tx) * ((1 - ty) * P.U[nd[0]] + ty * P.U[nd[1]]) + tx * ((1 - ty) * P.U[nd[2]] + ty * P.U[nd[3]]);
            } else {
                const double hx[4] = {1 - 3 * tx * tx + 2 * tx * tx * tx, X * tx * (1 - tx) * (1 - tx), 3 * tx * tx - 2 * tx * tx * tx,
                                      X * tx * tx * (tx - 1)};
                const double hy[4] = {1 - 3 * ty * ty + 2 * ty * ty * ty, Y * ty * (1 - ty) * (1 - ty), 3 * ty * ty - 2 * ty * ty * ty,
                                      Y * ty * ty * (ty - 1)};
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    const int ax = (n >> 1) * 2, ay = (n & 1) * 2;
                    res += P.U[nd[n]] * hx[ax] * hy[ay] + P.U[P.NLp + nd[n]] * hx[ax] * hy[ay + 1] + P.U[2 * P.NLp + nd[n]] * hx[ax + 1] * hy[ay];
                }
            }
        }
    }
    out[i] = res;
}

}  // namespace pdiff
