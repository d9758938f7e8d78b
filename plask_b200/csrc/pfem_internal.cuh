// pfem_internal.cuh — device data layout, reduction helpers and the brick-element stencil core.
//
// Layout in HBM.  The mesh is handled in INDEX space: I = minor (fastest),
// J = medium, K = major axis of RectangularMesh<3> (plask/mesh/rectilinear3d.cpp:20-32).  A node
// array is the reference's DataVector<double> with the rows padded to a pitch sJ = nI rounded up
// to 16 doubles (128-byte rows: coalescing, and the 16-byte stride rule of TMA tensor maps):
// node (i,j,k) at i + sJ*(j + nJ*k); the pad entries are always 0.
// Element arrays live on the SAME lattice: element (i,j,k) is stored at the index of its
// lowest corner node; the slots with i >= nI-1, j = nJ-1 or k = nK-1 are padding and hold ZERO
// conductivity.  Every array additionally has a zero guard band of G >= sK + sJ + 2 entries
// on both sides.  Together this makes the 27-point operator branch-free: a neighbour that does
// not exist is reached through a padding/guard slot whose conductivity is 0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef PFEM_MAT_EXCLUDED
#define PFEM_MAT_EXCLUDED 0xFFFFFFFFu   /* include/plaskfem_cuda.h: element outside the masked mesh */
#endif

namespace pfem {

typedef long long idx_t;

struct Grid {
    int nI, nJ, nK;      // nodes per index-space axis
    idx_t sJ, sK;        // node strides (sI = 1): sJ = row pitch >= nI, sK = sJ*nJ
    idx_t N;             // nI*nJ*nK (true node count)
    idx_t NP;            // sK*nK (length of a pitched node array)
    idx_t G;             // guard band length (multiple of 16)
    int vdim;            // index-space axis (0 I, 1 J, 2 K) that is the physical vertical axis 2
    int dim_of_phys[3];  // index-space axis of physical axis a
    idx_t ps[3];         // node stride of physical axis a
    int pn[3];           // node count of physical axis a
    idx_t es[3];         // element stride of physical axis a in the ABI (compact) element order
    int abi_dim[3];      // index-space axis (0 I, 1 J, 2 K) of the ABI's minor, medium, major axis (identity unless the
                         // context was given another internal layout with pfem_set_layout)
    idx_t abi_ns[3];     // ABI node stride of index-space axis I, J, K
    idx_t E;             // compact element count
    int kown0, kown1;    // node planes [kown0, kown1) along K owned by this context (slab mode; else 0, nK)
    // spacings per index-space axis and their reciprocals; each array has one guard entry in
    // front and one behind (value 1), so h[-1] and h[n-1] are readable.
    const double *hI, *hJ, *hK, *rI, *rJ, *rK;
    // geometric spacings for gradients (currents, fluxes, heat): equal to h unless pfem_set_axis_weight made h = w * spacing and
    // r = w / spacing on one axis (cylindrical 2-D solvers: every element integral carries the factor r of its midpoint)
    const double *uI, *uJ, *uK;
};

// L2 prefetch of one lattice row of up to 512 doubles (lane l touches the 128-byte line l): the warp-per-row line kernels issue it for
// the row they will solve NEXT right after the loads of the current row, so that row's HBM latency passes under the current solve
// (ncu r02: 61 % of the stall samples of k_line_ml<FINE> were long-scoreboard waits on the first use of a freshly loaded row).
__device__ __forceinline__ void prefetch_row_l2(const double* row, const int lane, const int n) {
    if (lane * 16 < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + lane * 16));
}

// ---- slab mode (one context per GPU, peers mapped with CUDA IPC over NVLink) ----------------
#define PFEM_MAX_RANKS 8
#define PFEM_COMM_NV 8
#define PFEM_COMM_VEC 1024   // doubles per rank of the vector exchange (multilevel top level: one or two vertical lines of <= 512 nodes)
// One per rank, in memory the rank exports: peers deposit their partial sums here.
struct Inbox {
    double data[2][PFEM_MAX_RANKS][PFEM_COMM_NV];
    unsigned long long flag[2][PFEM_MAX_RANKS];
    double vec[2][PFEM_MAX_RANKS][PFEM_COMM_VEC];
};
struct Comm {
    int rank, nranks;
    unsigned long long seq;            // cross-rank exchanges completed so far (identical on all ranks)
    Inbox* inbox[PFEM_MAX_RANKS];      // inbox of every rank; inbox[rank] is local
    int timeout;                       // a wait gave up: a peer died or the call sequences diverged
};
// Destination planes in the neighbours' arrays for the boundary planes of r', q', p' (null = no neighbour).
struct PeerOut {
    double *r_lo, *q_lo, *p_lo;        // upper halo plane of the lower neighbour
    double *r_hi, *q_hi, *p_hi;        // lower halo plane of the upper neighbour
    double *z_lo, *z_hi;               // the same for z of the line-Jacobi iteration (kernels_line.cuh)
};

// Scalars of the PCG iteration and of the nonlinear loop; device resident, one instance per
// context, mirrored to pinned host memory after every graph batch.
struct Scalars {
    double rho;       // r.z
    double rho_prev;
    double pq;        // p.Ap
    double alpha;
    double beta;
    double rr;        // r.r
    double bb;        // ||b_free||^2
    double bz;        // b_free . D^-1 b_free  (preconditioner norm of the rhs)
    double zz;        // ||D^-1 r||^2  (Jacobi pseudo-residual, NSPCG's stopping quantity)
    double xx;        // ||x||^2
    double tol2;      // lin_tol^2 (negative: never converge, benchmark mode)
    double red[4];    // generic reduction outputs (err, max, ...)
    long long argidx; // arg-max index of the current reduction
    double maxcur[3];
    int iter;
    int maxit;
    int done;         // 1: all iteration kernels return immediately
    int status;       // 0 running, 1 converged, 2 maxit, -1 breakdown (p.Ap <= 0), -2 non-finite
    int bench;        // 1: ignore breakdown / convergence
    int neg_diag;     // a free row has a negative diagonal (NSPCG ier = -4)
    int launch;       // fused kernel: launches since the start of the solve (launch m has applied m-1 updates)
    int surf;         // 1: boundary-face matrix terms exist (convection); k_surf_iter completes q and finalises alpha/beta
    unsigned int ticket[8];
    Comm* comm;       // null unless the context is one slab of a multi-GPU solve
    double qz, qdq;   // q.z and q.D^-1 q of the stiffness part, handed from k_fpcg to k_surf_iter when surf != 0
    int line;         // 1: line-Jacobi PCG (kernels_line.cuh): k_fpcg only forms p', q', x' and alpha; 2: multilevel (kernels_ml.cuh)
    int pad2_;
    double ml_rho, ml_rr;   // multilevel preconditioner: r'.z summed over the levels so far, |r'|^2 of the fine level
};

// Boundary-face terms (conditions of the 2nd / 3rd kind and radiation, therm3d.cpp:140-168,242-268) flattened by the
// host into rows = mesh nodes that receive a load term or carry matrix entries.  All indices are lattice indices.
struct Surf {
    int nrows;                 // 0: no boundary terms
    int nnz;                   // matrix entries (both triangles), 0 unless convection is present
    const idx_t* node;         // [nrows] ascending
    const double* lconst;      // [nrows] temperature-independent load: heat flux + convection coeff*ambient
    const int* radptr;         // [nrows+1] radiation terms of the row
    const idx_t* rad_src;      // node whose temperature the term reads
    const double* rad_coef;    // 0.25e-12 * area * emissivity * sigma_SB
    const double* rad_amb4;    // ambient^4
    double* radv;              // [nrows] radiation load evaluated at the start of the current loop
    const int* kptr;           // [nrows+1] CSR of the matrix terms
    const idx_t* kcol;
    const double* kval;
};

// ------------------------------------------------------------------ reductions ----------

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide reduction of NV values; result valid in thread 0.  `sh` needs 32*NV doubles.
template <int NV, bool MAX>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* sh) {
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthr = blockDim.x * blockDim.y * blockDim.z;
    const int lane = tid & 31, warp = tid >> 5, nwarp = (nthr + 31) >> 5;
#pragma unroll
    for (int a = 0; a < NV; ++a) v[a] = MAX ? warp_max(v[a]) : warp_sum(v[a]);
    __syncthreads();  // protect sh from a previous use
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < NV; ++a) sh[a * 32 + warp] = v[a];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int a = 0; a < NV; ++a) {
            double x = (lane < nwarp) ? sh[a * 32 + lane] : (MAX ? -1.7976931348623157e308 : 0.);
            v[a] = MAX ? warp_max(x) : warp_sum(x);
        }
    }
}

// Grid-wide deterministic reduction ("last block done"): every block stores its NV partials,
// the block that draws the last ticket reduces all of them in a fixed order.  Returns true in
// ALL threads of that last block; totals valid in its thread 0.  partials: nblocks*NV doubles.
template <int NV, bool MAX>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double* __restrict__ partials, unsigned int* ticket,
                                            double* sh, int* sh_flag, bool peer_writes = false) {
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthr = blockDim.x * blockDim.y * blockDim.z;
    const unsigned int nblk = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    block_reduce<NV, MAX>(v, sh);
    if (tid == 0) {
#pragma unroll
        for (int a = 0; a < NV; ++a) partials[(size_t)bid * NV + a] = v[a];
        if (peer_writes) __threadfence_system();   // this CTA's stores into a neighbour's halo planes go first
        else __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        *sh_flag = (t == nblk - 1);
    }
    __syncthreads();
    if (!*sh_flag) return false;
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int a = 0; a < NV; ++a) acc[a] = MAX ? -1.7976931348623157e308 : 0.;
    for (unsigned int b = tid; b < nblk; b += nthr) {
#pragma unroll
        for (int a = 0; a < NV; ++a) {
            double x = __ldcg(partials + (size_t)b * NV + a);
            acc[a] = MAX ? fmax(acc[a], x) : acc[a] + x;
        }
    }
    block_reduce<NV, MAX>(acc, sh);
    if (tid == 0) {
#pragma unroll
        for (int a = 0; a < NV; ++a) v[a] = acc[a];
        *ticket = 0u;
    }
    return true;
}

// Cross-rank all-reduce of NV <= PFEM_COMM_NV doubles through the peers' inboxes; called by ALL threads of
// one block (the last block of a grid reduction), input/result in thread 0's v.  Every rank writes its values
// into every inbox (its own too), raises a sequence flag after a system-scope fence, waits for all flags in its
// own inbox and sums in rank order, so all ranks get bit-identical results.  Slots alternate with the sequence
// number; a rank can only be one exchange ahead of the slowest, so two slots suffice.  It is also the
// inter-iteration barrier that orders the halo-plane stores of k_fpcg (fence cumulativity).
template <int NV, bool MAX>
__device__ __forceinline__ void rank_allreduce(double (&v)[NV], Comm* cm, double* sh) {
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int a = 0; a < NV; ++a) sh[a] = v[a];
    }
    __syncthreads();
    const unsigned long long s = *(volatile unsigned long long*)&cm->seq;
    const int slot = (int)(s & 1ull), me = cm->rank;
    if (tid < cm->nranks) {
        Inbox* dst = cm->inbox[tid];
#pragma unroll
        for (int a = 0; a < NV; ++a) *(volatile double*)&dst->data[slot][me][a] = sh[a];
        __threadfence_system();
        *(volatile unsigned long long*)&dst->flag[slot][me] = s + 1ull;
        const Inbox* mine = cm->inbox[me];
        const long long t0 = clock64();
        while (*(volatile const unsigned long long*)&mine->flag[slot][tid] != s + 1ull) {
            // ~10 s; once one exchange has given up every later one falls through at once, so that the call ends with
            // PFEM_ERR_CUDA after seconds instead of 10 s per exchange of the remaining kernels
            if (*(volatile const int*)&cm->timeout || clock64() - t0 > 20000000000ll) { cm->timeout = 1; break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (tid == 0) {
        const Inbox* mine = cm->inbox[me];
#pragma unroll
        for (int a = 0; a < NV; ++a) {
            double acc = *(volatile const double*)&mine->data[slot][0][a];
            for (int r = 1; r < cm->nranks; ++r) {
                const double x = *(volatile const double*)&mine->data[slot][r][a];
                acc = MAX ? fmax(acc, x) : acc + x;
            }
            v[a] = acc;
        }
        *(volatile unsigned long long*)&cm->seq = s + 1ull;
    }
    __syncthreads();
}

// The same exchange carrying a vector of n <= PFEM_COMM_VEC doubles besides the NV scalars: all threads of the block store the vector
// into every rank's inbox first (the system-scope fence of the scalar exchange publishes those stores, cumulativity through the
// block barrier), then the scalar exchange runs, then every rank adds the vectors in rank order -> bit-identical sums everywhere.
// One vector exchange per sequence number, same slot as the scalars of that exchange.
template <int NV>
__device__ __forceinline__ void rank_allreduce_vec(double (&v)[NV], double* vec, const int n, Comm* cm, double* sh) {
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nt = blockDim.x * blockDim.y * blockDim.z;
    __syncthreads();
    const unsigned long long s = *(volatile unsigned long long*)&cm->seq;
    const int slot = (int)(s & 1ull), me = cm->rank, nr = cm->nranks;
    for (int r = 0; r < nr; ++r) {
        double* dst = cm->inbox[r]->vec[slot][me];
        for (int i = tid; i < n; i += nt) *(volatile double*)&dst[i] = vec[i];
    }
    rank_allreduce<NV, false>(v, cm, sh);   // starts and ends with a block barrier; thread 0 advances cm->seq
    const Inbox* mine = cm->inbox[me];
    for (int i = tid; i < n; i += nt) {
        double acc = *(volatile const double*)&mine->vec[slot][0][i];
        for (int r = 1; r < nr; ++r) acc += *(volatile const double*)&mine->vec[slot][r][i];
        vec[i] = acc;
    }
    __syncthreads();
}

// ------------------------------------------------------- brick element coefficients -----

// Directional element conductances of the 8-node brick (therm3d.cpp:215-220,
// electr3d.cpp:310-324): k_d = 1e-6 * c_d * h_a*h_b/h_d with c_d = c_vert on the physical
// vertical axis and c_lat on the two lateral ones.  ei,ej,ek may be -1 or n-1 (guards).
__device__ __forceinline__ void elem_conductances(const Grid& g, double cl, double cv, int ei, int ej, int ek,
                                                  double& kI, double& kJ, double& kK) {
    const double hi = g.hI[ei], hj = g.hJ[ej], hk = g.hK[ek];
    const double cI = (g.vdim == 0 ? cv : cl) * 1e-6;
    const double cJ = (g.vdim == 1 ? cv : cl) * 1e-6;
    const double cK = (g.vdim == 2 ? cv : cl) * 1e-6;
    kI = cI * (hj * hk) * g.rI[ei];
    kJ = cJ * (hi * hk) * g.rJ[ej];
    kK = cK * (hi * hj) * g.rK[ek];
}

// The 8 distinct entries of the element matrix indexed by the XOR of the local node numbers
// (bit 0 = I, bit 1 = J, bit 2 = K differ); therm3d.cpp:227-237 with x,y,z -> I,J,K.
__device__ __forceinline__ void elem_matrix8(double kI, double kJ, double kK, double (&kv)[8]) {
    const double s = kI + kJ + kK;
    kv[0] = s * (1. / 9.);
    kv[1] = (s - 3. * kI) * (1. / 18.);
    kv[2] = (s - 3. * kJ) * (1. / 18.);
    kv[4] = (s - 3. * kK) * (1. / 18.);
    kv[3] = (3. * kK - 2. * s) * (1. / 36.);  // I,J differ: (-2kI-2kJ+kK)/36
    kv[5] = (3. * kJ - 2. * s) * (1. / 36.);  // I,K differ
    kv[6] = (3. * kI - 2. * s) * (1. / 36.);  // J,K differ
    kv[7] = -s * (1. / 36.);
}

}  // namespace pfem
