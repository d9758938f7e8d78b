// plaskfem_cuda.cu — C ABI (include/plaskfem_cuda.h) and host-side drivers of the device path:
// PCG loop (CUDA-graph batches, scalars device resident), the nonlinear loops of
// ThermalFem3DSolver::compute (solvers/thermal/static/therm3d.cpp:281-340) and
// ElectricalFem3DSolver::compute (solvers/electrical/shockley/electr3d.cpp:356-442).
// No CPU fallback anywhere: every numeric step is a kernel of this library.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/plaskfem_cuda.h"
#include "kernels_simple.cuh"
#include "kernels_tiled.cuh"
#include "kernels_tma.cuh"
#include "kernels_fused.cuh"
#include "kernels_surface.cuh"
#include "kernels_line.cuh"
#include "kernels_ml.cuh"

using namespace pfem;

// ------------------------------------------------------------------------ context -------

struct DevArr {
    void* base = nullptr;  // allocation start
    size_t bytes = 0;
};

struct pfem_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    bool have_mesh = false, have_materials = false, have_junctions = false, conds_valid = false;
    bool tables_iso = false;   // tab_lat == tab_vert entry by entry
    bool cond_iso = false;     // c_lat == c_vert in every element right now: the iteration kernels skip the c_vert stream
    Grid g;
    std::vector<DevArr> allocs;
    // node arrays (pointers already offset by the guard band)
    double *x = nullptr, *xprev = nullptr, *r = nullptr, *r2 = nullptr, *p = nullptr, *p2 = nullptr, *q = nullptr, *q2 = nullptr,
           *dinv = nullptr, *f = nullptr;
    uint8_t* fixed = nullptr;
    idx_t* bc_node = nullptr;   // de-duplicated Dirichlet nodes (device lattice indices) and values
    double* bc_val = nullptr;
    double* bc_val_first = nullptr;   // value of the FIRST condition naming each node; null unless some node is named with different values
    size_t nbc = 0;
    // element arrays on the node lattice
    double *cl = nullptr, *cv = nullptr, *Te = nullptr, *cur0 = nullptr, *cur1 = nullptr, *cur2 = nullptr;
    double *aux0 = nullptr, *aux1 = nullptr, *aux2 = nullptr;
    uint32_t *mat = nullptr, *junc = nullptr;
    uint8_t *role = nullptr, *noheat = nullptr;
    uint8_t* inactive = nullptr;   // masked mesh: nodes that touch no kept element (null unless elements are excluded)
    std::vector<uint8_t> h_excluded;   // host copy: element (ABI order) is outside the masked mesh (empty unless elements are excluded)
    bool has_excluded = false, source_set = false;
    // small arrays
    double* hbuf = nullptr;
    void* itab = nullptr;                // interpolation tables of the field exchange / foreign-mesh provider (grow-only, reused)
    size_t itab_bytes = 0;
    std::vector<double> h_host;          // host copy of hbuf: per index-space axis h, r (weighted) and u (geometric spacing)
    size_t h_off[3] = {0, 0, 0}, r_off[3] = {0, 0, 0}, u_off[3] = {0, 0, 0};
    bool weighted = false;               // pfem_set_axis_weight is in force (2-D cylindrical embedding)
    double *tab_lat = nullptr, *tab_vert = nullptr;
    uint32_t nmat = 0, nT = 0;
    double T0 = 0, dT = 1;
    JunctionDev* act = nullptr;
    int nact = 0;
    size_t ncol = 0;
    double *junc_cond = nullptr, *beta_col = nullptr, *js_col = nullptr;
    double pcond = 5., ncond = 50.;
    int stable = 0;
    int loopno = 0;
    // scalars / reductions
    Scalars* d_sc = nullptr;
    Scalars* h_sc = nullptr;  // pinned
    double* partials = nullptr;
    long long* partial_idx = nullptr;
    size_t n_partials = 0;
    // scratch for compact <-> lattice transfers
    void* stage = nullptr;
    size_t stage_bytes = 0;
    // cached CUDA graph of `graph_batch` PCG iterations
    cudaGraphExec_t graph = nullptr;
    int graph_batch = 0, graph_variant = -1, graph_precond = -1, graph_surf = -1, graph_iso = -1;
    // boundary-face terms (2nd / 3rd kind, radiation): flattened rows on the device, effective load vector
    Surf surf = {};
    double* fS = nullptr;
    int surf_iter = 0;          // 1: k_surf_iter follows k_fpcg (convection terms exist on some rank)
    bool surf_iter_known = false;
    // line-Jacobi preconditioner (pfem_opts::precond = 1): z, free-row mask, L D L^T factors of the vertical line blocks,
    // a tiny zero buffer whose tensor map stands in for q (every box is out of range -> zero fill, no traffic)
    double *lz = nullptr, *lmask = nullptr, *ll = nullptr, *ld = nullptr, *zero1 = nullptr;
    FusedPlan line_plan;
    int precond = 0;            // preconditioner of the PCG state prepared last
    // multilevel line preconditioner (pfem_opts::precond = 2, kernels_ml.cuh): level arrays, set-up partial sums
    MLDev ml = {};
    double* mlS = nullptr;
    idx_t ml_slen = 0;
    // internal layout (pfem_set_layout): 0 = the ABI's iteration order, 1 = vertical axis as the minor (I) axis
    int layout = 0;
    bool permuted = false;      // the lattice order differs from the ABI order: transfers go through permuting kernels
    size_t abi_n[3] = {0, 0, 0};   // node counts of the ABI's minor, medium, major axis
    long long launches = 0;
    double last_relres_pre = 0.;
    int sm_count = 148;
    std::vector<double> hax[3];   // host copy of the physical axes (interpolation tables of the field exchange)
    bool noheat_set = false;
    TiledPlan plan;
    TmaPlan tma;
    FusedPlan fused;
    // Dynamic3D (pfem_solve_dynamic): heat-capacity table, operator extension of the current time step (null outside a call),
    // scratch arrays allocated on first use
    double* tab_cprho = nullptr;
    uint32_t cap_nmat = 0, cap_nT = 0;
    const double *op_mass = nullptr, *op_cmass = nullptr;
    double *dyn_mass = nullptr, *dyn_ce = nullptr, *dyn_cl = nullptr, *dyn_cv = nullptr, *dyn_f = nullptr;
    int graph_mass = -1;
    // slab mode (multi-GPU): this context owns node planes [kown0, kown1) of its local mesh, the rest are halo planes
    int rank = 0, nranks = 1;
    Comm* d_comm = nullptr;
    Inbox* inbox = nullptr;                 // exported
    Inbox* peer_inbox[PFEM_MAX_RANKS] = {};  // imported (null for self)
    struct Neighbour { bool present = false; double* arr[9] = {}; long long G = 0, sK = 0, nK = 0; } nb_lo, nb_hi;
};

// arrays a neighbour may write / read: indices into pfem_ctx (order fixed by the blob layout)
enum { SA_R = 0, SA_R2, SA_Q, SA_Q2, SA_P, SA_P2, SA_DINV, SA_X, SA_LZ, SA_COUNT };
static_assert(SA_COUNT == 9, "pfem_ctx::Neighbour::arr holds SA_COUNT pointers");
static double* slab_array(pfem_ctx* ctx, int a) {
    switch (a) {
        case SA_R: return ctx->r; case SA_R2: return ctx->r2; case SA_Q: return ctx->q; case SA_Q2: return ctx->q2;
        case SA_P: return ctx->p; case SA_P2: return ctx->p2; case SA_DINV: return ctx->dinv; case SA_LZ: return ctx->lz; default: return ctx->x;
    }
}
struct SlabBlob {
    cudaIpcMemHandle_t arr[SA_COUNT];
    cudaIpcMemHandle_t inbox;
    long long G, sK, nK;
    int rank, nranks;
};

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            char b_[512];                                                                                \
            snprintf(b_, sizeof b_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            ctx->err = b_;                                                                               \
            return (e_ == cudaErrorMemoryAllocation) ? PFEM_ERR_NOMEM : PFEM_ERR_CUDA;                   \
        }                                                                                                \
    } while (0)

#define FAIL(code, ...)                          \
    do {                                         \
        char b_[512];                            \
        snprintf(b_, sizeof b_, __VA_ARGS__);    \
        ctx->err = b_;                           \
        return (code);                           \
    } while (0)

#define TRY(...)                   \
    do {                           \
        int rc_ = (__VA_ARGS__);   \
        if (rc_ < 0) return rc_;   \
    } while (0)

static void slab_release(pfem_ctx* ctx) {
    for (int r = 0; r < PFEM_MAX_RANKS; ++r)
        if (ctx->peer_inbox[r]) { cudaIpcCloseMemHandle(ctx->peer_inbox[r]); ctx->peer_inbox[r] = nullptr; }
    for (pfem_ctx::Neighbour* nb : {&ctx->nb_lo, &ctx->nb_hi}) {
        if (nb->present)
            for (int a = 0; a < SA_COUNT; ++a)
                if (nb->arr[a]) cudaIpcCloseMemHandle(nb->arr[a] - nb->G);
        *nb = pfem_ctx::Neighbour();
    }
    if (ctx->inbox) { cudaFree(ctx->inbox); ctx->inbox = nullptr; }
    if (ctx->d_comm) { cudaFree(ctx->d_comm); ctx->d_comm = nullptr; }
    ctx->rank = 0; ctx->nranks = 1;
}

static void free_all(pfem_ctx* ctx) {
    if (ctx->graph) { cudaGraphExecDestroy(ctx->graph); ctx->graph = nullptr; }
    slab_release(ctx);
    for (auto& a : ctx->allocs) cudaFree(a.base);
    ctx->allocs.clear();
    // every pointer below came from dev_alloc: a later pfem_set_mesh must not see stale addresses
    ctx->x = ctx->xprev = ctx->r = ctx->r2 = ctx->p = ctx->p2 = ctx->q = ctx->q2 = ctx->dinv = ctx->f = nullptr;
    ctx->fixed = nullptr; ctx->bc_node = nullptr; ctx->bc_val = nullptr; ctx->bc_val_first = nullptr; ctx->nbc = 0;
    ctx->cl = ctx->cv = ctx->Te = ctx->cur0 = ctx->cur1 = ctx->cur2 = ctx->aux0 = ctx->aux1 = ctx->aux2 = nullptr;
    ctx->mat = ctx->junc = nullptr; ctx->role = ctx->noheat = nullptr;
    ctx->inactive = nullptr; ctx->has_excluded = false; ctx->source_set = false; ctx->h_excluded.clear();
    ctx->hbuf = ctx->tab_lat = ctx->tab_vert = nullptr;
    ctx->act = nullptr; ctx->nact = 0; ctx->ncol = 0;
    ctx->junc_cond = ctx->beta_col = ctx->js_col = nullptr;
    ctx->partials = nullptr; ctx->partial_idx = nullptr; ctx->n_partials = 0;
    memset(&ctx->surf, 0, sizeof ctx->surf); ctx->fS = nullptr; ctx->surf_iter = 0; ctx->surf_iter_known = false;
    ctx->lz = ctx->lmask = ctx->ll = ctx->ld = ctx->zero1 = nullptr; ctx->line_plan.valid = false; ctx->precond = 0;
    memset(&ctx->ml, 0, sizeof ctx->ml); ctx->mlS = nullptr; ctx->ml_slen = 0;
    if (ctx->stage) { cudaFree(ctx->stage); ctx->stage = nullptr; ctx->stage_bytes = 0; }
    if (ctx->itab) { cudaFree(ctx->itab); ctx->itab = nullptr; ctx->itab_bytes = 0; }
    ctx->have_mesh = ctx->have_materials = ctx->have_junctions = ctx->conds_valid = false;
    ctx->noheat_set = false;
    ctx->tab_cprho = nullptr; ctx->cap_nmat = ctx->cap_nT = 0; ctx->op_mass = ctx->op_cmass = nullptr;
    ctx->dyn_mass = ctx->dyn_ce = ctx->dyn_cl = ctx->dyn_cv = ctx->dyn_f = nullptr; ctx->graph_mass = -1;
}

template <typename T>
static int dev_alloc(pfem_ctx* ctx, T** out, size_t count, size_t guard) {
    void* base = nullptr;
    size_t bytes = (count + 2 * guard) * sizeof(T);
    if (bytes == 0) bytes = sizeof(T);
    CU(cudaMalloc(&base, bytes));
    CU(cudaMemsetAsync(base, 0, bytes, ctx->stream));
    ctx->allocs.push_back({base, bytes});
    *out = reinterpret_cast<T*>(base) + guard;
    return PFEM_OK;
}

// release one array obtained from dev_alloc with guard 0 (re-set Dirichlet lists, tables ...)
template <typename T>
static void dev_release(pfem_ctx* ctx, T** p) {
    if (!*p) return;
    for (size_t a = 0; a < ctx->allocs.size(); ++a)
        if (ctx->allocs[a].base == (void*)*p) {
            cudaStreamSynchronize(ctx->stream);
            cudaFree(ctx->allocs[a].base);
            ctx->allocs.erase(ctx->allocs.begin() + a);
            break;
        }
    *p = nullptr;
}

// line-Jacobi preconditioner: z, free-row mask and the L D L^T factors (allocated on first use; in slab mode before the export)
static int alloc_line_arrays(pfem_ctx* ctx) {
    if (ctx->lz) return PFEM_OK;
    const size_t N = (size_t)ctx->g.NP, G = (size_t)ctx->g.G;
    TRY(dev_alloc(ctx, &ctx->lz, N, G));
    TRY(dev_alloc(ctx, &ctx->lmask, N, G));
    TRY(dev_alloc(ctx, &ctx->ll, N, G));
    TRY(dev_alloc(ctx, &ctx->ld, N, G));
    return PFEM_OK;
}

// grow-only device buffer for the interpolation tables: the meta loop exchanges fields every iteration and must not pay a
// cudaMalloc / cudaFree (a device-wide synchronisation) each time
static int ensure_itab(pfem_ctx* ctx, size_t bytes) {
    if (ctx->itab_bytes >= bytes) return PFEM_OK;
    if (ctx->itab) { CU(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->itab); ctx->itab = nullptr; ctx->itab_bytes = 0; }
    const size_t cap = (bytes + 4095) / 4096 * 4096;
    CU(cudaMalloc(&ctx->itab, cap));
    ctx->itab_bytes = cap;
    return PFEM_OK;
}

static int ensure_stage(pfem_ctx* ctx, size_t bytes) {
    if (ctx->stage_bytes >= bytes) return PFEM_OK;
    if (ctx->stage) { CU(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->stage); ctx->stage = nullptr; ctx->stage_bytes = 0; }
    CU(cudaMalloc(&ctx->stage, bytes));
    ctx->stage_bytes = bytes;
    return PFEM_OK;
}

static inline dim3 node_block() { return dim3(PFEM_NODE_BLOCK_X, PFEM_NODE_BLOCK_Y, 1); }
static inline dim3 node_grid(const Grid& g) {
    return dim3((g.nI + PFEM_NODE_BLOCK_X - 1) / PFEM_NODE_BLOCK_X, (g.nJ + PFEM_NODE_BLOCK_Y - 1) / PFEM_NODE_BLOCK_Y, g.nK);
}
static inline int vec_blocks(const pfem_ctx* ctx) { return ctx->sm_count * 8; }

// host <-> device copies of node arrays: the ABI side is dense (row length nI), the device side pitched
static int ensure_stage(pfem_ctx* ctx, size_t bytes);
static cudaError_t upload_nodes(pfem_ctx* ctx, double* dev, const double* host) {
    const Grid& g = ctx->g;
    if (!ctx->permuted)
        return cudaMemcpy2DAsync(dev, (size_t)g.sJ * 8, host, (size_t)g.nI * 8, (size_t)g.nI * 8, (size_t)g.nJ * g.nK,
                                 cudaMemcpyHostToDevice, ctx->stream);
    if (ensure_stage(ctx, (size_t)g.N * 8) < 0) return cudaErrorMemoryAllocation;
    cudaError_t e = cudaMemcpyAsync(ctx->stage, host, (size_t)g.N * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) return e;
    k_node_expand<<<dim3((g.nI + PFEM_NODE_BLOCK_X - 1) / PFEM_NODE_BLOCK_X, (g.nJ + PFEM_NODE_BLOCK_Y - 1) / PFEM_NODE_BLOCK_Y, g.nK),
                    dim3(PFEM_NODE_BLOCK_X, PFEM_NODE_BLOCK_Y, 1), 0, ctx->stream>>>(g, (const double*)ctx->stage, dev);
    ++ctx->launches;
    return cudaGetLastError();
}
static cudaError_t download_nodes(pfem_ctx* ctx, double* host, const double* dev) {
    const Grid& g = ctx->g;
    if (!ctx->permuted)
        return cudaMemcpy2DAsync(host, (size_t)g.nI * 8, dev, (size_t)g.sJ * 8, (size_t)g.nI * 8, (size_t)g.nJ * g.nK,
                                 cudaMemcpyDeviceToHost, ctx->stream);
    if (ensure_stage(ctx, (size_t)g.N * 8) < 0) return cudaErrorMemoryAllocation;
    k_node_compact<<<dim3((g.nI + PFEM_NODE_BLOCK_X - 1) / PFEM_NODE_BLOCK_X, (g.nJ + PFEM_NODE_BLOCK_Y - 1) / PFEM_NODE_BLOCK_Y, g.nK),
                     dim3(PFEM_NODE_BLOCK_X, PFEM_NODE_BLOCK_Y, 1), 0, ctx->stream>>>(g, dev, (double*)ctx->stage);
    ++ctx->launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(host, ctx->stage, (size_t)g.N * 8, cudaMemcpyDeviceToHost, ctx->stream);
}
// ABI node number -> pitched lattice index
static inline idx_t abi_to_lattice(const pfem_ctx* ctx, idx_t r) {
    const Grid& g = ctx->g;
    const idx_t st[3] = {1, g.sJ, g.sK};
    const idx_t m0 = r % (idx_t)ctx->abi_n[0], t = r / (idx_t)ctx->abi_n[0];
    const idx_t m1 = t % (idx_t)ctx->abi_n[1], m2 = t / (idx_t)ctx->abi_n[1];
    return m0 * st[g.abi_dim[0]] + m1 * st[g.abi_dim[1]] + m2 * st[g.abi_dim[2]];
}

#define LAUNCHED(n) (ctx->launches += (n))
#define KCHECK() CU(cudaGetLastError())

// ------------------------------------------------------------------------ life cycle ----

extern "C" int pfem_abi_version(void) { return PFEM_ABI_VERSION; }

extern "C" int pfem_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char* pfem_strerror(int s) {
    switch (s) {
        case PFEM_OK: return "ok";
        case PFEM_NOT_CONVERGED: return "linear solver failed to converge in maxit iterations";
        case PFEM_ERR_CUDA: return "CUDA runtime error";
        case PFEM_ERR_NO_DEVICE: return "no usable CUDA device (the CUDA algorithm has no CPU fallback)";
        case PFEM_ERR_BAD_INPUT: return "bad input";
        case PFEM_ERR_STATE: return "call sequence error: mesh/materials/field not set";
        case PFEM_ERR_NOT_SPD: return "stiffness matrix is not positive definite (p.Ap <= 0)";
        case PFEM_ERR_NOMEM: return "out of device memory";
        case PFEM_ERR_NAN: return "non-finite value in the iteration";
        default: return "unknown status";
    }
}

extern "C" const char* pfem_last_error(const pfem_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int pfem_create(pfem_ctx** out, int device) {
    if (!out) return PFEM_ERR_BAD_INPUT;
    *out = nullptr;
    int n = pfem_device_count();
    if (n <= 0 || device < 0 || device >= n) return PFEM_ERR_NO_DEVICE;
    pfem_ctx* ctx = new pfem_ctx();
    ctx->device = device;
    auto bail = [&](int rc) { delete ctx; return rc; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(PFEM_ERR_NO_DEVICE);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(PFEM_ERR_NO_DEVICE);
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(PFEM_ERR_CUDA);
    if (cudaMalloc(&ctx->d_sc, sizeof(Scalars)) != cudaSuccess) return bail(PFEM_ERR_NOMEM);
    cudaMemset(ctx->d_sc, 0, sizeof(Scalars));
    if (cudaMallocHost(&ctx->h_sc, sizeof(Scalars)) != cudaSuccess) return bail(PFEM_ERR_NOMEM);
    memset(ctx->h_sc, 0, sizeof(Scalars));
    *out = ctx;
    return PFEM_OK;
}

extern "C" void pfem_destroy(pfem_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    free_all(ctx);
    if (ctx->d_sc) cudaFree(ctx->d_sc);
    if (ctx->h_sc) cudaFreeHost(ctx->h_sc);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int pfem_set_layout(pfem_ctx* ctx, int layout) {
    if (!ctx) return PFEM_ERR_BAD_INPUT;
    if (layout != PFEM_LAYOUT_ABI && layout != PFEM_LAYOUT_VERTICAL_MINOR) FAIL(PFEM_ERR_BAD_INPUT, "unknown layout %d", layout);
    if (ctx->have_mesh && layout != ctx->layout) FAIL(PFEM_ERR_STATE, "pfem_set_layout must be called before pfem_set_mesh");
    ctx->layout = layout;
    return PFEM_OK;
}

extern "C" void pfem_default_opts(pfem_opts* o) {
    memset(o, 0, sizeof(*o));
    o->maxit = 10000;
    o->lin_tol = 1e-8;
    o->precond = 0;
    o->outer_tol = 0.05;
    o->loops = 0;
    o->batch = 0;
    o->variant = 3;
}

// ------------------------------------------------------------------------ mesh ----------

extern "C" int pfem_set_mesh(pfem_ctx* ctx, const size_t n[3], const double* ax0, const double* ax1, const double* ax2,
                             const size_t stride[3]) {
    if (!ctx) return PFEM_ERR_BAD_INPUT;
    CU(cudaSetDevice(ctx->device));
    if (!n || !ax0 || !ax1 || !ax2 || !stride) FAIL(PFEM_ERR_BAD_INPUT, "null mesh argument");
    for (int a = 0; a < 3; ++a)
        if (n[a] < 2) FAIL(PFEM_ERR_BAD_INPUT, "axis %d needs at least 2 points", a);
    // recover the iteration order from the strides (rectilinear3d.cpp:20-32)
    int minor = -1, medium = -1, major = -1;
    for (int a = 0; a < 3; ++a) if (stride[a] == 1) minor = a;
    if (minor < 0) FAIL(PFEM_ERR_BAD_INPUT, "no axis has stride 1");
    for (int a = 0; a < 3; ++a) if (a != minor && stride[a] == n[minor]) medium = a;
    if (medium < 0) FAIL(PFEM_ERR_BAD_INPUT, "strides are not a RectangularMesh<3> iteration order");
    major = 3 - minor - medium;
    if (stride[major] != n[minor] * n[medium]) FAIL(PFEM_ERR_BAD_INPUT, "strides are not a RectangularMesh<3> iteration order");
    const double* ax[3] = {ax0, ax1, ax2};
    for (int a = 0; a < 3; ++a)
        for (size_t i = 1; i < n[a]; ++i)
            if (!(ax[a][i] > ax[a][i - 1])) FAIL(PFEM_ERR_BAD_INPUT, "axis %d is not strictly increasing at %zu", a, i);
    double total = (double)n[0] * (double)n[1] * (double)n[2];
    if (total > 2.0e9) FAIL(PFEM_ERR_BAD_INPUT, "mesh too large for one device context (%g nodes)", total);

    CU(cudaStreamSynchronize(ctx->stream));
    free_all(ctx);
    Grid& g = ctx->g;
    memset(&g, 0, sizeof(g));
    for (int a = 0; a < 3; ++a) ctx->hax[a].assign(ax[a], ax[a] + n[a]);
    ctx->noheat_set = false;
    // internal layout: I, J, K = the ABI's minor, medium, major axis, or (pfem_set_layout) the vertical axis as I and the two
    // lateral axes in their ABI order behind it
    int im = minor, jm = medium, km = major;
    if (ctx->layout == 1 && minor != 2) {
        im = 2;
        jm = (medium == 2) ? minor : (minor == 2 ? medium : minor);
        km = 3 - im - jm;
        if (stride[jm] > stride[km]) { int t = jm; jm = km; km = t; }
    }
    ctx->permuted = !(im == minor && jm == medium && km == major);
    ctx->abi_n[0] = n[minor]; ctx->abi_n[1] = n[medium]; ctx->abi_n[2] = n[major];
    g.nI = (int)n[im]; g.nJ = (int)n[jm]; g.nK = (int)n[km];
    g.sJ = ((idx_t)g.nI + 15) / 16 * 16;   // 128-byte rows
    g.sK = g.sJ * g.nJ;
    g.N = (idx_t)g.nI * g.nJ * g.nK;
    g.NP = g.sK * g.nK;
    g.G = ((g.sK + g.sJ + 2 + 15) / 16) * 16;
    g.dim_of_phys[im] = 0; g.dim_of_phys[jm] = 1; g.dim_of_phys[km] = 2;
    g.abi_dim[0] = g.dim_of_phys[minor]; g.abi_dim[1] = g.dim_of_phys[medium]; g.abi_dim[2] = g.dim_of_phys[major];
    g.abi_ns[0] = (idx_t)stride[im]; g.abi_ns[1] = (idx_t)stride[jm]; g.abi_ns[2] = (idx_t)stride[km];
    g.vdim = g.dim_of_phys[2];
    for (int a = 0; a < 3; ++a) g.pn[a] = (int)n[a];
    g.ps[im] = 1; g.ps[jm] = g.sJ; g.ps[km] = g.sK;   // strides in the pitched device layout
    g.es[minor] = 1; g.es[medium] = (idx_t)n[minor] - 1; g.es[major] = (idx_t)(n[minor] - 1) * (idx_t)(n[medium] - 1);
    g.E = (idx_t)(g.nI - 1) * (g.nJ - 1) * (g.nK - 1);
    g.kown0 = 0; g.kown1 = g.nK;

    // spacing arrays with one guard entry (value 1) on both sides: h, r = 1/h and the geometric spacing u per index-space axis
    const int cnt[3] = {g.nI - 1, g.nJ - 1, g.nK - 1};
    const int phys_of_dim[3] = {im, jm, km};
    size_t tot = 0;
    for (int d = 0; d < 3; ++d) tot += 3 * (size_t)(cnt[d] + 2);
    std::vector<double>& hb = ctx->h_host;
    hb.assign(tot, 1.0);
    size_t off = 0;
    size_t* hoff = ctx->h_off; size_t* roff = ctx->r_off; size_t* uoff = ctx->u_off;
    for (int d = 0; d < 3; ++d) {
        hoff[d] = off + 1;
        for (int i = 0; i < cnt[d]; ++i) hb[hoff[d] + i] = ax[phys_of_dim[d]][i + 1] - ax[phys_of_dim[d]][i];
        off += cnt[d] + 2;
        roff[d] = off + 1;
        for (int i = 0; i < cnt[d]; ++i) hb[roff[d] + i] = 1.0 / hb[hoff[d] + i];
        off += cnt[d] + 2;
        uoff[d] = off + 1;
        for (int i = 0; i < cnt[d]; ++i) hb[uoff[d] + i] = hb[hoff[d] + i];
        off += cnt[d] + 2;
    }
    ctx->weighted = false;
    TRY(dev_alloc(ctx, &ctx->hbuf, tot, 0));
    CU(cudaMemcpyAsync(ctx->hbuf, hb.data(), tot * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    g.hI = ctx->hbuf + hoff[0]; g.rI = ctx->hbuf + roff[0]; g.uI = ctx->hbuf + uoff[0];
    g.hJ = ctx->hbuf + hoff[1]; g.rJ = ctx->hbuf + roff[1]; g.uJ = ctx->hbuf + uoff[1];
    g.hK = ctx->hbuf + hoff[2]; g.rK = ctx->hbuf + roff[2]; g.uK = ctx->hbuf + uoff[2];

    const size_t N = (size_t)g.NP, G = (size_t)g.G;
    TRY(dev_alloc(ctx, &ctx->x, N, G));
    TRY(dev_alloc(ctx, &ctx->xprev, N, G));
    TRY(dev_alloc(ctx, &ctx->r, N, G));
    TRY(dev_alloc(ctx, &ctx->p, N, G));
    TRY(dev_alloc(ctx, &ctx->p2, N, G));
    TRY(dev_alloc(ctx, &ctx->q, N, G));
    TRY(dev_alloc(ctx, &ctx->r2, N, G));
    TRY(dev_alloc(ctx, &ctx->q2, N, G));
    TRY(dev_alloc(ctx, &ctx->dinv, N, G));
    TRY(dev_alloc(ctx, &ctx->f, N, G));
    TRY(dev_alloc(ctx, &ctx->cl, N, G));
    TRY(dev_alloc(ctx, &ctx->cv, N, G));
    TRY(dev_alloc(ctx, &ctx->fixed, N, G));
    TRY(dev_alloc(ctx, &ctx->mat, N, G));
    // reduction partials: enough for the node-lattice grid and the persistent vector grid
    dim3 ng = node_grid(g);
    size_t nblk = (size_t)ng.x * ng.y * ng.z;
    size_t nb2 = (size_t)vec_blocks(ctx);
    ctx->n_partials = (nblk > nb2 ? nblk : nb2) + 1024;
    TRY(dev_alloc(ctx, &ctx->partials, ctx->n_partials * 4, 0));
    TRY(dev_alloc(ctx, &ctx->partial_idx, ctx->n_partials, 0));
    CU(cudaMemsetAsync(ctx->d_sc, 0, sizeof(Scalars), ctx->stream));
    ctx->loopno = 0;
    ctx->plan = make_tiled_plan(g, ctx->sm_count);
    ctx->tma = make_tma_plan(g, ctx->sm_count, ctx->p, ctx->p2, ctx->r, ctx->dinv, ctx->cl, ctx->cv);
    {
        double* const rr[2] = {ctx->r, ctx->r2};
        double* const qq[2] = {ctx->q, ctx->q2};
        double* const pp[2] = {ctx->p, ctx->p2};
        ctx->fused = make_fused_plan(g, ctx->sm_count, rr, qq, pp, ctx->dinv, ctx->cl, ctx->cv, ctx->x);
    }
    ctx->nbc = 0;
    ctx->bc_node = nullptr;
    ctx->bc_val = nullptr;
    ctx->bc_val_first = nullptr;
    ctx->have_mesh = true;
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

#define NEED_MESH() \
    do { if (!ctx) return PFEM_ERR_BAD_INPUT; if (!ctx->have_mesh) FAIL(PFEM_ERR_STATE, "pfem_set_mesh has not been called"); \
         CU(cudaSetDevice(ctx->device)); } while (0)

// upload a compact element array (NC interleaved components) onto the lattice
template <typename T, int NC>
static int upload_elem(pfem_ctx* ctx, const T* host, T* d0, T* d1, T* d2) {
    const Grid& g = ctx->g;
    size_t bytes = (size_t)g.E * NC * sizeof(T);
    TRY(ensure_stage(ctx, bytes));
    CU(cudaMemcpyAsync(ctx->stage, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    k_elem_expand<T, NC><<<vec_blocks(ctx), 256, 0, ctx->stream>>>(g, (const T*)ctx->stage, d0, d1, d2);
    KCHECK(); LAUNCHED(1);
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}
template <typename T, int NC>
static int download_elem(pfem_ctx* ctx, const T* s0, const T* s1, const T* s2, T* host) {
    const Grid& g = ctx->g;
    size_t bytes = (size_t)g.E * NC * sizeof(T);
    TRY(ensure_stage(ctx, bytes));
    k_elem_compact<T, NC><<<vec_blocks(ctx), 256, 0, ctx->stream>>>(g, s0, s1, s2, (T*)ctx->stage);
    KCHECK(); LAUNCHED(1);
    CU(cudaMemcpyAsync(host, ctx->stage, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

// Weight of the elements along one physical axis: every element integral of the operator, the load vector and the heat capacity is
// multiplied by w[index of the element along `axis`] — h := w * spacing, r := w / spacing in the spacing tables, while gradients
// keep the geometric spacing.  This is how the cylindrical 2-D solvers (therm2d.cpp:338-420: K_e and f_e times the midpoint
// radius r) run on the brick kernels: axis = the radial axis, w = element midpoint radii.  w == NULL removes the weights.
extern "C" int pfem_set_axis_weight(pfem_ctx* ctx, int axis, const double* w) {
    NEED_MESH();
    if (axis < 0 || axis > 2) FAIL(PFEM_ERR_BAD_INPUT, "axis must be 0, 1 or 2");
    if (ctx->surf.nrows) FAIL(PFEM_ERR_BAD_INPUT, "element weights and boundary conditions of the 2nd / 3rd kind cannot be combined");
    const Grid& g = ctx->g;
    const int d = g.dim_of_phys[axis];
    const int cnt = g.pn[axis] - 1;
    std::vector<double>& hb = ctx->h_host;
    for (int i = 0; i < cnt; ++i) {
        const double wi = w ? w[i] : 1.;
        if (!(wi > 0.) || !(wi < 1e300)) FAIL(PFEM_ERR_BAD_INPUT, "weight %d of axis %d is not a positive finite number", i, axis);
        const double u = hb[ctx->u_off[d] + i];
        hb[ctx->h_off[d] + i] = wi * u;
        hb[ctx->r_off[d] + i] = wi / u;
    }
    bool any = false;
    for (int dd = 0; dd < 3 && !any; ++dd) {
        const int c = (dd == 0 ? g.nI : dd == 1 ? g.nJ : g.nK) - 1;
        for (int i = 0; i < c && !any; ++i) any = hb[ctx->h_off[dd] + i] != hb[ctx->u_off[dd] + i];
    }
    ctx->weighted = any;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpyAsync(ctx->hbuf, hb.data(), hb.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

extern "C" int pfem_set_materials(pfem_ctx* ctx, const uint32_t* elem_mat, uint32_t nmat, double T0, double dT,
                                  uint32_t nT, const double* c_lat, const double* c_vert) {
    NEED_MESH();
    if (!elem_mat || !c_lat || !c_vert) FAIL(PFEM_ERR_BAD_INPUT, "null material argument");
    if (nmat == 0 || nT < 2 || !(dT > 0.)) FAIL(PFEM_ERR_BAD_INPUT, "need nmat >= 1, nT >= 2, dT > 0");
    bool excluded = false;
    for (idx_t e = 0; e < ctx->g.E; ++e) {
        if (elem_mat[e] == PFEM_MAT_EXCLUDED) { excluded = true; continue; }
        if (elem_mat[e] >= nmat) FAIL(PFEM_ERR_BAD_INPUT, "element %lld has material id %u >= nmat %u", (long long)e, elem_mat[e], nmat);
    }
    if (excluded && ctx->source_set)
        FAIL(PFEM_ERR_STATE, "elements are excluded: call pfem_set_materials before pfem_set_source");
    if (excluded != ctx->has_excluded && ctx->surf.nrows)
        FAIL(PFEM_ERR_STATE, "the masked mesh changed: call pfem_set_materials before pfem_set_boundary (boundary terms exist on kept elements only)");
    ctx->h_excluded.clear();
    if (excluded) {
        ctx->h_excluded.resize((size_t)ctx->g.E);
        for (idx_t e = 0; e < ctx->g.E; ++e) ctx->h_excluded[(size_t)e] = elem_mat[e] == PFEM_MAT_EXCLUDED;
    }
    TRY(upload_elem<uint32_t, 1>(ctx, elem_mat, ctx->mat, nullptr, nullptr));
    ctx->has_excluded = excluded;
    if (excluded) {   // masked mesh: mark the nodes that touch no kept element
        if (!ctx->inactive) TRY(dev_alloc(ctx, &ctx->inactive, (size_t)ctx->g.NP, (size_t)ctx->g.G));
        k_mark_inactive<<<node_grid(ctx->g), node_block(), 0, ctx->stream>>>(ctx->g, ctx->mat, ctx->inactive);
        KCHECK(); LAUNCHED(1);
    }
    size_t cnt = (size_t)nmat * nT;
    dev_release(ctx, &ctx->tab_lat);
    dev_release(ctx, &ctx->tab_vert);
    TRY(dev_alloc(ctx, &ctx->tab_lat, cnt, 0));
    TRY(dev_alloc(ctx, &ctx->tab_vert, cnt, 0));
    CU(cudaMemcpyAsync(ctx->tab_lat, c_lat, cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->tab_vert, c_vert, cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->nmat = nmat; ctx->nT = nT; ctx->T0 = T0; ctx->dT = dT;
    ctx->tables_iso = memcmp(c_lat, c_vert, cnt * sizeof(double)) == 0;
    ctx->have_materials = true;
    ctx->conds_valid = false;
    return PFEM_OK;
}

// x[node] = value and/or fixed[node] = 1 for the stored Dirichlet list
static int scatter_bc(pfem_ctx* ctx, double* x, uint8_t* fixed, bool first = false) {
    if (!ctx->nbc) return PFEM_OK;
    int blocks = (int)((ctx->nbc + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    k_scatter_dirichlet<<<blocks, 256, 0, ctx->stream>>>(ctx->nbc, ctx->bc_node, (first && ctx->bc_val_first) ? ctx->bc_val_first : ctx->bc_val, x, fixed);
    KCHECK(); LAUNCHED(1);
    return PFEM_OK;
}

extern "C" int pfem_set_dirichlet(pfem_ctx* ctx, size_t nd, const size_t* node, const double* value) {
    NEED_MESH();
    const Grid& g = ctx->g;
    if (nd && (!node || !value)) FAIL(PFEM_ERR_BAD_INPUT, "null Dirichlet argument");
    // De-duplicate.  setBC (iterative_matrix.hpp:462-485) handles the conditions one by one: the FIRST one that names node r
    // lifts its value into the neighbours' right-hand sides and clears the couplings, every later one only overwrites B[r].
    // So the free rows see the first value, the node itself ends with the last one; both are kept.
    std::vector<idx_t> nn;
    std::vector<double> vv, vfirst;
    bool conflict = false;
    nn.reserve(nd); vv.reserve(nd); vfirst.reserve(nd);
    {
        std::vector<long long> last;  // node -> position, lazily via sort-free map for big meshes
        std::vector<std::pair<idx_t, size_t>> order(nd);
        for (size_t m = 0; m < nd; ++m) {
            if (node[m] >= (size_t)g.N) FAIL(PFEM_ERR_BAD_INPUT, "Dirichlet node %zu out of range", node[m]);
            if (!(value[m] == value[m]) || isinf(value[m])) FAIL(PFEM_ERR_BAD_INPUT, "non-finite Dirichlet value");
            order[m] = {abi_to_lattice(ctx, (idx_t)node[m]), m};   // ABI index (dense rows) -> pitched lattice index
        }
        std::sort(order.begin(), order.end());
        size_t run0 = 0;
        for (size_t m = 0; m < nd; ++m) {
            if (m > 0 && order[m].first != order[m - 1].first) run0 = m;
            if (m + 1 < nd && order[m + 1].first == order[m].first) continue;  // keep the last occurrence
            nn.push_back(order[m].first);
            vv.push_back(value[order[m].second]);
            vfirst.push_back(value[order[run0].second]);
            if (vfirst.back() != vv.back()) conflict = true;
        }
    }
    // The values are NOT written into the field here: like the reference (applyBC works on the
    // matrix and B only, matrix.hpp:111-118) the field keeps its previous values until the solve,
    // so the conductivities of the first loop are evaluated from the un-constrained field.
    CU(cudaMemsetAsync(ctx->fixed, 0, (size_t)g.NP, ctx->stream));
    ctx->nbc = nn.size();
    dev_release(ctx, &ctx->bc_node);
    dev_release(ctx, &ctx->bc_val);
    dev_release(ctx, &ctx->bc_val_first);
    if (conflict) {
        TRY(dev_alloc(ctx, &ctx->bc_val_first, nn.size(), 0));
        CU(cudaMemcpyAsync(ctx->bc_val_first, vfirst.data(), vfirst.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!nn.empty()) {
        TRY(dev_alloc(ctx, &ctx->bc_node, nn.size(), 0));
        TRY(dev_alloc(ctx, &ctx->bc_val, nn.size(), 0));
        CU(cudaMemcpyAsync(ctx->bc_node, nn.data(), nn.size() * sizeof(idx_t), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->bc_val, vv.data(), vv.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        TRY(scatter_bc(ctx, nullptr, ctx->fixed));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

static int ensure_elem_arrays(pfem_ctx* ctx, bool shockley) {
    const size_t N = (size_t)ctx->g.NP, G = (size_t)ctx->g.G;
    if (!ctx->aux0) {
        TRY(dev_alloc(ctx, &ctx->aux0, N, G));
        TRY(dev_alloc(ctx, &ctx->aux1, N, G));
        TRY(dev_alloc(ctx, &ctx->aux2, N, G));
    }
    if (shockley && !ctx->cur0) {
        TRY(dev_alloc(ctx, &ctx->cur0, N, G));
        TRY(dev_alloc(ctx, &ctx->cur1, N, G));
        TRY(dev_alloc(ctx, &ctx->cur2, N, G));
    }
    return PFEM_OK;
}

extern "C" int pfem_set_source(pfem_ctx* ctx, const double* heat) {
    NEED_MESH();
    const Grid& g = ctx->g;
    ctx->source_set = true;
    if (!heat) {
        CU(cudaMemsetAsync(ctx->f, 0, (size_t)g.NP * sizeof(double), ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return PFEM_OK;
    }
    TRY(ensure_elem_arrays(ctx, false));
    CU(cudaMemsetAsync(ctx->aux0 - g.G, 0, (size_t)(g.NP + 2 * g.G) * sizeof(double), ctx->stream));
    TRY(upload_elem<double, 1>(ctx, heat, ctx->aux0, nullptr, nullptr));
    k_load_vector<<<node_grid(g), node_block(), 0, ctx->stream>>>(g, ctx->aux0, ctx->has_excluded ? ctx->mat : nullptr, ctx->f);
    KCHECK(); LAUNCHED(1);
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

// ------------------------------------------------ boundary conditions of the 2nd / 3rd kind, radiation (a7) ------
// setBoundaries (therm3d.cpp:140-168) and its three uses in setMatrix (:242-268), flattened once on the host into the
// row list `Surf` (kernels_surface.cuh).  The per-node optional values are what
// BoundaryConditionsWithMesh::getValue yields (first matching condition wins, boundary_conditions.hpp:182-186).

template <typename T>
static int surf_upload(pfem_ctx* ctx, const T** dst, const std::vector<T>& v) {
    T* d = nullptr;
    TRY(dev_alloc(ctx, &d, v.size() ? v.size() : 1, 0));
    if (!v.empty()) CU(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *dst = d;
    return PFEM_OK;
}

static void surf_release(pfem_ctx* ctx) {
    Surf& s = ctx->surf;
    dev_release(ctx, const_cast<idx_t**>(&s.node));
    dev_release(ctx, const_cast<double**>(&s.lconst));
    dev_release(ctx, const_cast<int**>(&s.radptr));
    dev_release(ctx, const_cast<idx_t**>(&s.rad_src));
    dev_release(ctx, const_cast<double**>(&s.rad_coef));
    dev_release(ctx, const_cast<double**>(&s.rad_amb4));
    dev_release(ctx, &s.radv);
    dev_release(ctx, const_cast<int**>(&s.kptr));
    dev_release(ctx, const_cast<idx_t**>(&s.kcol));
    dev_release(ctx, const_cast<double**>(&s.kval));
    memset(&s, 0, sizeof s);
    ctx->surf_iter = 0; ctx->surf_iter_known = false;
    if (ctx->graph) { cudaGraphExecDestroy(ctx->graph); ctx->graph = nullptr; }
}

// Boundary conditions of the 2nd / 3rd kind and radiation of the 2-D thermal solvers (ThermalFem2DSolver::setMatrix, therm2d.cpp) in the
// reference's own 2-D units: edge by edge like setBoundaries (:138-172) with the terms of :225-265 (Cartesian) and :371-413
// (cylindrical).  An edge carries a condition when both of its nodes have a value.  Unlike the 3-D solver the 2-D one accumulates
// into the right nodes (F1..F4 by name) and radiation reads the wall node itself; what b->verbatim reproduces here is (i) the convection
// MATRIX terms missing the 1e-6 (um -> m) that the load terms carry (:241,244 against :238) and (ii), cylindrical, their second
// factor r: they are added to k11.. before A(..) += r * k11 (:385-397, :415-426).  Pure host code (pfem_edges2d_host exposes it to
// the CPU tests).  Node (i1, i2) -> i1 * s1 + i2 * s2 (flags and values are read there), element (e1, e2) -> e1 * es1 + e2 * es2.
struct Edge2DTerms {
    std::vector<idx_t> lnode, rnode, krow, kcol;
    std::vector<double> lval, rcoef, ramb4, kval;    // load; radiation: load = -rcoef (T^4 - ramb4); matrix triplets (both orientations)
};
static void edges2d(idx_t n1, idx_t n2, const double* X, const double* Y, idx_t s1, idx_t s2, const uint8_t* excluded, idx_t es1, idx_t es2,
                    const pfem_boundary* b, bool cyl, Edge2DTerms& t) {
    const double SB = 5.670373e-8;   // plask/phys/constants.hpp:41
    const bool quirk = b->verbatim != 0;
    const double kunit = quirk ? 1. : 1e-6;
    auto flags = [&](idx_t n) {
        return (uint8_t)((b->has_flux && b->has_flux[n] ? 1 : 0) | (b->has_conv && b->has_conv[n] ? 2 : 0) | (b->has_rad && b->has_rad[n] ? 4 : 0));
    };
    for (idx_t e1 = 0; e1 + 1 < n1; ++e1)
        for (idx_t e2 = 0; e2 + 1 < n2; ++e2) {
            if (excluded && excluded[(size_t)(es1 * e1 + es2 * e2)]) continue;                       // :196: elements of the masked mesh only
            const idx_t c4[4] = {e1 * s1 + e2 * s2, (e1 + 1) * s1 + e2 * s2,                         // lower left, lower right,
                                 (e1 + 1) * s1 + (e2 + 1) * s2, e1 * s1 + (e2 + 1) * s2};            // upper right, upper left
            uint8_t mk4[4], any = 0;
            for (int l = 0; l < 4; ++l) { mk4[l] = flags(c4[l]); any |= mk4[l]; }
            if (!any) continue;
            const double width = X[e1 + 1] - X[e1], height = Y[e2 + 1] - Y[e2], rmid = 0.5 * (X[e1] + X[e1 + 1]);
            const double kr = (cyl && quirk) ? rmid : 1.;
            for (int kind = 0; kind < 3; ++kind) {           // heat flux, convection, radiation: the order of :225-265
                const uint8_t bit = (uint8_t)(1 << kind);
                for (int side = 0; side < 4; ++side) {       // bottom (1,2), right (2,3), top (3,4), left (4,1): :153-172
                    const int la = side, lb = (side + 1) & 3;
                    if (!(mk4[la] & mk4[lb] & bit)) continue;
                    const double len = (side & 1) ? height : width;
                    for (int pass = 0; pass < 2; ++pass) {   // the terms of node a with partner b, then of b with a
                        const int l1 = pass ? lb : la, l2 = pass ? la : lb;
                        const idx_t na = c4[l1], nb = c4[l2];
                        // radial factors: the radius of a vertical edge; on a horizontal one r -+ len/6 for the node at the lower /
                        // higher radius (:376-378), r for the off-diagonal entry (:397)
                        double rf = 1., rf_off = 1.;
                        if (cyl) {
                            const bool inner = (l1 == 0 || l1 == 3);
                            rf = side == 3 ? X[e1] : side == 1 ? X[e1 + 1] : rmid + (inner ? -len / 6. : len / 6.);
                            rf_off = side == 3 ? X[e1] : side == 1 ? X[e1 + 1] : rmid;
                        }
                        if (kind == 0) { t.lnode.push_back(na); t.lval.push_back(-0.5e-6 * len * b->flux[na] * rf); }
                        else if (kind == 2) {
                            double a = b->rad_ambient[na]; a = a * a;
                            t.rnode.push_back(na); t.rcoef.push_back(0.5e-6 * len * b->rad_emissivity[na] * SB * rf); t.ramb4.push_back(a * a);
                        } else {
                            const double csum = b->conv_coeff[na] + b->conv_coeff[nb];
                            t.lnode.push_back(na);
                            t.lval.push_back(cyl ? 0.125e-6 * len * csum * (b->conv_ambient[na] + b->conv_ambient[nb]) * rf
                                                 : 0.5e-6 * len * b->conv_coeff[na] * b->conv_ambient[na]);
                            t.krow.push_back(na); t.kcol.push_back(na); t.kval.push_back(kunit * kr * csum * len / 6. * rf);
                            t.krow.push_back(na); t.kcol.push_back(nb); t.kval.push_back(kunit * kr * csum * len / 12. * rf_off);
                        }
                    }
                }
            }
        }
}

extern "C" int pfem_edges2d_host(size_t n1, const double* x, size_t n2, const double* y, const pfem_boundary* b,
                                 double* load, double* rad_coef, double* rad_amb4, double* K) {
    if (!x || !y || !b || !load || !rad_coef || !rad_amb4 || !K || n1 < 2 || n2 < 2) return PFEM_ERR_BAD_INPUT;
    if (b->mode2d != 1 && b->mode2d != 2) return PFEM_ERR_BAD_INPUT;
    if ((b->has_flux && !b->flux) || (b->has_conv && (!b->conv_coeff || !b->conv_ambient)) || (b->has_rad && (!b->rad_emissivity || !b->rad_ambient)))
        return PFEM_ERR_BAD_INPUT;
    const size_t N = n1 * n2;
    Edge2DTerms t;
    edges2d((idx_t)n1, (idx_t)n2, x, y, (idx_t)n2, 1, nullptr, 0, 0, b, b->mode2d == 2, t);
    for (size_t k = 0; k < N; ++k) load[k] = rad_coef[k] = rad_amb4[k] = 0.;
    for (size_t k = 0; k < N * N; ++k) K[k] = 0.;
    for (size_t k = 0; k < t.lnode.size(); ++k) load[t.lnode[k]] += t.lval[k];
    for (size_t k = 0; k < t.rnode.size(); ++k) { rad_coef[t.rnode[k]] += t.rcoef[k]; rad_amb4[t.rnode[k]] = t.ramb4[k]; }
    for (size_t k = 0; k < t.krow.size(); ++k) K[(size_t)t.krow[k] * N + (size_t)t.kcol[k]] += t.kval[k];
    return PFEM_OK;
}

extern "C" int pfem_set_boundary(pfem_ctx* ctx, const pfem_boundary* b) {
    NEED_MESH();
    const Grid& g = ctx->g;
    CU(cudaStreamSynchronize(ctx->stream));
    surf_release(ctx);
    if (!b || (!b->has_flux && !b->has_conv && !b->has_rad)) return PFEM_OK;
    if ((b->has_flux && !b->flux) || (b->has_conv && (!b->conv_coeff || !b->conv_ambient)) ||
        (b->has_rad && (!b->rad_emissivity || !b->rad_ambient)))
        FAIL(PFEM_ERR_BAD_INPUT, "boundary condition flags without values");
    const int mode2d = b->mode2d;
    if (mode2d < 0 || mode2d > 2) FAIL(PFEM_ERR_BAD_INPUT, "pfem_boundary::mode2d must be 0 (brick), 1 (Cartesian 2-D) or 2 (cylindrical 2-D)");
    if (ctx->weighted && mode2d != 2)
        FAIL(PFEM_ERR_BAD_INPUT, "element weights (pfem_set_axis_weight) combine with boundary conditions of the 2nd / 3rd kind only in the cylindrical 2-D mode (mode2d = 2)");
    if (mode2d && (g.pn[0] != 2 || ctx->nranks > 1))
        FAIL(PFEM_ERR_BAD_INPUT, "mode2d needs the one-layer embedding of a 2-D mesh (2 nodes along physical axis 0) on a single device");
    if (mode2d == 2 && !(ctx->hax[1].front() >= 0.)) FAIL(PFEM_ERR_BAD_INPUT, "mode2d = 2: negative radius on physical axis 1");
    if (b->verbatim && b->has_rad && ctx->nranks > 1 && !mode2d)
        FAIL(PFEM_ERR_BAD_INPUT, "slab mode: verbatim radiation reads temperatures[0..7] of the whole mesh (therm3d.cpp:265); use the corrected form");
    // dense (ABI) and lattice strides of the physical axes; element extents along them
    idx_t dps[3], cnt[3] = {g.nI, g.nJ, g.nK};
    for (int a = 0; a < 3; ++a) dps[a] = g.abi_ns[g.dim_of_phys[a]];
    const size_t N = (size_t)g.N;
    std::vector<uint8_t> mask(N, 0);
    for (size_t n = 0; n < N; ++n)
        mask[n] = (uint8_t)((b->has_flux && b->has_flux[n] ? 1 : 0) | (b->has_conv && b->has_conv[n] ? 2 : 0) | (b->has_rad && b->has_rad[n] ? 4 : 0));
    auto lattice = [&](idx_t r) { return abi_to_lattice(ctx, r); };
    struct LoadT { idx_t node; double v; };
    struct RadT { idx_t node, src; double coef, amb4; };
    struct KT { idx_t row, col; double v; };
    std::vector<LoadT> loads;
    std::vector<RadT> rads;
    std::vector<KT> kts;
    static const int walls[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}};
    const double SB = 5.670373e-8;   // plask/phys/constants.hpp:41
    const int quirk = b->verbatim ? 1 : 0;
    if (mode2d) {
        // the brick operator and load restricted to one of the two node planes are  emb = d/2 * 1e-6  times their 2-D counterparts
        // (INTEGRATION.md 9): the edge terms of the 2-D solver go to both planes, scaled by emb
        const double emb = 0.5e-6 * (ctx->hax[0][1] - ctx->hax[0][0]);
        Edge2DTerms t;
        edges2d(g.pn[1], g.pn[2], ctx->hax[1].data(), ctx->hax[2].data(), dps[1], dps[2],
                ctx->h_excluded.empty() ? nullptr : ctx->h_excluded.data(), g.es[1], g.es[2], b, mode2d == 2, t);
        for (int plane = 0; plane < 2; ++plane) {
            const idx_t o = plane * dps[0];
            for (size_t k = 0; k < t.lnode.size(); ++k) loads.push_back({lattice(t.lnode[k] + o), emb * t.lval[k]});
            for (size_t k = 0; k < t.rnode.size(); ++k) rads.push_back({lattice(t.rnode[k] + o), lattice(t.rnode[k] + o), emb * t.rcoef[k], t.ramb4[k]});
            for (size_t k = 0; k < t.krow.size(); ++k) kts.push_back({lattice(t.krow[k] + o), lattice(t.kcol[k] + o), emb * t.kval[k]});
        }
    } else
    for (idx_t ek = 0; ek < cnt[2] - 1; ++ek)
        for (idx_t ej = 0; ej < cnt[1] - 1; ++ej)
            for (idx_t ei = 0; ei < cnt[0] - 1; ++ei) {
                const idx_t n0 = ei * g.abi_ns[0] + ej * g.abi_ns[1] + ek * g.abi_ns[2];   // dense (ABI) index of the lowest corner
                idx_t idx[8];
                uint8_t any = 0, mk[8];
                for (int l = 0; l < 8; ++l) {
                    idx[l] = n0 + ((l & 1) ? dps[0] : 0) + ((l & 2) ? dps[1] : 0) + ((l & 4) ? dps[2] : 0);
                    mk[l] = mask[idx[l]];
                    any |= mk[l];
                }
                if (!any) continue;
                const idx_t ix[3] = {ei, ej, ek};
                if (!ctx->h_excluded.empty()) {
                    // setBoundaries runs inside `for (elem : maskedMesh->elements())` (therm3d.cpp:186-268): an element outside
                    // the masked mesh adds nothing, even if all four nodes of one of its sides carry a condition
                    idx_t e = 0;
                    for (int a = 0; a < 3; ++a) e += g.es[a] * ix[g.dim_of_phys[a]];
                    if (ctx->h_excluded[(size_t)e]) continue;
                }
                double d[3];
                for (int a = 0; a < 3; ++a) { const idx_t t = ix[g.dim_of_phys[a]]; d[a] = ctx->hax[a][t + 1] - ctx->hax[a][t]; }
                const double areas[3] = {d[0] * d[1], d[1] * d[2], d[2] * d[0]};
                double F[8] = {0, 0, 0, 0, 0, 0, 0, 0}, K[8][8];
                bool anyK = false, anyF = false;
                memset(K, 0, sizeof K);
                for (int kind = 0; kind < 3; ++kind) {      // heat flux, convection, radiation: the order of :242-268
                    const uint8_t bit = (uint8_t)(1 << kind);
                    for (int side = 0; side < 6; ++side) {
                        const int* w = walls[side];
                        if (!(mk[w[0]] & mk[w[1]] & mk[w[2]] & mk[w[3]] & bit)) continue;   // all four nodes carry it, :153
                        const double area = areas[side / 2];
                        for (int i = 0; i < 4; ++i) {
                            const idx_t ni = idx[w[i]];
                            const int si = quirk ? i : w[i];   // slot that receives the term (:157)
                            if (kind == 0) { F[si] += -0.25e-12 * area * b->flux[ni]; anyF = true; }
                            else if (kind == 1) { F[si] += 0.25e-12 * area * b->conv_coeff[ni] * b->conv_ambient[ni]; anyF = true; }
                            else {
                                double a = b->rad_ambient[ni]; a = a * a;
                                // :265 reads temperatures[i] with the LOCAL node number; corrected: the wall node itself
                                const idx_t src = quirk ? (idx_t)w[i] : ni;
                                if ((size_t)src >= N) FAIL(PFEM_ERR_BAD_INPUT, "mesh has fewer than 8 nodes");
                                rads.push_back({lattice(idx[si]), lattice(src), 0.25e-12 * area * b->rad_emissivity[ni] * SB, a * a});
                            }
                            if (kind != 1) continue;
                            for (int j = 0; j <= i; ++j) {
                                const int sj = quirk ? j : w[j];
                                const int ij = quirk ? (i ^ j) : (w[i] ^ w[j]);
                                const bool edge = (ij == 1 || ij == 2 || ij == 4);
                                // :255 has 0.125e-12, a quarter of the consistent face mass matrix (rows sum to A c/16 against a
                                // load of A c Ta/4); kept verbatim, the corrected form uses the consistent matrix
                                double v = (quirk ? 0.125e-12 : 0.5e-12) * area * (b->conv_coeff[ni] + b->conv_coeff[idx[w[j]]]);
                                v = v / (w[j] == w[i] ? 9. : edge ? 18. : 36.);
                                if (si >= sj) K[si][sj] += v; else K[sj][si] += v;
                                anyK = true;
                            }
                        }
                    }
                }
                if (anyF)
                    for (int l = 0; l < 8; ++l)
                        if (F[l] != 0.) loads.push_back({lattice(idx[l]), F[l]});
                if (anyK)
                    for (int i = 0; i < 8; ++i)
                        for (int j = 0; j <= i; ++j)
                            if (K[i][j] != 0.) {
                                const idx_t r = lattice(idx[i]), c = lattice(idx[j]);
                                kts.push_back({r, c, K[i][j]});
                                if (r != c) kts.push_back({c, r, K[i][j]});
                            }
            }
    if (loads.empty() && rads.empty() && kts.empty()) return PFEM_OK;
    std::stable_sort(loads.begin(), loads.end(), [](const LoadT& a, const LoadT& c) { return a.node < c.node; });
    std::stable_sort(rads.begin(), rads.end(), [](const RadT& a, const RadT& c) { return a.node < c.node; });
    std::stable_sort(kts.begin(), kts.end(), [](const KT& a, const KT& c) { return a.row != c.row ? a.row < c.row : a.col < c.col; });
    std::vector<idx_t> rows;
    rows.reserve(loads.size() + rads.size() + kts.size());
    for (auto& t : loads) rows.push_back(t.node);
    for (auto& t : rads) rows.push_back(t.node);
    for (auto& t : kts) rows.push_back(t.row);
    std::sort(rows.begin(), rows.end());
    rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
    const size_t nr = rows.size();
    if (nr > 0x7fffffffull || kts.size() > 0x7fffffffull || rads.size() > 0x7fffffffull) FAIL(PFEM_ERR_BAD_INPUT, "too many boundary terms");
    std::vector<double> lconst(nr, 0.), rcoef, ramb, kval;
    std::vector<int> radptr(nr + 1, 0), kptr(nr + 1, 0);
    std::vector<idx_t> rsrc, kcol;
    size_t il = 0, ir = 0, ik = 0;
    for (size_t r = 0; r < nr; ++r) {
        const idx_t n = rows[r];
        for (; il < loads.size() && loads[il].node == n; ++il) lconst[r] += loads[il].v;
        radptr[r] = (int)rsrc.size();
        for (; ir < rads.size() && rads[ir].node == n; ++ir) { rsrc.push_back(rads[ir].src); rcoef.push_back(rads[ir].coef); ramb.push_back(rads[ir].amb4); }
        kptr[r] = (int)kcol.size();
        for (; ik < kts.size() && kts[ik].row == n; ++ik) {
            if ((int)kcol.size() > kptr[r] && kcol.back() == kts[ik].col) kval.back() += kts[ik].v;
            else { kcol.push_back(kts[ik].col); kval.push_back(kts[ik].v); }
        }
    }
    radptr[nr] = (int)rsrc.size();
    kptr[nr] = (int)kcol.size();
    Surf& s = ctx->surf;
    TRY(surf_upload(ctx, &s.node, rows));
    TRY(surf_upload(ctx, &s.lconst, lconst));
    TRY(surf_upload(ctx, &s.radptr, radptr));
    TRY(surf_upload(ctx, &s.rad_src, rsrc));
    TRY(surf_upload(ctx, &s.rad_coef, rcoef));
    TRY(surf_upload(ctx, &s.rad_amb4, ramb));
    TRY(surf_upload(ctx, &s.kptr, kptr));
    TRY(surf_upload(ctx, &s.kcol, kcol));
    TRY(surf_upload(ctx, &s.kval, kval));
    TRY(dev_alloc(ctx, &s.radv, nr, 0));
    if (!ctx->fS) TRY(dev_alloc(ctx, &ctx->fS, (size_t)g.NP, (size_t)g.G));
    s.nrows = (int)nr;
    s.nnz = (int)kcol.size();
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

static inline int surf_blocks(const pfem_ctx* ctx) {
    int b = (ctx->surf.nrows + 255) / 256;
    const int cap = ctx->sm_count * 4;
    return b < 1 ? 1 : (b > cap ? cap : b);
}

// radiation loads from the temperatures the loop starts with (call BEFORE the Dirichlet values are scattered into x)
static int surf_eval_rad(pfem_ctx* ctx) {
    if (!ctx->surf.nrows) return PFEM_OK;
    k_surf_rad<<<surf_blocks(ctx), 256, 0, ctx->stream>>>(ctx->surf, ctx->x);
    KCHECK(); LAUNCHED(1);
    return PFEM_OK;
}

// effective load vector for  M (f - A in):  f + boundary loads - S in  (ctx->f itself without boundary terms)
static int surf_rhs(pfem_ctx* ctx, const double* in, const double** feff) {
    *feff = ctx->f;
    if (!ctx->surf.nrows) return PFEM_OK;
    CU(cudaMemcpyAsync(ctx->fS, ctx->f, (size_t)ctx->g.NP * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    k_surf_rhs<<<surf_blocks(ctx), 256, 0, ctx->stream>>>(ctx->surf, ctx->f, in, ctx->fS);
    KCHECK(); LAUNCHED(1);
    *feff = ctx->fS;
    return PFEM_OK;
}

extern "C" int pfem_set_field(pfem_ctx* ctx, const double* x0) {
    NEED_MESH();
    if (!x0) FAIL(PFEM_ERR_BAD_INPUT, "null field");
    CU(upload_nodes(ctx, ctx->x, x0));
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

extern "C" int pfem_fill_field(pfem_ctx* ctx, double value) {
    NEED_MESH();
    k_fill_nodes<<<node_grid(ctx->g), node_block(), 0, ctx->stream>>>(ctx->g, ctx->x, value);
    KCHECK(); LAUNCHED(1);
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

extern "C" int pfem_set_elem_temperature(pfem_ctx* ctx, const double* T_elem, double uniform_T) {
    NEED_MESH();
    const Grid& g = ctx->g;
    if (!ctx->Te) TRY(dev_alloc(ctx, &ctx->Te, (size_t)g.NP, (size_t)g.G));
    if (T_elem) {
        TRY(upload_elem<double, 1>(ctx, T_elem, ctx->Te, nullptr, nullptr));
    } else {
        k_fill_nodes<<<node_grid(g), node_block(), 0, ctx->stream>>>(g, ctx->Te, uniform_T);
        KCHECK(); LAUNCHED(1);
        CU(cudaStreamSynchronize(ctx->stream));
    }
    ctx->conds_valid = false;
    return PFEM_OK;
}

extern "C" int pfem_set_junctions(pfem_ctx* ctx, uint32_t njunc, const pfem_junction* junc, const uint32_t* elem_junc,
                                  const uint8_t* elem_role, double pcond, double ncond, size_t ncol,
                                  const double* junc_cond, const double* beta_col, const double* js_col, int stable) {
    NEED_MESH();
    const Grid& g = ctx->g;
    if (njunc && (!junc || !elem_junc || !junc_cond || !beta_col || !js_col)) FAIL(PFEM_ERR_BAD_INPUT, "null junction argument");
    const size_t N = (size_t)g.NP, G = (size_t)g.G;
    if (!ctx->junc) TRY(dev_alloc(ctx, &ctx->junc, N, G));
    if (!ctx->role) TRY(dev_alloc(ctx, &ctx->role, N, G));
    CU(cudaMemsetAsync(ctx->junc, 0, N * sizeof(uint32_t), ctx->stream));
    CU(cudaMemsetAsync(ctx->role, 0, N, ctx->stream));
    if (elem_junc) {
        for (idx_t e = 0; e < g.E; ++e)
            if (elem_junc[e] > njunc) FAIL(PFEM_ERR_BAD_INPUT, "element %lld names junction %u > njunc", (long long)e, elem_junc[e]);
        TRY(upload_elem<uint32_t, 1>(ctx, elem_junc, ctx->junc, nullptr, nullptr));
    }
    if (elem_role) TRY(upload_elem<uint8_t, 1>(ctx, elem_role, ctx->role, nullptr, nullptr));
    std::vector<JunctionDev> hj(njunc);
    size_t need = 0;
    for (uint32_t a = 0; a < njunc; ++a) {
        const pfem_junction& s = junc[a];
        if (s.top <= s.bottom || s.top >= (size_t)g.pn[2] || s.right > (size_t)g.pn[1] - 1 || s.front > (size_t)g.pn[0] - 1 ||
            s.left > s.right || s.back > s.front || !(s.height > 0.))
            FAIL(PFEM_ERR_BAD_INPUT, "junction %u has inconsistent extents", a);
        hj[a] = {(idx_t)s.bottom, (idx_t)s.top, (idx_t)s.left, (idx_t)s.right, (idx_t)s.back, (idx_t)s.front, (idx_t)s.ld,
                 (idx_t)s.offset, s.height};
        if (s.right > s.left && s.front > s.back) {
            size_t last = (size_t)(s.offset + (ptrdiff_t)(s.ld * (s.right - 1) + (s.front - 1))) + 1;
            if (last > need) need = last;
        }
    }
    if (need > ncol) FAIL(PFEM_ERR_BAD_INPUT, "junction table too short: need %zu entries, got %zu", need, ncol);
    if (njunc) {
        // every element of junction k addresses table entry offset + ld * i1 + i0 (electr3d.cpp:213-215,250): it must exist,
        // also for elements outside [back,front) x [left,right) (setupActiveRegions does not widen the region for the
        // column that creates it, electr3d.cpp:120-135)
        int ord[3] = {0, 1, 2};
        std::sort(ord, ord + 3, [&](int a, int b) { return g.es[a] > g.es[b]; });   // major, medium, minor element axis
        for (idx_t e = 0; e < g.E; ++e) {
            if (!elem_junc[e]) continue;
            idx_t c[3], rem = e;
            for (int q = 0; q < 3; ++q) { c[ord[q]] = rem / g.es[ord[q]]; rem %= g.es[ord[q]]; }
            const pfem_junction& a = junc[elem_junc[e] - 1];
            const long long col = (long long)a.offset + (long long)a.ld * c[1] + c[0];
            if (col < 0 || col >= (long long)ncol)
                FAIL(PFEM_ERR_BAD_INPUT, "element %lld of junction %u addresses junction-table entry %lld outside [0, %zu)", (long long)e, elem_junc[e] - 1, col, ncol);
        }
    }
    ctx->nact = (int)njunc;
    ctx->ncol = ncol;
    dev_release(ctx, &ctx->act);
    dev_release(ctx, &ctx->junc_cond);
    dev_release(ctx, &ctx->beta_col);
    dev_release(ctx, &ctx->js_col);
    if (njunc) {
        TRY(dev_alloc(ctx, &ctx->act, njunc, 0));
        CU(cudaMemcpyAsync(ctx->act, hj.data(), njunc * sizeof(JunctionDev), cudaMemcpyHostToDevice, ctx->stream));
        TRY(dev_alloc(ctx, &ctx->junc_cond, 2 * ncol, 0));
        TRY(dev_alloc(ctx, &ctx->beta_col, ncol, 0));
        TRY(dev_alloc(ctx, &ctx->js_col, ncol, 0));
        CU(cudaMemcpyAsync(ctx->junc_cond, junc_cond, 2 * ncol * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->beta_col, beta_col, ncol * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->js_col, js_col, ncol * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->pcond = pcond; ctx->ncond = ncond; ctx->stable = stable;
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have_junctions = true;
    ctx->conds_valid = false;
    return PFEM_OK;
}

// ------------------------------------------------------------------------ slab mode -----
// One context per GPU / process; the local mesh given to pfem_set_mesh contains the owned node planes along the
// MAJOR axis plus one halo plane towards each existing neighbour.  Peers are mapped with CUDA IPC (NVLink P2P).
// One process per context is required: the kernels of a collective call spin on their peers, and any implicitly
// device-synchronising runtime call of a sibling context in the same process would dead-lock against them.

extern "C" size_t pfem_slab_blob_size(void) { return sizeof(SlabBlob); }

extern "C" int pfem_slab_configure(pfem_ctx* ctx, int rank, int nranks, size_t own_lo, size_t own_hi) {
    NEED_MESH();
    Grid& g = ctx->g;
    if (nranks < 1 || nranks > PFEM_MAX_RANKS || rank < 0 || rank >= nranks) FAIL(PFEM_ERR_BAD_INPUT, "bad rank %d of %d (max %d)", rank, nranks, PFEM_MAX_RANKS);
    if (own_lo >= own_hi || own_hi > (size_t)g.nK) FAIL(PFEM_ERR_BAD_INPUT, "owned plane range [%zu,%zu) outside the local mesh (%d planes)", own_lo, own_hi, g.nK);
    if (own_lo != (rank > 0 ? 1u : 0u) || (size_t)g.nK - own_hi != (rank < nranks - 1 ? 1u : 0u))
        FAIL(PFEM_ERR_BAD_INPUT, "the local mesh must hold exactly one halo plane towards each neighbour");
    if (!ctx->fused.valid) FAIL(PFEM_ERR_STATE, "slab mode needs the fused PCG kernel: %s", ctx->fused.why);
    if (g.abi_dim[2] != 2) FAIL(PFEM_ERR_BAD_INPUT, "slab mode cuts the ABI's major axis, which this internal layout does not keep as the slowest axis (vertical major axis with pfem_set_layout)");
    CU(cudaStreamSynchronize(ctx->stream));
    slab_release(ctx);
    ctx->rank = rank; ctx->nranks = nranks;
    g.kown0 = (int)own_lo; g.kown1 = (int)own_hi;
    if (ctx->graph) { cudaGraphExecDestroy(ctx->graph); ctx->graph = nullptr; }
    ctx->line_plan.valid = false;
    if (nranks > 1) TRY(alloc_line_arrays(ctx));   // z of the line-Jacobi iteration is exported to the neighbours
    {   // the fused plan chunks the OWNED planes
        double* const rr[2] = {ctx->r, ctx->r2};
        double* const qq[2] = {ctx->q, ctx->q2};
        double* const pp[2] = {ctx->p, ctx->p2};
        ctx->fused = make_fused_plan(g, ctx->sm_count, rr, qq, pp, ctx->dinv, ctx->cl, ctx->cv, ctx->x);
    }
    if (nranks == 1) return PFEM_OK;
    CU(cudaMalloc(&ctx->inbox, sizeof(Inbox)));
    CU(cudaMemset(ctx->inbox, 0, sizeof(Inbox)));
    CU(cudaMalloc(&ctx->d_comm, sizeof(Comm)));
    CU(cudaMemset(ctx->d_comm, 0, sizeof(Comm)));
    return PFEM_OK;
}

extern "C" int pfem_slab_export(pfem_ctx* ctx, void* blob) {
    NEED_MESH();
    if (!blob) FAIL(PFEM_ERR_BAD_INPUT, "null blob");
    if (ctx->nranks < 2 || !ctx->inbox) FAIL(PFEM_ERR_STATE, "pfem_slab_configure with nranks > 1 has not been called");
    SlabBlob b;
    memset(&b, 0, sizeof b);
    for (int a = 0; a < SA_COUNT; ++a) CU(cudaIpcGetMemHandle(&b.arr[a], slab_array(ctx, a) - ctx->g.G));
    CU(cudaIpcGetMemHandle(&b.inbox, ctx->inbox));
    b.G = ctx->g.G; b.sK = ctx->g.sK; b.nK = ctx->g.nK; b.rank = ctx->rank; b.nranks = ctx->nranks;
    memcpy(blob, &b, sizeof b);
    return PFEM_OK;
}

extern "C" int pfem_slab_connect(pfem_ctx* ctx, const void* blobs) {
    NEED_MESH();
    if (!blobs) FAIL(PFEM_ERR_BAD_INPUT, "null blobs");
    if (ctx->nranks < 2 || !ctx->inbox) FAIL(PFEM_ERR_STATE, "pfem_slab_configure with nranks > 1 has not been called");
    const SlabBlob* B = reinterpret_cast<const SlabBlob*>(blobs);
    Comm hc;
    memset(&hc, 0, sizeof hc);
    hc.rank = ctx->rank; hc.nranks = ctx->nranks;
    for (int r = 0; r < ctx->nranks; ++r) {
        if (B[r].rank != r || B[r].nranks != ctx->nranks) FAIL(PFEM_ERR_BAD_INPUT, "blob %d does not come from rank %d of %d", r, r, ctx->nranks);
        if (B[r].sK != ctx->g.sK) FAIL(PFEM_ERR_BAD_INPUT, "rank %d has a different plane size", r);
        if (r == ctx->rank) { hc.inbox[r] = ctx->inbox; continue; }
        void* ptr = nullptr;
        CU(cudaIpcOpenMemHandle(&ptr, B[r].inbox, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_inbox[r] = reinterpret_cast<Inbox*>(ptr);
        hc.inbox[r] = ctx->peer_inbox[r];
        pfem_ctx::Neighbour* nb = (r == ctx->rank - 1) ? &ctx->nb_lo : (r == ctx->rank + 1) ? &ctx->nb_hi : nullptr;
        if (!nb) continue;
        nb->G = B[r].G; nb->sK = B[r].sK; nb->nK = B[r].nK;
        for (int a = 0; a < SA_COUNT; ++a) {
            void* base = nullptr;
            CU(cudaIpcOpenMemHandle(&base, B[r].arr[a], cudaIpcMemLazyEnablePeerAccess));
            nb->arr[a] = reinterpret_cast<double*>(base) + nb->G;
        }
        nb->present = true;
    }
    CU(cudaMemcpy(ctx->d_comm, &hc, sizeof hc, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(&ctx->d_sc->comm, &ctx->d_comm, sizeof(Comm*), cudaMemcpyHostToDevice));
    ctx->surf_iter_known = false;
    return PFEM_OK;
}

// cross-rank barrier; returns (through *any) the OR of `flag` over the ranks
static int rank_barrier(pfem_ctx* ctx, int flag, int* any) {
    if (ctx->nranks < 2) { if (any) *any = flag; return PFEM_OK; }
    k_rank_barrier<<<1, 32, 0, ctx->stream>>>(ctx->d_sc, flag);
    KCHECK(); LAUNCHED(1);
    if (any) {
        CU(cudaMemcpyAsync(ctx->h_sc, ctx->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        *any = ctx->h_sc->red[3] != 0.;
    }
    return PFEM_OK;
}

// refresh the halo planes of array `a` on both neighbours from my first / last owned plane
static int halo_sync(pfem_ctx* ctx, int a) {
    if (ctx->nranks < 2) return PFEM_OK;
    const Grid& g = ctx->g;
    // Receiver-ready barrier first: the local kernels that produced array `a` also wrote (meaningless) values
    // into their own halo planes; a neighbour that runs ahead must not push before those kernels have finished.
    TRY(rank_barrier(ctx, 0, nullptr));
    const double* mine = slab_array(ctx, a);
    double* dlo = ctx->nb_lo.present ? ctx->nb_lo.arr[a] + (ctx->nb_lo.nK - 1) * ctx->nb_lo.sK : nullptr;
    double* dhi = ctx->nb_hi.present ? ctx->nb_hi.arr[a] : nullptr;
    k_halo_push<<<vec_blocks(ctx) / 4, 256, 0, ctx->stream>>>(g.sK, mine + g.sK * g.kown0, dlo, mine + g.sK * (g.kown1 - 1), dhi);
    KCHECK(); LAUNCHED(1);
    return rank_barrier(ctx, 0, nullptr);
}

static PeerOut peer_out(pfem_ctx* ctx, int out_parity) {
    PeerOut po;
    memset(&po, 0, sizeof po);
    if (ctx->nb_lo.present) {
        const long long off = (ctx->nb_lo.nK - 1) * ctx->nb_lo.sK;
        po.r_lo = ctx->nb_lo.arr[out_parity ? SA_R2 : SA_R] + off;
        po.q_lo = ctx->nb_lo.arr[out_parity ? SA_Q2 : SA_Q] + off;
        po.p_lo = ctx->nb_lo.arr[out_parity ? SA_P2 : SA_P] + off;
    }
    if (ctx->nb_hi.present) {
        po.r_hi = ctx->nb_hi.arr[out_parity ? SA_R2 : SA_R];
        po.q_hi = ctx->nb_hi.arr[out_parity ? SA_Q2 : SA_Q];
        po.p_hi = ctx->nb_hi.arr[out_parity ? SA_P2 : SA_P];
    }
    if (ctx->nb_lo.present && ctx->nb_lo.arr[SA_LZ]) po.z_lo = ctx->nb_lo.arr[SA_LZ] + (ctx->nb_lo.nK - 1) * ctx->nb_lo.sK;
    if (ctx->nb_hi.present && ctx->nb_hi.arr[SA_LZ]) po.z_hi = ctx->nb_hi.arr[SA_LZ];
    return po;
}

// ------------------------------------------------------------------ conductivities ------

extern "C" int pfem_update_conductivity_thermal(pfem_ctx* ctx) {
    NEED_MESH();
    if (!ctx->have_materials) FAIL(PFEM_ERR_STATE, "pfem_set_materials has not been called");
    const Grid& g = ctx->g;
    k_cond_thermal<<<node_grid(g), node_block(), 0, ctx->stream>>>(g, ctx->x, ctx->mat, ctx->nT, ctx->T0, ctx->dT,
                                                                    ctx->tab_lat, ctx->tab_vert, ctx->cl, ctx->cv);
    KCHECK(); LAUNCHED(1);
    ctx->conds_valid = true;
    ctx->cond_iso = ctx->tables_iso;
    return PFEM_OK;
}

extern "C" int pfem_update_conductivity_shockley(pfem_ctx* ctx) {
    NEED_MESH();
    if (!ctx->have_materials) FAIL(PFEM_ERR_STATE, "pfem_set_materials has not been called");
    const Grid& g = ctx->g;
    if (!ctx->Te) TRY(pfem_set_elem_temperature(ctx, nullptr, 300.));  // inTemperature = 300 (electr3d.cpp:35)
    k_cond_shockley<<<node_grid(g), node_block(), 0, ctx->stream>>>(
        g, ctx->mat, ctx->nact ? ctx->junc : nullptr, ctx->role, ctx->Te, ctx->nT, ctx->T0, ctx->dT, ctx->tab_lat,
        ctx->tab_vert, ctx->act, ctx->junc_cond, ctx->pcond, ctx->ncond, ctx->cl, ctx->cv);
    KCHECK(); LAUNCHED(1);
    ctx->conds_valid = true;
    ctx->cond_iso = ctx->tables_iso && ctx->nact == 0;   // a junction element is (0, sigma): anisotropic by construction
    return PFEM_OK;
}

extern "C" int pfem_set_conductivity(pfem_ctx* ctx, const double* cond) {
    NEED_MESH();
    if (!cond) FAIL(PFEM_ERR_BAD_INPUT, "null conductivity");
    TRY(upload_elem<double, 2>(ctx, cond, ctx->cl, ctx->cv, nullptr));
    ctx->conds_valid = true;
    ctx->cond_iso = true;
    for (idx_t e = 0; e < ctx->g.E && ctx->cond_iso; ++e) ctx->cond_iso = cond[2 * e] == cond[2 * e + 1];
    return PFEM_OK;
}

// masked mesh: the field is pinned to 0 on the nodes that are not part of it
static int mask_field(pfem_ctx* ctx) {
    if (!ctx->has_excluded) return PFEM_OK;
    // owned planes only: whether a node of a halo plane belongs to the masked mesh is the neighbour's knowledge (its value
    // arrives with the halo exchange)
    const idx_t off = ctx->g.sK * ctx->g.kown0, cnt = ctx->g.sK * (ctx->g.kown1 - ctx->g.kown0);
    k_zero_inactive<<<vec_blocks(ctx), 256, 0, ctx->stream>>>(cnt, ctx->inactive + off, ctx->x + off);
    KCHECK(); LAUNCHED(1);
    return PFEM_OK;
}

// ------------------------------------------------------------------------ PCG -----------

__global__ void k_set_params(Scalars* sc, double tol2, int maxit, int bench, int surf, int line) {
    sc->tol2 = tol2; sc->maxit = maxit; sc->bench = bench; sc->neg_diag = 0; sc->surf = surf; sc->line = line;
}

// line-Jacobi preconditioner: arrays and the tensor maps of its operator step (k_fpcg reading z for r, zeros for q,
// the free-row mask for D^-1)
static int ensure_line(pfem_ctx* ctx) {
    if (ctx->line_plan.valid) return PFEM_OK;
    const Grid& g = ctx->g;
    TRY(alloc_line_arrays(ctx));
    double* const zz[2] = {ctx->lz, ctx->lz};
    double* const qq[2] = {ctx->q, ctx->q2};
    double* const pp[2] = {ctx->p, ctx->p2};
    FusedPlan f = make_fused_plan(g, ctx->sm_count, zz, qq, pp, ctx->lmask, ctx->cl, ctx->cv, ctx->x);
    if (!f.valid) FAIL(PFEM_ERR_STATE, "line preconditioner: %s", f.why);
    ctx->line_plan = f;
    return PFEM_OK;
}

static inline int line_seg(const Grid& g) {   // nodes per lane of the I-line kernel
    for (int s = 2; s <= 16; s *= 2) if (32 * s >= g.nI) return s;
    return 0;
}

// L2 prefetch of the next row in the warp-per-row line kernels (k_line_I, k_line_ml); PFEM_LINE_PREFETCH=0 turns it off for A/B runs
static int line_prefetch() {
    static const int v = getenv("PFEM_LINE_PREFETCH") ? atoi(getenv("PFEM_LINE_PREFETCH")) : 1;
    return v;
}

// z = M^-1 (r - alpha q) for the line-Jacobi preconditioner; mode 1: only b.M^-1 b -> sc->bz
static int launch_line_solve(pfem_ctx* ctx, const double* r_in, const double* q_in, double* r_out, int mode) {
    const Grid& g = ctx->g;
    const PeerOut po = peer_out(ctx, 0);   // only the z pointers are used
    if (g.vdim == 0 && line_seg(g) > 0) {
        const idx_t rows = (idx_t)g.nJ * (g.kown1 - g.kown0);
        // balanced: every warp gets the same number of rows (8 warps per block)
        static const int cap_env = getenv("PFEM_LINE_BLOCKS") ? atoi(getenv("PFEM_LINE_BLOCKS")) : 0;
        const idx_t cap_warps = (idx_t)(cap_env > 0 ? cap_env : ctx->sm_count * 2) * 8;   // one resident wave (measured best)
        const idx_t rpw = (rows + cap_warps - 1) / cap_warps;
        int blocks = (int)(((rows + rpw - 1) / rpw + 7) / 8);
        if (blocks < 1) blocks = 1;
#define PFEM_LINE_CASE(S) case S: k_line_I<S><<<blocks, 256, 0, ctx->stream>>>(g, r_in, q_in, ctx->ll, ctx->ld, r_out, ctx->lz, ctx->d_sc, ctx->partials, mode, po, line_prefetch()); break;
        switch (line_seg(g)) {
            PFEM_LINE_CASE(2) PFEM_LINE_CASE(4) PFEM_LINE_CASE(8) PFEM_LINE_CASE(16)
            default: FAIL(PFEM_ERR_STATE, "no warp-per-row line kernel for %d nodes per line", g.nI);
        }
#undef PFEM_LINE_CASE
    } else {
        // lines along J or K, or along I when they are longer than the warp-per-row kernel holds (> 512 nodes)
        const idx_t lines = g.vdim == 0 ? (idx_t)g.nJ * (g.kown1 - g.kown0) : (idx_t)g.nI * (g.vdim == 1 ? g.kown1 - g.kown0 : g.nJ);
        int blocks = (int)std::min<idx_t>((lines + 127) / 128, (idx_t)ctx->sm_count * 16);
        if (blocks < 1) blocks = 1;
        k_line_strided<<<blocks, 128, 0, ctx->stream>>>(g, r_in, q_in, ctx->ll, ctx->ld, r_out, ctx->lz, ctx->d_sc, ctx->partials, mode, po);
    }
    KCHECK();   // counted by the caller (launch_iteration returns the kernels of an iteration)
    return PFEM_OK;
}
__global__ void k_force_running(Scalars* sc) { sc->done = 0; sc->status = 0; }

// ---- multilevel line preconditioner (kernels_ml.cuh) ----------------------------------------------------------
static int ml_blocks(const pfem_ctx* ctx, const LineDom& d) {
    const int cap = ctx->sm_count * 2;   // one resident wave of 256-thread blocks, like k_line_I
    const int na = ((d.nJ + PFEM_ML_C - 1) / PFEM_ML_C) * ((d.nK + d.koff + PFEM_ML_C - 1) / PFEM_ML_C);
    if (na <= cap) return na;
    const int per = (na + cap - 1) / cap;   // every block the same number of aggregates (+-1)
    return (na + per - 1) / per;
}

static int ensure_ml(pfem_ctx* ctx) {
    if (ctx->ml.nlev) return PFEM_OK;
    const Grid& g = ctx->g;
    if (g.vdim != 0) FAIL(PFEM_ERR_STATE, "the multilevel preconditioner needs the vertical axis as the fastest one: call pfem_set_layout(PFEM_LAYOUT_VERTICAL_MINOR) before pfem_set_mesh");
    if (line_seg(g) == 0) FAIL(PFEM_ERR_BAD_INPUT, "the multilevel preconditioner holds at most 512 nodes per vertical line (%d given)", g.nI);
    MLDev ml;
    memset(&ml, 0, sizeof ml);
    // Slab mode: the aggregates of the 4x4 and 16x16 levels must not straddle slabs, so that the hierarchy is the one of the whole
    // mesh cut into pieces (same aggregates, same line blocks) and only the top level — one column for the whole device — needs
    // an exchange.  Every rank with an upper neighbour therefore owns a multiple of 16 planes; the aggregate rows are counted from
    // 16 planes in front of the first owned plane (koff), which puts the halo planes into empty aggregates of their own.
    const bool slab = ctx->nranks > 1;
    const int koff = slab ? 16 - g.kown0 : 0;
    if (slab) {   // the refusal must be collective: a rank that went on alone would wait for the others in its first exchange
        const int mine = (ctx->nb_hi.present && (g.kown1 - g.kown0) % 16 != 0) || g.sJ * 2 > PFEM_COMM_VEC;
        int any = 0;
        TRY(rank_barrier(ctx, mine, &any));
        if (any)
            FAIL(PFEM_ERR_BAD_INPUT, "slab mode: the multilevel preconditioner needs a multiple of 16 owned planes on every rank but the last "
                 "(this rank: %d%s) and at most 512 nodes per vertical line", g.kown1 - g.kown0, mine ? ", not accepted" : "");
    }
    ml.dom[0] = LineDom{g.nI, g.nJ, g.nK, g.sJ, g.sK, g.kown0, g.kown1, koff};
    // aggregated levels (4x4, 16x16 lateral columns) as long as they hold more than one column, then the top level = one column
    int L = 0;
    for (int nJ = g.nJ, nK = g.nK + koff; L < PFEM_ML_MAXL - 1;) {
        nJ = (nJ + PFEM_ML_C - 1) / PFEM_ML_C; nK = (nK + PFEM_ML_C - 1) / PFEM_ML_C;
        if (nJ == 1 && nK == 1) break;
        ++L;
        ml.dom[L] = LineDom{g.nI, nJ, nK, g.sJ, g.sJ * nJ, 0, nK, 0};
    }
    ml.nagg = L;
    ml.nlev = L + 1;
    ml.dom[ml.nlev] = LineDom{g.nI, 1, 1, g.sJ, g.sJ, 0, 1, 0};
    // set-up partial sums live on the level-1 aggregates whether or not level 1 is a level of its own
    const int nJ1 = (g.nJ + PFEM_ML_C - 1) / PFEM_ML_C, nK1 = (g.nK + koff + PFEM_ML_C - 1) / PFEM_ML_C;
    const idx_t slen = (idx_t)nJ1 * nK1 * g.sJ;
    TRY(dev_alloc(ctx, &ctx->mlS, (size_t)(2 * ml.nlev * slen), 0));
    for (int l = 1; l <= ml.nlev; ++l) {
        const size_t len = (size_t)ml.dom[l].nJ * ml.dom[l].nK * g.sJ;
        if (l <= ml.nagg) TRY(dev_alloc(ctx, &ml.r[l], len, 0));
        TRY(dev_alloc(ctx, &ml.z[l], len, 0));
        if (l == 1) { ml.ld[1] = ctx->mlS; ml.ll[1] = ctx->mlS + slen; }   // level 1: the partial sums are the blocks themselves
        else { TRY(dev_alloc(ctx, &ml.ld[l], len, 0)); TRY(dev_alloc(ctx, &ml.ll[l], len, 0)); }
    }
    TRY(dev_alloc(ctx, &ml.part, (size_t)(ctx->sm_count * 2 + 1) * g.sJ, 0));
    TRY(dev_alloc(ctx, &ml.zero, (size_t)g.sJ, 0));
    ctx->ml = ml;
    ctx->ml_slen = slen;
    return PFEM_OK;
}

// tridiagonal blocks of the Galerkin operators of all levels and their L D L^T factors (after launch_diag: dinv is current)
static int ml_setup(pfem_ctx* ctx) {
    const Grid& g = ctx->g;
    const MLDev& ml = ctx->ml;
    const int L = ml.nlev;
    const int koff = ml.dom[0].koff;
    const int nJ1 = (g.nJ + PFEM_ML_C - 1) / PFEM_ML_C, nK1 = (g.nK + koff + PFEM_ML_C - 1) / PFEM_ML_C;
    const LineDom d1{g.nI, nJ1, nK1, g.sJ, g.sJ * nJ1, 0, nK1, 0};
    k_ml_rowsums<<<dim3((unsigned)((g.sJ + 31) / 32), nJ1, nK1), dim3(32, PFEM_ML_C, PFEM_ML_C), 0, ctx->stream>>>(
        g, ctx->cl, ctx->cv, ctx->dinv, L, nJ1, ctx->mlS, ctx->ml_slen, koff);
    KCHECK(); LAUNCHED(1);
    for (int l = 2; l <= L; ++l) {
        const int f = (l == L) ? (1 << 30) : PFEM_ML_C;   // the top level gathers everything
        k_ml_gather<<<dim3((unsigned)((g.sJ + 127) / 128), ml.dom[l].nJ, ml.dom[l].nK), 128, 0, ctx->stream>>>(
            d1, ml.dom[l], f, ctx->mlS + (size_t)(2 * (l - 1)) * ctx->ml_slen, ctx->mlS + (size_t)(2 * (l - 1) + 1) * ctx->ml_slen, ml.ld[l], ml.ll[l]);
        KCHECK(); LAUNCHED(1);
    }
    if (ctx->nranks > 1) {   // the top line is the 1-D problem of the whole device: its blocks are sums over the ranks
        k_ml_top_allreduce<<<1, 256, 0, ctx->stream>>>(ml.ld[L], ml.ll[L], (int)g.sJ, ctx->d_sc);
        KCHECK(); LAUNCHED(1);
    }
    for (int l = 1; l <= L; ++l) {
        const idx_t lines = (idx_t)ml.dom[l].nJ * ml.dom[l].nK;
        k_ml_factor<<<(unsigned)((lines + 63) / 64), 64, 0, ctx->stream>>>(ml.dom[l], ml.ll[l], ml.ld[l], ctx->d_sc);
        KCHECK(); LAUNCHED(1);
    }
    return PFEM_OK;
}

// the two gathers of k_fpcg<MODE 3>
static CoarseAdd ml_coarse_add(const pfem_ctx* ctx) {
    const MLDev& ml = ctx->ml;
    CoarseAdd ca;
    const double* zt = ml.z[ml.nlev];
    const int koff = ml.dom[0].koff;
    if (ml.nagg == 0) ca = CoarseAdd{ml.zero, 1, 31, ml.zero, 1, 31, zt, koff};                                   // top level only
    else if (ml.nagg == 1) ca = CoarseAdd{ml.z[1], ml.dom[1].nJ, PFEM_ML_SHIFT, ml.zero, 1, 31, zt, koff};
    else ca = CoarseAdd{ml.z[1], ml.dom[1].nJ, PFEM_ML_SHIFT, ml.z[2], ml.dom[2].nJ, 2 * PFEM_ML_SHIFT, zt, koff};
    return ca;
}

template <int SEG>
static void launch_ml_chain_seg(pfem_ctx* ctx, const double* r_in, const double* q_in, double* r_out, int mode) {
    const MLDev& ml = ctx->ml;
    const int La = ml.nagg, T = ml.nlev;
    auto top_of = [&](int l) {   // the kernel of the last aggregated level also does the top level
        MLTop t;
        memset(&t, 0, sizeof t);
        if (l != La) return t;
        t.part = ml.part; t.ll = ml.ll[T]; t.ld = ml.ld[T]; t.z = ml.z[T];
        return t;
    };
    k_line_ml<SEG, true><<<ml_blocks(ctx, ml.dom[0]), 256, 0, ctx->stream>>>(ml.dom[0], r_in, q_in, ctx->ll, ctx->ld, r_out, ctx->lz,
                                                                           La >= 1 ? ml.r[1] : nullptr, La >= 1 ? ml.dom[1].nJ : 0, ctx->d_sc,
                                                                           ctx->partials, mode, top_of(0), line_prefetch());
    for (int l = 1; l <= La; ++l)
        k_line_ml<SEG, false><<<ml_blocks(ctx, ml.dom[l]), 256, 0, ctx->stream>>>(ml.dom[l], ml.r[l], nullptr, ml.ll[l], ml.ld[l], nullptr, ml.z[l],
                                                                                l < La ? ml.r[l + 1] : nullptr, l < La ? ml.dom[l + 1].nJ : 0,
                                                                                ctx->d_sc, ctx->partials, mode, top_of(l), line_prefetch());
}

// z_0 .. z_top and the CG scalars of the multilevel preconditioner; mode 1: only b.M^-1 b -> sc->bz; mode 2: z only
static int launch_ml_chain(pfem_ctx* ctx, const double* r_in, const double* q_in, double* r_out, int mode) {
    switch (line_seg(ctx->g)) {
        case 2: launch_ml_chain_seg<2>(ctx, r_in, q_in, r_out, mode); break;
        case 4: launch_ml_chain_seg<4>(ctx, r_in, q_in, r_out, mode); break;
        case 8: launch_ml_chain_seg<8>(ctx, r_in, q_in, r_out, mode); break;
        case 16: launch_ml_chain_seg<16>(ctx, r_in, q_in, r_out, mode); break;
        default: FAIL(PFEM_ERR_STATE, "no warp-per-row line kernel for %d nodes per line", ctx->g.nI);
    }
    KCHECK();
    return PFEM_OK;
}

static int launch_diag(pfem_ctx* ctx) {
    const Grid& g = ctx->g;
    k_diag<<<node_grid(g), node_block(), 0, ctx->stream>>>(g, ctx->cl, ctx->cv, ctx->fixed, ctx->dinv, ctx->d_sc, ctx->op_mass, ctx->op_cmass);
    KCHECK(); LAUNCHED(1);
    if (ctx->surf.nnz) {   // convection adds a face mass matrix (therm3d.cpp:253-257)
        k_surf_diag<<<surf_blocks(ctx), 256, 0, ctx->stream>>>(ctx->surf, ctx->dinv);
        KCHECK(); LAUNCHED(1);
    }
    return PFEM_OK;
}

// out = M (f - A in) [MODE 1/2] or M A in [MODE 3]
template <int MODE>
static int launch_apply_simple(pfem_ctx* ctx, const double* in, double* out, const double* f = nullptr) {
    const Grid& g = ctx->g;
    k_apply_simple<MODE><<<node_grid(g), node_block(), 0, ctx->stream>>>(g, ctx->cl, ctx->cv, in, ctx->dinv, f ? f : ctx->f, out,
                                                                          ctx->d_sc, ctx->partials, ctx->op_mass, ctx->op_cmass);
    KCHECK(); LAUNCHED(1);
    return PFEM_OK;
}

static int kernels_per_iteration(const pfem_ctx* ctx, int variant);

// one PCG iteration's kernels on ctx->stream; returns the number of kernels launched.
// The tiled operator kernel reads p_old from one buffer and writes p_new to the other
// (halo nodes of p_old are read by neighbouring CTAs), so `parity` = iteration & 1 picks them.
static int launch_iteration(pfem_ctx* ctx, int variant, int parity, cudaEvent_t* ev /* 3 events or null */) {
    const Grid& g = ctx->g;
    int launched = 0;
    const double* pin = parity ? ctx->p2 : ctx->p;
    double* pout = parity ? ctx->p : ctx->p2;
    const double* pnew = (variant == 1) ? ctx->p : pout;
    if (ev) cudaEventRecord(ev[0], ctx->stream);
    if (variant == 3 && ctx->precond >= 1) {
        // line-Jacobi PCG: line solve (r in place, z) then the operator step of k_fpcg with z for r and zeros for q;
        // multilevel: the level chain instead of the single line solve, k_fpcg adds the coarse correction while it forms p'
        double* const qq[2] = {ctx->q, ctx->q2};
        double* const pp[2] = {ctx->p, ctx->p2};
        const bool mlp = ctx->precond == 2;
        if (mlp) launch_ml_chain(ctx, ctx->r, qq[parity], ctx->r, 0);
        else launch_line_solve(ctx, ctx->r, qq[parity], ctx->r, 0);
        const PeerOut po = peer_out(ctx, 1 - parity);   // slab mode: p' of the boundary planes goes to the neighbours
        if (mlp && ctx->nranks > 1) {   // z_0 + z_1 + z_2 of my boundary planes -> the neighbours' halo planes of z_0, then a rank barrier
            k_ml_halo<<<dim3((unsigned)((g.nI + 127) / 128), g.nJ), 128, 0, ctx->stream>>>(g, ctx->lz, ml_coarse_add(ctx), po.z_lo, po.z_hi);
            k_rank_barrier<<<1, 32, 0, ctx->stream>>>(ctx->d_sc, 0);
        }
        if (ev) cudaEventRecord(ev[1], ctx->stream);
        if (mlp) launch_fused_dispatch<3>(ctx->line_plan, g, parity, nullptr, qq[1 - parity], pp[1 - parity], ctx->x, ctx->d_sc, ctx->partials,
                                          po, ctx->stream, ml_coarse_add(ctx));
        else launch_fused_dispatch<2>(ctx->line_plan, g, parity, nullptr, qq[1 - parity], pp[1 - parity], ctx->x, ctx->d_sc, ctx->partials,
                                      po, ctx->stream);
        if (ctx->surf_iter) {  // q' += S p' on the boundary rows, alpha from the completed p'.q' (q is only read on owned rows: no push)
            PeerOut none;
            memset(&none, 0, sizeof none);
            k_surf_iter<<<surf_blocks(ctx), 256, 0, ctx->stream>>>(ctx->surf, g, pp[1 - parity], qq[1 - parity], ctx->r, ctx->dinv, ctx->d_sc,
                                                                    ctx->partials, none);
        }
        if (ev) cudaEventRecord(ev[2], ctx->stream);
        return kernels_per_iteration(ctx, 3);
    }
    if (variant == 3) {
        // the whole iteration in one kernel: inputs r,q,p[parity] -> outputs r,q,p[1-parity], x in place
        double* const rr[2] = {ctx->r, ctx->r2};
        double* const qq[2] = {ctx->q, ctx->q2};
        double* const pp[2] = {ctx->p, ctx->p2};
        launch_fused_dispatch<1>(ctx->fused, g, parity, rr[1 - parity], qq[1 - parity], pp[1 - parity], ctx->x, ctx->d_sc,
                                    ctx->partials, peer_out(ctx, 1 - parity), ctx->stream);
        if (ev) cudaEventRecord(ev[1], ctx->stream);
        if (ctx->surf_iter) {   // q' += S p' on the boundary rows, then the alpha / beta step
            k_surf_iter<<<surf_blocks(ctx), 256, 0, ctx->stream>>>(ctx->surf, g, pp[1 - parity], qq[1 - parity], rr[1 - parity], ctx->dinv,
                                                                    ctx->d_sc, ctx->partials, peer_out(ctx, 1 - parity));
            launched = 1;
        }
        if (ev) cudaEventRecord(ev[2], ctx->stream);
        return 1 + launched;
    }
    if (variant == 1) {
        k_pupdate<<<vec_blocks(ctx), 256, 0, ctx->stream>>>(g.NP, ctx->r, ctx->dinv, ctx->p, ctx->d_sc);
        k_apply_simple<0><<<node_grid(g), node_block(), 0, ctx->stream>>>(g, ctx->cl, ctx->cv, ctx->p, ctx->dinv, ctx->f,
                                                                         ctx->q, ctx->d_sc, ctx->partials, ctx->op_mass, ctx->op_cmass);
        launched += 2;
    } else if (variant == 2) {
        launch_apply_tiled(ctx->plan, g, ctx->cl, ctx->cv, ctx->r, ctx->dinv, pin, pout, ctx->q, ctx->d_sc,
                           ctx->partials, ctx->stream);
        launched += 1;
    } else {
        launch_tma_dispatch<true>(ctx->tma, g, ctx->tma.m_p[parity], ctx->dinv, pout, ctx->q, ctx->d_sc, ctx->partials,
                                  ctx->stream);
        launched += 1;
    }
    if (ev) cudaEventRecord(ev[1], ctx->stream);
    k_update<false><<<vec_blocks(ctx), 256, 0, ctx->stream>>>(g.NP, ctx->x, ctx->r, pnew, ctx->q, ctx->dinv, ctx->d_sc,
                                                               ctx->partials);
    launched += 1;
    if (ev) cudaEventRecord(ev[2], ctx->stream);
    return launched;
}

static int kernels_per_iteration(const pfem_ctx* ctx, int variant) {
    if (variant == 3 && ctx->precond == 2)   // level kernels (the top level rides on the last one) + k_fpcg (+ halo push and barrier in slab mode)
        return 2 + ctx->ml.nagg + ctx->surf_iter + (ctx->nranks > 1 ? 2 : 0);
    return variant == 1 ? 3 : (variant == 3 ? (ctx->precond == 1 ? 2 + ctx->surf_iter : 1 + ctx->surf_iter) : 2);
}  // variants 0 (TMA) and 2 (LDG tiled) fuse the p-update

static int build_graph(pfem_ctx* ctx, int batch, int variant, int precond) {
    if (ctx->graph && ctx->graph_batch == batch && ctx->graph_variant == variant && ctx->graph_precond == precond &&
        ctx->graph_surf == ctx->surf_iter && ctx->graph_iso == (int)ctx->fused.iso && ctx->graph_mass == (ctx->op_mass || ctx->op_cmass ? 1 : 0))
        return PFEM_OK;
    if (ctx->graph) { cudaGraphExecDestroy(ctx->graph); ctx->graph = nullptr; }
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    for (int it = 0; it < batch; ++it) launch_iteration(ctx, variant, it & 1, nullptr);
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    if (e != cudaSuccess) FAIL(PFEM_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&ctx->graph, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { ctx->graph = nullptr; FAIL(PFEM_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e)); }
    ctx->graph_batch = batch; ctx->graph_variant = variant; ctx->graph_precond = precond; ctx->graph_surf = ctx->surf_iter; ctx->graph_iso = (int)ctx->fused.iso;
    ctx->graph_mass = (ctx->op_mass || ctx->op_cmass) ? 1 : 0;
    return PFEM_OK;
}

static int read_scalars(pfem_ctx* ctx) {
    CU(cudaMemcpyAsync(ctx->h_sc, ctx->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

// Prepare the PCG state from the current x / conds: dinv, ||b_free||^2, r0, rho0.
static int pcg_prepare(pfem_ctx* ctx, const pfem_opts* o, int bench) {
    const Grid& g = ctx->g;
    if (!ctx->conds_valid) FAIL(PFEM_ERR_STATE, "conductivities have not been computed");
    TRY(mask_field(ctx));
    double tol2 = bench ? -1. : o->lin_tol * o->lin_tol;
    if (!ctx->surf_iter_known) {             // convection terms on ANY rank: every rank runs k_surf_iter (it is collective)
        int any = ctx->surf.nnz > 0;
        if (ctx->nranks > 1) TRY(rank_barrier(ctx, any, &any));
        ctx->surf_iter = any ? 1 : 0;
        ctx->surf_iter_known = true;
    }
    if (ctx->surf_iter && o->variant != 3) FAIL(PFEM_ERR_BAD_INPUT, "convection boundary terms run with the fused PCG kernel only (variant 3)");
    if (o->precond >= 1) TRY(ensure_line(ctx));
    if (o->precond == 2) TRY(ensure_ml(ctx));
    ctx->precond = o->precond;
    {   // isotropic conductivities: the iteration kernels do not stream c_vert (PFEM_NO_ISO=1 keeps the general kernel, for A/B runs)
        static const bool no_iso = getenv("PFEM_NO_ISO") != nullptr;
        ctx->fused.iso = ctx->cond_iso && !no_iso;
        ctx->line_plan.iso = ctx->fused.iso;
        ctx->fused.mass = ctx->op_mass;           // Dynamic3D time step: operator + lumped capacity diagonal
        ctx->line_plan.mass = ctx->op_mass;
    }
    k_set_params<<<1, 1, 0, ctx->stream>>>(ctx->d_sc, tol2, o->maxit, bench, ctx->surf_iter, o->precond);
    LAUNCHED(1);
    TRY(launch_diag(ctx));
    TRY(halo_sync(ctx, SA_DINV));            // slab mode: the diagonal of a halo plane needs the neighbour's elements
    TRY(surf_eval_rad(ctx));                 // radiation load from the temperatures of the previous loop (therm3d.cpp:262-267)
    TRY(scatter_bc(ctx, ctx->x, nullptr, true));   // x_D = v_D as the free rows see it (first condition per node, iterative_matrix.hpp:466-484)
    // ||b_free||^2 with b_free = M (f - A x_D): q <- x_D, p <- b_free (scratch)
    k_dirichlet_only<<<vec_blocks(ctx), 256, 0, ctx->stream>>>(g.NP, ctx->x, ctx->fixed, ctx->q);
    LAUNCHED(1);
    const double* feff = nullptr;
    TRY(surf_rhs(ctx, ctx->q, &feff));
    TRY(launch_apply_simple<2>(ctx, ctx->q, ctx->p, feff));
    if (o->precond >= 1) {   // factors of the line blocks, then ||b_free|| in the metric of this preconditioner
        k_line_factor<<<(unsigned)((g.N / (g.vdim == 0 ? g.nI : g.vdim == 1 ? g.nJ : g.nK) + 127) / 128), 128, 0, ctx->stream>>>(
            g, ctx->cl, ctx->cv, ctx->dinv, ctx->ll, ctx->ld, ctx->lmask, ctx->d_sc);
        KCHECK(); LAUNCHED(1);
        if (o->precond == 2) { TRY(ml_setup(ctx)); TRY(launch_ml_chain(ctx, ctx->p, nullptr, nullptr, 1)); LAUNCHED(1 + ctx->ml.nagg); }
        else { TRY(launch_line_solve(ctx, ctx->p, nullptr, nullptr, 1)); LAUNCHED(1); }
    }
    // r0 = M (f - A x)
    TRY(surf_rhs(ctx, ctx->x, &feff));
    TRY(launch_apply_simple<1>(ctx, ctx->x, ctx->r, feff));
    TRY(halo_sync(ctx, SA_R));
    if (ctx->bc_val_first) TRY(scatter_bc(ctx, ctx->x, nullptr));   // B[r] = val of the LAST condition (:463-464); fixed rows stay put in CG
    CU(cudaMemsetAsync(ctx->p, 0, (size_t)g.NP * sizeof(double), ctx->stream));
    CU(cudaMemsetAsync(ctx->q, 0, (size_t)g.NP * sizeof(double), ctx->stream));   // fused kernel: r' = r - 0*q on the first launch
    {   // initial scalars over the OWNED planes (all planes unless this is a slab)
        const idx_t off = g.sK * g.kown0, cnt = g.sK * (g.kown1 - g.kown0);
        k_update<true><<<vec_blocks(ctx), 256, 0, ctx->stream>>>(cnt, ctx->x + off, ctx->r + off, ctx->p + off, ctx->q + off,
                                                                  ctx->dinv + off, ctx->d_sc, ctx->partials);
    }
    LAUNCHED(1);
    KCHECK();
    return PFEM_OK;
}

static int pcg_solve(pfem_ctx* ctx, const pfem_opts* o, int* iters, double* relres, int* converged) {
    TRY(pcg_prepare(ctx, o, 0));
    int batch = o->batch > 0 ? o->batch : 32;
    batch += batch & 1;  // even: p ping-pong parity is preserved across graph launches
    TRY(build_graph(ctx, batch, o->variant, o->precond));
    TRY(read_scalars(ctx));
    int neg_diag = ctx->h_sc->neg_diag;
    if (ctx->nranks > 1) { TRY(rank_barrier(ctx, neg_diag, &neg_diag)); TRY(read_scalars(ctx)); }   // all ranks fail together
    if (neg_diag) FAIL(PFEM_ERR_NOT_SPD, "nonpositive diagonal element in stiffness matrix");
    while (!ctx->h_sc->done) {
        CU(cudaGraphLaunch(ctx->graph, ctx->stream));
        LAUNCHED((long long)batch * kernels_per_iteration(ctx, o->variant));
        TRY(read_scalars(ctx));
    }
    const Scalars& s = *ctx->h_sc;
    *iters = s.iter;
    *relres = (s.bb > 0.) ? sqrt(s.rr / s.bb) : sqrt(s.rr);
    ctx->last_relres_pre = (s.bz > 0.) ? sqrt(s.rho / s.bz) : sqrt(s.rho);
    *converged = (s.status == 1);
    if (s.status == -1) FAIL(PFEM_ERR_NOT_SPD, "p.Ap = %g <= 0 at iteration %d: stiffness matrix is not positive definite", s.pq, s.iter);
    if (s.status == -2) FAIL(PFEM_ERR_NAN, "non-finite value in the PCG iteration %d", s.iter);
    if (s.status == -3) FAIL(PFEM_ERR_CUDA, "slab mode: timed out waiting for a peer rank in PCG iteration %d", s.iter);
    return PFEM_OK;
}

static int check_opts(pfem_ctx* ctx, const pfem_opts* o) {
    if (!o) FAIL(PFEM_ERR_BAD_INPUT, "null options");
    if (o->maxit <= 0) FAIL(PFEM_ERR_BAD_INPUT, "maxit must be positive");
    if (!(o->lin_tol > 0.)) FAIL(PFEM_ERR_BAD_INPUT, "lin_tol must be positive");
    if (o->precond < 0 || o->precond > 2) FAIL(PFEM_ERR_BAD_INPUT, "preconditioner %d is not implemented (0 = Jacobi, 1 = line-Jacobi, 2 = multilevel line)", o->precond);
    if (o->precond >= 1 && o->variant != 3) FAIL(PFEM_ERR_BAD_INPUT, "the line preconditioner runs with kernel variant 3 only");
    if (o->precond >= 1 && ctx->nranks > 1 && ctx->g.vdim == 2)
        FAIL(PFEM_ERR_BAD_INPUT, "slab mode: the line preconditioner needs the vertical axis inside the slabs (a lateral major axis)");
    if (o->variant < 0 || o->variant > 3) FAIL(PFEM_ERR_BAD_INPUT, "unknown kernel variant %d", o->variant);
    if ((ctx->op_mass || ctx->op_cmass) && o->precond == 2) FAIL(PFEM_ERR_BAD_INPUT, "Dynamic3D: the multilevel preconditioner is not available (use jac or ljac)");
    if (ctx->op_mass && o->variant != 3 && o->variant != 1) FAIL(PFEM_ERR_BAD_INPUT, "Dynamic3D runs kernel variant 3 (fused) or 1 (simple)");
    if (ctx->op_cmass && (o->variant != 1 || o->precond != 0))
        FAIL(PFEM_ERR_BAD_INPUT, "Dynamic3D with a consistent capacity matrix (lumping = no) runs kernel variant 1 with the Jacobi preconditioner only");
    if (o->variant == 3 && !ctx->fused.valid) FAIL(PFEM_ERR_STATE, "fused PCG kernel unavailable: %s", ctx->fused.why);
    if (ctx->nranks > 1 && o->variant != 3) FAIL(PFEM_ERR_BAD_INPUT, "slab mode runs the fused PCG kernel only (variant 3)");
    if (o->variant == 0 && !ctx->tma.valid) FAIL(PFEM_ERR_STATE, "TMA operator kernel unavailable: %s", ctx->tma.why);
    if (o->variant == 2 && !ctx->plan.valid) FAIL(PFEM_ERR_STATE, "no tiled kernel plan for this mesh");
    return PFEM_OK;
}

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t s;
    explicit Timer(cudaStream_t st) : s(st) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s); }
    double stop() { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
    ~Timer() { cudaEventDestroy(a); cudaEventDestroy(b); }
};

extern "C" int pfem_solve_linear(pfem_ctx* ctx, const pfem_opts* o, pfem_stats* st) {
    NEED_MESH();
    TRY(check_opts(ctx, o));
    long long l0 = ctx->launches;
    Timer t(ctx->stream);
    int iters = 0, conv = 0;
    double relres = 0;
    TRY(pcg_solve(ctx, o, &iters, &relres, &conv));
    TRY(halo_sync(ctx, SA_X));
    if (st) {
        memset(st, 0, sizeof(*st));
        st->lin_iters = st->last_iters = iters;
        st->converged = conv; st->lin_relres = relres; st->loopno = ctx->loopno;
        st->lin_relres_precond = ctx->last_relres_pre;
        st->t_solve_ms = t.stop();
        st->kernel_launches = ctx->launches - l0;
    }
    return conv ? PFEM_OK : PFEM_NOT_CONVERGED;
}

// ------------------------------------------------------------------ nonlinear loops -----

extern "C" int pfem_solve_thermal(pfem_ctx* ctx, const pfem_opts* o, pfem_stats* st) {
    NEED_MESH();
    TRY(check_opts(ctx, o));
    if (!ctx->have_materials) FAIL(PFEM_ERR_STATE, "pfem_set_materials has not been called");
    const Grid& g = ctx->g;
    long long l0 = ctx->launches;
    Timer t(ctx->stream);
    int loop = 0, conv = 1, iters = 0;
    long long total_iters = 0;
    double err = 0., toterr = 0., maxT = 0., relres = 0.;
    TRY(mask_field(ctx));
    const int cap = o->loops > 0 ? o->loops : 100000;
    const idx_t own_off = g.sK * g.kown0, own_cnt = g.sK * (g.kown1 - g.kown0);
    do {
        TRY(halo_sync(ctx, SA_X));                             // slab mode: the solve updates owned planes only
        CU(cudaMemcpyAsync(ctx->xprev, ctx->x, (size_t)g.NP * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        TRY(pfem_update_conductivity_thermal(ctx));           // therm3d.cpp:204-220
        TRY(pcg_solve(ctx, o, &iters, &relres, &conv));        // setMatrix + A.solve, :314-315
        total_iters += iters;
        k_thermal_error<<<vec_blocks(ctx), 256, 0, ctx->stream>>>(own_cnt, ctx->x + own_off, ctx->xprev + own_off, ctx->d_sc,
                                                                   ctx->partials);
        KCHECK(); LAUNCHED(1);
        TRY(read_scalars(ctx));
        err = ctx->h_sc->red[0];
        maxT = ctx->h_sc->red[1];
        if (err > toterr) toterr = err;
        ++ctx->loopno;
        ++loop;
    } while ((!conv || err > o->outer_tol) && loop < cap);      // :334
    TRY(halo_sync(ctx, SA_X));
    if (st) {
        memset(st, 0, sizeof(*st));
        st->outer_loops = loop; st->loopno = ctx->loopno; st->lin_iters = total_iters; st->last_iters = iters;
        st->converged = conv; st->lin_relres = relres; st->err = err; st->toterr = toterr; st->maxval = maxT;
        st->lin_relres_precond = ctx->last_relres_pre;
        st->t_solve_ms = t.stop();
        st->kernel_launches = ctx->launches - l0;
    }
    return conv ? PFEM_OK : PFEM_NOT_CONVERGED;
}

// ------------------------------------------------------------------ Dynamic3D ------------

extern "C" int pfem_set_capacity(pfem_ctx* ctx, uint32_t nmat, uint32_t nT, const double* cp_dens) {
    NEED_MESH();
    if (!ctx->have_materials) FAIL(PFEM_ERR_STATE, "pfem_set_materials has not been called");
    if (!cp_dens) FAIL(PFEM_ERR_BAD_INPUT, "null capacity table");
    if (nmat != ctx->nmat || nT != ctx->nT) FAIL(PFEM_ERR_BAD_INPUT, "the capacity table must have the shape [nmat][nT] of pfem_set_materials (%u x %u)", ctx->nmat, ctx->nT);
    const size_t cnt = (size_t)nmat * nT;
    for (size_t a = 0; a < cnt; ++a)
        if (!(cp_dens[a] > 0.)) FAIL(PFEM_ERR_BAD_INPUT, "heat capacity table entry %zu is not positive", a);
    dev_release(ctx, &ctx->tab_cprho);
    TRY(dev_alloc(ctx, &ctx->tab_cprho, cnt, 0));
    CU(cudaMemcpyAsync(ctx->tab_cprho, cp_dens, cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->cap_nmat = nmat; ctx->cap_nT = nT;
    return PFEM_OK;
}

// setMatrix of DynamicThermalFem3DSolver (femT3d.cpp:127-255) in matrix-free form: conductivities and capacities at the current
// temperatures; ctx->cl/cv become methodparam * k (the stiffness part of A), the unscaled values go to dyn_cl/dyn_cv (the K of
// B = A - K); the capacity becomes a node diagonal (lumped) or stays an element array (consistent).
static int dynamic_set_matrix(pfem_ctx* ctx, const pfem_dynamic* d) {
    const Grid& g = ctx->g;
    TRY(pfem_update_conductivity_thermal(ctx));
    CU(cudaMemcpyAsync(ctx->dyn_cl, ctx->cl, (size_t)g.NP * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->dyn_cv, ctx->cv, (size_t)g.NP * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    k_scale2<<<vec_blocks(ctx), 256, 0, ctx->stream>>>(g.NP, d->methodparam, ctx->dyn_cl, ctx->dyn_cv, ctx->cl, ctx->cv);
    KCHECK(); LAUNCHED(1);
    k_elem_capacity<<<node_grid(g), node_block(), 0, ctx->stream>>>(g, ctx->x, ctx->mat, ctx->nT, ctx->T0, ctx->dT, ctx->tab_cprho,
                                                                     1. / d->timestep, ctx->dyn_ce);
    KCHECK(); LAUNCHED(1);
    if (d->lumping) {
        k_mass_lumped<<<node_grid(g), node_block(), 0, ctx->stream>>>(g, ctx->dyn_ce, ctx->dyn_mass);
        KCHECK(); LAUNCHED(1);
        ctx->op_mass = ctx->dyn_mass; ctx->op_cmass = nullptr;
    } else {
        ctx->op_mass = nullptr; ctx->op_cmass = ctx->dyn_ce;
    }
    return PFEM_OK;
}

static int dynamic_steps(pfem_ctx* ctx, const pfem_opts* o, const pfem_dynamic* d, pfem_stats* st, double* f_static) {
    const Grid& g = ctx->g;
    int steps = 0, conv = 1, iters = 0, all_conv = 1;
    long long total_iters = 0;
    double relres = 0., maxT = 0.;
    const idx_t own_off = g.sK * g.kown0, own_cnt = g.sK * (g.kown1 - g.kown0);   // slab mode: reductions over the owned planes
    TRY(mask_field(ctx));
    TRY(halo_sync(ctx, SA_X));                                     // slab mode: conductivities / capacities of the boundary elements
    TRY(dynamic_set_matrix(ctx, d));
    long long r = d->rebuildfreq;
    const double tend = d->time + d->timestep / 2.;
    for (double t = 0.; t < tend; t += d->timestep) {               // femT3d.cpp:271-272
        if (steps > 0) TRY(halo_sync(ctx, SA_X));                   // slab mode: the solve updates owned planes only
        if (d->rebuildfreq && r == 0) { TRY(dynamic_set_matrix(ctx, d)); r = d->rebuildfreq; }   // :274-278
        // right-hand side B T + F with B = A - K:  M (F - K T) + M (A T), then the Dirichlet rows take their values in the solve
        TRY(launch_diag(ctx));                                      // the row mask (dinv == 0) of the kernels below
        k_apply_simple<1><<<node_grid(g), node_block(), 0, ctx->stream>>>(g, ctx->dyn_cl, ctx->dyn_cv, ctx->x, ctx->dinv, f_static, ctx->p,
                                                                          ctx->d_sc, ctx->partials, nullptr, nullptr);
        KCHECK(); LAUNCHED(1);
        TRY(launch_apply_simple<3>(ctx, ctx->x, ctx->q));
        k_axpby<<<vec_blocks(ctx), 256, 0, ctx->stream>>>(g.NP, ctx->p, ctx->q, ctx->dyn_f);
        KCHECK(); LAUNCHED(1);
        ctx->f = ctx->dyn_f;
        const int rc = pcg_solve(ctx, o, &iters, &relres, &conv);   // A.solverhs(T, temperatures): warm start from T^n (:283)
        ctx->f = f_static;
        if (rc != PFEM_OK) return rc;
        total_iters += iters;
        all_conv = all_conv && conv;
        ++steps; --r;
        if (d->maxT_log && (size_t)(steps - 1) < d->maxT_log_len) {
            k_thermal_error<<<vec_blocks(ctx), 256, 0, ctx->stream>>>(own_cnt, ctx->x + own_off, ctx->x + own_off, ctx->d_sc, ctx->partials);
            KCHECK(); LAUNCHED(1);
            TRY(read_scalars(ctx));
            d->maxT_log[steps - 1] = ctx->h_sc->red[1];
        }
    }
    k_thermal_error<<<vec_blocks(ctx), 256, 0, ctx->stream>>>(own_cnt, ctx->x + own_off, ctx->x + own_off, ctx->d_sc, ctx->partials);
    KCHECK(); LAUNCHED(1);
    TRY(read_scalars(ctx));
    maxT = ctx->h_sc->red[1];
    TRY(halo_sync(ctx, SA_X));
    if (st) {
        st->outer_loops = steps; st->loopno = ctx->loopno; st->lin_iters = total_iters; st->last_iters = iters;
        st->converged = all_conv; st->lin_relres = relres; st->maxval = maxT;
        st->lin_relres_precond = ctx->last_relres_pre;
    }
    return all_conv ? PFEM_OK : PFEM_NOT_CONVERGED;
}

// The time loop of DynamicThermalFem3DSolver::compute (femT3d.cpp:258-305), corrected specification:
//   (methodparam K + C/dt) T^{n+1} = (C/dt - (1 - methodparam) K) T^n + F   on the free rows,   T^{n+1} = value on Dirichlet rows.
extern "C" int pfem_solve_dynamic(pfem_ctx* ctx, const pfem_opts* o, const pfem_dynamic* d, pfem_stats* st) {
    NEED_MESH();
    if (!d) FAIL(PFEM_ERR_BAD_INPUT, "null dynamic parameters");
    if (!ctx->have_materials) FAIL(PFEM_ERR_STATE, "pfem_set_materials has not been called");
    if (!ctx->tab_cprho) FAIL(PFEM_ERR_STATE, "pfem_set_capacity has not been called");
    if (!(d->timestep > 0.) || !(d->time >= 0.)) FAIL(PFEM_ERR_BAD_INPUT, "need timestep > 0 and time >= 0");
    if (!(d->methodparam >= 0. && d->methodparam <= 1.)) FAIL(PFEM_ERR_BAD_INPUT, "methodparam must lie in [0, 1]");
    if (d->rebuildfreq < 0) FAIL(PFEM_ERR_BAD_INPUT, "negative rebuildfreq");
    // slab mode: collective; every rank must pass the same time span, step, rebuild frequency and maxT_log_len (the per-step maximum
    // is a cross-rank reduction); the consistent capacity matrix runs the node-per-thread kernel, which is single-GPU
    if (ctx->nranks > 1 && !d->lumping) FAIL(PFEM_ERR_BAD_INPUT, "slab mode: Dynamic3D runs with a lumped capacity matrix only");
    if (ctx->surf.nrows) FAIL(PFEM_ERR_BAD_INPUT, "Dynamic3D has boundary conditions of the first kind only (femT3d.hpp)");
    const Grid& g = ctx->g;
    if (!ctx->dyn_f) {
        TRY(dev_alloc(ctx, &ctx->dyn_mass, (size_t)g.NP, (size_t)g.G));
        TRY(dev_alloc(ctx, &ctx->dyn_ce, (size_t)g.NP, (size_t)g.G));
        TRY(dev_alloc(ctx, &ctx->dyn_cl, (size_t)g.NP, (size_t)g.G));
        TRY(dev_alloc(ctx, &ctx->dyn_cv, (size_t)g.NP, (size_t)g.G));
        TRY(dev_alloc(ctx, &ctx->dyn_f, (size_t)g.NP, (size_t)g.G));
    }
    // the operator extension must be known to check_opts
    if (d->lumping) { ctx->op_mass = ctx->dyn_mass; ctx->op_cmass = nullptr; } else { ctx->op_mass = nullptr; ctx->op_cmass = ctx->dyn_ce; }
    int rc = check_opts(ctx, o);
    long long l0 = ctx->launches;
    if (st) memset(st, 0, sizeof(*st));
    double* const f_static = ctx->f;
    if (rc == PFEM_OK) {
        Timer t(ctx->stream);
        rc = dynamic_steps(ctx, o, d, st, f_static);
        if (st) { st->t_solve_ms = t.stop(); st->kernel_launches = ctx->launches - l0; }
    }
    // back to the static state: true conductivities in cl/cv (providers read them), no operator extension
    ctx->f = f_static;
    ctx->op_mass = nullptr; ctx->op_cmass = nullptr;
    ctx->fused.mass = nullptr; ctx->line_plan.mass = nullptr;
    if (ctx->conds_valid && ctx->dyn_cl) {
        cudaMemcpyAsync(ctx->cl, ctx->dyn_cl, (size_t)g.NP * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
        cudaMemcpyAsync(ctx->cv, ctx->dyn_cv, (size_t)g.NP * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
    }
    return rc;
}

extern "C" int pfem_solve_shockley(pfem_ctx* ctx, const pfem_opts* o, pfem_stats* st) {
    NEED_MESH();
    TRY(check_opts(ctx, o));
    if (!ctx->have_materials) FAIL(PFEM_ERR_STATE, "pfem_set_materials has not been called");
    const Grid& g = ctx->g;
    // slab mode: a junction must not be cut by the partition, i.e. the slab (major) axis has to be lateral
    if (ctx->nranks > 1 && ctx->nact && g.vdim == 2)
        FAIL(PFEM_ERR_BAD_INPUT, "slab mode with junctions needs a lateral major axis (iteration order 0xx or 1xx)");
    TRY(ensure_elem_arrays(ctx, true));
    long long l0 = ctx->launches;
    Timer t(ctx->stream);
    TRY(mask_field(ctx));
    TRY(pfem_update_conductivity_shockley(ctx));               // loadConductivity, electr3d.cpp:378
    const int noactive = (ctx->nact == 0);
    const double minj = 100e-7;                                 // :381
    int loop = 0, conv = 1, iters = 0;
    long long total_iters = 0;
    double err = 0., toterr = 0., mcur = 0., relres = 0.;
    const int cap = o->loops > 0 ? o->loops : 100000;
    TRY(halo_sync(ctx, SA_X));                                  // slab mode: the potential of the halo planes
    do {
        if (ctx->loopno != 0 && ctx->nact) {                    // :246-274
            k_junction_update<<<node_grid(g), node_block(), 0, ctx->stream>>>(g, ctx->junc, ctx->act, ctx->x, ctx->beta_col,
                                                                               ctx->js_col, ctx->stable, ctx->cl, ctx->cv);
            KCHECK(); LAUNCHED(1);
        }
        TRY(pcg_solve(ctx, o, &iters, &relres, &conv));        // assembly + applyBC + solve, :281-344,385
        total_iters += iters;
        TRY(halo_sync(ctx, SA_X));                             // slab mode: the solve updates owned planes only
        k_currents<<<node_grid(g), node_block(), 0, ctx->stream>>>(g, ctx->x, ctx->cl, ctx->cv, ctx->junc, noactive,
                                                                    ctx->cur0, ctx->cur1, ctx->cur2, ctx->d_sc,
                                                                    ctx->partials, ctx->partial_idx);
        k_fetch_maxcur<<<1, 32, 0, ctx->stream>>>(g, ctx->cur0, ctx->cur1, ctx->cur2, ctx->d_sc);
        KCHECK(); LAUNCHED(2);
        TRY(read_scalars(ctx));
        mcur = sqrt(ctx->h_sc->red[1]);
        err = 100. * sqrt(ctx->h_sc->red[0]) / (mcur > minj ? mcur : minj);   // :423-424
        if ((loop != 0 || mcur >= minj) && err > toterr) toterr = err;         // :425
        ++ctx->loopno;
        ++loop;
    } while ((!conv || err > o->outer_tol) && loop < cap);      // :433
    if (ctx->nact) {                                            // saveConductivity, :435
        k_junction_save<<<64, 256, 0, ctx->stream>>>(g, ctx->act, ctx->nact, ctx->cl, ctx->cv, ctx->junc_cond);
        KCHECK(); LAUNCHED(1);
    }
    if (st) {
        memset(st, 0, sizeof(*st));
        st->outer_loops = loop; st->loopno = ctx->loopno; st->lin_iters = total_iters; st->last_iters = iters;
        st->converged = conv; st->lin_relres = relres; st->err = err; st->toterr = toterr; st->maxval = mcur;
        st->lin_relres_precond = ctx->last_relres_pre;
        for (int a = 0; a < 3; ++a) st->maxcur[a] = ctx->h_sc->maxcur[a];
        st->t_solve_ms = t.stop();
        st->kernel_launches = ctx->launches - l0;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return conv ? PFEM_OK : PFEM_NOT_CONVERGED;
}

// ------------------------------------------------------------------------ results -------

extern "C" int pfem_get_field(pfem_ctx* ctx, double* x) {
    NEED_MESH();
    if (!x) FAIL(PFEM_ERR_BAD_INPUT, "null output");
    TRY(mask_field(ctx));
    CU(download_nodes(ctx, x, ctx->x));
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

extern "C" int pfem_get_elem(pfem_ctx* ctx, int what, const uint8_t* noheat, double* out) {
    NEED_MESH();
    if (!out) FAIL(PFEM_ERR_BAD_INPUT, "null output");
    const Grid& g = ctx->g;
    switch (what) {
        case PFEM_ELEM_COND:
            if (!ctx->conds_valid) FAIL(PFEM_ERR_STATE, "conductivities have not been computed");
            return download_elem<double, 2>(ctx, ctx->cl, ctx->cv, nullptr, out);
        case PFEM_ELEM_CURRENT:
            if (!ctx->cur0) FAIL(PFEM_ERR_STATE, "no current densities: pfem_solve_shockley has not run");
            return download_elem<double, 3>(ctx, ctx->cur0, ctx->cur1, ctx->cur2, out);
        case PFEM_ELEM_HEAT: {
            if (!ctx->conds_valid) FAIL(PFEM_ERR_STATE, "conductivities have not been computed");
            TRY(ensure_elem_arrays(ctx, false));
            if (noheat) TRY(pfem_set_noheat(ctx, noheat));
            const uint8_t* nh = ctx->noheat_set ? ctx->noheat : nullptr;
            k_gradient_fields<true><<<node_grid(g), node_block(), 0, ctx->stream>>>(g, ctx->x, ctx->cl, ctx->cv, nh, ctx->aux0,
                                                                                     ctx->aux1, ctx->aux2);
            KCHECK(); LAUNCHED(1);
            return download_elem<double, 1>(ctx, ctx->aux0, nullptr, nullptr, out);
        }
        case PFEM_ELEM_FLUX: {
            if (!ctx->have_materials) FAIL(PFEM_ERR_STATE, "pfem_set_materials has not been called");
            TRY(ensure_elem_arrays(ctx, false));
            // saveHeatFluxes re-evaluates thermk at the CURRENT temperatures (therm3d.cpp:362-370)
            TRY(pfem_update_conductivity_thermal(ctx));
            k_gradient_fields<false><<<node_grid(g), node_block(), 0, ctx->stream>>>(g, ctx->x, ctx->cl, ctx->cv, nullptr,
                                                                                      ctx->aux0, ctx->aux1, ctx->aux2);
            KCHECK(); LAUNCHED(1);
            return download_elem<double, 3>(ctx, ctx->aux0, ctx->aux1, ctx->aux2, out);
        }
        default: FAIL(PFEM_ERR_BAD_INPUT, "unknown element field %d", what);
    }
}

extern "C" int pfem_get_junction_cond(pfem_ctx* ctx, double* junc_cond) {
    NEED_MESH();
    if (!ctx->nact || !junc_cond) FAIL(PFEM_ERR_BAD_INPUT, "no junctions / null output");
    CU(cudaMemcpyAsync(junc_cond, ctx->junc_cond, 2 * ctx->ncol * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

extern "C" int pfem_get_elem_temperature(pfem_ctx* ctx, size_t n, const size_t* elem, double* T_out) {
    NEED_MESH();
    if (n && (!elem || !T_out)) FAIL(PFEM_ERR_BAD_INPUT, "null argument");
    if (!ctx->Te) FAIL(PFEM_ERR_STATE, "no element temperatures: pfem_set_elem_temperature / pfem_transfer_temperature has not been called");
    if (!n) return PFEM_OK;
    const Grid& g = ctx->g;
    int ord[3] = {0, 1, 2};
    std::sort(ord, ord + 3, [&](int a, int b) { return g.es[a] > g.es[b]; });   // major, medium, minor element axis
    std::vector<idx_t> slot(n);
    for (size_t m = 0; m < n; ++m) {
        if (elem[m] >= (size_t)g.E) FAIL(PFEM_ERR_BAD_INPUT, "element %zu out of range", elem[m]);
        idx_t rem = (idx_t)elem[m], s = 0;
        for (int q = 0; q < 3; ++q) { s += g.ps[ord[q]] * (rem / g.es[ord[q]]); rem %= g.es[ord[q]]; }
        slot[m] = s;    // an element lives at the lattice index of its lowest corner node
    }
    const size_t bytes = n * (sizeof(idx_t) + sizeof(double));
    TRY(ensure_stage(ctx, bytes));
    idx_t* d_slot = reinterpret_cast<idx_t*>(ctx->stage);
    double* d_out = reinterpret_cast<double*>(d_slot + n);
    CU(cudaMemcpyAsync(d_slot, slot.data(), n * sizeof(idx_t), cudaMemcpyHostToDevice, ctx->stream));
    k_gather<<<(unsigned)std::min<size_t>((n + 255) / 256, 1024), 256, 0, ctx->stream>>>(n, d_slot, ctx->Te, d_out);
    KCHECK(); LAUNCHED(1);
    CU(cudaMemcpyAsync(T_out, d_out, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

// ------------------------------------------------------------------ field exchange -------

extern "C" int pfem_set_noheat(pfem_ctx* ctx, const uint8_t* noheat) {
    NEED_MESH();
    if (!noheat) { ctx->noheat_set = false; return PFEM_OK; }
    if (!ctx->noheat) TRY(dev_alloc(ctx, &ctx->noheat, (size_t)ctx->g.NP, (size_t)ctx->g.G));
    TRY(upload_elem<uint8_t, 1>(ctx, noheat, ctx->noheat, nullptr, nullptr));
    ctx->noheat_set = true;
    return PFEM_OK;
}

// Bracketing tables of prepareInterpolationForAxis (plask/mesh/axis1d.cpp:99-156, no symmetry / periodicity) for the
// midpoints of `dst_ax` in the source axis `src` (node coordinates, or their midpoints when src_mid).
struct AxisTab { std::vector<int> ilo, ihi; std::vector<double> lo, hi, pt; };
// dst_mid: the targets are the midpoints of dst_ax (element mesh of a context), else the points of dst_ax themselves
static AxisTab make_axis_tab(const std::vector<double>& src_nodes, bool src_mid, const std::vector<double>& dst_ax, bool dst_mid = true) {
    std::vector<double> mid;
    if (src_mid) { mid.resize(src_nodes.size() - 1); for (size_t i = 0; i + 1 < src_nodes.size(); ++i) mid[i] = (src_nodes[i] + src_nodes[i + 1]) * 0.5; }
    const std::vector<double>& src = src_mid ? mid : src_nodes;
    const size_t n = src.size(), m = dst_mid ? dst_ax.size() - 1 : dst_ax.size();
    AxisTab t;
    t.ilo.resize(m); t.ihi.resize(m); t.lo.resize(m); t.hi.resize(m); t.pt.resize(m);
    for (size_t j = 0; j < m; ++j) {
        const double p = dst_mid ? (dst_ax[j] + dst_ax[j + 1]) * 0.5 : dst_ax[j];   // MidpointAxis::at, axis1d.cpp:51-53
        size_t up = std::upper_bound(src.begin(), src.end(), p) - src.begin();
        size_t ilo, ihi = up; double lo, hi;
        if (up == 0) { ilo = 0; lo = src[0] - 1.; } else { ilo = up - 1; lo = src[up - 1]; }
        if (up == n) { ihi = n - 1; hi = src[n - 1] + 1.; } else hi = src[up];
        t.ilo[j] = (int)ilo; t.ihi[j] = (int)ihi; t.lo[j] = lo; t.hi[j] = hi; t.pt[j] = p;
    }
    return t;
}

// dst_arr (element lattice of dst) <- linear interpolation of src_arr (node or element lattice of src)
static int interp_to_elems(pfem_ctx* dst, pfem_ctx* src, const double* src_arr, bool src_is_elem, double* dst_arr, const InterpOutside& outside) {
    pfem_ctx* ctx = dst;
    if (src->device != dst->device) FAIL(PFEM_ERR_BAD_INPUT, "field exchange needs both contexts on the same device");
    if (src->nranks != dst->nranks || src->rank != dst->rank) FAIL(PFEM_ERR_BAD_INPUT, "field exchange needs the same slab partition on both sides");
    AxisTab tab[3];
    size_t ni = 0, nd = 0;
    for (int a = 0; a < 3; ++a) {
        if (src_is_elem && src->hax[a].size() < 2) FAIL(PFEM_ERR_STATE, "source mesh not set");
        tab[a] = make_axis_tab(src->hax[a], src_is_elem, dst->hax[a]);
        ni += 2 * tab[a].ilo.size(); nd += 3 * tab[a].lo.size();
    }
    // one staging buffer: doubles first (8-byte aligned), then ints
    const size_t bytes = nd * sizeof(double) + ni * sizeof(int);
    std::vector<unsigned char> hb(bytes);
    double* hd = reinterpret_cast<double*>(hb.data());
    int* hi = reinterpret_cast<int*>(hb.data() + nd * sizeof(double));
    TRY(ensure_itab(dst, bytes));
    void* const dbuf = dst->itab;
    double* dd = reinterpret_cast<double*>(dbuf);
    int* di = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(dbuf) + nd * sizeof(double));
    InterpAxis ia[3];
    size_t od = 0, oi = 0;
    for (int a = 0; a < 3; ++a) {
        const size_t m = tab[a].lo.size();
        memcpy(hd + od, tab[a].lo.data(), m * 8); ia[a].lo = dd + od; od += m;
        memcpy(hd + od, tab[a].hi.data(), m * 8); ia[a].hi = dd + od; od += m;
        memcpy(hd + od, tab[a].pt.data(), m * 8); ia[a].pt = dd + od; od += m;
        memcpy(hi + oi, tab[a].ilo.data(), m * 4); ia[a].ilo = di + oi; oi += m;
        memcpy(hi + oi, tab[a].ihi.data(), m * 4); ia[a].ihi = di + oi; oi += m;
    }
    cudaError_t e = cudaStreamSynchronize(src->stream);            // the source field is final
    if (e == cudaSuccess) e = cudaMemcpyAsync(dbuf, hb.data(), bytes, cudaMemcpyHostToDevice, dst->stream);
    if (e == cudaSuccess) {
        const Grid& gs = src->g;
        k_interp_to_elems<<<node_grid(dst->g), node_block(), 0, dst->stream>>>(dst->g, gs.ps[0], gs.ps[1], gs.ps[2], src_arr,
                                                                                ia[0], ia[1], ia[2], dst_arr, outside);
        e = cudaGetLastError();
        dst->launches += 1;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(dst->stream);   // hb (host staging of the tables) goes out of scope
    CU(e);
    return PFEM_OK;
}

// Provider on a foreign rectilinear mesh: the unknown field interpolated linearly at the tensor-product points of the target axes.
extern "C" int pfem_interpolate_field(pfem_ctx* ctx, const size_t n[3], const double* ax0, const double* ax1, const double* ax2,
                                      const size_t stride[3], double* out) {
    NEED_MESH();
    if (!n || !ax0 || !ax1 || !ax2 || !stride || !out) FAIL(PFEM_ERR_BAD_INPUT, "null argument");
    // slab mode: COLLECTIVE (the halo planes are refreshed first); every rank interpolates on its local mesh (owned + halo planes),
    // so a target point is valid on the rank whose local extent along the slab axis contains it and continued as a constant elsewhere
    if (ctx->nranks > 1) TRY(halo_sync(ctx, SA_X));
    const double* ax[3] = {ax0, ax1, ax2};
    size_t total = 1, ni = 0, nd = 0;
    AxisTab tab[3];
    for (int a = 0; a < 3; ++a) {
        if (n[a] < 1) FAIL(PFEM_ERR_BAD_INPUT, "target axis %d is empty", a);
        total *= n[a];
        tab[a] = make_axis_tab(ctx->hax[a], false, std::vector<double>(ax[a], ax[a] + n[a]), false);
        ni += 2 * n[a]; nd += 3 * n[a];
    }
    if (total > 0x7fffffffull * 4) FAIL(PFEM_ERR_BAD_INPUT, "target mesh too large");
    {   // the output strides must be an iteration order of the target mesh (dense array of `total` values)
        int o[3] = {0, 1, 2};
        std::sort(o, o + 3, [&](int p, int q) { return stride[p] < stride[q]; });
        if (stride[o[0]] != 1 || stride[o[1]] != n[o[0]] || stride[o[2]] != n[o[0]] * n[o[1]])
            FAIL(PFEM_ERR_BAD_INPUT, "output strides are not an iteration order of the target mesh");
    }
    TRY(mask_field(ctx));
    const size_t tbytes = nd * sizeof(double) + ni * sizeof(int), off_out = ((tbytes + 15) / 16) * 16;
    std::vector<unsigned char> hb(tbytes);
    double* hd = reinterpret_cast<double*>(hb.data());
    int* hi = reinterpret_cast<int*>(hb.data() + nd * sizeof(double));
    TRY(ensure_itab(ctx, off_out + total * sizeof(double)));
    void* const dbuf = ctx->itab;
    double* dd = reinterpret_cast<double*>(dbuf);
    int* di = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(dbuf) + nd * sizeof(double));
    double* dout = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(dbuf) + off_out);
    InterpAxis ia[3];
    size_t od = 0, oi = 0;
    for (int a = 0; a < 3; ++a) {
        const size_t m = n[a];
        memcpy(hd + od, tab[a].lo.data(), m * 8); ia[a].lo = dd + od; od += m;
        memcpy(hd + od, tab[a].hi.data(), m * 8); ia[a].hi = dd + od; od += m;
        memcpy(hd + od, tab[a].pt.data(), m * 8); ia[a].pt = dd + od; od += m;
        memcpy(hi + oi, tab[a].ilo.data(), m * 4); ia[a].ilo = di + oi; oi += m;
        memcpy(hi + oi, tab[a].ihi.data(), m * 4); ia[a].ihi = di + oi; oi += m;
    }
    cudaError_t e = cudaMemcpyAsync(dbuf, hb.data(), tbytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const Grid& g = ctx->g;
        k_interp_to_points<<<vec_blocks(ctx), 256, 0, ctx->stream>>>((int)n[0], (int)n[1], (int)n[2], (idx_t)stride[0], (idx_t)stride[1],
                                                                      (idx_t)stride[2], g.ps[0], g.ps[1], g.ps[2], ctx->x, ia[0], ia[1], ia[2], dout);
        e = cudaGetLastError();
        ctx->launches += 1;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    CU(e);
    return PFEM_OK;
}

extern "C" int pfem_transfer_temperature(pfem_ctx* electrical, pfem_ctx* thermal) {
    pfem_ctx* ctx = electrical;
    NEED_MESH();
    if (!thermal || !thermal->have_mesh) FAIL(PFEM_ERR_STATE, "thermal context has no mesh");
    const Grid& g = ctx->g;
    if (!ctx->Te) TRY(dev_alloc(ctx, &ctx->Te, (size_t)g.NP, (size_t)g.G));
    InterpOutside out;
    memset(&out, 0, sizeof out);
    if (thermal->has_excluded) {   // masked thermal mesh: NaN outside the kept elements -> SafeData's 300 K (therm3d.cpp:391-392)
        out.src_mat = thermal->mat;
        out.fill = 300.;
    }
    TRY(interp_to_elems(electrical, thermal, thermal->x, false, ctx->Te, out));
    ctx->conds_valid = false;
    return PFEM_OK;
}

extern "C" int pfem_transfer_heat(pfem_ctx* thermal, pfem_ctx* electrical) {
    pfem_ctx* ctx = electrical;
    if (!thermal || !thermal->have_mesh) return PFEM_ERR_STATE;
    NEED_MESH();
    if (!ctx->conds_valid) FAIL(PFEM_ERR_STATE, "electrical conductivities have not been computed");
    {   // Joule heat of the electrical solution on its own element lattice (aux0)
        const Grid& g = ctx->g;
        TRY(ensure_elem_arrays(ctx, false));
        k_gradient_fields<true><<<node_grid(g), node_block(), 0, ctx->stream>>>(g, ctx->x, ctx->cl, ctx->cv,
                                                                                 ctx->noheat_set ? ctx->noheat : nullptr,
                                                                                 ctx->aux0, ctx->aux1, ctx->aux2);
        KCHECK(); LAUNCHED(1);
    }
    const double* heat = ctx->aux0;
    ctx = thermal;
    CU(cudaSetDevice(ctx->device));
    const Grid& g = ctx->g;
    TRY(ensure_elem_arrays(ctx, false));
    CU(cudaMemsetAsync(ctx->aux0 - g.G, 0, (size_t)(g.NP + 2 * g.G) * sizeof(double), ctx->stream));
    InterpOutside out;
    memset(&out, 0, sizeof out);
    out.bbox = 1;                  // no heat outside the electrical solver's domain (electr3d.cpp:545-548)
    for (int a = 0; a < 3; ++a) { out.bb[2 * a] = electrical->hax[a].front(); out.bb[2 * a + 1] = electrical->hax[a].back(); }
    TRY(interp_to_elems(thermal, electrical, heat, true, ctx->aux0, out));
    k_load_vector<<<node_grid(g), node_block(), 0, ctx->stream>>>(g, ctx->aux0, ctx->has_excluded ? ctx->mat : nullptr, ctx->f);
    KCHECK(); LAUNCHED(1);
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

// ---------------------------------------------------------------- test / bench hooks ----

extern "C" int pfem_apply(pfem_ctx* ctx, const double* p, double* q, int variant) {
    NEED_MESH();
    if (!p || !q) FAIL(PFEM_ERR_BAD_INPUT, "null vector");
    if (!ctx->conds_valid) FAIL(PFEM_ERR_STATE, "conductivities have not been computed");
    const Grid& g = ctx->g;
    TRY(launch_diag(ctx));
    // r <- p (as given), p <- M p, q <- M A M p, then q_D <- p_D
    CU(upload_nodes(ctx, ctx->r, p));
    CU(cudaMemsetAsync(ctx->q, 0, (size_t)g.NP * sizeof(double), ctx->stream));
    k_select_fixed<<<vec_blocks(ctx), 256, 0, ctx->stream>>>(g.NP, ctx->fixed, ctx->q, ctx->r, ctx->p);
    LAUNCHED(1);
    if (variant == 1) {
        TRY(launch_apply_simple<3>(ctx, ctx->p, ctx->q));
    } else if (variant == 2) {
        if (!ctx->plan.valid) FAIL(PFEM_ERR_STATE, "no tiled kernel plan for this mesh");
        CU(launch_apply_tiled_plain(ctx->plan, g, ctx->cl, ctx->cv, ctx->dinv, ctx->p, ctx->q, ctx->stream));
        LAUNCHED(1);
    } else if (variant == 3) {
        if (!ctx->fused.valid) FAIL(PFEM_ERR_STATE, "fused PCG kernel unavailable: %s", ctx->fused.why);
        PeerOut none;
        memset(&none, 0, sizeof none);
        CU(launch_fused_dispatch<0>(ctx->fused, g, 0, nullptr, ctx->q, nullptr, nullptr, nullptr, nullptr, none, ctx->stream));
        LAUNCHED(1);
    } else if (variant == 0) {
        if (!ctx->tma.valid) FAIL(PFEM_ERR_STATE, "TMA operator kernel unavailable: %s", ctx->tma.why);
        CU(launch_tma_dispatch<false>(ctx->tma, g, ctx->tma.m_p[0], ctx->dinv, nullptr, ctx->q, nullptr, nullptr, ctx->stream));
        LAUNCHED(1);
    } else FAIL(PFEM_ERR_BAD_INPUT, "unknown kernel variant %d", variant);
    if (ctx->surf.nnz) {
        k_surf_add<<<surf_blocks(ctx), 256, 0, ctx->stream>>>(ctx->surf, ctx->dinv, ctx->p, ctx->q);
        LAUNCHED(1);
    }
    k_select_fixed<<<vec_blocks(ctx), 256, 0, ctx->stream>>>(g.NP, ctx->fixed, ctx->r, ctx->q, ctx->q);
    KCHECK(); LAUNCHED(1);
    CU(download_nodes(ctx, q, ctx->q));
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

// z = M^-1 r for the preconditioner opts->precond built from the current conds / Dirichlet set (free rows; 0 on fixed rows)
extern "C" int pfem_apply_precond(pfem_ctx* ctx, const pfem_opts* o, const double* r, double* z) {
    NEED_MESH();
    TRY(check_opts(ctx, o));
    if (!r || !z) FAIL(PFEM_ERR_BAD_INPUT, "null vector");
    if (!ctx->conds_valid) FAIL(PFEM_ERR_STATE, "conductivities have not been computed");
    if (ctx->nranks > 1) FAIL(PFEM_ERR_BAD_INPUT, "not available in slab mode");
    const Grid& g = ctx->g;
    TRY(launch_diag(ctx));
    CU(upload_nodes(ctx, ctx->q, r));
    CU(cudaMemsetAsync(ctx->r, 0, (size_t)g.NP * sizeof(double), ctx->stream));
    k_select_fixed<<<vec_blocks(ctx), 256, 0, ctx->stream>>>(g.NP, ctx->fixed, ctx->r, ctx->q, ctx->p);   // p = free rows of r
    LAUNCHED(1);
    if (o->precond == 0) {
        k_pupdate_plain<<<vec_blocks(ctx), 256, 0, ctx->stream>>>(g.NP, ctx->p, ctx->dinv, ctx->q);
        KCHECK(); LAUNCHED(1);
    } else {
        TRY(ensure_line(ctx));
        k_line_factor<<<(unsigned)((g.N / (g.vdim == 0 ? g.nI : g.vdim == 1 ? g.nJ : g.nK) + 127) / 128), 128, 0, ctx->stream>>>(
            g, ctx->cl, ctx->cv, ctx->dinv, ctx->ll, ctx->ld, ctx->lmask, ctx->d_sc);
        KCHECK(); LAUNCHED(1);
        if (o->precond == 2) {
            TRY(ensure_ml(ctx));
            TRY(ml_setup(ctx));
            TRY(launch_ml_chain(ctx, ctx->p, nullptr, nullptr, 2));
            k_ml_prolong_add<<<dim3((unsigned)((g.nI + 127) / 128), g.nJ, g.nK), 128, 0, ctx->stream>>>(g, ctx->lz, ml_coarse_add(ctx), ctx->q);
            KCHECK(); LAUNCHED(2 + ctx->ml.nagg);
        } else {
            TRY(launch_line_solve(ctx, ctx->p, nullptr, nullptr, 1));
            CU(cudaMemcpyAsync(ctx->q, ctx->lz, (size_t)g.NP * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            LAUNCHED(1);
        }
    }
    CU(download_nodes(ctx, z, ctx->q));
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

extern "C" int pfem_get_rhs(pfem_ctx* ctx, double* b) {
    NEED_MESH();
    if (!b) FAIL(PFEM_ERR_BAD_INPUT, "null output");
    if (!ctx->conds_valid) FAIL(PFEM_ERR_STATE, "conductivities have not been computed");
    const Grid& g = ctx->g;
    TRY(launch_diag(ctx));
    // q <- v_D on Dirichlet nodes, 0 elsewhere; p <- M (f - A q); b = fixed ? v_D : p
    CU(cudaMemsetAsync(ctx->q, 0, (size_t)g.NP * sizeof(double), ctx->stream));
    TRY(scatter_bc(ctx, ctx->q, nullptr, true));   // the value the free rows see (first condition per node)
    const double* feff = nullptr;
    TRY(surf_eval_rad(ctx));
    TRY(surf_rhs(ctx, ctx->q, &feff));
    TRY(launch_apply_simple<1>(ctx, ctx->q, ctx->p, feff));
    if (ctx->bc_val_first) TRY(scatter_bc(ctx, ctx->q, nullptr));   // B[r] = the last value
    k_select_fixed<<<vec_blocks(ctx), 256, 0, ctx->stream>>>(g.NP, ctx->fixed, ctx->q, ctx->p, ctx->p);
    KCHECK(); LAUNCHED(1);
    CU(download_nodes(ctx, b, ctx->p));
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

extern "C" int pfem_get_diag(pfem_ctx* ctx, double* d) {
    NEED_MESH();
    if (!d) FAIL(PFEM_ERR_BAD_INPUT, "null output");
    if (!ctx->conds_valid) FAIL(PFEM_ERR_STATE, "conductivities have not been computed");
    const Grid& g = ctx->g;
    TRY(launch_diag(ctx));
    k_diag_from_dinv<<<vec_blocks(ctx), 256, 0, ctx->stream>>>(g.NP, ctx->dinv, ctx->fixed, ctx->q);
    KCHECK(); LAUNCHED(1);
    CU(download_nodes(ctx, d, ctx->q));
    CU(cudaStreamSynchronize(ctx->stream));
    return PFEM_OK;
}

extern "C" int pfem_get_info(pfem_ctx* ctx, int what, double* value) {
    NEED_MESH();
    if (!value) FAIL(PFEM_ERR_BAD_INPUT, "null output");
    switch (what) {
        case PFEM_INFO_COND_ISO: *value = (ctx->conds_valid && ctx->cond_iso && !getenv("PFEM_NO_ISO")) ? 1. : 0.; break;
        case PFEM_INFO_ML_LEVELS: *value = ctx->ml.nlev; break;
        case PFEM_INFO_FUSED_CTAS: *value = ctx->fused.valid ? (double)ctx->fused.tilesI * ctx->fused.tilesJ * ctx->fused.chunksK : 0.; break;
        case PFEM_INFO_DEVICE_BYTES: { double b = 0.; for (auto& a : ctx->allocs) b += (double)a.bytes; *value = b + (double)ctx->stage_bytes; break; }
        default: FAIL(PFEM_ERR_BAD_INPUT, "unknown info %d", what);
    }
    return PFEM_OK;
}

extern "C" int pfem_bench_pcg(pfem_ctx* ctx, const pfem_opts* o, int iters, int split_timing, double* ms, double* apply_ms,
                              double* update_ms, long long* launches) {
    NEED_MESH();
    TRY(check_opts(ctx, o));
    if (iters <= 0) FAIL(PFEM_ERR_BAD_INPUT, "iters must be positive");
    long long l0 = ctx->launches;
    TRY(pcg_prepare(ctx, o, 1));
    CU(cudaStreamSynchronize(ctx->stream));
    l0 = ctx->launches;
    double t_apply = 0., t_update = 0., t_total = 0.;
    if (split_timing) {
        std::vector<cudaEvent_t> ev((size_t)iters * 3);
        for (auto& e : ev) CU(cudaEventCreate(&e));
        for (int it = 0; it < iters; ++it) LAUNCHED(launch_iteration(ctx, o->variant, it & 1, &ev[(size_t)it * 3]));
        KCHECK();
        CU(cudaStreamSynchronize(ctx->stream));
        for (int it = 0; it < iters; ++it) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, ev[(size_t)it * 3], ev[(size_t)it * 3 + 1]);
            cudaEventElapsedTime(&b, ev[(size_t)it * 3 + 1], ev[(size_t)it * 3 + 2]);
            t_apply += a; t_update += b;
        }
        float tot = 0;
        cudaEventElapsedTime(&tot, ev[0], ev[(size_t)iters * 3 - 1]);
        t_total = tot;
        for (auto& e : ev) cudaEventDestroy(e);
    } else {
        int batch = o->batch > 0 ? o->batch : 32;
        batch += batch & 1;
        if (batch > iters) batch = iters & ~1;
        if (batch < 2) batch = 2;
        TRY(build_graph(ctx, batch, o->variant, o->precond));
        Timer t(ctx->stream);
        int done = 0;
        while (done + batch <= iters) {
            CU(cudaGraphLaunch(ctx->graph, ctx->stream));
            LAUNCHED((long long)batch * kernels_per_iteration(ctx, o->variant));
            done += batch;
        }
        for (; done < iters; ++done) LAUNCHED(launch_iteration(ctx, o->variant, done & 1, nullptr));
        KCHECK();
        t_total = t.stop();
    }
    TRY(read_scalars(ctx));
    k_set_params<<<1, 1, 0, ctx->stream>>>(ctx->d_sc, o->lin_tol * o->lin_tol, o->maxit, 0, ctx->surf_iter, 0);
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_sc->iter != iters) FAIL(PFEM_ERR_CUDA, "benchmark ran %d iterations instead of %d", ctx->h_sc->iter, iters);
    if (ms) *ms = t_total;
    if (apply_ms) *apply_ms = t_apply;
    if (update_ms) *update_ms = t_update;
    if (launches) *launches = ctx->launches - l0;
    return PFEM_OK;
}
