/* plaskfem_cuda.h — C ABI of libplaskfem_cuda.so: the B200 (sm_100a, FP64) replacement for
 * the hot path of PLaSK's thermal.static.Static3D and electrical.shockley.Shockley3D.
 *
 * What it replaces (paths relative to the PLaSK source tree):
 *   - the pair setMatrix(A,B,...) + A.solve(B,X) inside the nonlinear do..while of
 *     ThermalFem3DSolver::compute      solvers/thermal/static/therm3d.cpp:311-334
 *     ElectricalFem3DSolver::compute   solvers/electrical/shockley/electr3d.cpp:383-433
 *     i.e. element stiffness assembly (therm3d.cpp:186-276, electr3d.cpp:281-342), Dirichlet
 *     elimination (plask/common/fem/matrix.hpp:111-118, iterative_matrix.hpp:462-485) and
 *     the SPD solve (iterative_matrix.hpp:141-339 -> extlib/nspcg, or cholesky_matrix.hpp:90-111);
 *   - the conductivity updates around it (therm3d.cpp:204-220, electr3d.cpp:203-225,246-274,
 *     beta.hpp:43-46), the loop error reductions (therm3d.cpp:318-326, electr3d.cpp:387-425)
 *     and the element post-processing (therm3d.cpp:342-384, electr3d.cpp:444-478).
 * The solver plugin selects it with a new FemMatrixAlgorithm value (fem_solver.hpp:25-29);
 * see INTEGRATION.md for the plugin-side binding.
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary: every call returns a pfem_status
 *     (0 = ok, < 0 = error, > 0 = completed with a condition the caller maps to its
 *     `noconv` policy, iterative_matrix.hpp:299-314).
 *   - every pointer argument is HOST memory owned by the caller and is copied during the
 *     call (DataVector<T>::data(), plask/data.hpp); the library owns all device memory
 *     until pfem_destroy (solver onInvalidate, therm3d.cpp:118-122, electr3d.cpp:195-201).
 *   - node arrays have N = n0*n1*n2 entries indexed exactly like RectangularMesh<3>
 *     (plask/mesh/rectilinear3d.cpp:20-32); element arrays have E = (n0-1)(n1-1)(n2-1)
 *     entries indexed like the element mesh of the same iteration order
 *     (plask/mesh/rectangular3d.cpp:20-22).  Tensor2 / Vec<3> element arrays are interleaved
 *     (c00,c11) / (c0,c1,c2) like DataVector<Tensor2<double>> / DataVector<Vec<3>>.
 *   - all arithmetic is FP64.  Not thread-safe per context; no callbacks.
 *   - there is NO CPU fallback: without a CUDA device pfem_create fails with
 *     PFEM_ERR_NO_DEVICE.
 */
#ifndef PLASKFEM_CUDA_H
#define PLASKFEM_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFEM_ABI_VERSION 1

typedef struct pfem_ctx pfem_ctx;

typedef enum {
    PFEM_OK = 0,
    PFEM_NOT_CONVERGED = 1,   /* linear solve hit maxit (NSPCG ier = 1, iterative_matrix.hpp:299) */
    PFEM_ERR_CUDA = -1,       /* CUDA runtime error, text in pfem_last_error */
    PFEM_ERR_NO_DEVICE = -2,  /* no usable CUDA device: the GPU path has no CPU fallback */
    PFEM_ERR_BAD_INPUT = -3,  /* BadInput (plask/exceptions.hpp) */
    PFEM_ERR_STATE = -4,      /* call order: mesh / materials / field not set */
    PFEM_ERR_NOT_SPD = -5,    /* p.Ap <= 0: ComputationError "not positive definite" (ier -6/-7) */
    PFEM_ERR_NOMEM = -6,
    PFEM_ERR_NAN = -7         /* non-finite value met in the iteration */
} pfem_status;

/* ---- life cycle --------------------------------------------------------------------- */
int pfem_abi_version(void);
int pfem_device_count(void);                       /* 0 if no driver / no device */
int pfem_create(pfem_ctx** ctx, int device);
void pfem_destroy(pfem_ctx* ctx);
const char* pfem_strerror(int status);
const char* pfem_last_error(const pfem_ctx* ctx);  /* detail of the last failing call */

/* ---- internal layout (optional, before pfem_set_mesh) --------------------------------- */
/* The ABI always speaks the mesh's own iteration order (node index = rectilinear3d.cpp:20-32, element index
 * likewise).  PFEM_LAYOUT_ABI keeps that order in device memory.  PFEM_LAYOUT_VERTICAL_MINOR stores the fields with the
 * physical vertical axis fastest, whatever the mesh order: the lines of the line-Jacobi preconditioner are then
 * contiguous rows (warp-per-row solve instead of the strided one).  Transfers permute on the device; results do not
 * depend on the layout.  Slab mode needs the ABI's major axis to be lateral with this layout. */
#define PFEM_LAYOUT_ABI 0
#define PFEM_LAYOUT_VERTICAL_MINOR 1
int pfem_set_layout(pfem_ctx* ctx, int layout);

/* ---- problem description (host arrays, copied) --------------------------------------- */

/* Mesh: n[a], coordinates ax_a[0..n[a]-1] in um for the physical axes a = 0,1,2 (axis 2 is
 * vertical); stride[a] = linear-index stride of axis a, i.e. the iteration order of
 * RectilinearMesh3D (rectilinear3d.hpp:349; any of the 6 permutations).  Resets everything
 * else in the context. */
int pfem_set_mesh(pfem_ctx* ctx, const size_t n[3], const double* ax0, const double* ax1, const double* ax2,
                  const size_t stride[3]);

/* Element weights along one physical axis (after pfem_set_mesh): every element integral of the stiffness operator, the load
 * vector and the heat capacity is multiplied by w[i], i = index of the element along `axis` (n[axis] - 1 values, > 0; NULL
 * removes the weights); gradients (currents, fluxes, Joule heat) keep the geometric spacing.  This is what the cylindrical
 * 2-D solvers need — K_e and f_e carry the midpoint radius r of the element (therm2d.cpp:338-420, electr2d.cpp:219-230) — when
 * a 2-D mesh is handed over as a brick mesh of one element layer.  Call it before pfem_set_boundary, which
 * accepts weights only in its cylindrical 2-D mode (pfem_boundary::mode2d = 2). */
int pfem_set_axis_weight(pfem_ctx* ctx, int axis, const double* w);

/* Material id per element + per-id conductivity tables c_lat/c_vert[nmat][nT] sampled by the
 * host on the grid T0 + i*dT from material->thermk(T, thickness) (thermal, therm3d.cpp:213;
 * ids distinguish (material, layer thickness) pairs, therm3d.cpp:81-114) or material->cond(T)
 * (electrical, electr3d.cpp:221).  Interpolation is linear, clamped at both ends.
 *
 * Masked meshes (empty-elements="exclude", FemSolverWithMaskedMesh::setupMaskedMesh, fem_solver.hpp:182-189 — the
 * default of the reference's Cholesky path): pass PFEM_MAT_EXCLUDED for the elements that are not in the masked mesh
 * (material kind EMPTY).  They get no stiffness, no load, no current, heat or flux; mesh nodes that touch no kept element
 * drop out of the system (they are not nodes of RectangularMaskedMesh3D) and read back as 0 from pfem_get_field.  All
 * arrays of the ABI stay in FULL-mesh numbering; the plugin maps to its masked numbering (plaskfem::MaskedNumbering in
 * plaskfem_cuda.hpp).  Call pfem_set_materials before pfem_set_source when elements are excluded. */
#define PFEM_MAT_EXCLUDED 0xFFFFFFFFu
int pfem_set_materials(pfem_ctx* ctx, const uint32_t* elem_mat, uint32_t nmat, double T0, double dT, uint32_t nT,
                       const double* c_lat, const double* c_vert);

/* First-kind boundary conditions flattened in application order:
 * for (cond : bconds) for (r : cond.place) (r, cond.value)  — matrix.hpp:111-118. */
int pfem_set_dirichlet(pfem_ctx* ctx, size_t nd, const size_t* node, const double* value);

/* Volumetric heat source per element, W/m^3 (inHeat(elementMesh), therm3d.cpp:179,223);
 * NULL = zero load vector (electrical, electr3d.cpp:278). */
int pfem_set_source(pfem_ctx* ctx, const double* heat_per_elem);

/* Boundary conditions of the 2nd kind (heat flux), 3rd kind (convection) and radiation of ThermalFem3DSolver
 * (heatflux_boundary, convection_boundary, radiation_boundary: therm3d.hpp:79-82; setBoundaries therm3d.cpp:140-168,
 * applied in setMatrix :242-268).  Each kind is given as the per-node optional value that
 * BoundaryConditionsWithMesh::getValue(node) yields (first matching condition wins,
 * plask/mesh/boundary_conditions.hpp:182-186): has_*[N] flags (NULL = no condition of that kind) and value arrays
 * of length N in the node order of the mesh.  A side of an element carries a condition when all four of its nodes
 * have a value (:153).  Heat flux and radiation change the load vector only (radiation is re-evaluated from the
 * temperatures every nonlinear loop, :262-267), convection also adds a face mass matrix, applied matrix-free.
 * verbatim != 0 reproduces setBoundaries to the letter: it accumulates into the LOCAL slots F[i], K[i][j], i,j in 0..3
 * (:157-162), i.e. always into the element's four z-low nodes whatever the side, and radiation reads
 * temperatures[0..7] (:265), and it keeps the convection matrix as written (:255: 0.125e-12, a QUARTER of the
 * consistent face mass matrix, so that the solution relaxes towards 4*ambient).  verbatim == 0 is the corrected
 * form: terms go to the nodes of the wall, radiation reads the wall node, consistent face mass matrix.
 * b == NULL (or no flags) removes all boundary terms.
 *
 * mode2d != 0: the mesh is the one-layer embedding of a 2-D mesh (2 nodes along physical axis 0) and the
 * conditions are those of ThermalFem2DSolver (therm2d.cpp: setBoundaries :138-172, Cartesian terms :225-265, cylindrical
 * :371-413, mode2d = 2; the radius is physical axis 1).  An EDGE of a 2-D element carries a condition when both of its nodes
 * have a value; flags and values are read from the nodes of plane 0 and the terms are added on both planes, scaled like the
 * embedded operator.  The 2-D solver puts every term on the right node, so verbatim means something else here: the convection
 * matrix terms are taken as written, WITHOUT the 1e-6 (um -> m) their load terms carry (:241,244 against :238) and,
 * cylindrical, with the second factor r they pick up in A += r * k11 (:385-397 against :415-426) — the boundary is then pinned
 * to about 1e-6 * ambient (its matrix term is 1e6 times its own load term); verbatim == 0 restores the unit factor and the
 * single r.  mode2d = 2 is the one case that may be combined with pfem_set_axis_weight (set the weights first). */
typedef struct {
    const uint8_t* has_flux;  const double* flux;                                     /* W/m^2            */
    const uint8_t* has_conv;  const double* conv_coeff;     const double* conv_ambient; /* W/(m^2 K), K    */
    const uint8_t* has_rad;   const double* rad_emissivity; const double* rad_ambient;  /* -, K            */
    int verbatim;
    int mode2d;               /* 0 brick faces (therm3d.cpp), 1 Cartesian 2-D edges, 2 cylindrical 2-D edges (therm2d.cpp) */
    int reserved[2];
} pfem_boundary;
int pfem_set_boundary(pfem_ctx* ctx, const pfem_boundary* b);

/* Host-only view of what the 2-D mode adds (no device, no context): the edge terms of a 2-D mesh x[n1] (tran or r) by y[n2] (vert),
 * node (i1, i2) -> i1 * n2 + i2, in the 2-D solver's own units (before the embedding scale): load[N] (heat flux + convection),
 * radiation as load -= rad_coef[n] * (T[n]^4 - rad_amb4[n]), and the convection matrix K[N*N] (dense, row-major).  b->mode2d must
 * be 1 or 2.  Lets the CPU test suite compare the flattening with the oracle's restatement of therm2d.cpp. */
int pfem_edges2d_host(size_t n1, const double* x, size_t n2, const double* y, const pfem_boundary* b,
                      double* load, double* rad_coef, double* rad_amb4, double* K);

/* Unknown field (temperatures [K] / potential [V]): initial guess and warm start
 * (therm3d.cpp:79, electr3d.cpp:190; iterative_matrix.hpp:205-209). */
int pfem_set_field(pfem_ctx* ctx, const double* x0);
int pfem_fill_field(pfem_ctx* ctx, double value);

/* Temperature at element midpoints for sigma(T) (inTemperature(elementMesh), electr3d.cpp:204-205);
 * NULL = uniform `uniform_T`. */
int pfem_set_elem_temperature(pfem_ctx* ctx, const double* T_elem, double uniform_T);

/* One p-n junction, ElectricalFem3DSolver::Active (electr3d.hpp:28-78) after
 * setupActiveRegions (electr3d.cpp:89-183): node-plane indices bottom/top along axis 2,
 * element index ranges [left,right) along axis 1 and [back,front) along axis 0,
 * ld = front-back, offset into the junction table, height in um. */
typedef struct {
    size_t bottom, top, left, right, back, front, ld;
    ptrdiff_t offset;
    double height;
} pfem_junction;

/* Shockley description.  elem_junc[E]: 0 = not a junction, k+1 = junction k (isActive,
 * electr3d.hpp:148-171).  elem_role[E]: 0 none, 1 p-contact, 2 n-contact (electr3d.cpp:216-219),
 * may be NULL.  junc_cond[ncol][2]: junction_conductivity table (electr3d.hpp:87), ncol =
 * sum over junctions of (right-left)*(front-back).  beta_col/js_col[ncol]: Shockley
 * parameters per table entry; the host evaluates beta(T), js(T) (beta.hpp:50-77,
 * electr_python.cpp:67-110) at the mid-plane element temperature (electr3d.cpp:261-262).
 * stable != 0 selects CONVERGENCE_STABLE (electr3d.cpp:263-268). */
int pfem_set_junctions(pfem_ctx* ctx, uint32_t njunc, const pfem_junction* junc, const uint32_t* elem_junc,
                       const uint8_t* elem_role, double pcond, double ncond, size_t ncol, const double* junc_cond,
                       const double* beta_col, const double* js_col, int stable);

/* ---- slab mode: one context per GPU, z-slab partition along the MAJOR mesh axis ------------------------
 * (new with this library: the reference has no domain decomposition).  The plugin gives each
 * context the LOCAL mesh = its owned node planes along the slowest-varying index plus ONE halo plane towards each
 * existing neighbour (all arrays of the calls above in local indexing, Dirichlet nodes of halo planes included), then
 *     pfem_slab_configure(ctx, rank, nranks, own_lo, own_hi)       owned planes = local [own_lo, own_hi)
 *     pfem_slab_export(ctx, blob)                                  pfem_slab_blob_size() bytes
 *     (exchange the blobs between the processes: MPI / torch.distributed all-gather — plumbing)
 *     pfem_slab_connect(ctx, all_blobs)                            maps the peers' memory (CUDA IPC, NVLink P2P)
 * ONE PROCESS PER GPU.  Every solve call is then COLLECTIVE: all ranks must make the same calls in the same order.
 * Per PCG iteration the boundary planes of r, q, p are written straight into the neighbours' halo planes by the
 * iteration kernel and the 7 CG scalars are exchanged once through peer inboxes (no separate collective launch);
 * all ranks obtain bit-identical scalars.  pfem_get_field returns the local field with up-to-date halo planes.
 * Implemented for pfem_solve_thermal, pfem_solve_shockley, pfem_solve_dynamic and pfem_solve_linear (kernel variant 3) with all
 * three preconditioners.  The multilevel preconditioner (precond = 2) asks for slab boundaries at multiples of 16 planes (every
 * rank but the last owns a multiple of 16 planes): its lateral aggregates then coincide with those of the undivided mesh and the
 * iteration counts are those of one GPU; only its top level (one vertical line for the whole device) is summed over the ranks. */
size_t pfem_slab_blob_size(void);
int pfem_slab_configure(pfem_ctx* ctx, int rank, int nranks, size_t own_lo, size_t own_hi);
int pfem_slab_export(pfem_ctx* ctx, void* blob);
int pfem_slab_connect(pfem_ctx* ctx, const void* blobs);

/* ---- solve --------------------------------------------------------------------------- */

typedef struct {
    int maxit;        /* max PCG iterations per linear solve (iter_params.maxit)            */
    double lin_tol;   /* stop when ||r||2 <= lin_tol*||b||2 AND ||r||_D^-1 <= lin_tol*||b||_D^-1 (b = free rows of the rhs) */
    int precond;      /* 0 = Jacobi (NSPCG "jac"); 1 = line-Jacobi (NSPCG "ljac"): tridiagonal line blocks along the
                       * physical vertical axis, two kernels per iteration; in slab mode the vertical axis must not be the major one;
                       * 2 = additive multilevel line preconditioner (kernels_ml.cuh): line blocks + the line blocks of the Galerkin
                       * operators on 4^l x 4^l lateral aggregates up to one column — the counterpart of the strength of the
                       * reference's default "ic" (iterative_matrix.hpp:73); needs PFEM_LAYOUT_VERTICAL_MINOR (or a mesh order whose
                       * minor axis is vertical) and at most 512 nodes per vertical line; slab mode: boundaries at multiples of 16 planes */
    double outer_tol; /* maxerr of the nonlinear loop: K (thermal) or % (electrical)        */
    int loops;        /* max nonlinear loops in this call, 0 = until converged              */
    int batch;        /* PCG iterations per captured CUDA graph launch (0 = default)        */
    int variant;      /* 3 = production: fused single-kernel PCG iteration (default);
                       * 0 = two-kernel iteration, TMA operator kernel; 2 = same with LDG tiles;
                       * 1 = simple one-thread-per-node reference kernels (tests)            */
    int reserved[5];
} pfem_opts;

typedef struct {
    int outer_loops;         /* loops done in this call                                     */
    int loopno;              /* loops done since the last pfem_set_mesh (solver loopno)     */
    long long lin_iters;     /* PCG iterations summed over the call                         */
    int last_iters;          /* ... of the last linear solve (iter_params.iters)            */
    int converged;           /* last linear solve converged (iter_params.converged)         */
    double lin_relres;       /* ||r||2/||b_free||2 at exit of the last solve (iter_params.err) */
    double err;              /* last loop error (K or %)                                    */
    double toterr;           /* max loop error in this call (the value compute() returns)   */
    double maxval;           /* max T (thermal) / max |j| at the junction, kA/cm2           */
    double maxcur[3];        /* current density vector where |j| is largest (electr3d.hpp:184) */
    double t_solve_ms;       /* device time of the call, CUDA events                        */
    long long kernel_launches; /* kernels of this library launched during the call          */
    double lin_relres_precond; /* sqrt(r.D^-1 r / b.D^-1 b) at exit of the last solve          */
    double reserved[3];
} pfem_stats;

void pfem_default_opts(pfem_opts* opts);

/* The whole nonlinear loop of ThermalFem3DSolver::compute (therm3d.cpp:311-334) on the device. */
int pfem_solve_thermal(pfem_ctx* ctx, const pfem_opts* opts, pfem_stats* stats);
/* The whole nonlinear loop of ElectricalFem3DSolver::compute (electr3d.cpp:378-435),
 * including loadConductivity before and saveConductivity after it. */
int pfem_solve_shockley(pfem_ctx* ctx, const pfem_opts* opts, pfem_stats* stats);

/* ---- Dynamic3D: thermal.dynamic.Dynamic3D --------------------------------------------------------------
 * Volumetric heat capacity per material id on the temperature grid of pfem_set_materials:
 * cp_dens[nmat][nT] = material->cp(T) * material->dens(T)  [J/(m^3 K)]  (femT3d.cpp:176). */
int pfem_set_capacity(pfem_ctx* ctx, uint32_t nmat, uint32_t nT, const double* cp_dens);
typedef struct {
    double time;         /* ns to advance in this call (argument of compute, femT3d.cpp:258)                              */
    double timestep;     /* ns, <loop timestep=...> (:57)                                                                 */
    double methodparam;  /* theta of the time scheme: 0.5 Crank-Nicolson (default), 1 backward Euler (:63,203-204)        */
    int lumping;         /* != 0: lumped capacity matrix (default, :207-212); 0: consistent (:216-231, kernel variant 1)  */
    int rebuildfreq;     /* steps between re-evaluations of k(T), cp(T) dens(T); 0 = only at the start of the call (:274) */
    double* maxT_log;    /* host array (may be NULL): max T after step i, i < maxT_log_len (the LOG_RESULT of :287-291)    */
    size_t maxT_log_len;
    int reserved[4];
} pfem_dynamic;
/* The time loop of DynamicThermalFem3DSolver::compute (femT3d.cpp:258-305) on the device, one PCG solve per step (warm start
 * from T^n like A.solverhs(T, temperatures), :283), with the CORRECTED update (the reference overwrites the
 * solution with the right-hand side, :287, and leaves the Dirichlet rows in B, :203-204,236):
 *   (theta K + C/dt) T^{n+1} = (C/dt - (1 - theta) K) T^n + F  on the free rows,  T^{n+1} = value on the Dirichlet rows.
 * The field (pfem_set_field / pfem_fill_field = inittemp, :82) is advanced in place; stats: outer_loops = steps done,
 * lin_iters = PCG iterations of all steps, maxval = max T.  opts->precond 0 or 1, opts->variant 3 (lumped) or 1.
 * The caller keeps the elapsed time (elapstime, :293,297).  Slab mode: collective, lumped capacity only, every rank passes the
 * same parameters (maxT_log_len included: the per-step maximum is a cross-rank reduction). */
int pfem_solve_dynamic(pfem_ctx* ctx, const pfem_opts* opts, const pfem_dynamic* dyn, pfem_stats* stats);

/* ---- results ------------------------------------------------------------------------- */
int pfem_get_field(pfem_ctx* ctx, double* x);  /* temperatures / potential, N doubles */

/* Provider on a foreign mesh (getTemperatures / getVoltage with INTERPOLATION_LINEAR, therm3d.cpp:387-395,
 * electr3d.cpp:518-524): the field interpolated linearly, exactly like RectilinearMesh3D::interpolateLinear
 * (rectilinear3d.hpp:802-845; constant outside the mesh), at the tensor-product points of the target axes; out[] is dense
 * with the target mesh's own index strides (any of its 6 iteration orders).  Slab mode: the call is COLLECTIVE (halo planes are
 * refreshed first) and every rank interpolates on its LOCAL mesh: a target point is valid on the rank whose local extent along
 * the slab axis (owned planes + halo planes) contains it; elsewhere the rank returns the constant continuation of its own slab. */
int pfem_interpolate_field(pfem_ctx* ctx, const size_t n[3], const double* ax0, const double* ax1, const double* ax2,
                           const size_t stride[3], double* out);

typedef enum {
    PFEM_ELEM_COND = 0,     /* conds, E x (c00,c11)  [W/m/K or S/m]                          */
    PFEM_ELEM_CURRENT = 1,  /* current, E x 3 [kA/cm2]            (electr3d.cpp:399-411)     */
    PFEM_ELEM_HEAT = 2,     /* Joule heat, E [W/m3]               (electr3d.cpp:444-478)     */
    PFEM_ELEM_FLUX = 3      /* heat flux, E x 3 [W/m2]            (therm3d.cpp:342-384)      */
} pfem_elem_field;
/* noheat[E] (may be NULL = keep what pfem_set_noheat stored): elements with EMPTY material or the "noheat" role
 * get zero heat (electr3d.cpp:472); only read for PFEM_ELEM_HEAT. */
int pfem_get_elem(pfem_ctx* ctx, int what, const uint8_t* noheat, double* out);
int pfem_get_junction_cond(pfem_ctx* ctx, double* junc_cond /* [ncol][2] */);
/* Element temperatures T_elem (as set by pfem_set_elem_temperature or pfem_transfer_temperature) at `n` selected elements
 * (indices in the element order of the mesh).  The host needs them at the mid-plane element of every junction column to
 * evaluate beta(T), js(T) (electr3d.cpp:261-262, electr_python.cpp:103-110) when the temperature field came from another
 * context and never touched the host. */
int pfem_get_elem_temperature(pfem_ctx* ctx, size_t n, const size_t* elem, double* T_out);

/* ---- field exchange between two contexts on the same device ------------------------
 * The ThermoElectric meta loop (solvers/meta/shockley/thermoelectric.py:207-211) connects
 *     electrical.inTemperature <- thermal.outTemperature      and      thermal.inHeat <- electrical.outHeat.
 * In the reference both go through a per-point virtual LazyData::at on the host; here the data never leave HBM.
 * The two meshes may differ (thermoelectric.py:132-138): values are interpolated linearly exactly like
 * RectilinearMesh3D::interpolateLinear (rectilinear3d.hpp:802-845, constant outside the source mesh).
 * In slab mode both contexts must hold the same planes of the same partition (the exchange is then local). */
/* EMPTY-material / "noheat" elements get zero Joule heat (electr3d.cpp:472); kept until the next pfem_set_mesh. */
int pfem_set_noheat(pfem_ctx* ctx, const uint8_t* noheat /* E, may be NULL to clear */);
/* T_elem of `electrical` <- temperatures of `thermal` at the element midpoints of `electrical`
 * (getTemperatures, therm3d.cpp:385-393 -> loadConductivity, electr3d.cpp:203-205). */
int pfem_transfer_temperature(pfem_ctx* electrical, pfem_ctx* thermal);
/* load vector of `thermal` <- Joule heat density of `electrical` (saveHeatDensity, electr3d.cpp:444-478; getHeatDensity
 * :538-548) at the element midpoints of `thermal` (setMatrix, therm3d.cpp:179,223). */
int pfem_transfer_heat(pfem_ctx* thermal, pfem_ctx* electrical);

/* ---- building blocks exposed for parity tests and the benchmark ---------------------- */
/* conds from the current field (thermal, therm3d.cpp:204-213) */
int pfem_update_conductivity_thermal(pfem_ctx* ctx);
/* conds from T_elem / junction table / contacts (loadConductivity, electr3d.cpp:203-225) */
int pfem_update_conductivity_shockley(pfem_ctx* ctx);
/* conds given directly, E x (c00,c11) */
int pfem_set_conductivity(pfem_ctx* ctx, const double* cond);
/* q = A p with A the Dirichlet-eliminated stiffness matrix of the current conds
 * (== SparseBandMatrix::mult after applyBC, iterative_matrix.hpp:420-433,462-485) */
int pfem_apply(pfem_ctx* ctx, const double* p, double* q, int variant);
/* z = M^-1 r with the preconditioner opts->precond (0 Jacobi, 1 line-Jacobi, 2 multilevel) built from the current conds and
 * Dirichlet set; r is masked to the free rows first.  Lets the parity tests compare the preconditioner itself (not only the
 * converged solution, which no SPD preconditioner can change) with its construction from the assembled matrix. */
int pfem_apply_precond(pfem_ctx* ctx, const pfem_opts* opts, const double* r, double* z);
/* load vector after applyBC (B of therm3d.cpp:278 / electr3d.cpp:344) */
int pfem_get_rhs(pfem_ctx* ctx, double* b);
/* diagonal of the eliminated matrix (1 on Dirichlet rows) */
int pfem_get_diag(pfem_ctx* ctx, double* d);
/* one linear solve from the current field with the current conds */
int pfem_solve_linear(pfem_ctx* ctx, const pfem_opts* opts, pfem_stats* stats);
/* facts about the prepared state, for the benchmark's byte accounting and the tests */
typedef enum {
    PFEM_INFO_COND_ISO = 0,     /* 1: c_lat == c_vert in every element, the iteration kernels do not stream c_vert (80 instead of 88 B/DOF) */
    PFEM_INFO_ML_LEVELS = 1,    /* coarse levels of the multilevel preconditioner (0: not set up) */
    PFEM_INFO_FUSED_CTAS = 2,   /* grid size of the fused iteration kernel */
    PFEM_INFO_DEVICE_BYTES = 3  /* device memory held by the context */
} pfem_info;
int pfem_get_info(pfem_ctx* ctx, int what, double* value);
/* exactly `iters` PCG iterations (no convergence exit) from the current state, device
 * resident, timed with CUDA events on the launching stream; *ms = elapsed, *apply_ms /
 * *update_ms = summed duration of the operator kernel / the vector-update kernel when
 * split_timing != 0 (adds event records between kernels). */
int pfem_bench_pcg(pfem_ctx* ctx, const pfem_opts* opts, int iters, int split_timing, double* ms, double* apply_ms,
                   double* update_ms, long long* launches);

#ifdef __cplusplus
}
#endif
#endif /* PLASKFEM_CUDA_H */
