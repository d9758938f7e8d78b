/* fem3d_oracle.c — CPU restatement of PLaSK's 3D FEM thermal/electrical hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in plask_b200/ (the product) may include,
 * link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Every function restates one piece of the reference and cites it
 * (paths relative to /root/reference).  Plain C99, single thread, the same loop
 * order and the same arithmetic expressions as the reference so that rounding
 * differences against the real thing stay at the 1-ulp level.
 *
 * Pinning: see oracle/README.md.  Shockley3D is pinned by the analytic values of
 * solvers/electrical/shockley/tests/shockley3d.py:62-83 (tests/test_oracle_pin.py);
 * the linear solve is pinned by running the reference's own NSPCG (oracle/_ref,
 * compiled from extlib/nspcg/nspcg.c) and LAPACK dpbtrf/dpbtrs on the matrices
 * assembled here.  Static3D has no reference test at all ("parity unpinned" by the
 * reference; cross-checked by Cholesky-vs-NSPCG agreement and a manufactured
 * solution).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ mesh --- */

/* RectangularMesh<3> linearisation, plask/mesh/rectilinear3d.cpp:20-32:
 *   index = i_minor + n_minor*(i_medium + n_medium*i_major)
 * and the element mesh (midpoints) keeps the same iteration order,
 * plask/mesh/rectangular3d.cpp:20-22.  `ns`/`es` are the resulting strides of
 * the PHYSICAL axes 0,1,2 for nodes and for elements. */
typedef struct {
    size_t n[3];
    const double* ax[3];
    size_t ns[3];
    size_t es[3];
} orc_mesh;

void orc_mesh_init(orc_mesh* m, size_t n0, size_t n1, size_t n2, const double* a0, const double* a1,
                   const double* a2, const int order[3] /* major, medium, minor axis numbers */) {
    m->n[0] = n0; m->n[1] = n1; m->n[2] = n2;
    m->ax[0] = a0; m->ax[1] = a1; m->ax[2] = a2;
    int mj = order[0], md = order[1], mn = order[2];
    m->ns[mn] = 1;            m->ns[md] = m->n[mn];          m->ns[mj] = m->n[mn] * m->n[md];
    m->es[mn] = 1;            m->es[md] = m->n[mn] - 1;      m->es[mj] = (m->n[mn] - 1) * (m->n[md] - 1);
}

size_t orc_mesh_nodes(const orc_mesh* m) { return m->n[0] * m->n[1] * m->n[2]; }
size_t orc_mesh_elements(const orc_mesh* m) { return (m->n[0] - 1) * (m->n[1] - 1) * (m->n[2] - 1); }

/* Element iteration in element-index order (maskedMesh->elements() with selectAll,
 * plask/mesh/rectangular_masked_common.hpp:303-320): decode e -> (i0,i1,i2). */
static inline void elem_indices(const orc_mesh* m, size_t e, size_t ix[3]) {
    /* sort axes by stride, largest first */
    int o[3] = {0, 1, 2};
    for (int a = 0; a < 3; ++a)
        for (int b = a + 1; b < 3; ++b)
            if (m->es[o[b]] > m->es[o[a]]) { int t = o[a]; o[a] = o[b]; o[b] = t; }
    /* ties only when an axis has one element; then its index is 0 regardless */
    size_t rem = e;
    for (int a = 0; a < 3; ++a) {
        size_t cnt = m->n[o[a]] - 1;
        size_t q = (m->es[o[a]] ? rem / m->es[o[a]] : 0);
        if (cnt <= 1) q = 0; else if (q >= cnt) q = cnt - 1;
        ix[o[a]] = q;
        rem -= q * m->es[o[a]];
    }
}

/* idx[0..7] = LLL,ULL,LUL,UUL,LLU,ULU,LUU,UUU (bit0 = upper in axis 0, bit1 = axis 1,
 * bit2 = axis 2), solvers/thermal/static/therm3d.cpp:189-197 */
static inline void elem_nodes(const orc_mesh* m, const size_t ix[3], size_t idx[8]) {
    size_t base = ix[0] * m->ns[0] + ix[1] * m->ns[1] + ix[2] * m->ns[2];
    for (int l = 0; l < 8; ++l)
        idx[l] = base + ((l & 1) ? m->ns[0] : 0) + ((l & 2) ? m->ns[1] : 0) + ((l & 4) ? m->ns[2] : 0);
}

/* ------------------------------------------------- a2: layer thickness ----- */

/* ThermalFem3DSolver::onInitialize, therm3d.cpp:81-114: thickness[e] = height of the
 * maximal vertical (axis 2) run of identical material containing e.  Material identity
 * (`m == material`) becomes equality of the material id. */
void orc_thickness(const orc_mesh* m, const uint32_t* elem_mat, double* thickness) {
    size_t E = orc_mesh_elements(m);
    for (size_t e = 0; e < E; ++e) thickness[e] = NAN;
    for (size_t e = 0; e < E; ++e) {
        if (!isnan(thickness[e])) continue;
        size_t ix[3];
        elem_indices(m, e, ix);
        uint32_t material = elem_mat[e];
        size_t row = ix[2];
        double top = m->ax[2][row + 1], bottom = m->ax[2][row];
        size_t itop = row + 1, ibottom = row;
        size_t col = ix[0] * m->es[0] + ix[1] * m->es[1];
        for (size_t r = row; r > 0; r--) {
            if (elem_mat[col + (r - 1) * m->es[2]] == material) { bottom = m->ax[2][r - 1]; ibottom = r - 1; }
            else break;
        }
        for (size_t r = row + 1; r < m->n[2] - 1; r++) {
            if (elem_mat[col + r * m->es[2]] == material) { top = m->ax[2][r + 1]; itop = r + 1; }
            else break;
        }
        double h = top - bottom;
        for (size_t r = ibottom; r < itop; ++r) thickness[col + r * m->es[2]] = h;
    }
}

/* ------------------------------------------ material tables (host-sampled) -- */

/* The drop-in host samples material->thermk(T,h) / material->cond(T)
 * (plask/material/material.hpp:691,649) on a uniform T grid per material id; both
 * the oracle and the CUDA library interpolate these tables linearly, clamped at
 * the ends.  tab is [nmat][nT]. */
static inline double table_at(const double* tab, uint32_t mat, uint32_t nT, double T0, double dT, double T) {
    double t = (T - T0) / dT;
    if (!(t > 0.)) t = 0.;
    double tmax = (double)(nT - 1);
    if (t > tmax) t = tmax;
    uint32_t i = (uint32_t)t;
    if (i > nT - 2) i = nT - 2;
    double f = t - (double)i;
    const double* row = tab + (size_t)mat * nT;
    return row[i] + f * (row[i + 1] - row[i]);
}

double orc_table_at(const double* tab, uint32_t mat, uint32_t nT, double T0, double dT, double T) {
    return table_at(tab, mat, nT, T0, dT, T);
}

/* a3 (thermal): therm3d.cpp:204-213 — T_mean = 0.125*sum of the 8 node temperatures,
 * (k_lat,k_vert) = thermk(T_mean, thickness).  Thickness dependence is folded into the
 * material id by the host.  cond is [E][2] = (c00 lateral, c11 vertical). */
void orc_thermal_conds(const orc_mesh* m, const double* T, const uint32_t* elem_mat, uint32_t nT, double T0,
                       double dT, const double* k_lat, const double* k_vert, double* cond) {
    size_t E = orc_mesh_elements(m);
    for (size_t e = 0; e < E; ++e) {
        size_t ix[3], idx[8];
        elem_indices(m, e, ix);
        elem_nodes(m, ix, idx);
        double temp = 0.;
        for (int i = 0; i < 8; ++i) temp += T[idx[i]];
        temp *= 0.125;
        cond[2 * e] = table_at(k_lat, elem_mat[e], nT, T0, dT, temp);
        cond[2 * e + 1] = table_at(k_vert, elem_mat[e], nT, T0, dT, temp);
    }
}

/* --------------------------------------------- a4: element stiffness ------- */

/* therm3d.cpp:215-237 == electr3d.cpp:310-338.  K[i][j] depends only on which axis
 * bits differ between local nodes i and j, so the 8 distinct values are indexed by i^j. */
static inline void elem_stiffness(double c00, double c11, double dx, double dy, double dz, double kv[8]) {
    double kx, ky = c00, kz = c11;
    ky *= 1e-6; kz *= 1e-6;
    kx = ky;
    kx /= dx; kx *= dy; kx *= dz;
    ky *= dx; ky /= dy; ky *= dz;
    kz *= dx; kz *= dy; kz /= dz;
    kv[0] = (kx + ky + kz) / 9.;
    kv[1] = (-2. * kx + ky + kz) / 18.;       /* differ in axis 0:  K[1][0] */
    kv[2] = (kx - 2. * ky + kz) / 18.;        /* differ in axis 1:  K[2][0] */
    kv[4] = (kx + ky - 2. * kz) / 18.;        /* differ in axis 2:  K[4][0] */
    kv[6] = (kx - 2. * ky - 2. * kz) / 36.;   /* axes 1,2:          K[4][2] */
    kv[5] = (-2. * kx + ky - 2. * kz) / 36.;  /* axes 0,2:          K[4][1] */
    kv[3] = (-2. * kx - 2. * ky + kz) / 36.;  /* axes 0,1:          K[2][1] */
    kv[7] = -(kx + ky + kz) / 36.;            /* all:               K[7][0] */
}

/* ------------------------- a5: SparseBandMatrix (14 symmetric diagonals) ---- */

/* plask/common/fem/iterative_matrix.hpp:372-390: offsets of the stored diagonals,
 * `major` = n_minor*n_medium, `minor` = n_minor (fem_solver.hpp:236-237). */
void orc_sparse14_offsets(size_t major, size_t minor, int icords[14]) {
    icords[0] = 0;
    icords[1] = 1;
    icords[2] = (int)minor - 1;
    icords[3] = (int)minor;
    icords[4] = (int)minor + 1;
    icords[5] = (int)(major - minor) - 1;
    icords[6] = (int)(major - minor);
    icords[7] = (int)(major - minor) + 1;
    icords[8] = (int)major - 1;
    icords[9] = (int)major;
    icords[10] = (int)major + 1;
    icords[11] = (int)(major + minor) - 1;
    icords[12] = (int)(major + minor);
    icords[13] = (int)(major + minor) + 1;
}

/* SparseBandMatrix::operator(), iterative_matrix.hpp:412-418 (linear search over the
 * offsets, data[min(r,c) + rank*d]). */
static inline double* sparse14_at(double* data, size_t rank, const int icords[14], size_t r, size_t c) {
    if (r == c) return data + r;
    if (r < c) { size_t t = r; r = c; c = t; }
    int d = (int)(r - c);
    size_t i = 0;
    while (i < 14 && icords[i] != d) ++i;
    if (i == 14) return NULL; /* assert(i != mdim) in the reference */
    return data + c + rank * i;
}

/* Assembly loop: therm3d.cpp:186-276 (heat != NULL, load f = 0.125e-18*dx*dy*dz*heat) and
 * electr3d.cpp:281-342 (heat == NULL, B = 0).  A is 14*N doubles, cleared here
 * (A.clear(); B.fill(0.), therm3d.cpp:182-183).  Returns 0, or -1 if an entry fell
 * outside the stored diagonals. */
int orc_assemble_sparse14(const orc_mesh* m, const double* cond, const double* heat, double* A, double* B) {
    size_t N = orc_mesh_nodes(m), E = orc_mesh_elements(m);
    /* minor = smallest node stride > 1 ... derive from strides */
    size_t s[3] = {m->ns[0], m->ns[1], m->ns[2]};
    for (int a = 0; a < 3; ++a) for (int b = a + 1; b < 3; ++b) if (s[b] < s[a]) { size_t t = s[a]; s[a] = s[b]; s[b] = t; }
    int icords[14];
    orc_sparse14_offsets(s[2], s[1], icords);
    memset(A, 0, 14 * N * sizeof(double));
    memset(B, 0, N * sizeof(double));
    for (size_t e = 0; e < E; ++e) {
        size_t ix[3], idx[8];
        elem_indices(m, e, ix);
        elem_nodes(m, ix, idx);
        double dx = m->ax[0][ix[0] + 1] - m->ax[0][ix[0]];
        double dy = m->ax[1][ix[1] + 1] - m->ax[1][ix[1]];
        double dz = m->ax[2][ix[2] + 1] - m->ax[2][ix[2]];
        double kv[8];
        elem_stiffness(cond[2 * e], cond[2 * e + 1], dx, dy, dz, kv);
        double f = heat ? 0.125e-18 * dx * dy * dz * heat[e] : 0.;
        for (int i = 0; i < 8; ++i) {
            for (int j = 0; j <= i; ++j) {
                double* a = sparse14_at(A, N, icords, idx[i], idx[j]);
                if (!a) return -1;
                *a += kv[i ^ j];
            }
            B[idx[i]] += f;
        }
    }
    return 0;
}

/* a8: FemMatrix::applyBC -> SparseBandMatrix::setBC, matrix.hpp:111-118 and
 * iterative_matrix.hpp:462-485; nodes are processed one by one in the given order. */
void orc_apply_bc_sparse14(size_t rank, const int icords[14], double* data, double* B, size_t nd,
                           const size_t* node, const double* value) {
    for (size_t k = 0; k < nd; ++k) {
        size_t r = node[k];
        double val = value[k];
        data[r] = 1.;
        B[r] = val;
        for (ptrdiff_t i = 13; i > 0; --i) { /* above diagonal */
            ptrdiff_t c = (ptrdiff_t)r - icords[i];
            if (c >= 0) {
                ptrdiff_t ii = c + (ptrdiff_t)rank * i;
                B[c] -= data[ii] * val;
                data[ii] = 0.;
            }
        }
        for (ptrdiff_t i = 1; i < 14; ++i) { /* below diagonal */
            size_t c = r + (size_t)icords[i];
            if (c < rank) {
                size_t ii = r + rank * (size_t)i;
                B[c] -= data[ii] * val;
                data[ii] = 0.;
            }
        }
    }
}

/* SparseBandMatrix::addmult, iterative_matrix.hpp:420-433 (result += A*vector). */
void orc_addmult_sparse14(size_t rank, const int icords[14], const double* data, const double* vector,
                          double* result) {
    for (size_t r = 0; r < rank; ++r) result[r] += data[r] * vector[r];
    for (size_t d = 1; d < 14; ++d) {
        size_t sd = rank * d;
        for (size_t r = 0; r < rank; ++r) {
            size_t c = r + (size_t)icords[d];
            if (c >= rank) break;
            result[r] += data[r + sd] * vector[c];
            result[c] += data[r + sd] * vector[r];
        }
    }
}

/* ------------------------------- a10: DpbMatrix (LAPACK lower band storage) - */

/* Bandwidth and leading dimension: fem_solver.hpp:219-229 (kd = n_minor*(n_medium+1)+1)
 * and cholesky_matrix.hpp:67-68 (ld = kd+1 rounded up to 2 doubles, minus 1). */
void orc_dpb_dims(const orc_mesh* m, size_t* kd, size_t* ld) {
    size_t s[3] = {m->ns[0], m->ns[1], m->ns[2]};
    for (int a = 0; a < 3; ++a) for (int b = a + 1; b < 3; ++b) if (s[b] < s[a]) { size_t t = s[a]; s[a] = s[b]; s[b] = t; }
    size_t band = s[2] + s[1] + 1; /* n_minor*n_medium + n_minor + 1 */
    *kd = band;
    *ld = ((band + 1 + (15 / sizeof(double))) & ~(size_t)(15 / sizeof(double))) - 1;
}

/* DpbMatrix::index, cholesky_matrix.hpp:70-88 with UPLO='L': data[ld*c + r] for r >= c. */
static inline double* dpb_at(double* data, size_t ld, size_t r, size_t c) {
    return (r < c) ? data + ld * r + c : data + ld * c + r;
}

void orc_assemble_dpb(const orc_mesh* m, const double* cond, const double* heat, size_t ld, double* AB, double* B) {
    size_t N = orc_mesh_nodes(m), E = orc_mesh_elements(m);
    memset(AB, 0, N * (ld + 1) * sizeof(double));
    memset(B, 0, N * sizeof(double));
    for (size_t e = 0; e < E; ++e) {
        size_t ix[3], idx[8];
        elem_indices(m, e, ix);
        elem_nodes(m, ix, idx);
        double dx = m->ax[0][ix[0] + 1] - m->ax[0][ix[0]];
        double dy = m->ax[1][ix[1] + 1] - m->ax[1][ix[1]];
        double dz = m->ax[2][ix[2] + 1] - m->ax[2][ix[2]];
        double kv[8];
        elem_stiffness(cond[2 * e], cond[2 * e + 1], dx, dy, dz, kv);
        double f = heat ? 0.125e-18 * dx * dy * dz * heat[e] : 0.;
        for (int i = 0; i < 8; ++i) {
            for (int j = 0; j <= i; ++j) *dpb_at(AB, ld, idx[i], idx[j]) += kv[i ^ j];
            B[idx[i]] += f;
        }
    }
}

/* ---------------- masked mesh: empty-elements = "exclude" (fem_solver.hpp:182-189) -------
 * RectangularMaskedMesh3D keeps the elements whose material is not EMPTY and the nodes of those elements,
 * numbered in the order of the full mesh (compressed sets, plask/mesh/rectangular_masked3d.hpp).  This is the
 * DEFAULT of the reference's Cholesky path (EMPTY_ELEMENTS_DEFAULT + ALGORITHM_CHOLESKY, fem_solver.hpp:183-187).
 * included[E] != 0 marks the kept elements; nodemap[N] receives the masked node number or SIZE_MAX.
 * Returns the number of masked nodes. */
size_t orc_masked_nodes(const orc_mesh* m, const uint8_t* included, size_t* nodemap) {
    size_t N = orc_mesh_nodes(m), E = orc_mesh_elements(m), cnt = 0;
    for (size_t i = 0; i < N; ++i) nodemap[i] = 0;
    for (size_t e = 0; e < E; ++e) {
        if (!included[e]) continue;
        size_t ix[3], idx[8];
        elem_indices(m, e, ix);
        elem_nodes(m, ix, idx);
        for (int l = 0; l < 8; ++l) nodemap[idx[l]] = 1;
    }
    for (size_t i = 0; i < N; ++i) nodemap[i] = nodemap[i] ? cnt++ : (size_t)-1;
    return cnt;
}

/* FemSolverWithMaskedMesh<Geometry3D,...>::getMatrix, fem_solver.hpp:219-231: band = max over the masked elements
 * of (UpUpUp - LoLoLo) in masked numbering; DpbMatrix ctor (cholesky_matrix.hpp:64-65) for ld. */
void orc_dpb_dims_masked(const orc_mesh* m, const uint8_t* included, const size_t* nodemap, size_t* kd, size_t* ld) {
    size_t E = orc_mesh_elements(m), band = 0;
    for (size_t e = 0; e < E; ++e) {
        if (!included[e]) continue;
        size_t ix[3], idx[8];
        elem_indices(m, e, ix);
        elem_nodes(m, ix, idx);
        size_t span = nodemap[idx[7]] - nodemap[idx[0]];
        if (span > band) band = span;
    }
    *kd = band;
    *ld = ((band + 1 + (15 / sizeof(double))) & ~(size_t)(15 / sizeof(double))) - 1;
}

/* The assembly loops (therm3d.cpp:186-276, electr3d.cpp:281-342) over maskedMesh->elements(), indices in masked
 * numbering.  AB is Nm*(ld+1), B is Nm. */
void orc_assemble_dpb_masked(const orc_mesh* m, const double* cond, const double* heat, const uint8_t* included,
                             const size_t* nodemap, size_t Nm, size_t ld, double* AB, double* B) {
    size_t E = orc_mesh_elements(m);
    memset(AB, 0, Nm * (ld + 1) * sizeof(double));
    memset(B, 0, Nm * sizeof(double));
    for (size_t e = 0; e < E; ++e) {
        if (!included[e]) continue;
        size_t ix[3], idx[8];
        elem_indices(m, e, ix);
        elem_nodes(m, ix, idx);
        double dx = m->ax[0][ix[0] + 1] - m->ax[0][ix[0]];
        double dy = m->ax[1][ix[1] + 1] - m->ax[1][ix[1]];
        double dz = m->ax[2][ix[2] + 1] - m->ax[2][ix[2]];
        double kv[8];
        elem_stiffness(cond[2 * e], cond[2 * e + 1], dx, dy, dz, kv);
        double f = heat ? 0.125e-18 * dx * dy * dz * heat[e] : 0.;
        for (int i = 0; i < 8; ++i) {
            for (int j = 0; j <= i; ++j) *dpb_at(AB, ld, nodemap[idx[i]], nodemap[idx[j]]) += kv[i ^ j];
            B[nodemap[idx[i]]] += f;
        }
    }
}

/* BandMatrix::setBC, matrix.hpp:132-145. */
void orc_apply_bc_dpb(size_t rank, size_t kd, size_t ld, double* AB, double* B, size_t nd, const size_t* node,
                      const double* value) {
    for (size_t k = 0; k < nd; ++k) {
        size_t r = node[k];
        double val = value[k];
        B[r] = val;
        *dpb_at(AB, ld, r, r) = 1.;
        size_t start = (r > kd) ? r - kd : 0;
        size_t end = (r + kd < rank) ? r + kd + 1 : rank;
        for (size_t c = start; c < r; ++c) { double* a = dpb_at(AB, ld, r, c); B[c] -= *a * val; *a = 0.; }
        for (size_t c = r + 1; c < end; ++c) { double* a = dpb_at(AB, ld, r, c); B[c] -= *a * val; *a = 0.; }
    }
}

/* ---------------- a7: boundary conditions of the 2nd / 3rd kind, radiation ---- */

/* setBoundaries (therm3d.cpp:140-168) and its three uses in setMatrix (:242-268), element by
 * element.  The per-node optional values (`boundary_conditions.getValue(idx[i])`,
 * plask/mesh/boundary_conditions.hpp:182-186: FIRST matching condition wins) are given as
 * flag + value arrays of length N.  A side of an element carries the condition when all 4 of its
 * nodes have a value (:153).
 *
 * quirk != 0 — VERBATIM: the reference accumulates into the LOCAL slots `F[i]`, `K[i][j]` with
 *   i,j in 0..3 (:157-162), i.e. always into the element's z-low nodes idx[0..3] whatever the
 *   side, decides "edge" from the slot numbers (:159-160), and the radiation lambda reads
 *   `temperatures[i]` with the local node number i = wall[i] in 0..7 (:265).  For the side
 *   {0,1,2,3} (z-low face of the element) slots and wall nodes coincide.
 * quirk == 0 — CORRECTED: F[wall[i]], K[wall[i]][wall[j]], edge from wall[i]^wall[j],
 *   temperatures[idx[wall[i]]], and the convection matrix scaled to the consistent face mass matrix (see below).  The CUDA library takes a flattened face-term list built by the
 *   host in either mode.
 *
 * Output: B additions are applied to B directly; matrix additions A(rows[k], cols[k]) += vals[k]
 * (symmetric storage, one triplet per local (i,j), j <= i) are appended up to `cap`.  Returns the
 * number of triplets produced (call again with a larger cap if it exceeds it). */
size_t orc_boundary_terms(const orc_mesh* m, const double* T, const uint8_t* has_flux, const double* flux,
                          const uint8_t* has_conv, const double* conv_coeff, const double* conv_amb,
                          const uint8_t* has_rad, const double* rad_emis, const double* rad_amb, int quirk,
                          size_t cap, size_t* rows, size_t* cols, double* vals, double* B, const uint8_t* included) {
    static const int walls[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}};
    const double SB = 5.670373e-8; /* plask/phys/constants.hpp:41 */
    size_t E = orc_mesh_elements(m), nt = 0;
    for (size_t e = 0; e < E; ++e) {
        size_t ix[3], idx[8];
        /* the loop is `for (auto elem : this->maskedMesh->elements())`, therm3d.cpp:186: elements outside the masked mesh
         * (empty-elements="exclude") add nothing */
        if (included && !included[e]) continue;
        elem_indices(m, e, ix);
        elem_nodes(m, ix, idx);
        double dx = m->ax[0][ix[0] + 1] - m->ax[0][ix[0]];
        double dy = m->ax[1][ix[1] + 1] - m->ax[1][ix[1]];
        double dz = m->ax[2][ix[2] + 1] - m->ax[2][ix[2]];
        const double areas[3] = {dx * dy, dy * dz, dz * dx};
        double F[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        double K[8][8];
        memset(K, 0, sizeof K);
        for (int kind = 0; kind < 3; ++kind) { /* heat flux, convection, radiation — the order of :242-268 */
            const uint8_t* has = kind == 0 ? has_flux : kind == 1 ? has_conv : has_rad;
            if (!has) continue;
            for (int side = 0; side < 6; ++side) {
                const int* wall = walls[side];
                if (!(has[idx[wall[0]]] && has[idx[wall[1]]] && has[idx[wall[2]]] && has[idx[wall[3]]])) continue;
                double area = areas[side / 2];
                for (int i = 0; i < 4; ++i) {
                    size_t ni = idx[wall[i]];
                    int si = quirk ? i : wall[i]; /* slot that receives the term */
                    double fv;
                    if (kind == 0) fv = -0.25e-12 * area * flux[ni];
                    else if (kind == 1) fv = 0.25e-12 * area * conv_coeff[ni] * conv_amb[ni];
                    else {
                        double a = rad_amb[ni]; a = a * a;
                        double t = quirk ? T[wall[i]] : T[ni]; t = t * t;
                        fv = -0.25e-12 * area * rad_emis[ni] * SB * (t * t - a * a);
                    }
                    F[si] += fv;
                    if (kind != 1) continue;
                    for (int j = 0; j <= i; ++j) {
                        int sj = quirk ? j : wall[j];
                        int ij = quirk ? (i ^ j) : (wall[i] ^ wall[j]);
                        int edge = (ij == 1 || ij == 2 || ij == 4);
                        /* :255 has 0.125e-12: with the /9,/18,/36 weights that is a QUARTER of the consistent face
                         * mass matrix (rows sum to A c/16 while the load is A c Ta/4), so the verbatim form relaxes
                         * towards 4 Ta; the corrected form uses the consistent matrix (rows sum to A c/4). */
                        double v = (quirk ? 0.125e-12 : 0.5e-12) * area * (conv_coeff[ni] + conv_coeff[idx[wall[j]]]);
                        v = v / (wall[j] == wall[i] ? 9. : edge ? 18. : 36.);
                        if (si >= sj) K[si][sj] += v; else K[sj][si] += v;
                    }
                }
            }
        }
        for (int i = 0; i < 8; ++i) {
            for (int j = 0; j <= i; ++j)
                if (K[i][j] != 0.) {
                    if (nt < cap) { rows[nt] = idx[i]; cols[nt] = idx[j]; vals[nt] = K[i][j]; }
                    ++nt;
                }
            B[idx[i]] += F[i];
        }
    }
    return nt;
}

void orc_add_coo_sparse14(size_t rank, const int icords[14], double* data, size_t nt, const size_t* rows,
                          const size_t* cols, const double* vals) {
    for (size_t k = 0; k < nt; ++k) *sparse14_at(data, rank, icords, rows[k], cols[k]) += vals[k];
}
void orc_add_coo_dpb(size_t ld, double* AB, size_t nt, const size_t* rows, const size_t* cols, const double* vals) {
    for (size_t k = 0; k < nt; ++k) *dpb_at(AB, ld, rows[k], cols[k]) += vals[k];
}

/* ------------------------------------ a11: thermal outer-loop reductions --- */

/* therm3d.cpp:318-325: err = max|T - T_prev|, maxT = max T (maxT starts at 0). */
void orc_thermal_error(size_t N, const double* T, const double* T0, double* err, double* maxT) {
    double e = 0., mt = 0.;
    for (size_t i = 0; i < N; ++i) {
        double corr = fabs(T0[i] - T[i]);
        if (corr > e) e = corr;
        if (T[i] > mt) mt = T[i];
    }
    *err = e;
    *maxT = mt;
}

/* a14 (thermal): saveHeatFluxes, therm3d.cpp:342-384.  flux is [E][3] in W/m^2.  The
 * conductivity pair is supplied by the caller (the reference re-evaluates thermk with
 * the leaf bounding-box height, :366-370; the host folds that into the table id). */
void orc_heat_flux(const orc_mesh* m, const double* T, const double* cond, double* flux) {
    size_t E = orc_mesh_elements(m);
    for (size_t e = 0; e < E; ++e) {
        size_t ix[3], n[8];
        elem_indices(m, e, ix);
        elem_nodes(m, ix, n);
        double lll = T[n[0]], ull = T[n[1]], lul = T[n[2]], uul = T[n[3]], llu = T[n[4]], ulu = T[n[5]], luu = T[n[6]], uuu = T[n[7]];
        double kxy = cond[2 * e], kz = cond[2 * e + 1];
        flux[3 * e] = -0.25e6 * kxy * (-lll - llu - lul - luu + ull + ulu + uul + uuu) / (m->ax[0][ix[0] + 1] - m->ax[0][ix[0]]);
        flux[3 * e + 1] = -0.25e6 * kxy * (-lll - llu + lul + luu - ull - ulu + uul + uuu) / (m->ax[1][ix[1] + 1] - m->ax[1][ix[1]]);
        flux[3 * e + 2] = -0.25e6 * kz * (-lll + llu - lul + luu - ull + ulu - uul + uuu) / (m->ax[2][ix[2] + 1] - m->ax[2][ix[2]]);
    }
}

/* ------------------------------------------------------------- Shockley ---- */

/* ElectricalFem3DSolver::Active, electr3d.hpp:28-78.  left/right index axis 1 ("tra"),
 * back/front index axis 0 ("lon"), bottom/top index axis 2 (node planes);
 * ld = front-back, offset = tot - ld*left - back. */
typedef struct {
    size_t bottom, top, left, right, back, front, ld;
    ptrdiff_t offset;
    double height;
} orc_active;

/* setupActiveRegions, electr3d.cpp:89-183, from a per-element junction number
 * (0 = none, k+1 = junction k; isActive(), electr3d.hpp:148-171).  Returns the number of
 * junctions found (<= max_act), -1 if a junction does not have flat top/bottom,
 * -2 if max_act is too small.  *condsize receives the junction-table length. */
int orc_setup_active(const orc_mesh* m, const uint32_t* elem_junc, orc_active* act, int max_act, size_t* condsize) {
    typedef struct { size_t bottom, top, left, right, back, front; int used; } region;
    region* regs = (region*)calloc((size_t)max_act + 1, sizeof(region));
    size_t nreg = 0;
    size_t e0 = m->n[0] - 1, e1 = m->n[1] - 1, e2 = m->n[2] - 1;
    int rc = 0;
    for (size_t lon = 0; lon < e0 && rc == 0; ++lon) {
        for (size_t tra = 0; tra < e1 && rc == 0; ++tra) {
            size_t num = 0, start = 0;
            for (size_t ver = 0; ver <= e2; ++ver) {
                size_t cur = (ver < e2) ? elem_junc[lon * m->es[0] + tra * m->es[1] + ver * m->es[2]] : 0;
                if (ver == e2 && num == 0) break;
                if (cur != num) {
                    if (num) {
                        if ((int)num > max_act) { rc = -2; break; }
                        region* r = &regs[num];
                        if (!r->used) {
                            r->used = 1; r->bottom = start; r->top = ver;
                            r->left = (size_t)-1; r->right = 0; r->back = (size_t)-1; r->front = 0;
                            if (nreg < num) nreg = num;
                            /* note: the reference creates the region with left=max,right=0 and only
                             * widens it on LATER columns (electr3d.cpp:116-128); first column included
                             * here the same way it ends up after the full scan */
                        } else if (start != r->bottom || ver != r->top) { rc = -1; break; }
                        if (tra < r->left) r->left = tra;
                        if (tra >= r->right) r->right = tra + 1;
                        if (lon < r->back) r->back = lon;
                        if (lon >= r->front) r->front = lon + 1;
                    }
                    num = cur;
                    start = ver;
                }
            }
        }
    }
    if (rc) { free(regs); return rc; }
    size_t tot = 0;
    for (size_t k = 1; k <= nreg; ++k) {
        region* r = &regs[k];
        orc_active* a = &act[k - 1];
        memset(a, 0, sizeof(*a));
        if (!r->used) continue;
        a->bottom = r->bottom; a->top = r->top; a->left = r->left; a->right = r->right;
        a->back = r->back; a->front = r->front; a->ld = r->front - r->back;
        a->offset = (ptrdiff_t)tot - (ptrdiff_t)((r->front - r->back) * r->left) - (ptrdiff_t)r->back;
        a->height = m->ax[2][r->top] - m->ax[2][r->bottom];
        tot += (r->right - r->left) * (r->front - r->back);
    }
    *condsize = tot;
    free(regs);
    return (int)nreg;
}

/* loadConductivity, electr3d.cpp:203-225.  elem_role: 0 none, 1 p-contact, 2 n-contact.
 * sigma tables are cond(T) per material id; Te is inTemperature at element midpoints.
 * junc_cond is [condsize][2]. */
void orc_shockley_load_conds(const orc_mesh* m, const uint32_t* elem_mat, const uint32_t* elem_junc,
                             const uint8_t* elem_role, const double* Te, uint32_t nT, double T0, double dT,
                             const double* s_lat, const double* s_vert, const orc_active* act,
                             const double* junc_cond, double pcond, double ncond, double* cond) {
    size_t E = orc_mesh_elements(m);
    for (size_t e = 0; e < E; ++e) {
        size_t ix[3];
        elem_indices(m, e, ix);
        uint32_t actn = elem_junc[e];
        if (actn) {
            const orc_active* a = &act[actn - 1];
            size_t j = (size_t)(a->offset + (ptrdiff_t)(a->ld * ix[1] + ix[0]));
            cond[2 * e] = junc_cond[2 * j];
            cond[2 * e + 1] = junc_cond[2 * j + 1];
            if (isnan(cond[2 * e + 1]) || fabs(cond[2 * e + 1]) < 1e-16) cond[2 * e + 1] = 1e-16;
        } else if (elem_role && elem_role[e] == 1) {
            cond[2 * e] = cond[2 * e + 1] = pcond;
        } else if (elem_role && elem_role[e] == 2) {
            cond[2 * e] = cond[2 * e + 1] = ncond;
        } else {
            cond[2 * e] = table_at(s_lat, elem_mat[e], nT, T0, dT, Te[e]);
            cond[2 * e + 1] = table_at(s_vert, elem_mat[e], nT, T0, dT, Te[e]);
        }
    }
}

/* saveConductivity, electr3d.cpp:227-237. */
void orc_shockley_save_conds(const orc_mesh* m, const orc_active* act, int nact, const double* cond, double* junc_cond) {
    for (int n = 0; n < nact; ++n) {
        const orc_active* a = &act[n];
        size_t v = (a->top + a->bottom) / 2;
        for (size_t t = a->left; t != a->right; ++t) {
            ptrdiff_t offset = a->offset + (ptrdiff_t)(a->ld * t);
            for (size_t l = a->back; l != a->front; ++l) {
                size_t e = l * m->es[0] + t * m->es[1] + v * m->es[2];
                junc_cond[2 * (offset + (ptrdiff_t)l)] = cond[2 * e];
                junc_cond[2 * (offset + (ptrdiff_t)l) + 1] = cond[2 * e + 1];
            }
        }
    }
}

/* Junction update, electr3d.cpp:246-274, with BetaSolver::activeCond, beta.hpp:43-46.
 * beta/js are given per junction-table entry (per lateral column): the host evaluates a
 * T-dependent beta(T)/js(T) (electr_python.cpp:103-110) at the mid-plane element
 * temperature, which is constant during one compute() call.  stable != 0 selects
 * CONVERGENCE_STABLE. */
void orc_shockley_junction_update(const orc_mesh* m, const uint32_t* elem_junc, const orc_active* act,
                                  const double* potential, const double* beta_col, const double* js_col,
                                  int stable, double* cond) {
    size_t E = orc_mesh_elements(m);
    for (size_t e = 0; e < E; ++e) {
        uint32_t nact = elem_junc[e];
        if (!nact) continue;
        size_t ix[3];
        elem_indices(m, e, ix);
        const orc_active* a = &act[nact - 1];
        size_t back = ix[0], front = ix[0] + 1, left = ix[1], right = ix[1] + 1;
#define NODE(i0, i1, i2) ((i0) * m->ns[0] + (i1) * m->ns[1] + (i2) * m->ns[2])
        double U = 0.25 * (-potential[NODE(back, left, a->bottom)] - potential[NODE(front, left, a->bottom)]
                           - potential[NODE(back, right, a->bottom)] - potential[NODE(front, right, a->bottom)]
                           + potential[NODE(back, left, a->top)] + potential[NODE(front, left, a->top)]
                           + potential[NODE(back, right, a->top)] + potential[NODE(front, right, a->top)]);
#undef NODE
        double jy = 0.1 * cond[2 * e + 1] * U / a->height;
        size_t col = (size_t)(a->offset + (ptrdiff_t)(a->ld * ix[1] + ix[0]));
        jy = fabs(jy);
        double c00 = 0., c11 = 10. * jy * a->height * beta_col[col] / log(1e7 * jy / js_col[col] + 1.);
        if (stable) { c00 = 0.5 * (cond[2 * e] + c00); c11 = 0.5 * (cond[2 * e + 1] + c11); }
        cond[2 * e] = c00;
        cond[2 * e + 1] = c11;
        if (isnan(cond[2 * e + 1]) || fabs(cond[2 * e + 1]) < 1e-16) cond[2 * e + 1] = 1e-16;
    }
}

/* Current densities and loop error, electr3d.cpp:387-425.  current is [E][3] (kA/cm^2),
 * updated in place; returns err (%, already 100*sqrt(err)/max(mcur,minj)); *mcur_out =
 * sqrt(max |j|^2) over junction elements (all elements if noactive); maxcur[3]. */
double orc_shockley_currents(const orc_mesh* m, const uint32_t* elem_junc, int noactive, const double* potential,
                             const double* cond, double* current, double* mcur_out, double maxcur[3]) {
    size_t E = orc_mesh_elements(m);
    double err = 0., mcur = 0.;
    const double minj = 100e-7;
    for (size_t e = 0; e < E; ++e) {
        size_t ix[3], n[8];
        elem_indices(m, e, ix);
        elem_nodes(m, ix, n);
        const double* p = potential;
        double lll = p[n[0]], ull = p[n[1]], lul = p[n[2]], uul = p[n[3]], llu = p[n[4]], ulu = p[n[5]], luu = p[n[6]], uuu = p[n[7]];
        double c0 = -0.025 * cond[2 * e] * (-lll - llu - lul - luu + ull + ulu + uul + uuu) / (m->ax[0][ix[0] + 1] - m->ax[0][ix[0]]);
        double c1 = -0.025 * cond[2 * e] * (-lll - llu + lul + luu - ull - ulu + uul + uuu) / (m->ax[1][ix[1] + 1] - m->ax[1][ix[1]]);
        double c2 = -0.025 * cond[2 * e + 1] * (-lll + llu - lul + luu - ull + ulu - uul + uuu) / (m->ax[2][ix[2] + 1] - m->ax[2][ix[2]]);
        if (noactive || elem_junc[e]) {
            double acur = c0 * c0 + c1 * c1 + c2 * c2;
            if (acur > mcur) { mcur = acur; maxcur[0] = c0; maxcur[1] = c1; maxcur[2] = c2; }
        }
        double d0 = current[3 * e] - c0, d1 = current[3 * e + 1] - c1, d2 = current[3 * e + 2] - c2;
        double delta = d0 * d0 + d1 * d1 + d2 * d2;
        if (delta > err) err = delta;
        current[3 * e] = c0; current[3 * e + 1] = c1; current[3 * e + 2] = c2;
    }
    mcur = sqrt(mcur);
    *mcur_out = mcur;
    return 100. * sqrt(err) / (mcur > minj ? mcur : minj);
}

/* saveHeatDensity, electr3d.cpp:444-478.  noheat[e] != 0 marks EMPTY material or the
 * "noheat" role (:472). heat in W/m^3. */
void orc_shockley_heat(const orc_mesh* m, const double* potential, const double* cond, const uint8_t* noheat, double* heat) {
    size_t E = orc_mesh_elements(m);
    for (size_t e = 0; e < E; ++e) {
        size_t ix[3], n[8];
        elem_indices(m, e, ix);
        elem_nodes(m, ix, n);
        const double* p = potential;
        double lll = p[n[0]], ull = p[n[1]], lul = p[n[2]], uul = p[n[3]], llu = p[n[4]], ulu = p[n[5]], luu = p[n[6]], uuu = p[n[7]];
        double dvx = -0.25e6 * (-lll - llu - lul - luu + ull + ulu + uul + uuu) / (m->ax[0][ix[0] + 1] - m->ax[0][ix[0]]);
        double dvy = -0.25e6 * (-lll - llu + lul + luu - ull - ulu + uul + uuu) / (m->ax[1][ix[1] + 1] - m->ax[1][ix[1]]);
        double dvz = -0.25e6 * (-lll + llu - lul + luu - ull + ulu - uul + uuu) / (m->ax[2][ix[2] + 1] - m->ax[2][ix[2]]);
        if (noheat && noheat[e]) heat[e] = 0.;
        else heat[e] = cond[2 * e] * dvx * dvx + cond[2 * e] * dvy * dvy + cond[2 * e + 1] * dvz * dvz;
    }
}

/* integrateCurrent, electr3d.cpp:480-497 (without the symmetry doubling: the flat
 * problem description has no mirror planes).  Returns mA. */
double orc_integrate_current(const orc_mesh* m, const uint32_t* elem_junc, const double* current, size_t vindex, int onlyactive) {
    double result = 0.;
    for (size_t i = 0; i < m->n[0] - 1; ++i)
        for (size_t j = 0; j < m->n[1] - 1; ++j) {
            size_t e = i * m->es[0] + j * m->es[1] + vindex * m->es[2];
            if (!onlyactive || elem_junc[e])
                result += current[3 * e + 2] * (m->ax[0][i + 1] - m->ax[0][i]) * (m->ax[1][j + 1] - m->ax[1][j]);
        }
    return result * 0.01;
}

/* getTotalHeat, electr3d.cpp:612-624 (mW). */
double orc_total_heat(const orc_mesh* m, const double* heat) {
    size_t E = orc_mesh_elements(m);
    double W = 0.;
    for (size_t e = 0; e < E; ++e) {
        size_t ix[3];
        elem_indices(m, e, ix);
        double d0 = m->ax[0][ix[0] + 1] - m->ax[0][ix[0]];
        double d1 = m->ax[1][ix[1] + 1] - m->ax[1][ix[1]];
        double d2 = m->ax[2][ix[2] + 1] - m->ax[2][ix[2]];
        W += 1e-15 * d0 * d1 * d2 * heat[e];
    }
    return W;
}

/* getTotalEnergy, electr3d.cpp:568-600: eps is the relative permittivity per element
 * (material->eps(T)); returns J. epsilon0 = 1/(mu0 c^2) F/m (plask/phys/constants.hpp:33-35). */
double orc_total_energy(const orc_mesh* m, const double* potential, const double* eps) {
    size_t E = orc_mesh_elements(m);
    const double epsilon0 = 1. / (4e-7 * 3.14159265358979323846) / 299792458. / 299792458.;
    double W = 0.;
    for (size_t e = 0; e < E; ++e) {
        size_t ix[3], n[8];
        elem_indices(m, e, ix);
        elem_nodes(m, ix, n);
        const double* p = potential;
        double lll = p[n[0]], ull = p[n[1]], lul = p[n[2]], uul = p[n[3]], llu = p[n[4]], ulu = p[n[5]], luu = p[n[6]], uuu = p[n[7]];
        double d0 = m->ax[0][ix[0] + 1] - m->ax[0][ix[0]];
        double d1 = m->ax[1][ix[1] + 1] - m->ax[1][ix[1]];
        double d2 = m->ax[2][ix[2] + 1] - m->ax[2][ix[2]];
        double dvx = -0.25e6 * (-lll - llu - lul - luu + ull + ulu + uul + uuu) / d0;
        double dvy = -0.25e6 * (-lll - llu + lul + luu - ull - ulu + uul + uuu) / d1;
        double dvz = -0.25e6 * (-lll + llu - lul + luu - ull + ulu - uul + uuu) / d2;
        double w = eps[e] * (dvx * dvx + dvy * dvy + dvz * dvz);
        W += 0.5e-18 * epsilon0 * d0 * d1 * d2 * w;
    }
    return W;
}

/* ------------------------- plain Jacobi-PCG on the 14-diagonal storage ------ */

/* Textbook PCG (the recurrence of itcg, extlib/nspcg/nspcg.f:9217-9337, with the Jacobi
 * preconditioner of jac2 :1577 and a TRUE-residual stopping test ||r||2 <= tol*||b||2
 * instead of NSPCG's eigenvalue-scaled test #2).  Always available on the GPU box; used
 * as the "port" CPU baseline and as a cross-check of oracle/_ref.  Returns iterations
 * (>= 0), or -1 on breakdown.  *relres receives ||b - A u||2/||b||2 (recomputed). */
int orc_pcg_jacobi_sparse14(size_t rank, const int icords[14], const double* data, const double* rhs, double* u,
                            int itmax, double tol, double* relres) {
    double* r = (double*)malloc(rank * sizeof(double));
    double* z = (double*)malloc(rank * sizeof(double));
    double* p = (double*)malloc(rank * sizeof(double));
    double* q = (double*)malloc(rank * sizeof(double));
    double bnorm = 0.;
    for (size_t i = 0; i < rank; ++i) bnorm += rhs[i] * rhs[i];
    bnorm = sqrt(bnorm);
    if (bnorm == 0.) bnorm = 1.;
    memset(q, 0, rank * sizeof(double));
    orc_addmult_sparse14(rank, icords, data, u, q);
    double rho = 0., rr = 0.;
    for (size_t i = 0; i < rank; ++i) {
        r[i] = rhs[i] - q[i];
        z[i] = r[i] / data[i];
        p[i] = z[i];
        rho += r[i] * z[i];
        rr += r[i] * r[i];
    }
    int it = 0, rc = 0;
    while (sqrt(rr) > tol * bnorm && it < itmax) {
        memset(q, 0, rank * sizeof(double));
        orc_addmult_sparse14(rank, icords, data, p, q);
        double pq = 0.;
        for (size_t i = 0; i < rank; ++i) pq += p[i] * q[i];
        if (!(pq > 0.)) { rc = -1; break; }
        double alpha = rho / pq, rho1 = 0.;
        rr = 0.;
        for (size_t i = 0; i < rank; ++i) {
            u[i] += alpha * p[i];
            r[i] -= alpha * q[i];
            z[i] = r[i] / data[i];
            rho1 += r[i] * z[i];
            rr += r[i] * r[i];
        }
        double beta = rho1 / rho;
        rho = rho1;
        for (size_t i = 0; i < rank; ++i) p[i] = z[i] + beta * p[i];
        ++it;
    }
    memset(q, 0, rank * sizeof(double));
    orc_addmult_sparse14(rank, icords, data, u, q);
    rr = 0.;
    for (size_t i = 0; i < rank; ++i) { double d = rhs[i] - q[i]; rr += d * d; }
    if (relres) *relres = sqrt(rr) / bnorm;
    free(r); free(z); free(p); free(q);
    return rc ? rc : it;
}

/* ------------------------------------------------------------------ provider interpolation */
/* RectilinearMesh3D::interpolateLinear (plask/mesh/rectilinear3d.hpp:802-845) for a geometry without
 * symmetry/periodicity: per axis prepareInterpolationForAxis (plask/mesh/axis1d.cpp:99-156) — index_hi =
 * upper_bound(axis, p); outside the axis both indices coincide (constant extrapolation) and the fake
 * coordinate is axis end -/+ 1 — then interpolation::trilinear (plask/utils/interpolation.hpp:31-71).
 * src: data on a rectilinear mesh with axes sa0..2 (sizes sn) and linear index i0*ss[0]+i1*ss[1]+i2*ss[2];
 * destination points: the tensor product of da0 x da1 x da2 (sizes dn), written at j0*ds[0]+j1*ds[1]+j2*ds[2]. */
static void prep_axis(const double* ax, size_t n, double p, size_t* ilo, size_t* ihi, double* lo, double* hi) {
    size_t up = 0, cnt = n;            /* std::upper_bound */
    while (cnt > 0) {
        size_t step = cnt / 2, it = up + step;
        if (!(p < ax[it])) { up = it + 1; cnt -= step + 1; } else cnt = step;
    }
    *ihi = up;
    if (up == 0) { *ilo = 0; *lo = ax[0] - 1.; } else { *ilo = up - 1; *lo = ax[up - 1]; }
    if (up == n) { *ihi = n - 1; *hi = ax[n - 1] + 1.; } else *hi = ax[up];
}

void orc_interp_linear(const size_t sn[3], const double* sa0, const double* sa1, const double* sa2, const size_t ss[3],
                       const double* src, const size_t dn[3], const double* da0, const double* da1, const double* da2,
                       const size_t ds[3], double* dst) {
    for (size_t j0 = 0; j0 < dn[0]; ++j0) {
        size_t l0, h0; double back, front;
        prep_axis(sa0, sn[0], da0[j0], &l0, &h0, &back, &front);
        for (size_t j1 = 0; j1 < dn[1]; ++j1) {
            size_t l1, h1; double left, right;
            prep_axis(sa1, sn[1], da1[j1], &l1, &h1, &left, &right);
            for (size_t j2 = 0; j2 < dn[2]; ++j2) {
                size_t l2, h2; double bottom, top;
                prep_axis(sa2, sn[2], da2[j2], &l2, &h2, &bottom, &top);
                const double px = da0[j0], py = da1[j1], pz = da2[j2];
#define D(a, b, c) src[(a) * ss[0] + (b) * ss[1] + (c) * ss[2]]
                const double dxh = front - px, dxl = px - back;
                const double lo = ((D(l0, l1, l2) * dxh + D(h0, l1, l2) * dxl) * (right - py) +
                                   (D(l0, h1, l2) * dxh + D(h0, h1, l2) * dxl) * (py - left)) / (right - left) / (front - back);
                const double hi = ((D(l0, l1, h2) * dxh + D(h0, l1, h2) * dxl) * (right - py) +
                                   (D(l0, h1, h2) * dxh + D(h0, h1, h2) * dxl) * (py - left)) / (right - left) / (front - back);
#undef D
                dst[j0 * ds[0] + j1 * ds[1] + j2 * ds[2]] = lo + (pz - bottom) / (top - bottom) * (hi - lo);
            }
        }
    }
}
