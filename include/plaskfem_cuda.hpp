// plaskfem_cuda.hpp — header-only C++17 host adapter over the C ABI (plaskfem_cuda.h).
//
// This is the code the PLaSK solver plugin compiles in: it turns what
// ThermalFem3DSolver / ElectricalFem3DSolver hold after onInitialize() into the flat arrays of
// the C ABI and maps status codes back to the solver's exceptions, `iter_params` outputs and
// the `noconv` policy.  It deliberately depends on NOTHING from PLaSK (so it builds and is
// tested in this repository); the three places where the plugin supplies PLaSK objects are
// template parameters / callables:
//     material_key(e)      -> an integer identifying geometry->getMaterial(elem.getMidpoint())
//     thermk(id, T, h)     -> material->thermk(T, h)      (tabulated, no callbacks into Python
//     cond(id, T)          -> material->cond(T)            from the device)
//     junction_number(e)   -> isActive(elem) (electr3d.hpp:148-171)
//
// Reference behaviour restated here (paths relative to the PLaSK tree):
//     layer_thickness        therm3d.cpp:81-114          (row a2 of SURVEY.md §8)
//     setup_active_regions   electr3d.cpp:89-183         (row a12 set-up)
//     flatten_dirichlet      plask/common/fem/matrix.hpp:111-118
//     IterParams / noconv    plask/common/fem/iterative_matrix.hpp:25-94, 296-336
//     Solver loops           therm3d.cpp:281-340, electr3d.cpp:356-442 (on the device)
#ifndef PLASKFEM_CUDA_HPP
#define PLASKFEM_CUDA_HPP

#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "plaskfem_cuda.h"

namespace plaskfem {

// ---- errors: the plugin re-throws these as plask::ComputationError / plask::BadInput -----------------

struct ComputationError : std::runtime_error { using std::runtime_error::runtime_error; };
struct BadInput : std::invalid_argument { using std::invalid_argument::invalid_argument; };
struct NoDevice : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- mesh description (row a1) -----------------------------------------------------------------------

// Iteration orders of RectilinearMesh3D (rectilinear3d.hpp:349): ORDER_<major><medium><minor>.
enum IterationOrder { ORDER_012, ORDER_021, ORDER_102, ORDER_120, ORDER_201, ORDER_210 };

struct Mesh {
    std::vector<double> axis[3];   // coordinates in um, axis 2 vertical
    IterationOrder order = ORDER_012;

    size_t n(int a) const { return axis[a].size(); }
    size_t size() const { return n(0) * n(1) * n(2); }
    size_t elements() const { return (n(0) - 1) * (n(1) - 1) * (n(2) - 1); }
    void order_axes(int& major, int& medium, int& minor) const {
        static const int t[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
        major = t[order][0]; medium = t[order][1]; minor = t[order][2];
    }
    // node strides, rectilinear3d.cpp:20-32
    void strides(size_t s[3]) const {
        int mj, md, mn; order_axes(mj, md, mn);
        s[mn] = 1; s[md] = n(mn); s[mj] = n(mn) * n(md);
    }
    // element strides: the element mesh keeps the order (rectangular3d.cpp:20-22)
    void elem_strides(size_t s[3]) const {
        int mj, md, mn; order_axes(mj, md, mn);
        s[mn] = 1; s[md] = n(mn) - 1; s[mj] = (n(mn) - 1) * (n(md) - 1);
    }
    size_t node(size_t i0, size_t i1, size_t i2) const { size_t s[3]; strides(s); return i0 * s[0] + i1 * s[1] + i2 * s[2]; }
    size_t elem(size_t i0, size_t i1, size_t i2) const { size_t s[3]; elem_strides(s); return i0 * s[0] + i1 * s[1] + i2 * s[2]; }
    // setOptimalIterationOrder, rectilinear3d.cpp:74-85
    void set_optimal_order() {
        static const int t[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
        for (int o = 0; o < 6; ++o)
            if (n(t[o][2]) <= n(t[o][1]) && n(t[o][1]) <= n(t[o][0])) { order = (IterationOrder)o; return; }
        order = ORDER_210;
    }
};

// ---- row a2: thickness of the maximal vertical run of equal material ---------------------------------

// thickness[e] = height of the maximal run of vertically adjacent elements with the same material_key
// that contains e (therm3d.cpp:81-114).  material_key: callable (i0, i1, i2) -> comparable key.
template <typename KeyFn>
std::vector<double> layer_thickness(const Mesh& m, KeyFn&& material_key) {
    const size_t e0 = m.n(0) - 1, e1 = m.n(1) - 1, e2 = m.n(2) - 1;
    size_t es[3]; m.elem_strides(es);
    std::vector<double> th(m.elements(), std::numeric_limits<double>::quiet_NaN());
    for (size_t i0 = 0; i0 < e0; ++i0)
        for (size_t i1 = 0; i1 < e1; ++i1) {
            size_t start = 0;
            auto key = material_key(i0, i1, (size_t)0);
            for (size_t r = 1; r <= e2; ++r) {
                bool brk = (r == e2);
                decltype(key) k2 = key;
                if (!brk) { k2 = material_key(i0, i1, r); brk = !(k2 == key); }
                if (brk) {
                    const double h = m.axis[2][r] - m.axis[2][start];
                    for (size_t q = start; q < r; ++q) th[i0 * es[0] + i1 * es[1] + q * es[2]] = h;
                    start = r; key = k2;
                }
            }
        }
    return th;
}

// Table ids: one id per distinct (material key, thickness) pair, so that thermk(T, thickness) can be
// tabulated per id (pfem_set_materials).  Returns ids[E]; `reps` receives one representative element per id.
template <typename Key>
std::vector<uint32_t> material_ids(const std::vector<Key>& key_per_elem, const std::vector<double>& thickness,
                                   std::vector<size_t>* reps = nullptr) {
    std::map<std::pair<Key, double>, uint32_t> seen;
    std::vector<uint32_t> ids(key_per_elem.size());
    for (size_t e = 0; e < key_per_elem.size(); ++e) {
        auto it = seen.find({key_per_elem[e], thickness.empty() ? 0. : thickness[e]});
        if (it == seen.end()) {
            it = seen.emplace(std::make_pair(key_per_elem[e], thickness.empty() ? 0. : thickness[e]), (uint32_t)seen.size()).first;
            if (reps) reps->push_back(e);
        }
        ids[e] = it->second;
    }
    return ids;
}

// Conductivity tables on the grid T0 + i*dT: fn(id, T) -> (lateral, vertical) = Tensor2 (c00, c11).
struct Tables {
    double T0 = 250., dT = 0.25;
    uint32_t nmat = 0, nT = 0;
    std::vector<double> lat, vert;   // [nmat][nT]
};
template <typename Fn>
Tables sample_tables(uint32_t nmat, Fn&& fn, double T0 = 250., double dT = 0.25, uint32_t nT = 1601) {
    Tables t; t.T0 = T0; t.dT = dT; t.nmat = nmat; t.nT = nT;
    t.lat.resize((size_t)nmat * nT); t.vert.resize((size_t)nmat * nT);
    for (uint32_t id = 0; id < nmat; ++id)
        for (uint32_t i = 0; i < nT; ++i) {
            std::pair<double, double> c = fn(id, T0 + i * dT);
            t.lat[(size_t)id * nT + i] = c.first; t.vert[(size_t)id * nT + i] = c.second;
        }
    return t;
}

// ---- row a8: BoundaryConditionsWithMesh flattened in application order (matrix.hpp:111-118) -----------

struct Dirichlet {
    std::vector<size_t> node;
    std::vector<double> value;
    template <typename NodeRange>
    void add(const NodeRange& place, double v) { for (size_t r : place) { node.push_back(r); value.push_back(v); } }
    void add_node(size_t r, double v) { node.push_back(r); value.push_back(v); }
};

// ---- masked meshes: empty-elements = "exclude" (fem_solver.hpp:182-189) --------------------------------
//
// RectangularMaskedMesh3D keeps the elements whose material is not EMPTY and the nodes of those elements, both numbered in
// the order of the full mesh.  The C ABI stays in full-mesh numbering (excluded elements are marked PFEM_MAT_EXCLUDED in
// the material ids); this helper gives the plugin the two index maps to move its DataVectors across.
struct MaskedNumbering {
    static constexpr size_t NONE = std::numeric_limits<size_t>::max();
    std::vector<size_t> node_of_full, elem_of_full;   // full index -> masked index or NONE
    std::vector<size_t> full_of_node, full_of_elem;   // masked index -> full index
    // included(i0, i1, i2) -> true if the element is kept (material kind != EMPTY)
    template <typename Included>
    MaskedNumbering(const Mesh& m, Included included) {
        const size_t n0 = m.n(0), n1 = m.n(1), n2 = m.n(2);
        node_of_full.assign(m.size(), NONE);
        elem_of_full.assign(m.elements(), NONE);
        std::vector<uint8_t> used(m.size(), 0);
        for (size_t i0 = 0; i0 + 1 < n0; ++i0) for (size_t i1 = 0; i1 + 1 < n1; ++i1) for (size_t i2 = 0; i2 + 1 < n2; ++i2) {
            if (!included(i0, i1, i2)) continue;
            elem_of_full[m.elem(i0, i1, i2)] = 0;
            for (int l = 0; l < 8; ++l) used[m.node(i0 + (l & 1), i1 + ((l >> 1) & 1), i2 + ((l >> 2) & 1))] = 1;
        }
        for (size_t i = 0; i < used.size(); ++i) if (used[i]) { node_of_full[i] = full_of_node.size(); full_of_node.push_back(i); }
        for (size_t e = 0; e < elem_of_full.size(); ++e) if (elem_of_full[e] != NONE) { elem_of_full[e] = full_of_elem.size(); full_of_elem.push_back(e); }
    }
    // the same from a flag per element of the full mesh (element order of the mesh); an EMPTY vector means a full mesh
    MaskedNumbering(const Mesh& m, const std::vector<uint8_t>& included_per_elem)
        : MaskedNumbering(m, [&](size_t i0, size_t i1, size_t i2) { return included_per_elem.empty() || included_per_elem[m.elem(i0, i1, i2)] != 0; }) {
        if (!included_per_elem.empty() && included_per_elem.size() != m.elements()) throw BadInput("MaskedNumbering: one flag per element expected");
    }
    size_t node_to_full(size_t masked_node) const { return full_of_node[masked_node]; }
    size_t elem_to_full(size_t masked_elem) const { return full_of_elem[masked_elem]; }
    // material ids for pfem_set_materials: the caller's ids on kept elements, PFEM_MAT_EXCLUDED elsewhere
    std::vector<uint32_t> mark_excluded(std::vector<uint32_t> ids) const {
        for (size_t e = 0; e < ids.size(); ++e) if (elem_of_full[e] == NONE) ids[e] = PFEM_MAT_EXCLUDED;
        return ids;
    }
    // masked vector <-> full vector (NC interleaved components per entry)
    void nodes_to_masked(const double* full, double* masked) const { for (size_t k = 0; k < full_of_node.size(); ++k) masked[k] = full[full_of_node[k]]; }
    void nodes_to_full(const double* masked, double* full, double fill = 0.) const {
        for (size_t i = 0; i < node_of_full.size(); ++i) full[i] = node_of_full[i] == NONE ? fill : masked[node_of_full[i]];
    }
    void elems_to_masked(const double* full, double* masked, int nc = 1) const {
        for (size_t k = 0; k < full_of_elem.size(); ++k) for (int c = 0; c < nc; ++c) masked[k * nc + c] = full[full_of_elem[k] * nc + c];
    }
    void elems_to_full(const double* masked, double* full, int nc = 1, double fill = 0.) const {
        for (size_t e = 0; e < elem_of_full.size(); ++e) for (int c = 0; c < nc; ++c)
            full[e * nc + c] = elem_of_full[e] == NONE ? fill : masked[elem_of_full[e] * nc + c];
    }
};

// ---- the 2-D solvers on the brick kernels (therm2d.cpp, electr2d.cpp, femT2d.cpp) --------------------
//
// A RectangularMesh<2> (axis 0 = tran or r, axis 1 = vert) is handed to the library as a brick mesh with ONE element layer along
// a dummy longitudinal axis: for z-invariant data the brick operator, load and capacity restricted to a node plane are
// 0.5e-6 * thickness times the 4-node rectangle ones, so plane 0 of the result is the 2-D FEM solution.  This helper holds the
// index bookkeeping every 2-D solver needs: plane-0 numbering node (i0, i1) = i0 * n1 + i1, element (i0, i1) = i0 * (n1 - 1) + i1
// whatever the iteration order of the plugin's own mesh.
struct Embedding2D {
    size_t n0 = 0, n1 = 0;
    Mesh mesh;   // axis[0] = {0, thickness}, axis[1] = tran / r, axis[2] = vert, ORDER_012

    Embedding2D() {}
    Embedding2D(const std::vector<double>& x, const std::vector<double>& y, double thickness = 1.) : n0(x.size()), n1(y.size()) {
        if (n0 < 2 || n1 < 2) throw BadInput("Embedding2D: mesh needs at least 2 points per axis");
        if (!(thickness > 0.)) throw BadInput("Embedding2D: the dummy layer needs a positive thickness");
        mesh.axis[0] = {0., thickness};
        mesh.axis[1] = x;
        mesh.axis[2] = y;
        mesh.order = ORDER_012;
    }
    size_t plane() const { return n0 * n1; }                          // nodes per plane; the brick mesh has 2 * plane()
    size_t elements() const { return (n0 - 1) * (n1 - 1); }           // the brick mesh has the same elements
    size_t node(size_t i0, size_t i1) const { return i0 * n1 + i1; }  // on plane 0; + plane() on plane 1
    size_t elem(size_t i0, size_t i1) const { return i0 * (n1 - 1) + i1; }
    // element weights of the cylindrical solvers: midpoint.rad_r() of every element column (therm2d.cpp:353, electr2d.cpp:220),
    // for Context::set_axis_weight(1, ...)
    std::vector<double> radial_weights() const {
        std::vector<double> w(n0 - 1);
        for (size_t i = 0; i + 1 < n0; ++i) w[i] = 0.5 * (mesh.axis[1][i] + mesh.axis[1][i + 1]);
        return w;
    }
    // a condition of the first kind on a node of the 2-D mesh: both planes, application order kept (matrix.hpp:111-118)
    void add_dirichlet(Dirichlet& bc, size_t node2d, double value) const {
        if (node2d >= plane()) throw BadInput("Embedding2D: node outside the 2-D mesh");
        bc.add_node(node2d, value);
        bc.add_node(node2d + plane(), value);
    }
    // 2-D nodal field -> brick field (the same values on both planes) and back (plane 0)
    std::vector<double> lift(const double* f2d) const {
        std::vector<double> f(2 * plane());
        for (size_t i = 0; i < plane(); ++i) f[i] = f[i + plane()] = f2d[i];
        return f;
    }
    void restrict_to_plane(const double* brick, double* f2d) const { for (size_t i = 0; i < plane(); ++i) f2d[i] = brick[i]; }
    // element vectors of the library ([E][3], the longitudinal component is zero) -> the (tran / r, vert) pairs of the 2-D solvers
    void elem_vec2(const double* brick3, double* out2) const {
        for (size_t e = 0; e < elements(); ++e) { out2[2 * e] = brick3[3 * e + 1]; out2[2 * e + 1] = brick3[3 * e + 2]; }
    }
    // pfem_boundary::mode2d for Context::set_boundary: edge conditions of therm2d.cpp, Cartesian or cylindrical
    static int boundary_mode(bool cylindrical) { return cylindrical ? 2 : 1; }
};

// ---- row a7: boundary conditions of the 2nd / 3rd kind and radiation ----------------------------------
//
// heatflux_boundary / convection_boundary / radiation_boundary (therm3d.hpp:79-82) in the form setBoundaries reads
// them (therm3d.cpp:149): boundary_conditions.getValue(node), i.e. the value of the FIRST condition whose place
// contains the node (plask/mesh/boundary_conditions.hpp:182-186).  NV = number of scalars of the condition:
// 1 heat flux, 2 Convection {coeff, ambient} / Radiation {emissivity, ambient} (therm3d.hpp:28-57).
template <int NV>
struct NodeConditions {
    std::vector<uint8_t> has;
    std::vector<double> v[NV];
    bool empty() const { return has.empty(); }
    template <typename NodeRange>
    void add(size_t mesh_size, const NodeRange& place, const std::array<double, NV>& value) {
        if (has.empty()) { has.assign(mesh_size, 0); for (auto& a : v) a.assign(mesh_size, 0.); }
        for (size_t r : place) {
            if (r >= mesh_size) throw BadInput("boundary condition names a node outside the mesh");
            if (has[r]) continue;   // an earlier condition already covers the node
            has[r] = 1;
            for (int k = 0; k < NV; ++k) v[k][r] = value[k];
        }
    }
    void add_node(size_t mesh_size, size_t r, const std::array<double, NV>& value) {
        const size_t one[1] = {r};
        add(mesh_size, one, value);
    }
};

// ---- junctions (electr3d.hpp:28-78, electr3d.cpp:89-183) ---------------------------------------------

// junction_number(i0, i1, i2) -> 0 (not active) or k+1.  Throws like the reference when a junction does not
// have flat top/bottom.  Returns the Active list; *condsize = length of junction_conductivity.
template <typename JuncFn>
std::vector<pfem_junction> setup_active_regions(const Mesh& m, JuncFn&& junction_number, size_t* condsize,
                                                const std::string& solver_id = "") {
    struct Region { size_t bottom, top, left, right, back, front; };
    const size_t SZMAX = std::numeric_limits<size_t>::max();
    std::map<size_t, Region> regions;
    size_t nreg = 0;
    const size_t p0 = m.n(0) - 1, p1 = m.n(1) - 1, p2 = m.n(2) - 1;
    auto summarize = [&](size_t num, size_t start, size_t ver, size_t lon, size_t tra) {
        auto found = regions.find(num);
        if (found == regions.end()) {
            regions[num] = Region{start, ver, SZMAX, 0, SZMAX, 0};
            if (nreg < num) nreg = num;
        } else {
            Region& r = found->second;
            if (start != r.bottom || ver != r.top)
                throw ComputationError(solver_id + ": Junction " + std::to_string(num - 1) +
                                       " does not have top and bottom edges at constant heights");
            if (tra < r.left) r.left = tra;
            if (tra >= r.right) r.right = tra + 1;
            if (lon < r.back) r.back = lon;
            if (lon >= r.front) r.front = lon + 1;
        }
    };
    for (size_t lon = 0; lon < p0; ++lon)
        for (size_t tra = 0; tra < p1; ++tra) {
            size_t num = 0, start = 0;
            for (size_t ver = 0; ver < p2; ++ver) {
                size_t cur = junction_number(lon, tra, ver);
                if (cur != num) {
                    if (num) summarize(num, start, ver, lon, tra);
                    num = cur; start = ver;
                }
            }
            if (num) summarize(num, start, p2, lon, tra);
        }
    // NOTE (reference quirk kept): the column that CREATES a region does not widen it, so a junction made of a
    // single column keeps left = SIZE_MAX, right = 0; real junctions span many columns and are unaffected.
    std::vector<pfem_junction> active(nreg);
    for (auto& a : active) { a = pfem_junction{0, 0, 0, 0, 0, 0, 0, 0, 0.}; }
    size_t tot = 0;
    for (auto& ir : regions) {
        const Region& r = ir.second;
        pfem_junction a;
        a.bottom = r.bottom; a.top = r.top; a.left = r.left; a.right = r.right; a.back = r.back; a.front = r.front;
        if (r.left == SZMAX || r.back == SZMAX) { a.left = a.right = a.back = a.front = 0; }
        a.ld = a.front - a.back;
        a.offset = (ptrdiff_t)tot - (ptrdiff_t)(a.ld * a.left) - (ptrdiff_t)a.back;
        a.height = m.axis[2][r.top] - m.axis[2][r.bottom];
        active[ir.first - 1] = a;
        tot += (a.right - a.left) * (a.front - a.back);
    }
    if (condsize) *condsize = tot;
    return active;
}

// ---- iter_params (iterative_matrix.hpp:25-94) ---------------------------------------------------------

struct IterParams {
    enum NoConvergenceBehavior { NO_CONVERGENCE_ERROR, NO_CONVERGENCE_WARNING, NO_CONVERGENCE_CONTINUE };
    int maxit = 10000;
    double maxerr = 1e-8;   // relative residual ||r||/||b_free|| (north-star criterion)
    NoConvergenceBehavior no_convergence_behavior = NO_CONVERGENCE_WARNING;
    // outputs (:90-93)
    bool converged = true;
    int iters = 0;
    double err = 0.;
    // iter_params.preconditioner (:27-46): NSPCG's 'jac' and 'ljac', and the multilevel line preconditioner 'mlj' that stands in
    // for the reference's default 'ic' (a sequential factorisation; mlj reaches its strength with parallel line solves)
    enum Preconditioner { PRECOND_JAC = 0, PRECOND_LJAC = 1, PRECOND_MLJ = 2 };
    Preconditioner preconditioner = PRECOND_JAC;
};

typedef std::function<void(int level /*0 error .. 3 result, 4 detail*/, const std::string&)> LogFn;

// ---- RAII context -------------------------------------------------------------------------------------

class Context {
    pfem_ctx* ctx_ = nullptr;
    std::string id_;

    void check(int rc) const {
        if (rc >= 0) return;
        std::string msg = id_ + ": " + pfem_strerror(rc);
        const char* d = pfem_last_error(ctx_);
        if (d && *d) msg += std::string(" (") + d + ")";
        switch (rc) {
            case PFEM_ERR_BAD_INPUT: case PFEM_ERR_STATE: throw BadInput(msg);
            case PFEM_ERR_NO_DEVICE: throw NoDevice(msg);
            default: throw ComputationError(msg);
        }
    }

  public:
    explicit Context(int device = 0, std::string solver_id = "") : id_(std::move(solver_id)) {
        int rc = pfem_create(&ctx_, device);
        if (rc != PFEM_OK) { ctx_ = nullptr; throw NoDevice(id_ + ": " + pfem_strerror(rc)); }
    }
    ~Context() { if (ctx_) pfem_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    pfem_ctx* get() const { return ctx_; }

    void set_mesh(const Mesh& m) {
        if (m.n(0) < 2 || m.n(1) < 2 || m.n(2) < 2) throw BadInput(id_ + ": mesh needs at least 2 points per axis");
        size_t n[3] = {m.n(0), m.n(1), m.n(2)}, s[3];
        m.strides(s);
        check(pfem_set_mesh(ctx_, n, m.axis[0].data(), m.axis[1].data(), m.axis[2].data(), s));
    }
    // cylindrical 2-D solvers on a one-layer brick mesh: element weights = midpoint radii along the radial axis
    void set_axis_weight(int axis, const std::vector<double>& w) { check(pfem_set_axis_weight(ctx_, axis, w.empty() ? nullptr : w.data())); }
    void set_materials(const std::vector<uint32_t>& ids, const Tables& t) {
        check(pfem_set_materials(ctx_, ids.data(), t.nmat, t.T0, t.dT, t.nT, t.lat.data(), t.vert.data()));
    }
    void set_dirichlet(const Dirichlet& bc) { check(pfem_set_dirichlet(ctx_, bc.node.size(), bc.node.data(), bc.value.data())); }
    // verbatim = true reproduces setBoundaries to the letter (local slots, quarter mass matrix; plaskfem_cuda.h); mode2d = 1 / 2: the
    // edge conditions of ThermalFem2DSolver<Cartesian / Cylindrical> on the one-layer embedding (therm2d.cpp:138-172)
    void set_boundary(const NodeConditions<1>& heatflux, const NodeConditions<2>& convection, const NodeConditions<2>& radiation,
                      bool verbatim = true, int mode2d = 0) {
        if (heatflux.empty() && convection.empty() && radiation.empty()) { check(pfem_set_boundary(ctx_, nullptr)); return; }
        pfem_boundary b;
        memset(&b, 0, sizeof b);
        if (!heatflux.empty()) { b.has_flux = heatflux.has.data(); b.flux = heatflux.v[0].data(); }
        if (!convection.empty()) { b.has_conv = convection.has.data(); b.conv_coeff = convection.v[0].data(); b.conv_ambient = convection.v[1].data(); }
        if (!radiation.empty()) { b.has_rad = radiation.has.data(); b.rad_emissivity = radiation.v[0].data(); b.rad_ambient = radiation.v[1].data(); }
        b.verbatim = verbatim ? 1 : 0;
        b.mode2d = mode2d;
        check(pfem_set_boundary(ctx_, &b));
    }
    void set_source(const double* heat_per_elem) { check(pfem_set_source(ctx_, heat_per_elem)); }
    void set_field(const double* x0) { check(pfem_set_field(ctx_, x0)); }
    void fill_field(double v) { check(pfem_fill_field(ctx_, v)); }
    void set_elem_temperature(const double* Te, double uniform = 300.) { check(pfem_set_elem_temperature(ctx_, Te, uniform)); }
    void set_junctions(const std::vector<pfem_junction>& act, const std::vector<uint32_t>& elem_junc,
                       const std::vector<uint8_t>& elem_role, double pcond, double ncond, const std::vector<double>& junc_cond,
                       const std::vector<double>& beta_col, const std::vector<double>& js_col, bool stable) {
        check(pfem_set_junctions(ctx_, (uint32_t)act.size(), act.data(), elem_junc.data(), elem_role.empty() ? nullptr : elem_role.data(),
                                 pcond, ncond, beta_col.size(), junc_cond.data(), beta_col.data(), js_col.data(), stable ? 1 : 0));
    }
    // internal layout, before set_mesh: PFEM_LAYOUT_VERTICAL_MINOR keeps the vertical lines contiguous for the 'ljac' preconditioner
    void set_layout(int layout) { check(pfem_set_layout(ctx_, layout)); }
    // EMPTY-material / "noheat" elements produce no Joule heat (electr3d.cpp:472)
    void set_noheat(const std::vector<uint8_t>& noheat) { check(pfem_set_noheat(ctx_, noheat.empty() ? nullptr : noheat.data())); }
    // Device-resident field exchange of the ThermoElectric meta loop (thermoelectric.py:207-211): call on the RECEIVING solver.
    //   electrical.take_temperature_from(thermal)  ==  inTemperature(elementMesh) <- thermal.outTemperature   (electr3d.cpp:203-205)
    //   thermal.take_heat_from(electrical)         ==  inHeat(elementMesh)        <- electrical.outHeat      (therm3d.cpp:179)
    // Both interpolate linearly like RectilinearMesh3D::interpolateLinear (rectilinear3d.hpp:802-845); the meshes may differ.
    void take_temperature_from(const Context& thermal) { check(pfem_transfer_temperature(ctx_, thermal.ctx_)); }
    void take_heat_from(const Context& electrical) {
        int rc = pfem_transfer_heat(ctx_, electrical.ctx_);
        if (rc < 0) { const char* d = pfem_last_error(electrical.ctx_); if (d && *d) throw ComputationError(id_ + ": " + d); }
        check(rc);
    }
    // Slab mode (one Context per GPU / process): the local mesh holds the owned planes of the major axis plus one halo plane
    // per neighbour; exchange the blobs of all ranks with any host transport, then connect.  Solves are collective afterwards.
    void slab_configure(int rank, int nranks, size_t own_lo, size_t own_hi) { check(pfem_slab_configure(ctx_, rank, nranks, own_lo, own_hi)); }
    std::vector<unsigned char> slab_export() {
        std::vector<unsigned char> blob(pfem_slab_blob_size());
        check(pfem_slab_export(ctx_, blob.data()));
        return blob;
    }
    void slab_connect(const std::vector<unsigned char>& all_blobs) {
        if (all_blobs.size() % pfem_slab_blob_size() != 0) throw BadInput(id_ + ": slab blobs have the wrong size");
        check(pfem_slab_connect(ctx_, all_blobs.data()));
    }
    void get_field(double* x) { check(pfem_get_field(ctx_, x)); }
    // provider on a foreign rectilinear mesh (getTemperatures / getVoltage with INTERPOLATION_LINEAR): out in `dst`'s own order
    void interpolate_field(const Mesh& dst, double* out) {
        size_t n[3] = {dst.n(0), dst.n(1), dst.n(2)}, s[3];
        dst.strides(s);
        check(pfem_interpolate_field(ctx_, n, dst.axis[0].data(), dst.axis[1].data(), dst.axis[2].data(), s, out));
    }
    void get_elem(int what, double* out, const uint8_t* noheat = nullptr) { check(pfem_get_elem(ctx_, what, noheat, out)); }
    void get_junction_cond(double* jc) { check(pfem_get_junction_cond(ctx_, jc)); }

    struct LoopResult { int loops, loopno; double err, toterr, maxval; double maxcur[3]; long long lin_iters; double ms; };

    // The do..while of compute() on the device.  Fills iter_params like SparseMatrix::solverhs does
    // (iterative_matrix.hpp:330-336) and applies the noconv policy (:296-314).
    LoopResult solve(bool thermal, IterParams& ip, double maxerr, int loops, const LogFn& log = LogFn()) {
        pfem_opts o;
        pfem_default_opts(&o);
        o.maxit = ip.maxit; o.lin_tol = ip.maxerr; o.outer_tol = maxerr; o.loops = loops; o.precond = (int)ip.preconditioner;
        pfem_stats st;
        int rc = thermal ? pfem_solve_thermal(ctx_, &o, &st) : pfem_solve_shockley(ctx_, &o, &st);
        check(rc);
        ip.converged = st.converged != 0; ip.iters = st.last_iters; ip.err = st.lin_relres;
        if (rc == PFEM_NOT_CONVERGED) {
            char buf[160];
            snprintf(buf, sizeof buf, "Failed to converge in %d iterations (error %g)", ip.maxit, ip.err);
            switch (ip.no_convergence_behavior) {
                case IterParams::NO_CONVERGENCE_ERROR: throw ComputationError(id_ + ": " + buf);
                case IterParams::NO_CONVERGENCE_WARNING: if (log) log(1, buf); break;
                case IterParams::NO_CONVERGENCE_CONTINUE: if (log) log(4, buf); break;
            }
        } else if (log) {
            char buf[160];
            snprintf(buf, sizeof buf, "Conjugate gradient converged after %d iterations (error %g)", ip.iters, ip.err);
            log(4, buf);
        }
        LoopResult r{st.outer_loops, st.loopno, st.err, st.toterr, st.maxval, {st.maxcur[0], st.maxcur[1], st.maxcur[2]},
                     st.lin_iters, st.t_solve_ms};
        if (log) {
            char buf[200];
            if (thermal) snprintf(buf, sizeof buf, "Loop %d(%d): max(T) = %.3f K, error = %g K", r.loops, r.loopno, r.maxval, r.err);
            else snprintf(buf, sizeof buf, "Loop %d(%d): max(j%s) = %g kA/cm2, error = %g%%", r.loops, r.loopno, "@junc", r.maxval, r.err);
            log(3, buf);
        }
        return r;
    }

    // ---- Dynamic3D (thermal.dynamic.Dynamic3D, femT3d.cpp) ----------------------------------------------
    // cp(T)*dens(T) per material id on the temperature grid of the Tables given to set_materials (femT3d.cpp:176)
    void set_capacity(const Tables& t, const std::vector<double>& cp_dens) {
        if (cp_dens.size() != (size_t)t.nmat * t.nT) throw BadInput(id_ + ": capacity table must be [nmat][nT]");
        check(pfem_set_capacity(ctx_, t.nmat, t.nT, cp_dens.data()));
    }
    struct TimeResult { int steps; double maxT; long long lin_iters; double ms; };
    // The time loop of DynamicThermalFem3DSolver::compute(time) (femT3d.cpp:258-305, corrected update — plaskfem_cuda.h).
    // `elapstime` is advanced like :293,297 (steps * timestep - timestep: the loop makes time/timestep + 1 solves while the clock
    // shows `time`); every `logfreq` steps the LOG_RESULT line of :287-291 goes to `log`.
    TimeResult solve_dynamic(IterParams& ip, double time, double timestep, double methodparam, bool lumping, int rebuildfreq,
                             int logfreq, double& elapstime, const LogFn& log = LogFn()) {
        pfem_opts o;
        pfem_default_opts(&o);
        o.maxit = ip.maxit; o.lin_tol = ip.maxerr; o.precond = (int)ip.preconditioner;
        if (!lumping) o.variant = 1;     // the consistent capacity matrix runs the node-per-thread operator kernel
        pfem_dynamic d;
        memset(&d, 0, sizeof d);
        d.time = time; d.timestep = timestep; d.methodparam = methodparam; d.lumping = lumping ? 1 : 0; d.rebuildfreq = rebuildfreq;
        std::vector<double> maxlog;
        if (log && logfreq > 0) {
            maxlog.resize((size_t)((time + timestep / 2.) / timestep) + 2);
            d.maxT_log = maxlog.data(); d.maxT_log_len = maxlog.size();
        }
        pfem_stats st;
        int rc = pfem_solve_dynamic(ctx_, &o, &d, &st);
        check(rc);
        ip.converged = st.converged != 0; ip.iters = st.last_iters; ip.err = st.lin_relres;
        if (rc == PFEM_NOT_CONVERGED) {
            char buf[160];
            snprintf(buf, sizeof buf, "Failed to converge in %d iterations (error %g)", ip.maxit, ip.err);
            switch (ip.no_convergence_behavior) {
                case IterParams::NO_CONVERGENCE_ERROR: throw ComputationError(id_ + ": " + buf);
                case IterParams::NO_CONVERGENCE_WARNING: if (log) log(1, buf); break;
                case IterParams::NO_CONVERGENCE_CONTINUE: if (log) log(4, buf); break;
            }
        }
        if (log && logfreq > 0) {
            int l = logfreq;
            for (int i = 0; i < st.outer_loops && (size_t)i < maxlog.size(); ++i) {
                if (l == 0) {
                    char buf[120];
                    snprintf(buf, sizeof buf, "Time %.2f ns: max(T) = %.3f K", elapstime + i * timestep, maxlog[(size_t)i]);
                    log(3, buf);
                    l = logfreq;
                }
                --l;
            }
        }
        elapstime += st.outer_loops * timestep - (st.outer_loops ? timestep : 0.);
        return TimeResult{st.outer_loops, st.maxval, st.lin_iters, st.t_solve_ms};
    }
};

}  // namespace plaskfem
#endif  // PLASKFEM_CUDA_HPP
