// mock_plask.hpp — a MODEL of the slice of the PLaSK API that the code added by patches/plask-algorithm-cuda.diff touches
// (tests/test_patch_compiles.py).  The real tree cannot be built here (no Boost, no LAPACK), so the functions the patch adds
// to the 2-D solvers are compiled against this model instead: it catches wrong adapter signatures, missing members and
// template mistakes in the new code; it says nothing about the rest of PLaSK.  Names and shapes follow plask/*.hpp
// (mesh/rectangular2d.hpp, mesh/rectangular_masked2d.hpp, material/material.hpp, data.hpp, lazydata.hpp,
// mesh/boundary_conditions.hpp, common/fem/fem_solver.hpp, common/fem/iterative_matrix.hpp:27-80).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <limits>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#define PLASK_UNUSED(x)
#define PLASK_SOLVER_API

namespace plask {

using std::shared_ptr;
using std::isnan;
using std::abs;
using std::max;

enum LogLevel { LOG_CRITICAL_ERROR, LOG_ERROR, LOG_ERROR_DETAIL, LOG_WARNING, LOG_IMPORTANT, LOG_INFO, LOG_RESULT, LOG_DATA, LOG_DETAIL, LOG_DEBUG };

struct Exception : std::runtime_error {
    template <typename... A> Exception(const std::string& fmt, A&&...) : std::runtime_error(fmt) {}
};
struct ComputationError : Exception {
    template <typename... A> ComputationError(const std::string& id, const std::string& fmt, A&&...) : Exception(id + ": " + fmt) {}
};
struct BadInput : Exception {
    template <typename... A> BadInput(const std::string& id, const std::string& fmt, A&&...) : Exception(id + ": " + fmt) {}
};

template <typename T = double> struct Tensor2 {
    T c00, c11;
    Tensor2(T a = T(), T b = T()) : c00(a), c11(b) {}
};
template <int D, typename T = double> struct Vec;
template <typename T> struct Vec<2, T> { T c0, c1; };
template <typename T> struct Vec<3, T> { T c0, c1, c2; };
inline Vec<2, double> vec(double a, double b) { return Vec<2, double>{a, b}; }
inline Vec<3, double> vec(double a, double b, double c) { return Vec<3, double>{a, b, c}; }

struct Material {
    enum Kind { GENERIC, EMPTY, SEMICONDUCTOR, OXIDE, DIELECTRIC, METAL, LIQUID_CRYSTAL, MIXED };
    virtual ~Material() {}
    virtual Kind kind() const { return GENERIC; }
    virtual Tensor2<double> thermk(double, double = std::numeric_limits<double>::infinity()) const { return Tensor2<double>(1., 1.); }
    virtual Tensor2<double> cond(double) const { return Tensor2<double>(1., 1.); }
    virtual double cp(double) const { return 1.; }
    virtual double dens(double) const { return 1.; }
};

template <typename T> struct DataVector {
    std::vector<T> v;
    size_t size() const { return v.size(); }
    T& operator[](size_t i) { return v[i]; }
    const T& operator[](size_t i) const { return v[i]; }
    T* data() { return v.data(); }
    const T* data() const { return v.data(); }
    DataVector claim() const { return *this; }
    void reset() { v.clear(); }
    void reset(size_t n, const T& x = T()) { v.assign(n, x); }
    explicit operator bool() const { return !v.empty(); }
    typename std::vector<T>::iterator begin() { return v.begin(); }
    typename std::vector<T>::iterator end() { return v.end(); }
};
template <typename T> struct LazyData {
    std::vector<T> v;
    T operator[](size_t i) const { return v[i]; }
    size_t size() const { return v.size(); }
};

struct MeshAxis {
    std::vector<double> p;
    size_t size() const { return p.size(); }
    double at(size_t i) const { return p[i]; }
};
template <int D> struct MeshD { virtual ~MeshD() {} };
template <int D> struct RectangularMesh : MeshD<D> {
    struct Boundary {};
    enum IterationOrder { ORDER_012, ORDER_021, ORDER_102, ORDER_120, ORDER_201, ORDER_210 };   // rectilinear3d.hpp:349
    struct Element {
        size_t idx;
        size_t getIndex() const { return idx; }
    };
    shared_ptr<MeshAxis> axis[D];
    size_t size() const { size_t n = 1; for (int a = 0; a < D; ++a) n *= axis[a]->size(); return n; }
    size_t getElementsCount() const { size_t n = 1; for (int a = 0; a < D; ++a) n *= axis[a]->size() - 1; return n; }
    IterationOrder getIterationOrder() const { return ORDER_012; }
    Element element(size_t, size_t, size_t) const { return Element{0}; }
};

struct RectangularMaskedMesh2D : MeshD<2> {
    static constexpr size_t NOT_INCLUDED = std::numeric_limits<size_t>::max();
    struct Element {
        static constexpr size_t UNKNOWN_ELEMENT_INDEX = std::numeric_limits<size_t>::max();
        size_t i0, i1, idx, n1;
        size_t getIndex0() const { return i0; }
        size_t getIndex1() const { return i1; }
        size_t getIndex() const { return idx; }
        size_t getLoLoIndex() const { return i0 * n1 + i1; }
        size_t getUpLoIndex() const { return (i0 + 1) * n1 + i1; }
        size_t getLoUpIndex() const { return i0 * n1 + i1 + 1; }
        size_t getUpUpIndex() const { return (i0 + 1) * n1 + i1 + 1; }
        Vec<2, double> getMidpoint() const { return Vec<2, double>{0., 0.}; }
    };
    std::vector<Element> elems;
    size_t nodes = 0;
    const std::vector<Element>& elements() const { return elems; }
    size_t size() const { return nodes; }
    size_t getElementsCount() const { return elems.size(); }
    bool full() const { return true; }
    shared_ptr<MeshD<2>> getElementMesh() const { return shared_ptr<MeshD<2>>(); }
    Element element(size_t i0, size_t i1) const { return Element{i0, i1, 0, 0}; }
};

struct RectangularMaskedMesh3D : MeshD<3> {
    struct Element {
        static constexpr size_t UNKNOWN_ELEMENT_INDEX = std::numeric_limits<size_t>::max();
        size_t i0, i1, i2, idx;
        size_t getIndex0() const { return i0; }
        size_t getIndex1() const { return i1; }
        size_t getIndex2() const { return i2; }
        size_t getIndex() const { return idx; }
        Vec<3, double> getMidpoint() const { return Vec<3, double>{0., 0., 0.}; }
    };
    std::vector<Element> elems;
    size_t nodes = 0;
    const std::vector<Element>& elements() const { return elems; }
    size_t size() const { return nodes; }
    size_t getElementsCount() const { return elems.size(); }
    bool full() const { return true; }
    shared_ptr<MeshD<3>> getElementMesh() const { return shared_ptr<MeshD<3>>(); }
    Element element(size_t i0, size_t i1, size_t i2) const { return Element{i0, i1, i2, 0}; }
};
template <int D> struct MaskedOf;
template <> struct MaskedOf<2> { typedef RectangularMaskedMesh2D type; };
template <> struct MaskedOf<3> { typedef RectangularMaskedMesh3D type; };
template <int D> using RectangularMaskedMesh = typename MaskedOf<D>::type;
template <typename MeshT> struct DimOf;
template <int D> struct DimOf<RectangularMesh<D>> { static constexpr int value = D; };

template <typename BoundaryT, typename ValueT> struct BoundaryConditionsWithMesh {
    struct Condition { std::vector<size_t> place; ValueT value; };
    std::vector<Condition> conds;
    typename std::vector<Condition>::const_iterator begin() const { return conds.begin(); }
    typename std::vector<Condition>::const_iterator end() const { return conds.end(); }
};

struct Geometry2DCartesian {
    shared_ptr<Material> getMaterial(const Vec<2, double>&) const { return std::make_shared<Material>(); }
    std::set<std::string> getRolesAt(const Vec<2, double>&) const { return std::set<std::string>(); }
};
struct Geometry2DCylindrical : Geometry2DCartesian {};
struct Geometry3D {
    shared_ptr<Material> getMaterial(const Vec<3, double>&) const { return std::make_shared<Material>(); }
    std::set<std::string> getRolesAt(const Vec<3, double>&) const { return std::set<std::string>(); }
};

template <typename SpaceT> struct ReceiverModel {
    template <typename MeshPtr> LazyData<double> operator()(const MeshPtr&) const { return LazyData<double>(); }
    ReceiverModel& operator=(double) { return *this; }
};
struct ProviderModel { void fireChanged() {} };

enum FemMatrixAlgorithm { ALGORITHM_CHOLESKY, ALGORITHM_GAUSS, ALGORITHM_ITERATIVE, ALGORITHM_CUDA };

struct IterativeMatrixParams {
    enum Accelerator { ACCEL_CG };
    enum Preconditioner { PRECOND_JAC, PRECOND_NEU, PRECOND_LSP, PRECOND_SOR, PRECOND_SSOR, PRECOND_IC, PRECOND_MIC, PRECOND_LSP_, PRECOND_LJAC };
    enum NoConvergenceBehavior { NO_CONVERGENCE_ERROR, NO_CONVERGENCE_WARNING, NO_CONVERGENCE_CONTINUE };
    Preconditioner preconditioner = PRECOND_IC;
    NoConvergenceBehavior no_convergence_behavior = NO_CONVERGENCE_WARNING;
    int maxit = 1000;
    double maxerr = 1e-6;
    bool converged = true;
    int iters = 0;
    double err = 0.;
};

template <typename SpaceT, typename MeshT> struct FemSolverWithMaskedMesh {
    shared_ptr<SpaceT> geometry = std::make_shared<SpaceT>();
    shared_ptr<MeshT> mesh = std::make_shared<MeshT>();
    shared_ptr<RectangularMaskedMesh<DimOf<MeshT>::value>> maskedMesh = std::make_shared<RectangularMaskedMesh<DimOf<MeshT>::value>>();
    FemMatrixAlgorithm algorithm = ALGORITHM_CUDA;
    IterativeMatrixParams iter_params;
    virtual ~FemSolverWithMaskedMesh() {}
    std::string getId() const { return "mock"; }
    template <typename... A> void writelog(LogLevel, const std::string&, A&&...) const {}
};

namespace thermal { namespace tstatic {
    struct Convection { double coeff, ambient; };
    struct Radiation { double emissivity, ambient; };
}}
namespace electrical { namespace shockley {
    enum Convergence { CONVERGENCE_FAST, CONVERGENCE_STABLE };
}}

}  // namespace plask

// ---- slice of the API used by the Diffusion3DSolver hunks (solvers/electrical/diffusion/diffusion3d.{hpp,cpp}) -------------------
#include <complex>
#include <map>
namespace plask {

typedef std::complex<double> dcomplex;
using std::real;
template <typename T> inline Tensor2<T> operator*(double a, const Tensor2<T>& t) { return Tensor2<T>(a * t.c00, a * t.c11); }
struct InterpolationMethod { enum Value { INTERPOLATION_DEFAULT, INTERPOLATION_LINEAR, INTERPOLATION_SPLINE }; };
struct Gain { enum EnumType { GAIN, DGDN }; };
constexpr double inv_hc = 1.0e-9 / (6.62607015e-34 * 299792458.);

struct RectangularMesh2D : RectangularMesh<2> {
    enum IterationOrder2D { ORDER_10, ORDER_01 };
    IterationOrder2D getIterationOrder() const { return ORDER_10; }
};
struct LateralMaskedModel {
    static constexpr size_t NOT_INCLUDED = std::numeric_limits<size_t>::max();
    RectangularMesh2D fullMesh;
    size_t getElementIndexFromLowIndexes(size_t, size_t) const { return 0; }
};
}  // namespace plask
// the hunk names RectangularMaskedMesh2D::NOT_INCLUDED (plask/mesh/rectangular_masked_common.hpp)
namespace plask { namespace electrical { namespace diffusion {
struct SizedMesh { size_t n = 0; size_t size() const { return n; } };
struct QwMesh2 : SizedMesh { shared_ptr<LateralMaskedModel> lateralMesh = std::make_shared<LateralMaskedModel>(); };
struct ActiveRegion3D {
    shared_ptr<QwMesh2> mesh2 = std::make_shared<QwMesh2>();
    shared_ptr<SizedMesh> emesh2 = std::make_shared<SizedMesh>();
    DataVector<double> U;
    std::vector<double> modesP;
    double QWheight = 0.;
};
struct ElementParams3D {
    double X = 1., Y = 1.;
    ElementParams3D(const ActiveRegion3D&, size_t) {}
};
inline Tensor2<double> integrateBilinear(double, double, const Tensor2<double>*) { return Tensor2<double>(); }
struct WavelengthReceiverModel { dcomplex operator()(size_t) const { return dcomplex(980., 0.); } };
struct GainReceiverModel {
    template <typename MeshPtr> LazyData<Tensor2<double>> operator()(const MeshPtr&, double, InterpolationMethod::Value) const { return LazyData<Tensor2<double>>(); }
    template <typename MeshPtr> LazyData<Tensor2<double>> operator()(Gain::EnumType, const MeshPtr&, double, InterpolationMethod::Value) const {
        return LazyData<Tensor2<double>>();
    }
};
}}}  // namespace plask::electrical::diffusion
