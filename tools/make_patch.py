"""Generate patches/plask-algorithm-cuda.diff: the binding of libplaskfem_cuda.so into the PLaSK tree as a unified diff
(`patch -p1` from the PLaSK root).  The edits are expressed against the text of /root/reference, so the script fails loudly
when the reference moves; tests/test_patch_applies.py applies the committed diff to a scratch copy of the touched files.

    python tools/make_patch.py            # needs /root/reference (build container only)

What the patch does (INTEGRATION.md explains every hunk):
  * a new FemMatrixAlgorithm value ALGORITHM_CUDA ('cuda' in XPL / Python / the GUI schema)
  * ThermalFem3DSolver / ElectricalFem3DSolver: a plaskfem::Context member, built in onInitialize(), and a compute()
    branch that hands the whole nonlinear loop to the library (include/plaskfem_cuda.hpp)
  * BetaSolver / the Python Shockley class expose beta(T), js(T) per junction through one virtual, so that the host can
    evaluate them at the mid-plane temperature of every junction column (electr3d.cpp:261-262)
  * ThermalFem2DSolver / DynamicThermalFem2DSolver / ElectricalFem2DSolver <Cartesian / Cylindrical>: the same through the one-layer embedding of the 2-D mesh
    (plaskfem::Embedding2D, radial element weights, the edge conditions of the 2nd / 3rd kind and radiation in the library's 2-D mode)
  * DynamicThermalFem3DSolver: the same for the time loop of compute(time) (pfem_solve_dynamic, corrected update)
  * Diffusion3DSolver: compute() hands the whole loop of one active region to pdiff_compute (include/plaskdiff_cuda.hpp)
  * the four solver CMakeLists link plaskfem_cuda
"""
import os
import shutil
import subprocess
import sys
import tempfile

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "patches", "plask-algorithm-cuda.diff")

EDITS = {}


def edit(path, old, new, count=1):
    EDITS.setdefault(path, []).append((old, new, count))


# ---------------------------------------------------------------- plask/common/fem/fem_solver.hpp
F = "plask/common/fem/fem_solver.hpp"
edit(F, """    ALGORITHM_GAUSS,     ///< Gauss elimination of asymmetric matrix (slower but safer as it uses pivoting)
    ALGORITHM_ITERATIVE  ///< Conjugate gradient iterative solver
};""", """    ALGORITHM_GAUSS,     ///< Gauss elimination of asymmetric matrix (slower but safer as it uses pivoting)
    ALGORITHM_ITERATIVE, ///< Conjugate gradient iterative solver
    ALGORITHM_CUDA       ///< Matrix-free conjugate gradients on a CUDA device (libplaskfem_cuda, no FemMatrix object)
};""")
edit(F, """                            .value("iterative", ALGORITHM_ITERATIVE)
                            .get(algorithm);""", """                            .value("iterative", ALGORITHM_ITERATIVE)
                            .value("cuda", ALGORITHM_CUDA)
                            .get(algorithm);""")
edit(F, """        case ALGORITHM_ITERATIVE: return new SparseBandMatrix(this, this->mesh->size(), this->mesh->minorAxis()->size());
    }
    return nullptr;""", """        case ALGORITHM_ITERATIVE: return new SparseBandMatrix(this, this->mesh->size(), this->mesh->minorAxis()->size());
        case ALGORITHM_CUDA: throw NotImplemented(this->getId(), "matrix object for algorithm 'cuda' (the solve is matrix-free)");
    }
    return nullptr;""")
edit(F, """            return new SparseBandMatrix(this, this->mesh->size(), mesh->mediumAxis()->size() * mesh->minorAxis()->size(),
                                        mesh->minorAxis()->size());
    }
    return nullptr;""", """            return new SparseBandMatrix(this, this->mesh->size(), mesh->mediumAxis()->size() * mesh->minorAxis()->size(),
                                        mesh->minorAxis()->size());
        case ALGORITHM_CUDA: throw NotImplemented(this->getId(), "matrix object for algorithm 'cuda' (the solve is matrix-free)");
    }
    return nullptr;""")
edit(F, """        if (empty_elements == EMPTY_ELEMENTS_INCLUDED ||
            (this->algorithm == ALGORITHM_ITERATIVE && empty_elements == EMPTY_ELEMENTS_DEFAULT)) {""",
     """        if (empty_elements == EMPTY_ELEMENTS_INCLUDED ||
            ((this->algorithm == ALGORITHM_ITERATIVE || this->algorithm == ALGORITHM_CUDA) &&
             empty_elements == EMPTY_ELEMENTS_DEFAULT)) {""")
edit(F, """                return new SparseFreeMatrix(this, this->maskedMesh->size(), this->maskedMesh->elements().size() * 10);
    }
    return nullptr;""", """                return new SparseFreeMatrix(this, this->maskedMesh->size(), this->maskedMesh->elements().size() * 10);
        case ALGORITHM_CUDA: throw NotImplemented(this->getId(), "matrix object for algorithm 'cuda' (the solve is matrix-free)");
    }
    return nullptr;""")
edit(F, """                return new SparseFreeMatrix(this, this->maskedMesh->size(), this->maskedMesh->elements().size() * 36);
    }
    return nullptr;""", """                return new SparseFreeMatrix(this, this->maskedMesh->size(), this->maskedMesh->elements().size() * 36);
        case ALGORITHM_CUDA: throw NotImplemented(this->getId(), "matrix object for algorithm 'cuda' (the solve is matrix-free)");
    }
    return nullptr;""")

# ---------------------------------------------------------------- python enum, XPL schema
edit("python/plask/common/fem/fem.cpp", """        .value("ITERATIVE", ALGORITHM_ITERATIVE);""", """        .value("ITERATIVE", ALGORITHM_ITERATIVE)
        .value("CUDA", ALGORITHM_CUDA);""")
edit("plask/common/fem.yml", """      - gauss
      - iterative
    help: >
      Algorithm used for solving set of linear positive-definite equations.""", """      - gauss
      - iterative
      - cuda
    help: >
      Algorithm used for solving set of linear positive-definite equations. ``cuda`` runs the whole nonlinear loop
      of the 3D thermal and electrical solvers matrix-free on a CUDA device (conjugate gradients with the ``jac``,
      ``ljac`` or ``mlj`` preconditioner); it has no CPU fallback.""")

# ---------------------------------------------------------------- thermal.static Static3D
F = "solvers/thermal/static/therm3d.hpp"
edit(F, """#include <plask/plask.hpp>
#include <plask/common/fem.hpp>

#include "common.hpp"

namespace plask { namespace thermal { namespace tstatic {

/**
 * Solver performing calculations in 2D Cartesian or Cylindrical space using finite element method
 */
struct PLASK_SOLVER_API ThermalFem3DSolver:""", """#include <plask/plask.hpp>
#include <plask/common/fem.hpp>

#include "common.hpp"

namespace plaskfem { class Context; }   // plaskfem_cuda.hpp: host adapter of libplaskfem_cuda.so (algorithm 'cuda')

namespace plask { namespace thermal { namespace tstatic {

/**
 * Solver performing calculations in 2D Cartesian or Cylindrical space using finite element method
 */
struct PLASK_SOLVER_API ThermalFem3DSolver:""")
edit(F, """    DataVector<Vec<3,double>> fluxes;           ///< Computed (only when needed) heat fluxes on our own mesh
""", """    DataVector<Vec<3,double>> fluxes;           ///< Computed (only when needed) heat fluxes on our own mesh

    std::shared_ptr<plaskfem::Context> cuda;     ///< Device context, exists only for algorithm 'cuda'

    /// Create the device context: mesh, (material, layer thickness) ids and their thermk(T) tables
    void setupCuda();

    /// The nonlinear loop of compute() on the device
    double computeCuda(int loops,
                       const BoundaryConditionsWithMesh<RectangularMesh<3>::Boundary,double>& btemperature,
                       const BoundaryConditionsWithMesh<RectangularMesh<3>::Boundary,double>& bheatflux,
                       const BoundaryConditionsWithMesh<RectangularMesh<3>::Boundary,Convection>& bconvection,
                       const BoundaryConditionsWithMesh<RectangularMesh<3>::Boundary,Radiation>& bradiation);
""")

F = "solvers/thermal/static/therm3d.cpp"
edit(F, """#include "therm3d.hpp"
""", """#include "therm3d.hpp"

#include <plaskfem_cuda.hpp>
""")
edit(F, """            if (idx != RectangularMaskedMesh3D::Element::UNKNOWN_ELEMENT_INDEX)
                thickness[idx] = h;
        }
    }
}


void ThermalFem3DSolver::onInvalidate() {
    temperatures.reset();
    fluxes.reset();
    thickness.reset();
}
""", """            if (idx != RectangularMaskedMesh3D::Element::UNKNOWN_ELEMENT_INDEX)
                thickness[idx] = h;
        }
    }

    if (algorithm == ALGORITHM_CUDA) setupCuda();
}


/// Flat description of a RectangularMesh<3> for the device library (SURVEY Appendix B)
static plaskfem::Mesh cudaMesh(const RectangularMesh<3>& mesh) {
    plaskfem::Mesh fm;
    for (int a = 0; a < 3; ++a) {
        fm.axis[a].reserve(mesh.axis[a]->size());
        for (size_t i = 0; i != mesh.axis[a]->size(); ++i) fm.axis[a].push_back(mesh.axis[a]->at(i));
    }
    fm.order = plaskfem::IterationOrder(int(mesh.getIterationOrder()));  // same enumerators, rectilinear3d.hpp:349
    return fm;
}


void ThermalFem3DSolver::setupCuda() {
    try {
        cuda.reset(new plaskfem::Context(0, this->getId()));
        // the line preconditioners want the vertical axis contiguous whatever the iteration order of the mesh
        if (iter_params.preconditioner != IterativeMatrixParams::PRECOND_JAC && this->mesh->axis[2]->size() <= 512)
            cuda->set_layout(PFEM_LAYOUT_VERTICAL_MINOR);
        plaskfem::Mesh fm = cudaMesh(*this->mesh);
        cuda->set_mesh(fm);

        // one table id per (material, layer thickness) pair: thermk(T, thickness) is sampled on the host, the device
        // interpolates (no virtual calls from kernels); elements outside the masked mesh are marked excluded
        const size_t nfull = this->mesh->getElementsCount();
        std::vector<shared_ptr<Material>> materials(nfull);
        std::vector<const Material*> key(nfull, nullptr);
        std::vector<double> thick(nfull, 0.);
        std::vector<uint8_t> included(nfull, 0);
        for (auto elem: this->maskedMesh->elements()) {
            size_t e = this->mesh->element(elem.getIndex0(), elem.getIndex1(), elem.getIndex2()).getIndex();
            materials[e] = this->geometry->getMaterial(elem.getMidpoint());
            key[e] = materials[e].get();
            thick[e] = thickness[elem.getIndex()];
            included[e] = 1;
        }
        std::vector<size_t> reps;
        std::vector<uint32_t> ids = plaskfem::material_ids(key, thick, &reps);
        plaskfem::Tables tables = plaskfem::sample_tables(reps.size(), [&](uint32_t id, double T) {
            if (!materials[reps[id]]) return std::make_pair(0., 0.);
            auto k = materials[reps[id]]->thermk(T, thick[reps[id]]);
            return std::make_pair(k.c00, k.c11);
        });
        if (!this->maskedMesh->full()) ids = plaskfem::MaskedNumbering(fm, included).mark_excluded(ids);
        cuda->set_materials(ids, tables);
        cuda->fill_field(inittemp);
    } catch (const plaskfem::NoDevice& err) {
        throw ComputationError(this->getId(), "algorithm 'cuda' has no CPU fallback: {}", err.what());
    } catch (const plaskfem::BadInput& err) {
        throw BadInput(this->getId(), "{}", err.what());
    }
}


void ThermalFem3DSolver::onInvalidate() {
    temperatures.reset();
    fluxes.reset();
    thickness.reset();
    cuda.reset();
}
""")
edit(F, """    this->writelog(LOG_INFO, "Running thermal calculations");

    int loop = 0;
    size_t size = maskedMesh->size();

    std::unique_ptr<FemMatrix> pA(getMatrix());""", """    this->writelog(LOG_INFO, "Running thermal calculations");

    if (algorithm == ALGORITHM_CUDA) return computeCuda(loops, btemperature, bheatflux, bconvection, bradiation);

    int loop = 0;
    size_t size = maskedMesh->size();

    std::unique_ptr<FemMatrix> pA(getMatrix());""")
edit(F, """void ThermalFem3DSolver::saveHeatFluxes()
{""", """double ThermalFem3DSolver::computeCuda(int loops,
                   const BoundaryConditionsWithMesh<RectangularMesh<3>::Boundary,double>& btemperature,
                   const BoundaryConditionsWithMesh<RectangularMesh<3>::Boundary,double>& bheatflux,
                   const BoundaryConditionsWithMesh<RectangularMesh<3>::Boundary,Convection>& bconvection,
                   const BoundaryConditionsWithMesh<RectangularMesh<3>::Boundary,Radiation>& bradiation)
{
    if (!cuda) setupCuda();
    try {
        plaskfem::Mesh fm = cudaMesh(*this->mesh);
        const bool masked = !this->maskedMesh->full();
        std::vector<uint8_t> included;
        if (masked) {
            included.assign(this->mesh->getElementsCount(), 0);
            for (auto elem: this->maskedMesh->elements())
                included[this->mesh->element(elem.getIndex0(), elem.getIndex1(), elem.getIndex2()).getIndex()] = 1;
        }
        plaskfem::MaskedNumbering numbering(fm, included);   // identity maps on a full mesh
        const size_t nfull = this->mesh->size();
        auto full_node = [&](size_t masked_node) { return masked ? numbering.node_to_full(masked_node) : masked_node; };

        // boundary conditions of the 2nd / 3rd kind and radiation (setMatrix :242-268): per-node getValue() arrays,
        // the element loop of setBoundaries (:140-168) runs inside the library
        plaskfem::NodeConditions<1> hf;
        plaskfem::NodeConditions<2> cv, rd;
        for (auto cond: bheatflux) for (auto r: cond.place) hf.add_node(nfull, full_node(r), {cond.value});
        for (auto cond: bconvection) for (auto r: cond.place) cv.add_node(nfull, full_node(r), {cond.value.coeff, cond.value.ambient});
        for (auto cond: bradiation) for (auto r: cond.place) rd.add_node(nfull, full_node(r), {cond.value.emissivity, cond.value.ambient});
        cuda->set_boundary(hf, cv, rd, /*verbatim=*/true);

        plaskfem::Dirichlet bc;                                   // application order of matrix.hpp:111-118
        for (auto cond: btemperature) for (auto r: cond.place) bc.add_node(full_node(r), cond.value);
        cuda->set_dirichlet(bc);

        auto heats = inHeat(this->maskedMesh->getElementMesh());  // :179
        std::vector<double> heat(this->mesh->getElementsCount(), 0.);
        for (auto elem: this->maskedMesh->elements())
            heat[this->mesh->element(elem.getIndex0(), elem.getIndex1(), elem.getIndex2()).getIndex()] = heats[elem.getIndex()];
        cuda->set_source(heat.data());

        temperatures = temperatures.claim();
        std::vector<double> field(nfull, 0.);
        for (size_t i = 0; i != temperatures.size(); ++i) field[full_node(i)] = temperatures[i];
        cuda->set_field(field.data());                            // warm start, like iterative_matrix.hpp:205-209

        plaskfem::IterParams ip{iter_params.maxit, iter_params.maxerr,
                                plaskfem::IterParams::NoConvergenceBehavior(int(iter_params.no_convergence_behavior))};
        // jac / ljac as named; every other choice (ic is the default) -> the multilevel line preconditioner
        ip.preconditioner = iter_params.preconditioner == IterativeMatrixParams::PRECOND_JAC ? plaskfem::IterParams::PRECOND_JAC :
                            iter_params.preconditioner == IterativeMatrixParams::PRECOND_LJAC ? plaskfem::IterParams::PRECOND_LJAC :
                                                                                                plaskfem::IterParams::PRECOND_MLJ;
        auto result = cuda->solve(true, ip, maxerr, loops,
                                  [this](int level, const std::string& msg) { this->writelog(LogLevel(level), msg); });
        iter_params.converged = ip.converged; iter_params.iters = ip.iters; iter_params.err = ip.err;

        cuda->get_field(field.data());
        for (size_t i = 0; i != temperatures.size(); ++i) temperatures[i] = field[full_node(i)];
        loopno = result.loopno; maxT = result.maxval; toterr = result.toterr;
    } catch (const plaskfem::BadInput& err) {
        throw BadInput(this->getId(), "{}", err.what());
    } catch (const std::runtime_error& err) {
        throw ComputationError(this->getId(), "{}", err.what());
    }

    outTemperature.fireChanged();
    outHeatFlux.fireChanged();

    return toterr;
}

void ThermalFem3DSolver::saveHeatFluxes()
{""")

# ---------------------------------------------------------------- thermal.static Static2D / StaticCyl
F = "solvers/thermal/static/therm2d.hpp"
edit(F, """#include "common.hpp"

namespace plask { namespace thermal { namespace tstatic {
""", """#include "common.hpp"

#include <plaskfem_cuda.hpp>   // host adapter of libplaskfem_cuda.so (algorithm 'cuda'): header-only, plain C ABI underneath

namespace plask { namespace thermal { namespace tstatic {
""")
edit(F, """    DataVector<Vec<2, double>> fluxes;  ///< Computed (only when needed) heat fluxes on our own mesh
""", """    DataVector<Vec<2, double>> fluxes;  ///< Computed (only when needed) heat fluxes on our own mesh

    std::shared_ptr<plaskfem::Context> cuda;  ///< Device context, exists only for algorithm 'cuda'

    /// The 2-D mesh as a brick mesh of one element layer (index bookkeeping of the embedding)
    plaskfem::Embedding2D cudaEmbedding;

    /// Node of the masked mesh -> node of plane 0 of the brick mesh
    std::vector<size_t> cudaNode;

    /// Create the device context: embedded mesh, radial weights, (material, thickness) ids, thermk(T) tables
    void setupCuda();

    /// The nonlinear loop of compute() on the device
    double computeCuda(int loops,
                       const BoundaryConditionsWithMesh<RectangularMesh<2>::Boundary, double>& btemperature,
                       const BoundaryConditionsWithMesh<RectangularMesh<2>::Boundary, double>& bheatflux,
                       const BoundaryConditionsWithMesh<RectangularMesh<2>::Boundary, Convection>& bconvection,
                       const BoundaryConditionsWithMesh<RectangularMesh<2>::Boundary, Radiation>& bradiation);
""")

F = "solvers/thermal/static/therm2d.cpp"
edit(F, """            if (idx != RectangularMaskedMesh2D::Element::UNKNOWN_ELEMENT_INDEX)
                thickness[idx] = h;
        }
    }
}


template<typename Geometry2DType> void ThermalFem2DSolver<Geometry2DType>::onInvalidate() {
    temperatures.reset();
    fluxes.reset();
    thickness.reset();
}
""", """            if (idx != RectangularMaskedMesh2D::Element::UNKNOWN_ELEMENT_INDEX)
                thickness[idx] = h;
        }
    }

    if (this->algorithm == ALGORITHM_CUDA) setupCuda();
}


template<typename Geometry2DType>
void ThermalFem2DSolver<Geometry2DType>::setupCuda() {
    try {
        cuda.reset(new plaskfem::Context(0, this->getId()));
        // The library has no 2-D kernels: a brick mesh with ONE element layer along a dummy axis and z-invariant data gives, on each
        // of its two node planes, 0.5e-6 * d times the 4-node rectangle operator, load and capacity of setMatrix (INTEGRATION.md 9).
        std::vector<double> x, y;
        for (size_t i = 0; i != this->mesh->axis[0]->size(); ++i) x.push_back(this->mesh->axis[0]->at(i));
        for (size_t i = 0; i != this->mesh->axis[1]->size(); ++i) y.push_back(this->mesh->axis[1]->at(i));
        cudaEmbedding = plaskfem::Embedding2D(x, y);
        const plaskfem::Embedding2D& emb = cudaEmbedding;
        cuda->set_mesh(emb.mesh);
        // every element matrix and load of the cylindrical solver carries midpoint.rad_r()
        if (std::is_same<Geometry2DType, Geometry2DCylindrical>::value) cuda->set_axis_weight(1, emb.radial_weights());

        const size_t nfull = emb.elements();
        std::vector<shared_ptr<Material>> materials(nfull);
        std::vector<const Material*> key(nfull, nullptr);
        std::vector<double> thick(nfull, 0.);
        std::vector<uint8_t> included(nfull, 0);
        cudaNode.assign(this->maskedMesh->size(), 0);   // node of the masked mesh -> node of plane 0, whatever the mesh's iteration order
        for (auto elem: this->maskedMesh->elements()) {
            const size_t i0 = elem.getIndex0(), i1 = elem.getIndex1(), e = emb.elem(i0, i1);
            materials[e] = this->geometry->getMaterial(elem.getMidpoint());
            key[e] = materials[e].get();
            thick[e] = thickness[elem.getIndex()];
            included[e] = 1;
            cudaNode[elem.getLoLoIndex()] = emb.node(i0, i1);
            cudaNode[elem.getUpLoIndex()] = emb.node(i0 + 1, i1);
            cudaNode[elem.getLoUpIndex()] = emb.node(i0, i1 + 1);
            cudaNode[elem.getUpUpIndex()] = emb.node(i0 + 1, i1 + 1);
        }
        std::vector<size_t> reps;
        std::vector<uint32_t> ids = plaskfem::material_ids(key, thick, &reps);
        plaskfem::Tables tables = plaskfem::sample_tables(reps.size(), [&](uint32_t id, double T) {
            if (!materials[reps[id]]) return std::make_pair(1., 1.);
            auto k = materials[reps[id]]->thermk(T, thick[reps[id]]);
            return std::make_pair(k.c00, k.c11);
        });
        if (!this->maskedMesh->full()) ids = plaskfem::MaskedNumbering(emb.mesh, included).mark_excluded(ids);
        cuda->set_materials(ids, tables);
        cuda->fill_field(inittemp);
    } catch (const plaskfem::NoDevice& err) {
        throw ComputationError(this->getId(), "algorithm 'cuda' has no CPU fallback: {}", err.what());
    } catch (const plaskfem::BadInput& err) {
        throw BadInput(this->getId(), "{}", err.what());
    }
}


template<typename Geometry2DType> void ThermalFem2DSolver<Geometry2DType>::onInvalidate() {
    temperatures.reset();
    fluxes.reset();
    thickness.reset();
    cuda.reset();
    cudaNode.clear();
}
""")
edit(F, """    this->writelog(LOG_INFO, "Running thermal calculations");

    int loop = 0;
    size_t size = this->maskedMesh->size();

    std::unique_ptr<FemMatrix> pA(this->getMatrix());
    FemMatrix& A = *pA.get();

    double err;
    toterr = 0.;
""", """    this->writelog(LOG_INFO, "Running thermal calculations");

    if (this->algorithm == ALGORITHM_CUDA) return computeCuda(loops, btemperature, bheatflux, bconvection, bradiation);

    int loop = 0;
    size_t size = this->maskedMesh->size();

    std::unique_ptr<FemMatrix> pA(this->getMatrix());
    FemMatrix& A = *pA.get();

    double err;
    toterr = 0.;
""")
edit(F, """template<typename Geometry2DType>
void ThermalFem2DSolver<Geometry2DType>::saveHeatFluxes()
{""", """template<typename Geometry2DType>
double ThermalFem2DSolver<Geometry2DType>::computeCuda(int loops,
                   const BoundaryConditionsWithMesh<RectangularMesh<2>::Boundary,double>& btemperature,
                   const BoundaryConditionsWithMesh<RectangularMesh<2>::Boundary,double>& bheatflux,
                   const BoundaryConditionsWithMesh<RectangularMesh<2>::Boundary,Convection>& bconvection,
                   const BoundaryConditionsWithMesh<RectangularMesh<2>::Boundary,Radiation>& bradiation)
{
    if (!cuda) setupCuda();
    try {
        const plaskfem::Embedding2D& emb = cudaEmbedding;

        plaskfem::Dirichlet bc;                                   // application order of matrix.hpp:111-118, on both node planes
        for (auto cond: btemperature) for (auto r: cond.place) emb.add_dirichlet(bc, cudaNode[r], cond.value);
        cuda->set_dirichlet(bc);

        auto heats = inHeat(this->maskedMesh->getElementMesh());
        std::vector<double> heat(emb.elements(), 0.);
        for (auto elem: this->maskedMesh->elements()) heat[emb.elem(elem.getIndex0(), elem.getIndex1())] = heats[elem.getIndex()];
        cuda->set_source(heat.data());

        temperatures = temperatures.claim();
        std::vector<double> plane0(emb.plane(), 0.);
        for (size_t i = 0; i != temperatures.size(); ++i) plane0[cudaNode[i]] = temperatures[i];
        std::vector<double> field = emb.lift(plane0.data());
        cuda->set_field(field.data());                            // warm start, like iterative_matrix.hpp:205-209

        // Edge conditions of the 2nd / 3rd kind and radiation (setBoundaries + the lambdas of setMatrix): per-node getValue() arrays
        // on plane 0, flattened by the library in its 2-D mode (pfem_boundary::mode2d: 1 Cartesian, 2 cylindrical)
        const size_t nbrick = 2 * emb.plane();
        plaskfem::NodeConditions<1> hf;
        plaskfem::NodeConditions<2> cv, rd;
        for (auto cond: bheatflux) for (auto r: cond.place) hf.add_node(nbrick, cudaNode[r], {cond.value});
        for (auto cond: bconvection) for (auto r: cond.place) cv.add_node(nbrick, cudaNode[r], {cond.value.coeff, cond.value.ambient});
        for (auto cond: bradiation) for (auto r: cond.place) rd.add_node(nbrick, cudaNode[r], {cond.value.emissivity, cond.value.ambient});
        cuda->set_boundary(hf, cv, rd, /*verbatim=*/true,
                           plaskfem::Embedding2D::boundary_mode(std::is_same<Geometry2DType, Geometry2DCylindrical>::value));

        plaskfem::IterParams ip{this->iter_params.maxit, this->iter_params.maxerr,
                                plaskfem::IterParams::NoConvergenceBehavior(int(this->iter_params.no_convergence_behavior))};
        ip.preconditioner = this->iter_params.preconditioner == IterativeMatrixParams::PRECOND_JAC ? plaskfem::IterParams::PRECOND_JAC :
                            this->iter_params.preconditioner == IterativeMatrixParams::PRECOND_LJAC ? plaskfem::IterParams::PRECOND_LJAC :
                                                                                                      plaskfem::IterParams::PRECOND_MLJ;
        auto result = cuda->solve(true, ip, maxerr, loops,
                                  [this](int level, const std::string& msg) { this->writelog(LogLevel(level), msg); });
        this->iter_params.converged = ip.converged; this->iter_params.iters = ip.iters; this->iter_params.err = ip.err;

        cuda->get_field(field.data());
        for (size_t i = 0; i != temperatures.size(); ++i) temperatures[i] = field[cudaNode[i]];
        loopno = result.loopno; maxT = result.maxval; toterr = result.toterr;
    } catch (const plaskfem::BadInput& err) {
        throw BadInput(this->getId(), "{}", err.what());
    } catch (const std::runtime_error& err) {
        throw ComputationError(this->getId(), "{}", err.what());
    }

    outTemperature.fireChanged();
    outHeatFlux.fireChanged();

    return toterr;
}


template<typename Geometry2DType>
void ThermalFem2DSolver<Geometry2DType>::saveHeatFluxes()
{""")

# ---------------------------------------------------------------- thermal.dynamic Dynamic2D / DynamicCyl
F = "solvers/thermal/dynamic/femT2d.hpp"
edit(F, """#include <plask/common/fem.hpp>

namespace plask { namespace thermal { namespace dynamic {

/**
 * Solver performing calculations in 2D Cartesian or Cylindrical space using finite element method
 */
template<typename Geometry2DType>""", """#include <plask/common/fem.hpp>

#include <plaskfem_cuda.hpp>   // host adapter of libplaskfem_cuda.so (algorithm 'cuda'): header-only, plain C ABI underneath

namespace plask { namespace thermal { namespace dynamic {

/**
 * Solver performing calculations in 2D Cartesian or Cylindrical space using finite element method
 */
template<typename Geometry2DType>""")
edit(F, """    DataVector<Vec<2,double>> fluxes;           ///< Computed (only when needed) heat fluxes on our own mesh
""", """    DataVector<Vec<2,double>> fluxes;           ///< Computed (only when needed) heat fluxes on our own mesh

    std::shared_ptr<plaskfem::Context> cuda;  ///< Device context, exists only for algorithm 'cuda'

    /// The 2-D mesh as a brick mesh of one element layer (index bookkeeping of the embedding)
    plaskfem::Embedding2D cudaEmbedding;

    /// Node of the masked mesh -> node of plane 0 of the brick mesh
    std::vector<size_t> cudaNode;

    /// Create the device context: embedded mesh, radial weights, (material, thickness) ids, thermk(T) and cp(T)*dens(T) tables
    void setupCuda();

    /// The time loop of compute() on the device (pfem_solve_dynamic)
    double computeCuda(double time, const BoundaryConditionsWithMesh<RectangularMesh<2>::Boundary,double>& btemperature);
""")

F = "solvers/thermal/dynamic/femT2d.cpp"
edit(F, """            if (idx != RectangularMaskedMesh2D::Element::UNKNOWN_ELEMENT_INDEX)
                thickness[idx] = h;
        }
    }
}


template<typename Geometry2DType> void DynamicThermalFem2DSolver<Geometry2DType>::onInvalidate() {
    temperatures.reset();
    fluxes.reset();
}
""", """            if (idx != RectangularMaskedMesh2D::Element::UNKNOWN_ELEMENT_INDEX)
                thickness[idx] = h;
        }
    }

    if (this->algorithm == ALGORITHM_CUDA) setupCuda();
}


template<typename Geometry2DType>
void DynamicThermalFem2DSolver<Geometry2DType>::setupCuda() {
    try {
        cuda.reset(new plaskfem::Context(0, this->getId()));
        // The library has no 2-D kernels: a brick mesh with ONE element layer along a dummy axis and z-invariant data gives, on each
        // of its two node planes, 0.5e-6 * d times the 4-node rectangle operator, load and capacity of setMatrix (INTEGRATION.md 9).
        std::vector<double> x, y;
        for (size_t i = 0; i != this->mesh->axis[0]->size(); ++i) x.push_back(this->mesh->axis[0]->at(i));
        for (size_t i = 0; i != this->mesh->axis[1]->size(); ++i) y.push_back(this->mesh->axis[1]->at(i));
        cudaEmbedding = plaskfem::Embedding2D(x, y);
        const plaskfem::Embedding2D& emb = cudaEmbedding;
        cuda->set_mesh(emb.mesh);
        // every element matrix and load of the cylindrical solver carries midpoint.rad_r()
        if (std::is_same<Geometry2DType, Geometry2DCylindrical>::value) cuda->set_axis_weight(1, emb.radial_weights());

        const size_t nfull = emb.elements();
        std::vector<shared_ptr<Material>> materials(nfull);
        std::vector<const Material*> key(nfull, nullptr);
        std::vector<double> thick(nfull, 0.);
        std::vector<uint8_t> included(nfull, 0);
        cudaNode.assign(this->maskedMesh->size(), 0);   // node of the masked mesh -> node of plane 0, whatever the mesh's iteration order
        for (auto elem: this->maskedMesh->elements()) {
            const size_t i0 = elem.getIndex0(), i1 = elem.getIndex1(), e = emb.elem(i0, i1);
            materials[e] = this->geometry->getMaterial(elem.getMidpoint());
            key[e] = materials[e].get();
            thick[e] = thickness[elem.getIndex()];
            included[e] = 1;
            cudaNode[elem.getLoLoIndex()] = emb.node(i0, i1);
            cudaNode[elem.getUpLoIndex()] = emb.node(i0 + 1, i1);
            cudaNode[elem.getLoUpIndex()] = emb.node(i0, i1 + 1);
            cudaNode[elem.getUpUpIndex()] = emb.node(i0 + 1, i1 + 1);
        }
        std::vector<size_t> reps;
        std::vector<uint32_t> ids = plaskfem::material_ids(key, thick, &reps);
        plaskfem::Tables tables = plaskfem::sample_tables(reps.size(), [&](uint32_t id, double T) {
            if (!materials[reps[id]]) return std::make_pair(1., 1.);
            auto k = materials[reps[id]]->thermk(T, thick[reps[id]]);
            return std::make_pair(k.c00, k.c11);
        });
        if (!this->maskedMesh->full()) ids = plaskfem::MaskedNumbering(emb.mesh, included).mark_excluded(ids);
        cuda->set_materials(ids, tables);
        // one more table per id: cp(T) * dens(T) of the element capacity (setMatrix: cp * dens * 0.25e-12 * w * h / timestep / 1e-9)
        std::vector<double> cpdens(size_t(tables.nmat) * tables.nT, 1.);
        for (uint32_t id = 0; id != tables.nmat; ++id)
            if (materials[reps[id]])
                for (uint32_t i = 0; i != tables.nT; ++i) {
                    double T = tables.T0 + i * tables.dT;
                    cpdens[size_t(id) * tables.nT + i] = materials[reps[id]]->cp(T) * materials[reps[id]]->dens(T);
                }
        cuda->set_capacity(tables, cpdens);
        cuda->fill_field(inittemp);
    } catch (const plaskfem::NoDevice& err) {
        throw ComputationError(this->getId(), "algorithm 'cuda' has no CPU fallback: {}", err.what());
    } catch (const plaskfem::BadInput& err) {
        throw BadInput(this->getId(), "{}", err.what());
    }
}


template<typename Geometry2DType>
double DynamicThermalFem2DSolver<Geometry2DType>::computeCuda(double time,
                   const BoundaryConditionsWithMesh<RectangularMesh<2>::Boundary,double>& btemperature)
{
    if (!cuda) setupCuda();
    try {
        const plaskfem::Embedding2D& emb = cudaEmbedding;

        plaskfem::Dirichlet bc;                                   // application order of matrix.hpp:111-118, on both node planes
        for (auto cond: btemperature) for (auto r: cond.place) emb.add_dirichlet(bc, cudaNode[r], cond.value);
        cuda->set_dirichlet(bc);

        auto heats = inHeat(this->maskedMesh->getElementMesh());
        std::vector<double> heat(emb.elements(), 0.);
        for (auto elem: this->maskedMesh->elements()) heat[emb.elem(elem.getIndex0(), elem.getIndex1())] = heats[elem.getIndex()];
        cuda->set_source(heat.data());

        temperatures = temperatures.claim();
        std::vector<double> plane0(emb.plane(), 0.);
        for (size_t i = 0; i != temperatures.size(); ++i) plane0[cudaNode[i]] = temperatures[i];
        std::vector<double> field = emb.lift(plane0.data());
        cuda->set_field(field.data());                            // warm start, like iterative_matrix.hpp:205-209

        plaskfem::IterParams ip{this->iter_params.maxit, this->iter_params.maxerr,
                                plaskfem::IterParams::NoConvergenceBehavior(int(this->iter_params.no_convergence_behavior))};
        // jac as named, every other choice -> line-Jacobi (the multilevel preconditioner does not take the capacity diagonal yet)
        ip.preconditioner = this->iter_params.preconditioner == IterativeMatrixParams::PRECOND_JAC ? plaskfem::IterParams::PRECOND_JAC :
                                                                                                      plaskfem::IterParams::PRECOND_LJAC;
        // the corrected theta scheme (plaskfem_cuda.h): Dirichlet rows keep their values (setMatrix eliminates them in A and F only)
        auto result = cuda->solve_dynamic(ip, time, timestep, methodparam, lumping, int(rebuildfreq), int(logfreq), elapstime,
                                          [this](int level, const std::string& msg) { this->writelog(LogLevel(level), msg); });
        this->iter_params.converged = ip.converged; this->iter_params.iters = ip.iters; this->iter_params.err = ip.err;

        cuda->get_field(field.data());
        for (size_t i = 0; i != temperatures.size(); ++i) temperatures[i] = field[cudaNode[i]];
        maxT = result.maxT;
    } catch (const plaskfem::BadInput& err) {
        throw BadInput(this->getId(), "{}", err.what());
    } catch (const std::runtime_error& err) {
        throw ComputationError(this->getId(), "{}", err.what());
    }

    outTemperature.fireChanged();
    outHeatFlux.fireChanged();

    return 0.;
}


template<typename Geometry2DType> void DynamicThermalFem2DSolver<Geometry2DType>::onInvalidate() {
    temperatures.reset();
    fluxes.reset();
    cuda.reset();
    cudaNode.clear();
}
""")
edit(F, """    auto btemperature = temperature_boundary(this->maskedMesh, this->geometry);

    size_t size = this->maskedMesh->size();

    std::unique_ptr<FemMatrix> pA(this->getMatrix());
    FemMatrix& A = *pA.get();
    std::unique_ptr<FemMatrix> pB(this->getMatrix());
    FemMatrix& B = *pB.get();
""", """    auto btemperature = temperature_boundary(this->maskedMesh, this->geometry);

    if (this->algorithm == ALGORITHM_CUDA) return computeCuda(time, btemperature);

    size_t size = this->maskedMesh->size();

    std::unique_ptr<FemMatrix> pA(this->getMatrix());
    FemMatrix& A = *pA.get();
    std::unique_ptr<FemMatrix> pB(this->getMatrix());
    FemMatrix& B = *pB.get();
""")

# ---------------------------------------------------------------- electrical.shockley Shockley3D
F = "solvers/electrical/shockley/electr3d.hpp"
edit(F, """#include "common.hpp"

namespace plask { namespace electrical { namespace shockley {

/**
 * Solver performing calculations in 2D Cartesian or Cylindrical space using finite element method
 */
struct PLASK_SOLVER_API ElectricalFem3DSolver :""", """#include "common.hpp"

namespace plaskfem { class Context; }   // plaskfem_cuda.hpp: host adapter of libplaskfem_cuda.so (algorithm 'cuda')

namespace plask { namespace electrical { namespace shockley {

/**
 * Solver performing calculations in 2D Cartesian or Cylindrical space using finite element method
 */
struct PLASK_SOLVER_API ElectricalFem3DSolver :""")
edit(F, """    virtual Tensor2<double> activeCond(size_t n, double U, double jy, double T) = 0;
""", """    virtual Tensor2<double> activeCond(size_t n, double U, double jy, double T) = 0;

    /** Parameters of the Shockley law of junction \\p n at temperature \\p T, for solvers whose activeCond is
     *  10 |jy| beta h / ln(1e7 |jy| / js + 1) (BetaSolver and its Python subclass).  Algorithm 'cuda' evaluates that law
     *  on the device and therefore needs the parameters instead of the callback.
     *  \\return false if this solver has another junction model
     */
    virtual bool shockleyParameters(size_t PLASK_UNUSED(n), double PLASK_UNUSED(T), double& PLASK_UNUSED(beta),
                                    double& PLASK_UNUSED(js)) const { return false; }

    std::shared_ptr<plaskfem::Context> cuda;    ///< Device context, exists only for algorithm 'cuda'

    /// Create the device context: mesh and the cond(T) tables of the materials
    void setupCuda();

    /// The nonlinear loop of compute() on the device
    double computeCuda(unsigned loops, const BoundaryConditionsWithMesh<RectangularMesh<3>::Boundary, double>& bvoltage);
""")

F = "solvers/electrical/shockley/beta.hpp"
edit(F, """        return Tensor2<double>(0., 10. * jy * this->active[n].height * getBeta(n) / log(1e7 * jy / getJs(n) + 1.));
    }
""", """        return Tensor2<double>(0., 10. * jy * this->active[n].height * getBeta(n) / log(1e7 * jy / getJs(n) + 1.));
    }

    bool shockleyParameters(size_t n, double PLASK_UNUSED(T), double& beta, double& js) const override {
        beta = getBeta(n);
        js = getJs(n);
        return true;
    }
""")

F = "solvers/electrical/shockley/python/electr_python.cpp"
edit(F, """        jy = abs(jy);
        return Tensor2<double>(0., 10. * jy * beta * this->active[n].height / log(1e7 * jy / js + 1.));
    }
};
""", """        jy = abs(jy);
        return Tensor2<double>(0., 10. * jy * beta * this->active[n].height / log(1e7 * jy / js + 1.));
    }

    bool shockleyParameters(size_t n, double T, double& beta, double& js) const override {
        OmpLockGuard lock(python_omp_lock);
        beta = (n < beta_function.size() && !beta_function[n].is_none()) ? py::extract<double>(beta_function[n](T))
                                                                         : BetaSolver<GeometryT>::getBeta(n);
        js = (n < js_function.size() && !js_function[n].is_none()) ? py::extract<double>(js_function[n](T))
                                                                   : BetaSolver<GeometryT>::getJs(n);
        return true;
    }
};
""")

F = "solvers/electrical/shockley/electr3d.cpp"
edit(F, """#include "electr3d.hpp"
""", """#include "electr3d.hpp"

#include <plaskfem_cuda.hpp>
""")
edit(F, """    potential.reset(maskedMesh->size(), 0.);
    current.reset(maskedMesh->getElementsCount(), vec(0., 0., 0.));
    conds.reset(maskedMesh->getElementsCount());
}

void ElectricalFem3DSolver::onInvalidate() {
    conds.reset();
    potential.reset();
    current.reset();
    heat.reset();
    junction_conductivity.reset(1, default_junction_conductivity);
}
""", """    potential.reset(maskedMesh->size(), 0.);
    current.reset(maskedMesh->getElementsCount(), vec(0., 0., 0.));
    conds.reset(maskedMesh->getElementsCount());
    if (algorithm == ALGORITHM_CUDA) setupCuda();
}

void ElectricalFem3DSolver::onInvalidate() {
    conds.reset();
    potential.reset();
    current.reset();
    heat.reset();
    junction_conductivity.reset(1, default_junction_conductivity);
    cuda.reset();
}

/// Flat description of a RectangularMesh<3> for the device library (SURVEY Appendix B)
static plaskfem::Mesh cudaMesh(const RectangularMesh<3>& mesh) {
    plaskfem::Mesh fm;
    for (int a = 0; a < 3; ++a) {
        fm.axis[a].reserve(mesh.axis[a]->size());
        for (size_t i = 0; i != mesh.axis[a]->size(); ++i) fm.axis[a].push_back(mesh.axis[a]->at(i));
    }
    fm.order = plaskfem::IterationOrder(int(mesh.getIterationOrder()));  // same enumerators, rectilinear3d.hpp:349
    return fm;
}

void ElectricalFem3DSolver::setupCuda() {
    try {
        cuda.reset(new plaskfem::Context(0, this->getId()));
        if (iter_params.preconditioner != IterativeMatrixParams::PRECOND_JAC && this->mesh->axis[2]->size() <= 512)
            cuda->set_layout(PFEM_LAYOUT_VERTICAL_MINOR);
        plaskfem::Mesh fm = cudaMesh(*this->mesh);
        cuda->set_mesh(fm);
        // cond(T) of every distinct material, sampled on the host (loadConductivity :221)
        const size_t nfull = this->mesh->getElementsCount();
        std::vector<shared_ptr<Material>> materials(nfull);
        std::vector<const Material*> key(nfull, nullptr);
        std::vector<uint8_t> included(nfull, 0);
        for (auto elem : this->maskedMesh->elements()) {
            size_t e = this->mesh->element(elem.getIndex0(), elem.getIndex1(), elem.getIndex2()).getIndex();
            materials[e] = this->geometry->getMaterial(elem.getMidpoint());
            key[e] = materials[e].get();
            included[e] = 1;
        }
        std::vector<size_t> reps;
        std::vector<uint32_t> ids = plaskfem::material_ids(key, std::vector<double>(), &reps);
        plaskfem::Tables tables = plaskfem::sample_tables(reps.size(), [&](uint32_t id, double T) {
            if (!materials[reps[id]]) return std::make_pair(0., 0.);
            auto s = materials[reps[id]]->cond(T);
            return std::make_pair(s.c00, s.c11);
        });
        if (!this->maskedMesh->full()) ids = plaskfem::MaskedNumbering(fm, included).mark_excluded(ids);
        cuda->set_materials(ids, tables);
        cuda->fill_field(0.);
    } catch (const plaskfem::NoDevice& err) {
        throw ComputationError(this->getId(), "algorithm 'cuda' has no CPU fallback: {}", err.what());
    } catch (const plaskfem::BadInput& err) {
        throw BadInput(this->getId(), "{}", err.what());
    }
}
""")
edit(F, """    this->writelog(LOG_INFO, "Running electrical calculations");

    unsigned loop = 0;
    double err = 0.;
    toterr = 0.;

    std::unique_ptr<FemMatrix> pA(this->getMatrix());""", """    this->writelog(LOG_INFO, "Running electrical calculations");

    if (algorithm == ALGORITHM_CUDA) return computeCuda(loops, bvoltage);

    unsigned loop = 0;
    double err = 0.;
    toterr = 0.;

    std::unique_ptr<FemMatrix> pA(this->getMatrix());""")
edit(F, """void ElectricalFem3DSolver::saveHeatDensity() {""", """double ElectricalFem3DSolver::computeCuda(unsigned loops,
                                          const BoundaryConditionsWithMesh<RectangularMesh<3>::Boundary, double>& bvoltage) {
    if (!cuda) setupCuda();
    try {
        plaskfem::Mesh fm = cudaMesh(*this->mesh);
        const bool masked = !this->maskedMesh->full();
        const size_t nfull = this->mesh->size(), efull = this->mesh->getElementsCount();
        std::vector<uint8_t> included;
        if (masked) {
            included.assign(efull, 0);
            for (auto elem : this->maskedMesh->elements())
                included[this->mesh->element(elem.getIndex0(), elem.getIndex1(), elem.getIndex2()).getIndex()] = 1;
        }
        plaskfem::MaskedNumbering numbering(fm, included);
        auto full_node = [&](size_t i) { return masked ? numbering.node_to_full(i) : i; };
        auto full_elem = [&](const RectangularMaskedMesh<3>::Element& el) {
            return this->mesh->element(el.getIndex0(), el.getIndex1(), el.getIndex2()).getIndex();
        };

        // what loadConductivity (:203-225) and saveHeatDensity (:472) read from the geometry, flattened
        std::vector<uint32_t> elem_junc(efull, 0);
        std::vector<uint8_t> elem_role(efull, 0), noheat(efull, 0);
        std::vector<double> Te(efull, 300.);
        auto temperature = inTemperature(this->maskedMesh->getElementMesh());
        for (auto el : this->maskedMesh->elements()) {
            size_t e = full_elem(el);
            auto mid = el.getMidpoint();
            auto roles = this->geometry->getRolesAt(mid);
            elem_junc[e] = uint32_t(isActive(mid));
            elem_role[e] = roles.find("p-contact") != roles.end() ? 1 : roles.find("n-contact") != roles.end() ? 2 : 0;
            noheat[e] = this->geometry->getMaterial(mid)->kind() == Material::EMPTY || roles.find("noheat") != roles.end();
            Te[e] = temperature[el.getIndex()];
        }
        cuda->set_elem_temperature(Te.data());
        cuda->set_noheat(noheat);

        // junctions: Active maps 1:1 onto pfem_junction; beta(T), js(T) per junction-table entry at the temperature of the
        // mid-plane element of the column (setMatrix :261-262)
        std::vector<pfem_junction> junctions(active.size());
        const size_t ncol = junction_conductivity.size();
        std::vector<double> beta_col(ncol, 1.), js_col(ncol, 1.);
        for (size_t n = 0; n != active.size(); ++n) {
            const Active& act = active[n];
            junctions[n] = pfem_junction{act.bottom, act.top, act.left, act.right, act.back, act.front, act.ld, act.offset, act.height};
            const size_t mid = (act.bottom + act.top) / 2;
            for (size_t t = act.left; t != act.right; ++t)
                for (size_t l = act.back; l != act.front; ++l) {
                    const size_t tidx = this->maskedMesh->element(l, t, mid).getIndex();
                    const double T = tidx != RectangularMaskedMesh3D::Element::UNKNOWN_ELEMENT_INDEX ? temperature[tidx] : 300.;
                    const size_t col = act.offset + act.ld * t + l;
                    if (!shockleyParameters(n, T, beta_col[col], js_col[col]))
                        throw BadInput(this->getId(), "algorithm 'cuda' needs a Shockley junction (beta, js); this solver has a custom junction model");
                }
        }
        std::vector<double> jcond(2 * ncol);
        for (size_t i = 0; i != ncol; ++i) { jcond[2 * i] = junction_conductivity[i].c00; jcond[2 * i + 1] = junction_conductivity[i].c11; }
        cuda->set_junctions(junctions, elem_junc, elem_role, pcond, ncond, jcond, beta_col, js_col, convergence == CONVERGENCE_STABLE);

        plaskfem::Dirichlet bc;                                   // application order of matrix.hpp:111-118
        for (auto cond : bvoltage) for (auto r : cond.place) bc.add_node(full_node(r), cond.value);
        cuda->set_dirichlet(bc);
        cuda->set_source(nullptr);                                // zero load vector (:278)

        potential = potential.claim();
        std::vector<double> field(nfull, 0.);
        for (size_t i = 0; i != potential.size(); ++i) field[full_node(i)] = potential[i];
        cuda->set_field(field.data());

        plaskfem::IterParams ip{iter_params.maxit, iter_params.maxerr,
                                plaskfem::IterParams::NoConvergenceBehavior(int(iter_params.no_convergence_behavior))};
        ip.preconditioner = iter_params.preconditioner == IterativeMatrixParams::PRECOND_JAC ? plaskfem::IterParams::PRECOND_JAC :
                            iter_params.preconditioner == IterativeMatrixParams::PRECOND_LJAC ? plaskfem::IterParams::PRECOND_LJAC :
                                                                                                plaskfem::IterParams::PRECOND_MLJ;
        auto result = cuda->solve(false, ip, maxerr, int(loops),
                                  [this](int level, const std::string& msg) { this->writelog(LogLevel(level), msg); });
        iter_params.converged = ip.converged; iter_params.iters = ip.iters; iter_params.err = ip.err;

        cuda->get_field(field.data());
        for (size_t i = 0; i != potential.size(); ++i) potential[i] = field[full_node(i)];
        std::vector<double> cur(3 * efull), cnd(2 * efull);
        cuda->get_elem(PFEM_ELEM_CURRENT, cur.data());
        cuda->get_elem(PFEM_ELEM_COND, cnd.data());
        for (auto el : this->maskedMesh->elements()) {
            size_t e = full_elem(el), i = el.getIndex();
            current[i] = vec(cur[3 * e], cur[3 * e + 1], cur[3 * e + 2]);
            conds[i] = Tensor2<double>(cnd[2 * e], cnd[2 * e + 1]);
        }
        cuda->get_junction_cond(jcond.data());                    // saveConductivity (:227-237) happened on the device
        for (size_t i = 0; i != ncol; ++i) junction_conductivity[i] = Tensor2<double>(jcond[2 * i], jcond[2 * i + 1]);
        heat.reset();
        maxcur = vec(result.maxcur[0], result.maxcur[1], result.maxcur[2]);
        loopno = result.loopno;
        toterr = result.toterr;
    } catch (const plaskfem::BadInput& err) {
        throw BadInput(this->getId(), "{}", err.what());
    } catch (const std::runtime_error& err) {
        throw ComputationError(this->getId(), "{}", err.what());
    }

    outVoltage.fireChanged();
    outCurrentDensity.fireChanged();
    outHeat.fireChanged();

    return toterr;
}

void ElectricalFem3DSolver::saveHeatDensity() {""")

# ---------------------------------------------------------------- electrical.shockley Shockley2D / ShockleyCyl
# BetaSolver<GeometryT> (beta.hpp) derives from ElectricalFem2DSolver<GeometryT> for the 2-D geometries: the shockleyParameters
# virtual it overrides must exist in that base too.
F = "solvers/electrical/shockley/electr2d.hpp"
edit(F, """#include "common.hpp"
""", """#include "common.hpp"

#include <plaskfem_cuda.hpp>   // host adapter of libplaskfem_cuda.so (algorithm 'cuda'): header-only, plain C ABI underneath
""")
edit(F, """    virtual Tensor2<double> activeCond(size_t n, double U, double jy, double T) = 0;
""", """    virtual Tensor2<double> activeCond(size_t n, double U, double jy, double T) = 0;

    /** Parameters of the Shockley law of junction \\p n at temperature \\p T, for solvers whose activeCond is
     *  10 |jy| beta h / ln(1e7 |jy| / js + 1) (BetaSolver and its Python subclass).  Algorithm 'cuda' evaluates that law
     *  on the device and therefore needs the parameters instead of the callback.
     *  \\return false if this solver has another junction model
     */
    virtual bool shockleyParameters(size_t PLASK_UNUSED(n), double PLASK_UNUSED(T), double& PLASK_UNUSED(beta),
                                    double& PLASK_UNUSED(js)) const { return false; }

    std::shared_ptr<plaskfem::Context> cuda;  ///< Device context, exists only for algorithm 'cuda'

    /// The 2-D mesh as a brick mesh of one element layer (index bookkeeping of the embedding)
    plaskfem::Embedding2D cudaEmbedding;

    /// Node of the masked mesh -> node of plane 0 of the brick mesh
    std::vector<size_t> cudaNode;

    /// Create the device context: embedded mesh, radial weights, the cond(T) tables of the materials
    void setupCuda();

    /// The nonlinear loop of compute() on the device
    double computeCuda(unsigned loops, const BoundaryConditionsWithMesh<RectangularMesh<2>::Boundary, double>& bvoltage);
""")

F = "solvers/electrical/shockley/electr2d.cpp"
edit(F, """    potentials.reset(this->maskedMesh->size(), 0.);
    currents.reset(this->maskedMesh->getElementsCount(), vec(0., 0.));
    conds.reset(this->maskedMesh->getElementsCount());
}

template <typename Geometry2DType> void ElectricalFem2DSolver<Geometry2DType>::onInvalidate() {
    conds.reset();
    potentials.reset();
    currents.reset();
    heats.reset();
    junction_conductivity.reset(1, default_junction_conductivity);
}
""", """    potentials.reset(this->maskedMesh->size(), 0.);
    currents.reset(this->maskedMesh->getElementsCount(), vec(0., 0.));
    conds.reset(this->maskedMesh->getElementsCount());
    if (this->algorithm == ALGORITHM_CUDA) setupCuda();
}

template <typename Geometry2DType> void ElectricalFem2DSolver<Geometry2DType>::onInvalidate() {
    conds.reset();
    potentials.reset();
    currents.reset();
    heats.reset();
    junction_conductivity.reset(1, default_junction_conductivity);
    cuda.reset();
    cudaNode.clear();
}

template <typename Geometry2DType> void ElectricalFem2DSolver<Geometry2DType>::setupCuda() {
    try {
        cuda.reset(new plaskfem::Context(0, this->getId()));
        // one element layer along a dummy axis: on each node plane the brick operator is 0.5e-6 * d times the rectangle operator of
        // setMatrix, currents and Joule heat are element gradients and come out identical (INTEGRATION.md 9)
        std::vector<double> x, y;
        for (size_t i = 0; i != this->mesh->axis[0]->size(); ++i) x.push_back(this->mesh->axis[0]->at(i));
        for (size_t i = 0; i != this->mesh->axis[1]->size(); ++i) y.push_back(this->mesh->axis[1]->at(i));
        cudaEmbedding = plaskfem::Embedding2D(x, y);
        const plaskfem::Embedding2D& emb = cudaEmbedding;
        cuda->set_mesh(emb.mesh);
        // setLocalMatrix of the cylindrical solver multiplies every element matrix by midpoint.rad_r()
        if (std::is_same<Geometry2DType, Geometry2DCylindrical>::value) cuda->set_axis_weight(1, emb.radial_weights());

        // cond(T) of every distinct material, sampled on the host (loadConductivities)
        const size_t nfull = emb.elements();
        std::vector<shared_ptr<Material>> materials(nfull);
        std::vector<const Material*> key(nfull, nullptr);
        std::vector<uint8_t> included(nfull, 0);
        cudaNode.assign(this->maskedMesh->size(), 0);
        for (auto elem : this->maskedMesh->elements()) {
            const size_t i0 = elem.getIndex0(), i1 = elem.getIndex1(), e = emb.elem(i0, i1);
            materials[e] = this->geometry->getMaterial(elem.getMidpoint());
            key[e] = materials[e].get();
            included[e] = 1;
            cudaNode[elem.getLoLoIndex()] = emb.node(i0, i1);
            cudaNode[elem.getUpLoIndex()] = emb.node(i0 + 1, i1);
            cudaNode[elem.getLoUpIndex()] = emb.node(i0, i1 + 1);
            cudaNode[elem.getUpUpIndex()] = emb.node(i0 + 1, i1 + 1);
        }
        std::vector<size_t> reps;
        std::vector<uint32_t> ids = plaskfem::material_ids(key, std::vector<double>(), &reps);
        plaskfem::Tables tables = plaskfem::sample_tables(reps.size(), [&](uint32_t id, double T) {
            if (!materials[reps[id]]) return std::make_pair(0., 0.);
            auto s = materials[reps[id]]->cond(T);
            return std::make_pair(s.c00, s.c11);
        });
        if (!this->maskedMesh->full()) ids = plaskfem::MaskedNumbering(emb.mesh, included).mark_excluded(ids);
        cuda->set_materials(ids, tables);
        cuda->fill_field(0.);
    } catch (const plaskfem::NoDevice& err) {
        throw ComputationError(this->getId(), "algorithm 'cuda' has no CPU fallback: {}", err.what());
    } catch (const plaskfem::BadInput& err) {
        throw BadInput(this->getId(), "{}", err.what());
    }
}
""")
edit(F, """    this->writelog(LOG_INFO, "Running electrical calculations");

    unsigned loop = 0;

    std::unique_ptr<FemMatrix> pA(this->getMatrix());
    FemMatrix& A = *pA.get();

    double err = 0.;
    toterr = 0.;
""", """    this->writelog(LOG_INFO, "Running electrical calculations");

    if (this->algorithm == ALGORITHM_CUDA) return computeCuda(loops, vconst);

    unsigned loop = 0;

    std::unique_ptr<FemMatrix> pA(this->getMatrix());
    FemMatrix& A = *pA.get();

    double err = 0.;
    toterr = 0.;
""")
edit(F, """template <typename Geometry2DType> void ElectricalFem2DSolver<Geometry2DType>::saveHeatDensities() {""",
     """template <typename Geometry2DType>
double ElectricalFem2DSolver<Geometry2DType>::computeCuda(unsigned loops,
                                                          const BoundaryConditionsWithMesh<RectangularMesh<2>::Boundary, double>& bvoltage) {
    if (!cuda) setupCuda();
    try {
        const plaskfem::Embedding2D& emb = cudaEmbedding;
        const size_t efull = emb.elements();

        // what loadConductivities and saveHeatDensities read from the geometry, flattened
        std::vector<uint32_t> elem_junc(efull, 0);
        std::vector<uint8_t> elem_role(efull, 0), noheat(efull, 0);
        std::vector<double> Te(efull, 300.);
        auto temperature = inTemperature(this->maskedMesh->getElementMesh());
        for (auto el : this->maskedMesh->elements()) {
            const size_t e = emb.elem(el.getIndex0(), el.getIndex1());
            auto mid = el.getMidpoint();
            auto roles = this->geometry->getRolesAt(mid);
            elem_junc[e] = uint32_t(isActive(mid));
            elem_role[e] = roles.find("p-contact") != roles.end() ? 1 : roles.find("n-contact") != roles.end() ? 2 : 0;
            noheat[e] = this->geometry->getMaterial(mid)->kind() == Material::EMPTY || roles.find("noheat") != roles.end();
            Te[e] = temperature[el.getIndex()];
        }
        cuda->set_elem_temperature(Te.data());
        cuda->set_noheat(noheat);

        // junctions: a 2-D Active is a pfem_junction one element deep along the dummy axis (back = 0, front = 1, ld = 1), so that the
        // library's table index offset + ld * tran + lon is the solver's own act.offset + index0 (loadConductivities) and
        // junction_conductivity travels as it is; beta(T), js(T) at the temperature of the mid-row element of the column (setMatrix)
        std::vector<pfem_junction> junctions(active.size());
        const size_t ncol = junction_conductivity.size();
        std::vector<double> beta_col(ncol, 1.), js_col(ncol, 1.);
        for (size_t n = 0; n != active.size(); ++n) {
            const Active& act = active[n];
            junctions[n] = pfem_junction{act.bottom, act.top, act.left, act.right, 0, 1, 1, act.offset, act.height};
            const size_t mid = (act.bottom + act.top) / 2;
            for (size_t t = act.left; t != act.right; ++t) {
                const size_t tidx = this->maskedMesh->element(t, mid).getIndex();
                const double T = tidx != RectangularMaskedMesh2D::Element::UNKNOWN_ELEMENT_INDEX ? temperature[tidx] : 300.;
                const size_t col = act.offset + t;
                if (!shockleyParameters(n, T, beta_col[col], js_col[col]))
                    throw BadInput(this->getId(), "algorithm 'cuda' needs a Shockley junction (beta, js); this solver has a custom junction model");
            }
        }
        std::vector<double> jcond(2 * ncol);
        for (size_t i = 0; i != ncol; ++i) { jcond[2 * i] = junction_conductivity[i].c00; jcond[2 * i + 1] = junction_conductivity[i].c11; }
        cuda->set_junctions(junctions, elem_junc, elem_role, pcond, ncond, jcond, beta_col, js_col, convergence == CONVERGENCE_STABLE);

        plaskfem::Dirichlet bc;                                   // application order of matrix.hpp:111-118, on both node planes
        for (auto cond : bvoltage) for (auto r : cond.place) emb.add_dirichlet(bc, cudaNode[r], cond.value);
        cuda->set_dirichlet(bc);
        cuda->set_source(nullptr);                                // zero load vector

        potentials = potentials.claim();
        std::vector<double> plane0(emb.plane(), 0.);
        for (size_t i = 0; i != potentials.size(); ++i) plane0[cudaNode[i]] = potentials[i];
        std::vector<double> field = emb.lift(plane0.data());
        cuda->set_field(field.data());

        plaskfem::IterParams ip{this->iter_params.maxit, this->iter_params.maxerr,
                                plaskfem::IterParams::NoConvergenceBehavior(int(this->iter_params.no_convergence_behavior))};
        ip.preconditioner = this->iter_params.preconditioner == IterativeMatrixParams::PRECOND_JAC ? plaskfem::IterParams::PRECOND_JAC :
                            this->iter_params.preconditioner == IterativeMatrixParams::PRECOND_LJAC ? plaskfem::IterParams::PRECOND_LJAC :
                                                                                                      plaskfem::IterParams::PRECOND_MLJ;
        auto result = cuda->solve(false, ip, maxerr, int(loops),
                                  [this](int level, const std::string& msg) { this->writelog(LogLevel(level), msg); });
        this->iter_params.converged = ip.converged; this->iter_params.iters = ip.iters; this->iter_params.err = ip.err;

        cuda->get_field(field.data());
        for (size_t i = 0; i != potentials.size(); ++i) potentials[i] = field[cudaNode[i]];
        std::vector<double> cur3(3 * efull), cur2(2 * efull), cnd(2 * efull);
        cuda->get_elem(PFEM_ELEM_CURRENT, cur3.data());
        emb.elem_vec2(cur3.data(), cur2.data());                  // the longitudinal component is zero
        cuda->get_elem(PFEM_ELEM_COND, cnd.data());
        for (auto el : this->maskedMesh->elements()) {
            const size_t e = emb.elem(el.getIndex0(), el.getIndex1()), i = el.getIndex();
            currents[i] = vec(cur2[2 * e], cur2[2 * e + 1]);
            conds[i] = Tensor2<double>(cnd[2 * e], cnd[2 * e + 1]);
        }
        cuda->get_junction_cond(jcond.data());                    // saveConductivities happened on the device
        for (size_t i = 0; i != ncol; ++i) junction_conductivity[i] = Tensor2<double>(jcond[2 * i], jcond[2 * i + 1]);
        heats.reset();
        maxcur = vec(result.maxcur[1], result.maxcur[2]);
        loopno = result.loopno;
        toterr = result.toterr;
    } catch (const plaskfem::BadInput& err) {
        throw BadInput(this->getId(), "{}", err.what());
    } catch (const std::runtime_error& err) {
        throw ComputationError(this->getId(), "{}", err.what());
    }

    outVoltage.fireChanged();
    outCurrentDensity.fireChanged();
    outHeat.fireChanged();

    return toterr;
}

template <typename Geometry2DType> void ElectricalFem2DSolver<Geometry2DType>::saveHeatDensities() {""")

# ---------------------------------------------------------------- thermal.dynamic Dynamic3D
F = "solvers/thermal/dynamic/femT3d.hpp"
edit(F, """namespace plask { namespace thermal { namespace dynamic {

/**
 * Solver performing calculations in 2D Cartesian or Cylindrical space using finite element method
 */
struct PLASK_SOLVER_API DynamicThermalFem3DSolver:""", """namespace plaskfem { class Context; }   // plaskfem_cuda.hpp: host adapter of libplaskfem_cuda.so (algorithm 'cuda')

namespace plask { namespace thermal { namespace dynamic {

/**
 * Solver performing calculations in 2D Cartesian or Cylindrical space using finite element method
 */
struct PLASK_SOLVER_API DynamicThermalFem3DSolver:""")
edit(F, """    DataVector<Vec<3,double>> fluxes;      ///< Computed (only when needed) heat fluxes on our own mesh
""", """    DataVector<Vec<3,double>> fluxes;      ///< Computed (only when needed) heat fluxes on our own mesh

    std::shared_ptr<plaskfem::Context> cuda;   ///< Device context, exists only for algorithm 'cuda'

    /// Create the device context: mesh, (material, layer thickness) ids, thermk(T) and cp(T)*dens(T) tables
    void setupCuda();

    /// The time loop of compute() on the device (pfem_solve_dynamic)
    double computeCuda(double time, const BoundaryConditionsWithMesh<RectangularMesh<3>::Boundary,double>& btemperature);
""")

F = "solvers/thermal/dynamic/femT3d.cpp"
edit(F, """#include "femT3d.hpp"
""", """#include "femT3d.hpp"

#include <plaskfem_cuda.hpp>
""")
edit(F, """            if (idx != RectangularMaskedMesh3D::Element::UNKNOWN_ELEMENT_INDEX)
                thickness[idx] = h;
        }
    }
}


void DynamicThermalFem3DSolver::onInvalidate() {
    temperatures.reset();
    fluxes.reset();
    thickness.reset();
}
""", """            if (idx != RectangularMaskedMesh3D::Element::UNKNOWN_ELEMENT_INDEX)
                thickness[idx] = h;
        }
    }

    if (algorithm == ALGORITHM_CUDA) setupCuda();
}


void DynamicThermalFem3DSolver::setupCuda() {
    try {
        cuda.reset(new plaskfem::Context(0, this->getId()));
        if (iter_params.preconditioner != IterativeMatrixParams::PRECOND_JAC && this->mesh->axis[2]->size() <= 512)
            cuda->set_layout(PFEM_LAYOUT_VERTICAL_MINOR);
        plaskfem::Mesh fm;
        for (int a = 0; a < 3; ++a) {
            fm.axis[a].reserve(this->mesh->axis[a]->size());
            for (size_t i = 0; i != this->mesh->axis[a]->size(); ++i) fm.axis[a].push_back(this->mesh->axis[a]->at(i));
        }
        fm.order = plaskfem::IterationOrder(int(this->mesh->getIterationOrder()));
        cuda->set_mesh(fm);

        // ids per (material, layer thickness) pair as in ThermalFem3DSolver::setupCuda; one more table per id: cp(T) * dens(T) (:176)
        const size_t nfull = this->mesh->getElementsCount();
        std::vector<shared_ptr<Material>> materials(nfull);
        std::vector<const Material*> key(nfull, nullptr);
        std::vector<double> thick(nfull, 0.);
        std::vector<uint8_t> included(nfull, 0);
        for (auto elem: this->maskedMesh->elements()) {
            size_t e = this->mesh->element(elem.getIndex0(), elem.getIndex1(), elem.getIndex2()).getIndex();
            materials[e] = this->geometry->getMaterial(elem.getMidpoint());
            key[e] = materials[e].get();
            thick[e] = thickness[elem.getIndex()];
            included[e] = 1;
        }
        std::vector<size_t> reps;
        std::vector<uint32_t> ids = plaskfem::material_ids(key, thick, &reps);
        plaskfem::Tables tables = plaskfem::sample_tables(reps.size(), [&](uint32_t id, double T) {
            if (!materials[reps[id]]) return std::make_pair(1., 1.);
            auto k = materials[reps[id]]->thermk(T, thick[reps[id]]);
            return std::make_pair(k.c00, k.c11);
        });
        std::vector<double> cpdens(size_t(tables.nmat) * tables.nT, 1.);
        for (uint32_t id = 0; id != tables.nmat; ++id)
            if (materials[reps[id]])
                for (uint32_t i = 0; i != tables.nT; ++i) {
                    double T = tables.T0 + i * tables.dT;
                    cpdens[size_t(id) * tables.nT + i] = materials[reps[id]]->cp(T) * materials[reps[id]]->dens(T);
                }
        if (!this->maskedMesh->full()) ids = plaskfem::MaskedNumbering(fm, included).mark_excluded(ids);
        cuda->set_materials(ids, tables);
        cuda->set_capacity(tables, cpdens);
        cuda->fill_field(inittemp);
    } catch (const plaskfem::NoDevice& err) {
        throw ComputationError(this->getId(), "algorithm 'cuda' has no CPU fallback: {}", err.what());
    } catch (const plaskfem::BadInput& err) {
        throw BadInput(this->getId(), "{}", err.what());
    }
}


double DynamicThermalFem3DSolver::computeCuda(double time,
                   const BoundaryConditionsWithMesh<RectangularMesh<3>::Boundary,double>& btemperature)
{
    if (!cuda) setupCuda();
    try {
        plaskfem::Mesh fm;
        for (int a = 0; a < 3; ++a)
            for (size_t i = 0; i != this->mesh->axis[a]->size(); ++i) fm.axis[a].push_back(this->mesh->axis[a]->at(i));
        fm.order = plaskfem::IterationOrder(int(this->mesh->getIterationOrder()));
        const bool masked = !this->maskedMesh->full();
        std::vector<uint8_t> included;
        if (masked) {
            included.assign(this->mesh->getElementsCount(), 0);
            for (auto elem: this->maskedMesh->elements())
                included[this->mesh->element(elem.getIndex0(), elem.getIndex1(), elem.getIndex2()).getIndex()] = 1;
        }
        plaskfem::MaskedNumbering numbering(fm, included);
        const size_t nfull = this->mesh->size();
        auto full_node = [&](size_t masked_node) { return masked ? numbering.node_to_full(masked_node) : masked_node; };

        plaskfem::Dirichlet bc;
        for (auto cond: btemperature) for (auto r: cond.place) bc.add_node(full_node(r), cond.value);
        cuda->set_dirichlet(bc);

        auto heats = inHeat(this->maskedMesh->getElementMesh());  // :160
        std::vector<double> heat(this->mesh->getElementsCount(), 0.);
        for (auto elem: this->maskedMesh->elements())
            heat[this->mesh->element(elem.getIndex0(), elem.getIndex1(), elem.getIndex2()).getIndex()] = heats[elem.getIndex()];
        cuda->set_source(heat.data());

        temperatures = temperatures.claim();
        std::vector<double> field(nfull, 0.);
        for (size_t i = 0; i != temperatures.size(); ++i) field[full_node(i)] = temperatures[i];
        cuda->set_field(field.data());

        plaskfem::IterParams ip{iter_params.maxit, iter_params.maxerr,
                                plaskfem::IterParams::NoConvergenceBehavior(int(iter_params.no_convergence_behavior))};
        // jac as named, every other choice -> line-Jacobi (the multilevel preconditioner does not take the capacity diagonal yet)
        ip.preconditioner = iter_params.preconditioner == IterativeMatrixParams::PRECOND_JAC ? plaskfem::IterParams::PRECOND_JAC :
                                                                                                plaskfem::IterParams::PRECOND_LJAC;
        // the corrected theta scheme (plaskfem_cuda.h): no std::swap(temperatures, X) (:287), Dirichlet rows keep their values
        auto result = cuda->solve_dynamic(ip, time, timestep, methodparam, lumping, int(rebuildfreq), int(logfreq), elapstime,
                                          [this](int level, const std::string& msg) { this->writelog(LogLevel(level), msg); });
        iter_params.converged = ip.converged; iter_params.iters = ip.iters; iter_params.err = ip.err;

        cuda->get_field(field.data());
        for (size_t i = 0; i != temperatures.size(); ++i) temperatures[i] = field[full_node(i)];
        maxT = result.maxT;
    } catch (const plaskfem::BadInput& err) {
        throw BadInput(this->getId(), "{}", err.what());
    } catch (const std::runtime_error& err) {
        throw ComputationError(this->getId(), "{}", err.what());
    }

    outTemperature.fireChanged();
    outHeatFlux.fireChanged();

    return 0.;
}


void DynamicThermalFem3DSolver::onInvalidate() {
    temperatures.reset();
    fluxes.reset();
    thickness.reset();
    cuda.reset();
}
""")
edit(F, """    size_t size = this->maskedMesh->size();

    std::unique_ptr<FemMatrix> pA(this->getMatrix());
    FemMatrix& A = *pA.get();
    std::unique_ptr<FemMatrix> pB(this->getMatrix());
    FemMatrix& B = *pB.get();
""", """    if (algorithm == ALGORITHM_CUDA) return computeCuda(time, btemperature);

    size_t size = this->maskedMesh->size();

    std::unique_ptr<FemMatrix> pA(this->getMatrix());
    FemMatrix& A = *pA.get();
    std::unique_ptr<FemMatrix> pB(this->getMatrix());
    FemMatrix& B = *pB.get();
""")

# ---------------------------------------------------------------- electrical.diffusion Diffusion3D
F = "solvers/electrical/diffusion/diffusion3d.hpp"
edit(F, """#include <plask/common/fem.hpp>
#include <plask/plask.hpp>
""", """#include <plask/common/fem.hpp>
#include <plask/plask.hpp>

namespace plaskdiff { class Region; }
""")
edit(F, """    std::map<size_t, ActiveRegion3D> active;  ///< Active regions information
""", """    std::map<size_t, ActiveRegion3D> active;  ///< Active regions information

    /// Algorithm 'cuda': one device context per active region (holds K, F and U on the device)
    std::map<size_t, std::unique_ptr<plaskdiff::Region>> cuda;

    /// The while(true) loop of compute() handed to libplaskfem_cuda (pdiff_compute)
    void computeCuda(unsigned loops, size_t act, ActiveRegion3D& active,
                     const DataVector<double>& A, const DataVector<double>& B, const DataVector<double>& C, const DataVector<double>& D,
                     const DataVector<double>& J, size_t nmodes, const std::vector<DataVector<Tensor2<double>>>& Ps,
                     const std::vector<DataVector<double>>& nrs);
""")

F = "solvers/electrical/diffusion/diffusion3d.cpp"
edit(F, """#include "diffusion3d.hpp"

#define DEFAULT_MESH_SPACING 0.01  // µm
""", """#include "diffusion3d.hpp"

#include <plaskdiff_cuda.hpp>

#define DEFAULT_MESH_SPACING 0.01  // µm
""")
edit(F, """void Diffusion3DSolver::onInvalidate() { active.clear(); }
""", """void Diffusion3DSolver::onInvalidate() {
    active.clear();
    cuda.clear();
}

void Diffusion3DSolver::computeCuda(unsigned loops, size_t act, ActiveRegion3D& active,
                                    const DataVector<double>& A, const DataVector<double>& B, const DataVector<double>& C,
                                    const DataVector<double>& D, const DataVector<double>& J, size_t nmodes,
                                    const std::vector<DataVector<Tensor2<double>>>& Ps, const std::vector<DataVector<double>>& nrs) {
    const size_t nn = active.mesh2->size(), ne = active.emesh2->size();
    try {
        auto& region = cuda[act];
        if (!region) {
            const auto& lateral = *active.mesh2->lateralMesh;
            const auto& full = lateral.fullMesh;
            std::vector<double> ax0(full.axis[0]->size()), ax1(full.axis[1]->size());
            for (size_t i = 0; i != ax0.size(); ++i) ax0[i] = full.axis[0]->at(i);
            for (size_t i = 0; i != ax1.size(); ++i) ax1[i] = full.axis[1]->at(i);
            const int order = full.getIterationOrder() == RectangularMesh2D::ORDER_01 ? PDIFF_ORDER_01 : PDIFF_ORDER_10;
            region.reset(new plaskdiff::Region(this->getId(), ax0, ax1, order, [&](size_t i0, size_t i1) {
                return lateral.getElementIndexFromLowIndexes(i0, i1) != RectangularMaskedMesh2D::NOT_INCLUDED;
            }));
            region->set_U(active.U.data());
        }
        region->set_parameters(A.data(), B.data(), C.data(), D.data());
        region->set_current(J.data());
        std::vector<DataVector<double>> Pm(nmodes), Gm(nmodes), dGm(nmodes);
        std::vector<const double*> Pp(nmodes), Gp(nmodes), dGp(nmodes);
        std::fill(active.modesP.begin(), active.modesP.end(), 0.);
        for (size_t n = 0; n != nmodes; ++n) {
            double wavelength = real(inWavelength(n));
            double factor = inv_hc * wavelength;
            auto gain = inGain(active.emesh2, wavelength, InterpolationMethod::INTERPOLATION_SPLINE);
            auto dgdn = inGain(Gain::DGDN, active.emesh2, wavelength, InterpolationMethod::INTERPOLATION_SPLINE);
            Pm[n].reset(2 * nn); Gm[n].reset(2 * ne); dGm[n].reset(2 * ne);
            for (size_t i = 0; i != nn; ++i) { Pm[n][2 * i] = Ps[n][i].c00; Pm[n][2 * i + 1] = Ps[n][i].c11; }
            for (size_t ie = 0; ie != ne; ++ie) {
                ElementParams3D el(active, ie);
                Tensor2<double> g = nrs[n][ie] * gain[ie], dg = nrs[n][ie] * dgdn[ie];
                Tensor2<double> p = integrateBilinear(el.X, el.Y, Ps[n].data() + ie);
                active.modesP[n] += p.c00 * g.c00 + p.c11 * g.c11;
                Gm[n][2 * ie] = factor * g.c00; Gm[n][2 * ie + 1] = factor * g.c11;
                dGm[n][2 * ie] = factor * dg.c00; dGm[n][2 * ie + 1] = factor * dg.c11;
            }
            active.modesP[n] *= 1e-13 * active.QWheight;
            Pp[n] = Pm[n].data(); Gp[n] = Gm[n].data(); dGp[n] = dGm[n].data();
        }
        region->set_modes(Pp, Gp, dGp);
        pdiff_stats stats;
        int rc = region->compute(loops, maxerr, stats, true, int(iter_params.maxit));
        for (int i = 1; i < stats.loops && i < 64; ++i)
            this->writelog(LOG_RESULT, "Loop {:d}({:d}) @ active region {}: error = {:g}%", i, loopno + i, act, stats.err_log[i]);
        loopno += stats.loops;
        if (rc == PFEM_NOT_CONVERGED) {
            if (iter_params.no_convergence_behavior == IterativeMatrixParams::NO_CONVERGENCE_ERROR)
                throw ComputationError(this->getId(), "Iterative solver did not converge in {} iterations", iter_params.maxit);
            this->writelog(LOG_WARNING, "Iterative solver did not converge in {} iterations", iter_params.maxit);
        }
        region->get_U(active.U.data());
    } catch (const plaskdiff::NoDevice& err) {
        throw ComputationError(this->getId(), "algorithm 'cuda' has no CPU fallback: {}", err.what());
    } catch (const plaskdiff::BadInput& err) {
        throw BadInput(this->getId(), "{}", err.what());
    } catch (const plaskdiff::ComputationError& err) {
        throw ComputationError(this->getId(), "{}", err.what());
    }
}
""")
edit(F, """    unsigned loop = 0;

    std::unique_ptr<FemMatrix> K;

    toterr = 0.;
""", """    unsigned loop = 0;

    if (this->algorithm == ALGORITHM_CUDA) {
        // assembly, residual, decision and linear solve of every loop run on the device (include/plaskdiff_cuda.h)
        toterr = 0.;
        computeCuda(loops, act, active, A, B, C, D, J, nmodes, Ps, nrs);
        outCarriersConcentration.fireChanged();
        return toterr;
    }

    std::unique_ptr<FemMatrix> K;

    toterr = 0.;
""")
edit(F, """        case ALGORITHM_ITERATIVE: K.reset(new SparseFreeMatrix(this, N, 78 * ne)); break;
    }
""", """        case ALGORITHM_ITERATIVE: K.reset(new SparseFreeMatrix(this, N, 78 * ne)); break;
        case ALGORITHM_CUDA: break;  // handled above
    }
""")

# ---------------------------------------------------------------- build
for F, target in (("solvers/thermal/static/CMakeLists.txt", "thermal"), ("solvers/electrical/shockley/CMakeLists.txt", "electrical"),
                  ("solvers/thermal/dynamic/CMakeLists.txt", "dynamic"), ("solvers/electrical/diffusion/CMakeLists.txt", "diffusion")):
    edit(F, """# Build everything the default way.
# Call this macro unless you really know what you are doing!
make_default()""", """# Algorithm 'cuda': the header-only host adapter plaskfem_cuda.hpp and libplaskfem_cuda.so (a plain C ABI; the solver does not
# need nvcc).  Point PLASKFEM_CUDA_ROOT at the directory that holds include/ and libplaskfem_cuda.so.
set(PLASKFEM_CUDA_ROOT "" CACHE PATH "Root of the plaskfem_cuda library (include/plaskfem_cuda.h, libplaskfem_cuda.so)")
find_path(PLASKFEM_CUDA_INCLUDE_DIR plaskfem_cuda.hpp HINTS ${PLASKFEM_CUDA_ROOT}/include)
find_library(PLASKFEM_CUDA_LIBRARY plaskfem_cuda HINTS ${PLASKFEM_CUDA_ROOT} ${PLASKFEM_CUDA_ROOT}/lib ${PLASKFEM_CUDA_ROOT}/plask_b200)
if(NOT PLASKFEM_CUDA_INCLUDE_DIR OR NOT PLASKFEM_CUDA_LIBRARY)
    message(FATAL_ERROR "plaskfem_cuda not found: set PLASKFEM_CUDA_ROOT")
endif()
include_directories(${PLASKFEM_CUDA_INCLUDE_DIR})
set(SOLVER_LINK_LIBRARIES ${SOLVER_LINK_LIBRARIES} ${PLASKFEM_CUDA_LIBRARY})

# Build everything the default way.
# Call this macro unless you really know what you are doing!
make_default()""")


def main():
    if not os.path.isdir(REF):
        raise SystemExit("needs /root/reference")
    tmp = tempfile.mkdtemp(prefix="plaskpatch_")
    a, b = os.path.join(tmp, "a"), os.path.join(tmp, "b")
    for path, edits in EDITS.items():
        for side in (a, b):
            os.makedirs(os.path.dirname(os.path.join(side, path)), exist_ok=True)
            shutil.copyfile(os.path.join(REF, path), os.path.join(side, path))
        text = open(os.path.join(b, path)).read()
        for old, new, count in edits:
            if text.count(old) != count:
                raise SystemExit(f"{path}: expected {count} occurrence(s), found {text.count(old)} of:\n{old[:200]}")
            text = text.replace(old, new)
        open(os.path.join(b, path), "w").write(text)
    r = subprocess.run(["diff", "-ruN", "a", "b"], cwd=tmp, capture_output=True, text=True)
    if r.returncode not in (0, 1):
        raise SystemExit(r.stderr)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # drop the time stamps of the ---/+++ lines: the diff must not change from run to run
    lines = []
    for line in r.stdout.splitlines(keepends=True):
        if line.startswith(("--- a/", "+++ b/")):
            line = line.split("\t")[0] + "\n"
        lines.append(line)
    open(OUT, "w").write("".join(lines))
    shutil.rmtree(tmp)
    print(f"{OUT}: {len(lines)} lines, {len(EDITS)} files")


if __name__ == "__main__":
    main()
