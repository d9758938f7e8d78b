"""The code that patches/plask-algorithm-cuda.diff ADDS to the 2-D solver templates (setupCuda / computeCuda of ThermalFem2DSolver,
DynamicThermalFem2DSolver and ElectricalFem2DSolver) compiled — both geometry instantiations — against a model of the slice of the
PLaSK API it touches (tests/cpp/mock_plask.hpp) and against the real adapter header include/plaskfem_cuda.hpp.  PLaSK itself cannot be
built in this container (no Boost, no LAPACK); this catches what can be caught without it: wrong adapter signatures, members the patch
forgot to declare, template mistakes.  Build container only (/root/reference is absent on the GPU box)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCH = os.path.join(ROOT, "patches", "plask-algorithm-cuda.diff")
REF = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree")

# members of the unpatched classes that the added functions use (therm2d.hpp:28-100, femT2d.hpp:26-90, electr2d.hpp:25-190)
EXISTING = {
    "ThermalFem2DSolver": """
    int loopno; double maxT, toterr, inittemp, maxerr;
    DataVector<double> temperatures, thickness; DataVector<Vec<2, double>> fluxes;
    ReceiverModel<Geometry2DType> inHeat; ProviderModel outTemperature, outHeatFlux;
    void onInitialize(); void onInvalidate(); void saveHeatFluxes();""",
    "DynamicThermalFem2DSolver": """
    double maxT, inittemp, timestep, methodparam, elapstime; bool lumping; size_t rebuildfreq, logfreq;
    DataVector<double> temperatures, thickness; DataVector<Vec<2, double>> fluxes;
    ReceiverModel<Geometry2DType> inHeat; ProviderModel outTemperature, outHeatFlux;
    void onInitialize(); void onInvalidate();""",
    "ElectricalFem2DSolver": """
    struct Active { size_t left, right, bottom, top; ptrdiff_t offset; double height; };
    double pcond, ncond, maxerr, toterr; int loopno; Convergence convergence; Vec<2, double> maxcur;
    DataVector<Tensor2<double>> junction_conductivity, conds; Tensor2<double> default_junction_conductivity;
    DataVector<double> potentials, heats; DataVector<Vec<2, double>> currents; std::vector<Active> active;
    ReceiverModel<Geometry2DType> inTemperature; ProviderModel outVoltage, outCurrentDensity, outHeat;
    size_t isActive(const Vec<2, double>&) const { return 0; }
    virtual Tensor2<double> activeCond(size_t n, double U, double jy, double T) = 0;
    void onInitialize(); void onInvalidate(); void saveHeatDensities();""",
}
EXISTING.update({
    "ThermalFem3DSolver": """
    int loopno; double maxT, toterr, inittemp, maxerr;
    DataVector<double> temperatures, thickness; DataVector<Vec<3, double>> fluxes;
    ReceiverModel<Geometry3D> inHeat; ProviderModel outTemperature, outHeatFlux;
    void onInitialize(); void onInvalidate(); void saveHeatFluxes();""",
    "DynamicThermalFem3DSolver": """
    double maxT, inittemp, timestep, methodparam, elapstime; bool lumping; size_t rebuildfreq, logfreq;
    DataVector<double> temperatures, thickness; DataVector<Vec<3, double>> fluxes;
    ReceiverModel<Geometry3D> inHeat; ProviderModel outTemperature, outHeatFlux;
    void onInitialize(); void onInvalidate();""",
    "ElectricalFem3DSolver": """
    struct Active { size_t bottom, top, left, right, back, front, ld; ptrdiff_t offset; double height; };
    double pcond, ncond, maxerr, toterr; int loopno; Convergence convergence; Vec<3, double> maxcur;
    DataVector<Tensor2<double>> junction_conductivity, conds; Tensor2<double> default_junction_conductivity;
    DataVector<double> potential, heat; DataVector<Vec<3, double>> current; std::vector<Active> active;
    ReceiverModel<Geometry3D> inTemperature; ProviderModel outVoltage, outCurrentDensity, outHeat;
    size_t isActive(const Vec<3, double>&) const { return 0; }
    virtual Tensor2<double> activeCond(size_t n, double U, double jy, double T) = 0;
    void onInitialize(); void onInvalidate(); void saveHeatDensity();""",
})
CASES3D = [
    ("ThermalFem3DSolver", "solvers/thermal/static/therm3d", "namespace plask { namespace thermal { namespace tstatic {", "}}}"),
    ("DynamicThermalFem3DSolver", "solvers/thermal/dynamic/femT3d",
     "namespace plask { namespace thermal { namespace dynamic { using namespace plask::thermal::tstatic;", "}}}"),
    ("ElectricalFem3DSolver", "solvers/electrical/shockley/electr3d", "namespace plask { namespace electrical { namespace shockley {", "}}}"),
]
CASES = [
    ("ThermalFem2DSolver", "solvers/thermal/static/therm2d", "namespace plask { namespace thermal { namespace tstatic {", "}}}"),
    ("DynamicThermalFem2DSolver", "solvers/thermal/dynamic/femT2d",
     "namespace plask { namespace thermal { namespace dynamic { using namespace plask::thermal::tstatic;", "}}}"),
    ("ElectricalFem2DSolver", "solvers/electrical/shockley/electr2d", "namespace plask { namespace electrical { namespace shockley {", "}}}"),
]


def _added_lines(path):
    """the '+' lines of one file's diff"""
    out, take = [], False
    for line in open(PATCH):
        if line.startswith("+++ b/"):
            take = line.strip() == "+++ b/" + path
            continue
        if take and line.startswith("+"):
            out.append(line[1:])
    return out


def _function(text, cls, name):
    """the definition of cls<Geometry2DType>::name from the patched source: from its (template) head to the closing brace in column 0"""
    m = re.search(r"^[^\n]*\b%s(?:<Geometry2DType>)?::%s\(" % (cls, name), text, flags=re.M)
    assert m, (cls, name)
    start = m.start()
    prev = text.rfind("\n", 0, start - 1) + 1
    if text[prev:start].startswith("template"):
        start = prev
    end = text.index("\n}\n", m.end()) + 3
    return text[start:end]


@pytest.mark.parametrize("cls,stem,ns_open,ns_close", CASES)
def test_added_2d_solver_code_compiles_against_the_api_model(tmp_path, cls, stem, ns_open, ns_close):
    for ext in (".hpp", ".cpp"):
        os.makedirs(os.path.dirname(tmp_path / (stem + ext)), exist_ok=True)
        shutil.copyfile(os.path.join(REF, stem + ext), tmp_path / (stem + ext))
    # apply only this solver's part of the patch
    r = subprocess.run(["patch", "-p1", "--batch", "--forward", "-d", str(tmp_path), "-i", PATCH], capture_output=True, text=True)
    assert os.path.exists(tmp_path / (stem + ".cpp"))       # other files of the patch are missing here: their hunks are skipped
    src = open(tmp_path / (stem + ".cpp")).read()
    assert "::setupCuda()" in src and "::computeCuda(" in src, r.stdout[-2000:]
    members = "".join(l for l in _added_lines(stem + ".hpp") if not l.startswith("#include"))
    assert "void setupCuda();" in members and "cudaEmbedding" in members
    code = f"""#include "mock_plask.hpp"
#include "plaskfem_cuda.hpp"
{ns_open}
template <typename Geometry2DType>
struct {cls} : public FemSolverWithMaskedMesh<Geometry2DType, RectangularMesh<2>> {{
{EXISTING[cls]}
{members}
}};
{_function(src, cls, "setupCuda")}
{_function(src, cls, "computeCuda")}
template struct {cls}<Geometry2DCartesian>;
template struct {cls}<Geometry2DCylindrical>;
{ns_close}
"""
    gen = tmp_path / "gen.cpp"
    gen.write_text(code)
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wno-unused-variable", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "tests", "cpp"), str(gen)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[:6000]


@pytest.mark.parametrize("cls,stem,ns_open,ns_close", CASES3D)
def test_added_3d_solver_code_compiles_against_the_api_model(tmp_path, cls, stem, ns_open, ns_close):
    """the headline path: setupCuda / computeCuda (and the cudaMesh helper) the patch adds to therm3d.cpp, electr3d.cpp, femT3d.cpp"""
    for ext in (".hpp", ".cpp"):
        os.makedirs(os.path.dirname(tmp_path / (stem + ext)), exist_ok=True)
        shutil.copyfile(os.path.join(REF, stem + ext), tmp_path / (stem + ext))
    subprocess.run(["patch", "-p1", "--batch", "--forward", "-d", str(tmp_path), "-i", PATCH], capture_output=True, text=True)
    src = open(tmp_path / (stem + ".cpp")).read()
    assert "::setupCuda()" in src and "::computeCuda(" in src
    members = "".join(l for l in _added_lines(stem + ".hpp") if not l.startswith(("#include", "namespace plaskfem")))
    assert "void setupCuda();" in members
    helper = ""
    m = re.search(r"^static plaskfem::Mesh cudaMesh\(", src, flags=re.M)
    if m:
        helper = src[src.rfind("\n", 0, m.start() - 1) + 1:src.index("\n}\n", m.end()) + 3]
    code = f"""#include "mock_plask.hpp"
#include "plaskfem_cuda.hpp"
{ns_open}
struct {cls} : public FemSolverWithMaskedMesh<Geometry3D, RectangularMesh<3>> {{
{EXISTING[cls]}
{members}
}};
{helper}
{_function(src, cls, "setupCuda")}
{_function(src, cls, "computeCuda")}
{ns_close}
"""
    gen = tmp_path / "gen.cpp"
    gen.write_text(code)
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wno-unused-variable", "-Wno-unused-function", "-I",
                        os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "cpp"), str(gen)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[:6000]


def test_added_diffusion3d_code_compiles_against_the_api_model(tmp_path):
    """Diffusion3DSolver::computeCuda (the loop of one active region handed to pdiff_compute) against include/plaskdiff_cuda.hpp"""
    stem = "solvers/electrical/diffusion/diffusion3d"
    for ext in (".hpp", ".cpp"):
        os.makedirs(os.path.dirname(tmp_path / (stem + ext)), exist_ok=True)
        shutil.copyfile(os.path.join(REF, stem + ext), tmp_path / (stem + ext))
    subprocess.run(["patch", "-p1", "--batch", "--forward", "-d", str(tmp_path), "-i", PATCH], capture_output=True, text=True)
    src = open(tmp_path / (stem + ".cpp")).read()
    members = "".join(l for l in _added_lines(stem + ".hpp") if not l.startswith(("#include", "namespace plaskdiff")))
    assert "std::map<size_t, std::unique_ptr<plaskdiff::Region>> cuda;" in members
    code = f"""#include "mock_plask.hpp"
#include "plaskdiff_cuda.hpp"
namespace plask {{ namespace electrical {{ namespace diffusion {{
struct Diffusion3DSolver : public FemSolverWithMaskedMesh<Geometry3D, RectangularMesh<3>> {{
    double maxerr, toterr; unsigned loopno;
    std::map<size_t, ActiveRegion3D> active;
    WavelengthReceiverModel inWavelength; GainReceiverModel inGain; ProviderModel outCarriersConcentration;
    void onInvalidate();
{members}
}};
{_function(src, "Diffusion3DSolver", "computeCuda")}
}}}}}}
"""
    gen = tmp_path / "gen.cpp"
    gen.write_text(code)
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wno-unused-variable", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "tests", "cpp"), str(gen)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[:6000]


def test_beta_solver_override_matches_every_base(tmp_path):
    """BetaSolver<GeometryT> (beta.hpp) derives from ElectricalFem3DSolver for Geometry3D and from ElectricalFem2DSolver<GeometryT>
    otherwise; the `shockleyParameters(...) const override` the patch adds to it only compiles if BOTH bases declare that virtual."""
    beta = "".join(_added_lines("solvers/electrical/shockley/beta.hpp"))
    assert "bool shockleyParameters(size_t n, double PLASK_UNUSED(T), double& beta, double& js) const override" in beta

    def virtual_of(path):
        lines = _added_lines(path)
        i = next(k for k, l in enumerate(lines) if "virtual bool shockleyParameters(" in l)
        return lines[i] + lines[i + 1]
    v3, v2 = virtual_of("solvers/electrical/shockley/electr3d.hpp"), virtual_of("solvers/electrical/shockley/electr2d.hpp")
    code = f"""#include "mock_plask.hpp"
namespace plask {{
struct ElectricalFem3DSolver {{ virtual ~ElectricalFem3DSolver() {{}}
{v3}
}};
template <typename G> struct ElectricalFem2DSolver {{ virtual ~ElectricalFem2DSolver() {{}}
{v2}
}};
template <typename GeometryT>
struct BetaSolver : public std::conditional<std::is_same<GeometryT, Geometry3D>::value, ElectricalFem3DSolver, ElectricalFem2DSolver<GeometryT>>::type {{
    double getBeta(size_t) const {{ return 11.; }}
    double getJs(size_t) const {{ return 1.; }}
{beta}
}};
template struct BetaSolver<Geometry3D>;
template struct BetaSolver<Geometry2DCartesian>;
template struct BetaSolver<Geometry2DCylindrical>;
}}
"""
    gen = tmp_path / "gen.cpp"
    gen.write_text(code)
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "cpp"), str(gen)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[:4000]
