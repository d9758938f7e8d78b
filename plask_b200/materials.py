"""Material models used by the synthetic benchmark/test configurations.

Each function restates the formula of one PLaSK material class (cited) as a numpy expression;
the host samples them into the per-material-id tables that `pfem_set_materials` takes
(in the real plugin the host calls material->thermk(T,h) / ->cond(T) instead).
All return (lateral, vertical) pairs; thermal conductivity in W/(m K), electrical in S/m.
"""
import numpy as np

QE = 1.60217733e-19  # plask/phys/constants.hpp:29


def _iso(v):
    v = np.asarray(v, dtype=np.float64)
    return v, v


# ---- thermal conductivity  thermk(T, thickness) ------------------------------------------
def thermk_GaAs(T):      # materials/semiconductors35/arsenides/GaAs.cpp:225-228
    return _iso(45. * (300. / T) ** 1.28)


def thermk_AlAs(T):      # .../AlAs.cpp:225-228
    return _iso(91. * (300. / T) ** 1.375)


def thermk_AlGaAs(T, Al):  # .../AlGaAs.cpp:251-255 (harmonic mixing with bowing 0.32)
    Ga = 1. - Al
    k = 1. / (Al / thermk_AlAs(T)[0] + Ga / thermk_GaAs(T)[0] + Al * Ga * 0.32)
    return _iso(k)


def thermk_AlOx(T):      # materials/oxides/AlOx.cpp:36-38
    return _iso(np.full_like(np.asarray(T, dtype=np.float64), 0.7))


def thermk_Au(T):        # materials/metals/Au.cpp:70-73
    return _iso(-0.064 * (T - 300.) + 317.1)


def thermk_Cu(T):        # materials/metals/Cu.cpp:67-70
    return _iso(400.8 * (300. / T) ** 0.073)


def thermk_GaN(T, t):    # materials/semiconductors35/nitrides/GaN.cpp:38-44 — depends on the LAYER THICKNESS t [um]
    fun_t = np.tanh(0.001529 * t ** 0.984) ** 0.12
    return _iso(230. * (np.asarray(T, dtype=np.float64) / 300.) ** -1.43 * fun_t)


def thermk_air(T):       # plask/material/air.cpp (thermk fit, ~0.025 W/mK at 300 K)
    return _iso(0.0258 * (np.asarray(T, dtype=np.float64) / 300.) ** 0.8)


# ---- electrical conductivity  cond(T) ------------------------------------------------------
def cond_doped(T, N_cm3, mob_RT_cm2, expo):
    """sigma = qe * N[1/m3] * mob(T)[m2/Vs], mob = mob_RT (300/T)^expo
    (materials/semiconductors35/arsenides/GaAs_Si.cpp:50-53,84-88)."""
    return _iso(QE * N_cm3 * 1e6 * mob_RT_cm2 * (300. / T) ** expo * 1e-4)


def cond_GaAs(T):        # .../GaAs.cpp:234-237 (intrinsic)
    return _iso(1e2 * QE * 8000. * (300. / T) ** (2. / 3.) * 1e16)


def cond_AlOx(T):        # materials/oxides/AlOx.cpp:28-30
    return _iso(np.full_like(np.asarray(T, dtype=np.float64), 1e-7))


def cond_Au(T):          # materials/metals/Au.cpp:60-63
    return _iso(1. / (8.38e-11 * (T - 300.) + 2.279e-8))


def cond_Cu(T):          # materials/metals/Cu.cpp:57-60
    return _iso(1. / (6.81e-11 * (T - 300.) + 1.726e-8))


def cond_air(T):         # plask/material/air.cpp:43-46
    return _iso(np.full_like(np.asarray(T, dtype=np.float64), 0.55e-14))


# ---- volumetric heat capacity cp(T) * dens(T)  [J/(m^3 K)]  (Dynamic3D, femT3d.cpp:176) -------------
def cpdens_GaAs(T):      # .../GaAs.cpp:243 (dens 5317.49 kg/m^3), :249 (cp 327 J/(kg K))
    return np.full_like(np.asarray(T, dtype=np.float64), 0.327e3 * 5.31749e3)


def cpdens_AlGaAs(T, Al):  # .../AlGaAs.cpp:262-264, 271-273: linear in the composition; AlAs.cpp:234 (3730.16), :240 (424)
    Ga = 1. - Al
    return np.full_like(np.asarray(T, dtype=np.float64), (Al * 0.424e3 + Ga * 0.327e3) * (Al * 3.73016e3 + Ga * 5.31749e3))


def cpdens_const(cp, dens):
    """materials without cp / dens in the reference's database (metals, oxides, air): handbook constants, synthetic"""
    return lambda T: np.full_like(np.asarray(T, dtype=np.float64), cp * dens)


def sample_table1(models, T0=250., dT=0.25, nT=1601):
    """Sample a list of callables T -> value on the uniform grid T0 + i*dT (one table, e.g. cp*dens)."""
    T = T0 + dT * np.arange(nT)
    return np.stack([np.asarray(f(T), dtype=np.float64) * np.ones(nT) for f in models])


def sample_tables(models, T0=250., dT=0.25, nT=1601):
    """Sample a list of callables T -> (lat, vert) on the uniform grid T0 + i*dT."""
    T = T0 + dT * np.arange(nT)
    lat = np.empty((len(models), nT))
    vert = np.empty((len(models), nT))
    for m, f in enumerate(models):
        a, b = f(T)
        lat[m], vert[m] = a, b
    return T0, dT, lat, vert
