"""Golden vectors of the Diffusion3D element integrals, produced by the REFERENCE's own generated expressions
(solvers/electrical/diffusion/diffusion3d-eval.ipp, diffusion3d-eval-shb.ipp compiled where they lie into
oracle/_ref/libdiffusion_ref.so by `make -C oracle ref`).  Run in the build container:

    python tests/golden/make_golden_diffusion.py        ->  tests/golden/diffusion_elements.npz

Inputs are random but physically scaled (sizes in um, concentrations ~1e18..1e19 cm^-3, GaAs-like A, B, C, D)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import diffusion_oracle as d  # noqa: E402

rng = np.random.default_rng(20261018)
n = 48
X, Y = rng.uniform(0.01, 0.4, n), rng.uniform(0.01, 0.4, n)
X[:4] = Y[:4]                                   # a few square elements
A, B, C = 3e7 * rng.uniform(.3, 3, n), 1.7e-10 * rng.uniform(.3, 3, n), 6e-27 * rng.uniform(.3, 3, n)
D = 1e9 * rng.uniform(.1, 3, n)
U = np.zeros((n, 12))
U[:, 0::3] = 1e19 * rng.uniform(0.05, 2., (n, 4))
U[:, 1::3] = 1e19 * rng.normal(size=(n, 4))     # slopes per um
U[:, 2::3] = 1e19 * rng.normal(size=(n, 4))
U[4:8] = 0.                                     # the first loop starts from U = 0
J = 1e30 * rng.uniform(0, 1, (n, 4))
G, dG = 1e29 * rng.uniform(-1, 1, (n, 2)), 1e12 * rng.uniform(0, 1, (n, 2))
P = rng.uniform(0, 1, (n, 4, 2))
Ug = np.array([d.element_center(X[i], Y[i], U[i], verbatim=True) for i in range(n)])
K, F, Kb, Fb = np.zeros((n, 12, 12)), np.zeros((n, 12)), np.zeros((n, 12, 12)), np.zeros((n, 12))
for i in range(n):
    K[i], F[i] = d.ref_local_matrix(X[i], Y[i], A[i], B[i], C[i], D[i], U[i], J[i])
    Kb[i], Fb[i] = d.ref_local_burning(X[i], Y[i], G[i], dG[i], Ug[i], P[i])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "diffusion_elements.npz")
np.savez_compressed(out, X=X, Y=Y, A=A, B=B, C=C, D=D, U=U, J=J, G=G, dG=dG, P=P, Ug=Ug, K=K, F=F, Kb=Kb, Fb=Fb)
print("wrote", out, os.path.getsize(out), "bytes")
