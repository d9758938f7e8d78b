"""Design study for the next round (CPU only, not a test and not part of the product): PCG iteration counts on the first-loop
system of config B for Jacobi, the vertical line blocks, and an additive two-level preconditioner = line blocks + Galerkin
coarse correction on piecewise-constant aggregates of c x c lateral columns (vertical resolution kept).  Uses the oracle's
assembled matrix, hence lives under tests/.   python tests/study_twolevel.py   (results quoted in DESIGN.md §10)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from helpers import oracle_thermal
from oracle import oracle as orc
from plask_b200 import configs as cf

def system(n):
    p = cf.config_B(n, order="012")
    o = oracle_thermal(p, algorithm="iterative")
    A14 = orc.Sparse14(o.mesh); B = np.zeros(p.N)
    o.set_matrix(A14, B)
    N = p.N
    data = A14.data.reshape(14, N)
    ic = o.mesh.icords
    rows, cols, vals = [np.arange(N)], [np.arange(N)], [data[0].copy()]
    # sparse14: data[c + rank*i] = A(c + icords[i], c) for i>=1 (see sparse14_at: r>=c, d=r-c, data + c + rank*i)
    for i in range(14):
        d = int(ic[i])
        if d == 0: continue
        c = np.arange(N - d)
        v = data[i][:N - d]
        nz = v != 0
        rows += [c[nz] + d, c[nz]]; cols += [c[nz], c[nz] + d]; vals += [v[nz], v[nz]]
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
    return p, A, B

def pcg(A, b, Minv, tol=1e-8, maxit=20000):
    x = np.zeros_like(b); r = b.copy(); z = Minv(r); p = z.copy(); rz = r @ z; nb = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        q = A @ p; al = rz / (p @ q); x += al * p; r -= al * q
        if np.linalg.norm(r) <= tol * nb: return it
        z = Minv(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return maxit

for n in (32, 48):
    p, A, b = system(n)
    N = p.N; n0, n1, n2 = p.n
    ng = np.broadcast_to(p.node_index_grid(), p.n)      # order 012: index = i0*n1*n2 + i1*n2 + i2 (vertical fastest)
    D = A.diagonal()
    jac = lambda r: r / D
    # vertical line blocks: tridiagonal along i2 (contiguous)
    off = np.array(A.diagonal(1)); off[np.arange(1, N) % n2 == 0 - 0] = off[np.arange(1, N) % n2 == 0]  # keep
    mask = (np.arange(N - 1) + 1) % n2 != 0
    T = sp.diags([D, off * mask, off * mask], [0, 1, -1], format='csc')
    Tlu = spla.splu(T)
    line = lambda r: Tlu.solve(r)
    # coarse space: aggregate cxc lateral columns, keep vertical resolution (semi-coarsening), piecewise constant
    res = {}
    res['jac'] = pcg(A, b, jac)
    res['line'] = pcg(A, b, line)
    for c in (2, 4, 8):
        a0 = np.arange(n0) // c; a1 = np.arange(n1) // c
        m0, m1 = a0.max() + 1, a1.max() + 1
        agg = (a0[:, None, None] * m1 + a1[None, :, None]) * n2 + np.arange(n2)[None, None, :]
        P = sp.csr_matrix((np.ones(N), (ng.ravel(), np.broadcast_to(agg, p.n).ravel())), shape=(N, m0 * m1 * n2))
        Ac = (P.T @ A @ P).tocsc()
        Aclu = spla.splu(Ac)
        add = lambda r: Tlu.solve(r) + P @ Aclu.solve(P.T @ r)                  # additive two-level
        res[f'line+coarse{c}x{c} additive'] = pcg(A, b, add)
    print(n, N, res, flush=True)
