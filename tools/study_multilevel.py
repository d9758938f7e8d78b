"""Design study (CPU only, not part of the product, uses the oracle's assembled matrix): PCG iteration counts of
  jac        point Jacobi
  line       vertical line blocks (NSPCG ljac with blocks = mesh lines)
  ml{c}      ADDITIVE MULTILEVEL line preconditioner:  M^-1 = sum_l P_l T_l^-1 P_l^T,  P_l = piecewise-constant aggregation of
             c^l x c^l lateral columns (vertical resolution kept), T_l = vertical tridiagonal blocks of P_l^T A P_l.  SPD, so plain
             PCG applies; needs no coarse operator, only line factors per level.
on the first-loop system of config B / C.   python tools/study_multilevel.py [n ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from helpers import oracle_thermal, oracle_shockley
from oracle import oracle as orc
from plask_b200 import configs as cf


def system(p):
    if p.kind == "thermal":
        o = oracle_thermal(p, algorithm="iterative")
        A14 = orc.Sparse14(o.mesh); B = np.zeros(p.N)
        o.set_matrix(A14, B)
    else:
        o = oracle_shockley(p, algorithm="iterative")
        A14 = orc.Sparse14(o.mesh); B = np.zeros(p.N)
        o.load_conductivity()
        A14.assemble(o.conds, None, B)
        A14.apply_bc(B, o.bc_nodes, o.bc_values)
    N = p.N
    data = A14.data.reshape(14, N)
    ic = o.mesh.icords
    rows, cols, vals = [np.arange(N)], [np.arange(N)], [data[0].copy()]
    for i in range(14):
        d = int(ic[i])
        if d == 0: continue
        c = np.arange(N - d)
        v = data[i][:N - d]
        nz = v != 0
        rows += [c[nz] + d, c[nz]]; cols += [c[nz], c[nz] + d]; vals += [v[nz], v[nz]]
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
    return A, B


def pcg(A, b, Minv, tol=1e-8, maxit=20000):
    x = np.zeros_like(b); r = b.copy(); z = Minv(r); p = z.copy(); rz = r @ z; nb = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        q = A @ p; al = rz / (p @ q); x += al * p; r -= al * q
        if np.linalg.norm(r) <= tol * nb: return it
        z = Minv(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return maxit


def line_solver(Al, nz):
    """tridiagonal blocks along the fastest index (lines of nz nodes); zero-diagonal rows become identity rows"""
    N = Al.shape[0]
    D = np.array(Al.diagonal())
    off = np.array(Al.diagonal(1)) if N > 1 else np.zeros(0)
    mask = (np.arange(N - 1) + 1) % nz != 0
    dead = D <= 0
    D = np.where(dead, 1., D)
    off = off * mask
    off[dead[:-1] | dead[1:]] = 0.
    T = sp.diags([D, off, off], [0, 1, -1], format='csc')
    lu = spla.splu(T)
    return lambda r: lu.solve(np.where(dead, 0., r))


def levels(p, A, c, fixed, min_cols=1):
    """list of (P_l, line solve of level l); level 0 is the fine mesh.  Vertical axis must be the fastest (order xx2)."""
    n0, n1, n2 = p.n
    N = p.N
    out = [(None, line_solver(A, n2))]
    free = np.ones(N); free[fixed] = 0.
    f = c
    while True:
        a0 = np.arange(n0) // f; a1 = np.arange(n1) // f
        m0, m1 = a0.max() + 1, a1.max() + 1
        agg = (a0[:, None, None] * m1 + a1[None, :, None]) * n2 + np.arange(n2)[None, None, :]
        ng = np.broadcast_to(p.node_index_grid(), p.n)
        P = sp.csr_matrix((free[ng.ravel()], (ng.ravel(), np.broadcast_to(agg, p.n).ravel())), shape=(N, m0 * m1 * n2))
        Al = (P.T @ A @ P).tocsr()
        out.append((P, line_solver(Al, n2)))
        if m0 * m1 <= min_cols:
            break
        f *= c
    return out


def additive(lv):
    def M(r):
        z = lv[0][1](r)
        for P, s in lv[1:]:
            z = z + P @ s(P.T @ r)
        return z
    return M


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [32, 48]
    for n in sizes:
        for name, p in (("B", cf.config_B(n, order="012")), ("C", cf.config_C((n, n, 2 * n), order="012"))):
            A, b = system(p)
            fixed = np.asarray(p.bc_nodes, dtype=np.int64)
            res = {}
            D = A.diagonal()
            if n <= 48:
                res['jac'] = pcg(A, b, lambda r: r / D)
            res['line'] = pcg(A, b, line_solver(A, p.n[2]))
            for c in (2, 4):
                lv = levels(p, A, c, fixed)
                res[f'ml{c} ({len(lv)} levels)'] = pcg(A, b, additive(lv))
                if c == 4:
                    res[f'ml{c} 2 levels only'] = pcg(A, b, additive(lv[:2]))
                    res[f'ml{c} 3 levels only'] = pcg(A, b, additive(lv[:3]))
            print(name, n, p.N, res, flush=True)


if __name__ == "__main__":
    main()
