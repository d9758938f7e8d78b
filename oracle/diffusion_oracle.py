"""oracle/diffusion_oracle.py — CPU restatement of electrical.diffusion.Diffusion3D's FEM.  TEST INFRASTRUCTURE, not the product.

Follows solvers/electrical/diffusion/diffusion3d.{hpp,cpp}:
  * unknowns: 3 per node of the lateral (masked) mesh of the active region — value, d/dy, d/dx (ElementParams3D,
    diffusion3d.hpp:108-141: i00 = 3 n00, i01 = i00 + 1, i10 = i00 + 2, ...);
  * element: the 12 tensor products H_a(x) H_b(y) of the cubic Hermite functions that carry at most one slope
    (the interpolation formula of ConcentrationDataImpl, diffusion3d.cpp:437-457, spells them out);
  * Newton-linearised steady-state diffusion equation  -D lap(u) + A u + B u^2 + C u^3 = J:
        K_ij = D int grad(phi_i).grad(phi_j) + int (A + 2 B u + 3 C u^2) phi_i phi_j
        F_i  = int (J + B u^2 + 2 C u^3) phi_i ,     J bilinear from the nodal values
    (setLocalMatrix, diffusion3d.cpp:196-199; the reference has the integrals in closed form, code-generated into
    diffusion3d-eval.ipp — here they are evaluated by 7x7 Gauss-Legendre quadrature, exact for these polynomials);
  * spatial hole burning (addLocalBurningMatrix, diffusion3d.cpp:201-204, diffusion3d-eval-shb.ipp), per mode:
        K_ij += int s phi_i phi_j ,  F_i += int (Ug s - t) phi_i ,   s = P.dG, t = P.G bilinear from the nodal P,
    with Ug the element-centre value of u (diffusion3d.cpp:296-300);
  * the loop of compute() (diffusion3d.cpp:206-372): assemble, err = 100 |K U - F| / |F|, stop or U = K^-1 F.

Pinning: `tests/test_oracle_diffusion.py` checks the quadrature element matrices against the reference's own generated
expressions (oracle/_ref/libdiffusion_ref.so, built from the .ipp files where they lie; vectors committed under
tests/golden/diffusion_elements.npz) and the whole solver against the two analytic cases of the reference's own test
solvers/electrical/diffusion/tests/diffusion3d.py:86-119 (uniform: rtol 1e-5, gaussian: rtol 0.5e-3).
"""
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libdiffusion_ref.so")

QE = 1.60217733e-19          # phys::qe, plask/phys/constants.hpp [C]
Z0 = 376.73031346177         # phys::Z0 [Ohm]
H_J = 6.62607015e-34
C_LIGHT = 299792458.

# local unknown -> (Hermite index along x, along y); Hermite index: 0 value@lo, 1 slope@lo, 2 value@up, 3 slope@up.
# local node order n00, n01, n10, n11 (first digit = x side), per node: value, d/dy, d/dx
LOCAL = [(2 * ax + (1 if c == 2 else 0), 2 * ay + (1 if c == 1 else 0)) for ax in (0, 1) for ay in (0, 1) for c in (0, 1, 2)]

_GX, _GW = np.polynomial.legendre.leggauss(7)
GX = 0.5 * (_GX + 1.)        # points on [0, 1]
GW = 0.5 * _GW


def hermite(t, h):
    """values and derivatives (d/dx, x = h t) of the four cubic Hermite functions on an interval of length h; shape (4, len(t))"""
    t = np.asarray(t, dtype=float)
    H = np.array([1 - 3 * t**2 + 2 * t**3, h * t * (1 - t)**2, 3 * t**2 - 2 * t**3, h * t**2 * (t - 1)])
    dH = np.array([(-6 * t + 6 * t**2) / h, (1 - t) * (1 - 3 * t), (6 * t - 6 * t**2) / h, t * (3 * t - 2)])
    return H, dH


def basis(X, Y, tx=GX, ty=GX):
    """phi[12, nx, ny], dphi/dx, dphi/dy on the tensor grid of reference points tx x ty of an X x Y element"""
    Hx, dHx = hermite(tx, X)
    Hy, dHy = hermite(ty, Y)
    phi = np.array([np.outer(Hx[a], Hy[b]) for a, b in LOCAL])
    px = np.array([np.outer(dHx[a], Hy[b]) for a, b in LOCAL])
    py = np.array([np.outer(Hx[a], dHy[b]) for a, b in LOCAL])
    return phi, px, py


def bilinear(v4, tx=GX, ty=GX):
    """bilinear interpolation of the 4 nodal values (local node order n00, n01, n10, n11) on the tensor grid"""
    lx = np.array([1 - tx, tx])
    ly = np.array([1 - ty, ty])
    return (v4[0] * np.outer(lx[0], ly[0]) + v4[1] * np.outer(lx[0], ly[1]) + v4[2] * np.outer(lx[1], ly[0]) +
            v4[3] * np.outer(lx[1], ly[1]))


def local_matrix(X, Y, A, B, Cc, D, U12, J4):
    """K[12,12], F[12] of one element (setLocalMatrix)"""
    phi, px, py = basis(X, Y)
    w = np.outer(GW, GW) * X * Y
    u = np.tensordot(U12, phi, axes=1)
    c = A + 2 * B * u + 3 * Cc * u * u
    f = bilinear(J4) + B * u * u + 2 * Cc * u**3
    K = D * (np.einsum("iab,jab,ab->ij", px, px, w) + np.einsum("iab,jab,ab->ij", py, py, w)) + np.einsum("iab,jab,ab->ij", phi, phi, w * c)
    F = np.einsum("iab,ab->i", phi, w * f)
    return K, F


def local_burning(X, Y, G2, dG2, Ug, P42):
    """K[12,12], F[12] contribution of one mode (addLocalBurningMatrix); P42[4][2] nodal (c00, c11)"""
    phi, _, _ = basis(X, Y)
    w = np.outer(GW, GW) * X * Y
    P42 = np.asarray(P42, dtype=float)
    s = bilinear(P42[:, 0] * dG2[0] + P42[:, 1] * dG2[1])
    t = bilinear(P42[:, 0] * G2[0] + P42[:, 1] * G2[1])
    K = np.einsum("iab,jab,ab->ij", phi, phi, w * s)
    F = np.einsum("iab,ab->i", phi, w * (Ug * s - t))
    return K, F


def element_center(X, Y, U12, verbatim=True):
    """ug of diffusion3d.cpp:296-300.  The Hermite interpolant at the element centre is the mean of the corner values plus
    X/16 (difference of the d/dx unknowns) + Y/16 (difference of the d/dy unknowns); the reference multiplies the d/dy unknowns
    (i01, i03, i21, i23) by X and the d/dx ones by Y — identical on square elements.  verbatim=True keeps that."""
    u = U12
    # local order: n00 (0,1,2) n01 (3,4,5) n10 (6,7,8) n11 (9,10,11); value, d/dy, d/dx
    dy = u[1] - u[4] + u[7] - u[10]
    dx = u[2] + u[5] - u[8] - u[11]
    if verbatim:
        return 0.25 * (u[0] + u[3] + u[6] + u[9] + 0.25 * (X * dy + Y * dx))
    return 0.25 * (u[0] + u[3] + u[6] + u[9] + 0.25 * (X * dx + Y * dy))


# ---- the reference's generated expressions (oracle/_ref) -----------------------------------------------------------------

_ref = None


def ref_available():
    return os.path.exists(REF_LIB)


def _load_ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_LIB)
        dp = C.POINTER(C.c_double)
        lib.dref_local_matrix.argtypes = [C.c_double] * 6 + [dp] * 4
        lib.dref_local_matrix.restype = None
        lib.dref_local_burning.argtypes = [C.c_double, C.c_double, dp, dp, C.c_double, dp, dp, dp]
        lib.dref_local_burning.restype = None
        _ref = lib
    return _ref


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def ref_local_matrix(X, Y, A, B, Cc, D, U12, J4):
    U12, J4 = np.ascontiguousarray(U12, dtype=float), np.ascontiguousarray(J4, dtype=float)
    K, F = np.zeros((12, 12)), np.zeros(12)
    _load_ref().dref_local_matrix(X, Y, A, B, Cc, D, _p(U12), _p(J4), _p(K), _p(F))
    return K, F


def ref_local_burning(X, Y, G2, dG2, Ug, P42):
    G2, dG2, P42 = (np.ascontiguousarray(a, dtype=float) for a in (G2, dG2, P42))
    K, F = np.zeros((12, 12)), np.zeros(12)
    _load_ref().dref_local_burning(X, Y, _p(G2), _p(dG2), Ug, _p(P42), _p(K), _p(F))
    return K, F


# ---- the solver ----------------------------------------------------------------------------------------------------------

class Diffusion3DOracle:
    """One active region of Diffusion3DSolver on the FULL lateral grid numbering (node = i0 * n1 + i1, ORDER_01 of
    RectangularMesh2D, rectangular2d.cpp:31-33; element likewise on (n0-1) x (n1-1)).  `active[e]` marks the elements of the masked
    lateral mesh (ActiveRegion3D, diffusion3d.hpp:73-80); nodes that touch none are not unknowns (their U stays 0).

    A, B, C, D: per element (D already in um^2/s, diffusion3d.cpp:229); J: per node, already |js j_z| (diffusion3d.cpp:232-238).
    modes: list of dicts(P=[nn,2], G=[ne,2], dG=[ne,2]) with G, dG already nr * gain * factor (diffusion3d.cpp:289-294).
    """

    def __init__(self, ax0, ax1, active=None):
        self.ax0, self.ax1 = np.asarray(ax0, dtype=float), np.asarray(ax1, dtype=float)
        self.n0, self.n1 = len(self.ax0), len(self.ax1)
        self.nn, self.ne = self.n0 * self.n1, (self.n0 - 1) * (self.n1 - 1)
        self.active = np.ones(self.ne, dtype=bool) if active is None else np.asarray(active, dtype=bool).ravel()
        self.U = np.zeros(3 * self.nn)
        self.loopno = 0
        self.history = []
        act2 = self.active.reshape(self.n0 - 1, self.n1 - 1)
        nodes = np.zeros((self.n0, self.n1), dtype=bool)
        for d0 in (0, 1):
            for d1 in (0, 1):
                nodes[d0:self.n0 - 1 + d0, d1:self.n1 - 1 + d1] |= act2
        self.node_active = nodes.ravel()

    def elem_nodes(self, e):
        i0, i1 = divmod(e, self.n1 - 1)
        n00 = i0 * self.n1 + i1
        return np.array([n00, n00 + 1, n00 + self.n1, n00 + self.n1 + 1]), self.ax0[i0 + 1] - self.ax0[i0], self.ax1[i1 + 1] - self.ax1[i1]

    def assemble_slow(self, A, B, Cc, D, J, modes=(), verbatim=True):
        """element by element through local_matrix / local_burning (the functions pinned against the reference's expressions)"""
        rows, cols, vals = [], [], []
        F = np.zeros(3 * self.nn)
        for e in np.flatnonzero(self.active):
            nd, X, Y = self.elem_nodes(e)
            dof = (3 * nd[:, None] + np.arange(3)[None, :]).ravel()
            Ke, Fe = local_matrix(X, Y, A[e], B[e], Cc[e], D[e], self.U[dof], J[nd])
            for m in modes:
                ug = element_center(X, Y, self.U[dof], verbatim)
                Kb, Fb = local_burning(X, Y, m["G"][e], m["dG"][e], ug, m["P"][nd])
                Ke, Fe = Ke + Kb, Fe + Fb
            rows.append(np.repeat(dof, 12)); cols.append(np.tile(dof, 12)); vals.append(Ke.ravel())
            np.add.at(F, dof, Fe)
        K = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(3 * self.nn,) * 2)
        return K, F

    def assemble(self, A, B, Cc, D, J, modes=(), verbatim=True):
        """the same sums for all elements at once: phi = s * phi_hat with the unit-square functions phi_hat and the scale s = 1, Y, X of
        the unknown (value, d/dy, d/dx); tests/test_oracle_diffusion.py checks it against assemble_slow"""
        es = np.flatnonzero(self.active)
        i0, i1 = np.divmod(es, self.n1 - 1)
        n00 = i0 * self.n1 + i1
        nd = np.stack([n00, n00 + 1, n00 + self.n1, n00 + self.n1 + 1], axis=1)
        X, Y = self.ax0[i0 + 1] - self.ax0[i0], self.ax1[i1 + 1] - self.ax1[i1]
        dof = (3 * nd[:, :, None] + np.arange(3)[None, None, :]).reshape(len(es), 12)
        ph, pxh, pyh = (a.reshape(12, -1) for a in basis(1., 1.))
        w = np.outer(GW, GW).ravel()
        comp = np.tile(np.arange(3), 4)
        sc = np.where(comp[None, :] == 0, 1., np.where(comp[None, :] == 1, Y[:, None], X[:, None]))     # [ne, 12]
        Ue = self.U[dof] * sc
        u = Ue @ ph                                                                                     # [ne, 49]
        Ae, Be, Ce, De = A[es, None], B[es, None], Cc[es, None], D[es]
        c = Ae + 2 * Be * u + 3 * Ce * u * u
        lx = np.array([1 - GX, GX]); ly = lx
        bl = np.array([np.outer(lx[a], ly[b]).ravel() for a in (0, 1) for b in (0, 1)])                 # [4, 49]
        f = J[nd] @ bl + Be * u * u + 2 * Ce * u**3
        for m in modes:
            U12 = self.U[dof].T
            dy = U12[1] - U12[4] + U12[7] - U12[10]
            dx = U12[2] + U12[5] - U12[8] - U12[11]
            ug = 0.25 * (U12[0] + U12[3] + U12[6] + U12[9] + 0.25 * ((X * dy + Y * dx) if verbatim else (X * dx + Y * dy)))
            P = np.asarray(m["P"], dtype=float)[nd]                                                     # [ne, 4, 2]
            G, dG = np.asarray(m["G"], dtype=float)[es], np.asarray(m["dG"], dtype=float)[es]
            sv = (P[:, :, 0] * dG[:, None, 0] + P[:, :, 1] * dG[:, None, 1]) @ bl
            tv = (P[:, :, 0] * G[:, None, 0] + P[:, :, 1] * G[:, None, 1]) @ bl
            c = c + sv
            f = f + ug[:, None] * sv - tv
        area = (X * Y)[:, None]
        Kh = np.einsum("eq,iq,jq->eij", c * w * area, ph, ph)
        Kh += np.einsum("e,iq,jq->eij", De * Y / X, pxh * w, pxh) + np.einsum("e,iq,jq->eij", De * X / Y, pyh * w, pyh)
        Ke = Kh * sc[:, :, None] * sc[:, None, :]
        Fe = ((f * w * area) @ ph.T) * sc
        K = sp.csr_matrix((Ke.ravel(), (np.repeat(dof, 12, axis=1).ravel(), np.tile(dof, (1, 12)).ravel())), shape=(3 * self.nn,) * 2)
        F = np.zeros(3 * self.nn)
        np.add.at(F, dof.ravel(), Fe.ravel())
        return K, F

    def compute(self, A, B, Cc, D, J, loops=0, maxerr=0.05, modes=(), verbatim=True):
        """the while(true) of Diffusion3DSolver::compute (diffusion3d.cpp:283-366); returns the number of loops run"""
        A, B, Cc, D = (np.broadcast_to(np.asarray(a, dtype=float), (self.ne,)) for a in (A, B, Cc, D))
        J = np.broadcast_to(np.asarray(J, dtype=float), (self.nn,))
        dofs = np.flatnonzero(np.repeat(self.node_active, 3))
        loop = 0
        while True:
            K, F = self.assemble(A, B, Cc, D, J, modes, verbatim)
            resid = K @ self.U - F
            err = 100. * np.sqrt(resid @ resid / (F @ F))
            self.history.append(err)
            self.loopno += 1
            loop += 1
            if err < maxerr or (loops != 0 and loop >= loops):
                break
            self.U = self.solve(K, F, dofs)
        self.K, self.F = K, F
        return loop

    def solve(self, K, F, dofs):
        """K->solve(F, U) with the reference's default algorithm: symmetric band Cholesky (DpbMatrix, cholesky_matrix.hpp:90-111:
        dpbtrf + dpbtrs; band 3 nm + 5, diffusion3d.cpp:276).  Full-grid numbering here, so the rows of the nodes outside the masked
        mesh are unit rows."""
        import scipy.linalg as sla
        N = 3 * self.nn
        kd = 3 * self.n1 + 5
        coo = sp.triu(K).tocoo()
        ab = np.zeros((kd + 1, N))
        np.add.at(ab, (kd + coo.row - coo.col, coo.col), coo.data)
        free = np.zeros(N, dtype=bool)
        free[dofs] = True
        ab[kd, ~free] = 1.
        rhs = np.where(free, F, 0.)
        return sla.solveh_banded(ab, rhs, lower=False, check_finite=False)

    def concentration(self, x, y):
        """outCarriersConcentration with spline interpolation (diffusion3d.cpp:420-458); 0 outside the active region.  Points are
        expected inside [ax0[0], ax0[-1]] x [ax1[0], ax1[-1]] (the geometry's mirror wrapping is the caller's)"""
        x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
        i0 = np.clip(np.searchsorted(self.ax0, x, side="right") - 1, 0, self.n0 - 2)
        i1 = np.clip(np.searchsorted(self.ax1, y, side="right") - 1, 0, self.n1 - 2)
        out = np.zeros(x.shape)
        for k in range(x.size):
            e = i0.flat[k] * (self.n1 - 1) + i1.flat[k]
            if not self.active[e]:
                continue
            nd, X, Y = self.elem_nodes(e)
            dof = (3 * nd[:, None] + np.arange(3)[None, :]).ravel()
            phi, _, _ = basis(X, Y, np.array([(x.flat[k] - self.ax0[i0.flat[k]]) / X]), np.array([(y.flat[k] - self.ax1[i1.flat[k]]) / Y]))
            out.flat[k] = self.U[dof] @ phi[:, 0, 0]
        return out


def burned_power(ax0, ax1, active, P, g, qw_height, verbatim=True):
    """modesP of one mode before the mode loop's factors (diffusion3d.cpp:291-302): sum over elements of p.g with
    p = integrateBilinear(X, Y, P + ie).  verbatim=True reproduces diffusion3d.hpp:170-172 and the call as written: the four
    values are P[ie .. ie+3] (consecutive entries starting at the ELEMENT index, both in the numbering of the masked lateral mesh)
    and the area is X*X; verbatim=False takes the four corner nodes and X*Y.  P, g: full-grid numbering (ORDER_01)."""
    n0, n1 = len(ax0), len(ax1)
    P = np.asarray(P, dtype=float)
    active = np.asarray(active, dtype=bool).ravel()
    node_active = np.zeros(n0 * n1, dtype=bool)
    for e in np.flatnonzero(active):
        i0, i1 = divmod(e, n1 - 1)
        n00 = i0 * n1 + i1
        node_active[[n00, n00 + 1, n00 + n1, n00 + n1 + 1]] = True
    Pm = P[node_active]
    tot = 0.
    for ie, e in enumerate(np.flatnonzero(active)):
        i0, i1 = divmod(e, n1 - 1)
        X, Y = ax0[i0 + 1] - ax0[i0], ax1[i1 + 1] - ax1[i1]
        if verbatim:
            p = 0.25 * sum(Pm[k] for k in range(ie, ie + 4) if k < len(Pm)) * X * X
        else:
            n00 = i0 * n1 + i1
            p = 0.25 * (P[n00] + P[n00 + 1] + P[n00 + n1] + P[n00 + n1 + 1]) * X * Y
        tot += p[0] * g[e][0] + p[1] * g[e][1]
    return tot * 1e-13 * qw_height
