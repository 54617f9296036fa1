"""TEST INFRASTRUCTURE (not product code).  One- and two-electron integrals over contracted Cartesian s and p
Gaussians (McMurchie-Davidson), for the `pyscf` stand-in of tests/shims.

The reference takes S, T, V_nuc and the ERI from PySCF (grid.py:61-66: `mol.intor('int1e_ovlp' | 'int1e_kin' |
'int1e_nuc' | 'int2e')`), which cannot be installed here.  tests/scf_driver.py carries closed forms for s functions
(hydrogen chains); this module covers what STO-3G needs for the BASELINE configurations' first-row molecules -- s and
p shells -- so that the reference's unmodified dft.py can run `python dft.py LDA|B3LYP H2O` (configs C1 and C3)
against the engine.  Pure numpy, vectorised over the primitives of a shell quartet; H2O (5 shells, 21 primitive
functions) takes about a second.

Conventions are molgrid.Basis's: a shell is sum_k coef_k (x-A)^lx (y-A)^ly (z-A)^lz exp(-a_k |r-A|^2) with the
primitive normalisation already inside coef_k; AO order per p shell is x, y, z.

Pinned by tests/test_gauss_integrals.py: equality with the closed-form s-type integrals, p-type integrals as centre
derivatives of s-type ones (finite differences), S and T against grid quadrature of the AO evaluator's values, and the
symmetries of the ERI.
"""
import math

import numpy as np
from scipy.special import gamma, gammainc

from quantum_compute_dft_b200.molgrid import ATOMIC_NUMBER


def boys(nmax, x):
    """F_n(x) = int_0^1 t^(2n) exp(-x t^2) dt for n = 0..nmax; x any non-negative array.  Returns (nmax+1, *x.shape)."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty((nmax + 1,) + x.shape)
    small = x < 1e-6
    xs = np.where(small, 1.0, x)
    for n in range(nmax + 1):
        big = gammainc(n + 0.5, xs) * gamma(n + 0.5) / (2.0 * xs ** (n + 0.5))
        ser = 1.0 / (2 * n + 1) - x / (2 * n + 3) + x * x / (2.0 * (2 * n + 5))
        out[n] = np.where(small, ser, big)
    return out


def _hermite_e(i, j, t, Q, a, b):
    """Hermite expansion coefficient E_t^{ij} of a 1-D Gaussian product; Q = A - B; arrays broadcast."""
    p = a + b
    if t < 0 or t > i + j:
        return 0.0
    if i == j == t == 0:
        return np.exp(-a * b / p * Q * Q)
    if j == 0:
        return (_hermite_e(i - 1, j, t - 1, Q, a, b) / (2.0 * p) - (b / p) * Q * _hermite_e(i - 1, j, t, Q, a, b)
                + (t + 1) * _hermite_e(i - 1, j, t + 1, Q, a, b))
    return (_hermite_e(i, j - 1, t - 1, Q, a, b) / (2.0 * p) + (a / p) * Q * _hermite_e(i, j - 1, t, Q, a, b)
            + (t + 1) * _hermite_e(i, j - 1, t + 1, Q, a, b))


def _hermite_r(t, u, v, n, alpha, X, Y, Z, F):
    """Hermite Coulomb integral R^n_{tuv}; F = boys(nmax, alpha R^2) precomputed."""
    if t < 0 or u < 0 or v < 0:
        return 0.0
    if t == u == v == 0:
        return (-2.0 * alpha) ** n * F[n]
    if t > 0:
        return (t - 1) * _hermite_r(t - 2, u, v, n + 1, alpha, X, Y, Z, F) + X * _hermite_r(t - 1, u, v, n + 1, alpha, X, Y, Z, F)
    if u > 0:
        return (u - 1) * _hermite_r(t, u - 2, v, n + 1, alpha, X, Y, Z, F) + Y * _hermite_r(t, u - 1, v, n + 1, alpha, X, Y, Z, F)
    return (v - 1) * _hermite_r(t, u, v - 2, n + 1, alpha, X, Y, Z, F) + Z * _hermite_r(t, u, v - 1, n + 1, alpha, X, Y, Z, F)


class _Shell:
    def __init__(self, basis, s):
        k0, k1 = basis.shell_prim_off[s], basis.shell_prim_off[s] + basis.shell_nprim[s]
        self.exp = np.asarray(basis.prim_exp[k0:k1], dtype=np.float64)
        self.coef = np.asarray(basis.prim_coef[k0:k1], dtype=np.float64)
        self.center = np.asarray(basis.shell_xyz[s], dtype=np.float64)
        self.l = int(basis.shell_l[s])
        self.off = int(basis.shell_ao_off[s])
        self.powers = [(0, 0, 0)] if self.l == 0 else [(1, 0, 0), (0, 1, 0), (0, 0, 1)]


def _pair(sa, sb):
    """Primitive-pair quantities of two shells, shaped (na, nb)."""
    a, b = sa.exp[:, None], sb.exp[None, :]
    p = a + b
    P = (a[..., None] * sa.center + b[..., None] * sb.center) / p[..., None]
    return a, b, p, P, sa.center - sb.center, sa.coef[:, None] * sb.coef[None, :]


def _overlap_1d(i, j, Q, a, b):
    return _hermite_e(i, j, 0, Q, a, b)


def one_electron(mol, basis):
    """S, T, V_nuc as (nao, nao) matrices."""
    shells = [_Shell(basis, s) for s in range(basis.nshell)]
    n = basis.nao
    S, T, V = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
    charges = [(ATOMIC_NUMBER[sym], np.asarray(R, dtype=np.float64)) for sym, R in zip(mol.symbols, mol.coords)]
    for sa in shells:
        for sb in shells:
            a, b, p, P, Q, cc = _pair(sa, sb)
            norm = cc * (np.pi / p) ** 1.5
            for ia, la in enumerate(sa.powers):
                for ib, lb in enumerate(sb.powers):
                    def s1(d, jb):      # 1-D overlap factor in direction d with the ket's power replaced by jb
                        return _overlap_1d(la[d], jb, Q[d], a, b) if jb >= 0 else 0.0
                    sx, sy, sz = s1(0, lb[0]), s1(1, lb[1]), s1(2, lb[2])
                    S[sa.off + ia, sb.off + ib] = np.sum(norm * sx * sy * sz)
                    # -1/2 <a| d^2/dx^2 |b> per direction: b (2 l + 1) S(l) - 2 b^2 S(l + 2) - 1/2 l (l - 1) S(l - 2)
                    kin = 0.0
                    for d in range(3):
                        l = lb[d]
                        kd = b * (2 * l + 1) * s1(d, l) - 2.0 * b * b * s1(d, l + 2) - 0.5 * l * (l - 1) * s1(d, l - 2)
                        others = [sx, sy, sz]
                        others[d] = kd
                        kin = kin + others[0] * others[1] * others[2]
                    T[sa.off + ia, sb.off + ib] = np.sum(norm * kin)
                    v = 0.0
                    L = sum(la) + sum(lb)
                    for Zc, R in charges:
                        X, Y, Zz = P[..., 0] - R[0], P[..., 1] - R[1], P[..., 2] - R[2]
                        F = boys(L, p * (X * X + Y * Y + Zz * Zz))
                        acc = 0.0
                        for t in range(la[0] + lb[0] + 1):
                            et = _hermite_e(la[0], lb[0], t, Q[0], a, b)
                            for u in range(la[1] + lb[1] + 1):
                                eu = _hermite_e(la[1], lb[1], u, Q[1], a, b)
                                for w in range(la[2] + lb[2] + 1):
                                    ew = _hermite_e(la[2], lb[2], w, Q[2], a, b)
                                    acc = acc + et * eu * ew * _hermite_r(t, u, w, 0, p, X, Y, Zz, F)
                        v = v - Zc * acc
                    V[sa.off + ia, sb.off + ib] = np.sum(cc * 2.0 * np.pi / p * v)
    return S, T, V


def _pair_hermite(sa, sb):
    """For every component pair of two shells: list of (t, u, v, coefficient array (na*nb,)) Hermite terms, with the
    contraction coefficients folded in; plus p and P flattened."""
    a, b, p, P, Q, cc = _pair(sa, sb)
    comps = {}
    for ia, la in enumerate(sa.powers):
        for ib, lb in enumerate(sb.powers):
            terms = []
            for t in range(la[0] + lb[0] + 1):
                et = _hermite_e(la[0], lb[0], t, Q[0], a, b)
                for u in range(la[1] + lb[1] + 1):
                    eu = _hermite_e(la[1], lb[1], u, Q[1], a, b)
                    for w in range(la[2] + lb[2] + 1):
                        ew = _hermite_e(la[2], lb[2], w, Q[2], a, b)
                        terms.append((t, u, w, (cc * et * eu * ew).reshape(-1)))
            comps[(ia, ib)] = terms
    return comps, p.reshape(-1), P.reshape(-1, 3)


def eri(basis):
    """(nao, nao, nao, nao) electron-repulsion integrals (ab|cd), chemists' notation."""
    shells = [_Shell(basis, s) for s in range(basis.nshell)]
    n = basis.nao
    out = np.zeros((n, n, n, n))
    ns = len(shells)
    pairs = {(i, j): _pair_hermite(shells[i], shells[j]) for i in range(ns) for j in range(ns)}
    for i in range(ns):
        for j in range(i + 1):
            bra, p, P = pairs[(i, j)]
            for k in range(ns):
                for l in range(k + 1):
                    if (k, l) > (i, j):       # (ab|cd) = (cd|ab): fill from the other triangle below
                        continue
                    ket, q, Qc = pairs[(k, l)]
                    alpha = p[:, None] * q[None, :] / (p[:, None] + q[None, :])
                    D = P[:, None, :] - Qc[None, :, :]
                    X, Y, Z = D[..., 0], D[..., 1], D[..., 2]
                    L = shells[i].l + shells[j].l + shells[k].l + shells[l].l
                    F = boys(L, alpha * (X * X + Y * Y + Z * Z))
                    pref = 2.0 * np.pi ** 2.5 / (p[:, None] * q[None, :] * np.sqrt(p[:, None] + q[None, :]))
                    rcache = {}
                    for (ia, ib), bterms in bra.items():
                        for (ic, id_), kterms in ket.items():
                            acc = 0.0
                            for (t, u, v, eb) in bterms:
                                for (tt, uu, vv, ek) in kterms:
                                    key = (t + tt, u + uu, v + vv)
                                    if key not in rcache:
                                        rcache[key] = pref * _hermite_r(key[0], key[1], key[2], 0, alpha, X, Y, Z, F)
                                    sign = -1.0 if (tt + uu + vv) & 1 else 1.0
                                    acc = acc + sign * np.einsum("p,pq,q->", eb, rcache[key], ek)
                            A, B = shells[i].off + ia, shells[j].off + ib
                            C, Dd = shells[k].off + ic, shells[l].off + id_
                            for (w, x) in ((A, B), (B, A)):
                                for (y, z) in ((C, Dd), (Dd, C)):
                                    out[w, x, y, z] = acc
                                    out[y, z, w, x] = acc
    return out


def nuclear_repulsion(mol):
    e = 0.0
    for i in range(mol.natm):
        for j in range(i):
            e += ATOMIC_NUMBER[mol.symbols[i]] * ATOMIC_NUMBER[mol.symbols[j]] / \
                math.sqrt(((mol.coords[i] - mol.coords[j]) ** 2).sum())
    return e


def sp_integrals(mol, basis):
    """S, Hcore = T + V_nuc, ERI, E_nuc -- the tuple scf_driver.s_integrals returns, for s and p shells."""
    if np.any(basis.shell_l > 1):
        raise ValueError("sp_integrals: s and p shells only")
    S, T, V = one_electron(mol, basis)
    return S, T + V, eri(basis), nuclear_repulsion(mol)


def sp_kinetic(mol, basis):
    return one_electron(mol, basis)[1]
