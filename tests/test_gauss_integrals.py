"""Pins tests/gauss_integrals.py (test infrastructure: the s/p integrals behind the `pyscf` stand-in) from four
independent sides: the closed-form s-type integrals, p functions as centre derivatives of s functions, grid quadrature
of the AO evaluator's values, and invariances of the H2O/STO-3G Hartree-Fock energy."""
import os
import sys

import numpy as np
import pytest
from scipy.linalg import eigh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gauss_integrals as G  # noqa: E402
import scf_driver  # noqa: E402
from quantum_compute_dft_b200 import molgrid as M  # noqa: E402


def _basis(shells):
    """shells: list of (l, centre, exponent) single-primitive shells with unit coefficients."""
    xyz, ls, off, poff, npr, ex, co, at = [], [], [], [], [], [], [], []
    nao = 0
    for k, (l, c, a) in enumerate(shells):
        xyz.append(c); ls.append(l); off.append(nao); poff.append(k); npr.append(1); ex.append(a); co.append(1.0); at.append(0)
        nao += 1 if l == 0 else 3
    return M.Basis(len(shells), nao, np.array(xyz, dtype=np.float64), np.array(ls, dtype=np.int32),
                   np.array(off, dtype=np.int32), np.array(poff, dtype=np.int32), np.array(npr, dtype=np.int32),
                   np.array(ex), np.array(co), np.array(at, dtype=np.int32))


def _h2o():
    mol = M.load_molecule("H2O")
    return mol, M.sto3g_basis(mol)


def test_s_shells_equal_the_closed_forms():
    z = np.array([0.0, 1.3, 3.1, 4.6])
    mol = M.Molecule("H4", ["H"] * 4, np.stack([0 * z, 0 * z, z], axis=1))
    basis = M.sto3g_basis(mol)
    S0, H0, eri0, en0 = scf_driver.s_integrals(mol, basis)
    S, H, eri, en = G.sp_integrals(mol, basis)
    np.testing.assert_allclose(S, S0, rtol=0, atol=1e-13)
    np.testing.assert_allclose(H, H0, rtol=0, atol=1e-12)
    np.testing.assert_allclose(eri, eri0, rtol=0, atol=1e-12)
    np.testing.assert_allclose(G.sp_kinetic(mol, basis), scf_driver.s_kinetic(mol, basis), rtol=0, atol=1e-13)
    assert abs(en - en0) < 1e-14


def test_p_functions_are_centre_derivatives_of_s_functions():
    """x exp(-a r_A^2) = (1 / 2a) d/dA_x exp(-a r_A^2): every integral with p functions is a centre derivative of the
    integral with s functions in their place.  Central differences on the all-s integrals (an independent code path:
    the Hermite recursions are not exercised at l = 0) against the analytic p-type values."""
    rng = np.random.default_rng(1)
    A, B, C, D = (rng.uniform(-1.0, 1.0, 3) for _ in range(4))
    a, b, c, d = 0.9, 1.7, 0.6, 1.2
    mol = M.Molecule("X", ["O", "H"], np.array([[0.3, -0.2, 0.5], [-0.6, 0.4, -0.1]]))
    h = 1e-3

    def shifted(P, ax, s):
        Q = P.copy(); Q[ax] += s
        return Q

    # analytic: shells p(A,a), p(B,b), s(C,c), s(D,d)
    bp = _basis([(1, A, a), (1, B, b), (0, C, c), (0, D, d)])
    S, T, V = G.one_electron(mol, bp)
    eri = G.eri(bp)

    def ss(Ap, Bp):
        bs = _basis([(0, Ap, a), (0, Bp, b), (0, C, c), (0, D, d)])
        return G.one_electron(mol, bs) + (G.eri(bs),)

    for ax in range(3):
        # one p function: first derivative with respect to A
        plus, minus = ss(shifted(A, ax, h), B), ss(shifted(A, ax, -h), B)
        for mat_p, k in ((S, 0), (T, 1), (V, 2)):
            fd = (plus[k][0, 2] - minus[k][0, 2]) / (2 * h) / (2 * a)       # <p_A | s_C>
            assert abs(mat_p[ax, 6] - fd) < 2e-6 * max(1.0, abs(fd)), (k, ax)
        fd = (plus[3][0, 2, 1, 3] - minus[3][0, 2, 1, 3]) / (2 * h) / (2 * a)  # (p_A s_C | s_B s_D)
        # the s_B in that integral is the l = 0 function at B: build it explicitly
        bmix = _basis([(1, A, a), (0, B, b), (0, C, c), (0, D, d)])
        assert abs(G.eri(bmix)[ax, 4, 3, 5] - fd) < 2e-6
        for ay in range(3):
            # two p functions: mixed second derivative with respect to A and B
            pp, pm = ss(shifted(A, ax, h), shifted(B, ay, h)), ss(shifted(A, ax, h), shifted(B, ay, -h))
            mp, mm = ss(shifted(A, ax, -h), shifted(B, ay, h)), ss(shifted(A, ax, -h), shifted(B, ay, -h))
            for mat_p, k in ((S, 0), (T, 1), (V, 2)):
                fd = (pp[k][0, 1] - pm[k][0, 1] - mp[k][0, 1] + mm[k][0, 1]) / (4 * h * h) / (4 * a * b)   # <p_A | p_B>
                assert abs(mat_p[ax, 3 + ay] - fd) < 5e-6 * max(1.0, abs(fd)), (k, ax, ay)
            fd = (pp[3][0, 1, 2, 3] - pm[3][0, 1, 2, 3] - mp[3][0, 1, 2, 3] + mm[3][0, 1, 2, 3]) / (4 * h * h) / (4 * a * b)
            assert abs(eri[ax, 3 + ay, 6, 7] - fd) < 5e-6, (ax, ay)           # (p_A p_B | s_C s_D)
            fd = (pp[3][0, 2, 1, 3] - pm[3][0, 2, 1, 3] - mp[3][0, 2, 1, 3] + mm[3][0, 2, 1, 3]) / (4 * h * h) / (4 * a * b)
            assert abs(eri[ax, 6, 3 + ay, 7] - fd) < 5e-6, (ax, ay)           # (p_A s_C | p_B s_D)


def test_h2o_overlap_and_kinetic_against_grid_quadrature():
    """S_ij = sum_g w phi_i phi_j and T_ij = 1/2 sum_g w grad phi_i . grad phi_j on the repo's level-3 grid, with the
    AO evaluator's host statement: the analytic integrals and the AO code agree on what the basis functions are."""
    mol, basis = _h2o()
    S, T, V = G.one_electron(mol, basis)
    coords, w, _ = M.make_grid(mol, scale=1.0)
    ao, grad = M.eval_ao_numpy(coords, basis, deriv=1)
    Sq = np.einsum("g,gi,gj->ij", w, ao, ao)
    Tq = 0.5 * np.einsum("g,cgi,cgj->ij", w, grad, grad)
    Vq = np.zeros_like(S)
    for sym, R in zip(mol.symbols, mol.coords):
        r = np.sqrt(((coords - R) ** 2).sum(1))
        Vq -= M.ATOMIC_NUMBER[sym] * np.einsum("g,gi,gj->ij", w / np.maximum(r, 1e-12), ao, ao)
    np.testing.assert_allclose(np.diag(S), 1.0, rtol=0, atol=1e-12)      # sto3g_basis renormalises every function
    assert abs(S[0, 1] - 0.2367) < 5e-5                                  # <1s|2s> of STO-3G oxygen (Szabo & Ostlund)
    # measured quadrature errors on this grid: 1.7e-5 (S), 7.4e-5 (T), 7.4e-5 (V, elements up to 61.7)
    np.testing.assert_allclose(Sq, S, rtol=0, atol=1e-4)
    np.testing.assert_allclose(Tq, T, rtol=0, atol=5e-4)
    np.testing.assert_allclose(Vq, V, rtol=0, atol=5e-4)


def _rhf(S, H, eri, e_nuc, nocc, iters=200):
    _, C = eigh(H, S)
    dm = 2.0 * C[:, :nocc] @ C[:, :nocc].T
    e_old = 0.0
    for _ in range(iters):
        J = np.einsum("ijkl,kl->ij", eri, dm)
        K = np.einsum("ikjl,kl->ij", eri, dm)
        F = H + J - 0.5 * K
        e = 0.5 * np.sum(dm * (H + F)) + e_nuc
        _, C = eigh(F, S)
        dm_new = 2.0 * C[:, :nocc] @ C[:, :nocc].T
        if abs(e - e_old) < 1e-11 and np.abs(dm_new - dm).max() < 1e-9:
            return e
        dm, e_old = 0.5 * dm + 0.5 * dm_new, e
    raise AssertionError("RHF did not converge")


def test_h2o_hartree_fock_energy_and_invariances():
    mol, basis = _h2o()
    S, H, eri, e_nuc = G.sp_integrals(mol, basis)
    n = basis.nao
    assert n == 7
    # 8-fold permutational symmetry of real ERIs
    for perm in ((1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1)):
        np.testing.assert_allclose(eri, eri.transpose(perm), rtol=0, atol=1e-13)
    assert np.all(np.einsum("iiii->i", eri) > 0)
    e = _rhf(S, H, eri, e_nuc, mol.nocc)
    # RHF/STO-3G water is -74.96 Ha near its equilibrium geometry (-74.963 at r = 0.96 A, 104.5 deg); the reference's
    # atom_txt/H2O.xyz has r = 0.99 A, 100 deg
    assert -75.00 < e < -74.93, e          # (measured: -74.96590116)
    # rigid rotation + translation of the molecule: p shells mix, the energy must not move
    rng = np.random.default_rng(2)
    Q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    mol2 = M.Molecule("H2O'", mol.symbols, mol.coords @ Q.T + np.array([0.3, -1.1, 0.7]))
    S2, H2, eri2, e_nuc2 = G.sp_integrals(mol2, M.sto3g_basis(mol2))
    e2 = _rhf(S2, H2, eri2, e_nuc2, mol.nocc)
    assert abs(e2 - e) < 1e-9, (e, e2)
    assert np.abs(eri2 - eri).max() > 1e-3       # (the tensors themselves do differ: the test is not vacuous)
