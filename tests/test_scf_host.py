"""Host SCF driver (tests/scf_driver.py, the mirror of dft.py:183-266): closed-form s-type
integrals pinned against literature values, and the loop itself with the CPU oracle as backend."""
import numpy as np
import pytest

from quantum_compute_dft_b200 import molgrid as M
import scf_driver as scf
from scf_backends import OracleBackend, h_chain


def test_s_integrals_against_literature():
    """H2 / STO-3G (zeta = 1.24) at R = 1.4 Bohr, Szabo & Ostlund section 3.5.2: S12 = 0.6593, (11|11) = 0.7746,
    E_RHF = -1.1167 Ha; hydrogen atom <h> = -0.46658 Ha."""
    mol, basis = h_chain([0.0, 1.4])
    S, H, eri, e_nuc = scf.s_integrals(mol, basis)
    assert abs(S[0, 1] - 0.6593) < 1e-4 and abs(S[0, 0] - 1.0) < 1e-7
    assert abs(eri[0, 0, 0, 0] - 0.7746) < 1e-4
    assert abs(eri[0, 0, 1, 1] - 0.5697) < 1e-4 and abs(eri[1, 0, 0, 0] - 0.4441) < 1e-4 and abs(eri[1, 0, 1, 0] - 0.2970) < 1e-4
    assert abs(e_nuc - 1.0 / 1.4) < 1e-14
    assert abs(scf.hartree_fock(S, H, eri, e_nuc, 1) - (-1.1167)) < 5e-5
    mol1, basis1 = h_chain([0.0])
    S1, H1, _, _ = scf.s_integrals(mol1, basis1)
    assert abs(H1[0, 0] / S1[0, 0] - (-0.46658)) < 1e-5
    # the 8-fold permutational symmetry of real ERIs
    np.testing.assert_allclose(eri, eri.transpose(1, 0, 2, 3), atol=1e-14)
    np.testing.assert_allclose(eri, eri.transpose(2, 3, 0, 1), atol=1e-14)


@pytest.mark.parametrize("functional", ["LDA", "GGA", "B3LYP"])
def test_scf_loop_with_oracle_backend(oracle, functional):
    """An asymmetric H4 chain (4 AOs, 2 occupied orbitals: the density really iterates) converges, the
    energy is variationally sensible (below the sum of atoms' LDA-ish energies is not required -- just bound
    and reproducible) and the run is deterministic."""
    mol, basis = h_chain([0.0, 1.3, 3.1, 4.6])
    S, H, eri, e_nuc = scf.s_integrals(mol, basis)
    coords, weights, _ = M.make_grid(mol, scale=0.3)
    be = OracleBackend(oracle, functional, basis, coords, weights, eri)
    e1, dm1, n1, ok1 = scf.run_scf(S, H, e_nuc, mol.nocc, be, functional)
    e2, dm2, n2, ok2 = scf.run_scf(S, H, e_nuc, mol.nocc, be, functional)
    assert ok1 and ok2 and n1 == n2 and n1 > 3
    assert e1 == e2
    assert -2.6 < e1 < -1.6          # four hydrogen atoms, ~ -0.45..-0.55 Ha each
    assert abs(np.trace(dm1 @ S) - mol.nelectron) < 1e-10
