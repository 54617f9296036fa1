"""gto.Mole as grid.py:43-47,61-66 and dft.py:273-279 use it."""
import os

import numpy as np

from quantum_compute_dft_b200 import molgrid as M


def _parse_atoms(atom):
    """`atom` is an XYZ file path (grid.py:44 passes the path; the first two lines are skipped like dft.py:97-99) or
    the XYZ body as a string (dft.py:274 / grid.py:26)."""
    if isinstance(atom, str) and os.path.exists(atom):
        with open(atom) as f:
            lines = f.readlines()[2:]
    else:
        lines = str(atom).replace(";", "\n").splitlines()
    syms, xyz = [], []
    for ln in lines:
        p = ln.split()
        if len(p) >= 4:
            syms.append(p[0].capitalize())
            xyz.append([float(p[1]), float(p[2]), float(p[3])])
    return syms, np.array(xyz, dtype=np.float64) * M.BOHR_PER_ANGSTROM


class Mole:
    def __init__(self, atom=None, basis="sto-3g", spin=0, verbose=0):
        self.atom, self.basis, self.spin, self.verbose = atom, basis, spin, verbose
        self._built = False

    def build(self):
        if str(self.basis).lower() != "sto-3g":
            raise NotImplementedError("shim: sto-3g only")
        syms, xyz = _parse_atoms(self.atom)
        self._mol = M.Molecule("shim", syms, xyz)
        # PySCF renormalises every contracted function to 1
        self._basis = M.sto3g_basis(self._mol, renormalize=True)
        self.nelec = (self._mol.nelectron // 2 + self.spin, self._mol.nelectron // 2)
        self._ints = None
        self._built = True
        return self

    def _integrals(self):
        if self._ints is None:
            import scf_driver
            if np.all(self._basis.shell_l == 0):      # hydrogen chains: closed forms
                self._ints = scf_driver.s_integrals(self._mol, self._basis)   # S, Hcore, eri, E_nuc
                self._kin = scf_driver.s_kinetic(self._mol, self._basis)
            else:                                     # s and p shells (H2O, ...): McMurchie-Davidson, tests/gauss_integrals.py
                import gauss_integrals
                self._ints = gauss_integrals.sp_integrals(self._mol, self._basis)
                self._kin = gauss_integrals.sp_kinetic(self._mol, self._basis)
        return self._ints

    def nao_nr(self):
        return self._basis.nao

    def intor(self, name):
        S, H, eri, _ = self._integrals()
        if name == "int1e_ovlp":
            return S.copy()
        if name == "int1e_kin":
            return self._kin.copy()
        if name == "int1e_nuc":
            return H - self._kin
        if name == "int2e":
            return eri.copy()
        raise NotImplementedError(name)

    def energy_nuc(self):
        return self._integrals()[3]
